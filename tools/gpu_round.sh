#!/bin/bash
# One gpurun call of the build -> measure loop: GPU tests, bench, A/B of a second library build, per-kernel profile,
# ncu launch list.  Everything lands in gpurun_out/$TAG_*.   usage: tools/gpu_round.sh TAG [stages...]
TAG=${1:-run}; shift
STAGES=${@:-tests bench ab profile launches}
O=gpurun_out; mkdir -p $O
for st in $STAGES; do
  case $st in
    tests)    timeout 1500 python -m pytest tests -m gpu -q -s > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_summary.txt; tail -5 $O/${TAG}_tests.log >> $O/${TAG}_summary.txt ;;
    bench)    timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" >> $O/${TAG}_summary.txt ;;
    ab)       SARAGAN_B200_LIB=$PWD/saragan_b200/libsaragan_b200_nowd.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_nowd.json 2> $O/${TAG}_bench_nowd.err; echo "ab rc=$?" >> $O/${TAG}_summary.txt ;;
    profile)  timeout 600 python tools/profile_step.py --config cfg3 > $O/${TAG}_insitu_cfg3.txt 2>&1; echo "profile rc=$?" >> $O/${TAG}_summary.txt ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${TAG}_launches_cfg3.csv python tools/one_step.py > $O/${TAG}_ncu.log 2>&1; echo "launches rc=$?" >> $O/${TAG}_summary.txt
              python tools/summarize_launches.py $O/${TAG}_launches_cfg3.csv > $O/${TAG}_launches_cfg3_summary.txt 2>&1 ;;
    membound) timeout 300 python tools/membound_bench.py > $O/${TAG}_membound.txt 2>&1; echo "membound rc=$?" >> $O/${TAG}_summary.txt ;;
    cfgs)     for c in cfg1 cfg2 cfg4 cfg5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err; echo "$c rc=$?" >> $O/${TAG}_summary.txt; done ;;
  esac
done
cat $O/${TAG}_summary.txt
