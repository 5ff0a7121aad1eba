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
    progression) for ph in 1 2 3 4 5; do timeout 300 python bench.py --config cfg2 --phase $ph --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_cfg2_phase$ph.json 2> $O/${TAG}_bench_cfg2_phase$ph.err; echo "cfg2 phase $ph rc=$?" >> $O/${TAG}_summary.txt; done ;;
    tf32bench) timeout 600 python bench.py --precision tf32 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_cfg3_tf32.json 2> $O/${TAG}_bench_cfg3_tf32.err; echo "tf32 bench rc=$?" >> $O/${TAG}_summary.txt ;;
    policy)   for mv in 16 8192; do
                SARAGAN_TF32_MAX_VOXELS=$mv timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_cfg3_policy$mv.json 2> $O/${TAG}_bench_cfg3_policy$mv.err; echo "policy $mv bench rc=$?" >> $O/${TAG}_summary.txt
                SARAGAN_TF32_MAX_VOXELS=$mv timeout 600 python -m pytest tests/test_step_gpu.py tests/test_fullsize_gpu.py tests/test_trajectory_gpu.py -m gpu -q -s -k "reduced_precision or fullsize or trajectory or layer_local" 2>&1 | grep -a "^\[\|passed\|failed\|median\|worst\|^bf16\|^tf32" > $O/${TAG}_policy${mv}_errors.txt
              done
              timeout 600 python -m pytest tests/test_step_gpu.py tests/test_fullsize_gpu.py tests/test_trajectory_gpu.py -m gpu -q -s -k "reduced_precision or fullsize or trajectory or layer_local" 2>&1 | grep -a "^\[\|passed\|failed\|median\|worst\|^bf16\|^tf32" > $O/${TAG}_policy1024_errors.txt ;;
    ncumem)   ncu --set full --clock-control none -k regex:'k_up2|k_down2|k_pw_expand|k_pw_wgrad|k_pw_reduce|k_adam_multi|k_lrelu|k_lincomb|k_pixelnorm' -c 45 -o $O/${TAG}_ncu_membound python tools/membound_bench.py --reps 1 > $O/${TAG}_ncu_membound.log 2>&1
              ncu -i $O/${TAG}_ncu_membound.ncu-rep --page raw --csv > $O/${TAG}_ncu_membound.csv 2>/dev/null; rm -f $O/${TAG}_ncu_membound.ncu-rep; echo "ncumem rc=$?" >> $O/${TAG}_summary.txt ;;
  esac
done
cat $O/${TAG}_summary.txt
