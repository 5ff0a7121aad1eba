#!/bin/bash
# ncu --set full captures of the hot kernels alone on the cfg3 top-level shapes (one GPU; run under gpurun, never with a
# multi-rank command): the three convolution kernels, the weight gradient, and the memory-bound kernels.
# Output: gpurun_out/${TAG}_ncu_<name>.csv (raw page) -> tools/ncu_traffic.py -> profiles/.
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
run() {  # name, kernel regex, command...
  local name=$1 re=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$re -c 6 -o $O/${TAG}_ncu_$name "$@" > $O/${TAG}_ncu_$name.log 2>&1
  ncu -i $O/${TAG}_ncu_$name.ncu-rep --page raw --csv > $O/${TAG}_ncu_$name.csv 2>/dev/null
  rm -f $O/${TAG}_ncu_$name.ncu-rep
}
run conv_32_32 'k_conv_tc' python tools/run_conv.py --n 4 --cin 32 --cout 32 --vol 32 128 128 --what fprop --reps 1
run conv_64_32 'k_conv_tc' python tools/run_conv.py --n 4 --cin 64 --cout 32 --vol 32 128 128 --what fprop --reps 1
run conv_32_64 'k_conv_tc' python tools/run_conv.py --n 4 --cin 32 --cout 64 --vol 32 128 128 --what fprop --reps 1
run wgrad_32_64 'k_wgrad_tc' python tools/run_conv.py --n 4 --cin 32 --cout 64 --vol 32 128 128 --what wgrad --reps 1
run wgrad_32_32 'k_wgrad_tc' python tools/run_conv.py --n 4 --cin 32 --cout 32 --vol 32 128 128 --what wgrad --reps 1
run membound 'k_up2|k_down2|k_pw_expand|k_pw_wgrad|k_pw_reduce|k_adam_multi|k_lrelu|k_lincomb|k_pixelnorm' python tools/membound_bench.py --reps 1
ls -la $O | grep ${TAG}_ncu
