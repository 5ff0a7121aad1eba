#!/bin/bash
# event-timed tcgen05 conv launches (cold L2) for the low-resolution layer shapes of cfg3
for spec in "4 512 512 2 8 8" "8 512 512 2 8 8" "4 256 256 4 16 16" "4 512 256 4 16 16" "4 128 128 8 32 32" "4 256 128 8 32 32" "4 64 64 16 64 64" "4 32 64 32 128 128" "4 32 32 32 128 128" "4 64 32 32 128 128"; do
  set -- $spec
  timeout 120 python tools/run_conv.py --n $1 --cin $2 --cout $3 --vol $4 $5 $6 --reps 5
done
