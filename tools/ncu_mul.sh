#!/bin/bash
# ncu pipe utilisation of the two multiplier schedules (dev aid)
for v in m0 m1; do
  ncu --set full --clock-control none -k regex:k_mul_chain -s 6 -c 1 -o gpurun_out/prof_mul_$v ./tools/microbench_$v > gpurun_out/ncu_mul_$v.log 2>&1
done
