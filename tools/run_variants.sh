#!/bin/bash
# runs quickbench (2^20 only) for every experiment variant of the library
mkdir -p gpurun_out
for f in distributed_groth16_b200/variants/lib_*.so; do
  echo "== $f"
  B200ZK_LIB=$PWD/$f timeout 120 python tools/quickbench.py 20 2>&1 | grep -E 'MSM G1|per-call|NTT 2\^22'
done
