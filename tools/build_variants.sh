#!/bin/bash
# Builds alternative libtcb200.so files for tools/variant_sweep.py: tools/build_variants.sh name "-DFLAG=1 ..." [name flags]...
# (development aid; the variants travel to the GPU box with the snapshot, tools/_bin is git-ignored)
# Knobs of bc_device.cuh: TCB_PRUNE (0 = exhaustive 4-colour sweeps, the `base` build every sweep is hash-checked against),
# TCB_PRUNE3, TCB_KSCALE (margin scale), TCB_PRUNE_AUDIT, TCB_NO_F32X2, TCB_PACK_*, TCB_MIN_CTAS.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
      -DTCB_DEV_FEW_KERNELS=1 $flags -o tools/_bin/variants/lib_$name.so tileconv_b200/csrc/tcb200.cu && echo built $name ) &
done
wait
