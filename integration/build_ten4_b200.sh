#!/bin/bash
# build_ten4_b200.sh [reference root] [output dir]
# Builds `ten4_b200`: the reference's Forth VM / MMU / model / IO sources, UNMODIFIED and compiled where they lie, on top
# of libt4k.so instead of the reference's own kernels:
#   * src/mu/tensor.cu, src/nn/gradient.cu, src/nn/debug.cu  — compiled with `-include integration/ref_fork_shim.h`
#     (the FORK* launch macros become C-ABI calls)
#   * src/nn/forward.cu, src/nn/backprop.cu                  — replaced by integration/model_shim.cu
#   * src/nn/nmath.cu (NN kernels)                           — dropped (nothing references it any more)
#   * src/t4math.cu                                          — kept only for the out-of-scope LA kernels (inverse / LU / det)
#   * src/mu/mmu.cu, src/ten4.cu                             — compiled with `-include integration/arena_shim.h` (object store sized for the
#     device, device-preferred managed memory); src/mu/tlsf.cpp replaced by integration/arena_shim.cpp (64-bit host-side allocator, §8f row 3)
# Same CMake bypass and `-include iostream` fix-up as oracle/ref/build_ref.sh.  No reference source is copied.
set -e
REF=${1:-/root/reference}; HERE=$(cd $(dirname $0) && pwd); OUT=${2:-$HERE/_build}
ROOT=$(cd $HERE/.. && pwd)
R=$REF/src; O=$OUT/obj
mkdir -p $O
ARCH="-gencode arch=compute_100a,code=sm_100a"
NV="nvcc -std=c++17 -O2 -I$R --device-c --expt-extended-lambda $ARCH -w"
SHIM="-include $HERE/ref_fork_shim.h"
ARENA="-include $HERE/arena_shim.h"
for f in util t4math; do $NV -c $R/$f.cu -o $O/$f.o & done
$NV $ARENA -c $R/ten4.cu -o $O/ten4.o &
$NV $ARENA -c $R/mu/mmu.cu -o $O/mu_mmu.o &
$NV -c $R/mu/dataset.cu -o $O/mu_dataset.o &
$NV -include iostream $SHIM -c $R/mu/tensor.cu -o $O/mu_tensor.o &
$NV $SHIM -c $R/nn/gradient.cu -o $O/nn_gradient.o &
$NV $SHIM -c $R/nn/debug.cu -o $O/nn_debug.o &
$NV -c $HERE/model_shim.cu -o $O/model_shim.o &
CX="g++ -std=c++17 -O2 -I$R -I/usr/local/cuda/include -w"
for f in sys debug; do $CX -c $R/$f.cpp -o $O/$f.cpp.o & done
$CX -c $R/mu/mpool.cpp -o $O/mu_mpool.cpp.o &
$CX -c $HERE/arena_shim.cpp -o $O/arena_shim.cpp.o &
for f in $R/io/aio*.cpp $R/vm/*.cpp $R/ld/*.cpp $R/nn/loss.cpp $R/nn/model.cpp $R/tb/summary.cpp; do
  b=$(echo $f | sed "s#$R/##; s#/#_#g"); $CX -c $f -o $O/$b.o & done
wait
nvcc $ARCH -cudart shared -Xnvlink --suppress-stack-size-warning $O/*.o -o $OUT/ten4_b200 \
     -L$ROOT/tensorforth_b200 -lt4k -Xlinker -rpath -Xlinker '$ORIGIN/../../tensorforth_b200'
rm -rf $O
ls -la $OUT
