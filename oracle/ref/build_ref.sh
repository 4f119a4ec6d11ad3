#!/bin/bash
# build_ref.sh <reference root> <output dir>
# Compiles the UNMODIFIED reference (chochain/tensorForth) from its sources where they lie,
# for sm_100, into <output dir> (oracle/_ref/, git-ignored).  Recipe = SURVEY.md §8c:
# the reference's CMake is bypassed (GL/GLUT/SDL3 FetchContent need X11 + network);
# src/vu (GL viewer) is skipped as the reference's own Makefile does (Makefile:60);
# one fix-up: `-include iostream` for src/mu/tensor.cu (std::ostream undeclared, tensor.h:157).
# Outputs:
#   ten4     the whole reference program (REPL)           -> bench.py --impl reference, script goldens
#   refkern  oracle/ref/refkern.cu + reference kernel TUs  -> kernel-level goldens at full FP32
set -e
REF=${1:-/root/reference}; OUT=${2:-$(dirname $0)/../_ref}
HERE=$(cd $(dirname $0) && pwd)
R=$REF/src; O=$OUT/obj
mkdir -p $O
ARCH="-gencode arch=compute_100,code=sm_100"
NV="nvcc -std=c++17 -O2 -I$R --device-c --expt-extended-lambda $ARCH -w"
for f in util t4math ten4; do $NV -c $R/$f.cu -o $O/$f.o & done
for f in mmu dataset; do $NV -c $R/mu/$f.cu -o $O/mu_$f.o & done
$NV -include iostream -c $R/mu/tensor.cu -o $O/mu_tensor.o &
for f in nmath forward backprop gradient debug; do $NV -c $R/nn/$f.cu -o $O/nn_$f.o & done
$NV -I$R/nn -c $HERE/refkern.cu -o $O/refkern.o &
CX="g++ -std=c++17 -O2 -I$R -I/usr/local/cuda/include -w"
for f in sys debug; do $CX -c $R/$f.cpp -o $O/$f.cpp.o & done
for f in tlsf mpool; do $CX -c $R/mu/$f.cpp -o $O/mu_$f.cpp.o & done
for f in $R/io/aio*.cpp $R/vm/*.cpp $R/ld/*.cpp $R/nn/loss.cpp $R/nn/model.cpp $R/tb/summary.cpp; do
  b=$(echo $f | sed "s#$R/##; s#/#_#g"); $CX -c $f -o $O/$b.o & done
wait
nvcc $ARCH -Xnvlink --suppress-stack-size-warning $(ls $O/*.o | grep -v refkern.o) -o $OUT/ten4
nvcc $ARCH $O/refkern.o $O/t4math.o $O/nn_nmath.o -o $OUT/refkern
# the reference's example scripts, staged beside the binaries for integration/run_side_by_side.sh
# (oracle/_ref/ is git-ignored: they travel to the GPU box, they never enter this repository)
mkdir -p $OUT/examples && cp $REF/examples/*.4th $OUT/examples/
rm -rf $O
ls -la $OUT
