#!/bin/bash
# run_side_by_side.sh [out dir]   (GPU box; needs oracle/_ref/ten4 and integration/_build/ten4_b200)
# Feeds the same Forth text to the reference build and to the reference VM on libt4k.so, keeps both outputs,
# then integration/diff_outputs.py compares them token by token (numbers within 1e-4 rel / 2e-4 abs — the
# printer shows 4 decimals).  The reference runs twice: a number that differs between its own two runs (random
# init, wall-clock seeds, float atomics in a chaotic trajectory) is held to 10x that spread instead.  Scripts: integration/scripts/*.4th (ours, deterministic) and the reference's own
# examples staged by oracle/ref/build_ref.sh into oracle/_ref/examples/ (git-ignored; never committed).
HERE=$(cd $(dirname $0) && pwd); ROOT=$(cd $HERE/.. && pwd)
OUT=${1:-$ROOT/gpurun_out/side_by_side}
REF=$ROOT/oracle/_ref/ten4; NEW=$HERE/_build/ten4_b200
mkdir -p $OUT/ref $OUT/ref2 $OUT/b200
rc=0
for f in $HERE/scripts/*.4th $ROOT/oracle/_ref/examples/t4_10a.4th $ROOT/oracle/_ref/examples/t4_20a.4th \
         $ROOT/oracle/_ref/examples/t4_30a.4th $ROOT/oracle/_ref/examples/t4_30b.4th $ROOT/oracle/_ref/examples/t4_30c.4th \
         $ROOT/oracle/_ref/examples/t4_30d.4th; do
  [ -f $f ] || continue
  b=$(basename $f .4th)
  timeout 120 $REF < $f > $OUT/ref/$b.out 2> $OUT/ref/$b.err;  echo "ref  $b rc=$?"
  sleep 1.1    # the reference seeds its RNG with time() (1 s resolution): the second run must not share the first one's seed
  timeout 120 $REF < $f > $OUT/ref2/$b.out 2> $OUT/ref2/$b.err   # second run of the reference: its own run-to-run spread
  sleep 1.1
  timeout 120 $NEW < $f > $OUT/b200/$b.out 2> $OUT/b200/$b.err; echo "b200 $b rc=$?"
done
python $HERE/diff_outputs.py $OUT/ref $OUT/b200 $OUT/ref2 | tee $OUT/summary.txt
