#!/bin/bash
# Same-box A/B of library builds: scripts/ab_libs.sh <tag> <lib1.so> <lib2.so> ...   (two alternating rounds, cfg3, 10 timed steps)
# Each candidate is copied over admm_lstm_b200/libadmm_lstm_b200.so (the source-hash stamp is untouched, so nothing rebuilds).
tag=$1; shift
cp admm_lstm_b200/libadmm_lstm_b200.so /tmp/lib_keep.so
for round in 1 2; do
  for lib in "$@"; do
    name=$(basename $lib .so)
    cp $lib admm_lstm_b200/libadmm_lstm_b200.so
    python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_${name}_r${round}.json 2> gpurun_out/${tag}_${name}_r${round}.err
  done
done
cp /tmp/lib_keep.so admm_lstm_b200/libadmm_lstm_b200.so
