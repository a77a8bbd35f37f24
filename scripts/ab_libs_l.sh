#!/bin/bash
# Same-box A/B of library builds on ADMM-LSTM-L (cfg4 shape): scripts/ab_libs_l.sh <tag> <lib1.so> <lib2.so> ...
tag=$1; shift
cp admm_lstm_b200/libadmm_lstm_b200.so /tmp/lib_keep.so
for round in 1 2; do
  for lib in "$@"; do
    name=$(basename $lib .so)
    cp $lib admm_lstm_b200/libadmm_lstm_b200.so
    python bench.py --workload cfg4 --variant admm_l --steps 10 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_${name}_r${round}.json 2> gpurun_out/${tag}_${name}_r${round}.err
  done
done
cp /tmp/lib_keep.so admm_lstm_b200/libadmm_lstm_b200.so
