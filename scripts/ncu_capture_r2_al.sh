# ncu --set full of one launch each of the kernels changed at the end of round 2 (scripts/ncu_target.py: N=16384, T=16, D=64,
# H=1024, third step): the fused moment pass (x and h phase), grad_from_z, atr_tc<64>.  Same recipe as ncu_capture_kernels.sh.
set -x
python scripts/ncu_target.py > gpurun_out/plain.log 2>&1 || exit 1
cap() { # name, kernel regex, skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 python scripts/ncu_target.py > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r2_al_$1_raw.csv 2>/dev/null
}
cap moments_x gate_gemm_tc_persistent 82
cap moments_h gate_gemm_tc_persistent 90
cap grad_from_z grad_from_z_kernel 4
cap atr64 atr_tc_kernel 8
ls -la gpurun_out/r2_al_*
