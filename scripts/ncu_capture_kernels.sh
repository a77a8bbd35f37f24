# ncu --set full of one launch of each main kernel (scripts/ncu_target.py: N=16384, T=16, D=64, H=1024, third step); the
# reports stay on the box, their raw / source pages come back as csv.  Persistent-kernel launch order per step: 2 GRAD:full
# (exit at once), 2 MOMENTS x, 4 RAWZ (exit), 2 GRAD refresh, 2 MOMENTS h, 4 RAWZ (exit), 16 SWEEP; 16 FORWARD before step 1.
set -x
python scripts/ncu_target.py > gpurun_out/plain.log 2>&1 || exit 1
cap() { # name, kernel regex, skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 python scripts/ncu_target.py > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r2_v_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/r2_v_$1_src.csv 2>/dev/null
}
cap moments_x gate_gemm_tc_persistent 82
cap grad_staged gate_gemm_tc_persistent 88
cap moments_h gate_gemm_tc_persistent 90
cap sweep gate_gemm_tc_persistent 100
cap atr256 atr_tc_kernel 10
ls -la gpurun_out/r2_v_*
