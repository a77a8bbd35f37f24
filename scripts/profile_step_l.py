"""ADMM-LSTM-L twin of profile_step.py: one steady-state iteration of a cfg3-shaped problem (T=16) bracketed by
cudaProfilerStart/Stop, the target of `ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_*`
(profiles/r01_ncu_l_launches_*.txt via scripts/launch_traffic.py)."""
import sys, torch
sys.path.insert(0, '.')
from bench import make_data, make_l_weights
from admm_lstm_b200.admm_l import ADMMLOptimizer
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T, D, H = 16, 64, 1024
x, y, _ = make_data(N, T, D, H, 1, 1, False)
w = {k: torch.from_numpy(v) for k, v in make_l_weights(D, H).items()}
opt = ADMMLOptimizer(w, torch.from_numpy(x), torch.from_numpy(y), n_norm=float(N))
for s in range(6):
    if s == 5:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    opt.step()
    torch.cuda.synchronize()
    if s == 5:
        torch.cuda.profiler.stop()
    print(s, {k: float(v) for k, v in opt.thetas.items()})
