"""One small cfg3-shaped problem (T=16 -> a single time chunk): steady-state step bracketed by cudaProfilerStart/Stop,
the target of `ncu --profile-from-start off --set full` (profiles/r01_ncu_full_*.txt)."""
import sys, torch
sys.path.insert(0, '.')
from bench import make_data, bench_params
from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T, D, H, O = 16, 64, 1024, 1
x, y, w = make_data(N, T, D, H, O, 1, False)
model = LSTM(D, H, O)
with torch.no_grad():
    for k, v in w.items(): getattr(model, k).copy_(torch.from_numpy(v))
opt = ADMMBasedOptimizer(model, (torch.from_numpy(x), torch.from_numpy(y)), bench_params("GoogleStock", N, H), verbose=False)
for s in range(6):
    if s == 5:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    opt.step()
    torch.cuda.synchronize()
    if s == 5:
        torch.cuda.profiler.stop()
    print(s, opt.theta_trace())
