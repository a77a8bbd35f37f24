"""One small cfg3-shaped problem (T=16 -> a single time chunk) stepped 4 times: the target of the ncu captures.
Kernel launches of {probe_eval, gate_gemm_tc, atr_tc}: 16 forward + 28 + 28 (steps without a theta hint) + 36 per
steady-state step, so `-s 108 -c 22` captures the weight phase of step 3 and its first two sweep launches."""
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
for s in range(4):
    opt.step()
    torch.cuda.synchronize()
    print(s, opt.theta_trace())
