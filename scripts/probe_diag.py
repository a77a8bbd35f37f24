"""Per-step backtracking diagnostics of a workload: which gates needed a fallback probe pass and why
(done[4+g]: 1 = a lower-bound proof was not conclusive, 2 = the window was exhausted; done[8+g] = exponent)."""
import sys, torch
sys.path.insert(0, '.')
from bench import WORKLOADS, make_data, bench_params
from admm_lstm_b200 import _lib
from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
n_gpu, T, D, H, O, pname, cpu_n, cls = WORKLOADS[name]
x, y, w = make_data(N, T, D, H, O, 1, cls)
model = LSTM(D, H, O)
with torch.no_grad():
    for k, v in w.items(): getattr(model, k).copy_(torch.from_numpy(v))
opt = ADMMBasedOptimizer(model, (torch.from_numpy(x), torch.from_numpy(y)), bench_params(pname, N, H), verbose=False)
orig = opt._ADMMBasedOptimizer__update_weights
log = []
def hooked(src, st):
    orig(src, st)
    log.append((src, opt._done.cpu().tolist(), [opt._hint[src] if opt._hint else None]))
opt._ADMMBasedOptimizer__update_weights = hooked
for s in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); opt.step(); e1.record(); torch.cuda.synchronize()
    th = opt.theta_trace()
    print(f"step {s}: {e0.elapsed_time(e1):7.1f} ms  exits x {[int(th['x2'+g]).bit_length() for g in 'ifgo']} h {[int(th['h2'+g]).bit_length() for g in 'ifgo']}")
    for src, d, hint in log:
        print(f"    src {'xh'[src]} hint {hint[0]} undecided-after-window {d[4:8]} at exponent {d[8:12]}")
    log.clear()
