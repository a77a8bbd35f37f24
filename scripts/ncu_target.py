"""Small fixed program for ncu captures: K steps of the main optimizer at a bench (D, H) with few timesteps.

    python scripts/ncu_target.py [N] [T] [D] [H] [steps] [variant]
Launch order of gate_gemm_tc_persistent per step (T = 16, two chunks of 8, moment probes): x-phase 6 RAWZ (2 chunks x 3
passes, the speculative ones exit at once), h-phase 2 GRAD + 6 RAWZ, then 16 SWEEP; the forward initialisation adds T
FORWARD launches before the first step."""
import sys
import torch
sys.path.insert(0, '.')
from bench import make_data, bench_params
from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer

a = sys.argv[1:]
N = int(a[0]) if len(a) > 0 else 16384
T = int(a[1]) if len(a) > 1 else 16
D = int(a[2]) if len(a) > 2 else 64
H = int(a[3]) if len(a) > 3 else 1024
steps = int(a[4]) if len(a) > 4 else 3
variant = a[5] if len(a) > 5 else "admm"
x, y, w = make_data(N, T, D, H, 1, 1, False)
model = LSTM(D, H, 1)
with torch.no_grad():
    for k, v in w.items():
        getattr(model, k).copy_(torch.from_numpy(v))
opt = ADMMBasedOptimizer(model, (torch.from_numpy(x), torch.from_numpy(y)), bench_params("GoogleStock", N, H), verbose=False,
                         variant=variant, scratch_bytes=8 * 32 * H * ((N + 127) // 128 * 128))
for _ in range(steps):
    opt.step()
torch.cuda.synchronize()
print("ok", opt.metrics())
