"""Small fixed program for compute-sanitizer (memcheck / racecheck): two steps of the main optimizer on a tensor-core shape
(staged GRAD refresh, fused MOMENTS, SWEEP, fp16 A^T R) and on a CUDA-core shape, and two ADMM-LSTM-L iterations.

    compute-sanitizer --tool memcheck python scripts/sanitizer_target.py
"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from helpers import GOOGLE, HAR, synthetic_problem
from gpu_utils import make_opt
from admm_lstm_b200.admm_l import ADMMLOptimizer

for shape, params, variant, tc in (((256, 3, 16, 64, 1), GOOGLE, "admm", True), ((300, 2, 9, 128, 2), HAR, "no_dual_y", True),
                                   ((200, 5, 3, 12, 2), GOOGLE, "admm", False)):
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=5, classification=(o > 1))
    _, opt = make_opt(w, x, y, params, variant, use_tensor_cores=tc)
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    print(shape, variant, "tc" if tc else "simt", opt.metrics())
rng = np.random.default_rng(3)
n, t, d, h = 300, 3, 16, 128
wt = {"Wy": torch.from_numpy((rng.standard_normal((h, 1)) * 0.1).astype(np.float32))}
for g in "fiog":
    wt["W" + g] = torch.from_numpy((rng.standard_normal((d, h)) * 0.1).astype(np.float32))
    wt["U" + g] = torch.from_numpy((rng.standard_normal((h, h)) * 0.1).astype(np.float32))
lopt = ADMMLOptimizer(wt, torch.from_numpy(rng.random((n, t, d), dtype=np.float32)), torch.from_numpy(rng.random((n, 1), dtype=np.float32)),
                      n_norm=float(n))
for _ in range(2):
    lopt.step()
torch.cuda.synchronize()
print("admm_l ok", float(lopt.weights()["Wy"].abs().max()))
