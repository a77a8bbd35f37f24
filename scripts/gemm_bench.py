"""Time the tensor-core gate GEMM alone (RAWZ epilogue: Z -> scratch) on a cfg3-shaped slab."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bench import make_data, bench_params
from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer
N, T, D, H, O = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 2, 64, int(sys.argv[2]) if len(sys.argv) > 2 else 1024, 1
x, y, w = make_data(N, T, D, H, O, 1, False)
model = LSTM(D, H, O)
with torch.no_grad():
    for k, v in w.items(): getattr(model, k).copy_(torch.from_numpy(v))
opt = ADMMBasedOptimizer(model, (torch.from_numpy(x), torch.from_numpy(y)), bench_params("GoogleStock", N, H), verbose=False,
                         keep_preactivations=False)
out = torch.empty(4 * H * opt.ldn, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for mode, name in ((1, "tcgen05 3xFP16"), (0, "fp32 CUDA-core")):
    for _ in range(3): opt._call("admm_debug_preact", opt._pp, 2, out.data_ptr(), mode, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if mode else 3
    e0.record()
    for _ in range(reps): opt._call("admm_debug_preact", opt._pp, 2, out.data_ptr(), mode, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 8.0 * H * (D + H) * N
    print(f"{name}: {ms:.3f} ms per launch, {fl / ms / 1e9:.1f} TFLOP/s useful, operand fetch {(N/128)*(H/64)*(128+256)*(D+H)*8/ms/1e9:.2f} TB/s")
# sweep epilogue
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): opt._call("admm_sweep_t", opt._pp, 1, 0, s)
e0.record()
for _ in range(20): opt._call("admm_sweep_t", opt._pp, 1, 0, s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"sweep kernel: {ms:.3f} ms per launch, {8.0 * H * (D + H) * N / ms / 1e9:.1f} TFLOP/s useful")
