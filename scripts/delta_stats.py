"""How large are the probe perturbations delta = Q 2^-k at the bottom of the probe window?  (feasibility of evaluating
the window candidates by a Taylor expansion around Z0 instead of one activation per candidate)"""
import sys, ctypes as C, torch
sys.path.insert(0, '.')
from bench import WORKLOADS, make_data, bench_params
from admm_lstm_b200 import _lib
from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n_gpu, T, D, H, O, pname, cpu_n, cls = WORKLOADS[name]
x, y, w = make_data(N, T, D, H, O, 1, cls)
model = LSTM(D, H, O)
with torch.no_grad():
    for k, v in w.items(): getattr(model, k).copy_(torch.from_numpy(v))
opt = ADMMBasedOptimizer(model, (torch.from_numpy(x), torch.from_numpy(y)), bench_params(pname, N, H), verbose=False)
orig = opt._call
state = {"seen": set()}
def hooked(name_, *args):
    orig(name_, *args)
    if name_ == "admm_weight_probe" and opt._step_index >= 4:
        src, t0, tc = args[1], args[2], args[3]
        plan = C.cast(args[6], C.POINTER(_lib.ProbePlan)).contents
        key = (opt._step_index, src)
        if key in state["seen"] or t0 != 0 or not plan.proof:
            return
        state["seen"].add(key)
        torch.cuda.synchronize()
        q = opt._scratch[4 * H * tc * opt.ldn: 8 * H * tc * opt.ldn].view(4, H, tc, opt.ldn)[:, :, :, :N]
        for g in range(4):
            k0 = plan.k0[g]
            d = q[g].abs() * 2.0 ** (-k0)
            print(f"step {opt._step_index} src {'xh'[src]} gate {'ifgo'[g]} k0={k0} max|delta|={float(d.max()):.3g} "
                  f"frac>0.1: {float((d > 0.1).float().mean()):.2e} frac>0.05: {float((d > 0.05).float().mean()):.2e} "
                  f"frac>0.02: {float((d > 0.02).float().mean()):.2e} median {float(d.median()):.2e}")
opt._call = hooked
for s in range(6):
    opt.step()
torch.cuda.synchronize()
