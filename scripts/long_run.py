"""Stability / speed check at bench scale: S iterations of a workload; every 5 steps the objective, residuals, thetas, the probe
diagnostics (per gate: 0 = decided by the moment pass) and the mean step time of the last 5 steps."""
import sys, math, torch
sys.path.insert(0, '.')
from bench import WORKLOADS, make_data, bench_params
from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
variant = sys.argv[4] if len(sys.argv) > 4 else "admm"
n_gpu, T, D, H, O, pname, cpu_n, cls = WORKLOADS[name]
x, y, w = make_data(N, T, D, H, O, 1000, cls)
model = LSTM(D, H, O)
with torch.no_grad():
    for k, v in w.items(): getattr(model, k).copy_(torch.from_numpy(v))
opt = ADMMBasedOptimizer(model, (torch.from_numpy(x), torch.from_numpy(y)), bench_params(pname, N, H), verbose=False, variant=variant,
                         sharding="presharded")
bad = 0
ev = torch.cuda.Event(enable_timing=True); ev.record()
for s in range(steps):
    opt.step()
    if s % 5 == 4 or s == steps - 1:
        e2 = torch.cuda.Event(enable_timing=True); e2.record(); torch.cuda.synchronize()
        ms = ev.elapsed_time(e2) / 5
        m = opt.metrics()
        th = opt.theta_trace()
        fin = all(math.isfinite(v) for v in m.values()) and all(torch.isfinite(p).all().item() for p in model.parameters())
        bad += not fin
        print(f"step {s + 1}: {ms:7.1f} ms/step  objective {m['objective']:.6g} primal {m['primal_residual']:.5g} dual {m['dual_residual']:.5g} "
              f"loss {m['loss_term']:.4g} finite {fin} diag {opt._done.cpu().tolist()[4:12]} qmax {[round(v, 3) for v in opt._qmax_w.tolist()]} "
              f"plan_x {opt._probe_plans(0)[0]} plan_h {opt._probe_plans(1)[0]} "
              f"exits x {[int(th['x2' + g]).bit_length() for g in 'ifgo']} h {[int(th['h2' + g]).bit_length() for g in 'ifgo']}", flush=True)
        ev = torch.cuda.Event(enable_timing=True); ev.record()
print("OK" if not bad else "NON-FINITE VALUES")
