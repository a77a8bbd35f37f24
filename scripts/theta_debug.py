import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bench import make_data, WORKLOADS, bench_params
from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer
from admm_lstm_b200.parameters import example_parameter_dictionary as epd
n_gpu, T, D, H, O, pname, cpu_n, cls = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
n_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
x, y, w = make_data(n_gpu, T, D, H, O, 1000, cls)
model = LSTM(D, H, O)
with torch.no_grad():
    for k, v in w.items(): getattr(model, k).copy_(torch.from_numpy(v))
opt = ADMMBasedOptimizer(model, (torch.from_numpy(x), torch.from_numpy(y)), bench_params(pname, n_gpu, H), verbose=False, sharding="presharded")
for s in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.step()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tr = opt.theta_trace(); m = opt.metrics(); dn = opt._done.cpu().tolist()
    print(s + 1, f"{dt*1e3:8.1f} ms hint={opt._hint}", {k: (int(np.log2(v)) if v >= 1 else v) for k, v in tr.items()},
          f"obj={m['objective']:.4g} prim={m['primal_residual']:.4g} diag(h-phase)={dn[4:]}")
