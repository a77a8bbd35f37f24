"""oracle/ref_runner.py -- time the UNMODIFIED reference (staged in oracle/_ref/ by oracle/make_ref.py).

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.  Runs in its own process (bench.py spawns it) so that
  * `import admm` resolves to the reference's admm.py (oracle/_ref is first on sys.path; the repo root with its drop-in
    admm.py is NOT on the path: this file lives in oracle/),
  * the device the reference picks at import (`_global.device`, reference _global.py:217: CUDA if available) can be
    controlled from outside through CUDA_VISIBLE_DEVICES, and the BLAS / torch thread count through --threads,
  * its `logs/` directory goes to a scratch CWD.

    python oracle/ref_runner.py --data problem.npz --variant admm|no_dual_y|admm_l --steps K --warmup W [--threads C]

problem.npz: x [N,T,D], y [N,O], the nine weight tensors (x2i..h2o, out) or the ADMM-LSTM-L ones (W*, U*, Wy), and
`params_json` (the rho / beta dictionary) -- written by bench.py with the same generator as the B200 arm.
Prints ONE JSON line: {"step_s": [...], "threads": C, "device": "...", "torch": "..."}; step_s are the K timed steps.
What is timed is exactly what demo.py:354-356 times: `optimizer.step()` (main.py:139-188's loop body for admm_l).
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True)
    ap.add_argument("--variant", default="admm", choices=["admm", "no_dual_y", "admm_l"])
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--dump", default="", help="write the reference's weights, gates and duals after the last step to this .npz "
                    "(tests/test_oracle_live_reference.py; not for admm_l)")
    a = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "admm.py")):
        print(json.dumps({"unavailable": "oracle/_ref is not staged (run python oracle/make_ref.py where /root/reference exists)"}))
        return 0
    if a.threads > 0:                      # before torch / MKL / OpenMP initialise (torchrun exports OMP_NUM_THREADS=1)
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ[k] = str(a.threads)
    data_path = os.path.abspath(a.data)
    dump_path = os.path.abspath(a.dump) if a.dump else ""
    os.chdir(tempfile.mkdtemp(prefix="admm_ref_run_"))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(HERE)]
    sys.path.insert(0, REF)
    import numpy as np
    import torch
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    d = np.load(data_path)
    params = json.loads(str(d["params_json"]))
    import _global                                            # reference module
    dev = _global.device
    x, y = torch.from_numpy(d["x"]).to(dev), torch.from_numpy(d["y"]).to(dev)

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    times = []
    if a.variant == "admm_l":
        # main.py imports matplotlib (TkAgg) and demo.save_model at module level; neither is used by the iteration
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *args, **kw: None
        sys.modules.setdefault("matplotlib", mpl)
        demo = types.ModuleType("demo")
        demo.save_model = lambda *args, **kw: None
        sys.modules.setdefault("demo", demo)
        sys.path.insert(0, os.path.join(REF, "comparison_experiment"))
        import comparison_experiment.admm_l.main as ref_l      # reference module
        marks = []

        def tick(_msg):                                        # main.py calls info() once before the loop and once per epoch
            sync()
            marks.append(time.perf_counter())
        ref_l.info = tick
        # the reference draws its weights inside admm_l_demo (randn * 0.1, main.py:75-83); same distribution as the B200 arm
        torch.manual_seed(12345)
        ref_l.admm_l_demo(a.warmup + a.steps, int(d["Wy"].shape[0]), x, y, x[:1], y[:1], save=False)
        per = [marks[i + 1] - marks[i] for i in range(len(marks) - 1)]
        times = per[a.warmup:]
    else:
        if a.variant == "admm":
            import admm as ref_mod                             # reference module
        else:
            spec = importlib.util.spec_from_file_location("admm_fast", os.path.join(REF, "admm.no_dual_y.py"))
            ref_mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref_mod)
        from blocks.lstm import LSTM                           # reference module
        dd, h = d["x2i"].shape
        model = LSTM(int(dd), int(h), int(d["out"].shape[1]))
        with torch.no_grad():
            for k in [f"{s}2{g}" for g in "ifgo" for s in "xh"] + ["out"]:
                getattr(model, k).copy_(torch.from_numpy(d[k]))
        opt = ref_mod.ADMMBasedOptimizer(model, (x, y), params, verbose=False)
        for s in range(a.warmup + a.steps):
            sync()
            t0 = time.perf_counter()
            opt.step()
            sync()
            if s >= a.warmup:
                times.append(time.perf_counter() - t0)
        if a.dump:
            rec = {f"w_{k}": getattr(model, k).detach().cpu().numpy() for k in [f"{s}2{g}" for g in "ifgo" for s in "xh"] + ["out"]}
            rec.update({f"gate_{k}": v.detach().cpu().numpy() for k, v in opt.gates.items()})
            rec.update({f"dual_{k}": v.detach().cpu().numpy() for k, v in opt.duals.items()})
            np.savez(dump_path, **rec)
    out = {"step_s": times, "threads": int(torch.get_num_threads()), "device": str(dev), "torch": torch.__version__,
           "n": int(d["x"].shape[0])}
    if dev.type == "cuda":
        out["gpu"] = torch.cuda.get_device_name(0)
        out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
