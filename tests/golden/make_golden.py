"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  Usage:
    python tests/golden/make_golden.py [--out tests/golden]
It imports /root/reference/admm.py and admm.no_dual_y.py as they are, drives their public
`step()` and (through the name-mangled attributes) their private per-function updates, and
stores inputs + outputs as .npz.  The reference has no tests/golden vectors of its own
(SURVEY.md section 4), so these files are the parity pin for oracle/ and for the CUDA path.
"""
import argparse
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference():
    work = tempfile.mkdtemp(prefix="admm_ref_")      # the reference writes logs/ under CWD
    os.chdir(work)
    sys.path.insert(0, REF)
    import torch
    torch.manual_seed(0)
    import admm as ref_admm                            # noqa: E402  (reference module)
    spec = importlib.util.spec_from_file_location("admm_fast", os.path.join(REF, "admm.no_dual_y.py"))
    ref_fast = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_fast)
    from blocks.lstm import LSTM                       # noqa: E402
    from parameters import example_parameter_dictionary  # noqa: E402
    return torch, ref_admm, ref_fast, LSTM, example_parameter_dictionary


def weights_of(model):
    out = {}
    for g in "ifgo":
        out["x2" + g] = getattr(model, "x2" + g).detach().numpy().copy()
        out["h2" + g] = getattr(model, "h2" + g).detach().numpy().copy()
    out["out"] = model.out.detach().numpy().copy()
    return out


def state_of(opt, prefix, rows=None):
    out = {}
    sl = slice(None) if rows is None else slice(0, rows)
    for k, v in opt.gates.items():
        out[f"{prefix}gate_{k}"] = v.detach().numpy()[sl].copy()
    for k, v in opt.duals.items():
        out[f"{prefix}dual_{k}"] = v.detach().numpy()[sl].copy()
    return out


def flat_w(prefix, w):
    return {f"{prefix}w_{k}": v for k, v in w.items()}


def mse(torch, model, x, y):
    with torch.no_grad():
        return float(torch.nn.functional.mse_loss(model(x), y))


def synthetic_case(torch, mod, LSTM, params, variant, with_dual_y, n=37, t=5, d=3, h=6, o=2, steps=4, seed=3):
    """Trajectory fixture: `steps` full step() calls on a tiny random problem."""
    torch.manual_seed(seed)
    x = torch.rand(n, t, d)
    y = torch.rand(n, o)
    model = LSTM(d, h, o)
    rec = {"x": x.numpy().copy(), "y": y.numpy().copy()}
    rec.update(flat_w("init_", weights_of(model)))
    if hasattr(mod, "with_dual_y"):
        mod.with_dual_y = with_dual_y
    opt = mod.ADMMBasedOptimizer(model, (x, y), params, verbose=False)
    rec.update(state_of(opt, "s0_"))
    losses = [mse(torch, model, x, y)]
    for s in range(1, steps + 1):
        opt.step()
        rec.update(flat_w(f"s{s}_", weights_of(model)))
        rec.update(state_of(opt, f"s{s}_"))
        losses.append(mse(torch, model, x, y))
    rec["losses"] = np.array(losses, dtype=np.float64)
    if hasattr(mod, "with_dual_y"):
        mod.with_dual_y = False
    return rec


def per_function_case(torch, mod, LSTM, params, variant, n=29, t=4, d=2, h=5, o=3, seed=11):
    """Per-function fixture: every private update called from the same random (non-trivial) state."""
    torch.manual_seed(seed)
    x = torch.rand(n, t, d)
    y = torch.rand(n, o)
    model = LSTM(d, h, o)
    base_w = weights_of(model)
    opt = mod.ADMMBasedOptimizer(model, (x, y), params, verbose=False)
    # a generic interior state: forward values perturbed, duals non-zero everywhere
    gen = torch.Generator().manual_seed(seed + 1)
    for k in "ifgoch":
        opt.gates[k] = (opt.gates[k].detach() + 0.1 * torch.randn(opt.gates[k].shape, generator=gen)).clone()
        opt.gates[k][:, 0, :] = 0
        opt.duals[k] = 0.05 * torch.randn(opt.duals[k].shape, generator=gen)
        opt.duals[k][:, 0, :] = 0
    opt.gates["a"] = (opt.gates["a"].detach() + 0.1 * torch.randn(opt.gates["a"].shape, generator=gen)).clone()
    base_g = {k: v.detach().clone() for k, v in opt.gates.items()}
    base_d = {k: v.detach().clone() for k, v in opt.duals.items()}
    rec = {"x": x.numpy().copy(), "y": y.numpy().copy()}
    rec.update(flat_w("base_", base_w))
    rec.update({f"base_gate_{k}": v.numpy().copy() for k, v in base_g.items()})
    rec.update({f"base_dual_{k}": v.numpy().copy() for k, v in base_d.items()})

    def restore():
        for k, v in base_g.items():
            opt.gates[k] = v.clone()
        for k, v in base_d.items():
            opt.duals[k] = v.clone()
        for name, val in base_w.items():
            setattr(model, name, torch.nn.Parameter(torch.tensor(val)))

    P = "_ADMMBasedOptimizer__"
    restore()
    getattr(opt, P + "update_wy")()
    rec["fn_wy"] = model.out.detach().numpy().copy()
    for src, g in (("x", "i"), ("h", "i"), ("x", "g"), ("h", "g"), ("x", "f"), ("h", "o")):
        restore()
        getattr(opt, P + "update_weights")(src, g)
        rec[f"fn_w_{src}2{g}"] = getattr(model, f"{src}2{g}").detach().numpy().copy()
    tt = 2
    for g in "ifgo":
        restore()
        getattr(opt, P + "update_primal_i_f_g_o")(g, tt)
        rec[f"fn_primal_{g}"] = opt.gates[g][:, tt, :].detach().numpy().copy()
    restore()
    getattr(opt, P + "update_primal_c")(tt)
    rec["fn_primal_c"] = opt.gates["c"][:, tt, :].detach().numpy().copy()
    restore()
    getattr(opt, P + "update_primal_h")(tt)
    rec["fn_primal_h_mid"] = opt.gates["h"][:, tt, :].detach().numpy().copy()
    restore()
    getattr(opt, P + "update_primal_h")(t)
    rec["fn_primal_h_last"] = opt.gates["h"][:, t, :].detach().numpy().copy()
    restore()
    getattr(opt, P + "update_primal_a")()
    rec["fn_primal_a"] = opt.gates["a"].detach().numpy().copy()
    for g in "ifgo":
        restore()
        getattr(opt, P + "update_dual_i_f_g_o")(g, tt)
        rec[f"fn_dual_{g}"] = opt.duals[g][:, tt, :].detach().numpy().copy()
    restore()
    getattr(opt, P + "update_dual_c")(tt)
    rec["fn_dual_c"] = opt.duals["c"][:, tt, :].detach().numpy().copy()
    restore()
    getattr(opt, P + "update_dual_h")(t)
    rec["fn_dual_h"] = opt.duals["h"][:, t, :].detach().numpy().copy()
    rec["tt"] = np.array(tt)
    return rec


def google_stock(torch):
    sys.path.insert(0, HERE)
    import xls_reader
    sheet = xls_reader.open_workbook(os.path.join(REF, "datasets/GoogleStock/GOOG.xls")).sheet_by_index(0)
    # dataset.py:404-440, restated (col 5 -> X, col 4 -> Y, each / its max, windows of 10)
    X = torch.tensor([sheet.cell_value(i, 5) for i in range(1, 4706)], dtype=torch.float32)
    Y = torch.tensor([sheet.cell_value(i, 4) for i in range(1, 4706)], dtype=torch.float32)
    X, Y = X / X.max(), Y / Y.max()
    train_x = torch.stack([X[i - 10:i] for i in range(10, 4234)]).unsqueeze(2)
    train_y = torch.stack([Y[i] for i in range(10, 4234)]).reshape(4224, 1)
    val_x = torch.stack([X[i - 10:i] for i in range(4244, 4705)]).unsqueeze(2)
    val_y = torch.stack([Y[i] for i in range(4244, 4705)]).reshape(461, 1)
    return train_x, train_y, val_x, val_y


def real_case(torch, mod, LSTM, params, data, hidden, iters, snap_iters, rows=48):
    """demo.py:383-409 + admm_demo (demo.py:311-376) restated around the unmodified optimizer."""
    train_x, train_y, val_x, val_y = data
    torch.manual_seed(0)                                      # demo.py:34,282
    model = LSTM(train_x.size(2), hidden, train_y.size(1))
    rec = flat_w("init_", weights_of(model))
    opt = mod.ADMMBasedOptimizer(model=model, training_samples=(train_x, train_y), parameter_dictionary=params)
    tr, va = [mse(torch, model, train_x, train_y)], [mse(torch, model, val_x, val_y)]
    wtraj = {k: [v] for k, v in weights_of(model).items()}
    for it in range(1, iters + 1):
        opt.step()
        tr.append(mse(torch, model, train_x, train_y))
        va.append(mse(torch, model, val_x, val_y))
        for k, v in weights_of(model).items():
            wtraj[k].append(v)
        if it in snap_iters:
            rec.update(state_of(opt, f"it{it}_", rows=rows))
    rec["train_loss"] = np.array(tr)
    rec["val_loss"] = np.array(va)
    for k, v in wtraj.items():
        rec["wtraj_" + k] = np.stack(v)
    return rec


def gefcom_standin(torch):
    """Load_history.csv is absent; temperature_history.csv has the same format and
    loads through the unmodified dataset.GEFCom2012 class."""
    for name in ("av", "cv2", "torchvision", "torchvision.transforms", "tqdm", "matplotlib", "matplotlib.pyplot",
                 "xlrd", "yfinance", "wfdb", "sklearn", "sklearn.preprocessing", "sklearn.model_selection"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__.setdefault("__path__", [])
                sys.modules[name] = m
    av = sys.modules["av"]
    if not hasattr(av, "container"):
        av.container = types.SimpleNamespace(InputContainer=object)
        av.InvalidDataError = Exception
    if not hasattr(sys.modules["tqdm"], "tqdm"):
        sys.modules["tqdm"].tqdm = lambda it, *a, **k: it
    from dataset import GEFCom2012
    path = os.path.join(REF, "datasets/GEFCOM2012_Data")

    def load(obj):
        samples = GEFCom2012(path=path, load_object=obj).sample_dict["temperature_history"]
        xs, ys = zip(*[(x, y) for x, y in samples])
        return torch.stack(xs), torch.stack(ys)

    tx, ty = load("temperature_history/1~20")
    vx, vy = load("temperature_history/21~30")
    return tx, ty, vx, vy


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--skip-real", action="store_true")
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    torch, ref_admm, ref_fast, LSTM, epd = load_reference()
    torch.set_num_threads(1)                                   # reproducible reductions
    gs = epd["GoogleStock"]
    har = epd["HAR"]

    for variant, mod in (("admm", ref_admm), ("no_dual_y", ref_fast)):
        np.savez_compressed(os.path.join(out, f"fn_{variant}.npz"),
                            **per_function_case(torch, mod, LSTM, gs, variant))
        np.savez_compressed(os.path.join(out, f"traj_{variant}.npz"),
                            **synthetic_case(torch, mod, LSTM, gs, variant, False))
        np.savez_compressed(os.path.join(out, f"traj_har_{variant}.npz"),
                            **synthetic_case(torch, mod, LSTM, har, variant, False, n=41, t=6, d=4, h=8, o=3, seed=5))
    np.savez_compressed(os.path.join(out, "traj_admm_dualy.npz"),
                        **synthetic_case(torch, ref_admm, LSTM, gs, "admm", True))
    if args.skip_real:
        return
    data = google_stock(torch)
    np.savez_compressed(os.path.join(out, "googlestock_data.npz"),
                        train_x=data[0].numpy(), train_y=data[1].numpy(), val_x=data[2].numpy(), val_y=data[3].numpy())
    for variant, mod in (("admm", ref_admm), ("no_dual_y", ref_fast)):
        rec = real_case(torch, mod, LSTM, gs, data, hidden=10, iters=50, snap_iters=(1, 10, 50))
        np.savez_compressed(os.path.join(out, f"googlestock_{variant}.npz"), **rec)
        print(variant, "GoogleStock train", rec["train_loss"][[0, 1, 2, 30, 50]], "val", rec["val_loss"][[0, 30, 50]])
    try:
        gdata = gefcom_standin(torch)
        np.savez_compressed(os.path.join(out, "gefcom_standin_data.npz"),
                            train_x=gdata[0].numpy(), train_y=gdata[1].numpy(),
                            val_x=gdata[2].numpy(), val_y=gdata[3].numpy())
        fast_rho = dict(epd["GEFCOM2012"]["rho"], h=0.001, y=0.0001)     # parameters.py:24 ("without dual y")
        for variant, mod, params in (("admm", ref_admm, epd["GEFCOM2012"]),
                                     ("no_dual_y", ref_fast, {"rho": fast_rho, "beta": epd["GEFCOM2012"]["beta"]})):
            rec = real_case(torch, mod, LSTM, params, gdata, hidden=10, iters=50, snap_iters=(1, 50))
            np.savez_compressed(os.path.join(out, f"gefcom_standin_{variant}.npz"), **rec)
            print(variant, "GEFCOM stand-in train", rec["train_loss"][[0, 1, 2, 50]])
    except Exception as exc:                                    # data loader needs stubs; not fatal
        print("GEFCOM stand-in skipped:", repr(exc))


if __name__ == "__main__":
    main()
