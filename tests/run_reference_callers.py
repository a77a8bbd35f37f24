"""Runs the reference's CALLERS unchanged against this repo's drop-in modules (helper of tests/test_gpu_callers.py; runs in
its own process so that module resolution is exactly what a maintainer would get).

Integration scenario of INTEGRATION.md: the reference tree as it is (here: the byte-for-byte copy staged in oracle/_ref by
oracle/make_ref.py) with ONLY the hot-path modules replaced --
    admm.py                                  <- this repo's admm.py (or admm.no_dual_y.py copied over it, the way the
                                                reference itself selects the Fast variant)
    comparison_experiment/admm_l/main.py     <- this repo's drop-in
and the admm_lstm_b200 package importable.  demo.py, comparison.py, blocks/lstm.py, _global.py, parameters.py, dataset.py,
data_plot.py, grad_based.py are the reference's own files.  Stubs only for what this image lacks: matplotlib (plots),
av (video datasets), xlrd (an .xls reader: tests/golden/xls_reader.py implements the three calls dataset.py:392-405 makes).

    python tests/run_reference_callers.py demo|comparison --variant admm|no_dual_y --epochs K --out result.json
"""
import argparse
import json
import os
import runpy
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")


class _Anything:
    """Stands in for matplotlib objects: every attribute / call / unpacking yields another one."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))


def _stub(name, **attrs):
    m = types.ModuleType(name)

    def module_getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()
    m.__getattr__ = module_getattr
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("caller", choices=["demo", "comparison"])
    ap.add_argument("--variant", default="admm", choices=["admm", "no_dual_y"])
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--out", required=True)
    ap.add_argument("--self-check", action="store_true",
                    help="overlay the REFERENCE's own admm.py / admm_l/main.py instead of this repo's: checks the harness (and "
                         "reproduces the golden curves) on a CPU-only box")
    a = ap.parse_args()
    out_path = os.path.abspath(a.out)

    work = tempfile.mkdtemp(prefix="ref_callers_")
    overlay = os.path.join(work, "overlay")
    os.makedirs(os.path.join(overlay, "comparison_experiment", "admm_l"))
    src_tree = REF if a.self_check else ROOT
    shutil.copyfile(os.path.join(src_tree, "admm.py" if a.variant == "admm" else "admm.no_dual_y.py"), os.path.join(overlay, "admm.py"))
    shutil.copyfile(os.path.join(src_tree, "comparison_experiment", "admm_l", "main.py"),
                    os.path.join(overlay, "comparison_experiment", "admm_l", "main.py"))
    os.symlink(os.path.join(REF, "datasets"), os.path.join(work, "datasets"))
    os.symlink(os.path.join(ROOT, "admm_lstm_b200"), os.path.join(overlay, "admm_lstm_b200"))     # the package, dropped into the tree
    os.chdir(work)                                   # logs/, plots/, SAVED_MODELS/ and the relative dataset path

    # module resolution: overlay first, then the reference tree.  The repo root is NOT on the path (its blocks/ and
    # comparison_experiment/ are regular packages and would shadow the reference's namespace packages wherever they sit).
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in (ROOT, HERE)]
    sys.path[0:0] = [overlay, os.path.join(overlay, "comparison_experiment"), REF, os.path.join(REF, "comparison_experiment")]

    mpl = _stub("matplotlib", use=lambda *x, **k: None)
    mpl.pyplot = _stub("matplotlib.pyplot")
    av = _stub("av", InvalidDataError=type("InvalidDataError", (Exception,), {}))
    av.container = _stub("av.container", InputContainer=type("InputContainer", (), {}))
    sys.path.append(os.path.join(HERE, "golden"))
    import xls_reader
    sys.modules["xlrd"] = xls_reader

    import torch
    result = {}
    if a.caller == "demo":
        sys.argv = ["demo.py", "--dataset", "GoogleStock", "--epoch", str(a.epochs), "--yes"]
        import demo                                   # the reference's demo.py
        import admm                                   # must be the overlay (this repo's module)
        import blocks.lstm
        assert os.path.dirname(os.path.abspath(demo.__file__)) == REF, demo.__file__
        assert os.path.dirname(os.path.abspath(admm.__file__)) == overlay, admm.__file__
        assert os.path.dirname(os.path.abspath(blocks.lstm.__file__)) == os.path.join(REF, "blocks")
        result = demo.training_demo(plot=False)       # demo.py:383-409, unchanged
        result["optimizer_class"] = f"{demo.ADMMBasedOptimizer.__module__}.{demo.ADMMBasedOptimizer.__mro__[1].__module__}"
    else:
        sys.argv = ["comparison.py", "--dataset", "GoogleStock", "--epoch", str(a.epochs), "--yes", "--save"]
        captured = {}
        import demo
        import admm_l.main as lmain                   # resolves to the overlay (this repo's drop-in)
        assert os.path.abspath(lmain.__file__).startswith(overlay), lmain.__file__
        real_admm_demo, real_l_demo = demo.admm_demo, lmain.admm_l_demo

        def rec_admm(*args, **kw):
            captured["fast"] = real_admm_demo(*args, **kw)
            return captured["fast"]

        def rec_l(*args, **kw):
            captured["admm_l"] = real_l_demo(*args, **kw)
            return captured["admm_l"]
        demo.admm_demo, lmain.admm_l_demo = rec_admm, rec_l            # observers only: same callables, same arguments
        runpy.run_path(os.path.join(REF, "comparison_experiment", "comparison.py"), run_name="__main__")   # :141-210
        result = {"fast": captured.get("fast"), "admm_l": captured.get("admm_l"),
                  "saved": sorted(os.listdir(os.path.join(work, "SAVED_MODELS"))) if os.path.isdir(os.path.join(work, "SAVED_MODELS")) else []}
        # SAVED_MODELS layout: the reference's reader (visualization.py:47-54) uses torch.load(weights_only=False)
        m = torch.load(os.path.join(work, "SAVED_MODELS", "Fast ADMM-LSTM.pt"), weights_only=False, map_location="cpu")
        result["fast_pickle_class"] = f"{type(m).__module__}.{type(m).__name__}"
        result["fast_pickle_params"] = [n for n, _ in m.named_parameters()]
        ml = torch.load(os.path.join(work, "SAVED_MODELS", "ADMM-LSTM-L.pt"), weights_only=False, map_location="cpu")
        result["l_pickle_class"] = f"{type(ml).__module__}.{type(ml).__name__}"
    with open(out_path, "w") as f:
        json.dump(result, f)
    print("callers ok:", a.caller, a.variant)


if __name__ == "__main__":
    main()
