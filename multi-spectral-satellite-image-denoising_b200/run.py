"""Launcher: run one of the reference's scripts (scripts/train.py, evaluate.py, export.py) unchanged, with its hot-path
names rebound to libmsd_b200's implementations (SURVEY 8b "wiring").

    python -m msd_b200.run /path/to/reference/scripts/evaluate.py --checkpoint best.pth

The reference binds names with `from ... import ...` at import time, so every module that holds a copy is patched:
src.models(.satmae_restoration).SatMAERestoration, src.data(.transforms|.dataloader).AddSensorNoise,
src.training(.losses).CombinedLoss/SSIMLoss, src.training(.metrics|.trainer|.eval).MetricsTracker / calculate_*,
src.training.eval.evaluate_model.  Optional third-party modules the reference imports but the hot path never uses
(matplotlib, IPython, onnx, onnxruntime) are stubbed when absent so that `import src.utils` succeeds.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import runpy
import sys
import types
from pathlib import Path

OPTIONAL = ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "IPython", "IPython.display", "onnx", "onnxruntime")


class _Stub(types.ModuleType):
    """Stand-in for an absent optional dependency: importable, attribute access and calls return inert stubs."""

    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        child = _Stub(f"{self.__name__}.{attr}")
        setattr(self, attr, child)
        return child

    def __call__(self, *args, **kwargs):
        return self


def _stub_missing():
    for name in OPTIONAL:
        try:
            importlib.import_module(name)
        except ImportError:
            mod = _Stub(name)
            mod.__path__ = []
            # importlib.util.find_spec() on a module already in sys.modules returns its __spec__ and raises if that is None
            # (torch._dynamo.trace_rules probes "onnx" that way the first time an optimizer is built)
            mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=True)
            sys.modules[name] = mod
            if "." in name:
                setattr(sys.modules[name.rsplit(".", 1)[0]], name.rsplit(".", 1)[1], mod)


def rebind(ref_root: Path) -> list[str]:
    """Import the reference's `src` package from ref_root and patch the hot-path names.  Returns what was patched."""
    sys.path.insert(0, str(ref_root))
    _stub_missing()
    from msd_b200.data import AddSensorNoise
    from msd_b200.models import SatMAERestoration
    from msd_b200.training import CombinedLoss, SSIMLoss, evaluate_model
    from msd_b200.training import metrics as M

    table = {
        "SatMAERestoration": (SatMAERestoration, ["src.models", "src.models.satmae_restoration"]),
        "AddSensorNoise": (AddSensorNoise, ["src.data", "src.data.transforms", "src.data.dataloader"]),
        "CombinedLoss": (CombinedLoss, ["src.training", "src.training.losses"]),
        "SSIMLoss": (SSIMLoss, ["src.training", "src.training.losses"]),
        "MetricsTracker": (M.MetricsTracker, ["src.training", "src.training.metrics", "src.training.trainer", "src.training.eval"]),
        "evaluate_model": (evaluate_model, ["src.training.eval"]),
    }
    for fn in ("calculate_psnr", "calculate_ssim", "calculate_sam", "calculate_rmse", "calculate_mae"):
        table[fn] = (getattr(M, fn), ["src.training", "src.training.metrics"])
    done = []
    for name, (obj, modules) in table.items():
        for modname in modules:
            try:
                mod = importlib.import_module(modname)
            except ImportError:
                continue
            if hasattr(mod, name):
                setattr(mod, name, obj)
                done.append(f"{modname}.{name}")
    return done


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) >= 2 and argv[0] == "--log-level":
        # The reference's dataset module logs a warning at import time, which configures the root logger at WARNING before the
        # script's own basicConfig(level=INFO) runs (a no-op by then): its INFO lines -- the evaluation results among them -- are
        # then invisible.  This launcher option configures logging first; the script itself is untouched.
        import logging
        logging.basicConfig(level=getattr(logging, argv[1].upper()), force=True)
        argv = argv[2:]
    if not argv:
        raise SystemExit("usage: python -m msd_b200.run [--log-level LEVEL] <reference>/scripts/<script>.py [script args]")
    script = Path(argv[0]).resolve()
    patched = rebind(script.parent.parent)
    print(f"[msd_b200.run] rebound {len(patched)} names to libmsd_b200", file=sys.stderr)
    sys.argv = [str(script)] + argv[1:]
    if script.name == "export.py":
        # scripts/export.py builds the model on the CPU, traces it and evaluates it there as the ONNX ground truth
        # (src/deployment/export_onnx.py:40-69, 216-219): the only script that needs the torch-op export graph
        from msd_b200.models import export_mode
        with export_mode():
            runpy.run_path(str(script), run_name="__main__")
        return
    runpy.run_path(str(script), run_name="__main__")


if __name__ == "__main__":
    main()
