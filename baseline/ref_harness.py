"""Run-time access to the UNMODIFIED reference sources (SDU-Gary/PCSS-Unet) -- measurement / test infrastructure only.

The reference is a flat directory of Python scripts (no setup.py / pyproject, so `pip install --target baseline/_ref`
has nothing to install) and `/root/reference` does not exist on the GPU box.  `stage()` therefore copies the handful of
reference files the hot path and its call sites consist of into the git-ignored `baseline/_ref/` (never committed,
but shipped to the GPU box by gpurun like the built .so).  `__graft_entry__.build()` calls it in the build container.

Users: `bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg (time the reference's own `Unet` on the host
cores), `tests/` (drive the reference's `main.train_model` / `infer.main` call sites against the drop-in).  Nothing
under `pcss-unet_b200/` imports this file.

Modules the reference imports but never uses on the path and that are not installed here (graphviz, pytorch_msssim,
OpenEXR, Imath, colorama, matplotlib) are replaced by empty stubs, as SURVEY.md 8c describes.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"
FILES = ("Unetmodel.py", "customLoss.py", "pert_loss.py", "calculate_dataset_stats.py", "setdata.py", "main.py",
         "infer.py", "visualize.py", "config.ini")


def stage(source: str = SOURCE) -> bool:
    """Copy the reference files into baseline/_ref/ (byte for byte).  Returns True when baseline/_ref is usable."""
    if os.path.isdir(source):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in FILES:
            src, dst = os.path.join(source, f), os.path.join(REF_DIR, f)
            if os.path.exists(src) and (not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read()):
                shutil.copyfile(src, dst)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("Unetmodel.py", "customLoss.py", "pert_loss.py"))


def install_stubs() -> None:
    """Empty stand-ins for third-party modules the reference imports at module scope but does not use on the path."""
    def stub(name, **attrs):
        try:
            if name not in sys.modules:
                importlib.import_module(name)
            return
        except Exception:
            pass
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m

    stub("graphviz", Digraph=object)
    stub("pytorch_msssim", ssim=None)
    stub("OpenEXR")
    stub("Imath")

    class _Codes:
        def __getattr__(self, k):
            return ""
    stub("colorama", init=lambda *a, **k: None, Fore=_Codes(), Style=_Codes(), Back=_Codes())
    stub("matplotlib")
    if "matplotlib.pyplot" not in sys.modules and isinstance(sys.modules.get("matplotlib"), types.ModuleType) \
            and not hasattr(sys.modules["matplotlib"], "__path__"):
        sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")


def load_module(filename: str, as_name: str):
    """Import one reference file under a private module name (so that it can live next to the drop-in module of the
    same name).  Only for files that import nothing else from the reference but visualize.py (Unetmodel.py)."""
    if not available():
        raise RuntimeError("baseline/_ref is empty: run __graft_entry__.build() in the build container first")
    install_stubs()
    if as_name in sys.modules:
        return sys.modules[as_name]
    if "visualize" not in sys.modules:
        vspec = importlib.util.spec_from_file_location("visualize", os.path.join(REF_DIR, "visualize.py"))
        vmod = importlib.util.module_from_spec(vspec)
        sys.modules["visualize"] = vmod
        vspec.loader.exec_module(vmod)
    spec = importlib.util.spec_from_file_location(as_name, os.path.join(REF_DIR, filename))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[as_name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_unet_class():
    """The reference's own `Unet` (Unetmodel.py:36-149), unmodified."""
    return load_module("Unetmodel.py", "_nsm_reference_Unetmodel").Unet
