"""Import the UNMODIFIED reference package from /root/reference (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so nothing that runs there may call
this module; it is used by ``oracle/make_golden.py`` and by the CPU-only tests that pin ``oracle/smc_oracle.py``
against the real reference (they skip when the reference tree is absent).
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("SMCB_REFERENCE_ROOT", "/root/reference")
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyfilter"))


def load_reference():
    """Returns the reference ``pyfilter`` module (imported with stand-ins for stochproc/pyro/matplotlib/statsmodels)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_STANDINS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("pyfilter")


def load_reference_resampling():
    """``pyfilter/{constants,utils,resampling}.py`` need torch only: load them by path, no stand-ins involved."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib.util
    import types

    pkg = types.ModuleType("_ref_pyfilter")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pyfilter")]
    sys.modules.setdefault("_ref_pyfilter", pkg)
    mods = {}
    for name in ("constants", "utils", "resampling"):
        spec = importlib.util.spec_from_file_location(
            f"_ref_pyfilter.{name}", os.path.join(REFERENCE_ROOT, "pyfilter", f"{name}.py")
        )
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"_ref_pyfilter.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods
