"""Import the reference's own util modules from /root/reference WITHOUT importing `marigold/__init__`
(which needs diffusers). Only usable in the build container; used by make_golden.py to produce the
committed fixtures. matplotlib is stubbed (only colorize_depth_maps uses it)."""
import importlib.util
import sys
import types
from pathlib import Path

REF = Path("/root/reference")


def load_reference_utils():
    if not REF.exists():
        raise RuntimeError("/root/reference is not available (fixtures are generated in the build container only)")
    if "matplotlib" not in sys.modules:
        sys.modules["matplotlib"] = types.ModuleType("matplotlib")
    pkg = types.ModuleType("refmarigold")
    pkg.__path__ = [str(REF / "marigold")]
    sys.modules["refmarigold"] = pkg
    util = types.ModuleType("refmarigold.util")
    util.__path__ = [str(REF / "marigold" / "util")]
    sys.modules["refmarigold.util"] = util
    mods = {}
    for name in ("image_util", "ensemble", "batchsize"):
        spec = importlib.util.spec_from_file_location(f"refmarigold.util.{name}", REF / "marigold" / "util" / f"{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods
