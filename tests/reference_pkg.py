"""Test infrastructure: imports the reference's UNMODIFIED Python package (LoFloat/{__init__,_custom_ops,formats,layers,
utils}.py) on top of this repository's compiled extension.

The package is read from oracle/_ref/lofloat_ref_py.zip -- an archive `make -C oracle refpy` builds from /root/reference
where that tree exists (git-ignored build output like the compiled reference shim; it travels to the GPU box, the sources are
never unpacked into the repository) -- with `LoFloat.LoFloat`, the name of the reference's compiled module
(src/LoPy_bind.cpp), bound to lofloat_b200/LoFloat.so."""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ZIP = ROOT / "oracle" / "_ref" / "lofloat_ref_py.zip"
EXT = ROOT / "lofloat_b200" / "LoFloat.so"


def available() -> bool:
    if not ZIP.exists() and Path("/root/reference/LoFloat/layers.py").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "refpy"], check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return ZIP.exists() and EXT.exists()


def load_extension():
    """lofloat_b200/LoFloat.so as the module `LoFloat.LoFloat` (PyInit_LoFloat)."""
    name = "LoFloat.LoFloat"
    if name in sys.modules:
        return sys.modules[name]
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    loader = importlib.machinery.ExtensionFileLoader(name, str(EXT))
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def load():
    """-> (the reference's `LoFloat` package, the compiled extension module)."""
    ext = load_extension()
    if "LoFloat" not in sys.modules:
        sys.path.insert(0, str(ZIP))
        try:
            importlib.import_module("LoFloat")
        finally:
            sys.path.remove(str(ZIP))
    return sys.modules["LoFloat"], ext
