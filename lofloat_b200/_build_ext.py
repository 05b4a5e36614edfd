"""Builds lofloat_b200/LoFloat.so -- the compiled PyTorch extension (csrc/ext/lofloat_ext.cpp, pybind11) that carries the
Python surface of the reference's `LoFloat.LoFloat` module on top of liblofloat_b200.so.  In-tree, plain g++ (the
extension has no device code of its own); called by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "ext" / "lofloat_ext.cpp"
OUT = PKG / "LoFloat.so"
LIB = PKG / "liblofloat_b200.so"


def build_extension(force: bool = False, verbose: bool = False) -> Path:
    deps = [SRC, PKG.parent / "include" / "lofloat_b200.h", LIB]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps if d.exists()):
        return OUT
    if not LIB.exists():
        raise RuntimeError(f"{LIB} must be built first (make -C lofloat_b200/csrc)")
    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=LoFloat",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for name in ("COMPILER_TYPE", "STDLIB", "BUILD_ABI"):                      # match torch's pybind11 internals where it states them
        val = getattr(torch._C, f"_PYBIND11_{name}", None)
        if val is not None:
            cmd.append(f'-DPYBIND11_{name}="{val}"')
    for inc in ce.include_paths() + [sysconfig.get_paths()["include"], f"{cuda_home}/include", str(PKG.parent / "include")]:
        cmd += ["-isystem" if "torch" in inc or "cuda" in inc else "-I", inc]
    cmd += [str(SRC), "-o", str(OUT)]
    for lp in ce.library_paths():
        cmd += [f"-L{lp}", f"-Wl,-rpath,{lp}"]
    cmd += [f"-L{PKG}", "-Wl,-rpath,$ORIGIN", "-llofloat_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0 or verbose:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("building lofloat_b200/LoFloat.so failed")
    return OUT


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
