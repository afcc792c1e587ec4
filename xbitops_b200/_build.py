"""In-tree builds (sm_100a only). The built .so files are git-ignored but travel to the GPU box.

  libxbitops_b200.so   nvcc, torch-free: the C ABI (include/xbitops_b200.h) + all kernels
  XbitOps*.so          g++ (no nvcc, no kernels): the pybind11/ATen shim over the C ABI, i.e. the
                       drop-in for the reference's extension module of the same name
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libxbitops_b200.so"
LIB_DEV = PKG / "libxbitops_b200_dev.so"      # -DXBIT_DEVTOOLS: phase stamps / skip-math knobs for tools/*.py only
OBJ = PKG / "_obj"
CUDA_SOURCES = ["xbit_capi.cu", "dq_sm100.cu", "gemv_sm100.cu", "gemv_w4p_sm100.cu"]
CUDA_DEPS = CUDA_SOURCES + ["unpack.cuh", "gemv_prims.cuh", "xbit_internal.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden"]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).exists() and Path(d).stat().st_mtime > t for d in deps)


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (Path(c).exists() or c == "nvcc"):
            return c
    return "nvcc"


def build_lib(force: bool = False, verbose: bool = False, dev: bool = False) -> Path:
    """One object per .cu (compiled in parallel), then one shared library."""
    from concurrent.futures import ThreadPoolExecutor
    lib = LIB_DEV if dev else LIB
    hdrs = [CSRC / d for d in CUDA_DEPS if not d.endswith(".cu")] + [ROOT / "include" / "xbitops_b200.h"]
    OBJ.mkdir(exist_ok=True)
    extra = ["-DXBIT_DEVTOOLS"] if dev else []
    jobs, objs = [], []
    for s in CUDA_SOURCES:
        o = OBJ / (s[:-3] + ("_dev.o" if dev else ".o"))
        objs.append(o)
        if force or _newer(o, [CSRC / s] + hdrs):
            jobs.append([nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", str(o), str(CSRC / s)])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(len(jobs)) as ex:
            logs = list(ex.map(run, jobs))
        if verbose:
            print("\n".join(logs))
    if jobs or force or _newer(lib, objs):
        run([nvcc_path(), "-shared", "-o", str(lib)] + [str(o) for o in objs])
    return lib


def torch_ext_path() -> Path | None:
    hits = sorted(PKG.glob("XbitOps*.so"))
    return hits[0] if hits else None


def build_torch_ext(force: bool = False) -> Path:
    """pybind11/ATen shim `XbitOps` (same module name as the reference extension), linked against
    libxbitops_b200.so with an $ORIGIN rpath."""
    build_lib()
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    out = PKG / f"XbitOps{suffix}"
    src = CSRC / "dq_torch_ops.cc"
    if not (force or _newer(out, [src, ROOT / "include" / "xbitops_b200.h"])):
        return out
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    libdirs = [f"-L{p}" for p in ce.library_paths("cuda")]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-fvisibility=hidden", "-Wl,-Bsymbolic", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
           "-DTORCH_EXTENSION_NAME=XbitOps", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DUSE_CUDA",
           *inc, str(src), "-o", str(out), *libdirs, f"-L{PKG}", "-lxbitops_b200",
           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
           "-Wl,-rpath,$ORIGIN", *[f"-Wl,-rpath,{p}" for p in ce.library_paths("cuda")]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return out


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--dev" in sys.argv:
        print(build_lib(force="--force" in sys.argv, dev=True))
    if "--torch" in sys.argv:
        print(build_torch_ext(force="--force" in sys.argv))
