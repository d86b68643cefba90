#!/usr/bin/env python
"""Compile the UNMODIFIED reference `op/` extensions into `oracle/_ref/` (test infrastructure).

This is *checker* code: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may load what it produces.  Nothing is copied from the
reference: the sources are compiled where they lie under ``/root/reference/op`` and only the
resulting shared objects land in ``oracle/_ref/`` (git-ignored, but shipped to the GPU box).

What gets built (explicit nvcc / g++ commands, no JIT cache, no reference build system):

* ``ref_rasterize``  <- op/rasterize.cpp + op/rasterize.cu (+ op/rasterize.h)
* ``ref_fused``      <- op/fused_bias_act.cpp + op/fused_bias_act_kernel.cu
* ``ref_upfirdn2d``  <- op/upfirdn2d.cpp + op/upfirdn2d_kernel.cu

The host-side math of the rasterizer is emitted by nvcc's host pass from rasterize.cu
(reference op/rasterize.cu:140-160), exactly as `torch.utils.cpp_extension.load` does in
op/rasterize.py:10-16, i.e. without -O / -march flags, hence without FMA contraction.
The device code is compiled for sm_100a so that on the B200 box the same modules double as the
"reference CUDA kernels recompiled for sm_100a" timing baseline.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_OP = os.environ.get("STYLERENDERER_REFERENCE", "/root/reference") + "/op"

MODULES = {
    "ref_rasterize": ["rasterize.cpp", "rasterize.cu"],
    "ref_fused": ["fused_bias_act.cpp", "fused_bias_act_kernel.cu"],
    "ref_upfirdn2d": ["upfirdn2d.cpp", "upfirdn2d_kernel.cu"],
}


def _torch_flags():
    import torch  # noqa: F401  (needed for the path helpers)
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    lib = ce.library_paths()
    return inc, lib


def reference_available() -> bool:
    return os.path.isdir(REF_OP)


def built(name: str) -> str:
    return os.path.join(OUT, name + ".so")


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when every module exists after the call."""
    if not reference_available():
        return all(os.path.exists(built(n)) for n in MODULES)
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    inc, lib = _torch_flags()
    inc_flags = [f"-I{p}" for p in inc]
    common = ["-std=c++17", "-D_GLIBCXX_USE_CXX11_ABI=1", "-DTORCH_API_INCLUDE_EXTENSION_H", "-w"]
    for name, srcs in MODULES.items():
        target = built(name)
        src_paths = [os.path.join(REF_OP, s) for s in srcs]
        if not force and os.path.exists(target) and all(
            os.path.getmtime(target) >= os.path.getmtime(s) for s in src_paths
        ):
            continue
        objs = []
        for s in src_paths:
            o = os.path.join(OUT, "obj", name + "_" + os.path.basename(s).replace(".", "_") + ".o")
            if s.endswith(".cu"):
                cmd = ["nvcc", "-c", s, "-o", o, "-gencode", "arch=compute_100a,code=sm_100a",
                       "--compiler-options", "-fPIC", f"-DTORCH_EXTENSION_NAME={name}",
                       "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                       "--expt-relaxed-constexpr"] + common + inc_flags
            else:
                cmd = ["g++", "-c", s, "-o", o, "-fPIC", f"-DTORCH_EXTENSION_NAME={name}"] + common + inc_flags
            if verbose:
                print("[oracle/_ref]", " ".join(cmd[:6]), "...", flush=True)
            subprocess.check_call(cmd)
            objs.append(o)
        link = ["g++", "-shared", "-o", target] + objs + [f"-L{p}" for p in lib] + [
            "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-ltorch_python", "-lcudart"]
        subprocess.check_call(link)
        if verbose:
            print("[oracle/_ref] built", target, flush=True)
    return all(os.path.exists(built(n)) for n in MODULES)


def load(name: str):
    """Import one of the compiled reference modules (torch must be importable)."""
    import importlib.util

    import torch  # noqa: F401  loads libtorch*.so the module links against

    path = built(name)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref complete:", ok)
    sys.exit(0 if ok else 1)
