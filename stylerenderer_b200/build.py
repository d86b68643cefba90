"""Build libstylerenderer_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA C++ with an `extern "C"` surface
(include/stylerenderer_b200.h); Python reaches it through ctypes (stylerenderer_b200/_lib.py).
nvcc cross-compiles without a GPU, so this runs in the authoring container and the resulting .so
travels to the B200 box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libstylerenderer_b200.so")
SOURCES = ["lib.cu", "fused_bias_act.cu", "upfirdn2d.cu", "rasterize.cu", "modconv.cu", "styled_ops.cu", "style_ops.cu", "mesh_ops.cu", "stylemap_net.cu", "stem_conv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--compiler-options", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "stylerenderer_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = ["nvcc", "-c", s, "-o", o] + NVCC_FLAGS
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(f"--- nvcc {src}\n{out}")
        with open(os.path.join(objdir, src + ".ptxas.log"), "w") as f:
            f.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call(["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                       "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
