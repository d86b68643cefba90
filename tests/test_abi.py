"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/stylerenderer_b200.h declares; the kernel entry points refuse CPU tensors (no fallback) and the product never
imports oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "stylerenderer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from stylerenderer_b200 import _lib, build
    build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert set(_lib.EXPORTS) == set(names), "ctypes signature table and header disagree"
    assert _lib.lib().sr_abi_version() >= 1


def test_sm100a_code_is_in_the_library():
    from stylerenderer_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_and_argument_errors():
    """The kernel entry points refuse CPU tensors (the C-ABI path never falls back).  The two dispatchers the reference
    itself routes by device -- op.upfirdn2d / op.fused_leaky_relu (reference op/upfirdn2d.py:146-150, op/fused_act.py:87-94)
    -- take plain torch ops for CPU tensors like the reference does (tests/test_cpu_dispatch.py)."""
    from stylerenderer_b200 import op
    x = torch.zeros(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        op.upfirdn2d_raw(x.view(3, 8, 8, 1), torch.ones(4, 4), 1, 1, 1, 1, 0, 0, 0, 0)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        op.fused_bias_act(x, torch.zeros(3), None, 3, 0, 0.2, 1.0)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        op.rasterize(torch.zeros(1, 3, 3), torch.zeros(1, 3, 2), torch.zeros(1, 3, dtype=torch.int64), 4)
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    L = __import__("stylerenderer_b200")._lib.lib()
    assert L.sr_upfirdn2d_f32(None, None, None, 1, 4, 4, 1, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, None) == -1
    assert b"up/down" in L.sr_last_error()
    assert L.sr_rasterize_forward_f32(1, 3, 1, 4, 8, 0, 1, 0, None, None, None, None, None, 1e-6, None, 0, None, None) == -1
    assert b"square" in L.sr_last_error()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        op.rasterize_pyramid(torch.zeros(1, 3, 3), torch.zeros(1, 3, 2), torch.zeros(1, 3, dtype=torch.int64), [4, 8])
    assert L.sr_rasterize_pyramid_forward_f32(1, 3, 1, 0, None, 0, 1, 0, None, None, None, 1e-6, None, 0, None) == -1
    assert b"levels" in L.sr_last_error()
    from stylerenderer_b200.op.rasterize import RasterLevel
    lv = (RasterLevel * 1)()
    lv[0].size = 0
    assert L.sr_rasterize_pyramid_forward_f32(1, 3, 1, 1, lv, 0, 1, 0, None, None, None, 1e-6, None, 0, None) == -1
    assert b"level size" in L.sr_last_error()
    import ctypes
    sizes = (ctypes.c_int64 * 3)(4, 8, 16)
    # keys (8 B per pixel of every level) + pre-pass products (projected vertices per level, int32 triangles) + slack
    assert L.sr_rasterize_pyramid_workspace_bytes(2, 5, 7, 3, sizes) == 2 * (16 + 64 + 256) * 8 + 2 * 5 * 12 * 3 + 8 + 2 * 7 * 12 + 8 + 16


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "stylerenderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "libsr_oracle" not in src and "oracle/_ref" not in src, f
    code = "import sys; import stylerenderer_b200.op, stylerenderer_b200.layers, stylerenderer_b200.model; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_fastdiv_magic_numbers():
    """Host model of sr::FastDiv (csrc/common.cuh) -- exact for every 32-bit numerator we can hit."""
    import random

    def magic(d):
        if d == 1:
            return 0, 0
        l = 0
        while (1 << l) < d:
            l += 1
        return (((1 << 32) * ((1 << l) - d)) // d + 1) & 0xffffffff, l

    def div(n, d, m, s):
        if d == 1:
            return n
        t = (n * m) >> 32
        return ((t + ((n - t) >> 1)) & 0xffffffff) >> (s - 1)

    rnd = random.Random(0)
    for d in list(range(1, 70)) + [127, 128, 129, 255, 256, 257, 512, 4096, 65536, 16641, 66049, 2 ** 31 - 1] + \
            [rnd.randrange(1, 2 ** 31) for _ in range(200)]:
        m, s = magic(d)
        for n in [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 32 - 1, 2 ** 31] + [rnd.randrange(0, 2 ** 32) for _ in range(200)]:
            n &= 0xffffffff
            assert div(n, d, m, s) == n // d, (n, d)
