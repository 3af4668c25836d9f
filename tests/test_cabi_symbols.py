"""The C-ABI library loads on a CPU-only machine and exports exactly what include/vaegam.h
declares; the ctypes signature table covers every declaration.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vaegam.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from vaegam import native
    if not os.path.isfile(native.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "vae-gam_b200", "csrc"), "-j8"], check=True)
    return native


def test_header_declares_expected_surface():
    names = declared_symbols()
    for must in ("vg_conv_fwd", "vg_conv_dgrad", "vg_conv_wgrad", "vg_bn_stats", "vg_linear_fwd", "vg_linear_bwd",
                 "vg_latent_fwd", "vg_latent_bwd", "vg_gain_fwd", "vg_gain_bwd", "vg_recon_loss_fwd",
                 "vg_recon_loss_bwd", "vg_adam_step", "vg_step_fwd", "vg_step_bwd", "vg_version", "vg_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in vaegam.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    assert sorted(built_lib.SIGNATURES) == declared_symbols()


def test_library_identity_without_gpu(built_lib):
    lib = built_lib.load()
    assert lib.vg_version() == 100
    assert lib.vg_last_error() is not None
    assert lib.vg_sm_count() > 0          # 148 fallback without a device
    assert lib.vg_launch_count() >= 0


def test_sass_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_sizes_match_c_layout(built_lib):
    """ctypes mirrors of the C structs (compiled check through a tiny C program)."""
    import tempfile
    src = r'''
#include <stdio.h>
#include "vaegam.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(VgConvDesc), sizeof(VgGainParams), sizeof(VgGainGrads),
  sizeof(VgStepConfig), sizeof(VgStepIO), sizeof(VgMlp)); return 0; }'''
    d = tempfile.mkdtemp()
    open(os.path.join(d, "s.c"), "w").write(src)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")], check=True)
    sizes = [int(v) for v in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True).stdout.split()]
    n = built_lib
    assert sizes == [ctypes.sizeof(n.VgConvDesc), ctypes.sizeof(n.VgGainParams), ctypes.sizeof(n.VgGainGrads),
                     ctypes.sizeof(n.VgStepConfig), ctypes.sizeof(n.VgStepIO), ctypes.sizeof(n.VgMlp)]


def test_fused_loss_work_decomposition(built_lib):
    """Host side of the fused reconstruction pass (csrc/recon_loss.cu: make_plan / recon_finalize): for every
    launch shape the spans tile the (row group, column) items exactly once, a CTA's span crosses at most one
    row-group boundary (it writes two partial segments), and the finalize ranges visit every span of a row group."""
    lib = built_lib.load()
    out = (ctypes.c_int * 6)()
    V = 41 * 49 * 35
    for variant in (0, 1, 2, 3):
        for b in (1, 2, 3, 4, 5, 31, 32, 33, 128, 129, 512, 4096):
            for v in (V, 256, 1000, 8 * 32 * 4 * 3 + 1):
                assert lib.vg_recon_plan(b, v, variant, out) == 0
                n_rg, n_c8, n_items, span, grid, warps = list(out)
                assert n_rg == -(-b // 4) and n_c8 == -(-(-(-v // 4)) // 8) and n_items == n_rg * n_c8
                assert span % warps == 0 and (span <= n_c8 or n_c8 < warps) and grid == -(-n_items // span)
                assert (grid - 1) * span < n_items <= grid * span
                if n_c8 < warps:
                    continue
                for c in range(0, grid, max(1, grid // 50)):          # a CTA's items span <= 2 row groups
                    first, last = c * span, min((c + 1) * span, n_items) - 1
                    assert last // n_c8 - first // n_c8 <= 1
                for rg in range(0, n_rg, max(1, n_rg // 50)):         # finalize: CTAs meeting row group rg, segment 0 / 1
                    c_lo, c_hi = (rg * n_c8) // span, (min((rg + 1) * n_c8, n_items) - 1) // span
                    assert 0 <= c_lo <= c_hi < grid
                    for c in (c_lo, c_hi):
                        assert rg - (c * span) // n_c8 in (0, 1)
                    # CTAs outside [c_lo, c_hi] hold no item of this row group
                    if c_lo > 0:
                        assert (c_lo * span - 1) // n_c8 < rg
                    if c_hi + 1 < grid:
                        assert ((c_hi + 1) * span) // n_c8 > rg
    assert lib.vg_recon_plan(4, 100, 3, out) != 0                     # fewer than 256 voxels is rejected


def test_fused_mlp_validation_without_gpu(built_lib):
    """vg_mlp_fwd / vg_mlp_bwd reject malformed chains before any device work."""
    n = built_lib
    lib = n.load()
    m = n.VgMlp()
    assert lib.vg_mlp_fwd(ctypes.byref(m), None) == n.VG_EINVAL                       # empty
    m.nlayers, m.nbufs, m.rows, m.rows_per_cta = 1, 2, 8, 3
    assert lib.vg_mlp_fwd(ctypes.byref(m), None) == n.VG_EINVAL                       # rows_per_cta not in {1,2,4,8}
    m.rows_per_cta = 4
    m.buf[0].width, m.buf[1].width = 300, 10
    m.buf[0].act = m.buf[1].act = 16                                                  # non-null placeholders, never dereferenced
    m.layer[0].w, m.layer[0].n, m.layer[0].k, m.layer[0].in_, m.layer[0].out = 16, 10, 300, 0, 1
    assert lib.vg_mlp_fwd(ctypes.byref(m), None) == n.VG_EINVAL                       # k > 224
    assert b"224" in lib.vg_last_error()


def test_conv_planner_dispatch_is_host_only_and_stable():
    """vg_conv_describe is host code (no device work): which kernel serves the decoder's layer passes in the tensor-core
    arithmetic, checked without a GPU.  Guards the planner's table sizes (convt2's forward needs 153 MMA entries for its
    8 parity phases) and the TMA / live-row-block geometry of the largest layers."""
    import ctypes as C

    from vaegam import native
    lib = native.load()

    def describe(kind, *a, **k):
        d = native.conv_desc(*a, **k)
        buf = C.create_string_buffer(8192)
        n = lib.vg_conv_describe(C.byref(d), kind, buf, len(buf))
        return n, buf.value.decode()

    bf = native.ARITH_BF16
    # convt2 (16 -> 16, stride 2, asymmetric padding): forward = ONE plane-folded launch over the 8 parity phases
    n, s = describe(0, True, 16, 16, (3, 3, 3), 2, (8, 10, 7), 288, 32, pad=(1, 0, 1), opad=(1, 0, 1), arith=bf)
    assert n == 1 and s.startswith("tc2 ") and "phases=8" in s, s
    # convt5's data gradient: 1 -> 8 channels, 1551 rows per plane in 512-row tiles (the last one holds 15 live rows)
    n, s = describe(1, True, 8, 1, (3, 3, 3), 1, (39, 47, 33), 288, 32, arith=bf, bf16_mask=native.BF16_X)
    assert n == 1 and "cin=1 cout=8" in s and "RTOT=1551" in s and "TR=512" in s and "ntiles=4" in s, s
    # convt5 / convt4 forward read bf16 activations: TMA-direct staging (tma_hb > 0)
    for args, kw in (((True, 8, 1, (3, 3, 3), 1, (39, 47, 33), 288, 32), dict(bf16_mask=native.BF16_X)),
                     ((True, 8, 8, (5, 3, 3), 2, (18, 22, 15), 288, 32), dict(bf16_mask=native.BF16_X | native.BF16_Y))):
        n, s = describe(0, *args, arith=bf, **kw)
        hb = int(s.split("tma_hb=")[1].split()[0])
        assert n == 1 and s.startswith("tc2 ") and hb > 0, s
    # fp32 arithmetic never reaches a tensor-core kernel
    n, s = describe(0, True, 8, 1, (3, 3, 3), 1, (39, 47, 33), 288, 32, arith=native.ARITH_FP32)
    assert "fp32" in s and "tc" not in s.replace("fp32", ""), s
