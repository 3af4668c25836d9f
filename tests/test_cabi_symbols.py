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
