"""Per-kernel parity through the C ABI (ctypes) on a real GPU.

Dense layers are compared with PyTorch's own fp32 operators (TF32 disabled) on the same
device — those ARE the reference's implementation of these layers.  The closed-
form stages (latent, gains, fused loss) are compared with the CPU oracle in fp64.
Tolerances: fp32 kernels 1e-5 relative (north star "fp32 check mode"); fp64 gain stage 1e-6.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu

LAYERS = {
    # name: (transposed, cin, cout, k, stride, in, pad, opad)  — vae_reg_GP.py:189-193, 211-215
    "conv1": (0, 1, 8, (3, 3, 3), 1, (41, 49, 35), (0, 0, 0), (0, 0, 0)),
    "conv2": (0, 8, 8, (3, 3, 3), 2, (39, 47, 33), (0, 0, 0), (0, 0, 0)),
    "conv3": (0, 8, 16, (3, 3, 3), 1, (19, 23, 16), (0, 0, 0), (0, 0, 0)),
    "conv4": (0, 16, 16, (3, 3, 3), 2, (17, 21, 14), (0, 0, 0), (0, 0, 0)),
    "conv5": (0, 16, 16, (3, 3, 3), 1, (8, 10, 6), (0, 0, 0), (0, 0, 0)),
    "convt1": (1, 16, 16, (3, 3, 3), 1, (6, 8, 5), (0, 0, 0), (0, 0, 0)),
    "convt2": (1, 16, 16, (3, 3, 3), 2, (8, 10, 7), (1, 0, 1), (1, 0, 1)),
    "convt3": (1, 16, 8, (3, 3, 3), 1, (16, 21, 14), (0, 0, 0), (0, 0, 0)),
    "convt4": (1, 8, 8, (5, 3, 3), 2, (18, 23, 16), (0, 0, 0), (0, 0, 0)),
    "convt5": (1, 8, 1, (3, 3, 3), 1, (39, 47, 33), (0, 0, 0), (0, 0, 0)),
    # small irregular shapes: ragged extents, every parity phase exercised
    "tiny_conv_s2": (0, 8, 16, (3, 3, 3), 2, (6, 7, 9), (0, 0, 0), (0, 0, 0)),
    "tiny_convt_s2": (1, 16, 8, (5, 3, 3), 2, (3, 4, 2), (1, 0, 1), (1, 0, 1)),
}


@pytest.fixture(scope="module")
def lib():
    from vaegam import native
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return native.load()


def nat():
    from vaegam import native
    return native


def to_cl(t):      # (N,C,D,H,W) -> (N,D,H,W,C) contiguous
    return t.permute(0, 2, 3, 4, 1).contiguous()


def from_cl(t):    # (N,D,H,W,C) -> (N,C,D,H,W)
    return t.permute(0, 4, 1, 2, 3).contiguous()


def torch_layer(spec, x, w, b):
    tr, cin, cout, k, s, in_, pad, opad = spec
    if tr:
        return F.conv_transpose3d(x, w, b, stride=s, padding=pad, output_padding=opad)
    return F.conv3d(x, w, b, stride=s)


@pytest.fixture(autouse=True)
def _fp32_mode_by_default(lib):
    """Tight fp32 parity runs on the CUDA-core kernels; tensor-core tests opt in explicitly."""
    lib.vg_set_conv_mode(0)
    yield
    lib.vg_set_conv_mode(0)
    lib.vg_set_conv_tuning(b"t2_min_voxels", 400000)


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def dispatch(lib, d, kind):
    """Which kernels serve this layer pass under d.arith: set of {"tc2", "tc1", "fp32"} (vg_conv_describe)."""
    buf = C.create_string_buffer(8192)
    n = lib.vg_conv_describe(C.byref(d), kind, buf, len(buf))
    lines = [l for l in buf.value.decode().strip().splitlines() if l]
    assert n >= 1 and lines
    return {l.split()[0] for l in lines}


def check_rounding(kinds, err_round, err_exact, tol, what):
    """A tensor-core launch must reproduce PyTorch fp32 fed the bf16-ROUNDED operands; an fp32 launch the exact
    ones.  Only a pass that mixes both kinds of launches (parity phases of a small stride-2 layer) may match either."""
    if kinds <= {"tc1", "tc2"}:
        assert err_round < tol, (what, "tensor-core kernel", kinds, err_round, err_exact)
    elif kinds == {"fp32"}:
        assert err_exact < tol, (what, "fp32 kernel", kinds, err_round, err_exact)
    else:
        assert min(err_round, err_exact) < tol, (what, kinds, err_round, err_exact)


# layer passes the plane-folded tcgen05 kernel serves in production
TC2_FWD = {"conv1", "convt3", "convt4", "convt5"}
TC2_DGRAD = {"conv1", "conv2", "convt3", "convt4", "convt5"}


@pytest.mark.parametrize("big", [True, False], ids=["plane_folded", "default_dispatch"])
@pytest.mark.parametrize("name", list(LAYERS))
def test_conv_tensor_core_path(lib, name, big):
    """tcgen05 implicit-GEMM kernels (mode 1); `big` forces the persistent plane-folded kernel onto the small
    test problems (production dispatch uses it from 4e5 output voxels per launch).  The kernel rounds the (BN-folded) input and the
    weights to bf16 and accumulates in fp32, so a PyTorch fp32 reference fed the SAME rounded
    operands must agree to fp32 accumulation error; layers the tensor-core kernel does not cover
    (1 input channel, strided gathers) fall back to the fp32 kernel and still have to pass."""
    native = nat()
    spec = LAYERS[name]
    tr, cin, cout, k, s, in_, pad, opad = spec
    N, group = 4, 2
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(7 + sum(map(ord, name)))
    x = torch.randn(N, cin, *in_, device=dev, generator=gen)
    wshape = (cin, cout, *k) if tr else (cout, cin, *k)
    w = torch.randn(*wshape, device=dev, generator=gen) * 0.2
    b = torch.randn(cout, device=dev, generator=gen)
    scale = torch.rand(N // group, cin, device=dev, generator=gen) + 0.5
    shift = torch.randn(N // group, cin, device=dev, generator=gen)
    d = native.conv_desc(tr, cin, cout, k, s, in_, N, group, pad, opad, arith=native.ARITH_BF16)   # per-call arithmetic
    assert lib.vg_get_conv_mode() == 0            # the process default stays fp32: the descriptor decides
    native.check(lib.vg_set_conv_tuning(b"t2_min_voxels", 0 if big else 400000))
    k_fwd, k_dgrad = dispatch(lib, d, 0), dispatch(lib, d, 1)
    if big and name in TC2_FWD:
        assert k_fwd == {"tc2"}, (name, k_fwd)
    if big and name in TC2_DGRAD:
        assert k_dgrad == {"tc2"}, (name, k_dgrad)
    d32 = native.conv_desc(tr, cin, cout, k, s, in_, N, group, pad, opad, arith=native.ARITH_FP32)
    assert dispatch(lib, d32, 0) == {"fp32"} and dispatch(lib, d32, 1) == {"fp32"}
    xa = torch.addcmul(shift.repeat_interleave(group, 0)[:, :, None, None, None], x,
                       scale.repeat_interleave(group, 0)[:, :, None, None, None])       # fma, as the kernel does
    x_cl = to_cl(x)
    y = torch.empty(N, *tuple(d.out), cout, device=dev)
    stats = torch.zeros(N // group, cout, 2, dtype=torch.float64, device=dev)
    st = native.stream_ptr()
    native.check(lib.vg_conv_fwd(C.byref(d), native.ptr(x_cl), native.ptr(w), native.ptr(b), native.ptr(scale),
                                 native.ptr(shift), native.ptr(y), native.ACT_RELU, native.ptr(stats), st))
    torch.cuda.synchronize()
    y_exact = torch.relu(torch_layer(spec, xa, w, b))
    y_round = torch.relu(torch_layer(spec, bf16r(xa), bf16r(w), b))
    err_round, err_exact = rel_err(from_cl(y).cpu(), y_round.cpu()), rel_err(from_cl(y).cpu(), y_exact.cpu())
    check_rounding(k_fwd, err_round, err_exact, 2e-5, "fwd")
    assert err_exact < 1e-2                                             # and bf16 stays within 1e-2 of fp32
    ref_stats = torch.stack([from_cl(y).permute(0, 1, 2, 3, 4).double().reshape(N // group, group, cout, -1).sum((1, 3)),
                             (from_cl(y).double() ** 2).reshape(N // group, group, cout, -1).sum((1, 3))], -1)
    assert rel_err(stats.cpu(), ref_stats.cpu()) < 1e-5
    # data gradient: dy and w rounded
    dy = torch.randn(N, cout, *tuple(d.out), device=dev, generator=gen)
    xg = xa.clone().requires_grad_(True)
    torch_layer(spec, xg, bf16r(w), b).backward(bf16r(dy))
    gx_round = xg.grad.clone()
    xg.grad = None
    torch_layer(spec, xg, w, b).backward(dy)
    gx_exact = xg.grad
    dy_cl = to_cl(dy)
    dx = torch.empty_like(x_cl)
    act = torch.randn(N, cin, *in_, device=dev, generator=gen)
    act_cl = to_cl(act)
    native.check(lib.vg_conv_dgrad(C.byref(d), native.ptr(dy_cl), native.ptr(w), native.ptr(dx), native.ptr(act_cl),
                                   None, None, None, None, st))
    torch.cuda.synchronize()
    mask = (act > 0)
    e_r, e_e = rel_err(from_cl(dx).cpu(), (gx_round * mask).cpu()), rel_err(from_cl(dx).cpu(), (gx_exact * mask).cpu())
    check_rounding(k_dgrad, e_r, e_e, 2e-5, "dgrad")
    assert e_e < 1e-2
    # BatchNorm-backward sums epilogue
    istd = torch.rand(N // group, cin, device=dev, generator=gen) + 0.5
    mistd = torch.randn(N // group, cin, device=dev, generator=gen)
    sums = torch.zeros(N // group, cin, 2, dtype=torch.float64, device=dev)
    dxb = torch.empty_like(x_cl)
    native.check(lib.vg_conv_dgrad(C.byref(d), native.ptr(dy_cl), native.ptr(w), native.ptr(dxb), None,
                                   native.ptr(act_cl), native.ptr(istd), native.ptr(mistd), native.ptr(sums), st))
    torch.cuda.synchronize()
    got = from_cl(dxb)
    xh = act * istd.repeat_interleave(group, 0)[:, :, None, None, None] - mistd.repeat_interleave(group, 0)[:, :, None, None, None]
    ref_sums = torch.stack([got.double().reshape(N // group, group, cin, -1).sum((1, 3)),
                            (got.double() * xh.double()).reshape(N // group, group, cin, -1).sum((1, 3))], -1)
    assert rel_err(sums.cpu(), ref_sums.cpu()) < 2e-5
    check_rounding(k_dgrad, rel_err(got.cpu(), gx_round.cpu()), rel_err(got.cpu(), gx_exact.cpu()), 2e-5, "dgrad+bn")
    # weight gradient (bf16 mma.sync kernel): folded input and dy rounded to bf16, fp32 accumulation
    wr = w.clone().requires_grad_(True)
    torch_layer(spec, bf16r(xa), wr, b).backward(bf16r(dy))
    gw_round = wr.grad.clone()
    wr.grad = None
    br = b.clone().requires_grad_(True)
    torch_layer(spec, xa, wr, br).backward(dy)
    gw_exact, gb_exact = wr.grad, br.grad
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    native.check(lib.vg_conv_wgrad(C.byref(d), native.ptr(x_cl), native.ptr(dy_cl), native.ptr(scale),
                                   native.ptr(shift), native.ptr(dw), native.ptr(db), st))
    torch.cuda.synchronize()
    e_r, e_e = rel_err(dw.cpu(), gw_round.cpu()), rel_err(dw.cpu(), gw_exact.cpu())
    assert min(e_r, e_e) < 5e-5, (e_r, e_e)
    assert e_e < 1e-2
    assert rel_err(db.cpu(), gb_exact.cpu()) < 2e-5


@pytest.mark.parametrize("name", ["convt5", "convt4", "conv2", "convt3"])
def test_conv_bf16_activation_storage(lib, name):
    """VgConvDesc.bf16_mask: the big activations may live in HBM as bf16 (tensor-core arithmetic).  A kernel fed
    bf16 x / dy must give exactly what it gives for the same values held in fp32 (they round to themselves), and a
    bf16 output is the fp32 output rounded once."""
    native = nat()
    spec = LAYERS[name]
    tr, cin, cout, k, s, in_, pad, opad = spec
    N, group = 4, 2
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(11 + sum(map(ord, name)))
    x = bf16r(torch.randn(N, cin, *in_, device=dev, generator=gen))
    wshape = (cin, cout, *k) if tr else (cout, cin, *k)
    w = torch.randn(*wshape, device=dev, generator=gen) * 0.2
    b = torch.randn(cout, device=dev, generator=gen)
    scale = torch.rand(N // group, cin, device=dev, generator=gen) + 0.5
    shift = torch.randn(N // group, cin, device=dev, generator=gen)
    affine = s == 1          # the model never folds a BatchNorm into a stride-2 layer (and the TMA path does not either)
    p_scale, p_shift = (native.ptr(scale), native.ptr(shift)) if affine else (None, None)
    native.check(lib.vg_set_conv_tuning(b"t2_min_voxels", 0))
    st = native.stream_ptr()
    mk = lambda mask: native.conv_desc(tr, cin, cout, k, s, in_, N, group, pad, opad, arith=native.ARITH_BF16, bf16_mask=mask)
    d0 = mk(0)
    x_cl = to_cl(x)
    x16 = x_cl.to(torch.bfloat16)
    y0 = torch.empty(N, *tuple(d0.out), cout, device=dev)
    native.check(lib.vg_conv_fwd(C.byref(d0), native.ptr(x_cl), native.ptr(w), native.ptr(b), p_scale,
                                 p_shift, native.ptr(y0), native.ACT_RELU, None, st))
    # forward: bf16 x (8 input channels), bf16 y (8 / 16 output channels)
    for mask in ([native.BF16_X] if cin == 8 else []) + ([native.BF16_Y] if cout % 8 == 0 else []) + \
            ([native.BF16_X | native.BF16_Y] if cin == 8 and cout % 8 == 0 else []):
        y = torch.empty(N, *tuple(d0.out), cout, device=dev, dtype=torch.bfloat16 if mask & native.BF16_Y else torch.float32)
        dm = mk(mask)
        native.check(lib.vg_conv_fwd(C.byref(dm), native.ptr(x16 if mask & native.BF16_X else x_cl), native.ptr(w), native.ptr(b),
                                     p_scale, p_shift, native.ptr(y), native.ACT_RELU, None, st))
        torch.cuda.synchronize()
        want = y0.to(torch.bfloat16).float() if mask & native.BF16_Y else y0
        buf = C.create_string_buffer(4096)
        lib.vg_conv_describe(C.byref(dm), 0, buf, len(buf))
        tma = "tma_hb=0" not in buf.value.decode() and "tma_hb=" in buf.value.decode()
        if tma and affine and (mask & native.BF16_X):
            # TMA-direct staging cannot touch the operand, so the BatchNorm fold moves: scale into the (bf16) weights,
            # shift into an exact fp32 bias over the taps that fall inside the input.  Same function, other roundings:
            ref = []
            for gi in range(N // group):
                xs = x[gi * group:(gi + 1) * group]
                wsc = w * (scale[gi].view(-1, 1, 1, 1, 1) if tr else scale[gi].view(1, -1, 1, 1, 1))
                shift_img = shift[gi].view(1, -1, 1, 1, 1).expand(group, cin, *in_).contiguous()
                ref.append(torch_layer(spec, xs, bf16r(wsc), None) + torch_layer(spec, shift_img, w, b))
            ref = to_cl(torch.relu(torch.cat(ref)))
            want = ref.to(torch.bfloat16).float() if mask & native.BF16_Y else ref
            assert rel_err(y.float().cpu(), want.cpu()) < (4e-3 if mask & native.BF16_Y else 2e-5), ("fwd tma", mask)
            assert rel_err(y.float().cpu(), y0.cpu()) < 1e-2, ("fwd tma vs producer-staged", mask)
        else:
            assert rel_err(y.float().cpu(), want.cpu()) < 1e-6, ("fwd", mask)
    # data gradient: bf16 dy (8 output channels), bf16 dx and bf16 saved activation (8 / 16 input channels)
    dy = bf16r(torch.randn(N, cout, *tuple(d0.out), device=dev, generator=gen))
    dy_cl = to_cl(dy)
    dy16 = dy_cl.to(torch.bfloat16)
    act = bf16r(torch.randn(N, cin, *in_, device=dev, generator=gen))
    act_cl = to_cl(act)
    act16 = act_cl.to(torch.bfloat16)
    istd = torch.rand(N // group, cin, device=dev, generator=gen) + 0.5
    mistd = torch.randn(N // group, cin, device=dev, generator=gen)
    dx0 = torch.empty_like(x_cl)
    sums0 = torch.zeros(N // group, cin, 2, dtype=torch.float64, device=dev)
    native.check(lib.vg_conv_dgrad(C.byref(d0), native.ptr(dy_cl), native.ptr(w), native.ptr(dx0), None, native.ptr(act_cl),
                                   native.ptr(istd), native.ptr(mistd), native.ptr(sums0), st))
    masks = []
    if cout == 8:
        masks.append(native.BF16_Y)
    if cin % 8 == 0:
        masks += [native.BF16_X, native.BF16_DX, native.BF16_X | native.BF16_DX]
        if cout == 8:
            masks.append(native.BF16_X | native.BF16_Y | native.BF16_DX)
    for mask in masks:
        dx = torch.empty_like(x_cl, dtype=torch.bfloat16 if mask & native.BF16_DX else torch.float32)
        sums = torch.zeros_like(sums0)
        dm = mk(mask)
        native.check(lib.vg_conv_dgrad(C.byref(dm), native.ptr(dy16 if mask & native.BF16_Y else dy_cl), native.ptr(w),
                                       native.ptr(dx), None, native.ptr(act16 if mask & native.BF16_X else act_cl),
                                       native.ptr(istd), native.ptr(mistd), native.ptr(sums), st))
        torch.cuda.synchronize()
        want = dx0.to(torch.bfloat16).float() if mask & native.BF16_DX else dx0
        assert rel_err(dx.float().cpu(), want.cpu()) < 1e-6, ("dgrad", mask)
        assert rel_err(sums.cpu(), sums0.cpu()) < 1e-6, ("dgrad sums", mask)
    # weight gradient: bf16 x / dy with 8 or 16 channels
    dw0, db0 = torch.zeros_like(w), torch.zeros_like(b)
    native.check(lib.vg_conv_wgrad(C.byref(d0), native.ptr(x_cl), native.ptr(dy_cl), p_scale, p_shift,
                                   native.ptr(dw0), native.ptr(db0), st))
    for mask in ([native.BF16_X] if cin % 8 == 0 else []) + ([native.BF16_Y] if cout % 8 == 0 else []):
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        dm = mk(mask)
        native.check(lib.vg_conv_wgrad(C.byref(dm), native.ptr(x16 if mask & native.BF16_X else x_cl),
                                       native.ptr(dy16 if mask & native.BF16_Y else dy_cl), p_scale, p_shift,
                                       native.ptr(dw), native.ptr(db), st))
        torch.cuda.synchronize()
        assert rel_err(dw.cpu(), dw0.cpu()) < 2e-5, ("wgrad", mask)       # fp32 atomics: accumulation order only
        assert rel_err(db.cpu(), db0.cpu()) < 2e-5, ("wgrad bias", mask)
        dw2 = torch.zeros_like(w)                                          # no bias gradient: bf16 dy is staged with cp.async
        native.check(lib.vg_conv_wgrad(C.byref(dm), native.ptr(x16 if mask & native.BF16_X else x_cl),
                                       native.ptr(dy16 if mask & native.BF16_Y else dy_cl), p_scale, p_shift,
                                       native.ptr(dw2), None, st))
        torch.cuda.synchronize()
        assert rel_err(dw2.cpu(), dw0.cpu()) < 2e-5, ("wgrad without bias", mask)


@pytest.mark.parametrize("in_", [(9, 11, 7), (39, 47, 33)], ids=["small", "bnt5_convt5"])
def test_fused_bn_backward_junction(lib, in_):
    """relu -> BatchNorm(batch statistics per group) -> ConvTranspose3d(8 -> 1, k3): the fused backward
    (vg_box_sums + vg_conv_wgrad_grouped + vg_bn_fused_finalize + vg_conv_dgrad_bn_apply, x stored as bf16) against
    PyTorch autograd of the same junction in fp32.  The tensor-core passes round dy and W to bf16: 1e-2."""
    native = nat()
    dev = "cuda"
    N, group, cin = 4, 2, 8
    G = N // group
    gen = torch.Generator(device=dev).manual_seed(5)
    x = bf16r(torch.relu(torch.randn(N, cin, *in_, device=dev, generator=gen) + 0.3))       # the stored activation
    w = torch.randn(cin, 1, 3, 3, 3, device=dev, generator=gen) * 0.2
    gamma = torch.rand(cin, device=dev, generator=gen) + 0.5
    beta = torch.randn(cin, device=dev, generator=gen) * 0.3
    out_dims = tuple(i + 2 for i in in_)
    dpre = torch.randn(N, 1, *out_dims, device=dev, generator=gen)
    # ---- reference: autograd
    xr = x.clone().requires_grad_(True)
    wr, gr, br = w.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bias = torch.zeros(1, device=dev, requires_grad=True)
    outs = []
    for gi in range(G):
        xn = F.batch_norm(xr[gi * group:(gi + 1) * group], None, None, gr, br, True, 0.0, 1e-5)
        outs.append(F.conv_transpose3d(xn, wr, bias))
    (torch.cat(outs) * dpre).sum().backward()
    du_ref = xr.grad * (x > 0)
    # ---- BatchNorm coefficients per (group, channel), as vg_bn_finalize gives them
    xg = x.reshape(G, group, cin, -1)
    mean = xg.mean((1, 3)); var = xg.var((1, 3), unbiased=False)
    istd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (gamma[None] * istd).contiguous(); shift = (beta[None] - mean * scale).contiguous()
    mistd = (mean * istd).contiguous(); istd = istd.contiguous()
    spatial = in_[0] * in_[1] * in_[2]
    count = float(group * spatial)
    # ---- fused path
    st = native.stream_ptr()
    d = native.conv_desc(1, cin, 1, (3, 3, 3), 1, in_, N, group, arith=native.ARITH_BF16,
                         bf16_mask=native.BF16_X | native.BF16_DX)
    x16 = to_cl(x).to(torch.bfloat16)
    dpre_cl = to_cl(dpre)
    box = torch.zeros(G, 28, dtype=torch.float64, device=dev)
    I3 = C.c_int32 * 3
    native.check(lib.vg_box_sums(native.ptr(dpre_cl), N, group, I3(*out_dims), 0, I3(*in_), I3(0, 0, 0), native.ptr(box), st))
    raw = torch.zeros(G, 27, 1, cin, device=dev)
    native.check(lib.vg_conv_wgrad_grouped(C.byref(d), native.ptr(x16), native.ptr(dpre_cl), native.ptr(raw), st))
    dw, dbias = torch.zeros_like(w), torch.zeros(1, device=dev)
    dgamma, dbeta = torch.zeros(cin, device=dev), torch.zeros(cin, device=dev)
    coef = torch.empty(G, cin, 3, device=dev)
    native.check(lib.vg_bn_fused_finalize(native.ptr(raw), native.ptr(box), native.ptr(w), native.ptr(scale), native.ptr(shift),
                                          native.ptr(istd), native.ptr(mistd), G, cin, count, native.ptr(dw), native.ptr(dbias),
                                          native.ptr(dgamma), native.ptr(dbeta), native.ptr(coef), st))
    du = torch.empty(N, *in_, cin, device=dev, dtype=torch.bfloat16)
    csum = torch.zeros(cin, device=dev)
    native.check(lib.vg_conv_dgrad_bn_apply(C.byref(d), native.ptr(dpre_cl), native.ptr(w), native.ptr(du), native.ptr(x16),
                                            native.ptr(coef), native.ptr(csum), st))
    torch.cuda.synchronize()
    # box sums are exact fp32 sums
    want_total = dpre.reshape(G, -1).double().sum(1)
    assert rel_err(box[:, 27].cpu(), want_total.cpu()) < 1e-5
    assert rel_err(box[:, 0].cpu(), dpre[:, 0, :in_[0], :in_[1], :in_[2]].reshape(G, -1).double().sum(1).cpu()) < 1e-5
    assert rel_err(dw.cpu(), wr.grad.cpu()) < 1e-2, rel_err(dw.cpu(), wr.grad.cpu())
    assert rel_err(dbias.cpu(), bias.grad.cpu()) < 1e-4
    assert rel_err(dgamma.cpu(), gr.grad.cpu()) < 1e-2 and rel_err(dbeta.cpu(), br.grad.cpu()) < 1e-2
    got = from_cl(du.float())
    assert rel_err(got.cpu(), du_ref.cpu()) < 1.5e-2, rel_err(got.cpu(), du_ref.cpu())
    assert rel_err(csum.cpu(), got.sum((0, 2, 3, 4)).cpu()) < 1e-2        # the stored values are rounded to bf16, the sum is not


@pytest.mark.parametrize("name", list(LAYERS))
def test_conv_forward_dgrad_wgrad(lib, name):
    native = nat()
    spec = LAYERS[name]
    tr, cin, cout, k, s, in_, pad, opad = spec
    N, group = 4, 2
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(sum(map(ord, name)))
    x = torch.randn(N, cin, *in_, device=dev, generator=gen)
    wshape = (cin, cout, *k) if tr else (cout, cin, *k)
    w = torch.randn(*wshape, device=dev, generator=gen) * 0.2
    b = torch.randn(cout, device=dev, generator=gen)
    scale = torch.rand(N // group, cin, device=dev, generator=gen) + 0.5
    shift = torch.randn(N // group, cin, device=dev, generator=gen)
    d = native.conv_desc(tr, cin, cout, k, s, in_, N, group, pad, opad)
    out_shape = tuple(d.out)

    # reference: affine per (group, channel), then the torch operator, ReLU
    xa = (x * scale.repeat_interleave(group, 0)[:, :, None, None, None]
          + shift.repeat_interleave(group, 0)[:, :, None, None, None]).requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pre = torch_layer(spec, xa, wr, br)
    y_ref = torch.relu(pre)

    x_cl = to_cl(x)
    y = torch.empty(N, *out_shape, cout, device=dev)
    stats = torch.zeros(N // group, cout, 2, dtype=torch.float64, device=dev)
    st = native.stream_ptr()
    native.check(lib.vg_conv_fwd(C.byref(d), native.ptr(x_cl), native.ptr(w), native.ptr(b), native.ptr(scale),
                                 native.ptr(shift), native.ptr(y), native.ACT_RELU, native.ptr(stats), st))
    torch.cuda.synchronize()
    assert rel_err(from_cl(y).cpu(), y_ref.detach().cpu()) < 1e-5
    ref_stats = torch.stack([y_ref.detach().double().reshape(N // group, group, cout, -1).sum((1, 3)),
                             (y_ref.detach().double() ** 2).reshape(N // group, group, cout, -1).sum((1, 3))], -1)
    assert rel_err(stats.cpu(), ref_stats.cpu()) < 1e-5

    # backward: dy w.r.t. the pre-activation
    dy = torch.randn_like(pre)
    pre.backward(dy)
    dy_cl = to_cl(dy)
    dx = torch.empty_like(x_cl)
    native.check(lib.vg_conv_dgrad(C.byref(d), native.ptr(dy_cl), native.ptr(w), native.ptr(dx), None, None, None,
                                   None, None, st))
    dw = torch.zeros_like(w)
    db = torch.zeros_like(b)
    native.check(lib.vg_conv_wgrad(C.byref(d), native.ptr(x_cl), native.ptr(dy_cl), native.ptr(scale),
                                   native.ptr(shift), native.ptr(dw), native.ptr(db), st))
    torch.cuda.synchronize()
    assert rel_err(from_cl(dx).cpu(), xa.grad.cpu()) < 1e-5          # gradient w.r.t. the affine-folded input
    assert rel_err(dw.cpu(), wr.grad.cpu()) < 2e-5
    assert rel_err(db.cpu(), br.grad.cpu()) < 2e-5

    # dgrad epilogues: ReLU mask and BatchNorm-backward sums
    act = torch.randn(N, cin, *in_, device=dev, generator=gen)
    act_cl = to_cl(act)
    dxm = torch.empty_like(x_cl)
    native.check(lib.vg_conv_dgrad(C.byref(d), native.ptr(dy_cl), native.ptr(w), native.ptr(dxm), native.ptr(act_cl),
                                   None, None, None, None, st))
    istd = torch.rand(N // group, cin, device=dev, generator=gen) + 0.5
    mistd = torch.randn(N // group, cin, device=dev, generator=gen)
    sums = torch.zeros(N // group, cin, 2, dtype=torch.float64, device=dev)
    dxb = torch.empty_like(x_cl)
    native.check(lib.vg_conv_dgrad(C.byref(d), native.ptr(dy_cl), native.ptr(w), native.ptr(dxb), None,
                                   native.ptr(act_cl), native.ptr(istd), native.ptr(mistd), native.ptr(sums), st))
    torch.cuda.synchronize()
    gx = xa.grad
    assert rel_err(from_cl(dxm).cpu(), (gx * (act > 0)).cpu()) < 1e-5
    assert rel_err(from_cl(dxb).cpu(), gx.cpu()) < 1e-5
    xh = act * istd.repeat_interleave(group, 0)[:, :, None, None, None] - mistd.repeat_interleave(group, 0)[:, :, None, None, None]
    ref_sums = torch.stack([gx.double().reshape(N // group, group, cin, -1).sum((1, 3)),
                            (gx.double() * xh.double()).reshape(N // group, group, cin, -1).sum((1, 3))], -1)
    assert rel_err(sums.cpu(), ref_sums.cpu()) < 2e-5


def test_conv_rejects_bad_descriptor(lib):
    native = nat()
    d = native.conv_desc(0, 8, 8, (3, 3, 3), 1, (5, 5, 5), 2, 1)
    d.out[0] = 7   # violates the size formula
    t = torch.zeros(8, device="cuda")
    rc = lib.vg_conv_fwd(C.byref(d), native.ptr(t), native.ptr(t), None, None, None, native.ptr(t), 0, None,
                         native.stream_ptr())
    assert rc == -1 and b"output size" in lib.vg_last_error()
    d2 = native.conv_desc(0, 3, 8, (3, 3, 3), 1, (5, 5, 5), 2, 1)   # unsupported channel count
    rc = lib.vg_conv_fwd(C.byref(d2), native.ptr(t), native.ptr(t), None, None, None, native.ptr(t), 0, None,
                         native.stream_ptr())
    assert rc == -1


@pytest.mark.parametrize("c,spatial", [(1, 70315), (8, 6992), (16, 240)])
def test_batchnorm_helpers(lib, c, spatial):
    """stats -> finalize -> folded affine == F.batch_norm(training=True); backward apply == autograd."""
    native = nat()
    dev = "cuda"
    N, group = 6, 3
    G = N // group
    gen = torch.Generator(device=dev).manual_seed(c)
    x = torch.relu(torch.randn(N, spatial, c, device=dev, generator=gen) + 0.3)
    gamma = torch.rand(c, device=dev, generator=gen) + 0.5
    beta = torch.randn(c, device=dev, generator=gen)
    stats = torch.zeros(G, c, 2, dtype=torch.float64, device=dev)
    st = native.stream_ptr()
    native.check(lib.vg_bn_stats(native.ptr(x), N, group, spatial, c, native.ptr(stats), st))
    coef = [torch.empty(G, c, device=dev) for _ in range(4)]
    count = float(group * spatial)
    native.check(lib.vg_bn_finalize(native.ptr(stats), native.ptr(gamma), native.ptr(beta), G, c, count,
                                    *[native.ptr(t) for t in coef], st))
    scale, shift, istd, mistd = coef
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ys = []
    for g in range(G):   # one F.batch_norm call per group, channels-first view
        xg = xr[g * group:(g + 1) * group].permute(0, 2, 1)
        ys.append(F.batch_norm(xg, None, None, gr, br, True, 0.0, 1e-5).permute(0, 2, 1))
    y_ref = torch.cat(ys)
    y = x.reshape(G, group, spatial, c) * scale[:, None, None, :] + shift[:, None, None, :]
    torch.cuda.synchronize()
    assert rel_err(y.reshape(N, spatial, c).cpu(), y_ref.detach().cpu()) < 1e-5
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    xh = x.reshape(G, group, spatial, c) * istd[:, None, None, :] - mistd[:, None, None, :]
    sums = torch.stack([dy.double().reshape(G, group, spatial, c).sum((1, 2)),
                        (dy.double().reshape(G, group, spatial, c) * xh.double()).sum((1, 2))], -1).contiguous()
    dx = torch.empty_like(x)
    dgamma, dbeta = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    native.check(lib.vg_bn_bwd_apply(native.ptr(dy), native.ptr(x), native.ptr(sums), native.ptr(scale),
                                     native.ptr(istd), native.ptr(mistd), N, group, spatial, c, count, 0, 0,
                                     native.ptr(dx), native.ptr(dgamma), native.ptr(dbeta), None, st))
    torch.cuda.synchronize()
    assert rel_err(dx.cpu(), xr.grad.cpu()) < 2e-5
    assert rel_err(dgamma.cpu(), gr.grad.cpu()) < 2e-5
    assert rel_err(dbeta.cpu(), br.grad.cpu()) < 2e-5
    # relu-masked variant
    dxm = torch.empty_like(x)
    native.check(lib.vg_bn_bwd_apply(native.ptr(dy), native.ptr(x), native.ptr(sums), native.ptr(scale),
                                     native.ptr(istd), native.ptr(mistd), N, group, spatial, c, count, 1, 0,
                                     native.ptr(dxm), None, None, None, st))
    torch.cuda.synchronize()
    assert rel_err(dxm.cpu(), (xr.grad * (x > 0)).cpu()) < 2e-5
    if c % 8 == 0:
        # bf16 storage of x and / or dx: same arithmetic on the bf16-rounded x, result rounded once on store
        xb = x.to(torch.bfloat16)
        ref16 = torch.empty_like(x)
        native.check(lib.vg_bn_bwd_apply(native.ptr(dy), native.ptr(xb.float().contiguous()), native.ptr(sums), native.ptr(scale),
                                         native.ptr(istd), native.ptr(mistd), N, group, spatial, c, count, 1, 0,
                                         native.ptr(ref16), None, None, None, st))
        for mask in (native.BF16_X, native.BF16_X | native.BF16_DX, native.BF16_DX):
            out = torch.empty_like(x, dtype=torch.bfloat16 if mask & native.BF16_DX else torch.float32)
            xin = xb if mask & native.BF16_X else xb.float().contiguous()
            csum = torch.zeros(c, device=dev)
            native.check(lib.vg_bn_bwd_apply(native.ptr(dy), native.ptr(xin), native.ptr(sums), native.ptr(scale),
                                             native.ptr(istd), native.ptr(mistd), N, group, spatial, c, count, 1, mask,
                                             native.ptr(out), None, None, native.ptr(csum), st))
            torch.cuda.synchronize()
            want = ref16.to(torch.bfloat16).float() if mask & native.BF16_DX else ref16
            assert rel_err(out.float().cpu(), want.cpu()) < 1e-6, mask
            assert rel_err(csum.cpu(), ref16.double().reshape(-1, c).sum(0).float().cpu()) < 1e-4, mask


def test_layout_transposes(lib):
    native = nat()
    x = torch.randn(5, 16, 240, device="cuda")
    y = torch.empty(5, 240, 16, device="cuda")
    native.check(lib.vg_nchw_to_nhwc(native.ptr(x), native.ptr(y), 5, 16, 240, native.stream_ptr()))
    z = torch.empty_like(x)
    native.check(lib.vg_nhwc_to_nchw(native.ptr(y), native.ptr(z), 5, 16, 240, native.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(y, x.permute(0, 2, 1).contiguous()) and torch.equal(z, x)


@pytest.mark.parametrize("m,n,k", [(2, 200, 3072), (32, 50, 100), (288, 3840, 200), (1, 32, 50), (36, 50, 41)])
def test_linear_forward_backward(lib, m, n, k):
    native = nat()
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(m + n + k)
    x = torch.randn(m, k, device=dev, generator=gen)
    w = torch.randn(n, k, device=dev, generator=gen) * 0.1
    b = torch.randn(n, device=dev, generator=gen)
    y = torch.empty(m, n, device=dev)
    st = native.stream_ptr()
    native.check(lib.vg_linear_fwd(native.ptr(x), native.ptr(w), native.ptr(b), native.ptr(y), m, n, k,
                                   native.ACT_RELU, st))
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    y_ref = torch.relu(F.linear(xr, wr, br))
    torch.cuda.synchronize()
    assert rel_err(y.cpu(), y_ref.detach().cpu()) < 1e-5
    dy = torch.randn_like(y)
    y_ref.backward(dy)
    dx = torch.empty_like(x)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    native.check(lib.vg_linear_bwd(native.ptr(dy), native.ptr(y), native.ptr(x), native.ptr(w), native.ptr(dx),
                                   native.ptr(dw), native.ptr(db), m, n, k, st))
    torch.cuda.synchronize()
    assert rel_err(dx.cpu(), xr.grad.cpu()) < 1e-5
    assert rel_err(dw.cpu(), wr.grad.cpu()) < 1e-5
    assert rel_err(db.cpu(), br.grad.cpu()) < 1e-5


def _mlp_case(native, chain, rows, dev, gen):
    """Builds a VgMlp for one of the two chains of the step plus its torch twin."""
    if chain == "enc_heads":      # fc2 -> fc31/32/33 -> fc41/42/43 (vae_reg_GP.py:245-251)
        widths = [200, 100, 50, 50, 50, 32, 32, 32]
        layers = [(100, 200, 0, 1, True), (50, 100, 1, 2, True), (50, 100, 1, 3, True), (50, 100, 1, 4, True),
                  (32, 50, 2, 5, False), (32, 50, 3, 6, False), (32, 50, 4, 7, False)]
        grad_in, rpc = [5, 6, 7], 4
    else:                         # fc5 -> fc6 -> fc7 (vae_reg_GP.py:255-257)
        widths = [41, 50, 100, 200]
        layers = [(50, 41, 0, 1, True), (100, 50, 1, 2, True), (200, 100, 2, 3, True)]
        grad_in, rpc = [3], 8
    acts = [torch.randn(rows, w, device=dev, generator=gen) if i == 0 else torch.full((rows, w), float("nan"), device=dev)
            for i, w in enumerate(widths)]
    grads = [torch.randn(rows, w, device=dev, generator=gen) if i in grad_in else torch.full((rows, w), float("nan"), device=dev)
             for i, w in enumerate(widths)]
    ws = [torch.randn(n, k, device=dev, generator=gen) * (1.5 / k ** 0.5) for n, k, *_ in layers]
    bs = [torch.randn(n, device=dev, generator=gen) * 0.3 for n, *_ in layers]
    dws = [torch.randn_like(w) for w in ws]          # gradients are ACCUMULATED on top of these
    dbs = [torch.randn_like(b) for b in bs]
    m = native.VgMlp()
    m.nlayers, m.nbufs, m.rows, m.rows_per_cta = len(layers), len(widths), rows, rpc
    for i, w in enumerate(widths):
        m.buf[i].act, m.buf[i].grad, m.buf[i].width = native.ptr(acts[i]), native.ptr(grads[i]), w
        m.buf[i].role = (native.MLP_INPUT | native.MLP_GRAD_OUT if i == 0 else 0) | (native.MLP_GRAD_IN if i in grad_in else 0)
    for l, (n, k, i, o, relu) in enumerate(layers):
        L = m.layer[l]
        L.w, L.b, L.dw, L.db = native.ptr(ws[l]), native.ptr(bs[l]), native.ptr(dws[l]), native.ptr(dbs[l])
        L.n, L.k, L.in_, L.out, L.act = n, k, i, o, native.ACT_RELU if relu else native.ACT_NONE
    return m, layers, grad_in, acts, grads, ws, bs, dws, dbs


@pytest.mark.parametrize("chain,rows,rpc", [("enc_heads", 1, 1), ("enc_heads", 5, 4), ("enc_heads", 32, 1), ("enc_heads", 33, 2),
                                            ("dec_stem", 9, 2), ("dec_stem", 45, 8), ("dec_stem", 288, 2), ("dec_stem", 288, 4)])
def test_fused_mlp_chains(lib, chain, rows, rpc):
    """vg_mlp_fwd / vg_mlp_bwd (the fused small fully-connected layers) vs torch fp32 autograd, including the
    three-way fan-in at h2, ragged last row blocks and gradient accumulation into dw / db."""
    native = nat()
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(rows)
    m, layers, grad_in, acts, grads, ws, bs, dws, dbs = _mlp_case(native, chain, rows, dev, gen)
    m.rows_per_cta = rpc
    dw0, db0 = [t.clone() for t in dws], [t.clone() for t in dbs]
    st = native.stream_ptr()
    native.check(lib.vg_mlp_fwd(C.byref(m), st))
    wr = [w.clone().requires_grad_(True) for w in ws]
    br = [b.clone().requires_grad_(True) for b in bs]
    x0 = acts[0].clone().requires_grad_(True)
    ref = {0: x0}
    for l, (n, k, i, o, relu) in enumerate(layers):
        y = F.linear(ref[i], wr[l], br[l])
        ref[o] = torch.relu(y) if relu else y
    torch.cuda.synchronize()
    for o in range(1, len(acts)):
        assert rel_err(acts[o].cpu(), ref[o].detach().cpu()) < 1e-5, o
    torch.autograd.backward([ref[o] for o in grad_in], [grads[o] for o in grad_in])
    native.check(lib.vg_mlp_bwd(C.byref(m), st))
    torch.cuda.synchronize()
    assert rel_err(grads[0].cpu(), x0.grad.cpu()) < 1e-5
    for l in range(len(layers)):
        assert rel_err((dws[l] - dw0[l]).cpu(), wr[l].grad.cpu()) < 2e-5, l
        assert rel_err((dbs[l] - db0[l]).cpu(), br[l].grad.cpu()) < 2e-5, l


def test_fused_mlp_rejects_bad_wiring(lib):
    native = nat()
    gen = torch.Generator(device="cuda").manual_seed(0)
    m, *_keep = _mlp_case(native, "dec_stem", 8, "cuda", gen)
    m.layer[1].k = 60                                   # does not match the width of its input buffer
    assert lib.vg_mlp_fwd(C.byref(m), native.stream_ptr()) == native.VG_EINVAL


@pytest.mark.parametrize("B,small_d", [(1, False), (4, False), (32, True), (70, False)])
def test_latent_sample_kl(lib, B, small_d):
    native = nat()
    from oracle import ref_port as rp
    dev = "cuda"
    gen = torch.Generator().manual_seed(B)
    heads = torch.randn(3, B, 32, generator=gen) * 0.5
    if small_d:
        heads[2, 0, 0] = -20.0     # d < 1e-6 somewhere -> every d gets + 1e-6 (vae_reg_GP.py:321-323)
    eps_w, eps_d = torch.randn(B, 1, generator=gen), torch.randn(B, 32, generator=gen)
    hd = heads.double().requires_grad_(True)
    z_ref, klz_ref, d_ref = rp.latent_sample_kl(hd[0], hd[1], torch.exp(hd[2]), eps_w.double(), eps_d.double())
    hg, ew, ed = heads.to(dev), eps_w.to(dev), eps_d.to(dev)
    z = torch.empty(B, 32, device=dev); klz = torch.empty(B, device=dev); dd = torch.empty(B, 32, device=dev)
    zcat = torch.empty(9, B, 41, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
    st = native.stream_ptr()
    native.check(lib.vg_latent_fwd(native.ptr(hg), native.ptr(ew), native.ptr(ed), B, native.ptr(z), native.ptr(klz),
                                   native.ptr(dd), native.ptr(zcat), native.ptr(flag), st))
    torch.cuda.synchronize()
    assert int(flag) == int(small_d)
    assert rel_err(z.cpu(), z_ref.detach()) < 1e-6
    assert rel_err(klz.cpu(), klz_ref.detach()) < 1e-5
    assert torch.equal(zcat[:, :, :32], z.expand(9, B, 32))
    assert torch.equal(zcat[:, :, 32:], torch.eye(9, device=dev)[:, None, :].expand(9, B, 9))
    dzcat = torch.randn(9, B, 41, generator=gen)
    dklz = torch.rand(B, generator=gen)
    (z_ref * dzcat[:, :, :32].double().sum(0)).sum().add((klz_ref * dklz.double()).sum()).backward()
    dheads = torch.empty(3, B, 32, device=dev)
    dzc, dkl = dzcat.to(dev), dklz.to(dev)          # keep the device copies alive across the call
    native.check(lib.vg_latent_bwd(native.ptr(hg), native.ptr(ew), native.ptr(ed), native.ptr(dd),
                                   native.ptr(dzc), native.ptr(dkl), B, native.ptr(dheads), st))
    torch.cuda.synchronize()
    assert rel_err(dheads.cpu(), hd.grad) < 2e-5


def _gain_setup(B, m, seed):
    from oracle import ref_port as rp
    gen = torch.Generator().manual_seed(seed)
    cov = torch.randn(B, 8, generator=gen) * 1.2
    eps = torch.randn(8, B, generator=gen)
    P = {}
    for i, key in enumerate(rp.GP_KEYS):
        P["sa_" + key] = torch.normal(1, 1, size=(1, 1), generator=gen)
        P["logstd_" + key] = torch.normal(0, 1, size=(1, 1), generator=gen) * 0.5
        if rp.has_gp(i + 1):
            P["qu_m_" + key] = torch.randn(1, m, generator=gen)
            S = torch.randn(m, m, generator=gen) * 0.3
            P["qu_S_" + key] = (2 * torch.eye(m) + S @ S.T).float()
            P["logkvar_" + key] = torch.tensor(0.1 * i)
            P["logls_" + key] = torch.tensor(-0.2 + 0.1 * i)
            P["xu_" + key] = torch.linspace(-3.3, 3.4, m)
    return cov, eps, P


@pytest.mark.parametrize("B,m,neural", [(2, 6, True), (32, 6, True), (32, 6, False), (7, 4, True), (128, 6, True)])
def test_gain_stage(lib, B, m, neural):
    """vg_gain_fwd/bwd vs the fp64 oracle with autograd (G1-G7)."""
    native = nat()
    from oracle import ref_port as rp
    dev = "cuda"
    cov, eps, P = _gain_setup(B, m, 100 + B + m)
    Pd = {k: v.double().clone().requires_grad_(not k.startswith("xu_")) for k, v in P.items()}
    g_ref, kl_ref, aux = rp.gains(Pd, cov.double(), eps.double(), neural)
    dg = torch.randn(8, B, generator=torch.Generator().manual_seed(5))
    kl_scale = 10.0
    ((g_ref * dg.double()).sum() + kl_scale * kl_ref).backward()

    Pg = {k: v.to(dev).contiguous() for k, v in P.items()}
    Gg = {k: torch.zeros_like(v) for k, v in Pg.items()}
    gp, gg = native.VgGainParams(), native.VgGainGrads()
    for i, key in enumerate(rp.GP_KEYS):
        gp.sa[i], gp.logstd[i] = native.ptr(Pg["sa_" + key]), native.ptr(Pg["logstd_" + key])
        gg.sa[i], gg.logstd[i] = native.ptr(Gg["sa_" + key]), native.ptr(Gg["logstd_" + key])
        gp.hrf[i] = int(neural and i == 0)
        if rp.has_gp(i + 1):
            gp.has_gp[i] = 1
            for f, pk in (("qu_m", "qu_m_"), ("qu_S", "qu_S_"), ("logkvar", "logkvar_"), ("logls", "logls_")):
                getattr(gp, f)[i] = native.ptr(Pg[pk + key])
                getattr(gg, f)[i] = native.ptr(Gg[pk + key])
            gp.xu[i] = native.ptr(Pg["xu_" + key])
    taps = rp.hrf_taps().to(dev)
    covg, epsg, dgg = cov.to(dev), eps.to(dev), dg.to(dev)
    g = torch.empty(8, B, device=dev); kl = torch.zeros(8, 2, dtype=torch.float64, device=dev)
    bm = torch.empty(8, B, device=dev); bv = torch.empty(8, B, device=dev)
    status = torch.zeros(8, dtype=torch.int32, device=dev)
    nbytes = int(lib.vg_gain_workspace_bytes(B, m))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = native.stream_ptr()
    native.check(lib.vg_gain_fwd(C.byref(gp), native.ptr(covg), native.ptr(epsg), native.ptr(taps), B, m, native.ptr(g),
                                 native.ptr(kl), native.ptr(bm), native.ptr(bv), native.ptr(status), native.ptr(ws),
                                 nbytes, st))
    native.check(lib.vg_gain_bwd(C.byref(gp), C.byref(gg), native.ptr(covg), native.ptr(epsg), native.ptr(taps),
                                 native.ptr(dgg), kl_scale, B, m, native.ptr(ws), nbytes, st))
    torch.cuda.synchronize()
    assert status.abs().sum().item() == 0
    assert np.abs(g.cpu().numpy() - g_ref.detach().numpy()).max() < 2e-6 * max(1.0, float(g_ref.abs().max()))
    assert abs(float(kl.sum()) - float(kl_ref)) < 1e-9 * max(1.0, abs(float(kl_ref)))
    assert np.abs(bm.cpu().numpy() - aux["mean"].detach().numpy()).max() < 2e-6 * max(1.0, float(aux["mean"].abs().max()))
    diag = torch.diagonal(aux["cov"], dim1=1, dim2=2).detach()
    assert rel_err(bv.cpu(), diag) < 1e-6
    for k, v in Gg.items():
        if k.startswith("xu_"):
            continue
        ref = Pd[k].grad
        if k.startswith("logkvar"):   # tiny by cancellation: absolute tolerance relative to the other GP grads
            assert abs(float(v.cpu()) - float(ref)) < 1e-5 * max(1.0, float(Pd["qu_m_" + k[8:]].grad.abs().max()))
        else:
            assert rel_err(v.cpu(), ref) < 5e-6, k


def test_gain_reports_non_pd(lib):
    native = nat()
    from oracle import ref_port as rp
    dev = "cuda"
    B, m = 4, 6
    cov, eps, P = _gain_setup(B, m, 3)
    P["qu_S_x"] = -torch.eye(m)            # not positive definite -> status, no NaN trap / crash
    Pg = {k: v.to(dev).contiguous() for k, v in P.items()}
    gp = native.VgGainParams()
    for i, key in enumerate(rp.GP_KEYS):
        gp.sa[i], gp.logstd[i] = native.ptr(Pg["sa_" + key]), native.ptr(Pg["logstd_" + key])
        if rp.has_gp(i + 1):
            gp.has_gp[i] = 1
            gp.qu_m[i], gp.qu_S[i] = native.ptr(Pg["qu_m_" + key]), native.ptr(Pg["qu_S_" + key])
            gp.logkvar[i], gp.logls[i], gp.xu[i] = native.ptr(Pg["logkvar_" + key]), native.ptr(Pg["logls_" + key]), native.ptr(Pg["xu_" + key])
    g = torch.empty(8, B, device=dev); kl = torch.zeros(8, 2, dtype=torch.float64, device=dev)
    status = torch.zeros(8, dtype=torch.int32, device=dev)
    nbytes = int(lib.vg_gain_workspace_bytes(B, m)); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    covg, epsg, tapsg = cov.to(dev), eps.to(dev), rp.hrf_taps().to(dev)
    native.check(lib.vg_gain_fwd(C.byref(gp), native.ptr(covg), native.ptr(epsg),
                                 native.ptr(tapsg), B, m, native.ptr(g), native.ptr(kl), None, None,
                                 native.ptr(status), native.ptr(ws), nbytes, native.stream_ptr()))
    torch.cuda.synchronize()
    assert status[1].item() != 0 and status[0].item() == 0


@pytest.mark.parametrize("B,cps", [(1, 3), (2, 3), (5, 0), (5, 1), (5, 2), (5, 3), (32, 0), (32, 1), (32, 2), (32, 3), (33, 3)])
def test_fused_recon_loss(lib, B, cps):
    """vg_recon_loss_fwd/bwd vs the fp64 oracle (R1-R4 and the fused-pass gradients); cps = the
    backward kernel's launch-shape variant (vg_recon_tune)."""
    native = nat()
    lib.vg_recon_tune(cps)
    from oracle import ref_port as rp
    dev = "cuda"
    V, VP = native.V, native.VP
    gen = torch.Generator().manual_seed(40 + B)
    maps = torch.rand(9, B, V, generator=gen) * 0.98 + 0.01
    pre = torch.log(maps / (1 - maps)).double().requires_grad_(True)          # pre-sigmoid
    g = (torch.randn(8, B, generator=gen)).double().requires_grad_(True)
    x = torch.rand(B, V, generator=gen)
    eps = (torch.randn(V, generator=gen) * 0.3 - 2.0).double().requires_grad_(True)
    glm = torch.rand(V, 8, generator=gen)
    lam = 0.7
    rl = rp.recon_loss(torch.sigmoid(pre), g, x.double(), eps, glm.double(), lam)
    tot = -rl["logp"].mean() + lam * rl["glm_reg"]
    tot.backward()

    mp = torch.zeros(9, B, VP, device=dev); mp[:, :, :V] = maps.to(dev)
    mp[:, :, V:] = float("nan")                                                 # padding must never leak
    ep = torch.zeros(VP, device=dev); ep[:V] = eps.detach().float().to(dev)
    gl = torch.zeros(8, VP, device=dev); gl[:, :V] = glm.t().to(dev)
    gd, xd = g.detach().float().to(dev).contiguous(), x.to(dev)
    logp = torch.empty(B, device=dev); norms = torch.empty(8, B, device=dev)
    cons = torch.empty(8, B, V, device=dev); xrec = torch.empty(B, V, device=dev)
    nbytes = int(lib.vg_recon_workspace_bytes(B, V)); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = native.stream_ptr()
    native.check(lib.vg_recon_loss_fwd(native.ptr(mp), native.ptr(gd), native.ptr(xd), native.ptr(ep), native.ptr(gl), B,
                                       V, native.ptr(logp), native.ptr(norms), native.ptr(cons), native.ptr(xrec),
                                       native.ptr(ws), nbytes, st))
    dpre = torch.zeros(9, B, VP, device=dev); dg = torch.empty(8, B, device=dev); deps = torch.empty(VP, device=dev)
    native.check(lib.vg_recon_loss_bwd(native.ptr(mp), native.ptr(gd), native.ptr(xd), native.ptr(ep), native.ptr(gl),
                                       native.ptr(norms), B, V, lam, native.ptr(dpre), native.ptr(dg), native.ptr(deps),
                                       native.ptr(ws), nbytes, st))
    torch.cuda.synchronize()
    lib.vg_recon_tune(-1)                                                       # back to the default
    assert rel_err(logp.cpu(), rl["logp"].detach()) < 2e-6
    assert rel_err(norms.cpu(), rl["glm_norms"].detach()) < 2e-6
    assert rel_err(cons.cpu(), rl["cons"].detach()) < 1e-6
    assert rel_err(xrec.cpu(), rl["x_rec"].detach()) < 1e-6
    assert rel_err(dpre[:, :, :V].cpu(), pre.grad) < 2e-5
    assert rel_err(dg.cpu(), g.grad) < 2e-5
    assert rel_err(deps[:V].cpu(), eps.grad) < 2e-5


def test_fused_adam_matches_torch(lib):
    native = nat()
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(0)
    p32 = torch.randn(10007, device=dev, generator=gen); p64 = torch.randn(501, device=dev, dtype=torch.float64, generator=gen)
    r32, r64 = p32.clone().requires_grad_(True), p64.clone().requires_grad_(True)
    opt = torch.optim.Adam([r32, r64], lr=1e-3)
    m32, v32 = torch.zeros_like(p32), torch.zeros_like(p32)
    m64, v64 = torch.zeros_like(p64), torch.zeros_like(p64)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for it in range(5):
        g32 = torch.randn(10007, device=dev, generator=gen); g64 = torch.randn(501, device=dev, dtype=torch.float64, generator=gen)
        r32.grad, r64.grad = g32.clone(), g64.clone()
        opt.step()
        native.check(lib.vg_adam_step(native.ptr(p32), native.ptr(g32), native.ptr(m32), native.ptr(v32), p32.numel(),
                                      native.ptr(p64), native.ptr(g64), native.ptr(m64), native.ptr(v64), p64.numel(),
                                      1e-3, 0.9, 0.999, 1e-8, 1.0, native.ptr(step), None, 0, native.stream_ptr()))
    torch.cuda.synchronize()
    assert int(step) == 5
    assert rel_err(p32.cpu(), r32.detach().cpu()) < 1e-6
    assert rel_err(p64.cpu(), r64.detach().cpu()) < 1e-13
    # skip flags: any non-zero flag freezes parameters, moments and the step count (non-PD minibatch)
    flags = torch.zeros(8, dtype=torch.int32, device=dev)
    keep = (p32.clone(), m32.clone(), v32.clone(), p64.clone())
    for bad in (True, False):
        flags[5] = 3 if bad else 0
        native.check(lib.vg_adam_step(native.ptr(p32), native.ptr(g32), native.ptr(m32), native.ptr(v32), p32.numel(),
                                      native.ptr(p64), native.ptr(g64), native.ptr(m64), native.ptr(v64), p64.numel(),
                                      1e-3, 0.9, 0.999, 1e-8, 1.0, native.ptr(step), native.ptr(flags), 8, native.stream_ptr()))
        torch.cuda.synchronize()
        frozen = all(torch.equal(a, b) for a, b in zip(keep, (p32, m32, v32, p64)))
        assert frozen == bad and int(step) == (5 if bad else 6)


def test_gp_posterior_matches_oracle(lib):
    import gp as gpmod
    from oracle import ref_port as rp
    dev = "cuda"
    m, nq = 6, 300
    gen = torch.Generator().manual_seed(9)
    xu = torch.linspace(-3, 3, m); qm = torch.randn(1, m, generator=gen)
    S = torch.randn(m, m, generator=gen) * 0.3; qs = 2 * torch.eye(m) + S @ S.T
    kv, ls = torch.tensor(1.1), torch.tensor(2.3)
    xq = torch.randn(nq, generator=gen) * 1.5
    f_ref, s_ref = rp.gp_posterior(xu.double(), kv.double(), ls.double(), qm.double().reshape(-1), qs.double(), xq.double())
    reg = gpmod.GP(xu.to(dev), kv.to(dev), ls.to(dev), qm.to(dev), qs.to(dev))
    f, s = reg.evaluate_posterior(xq.to(dev))
    f2, var = reg.evaluate_posterior_diag(xq.to(dev))
    assert rel_err(f.cpu(), f_ref) < 1e-6 and rel_err(s.cpu(), s_ref) < 1e-6
    assert rel_err(var.cpu(), torch.diagonal(s_ref)) < 1e-6 and torch.equal(f, f2)
    kl = reg.compute_GP_kl(m)
    assert abs(float(kl) - float(rp.gp_kl(qm.double().reshape(-1), qs.double()))) < 1e-5
