#!/usr/bin/env python
"""TEST / DESIGN TOOL (CPU, uses the oracle): which operand rounding drives the gradient deviation of the
tensor-core convolution mode?

The whole step is evaluated in fp64 by oracle/ref_port.py with its convolutions replaced by an autograd
function that rounds the operands of each pass (forward, data gradient, weight gradient) to a chosen format —
bf16 (8-bit mantissa), tf32 (11-bit), bf16x2 (hi + lo split, ~16-bit) or none — before an otherwise exact fp64
convolution.  Everything else (BatchNorm, linear layers, gains, loss) stays fp64, as the native path keeps
them fp32/fp64.  Output: per-parameter relative gradient error against the un-rounded fp64 step, and the
loss-term / map deviations, per rounding recipe.

    python tools/precision_study.py [B] [case]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-gam_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import ref_port as rp  # noqa: E402


def rnd(t, fmt):
    if fmt in (None, "none"):
        return t
    if fmt == "bf16":
        return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)
    if fmt == "bf16x2":
        f = t.to(torch.float32)
        hi = f.to(torch.bfloat16).to(torch.float32)
        lo = (f - hi).to(torch.bfloat16).to(torch.float32)
        return (hi + lo).to(t.dtype)
    if fmt == "tf32":      # round to nearest, 10 explicit mantissa bits
        f = t.to(torch.float32).contiguous()
        i = f.view(torch.int32)
        i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
        return i.view(torch.float32).to(t.dtype)
    if fmt == "fp32":
        return t.to(torch.float32).to(t.dtype)
    raise ValueError(fmt)


RECIPE = {"fwd": None, "dgrad": None, "wgrad": None, "store": None, "fwd_layers": None, "fwd_else": None}
CALL = [0]


def layer_name(idx):
    return f"conv{idx + 1}" if idx < 5 else f"convt{(idx - 5) % 5 + 1}"


class ConvEmu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, fn, kw):
        ctx.fn, ctx.kw = fn, kw
        ctx.save_for_backward(x, w)
        name = layer_name(CALL[0])
        CALL[0] += 1
        fmt = RECIPE["fwd"]
        if RECIPE["fwd_layers"] is not None and name not in RECIPE["fwd_layers"]:
            fmt = RECIPE["fwd_else"]
        y = fn(rnd(x, fmt), rnd(w, fmt), None, **kw) + b.view(1, -1, 1, 1, 1)
        return rnd(y, RECIPE["store"])

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        fn, kw = ctx.fn, ctx.kw
        dy = rnd(dy, RECIPE["store"])
        with torch.enable_grad():
            xd = x.detach().requires_grad_(True)
            y = fn(xd, rnd(w.detach(), RECIPE["dgrad"]), None, **kw)
            (dx,) = torch.autograd.grad(y, xd, rnd(dy, RECIPE["dgrad"]))
            wd = w.detach().requires_grad_(True)
            y2 = fn(rnd(x.detach(), RECIPE["wgrad"]), wd, None, **kw)
            (dw,) = torch.autograd.grad(y2, wd, rnd(dy, RECIPE["wgrad"]))
        db = dy.sum((0, 2, 3, 4))
        return dx, dw, db, None, None


def conv3d_emu(x, w, b, stride=1):
    return ConvEmu.apply(x, w, b, F.conv3d, {"stride": stride})


def convt3d_emu(x, w, b, stride=1, padding=0, output_padding=0):
    return ConvEmu.apply(x, w, b, F.conv_transpose3d, {"stride": stride, "padding": padding, "output_padding": output_padding})


class FEmu:
    """Stand-in for torch.nn.functional inside oracle.ref_port."""

    def __getattr__(self, name):
        return getattr(F, name)

    conv3d = staticmethod(conv3d_emu)
    conv_transpose3d = staticmethod(convt3d_emu)


def run(recipe, model, x, cov, noise, rc):
    RECIPE.update({"fwd": None, "dgrad": None, "wgrad": None, "store": None, "fwd_layers": None, "fwd_else": None})
    RECIPE.update(recipe)
    CALL[0] = 0
    P = {k: v.cpu() for k, v in rp.params_from_module(model).items()}
    Pd = rp.cast_params(P, torch.float64, requires_grad=True)
    out = rp.step(Pd, x.double(), cov.double(), noise, rc["gp_kl_scale"], rc["glm_reg_scale"], rc["neural"])
    out["tot"].backward()
    return out, Pd


def main():
    from helpers import build_case, load_golden
    case = sys.argv[2] if len(sys.argv) > 2 else "b4_m4_neural"
    g, rc = load_golden(case)
    if len(sys.argv) > 1:
        rc = dict(rc, B=int(sys.argv[1]))
    model, x, cov, ids = build_case(rc, device_name="cpu")
    noise = rp.draw_noise(rc["B"], seed=rc["noise_seed"])
    saved = rp.F
    rp.F = FEmu()
    try:
        ref, Pref = run({}, model, x, cov, noise, rc)
        recipes = [] if os.environ.get("STUDY") == "layers" else [
            ("fp32 everywhere", {"fwd": "fp32", "dgrad": "fp32", "wgrad": "fp32"}),
            ("bf16 all (round 1)", {"fwd": "bf16", "dgrad": "bf16", "wgrad": "bf16"}),
            ("bf16 fwd only", {"fwd": "bf16"}),
            ("bf16 dgrad only", {"dgrad": "bf16"}),
            ("bf16 wgrad only", {"wgrad": "bf16"}),
            ("bf16 fwd, tf32 dgrad+wgrad", {"fwd": "bf16", "dgrad": "tf32", "wgrad": "tf32"}),
            ("tf32 all", {"fwd": "tf32", "dgrad": "tf32", "wgrad": "tf32"}),
            ("bf16 fwd, bf16x2 dgrad+wgrad", {"fwd": "bf16", "dgrad": "bf16x2", "wgrad": "bf16x2"}),
            ("bf16x2 all", {"fwd": "bf16x2", "dgrad": "bf16x2", "wgrad": "bf16x2"}),
            ("bf16 all + bf16 storage", {"fwd": "bf16", "dgrad": "bf16", "wgrad": "bf16", "store": "bf16"}),
        ]
        if os.environ.get("STUDY") == "layers":
            enc = ["conv1", "conv2", "conv3", "conv4", "conv5"]
            dec = ["convt1", "convt2", "convt3", "convt4", "convt5"]
            bw = {"dgrad": "bf16", "wgrad": "bf16"}
            recipes = [("bf16 fwd encoder only", dict(bw, fwd="bf16", fwd_layers=set(enc))),
                       ("bf16 fwd decoder only", dict(bw, fwd="bf16", fwd_layers=set(dec)))]
            for l in enc + dec:
                recipes.append((f"bf16 fwd {l} only", dict(bw, fwd="bf16", fwd_layers={l})))
            recipes.append(("bf16 everywhere except tf32 conv1..3 fwd", dict(bw, fwd="tf32", fwd_layers={"conv1", "conv2", "conv3"}, fwd_else="bf16")))
            recipes.append(("bf16 everywhere except bf16x2 encoder fwd", dict(bw, fwd="bf16x2", fwd_layers=set(enc), fwd_else="bf16")))
            recipes.append(("bf16 everywhere except bf16x2 decoder fwd", dict(bw, fwd="bf16x2", fwd_layers=set(dec), fwd_else="bf16")))
        if os.environ.get("STUDY") == "mix":
            enc = {"conv1", "conv2", "conv3", "conv4", "conv5"}
            bw = {"dgrad": "bf16", "wgrad": "bf16"}
            recipes = [("bf16 all", dict(bw, fwd="bf16")),
                       ("bf16x2 encoder fwd, rest bf16", dict(bw, fwd="bf16x2", fwd_layers=enc, fwd_else="bf16")),
                       ("bf16x2 encoder fwd, tf32 decoder fwd, bf16 bwd", dict(bw, fwd="bf16x2", fwd_layers=enc, fwd_else="tf32")),
                       ("tf32 fwd, bf16 bwd", dict(bw, fwd="tf32")),
                       ("bf16x2 fwd, bf16 bwd", dict(bw, fwd="bf16x2")),
                       ("bf16x2 fwd, bf16 bwd, bf16 storage", dict(bw, fwd="bf16x2", store="bf16"))]
        names = [n for n, _ in model.named_parameters()]
        gmax = max(float(Pref[n].grad.norm()) for n in names)
        for title, rec in recipes:
            out, Pd = run(rec, model, x, cov, noise, rc)
            errs = {n: float((Pd[n].grad - Pref[n].grad).norm() / (Pref[n].grad.norm() + 1e-4 * gmax)) for n in names}
            worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
            terms = {k: abs(float(out[k]) - float(ref[k])) / abs(float(ref[k])) for k in ("tot", "neg_elbo", "glm_reg")}
            dm = (out["maps"] - ref["maps"]).abs()
            print(f"== {title}")
            print("   terms rel:", {k: f"{v:.1e}" for k, v in terms.items()}, f" maps mean {float(dm.mean()):.1e} max {float(dm.max()):.1e}")
            print("   worst grads:", ", ".join(f"{n} {e:.3f}" for n, e in worst))
            enc = max(e for n, e in errs.items() if n.startswith(("conv1", "conv2", "conv3", "conv4", "conv5", "bn1", "bn3", "bn5", "fc1", "fc2", "fc3", "fc4")))
            dec = max(e for n, e in errs.items() if n.startswith(("convt", "bnt", "fc5", "fc6", "fc7", "fc8")))
            print(f"   max encoder {enc:.4f}  max decoder {dec:.4f}  epsilon {errs['epsilon']:.2e}")
    finally:
        rp.F = saved


if __name__ == "__main__":
    main()
