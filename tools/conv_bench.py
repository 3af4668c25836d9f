#!/usr/bin/env python
"""Times the convolution launches of chosen layers at production size through the C ABI (CUDA events),
or runs them once for use under ncu.   python tools/conv_bench.py [--layers convt5,convt4] [--ops fwd,dgrad,wgrad]
[--batch 32] [--iters 10] [--mode 1]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-gam_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from vaegam import native  # noqa: E402
from test_gpu_kernels import LAYERS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", default="convt5,convt4,convt3,conv1,conv2,conv3")
ap.add_argument("--ops", default="fwd,dgrad,wgrad")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--mode", type=int, default=1)
args = ap.parse_args()
lib = native.load()
lib.vg_set_conv_mode(args.mode)
dev = "cuda"
st = native.stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in args.layers.split(","):
    tr, cin, cout, k, s, in_, pad, opad = LAYERS[name]
    group = args.batch
    N = 9 * args.batch if name.startswith("convt") else args.batch
    d = native.conv_desc(tr, cin, cout, k, s, in_, N, group, pad, opad)
    out = tuple(d.out)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(N, *in_, cin, device=dev, generator=g)
    w = torch.randn(*((cin, cout, *k) if tr else (cout, cin, *k)), device=dev, generator=g) * 0.2
    b = torch.randn(cout, device=dev, generator=g)
    sc = torch.rand(N // group, cin, device=dev, generator=g) + 0.5
    sh = torch.randn(N // group, cin, device=dev, generator=g)
    y = torch.empty(N, *out, cout, device=dev)
    dy = torch.randn(N, *out, cout, device=dev, generator=g)
    dx = torch.empty_like(x)
    stats = torch.zeros(N // group, cout, 2, dtype=torch.float64, device=dev)
    sums = torch.zeros(N // group, cin, 2, dtype=torch.float64, device=dev)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    nbytes = 4.0 * (x.numel() + y.numel())
    fns = {
        "fwd": lambda: native.check(lib.vg_conv_fwd(C.byref(d), native.ptr(x), native.ptr(w), native.ptr(b), native.ptr(sc),
                                                    native.ptr(sh), native.ptr(y), native.ACT_RELU, native.ptr(stats), st)),
        "dgrad": lambda: native.check(lib.vg_conv_dgrad(C.byref(d), native.ptr(dy), native.ptr(w), native.ptr(dx), None,
                                                        native.ptr(x), native.ptr(sc), native.ptr(sh), native.ptr(sums), st)),
        "wgrad": lambda: native.check(lib.vg_conv_wgrad(C.byref(d), native.ptr(x), native.ptr(dy), native.ptr(sc),
                                                        native.ptr(sh), native.ptr(dw), native.ptr(db), st)),
    }
    for op in args.ops.split(","):
        fn = fns[op]
        fn()
        torch.cuda.synchronize()
        if args.iters <= 0:
            continue
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        med = ms[len(ms) // 2]
        extra = 4.0 * x.numel() if op == "dgrad" else 0.0      # the BN-backward epilogue also reads x
        print(f"{name}.{op}: {med*1e3:8.1f} us  {(nbytes + extra) / med / 1e6:8.1f} GB/s (fp32 in+out{'+aux' if extra else ''} bytes) N={N}")
print("ok")
