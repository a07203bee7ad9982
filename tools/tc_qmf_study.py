#!/usr/bin/env python3
"""Tensor cores for the QMF cosine modulation?  (BASELINE.json north_star: "only if ncu shows the QMF cosine-modulation
batched GEMM beats the FP32 pipe within tolerance".)

The 64-band synthesis bank turns every time slot's 64 complex sub-band samples into 128 real values
    v[n] = sum_k Re{ X[k] / 64 * exp(j pi/128 (k + 1/2)(2n - 255)) },   n = 0..127
(ISO/IEC 14496-3 4.6.18.4.2; JAAD evaluates it with two 64-point DCT-IV kernels, sbr/SynthesisFilterbank64.java:44-77).  As
a GEMM that is [slots x 128] x [128 x 128] per batch -- 32.8 kflop per slot instead of the ~2.6 kflop of the fast transform.
This script measures what the tensor-core route could give at best, with cuBLAS as the stand-in for a hand-written
tcgen05 pipeline (at M = millions, N = K = 128 a library GEMM runs at the memory roofline, which no fused kernel of the same
traffic can beat): time and error of fp32 (CUDA cores), TF32, 3xTF32 (hi/lo split, three products) and 3xBF16 against a
float64 reference, for one tile of BASELINE config 3 (4096 streams x 2 channels x 12 frames x 32 slots).
One JSON line on stdout.
"""
import json

import numpy as np
import torch


def main():
    torch.manual_seed(0)
    dev = "cuda"
    slots = 4096 * 2 * 12 * 32
    n = torch.arange(128, dtype=torch.float64)
    k = torch.arange(64, dtype=torch.float64)
    ang = np.pi / 128.0 * (k[:, None] + 0.5) * (2 * n[None, :] - 255.0)
    # real GEMM: [Re X | Im X] (128) x [[cos], [-sin]] / 64
    W64 = torch.cat([torch.cos(ang), -torch.sin(ang)], 0) / 64.0            # [128, 128] float64
    W = W64.to(dev)
    # sub-band samples in JAAD's +-32768 domain: QMF analysis of full-scale audio gives |X| up to ~ 32768 * 32
    X = (torch.randn(slots, 128, device=dev, dtype=torch.float32) * 8192.0)
    ref = (X[: 1 << 16].double() @ W)                                        # float64 reference on a sample of the slots
    full_scale = 32768.0
    out = {"slots": slots, "flop_gemm_per_slot": 2 * 128 * 128, "flop_dct4_per_slot": 2 * 1300, "rows": {}}

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            y = fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps, y

    W32 = W.float()

    def split(a, dt):
        hi = a.to(dt).float()
        return hi, a - hi

    def run(name, fn):
        ms, y = timeit(fn)
        err = float((y[: 1 << 16].double() - ref).abs().max())
        out["rows"][name] = {"ms": ms, "max_abs_err": err, "err_over_full_scale": err / full_scale, "within_1e-5_fs": err / full_scale <= 1e-5,
                             "GB_s": 2 * slots * 128 * 4 / ms / 1e6}

    torch.backends.cuda.matmul.allow_tf32 = False
    run("fp32_cuda_cores", lambda: X @ W32)
    torch.backends.cuda.matmul.allow_tf32 = True
    run("tf32", lambda: X @ W32)

    def tf32x3():
        # a = a_hi + a_lo with a_hi exactly representable in TF32 (10 mantissa bits): three tensor-core products
        xh = (X.view(torch.int32) & -8192).view(torch.float32)
        xl = X - xh
        wh = (W32.view(torch.int32) & -8192).view(torch.float32)
        wl = W32 - wh
        return xh @ wh + (xh @ wl + xl @ wh)
    run("tf32_x3_split", tf32x3)
    torch.backends.cuda.matmul.allow_tf32 = False

    def bf16x3():
        xh, xl = split(X, torch.bfloat16)
        wh, wl = split(W32, torch.bfloat16)
        f = lambda a, b: (a.bfloat16() @ b.bfloat16()).float()
        return f(xh, wh) + (f(xh, wl) + f(xl, wh))
    run("bf16_x3_split", bf16x3)
    # the copy both routes cannot avoid if the modulation is its own pass: read X, write v
    ms, _ = timeit(lambda: X.clone())
    out["copy_same_traffic_ms"] = ms
    print(json.dumps(out))


if __name__ == "__main__":
    main()
