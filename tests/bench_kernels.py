"""Micro-benchmarks of single kernels (CUDA events, device-resident operands).  Not a test; run on the GPU box:
    python tests/bench_kernels.py rec        # recurrence kernel: time per step vs episodes in flight
    python tests/bench_kernels.py gemm       # 3xTF32 tcgen05 GEMM vs the SIMT fp32 GEMM
    python tests/bench_kernels.py rec_one B T   # a single launch (for ncu)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodaltopicsegmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def rec_setup(B, T, H=256, save=False, n_enc=1):
    g = torch.Generator(device=dev).manual_seed(0)
    gx = torch.randn((n_enc, B * T, 8 * H), device=dev, generator=g) * 0.5
    whh = torch.randn((n_enc, 2, 4 * H, H), device=dev, generator=g) * 0.05
    lens = ops.Lengths([T] * B, dev, T)
    y = torch.empty((B, T, n_enc * 2 * H), device=dev)
    gates = torch.empty((n_enc, 2, B, T, 5, H), device=dev) if save else None
    return gx, whh, lens, y, gates


REC_NAME = "mts_lstm_rec_fwd" if os.environ.get("MTS_REC_IMPL", "tc") == "fma" else "mts_lstm_rec_fwd_tc"


def rec_call(gx, whh, lens, y, gates, B, T, H=256, n_enc=1):
    ops._call(REC_NAME, gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), n_enc, B, T,
              H, y.data_ptr(), 0 if gates is None else gates.data_ptr(), *((0,) if REC_NAME.endswith("_tc") else ()), ops._stream())


def bench_rec():
    print("B     T    ms/launch   us/step   sentences/s(layer)   algorithmic GB/s")
    for B in (8, 16, 32, 64, 72, 128, 256, 1024):
        for T in (300,):
            args = rec_setup(B, T)
            ms = timeit(lambda: rec_call(*args, B, T))
            print(f"{B:5d} {T:5d} {ms:9.3f} {ms * 1e3 / T:9.2f} {B * T / ms * 1e3:14.3e} {B * T * 10240 / ms / 1e6:12.1f}")
    args = rec_setup(64, 300, save=True)
    ms = timeit(lambda: rec_call(*args, 64, 300))
    print(f"training forward (gates saved) B=64 T=300: {ms:.3f} ms")
    # backward
    B, T, H = 64, 300, 256
    gx, whh, lens, y, gates = rec_setup(B, T, save=True)
    rec_call(gx, whh, lens, y, gates, B, T)
    dy = torch.randn_like(y)
    dgx = torch.empty_like(gx)
    whh_t = whh.transpose(2, 3).contiguous()
    ms = timeit(lambda: ops._call("mts_lstm_rec_bwd", dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), whh_t.data_ptr(),
                                  lens.dev.data_ptr(), lens.order.data_ptr(), 1, B, T, H, dgx.data_ptr(), ops._stream()))
    print(f"backward, packed-FMA kernel (mts_lstm_rec_bwd) B=64 T=300: {ms:.3f} ms  ({ms * 1e3 / T:.2f} us/step)")
    for Bb in (8, 10, 16, 64, 128):
        gx, whh, lens, y, gates = rec_setup(Bb, T, save=True)
        rec_call(gx, whh, lens, y, gates, Bb, T)
        dy = torch.randn_like(y)
        dgx = torch.empty_like(gx)
        for name in ("mts_lstm_rec_bwd_h3", "mts_lstm_rec_bwd_tf32"):
            ms = timeit(lambda: ops._call(name, dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(),
                                          lens.order.data_ptr(), 1, Bb, T, H, dgx.data_ptr(), ops._stream()))
            print(f"backward, tensor-core kernel ({name}) B={Bb} T=300: {ms:.3f} ms  ({ms * 1e3 / T:.2f} us/step)")


def bench_gemm():
    print("M      N     K     tf32x3 ms  TFLOP/s(fp32-equiv)   simt ms  TFLOP/s")
    for M, N, K in ((19200, 2048, 896), (19200, 2048, 512), (65536, 2048, 896), (245760, 2688, 896), (4096, 256, 256)):
        a = torch.randn(M, K, device=dev)
        b = torch.randn(N, K, device=dev)
        bias = torch.randn(N, device=dev)
        c = torch.empty(M, N, device=dev)
        a_hi, a_lo = ops.split_tf32(a)
        b_hi, b_lo = ops.split_tf32(b, side=ops.B_SIDE)
        ms = timeit(lambda: ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, c, M, N, epilogue=1))
        ms2 = timeit(lambda: ops.gemm_f32(a.data_ptr(), K, b.data_ptr(), K, bias, c.data_ptr(), N, M, N, K, layout=0,
                                          epilogue=1), iters=3, warmup=1)
        fl = 2.0 * M * N * K
        print(f"{M:6d} {N:5d} {K:5d} {ms:9.3f} {fl / ms / 1e9:10.1f} {ms2:16.3f} {fl / ms2 / 1e9:8.1f}")
        ms3 = timeit(lambda: ops.split_tf32(a))
        print(f"       split_tf32 of A: {ms3:.3f} ms ({M * K * 12 / ms3 / 1e6:.0f} GB/s)")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "rec"
    if what == "rec":
        bench_rec()
    elif what == "gemm":
        bench_gemm()
    elif what == "gemm_big":
        M, N, K = 65536, 2048, 896
        a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev); bias = torch.randn(N, device=dev)
        c = torch.empty(M, N, device=dev)
        a_hi, a_lo = ops.split_tf32(a); b_hi, b_lo = ops.split_tf32(b, side=ops.B_SIDE)
        ms = timeit(lambda: ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, c, M, N, epilogue=1))
        print(f"{M} x {N} x {K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s fp32-equivalent")
    elif what == "xf_gemms":  # the four dense layers of a configs[2] encoder layer on its 135 428 valid tokens
        M = 135428
        tot = 0.0
        for name, N, K in (("qkv", 2688, 896), ("out-proj", 896, 896), ("ffn1", 256, 896), ("ffn2", 896, 256)):
            a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev) * 0.05; bias = torch.randn(N, device=dev)
            c = torch.empty(M, N, device=dev)
            a_hi, a_lo = ops.split_tf32(a); b_hi, b_lo = ops.split_tf32(b, side=ops.B_SIDE)
            ms = timeit(lambda: ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, c, M, N, epilogue=1))
            rows = torch.randint(0, M, (512,), device=dev)
            rows[0], rows[1] = M - 1, 0
            ref = a[rows].double() @ b.double().t() + bias.double()
            err = float((c[rows].double() - ref).abs().max() / ref.abs().max())
            tot += ms
            print(f"{name:9s} {M} x {N} x {K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s fp32-equivalent  norm-wise err {err:.2e}")
        print(f"sum per layer {tot:.3f} ms, x 6 layers = {6 * tot:.2f} ms")
    elif what == "gemm_f16x3":  # timing + accuracy probe of the fp16-split product against the TF32 + bf16 one
        for M, N, K in ((19200, 2048, 896), (19200, 2048, 512), (135428, 2688, 896), (135428, 896, 896), (135428, 256, 896), (135428, 896, 256)):
            a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev) * 0.05; bias = torch.randn(N, device=dev)
            c = torch.empty(M, N, device=dev); c2 = torch.empty(M, N, device=dev)
            a_hi, a_lo = ops.split_tf32(a); b_hi, b_lo = ops.split_tf32(b, side=ops.B_SIDE)
            ms0 = timeit(lambda: ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, c, M, N, epilogue=1))
            (ap, asc), (bp, bsc) = ops.f16_pieces(a), ops.f16_pieces(b)
            call = lambda: ops.gemm_f16x3(ap, asc, bp, bsc, bias, c2, M, N, epilogue=1)
            ms1 = timeit(call)
            rows = torch.randint(0, M, (512,), device=dev)
            ref = a[rows].double() @ b.double().t() + bias.double()
            e0 = float((c[rows].double() - ref).abs().max() / ref.abs().max())
            e1 = float((c2[rows].double() - ref).abs().max() / ref.abs().max())
            print(f"{M} x {N} x {K}: tf32+bf16 {ms0:.3f} ms ({2.0 * M * N * K / ms0 / 1e9:.0f} TF, err {e0:.1e})   "
                  f"fp16x3 {ms1:.3f} ms ({2.0 * M * N * K / ms1 / 1e9:.0f} TF, err {e1:.1e})", flush=True)
    elif what == "gemm_one":
        M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (19200, 2048, 896)
        a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev); bias = torch.randn(N, device=dev)
        c = torch.empty(M, N, device=dev)
        a_hi, a_lo = ops.split_tf32(a); b_hi, b_lo = ops.split_tf32(b, side=ops.B_SIDE)
        for _ in range(3):
            ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, c, M, N, epilogue=1)
        torch.cuda.synchronize()
    elif what == "attn_one":  # cfg3-shaped banded attention, one layer (reach w), for ncu
        w = int(sys.argv[2]) if len(sys.argv) > 2 else 48
        B, S, h, hd = 256, 960, 8, 112
        d = h * hd
        g = torch.Generator(device=dev).manual_seed(0)
        qkv = torch.randn(B * S, 3 * d, device=dev, generator=g)
        lengths = torch.randint(100, S + 1, (B,), generator=torch.Generator().manual_seed(7))
        lengths[0] = S
        lens = ops.Lengths(lengths, dev, S)
        hl = torch.empty(2, B * S, d, device=dev)
        for _ in range(3):
            ops._call("mts_band_attn_fwd", qkv.data_ptr(), 3 * d, lens.dev.data_ptr(), 0, B, S, h, hd, w, 0, hl[0].data_ptr(),
                      hl[1].data_ptr(), d, 0, ops._stream())
        torch.cuda.synchronize()
        ms = timeit(lambda: ops._call("mts_band_attn_fwd", qkv.data_ptr(), 3 * d, lens.dev.data_ptr(), 0, B, S, h, hd, w, 0,
                                      hl[0].data_ptr(), hl[1].data_ptr(), d, 0, ops._stream()), iters=3, warmup=0)
        n = int(lengths.sum())
        print(f"band_attn_fwd w={w}: {ms:.3f} ms, {n * 16 * d / ms / 1e6:.1f} GB/s algorithmic ({n} valid tokens)")
    elif what == "rec_one":
        B, T = int(sys.argv[2]), int(sys.argv[3])
        args = rec_setup(B, T)
        for _ in range(3):
            rec_call(*args, B, T)
        torch.cuda.synchronize()
