"""First-contact check of the tensor-core recurrence (run on the GPU box, under `timeout`):
  1. what kind::tf32 keeps of an fp32 operand (truncation vs rounding) -- the kernel's hi/lo split relies on it;
  2. lstm_fwd_tc_kernel against the exact-fp32 packed-FMA cluster kernel and a float64 torch recurrence;
  3. time per step of both kernels.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodaltopicsegmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
# which tensor-core forward kernel the checks exercise: "mts_lstm_rec_fwd_tc" (TF32 + bf16 correction) or
# "mts_lstm_rec_fwd_h3" (fp16-split operands)
TC_NAME = os.environ.get("REC_TC_NAME", "mts_lstm_rec_fwd_h3")
PROF_NAME = ("mts_debug_rec_profile_h3p" if TC_NAME.endswith("_h3p") else "mts_debug_rec_profile_h3" if TC_NAME.endswith("_h3")
             else "mts_debug_rec_profile")


def _extra(name):
    """trailing arguments between `gates` and the stream"""
    if name.endswith(("_h3", "_h3p")):
        return (0, int(os.environ.get("REC_H3_PRECISION", "0")))          # y_corr, precision (+ debug bits)
    if name.endswith("_tc"):
        return (0,)            # y_corr
    return ()


def trunc_check():
    g = torch.Generator().manual_seed(0)
    M, N, K = 256, 128, 64
    a = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, K, generator=g).to(dev)
    b_hi = (b.view(torch.int32) & ~0x1FFF).view(torch.float32).contiguous()   # TF32-representable B
    zeros_a, zeros_b = torch.zeros_like(a), torch.zeros_like(b_hi)
    c = torch.empty(M, N, device=dev)
    ops.gemm_tf32x3(a.contiguous(), zeros_a, b_hi, zeros_b, None, c, M, N)     # zero correction operands: hi x hi only
    a_tr = (a.view(torch.int32) & ~0x1FFF).view(torch.float32)
    a_rn = ((a.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)      # round-to-nearest TF32 (ties away)
    ref_tr = a_tr.double() @ b_hi.double().T
    ref_rn = a_rn.double() @ b_hi.double().T
    e_tr = float((c.double() - ref_tr).abs().max())
    e_rn = float((c.double() - ref_rn).abs().max())
    print(f"tf32 operand read: |c - trunc| = {e_tr:.3e}   |c - round| = {e_rn:.3e}  ->",
          "TRUNCATES" if e_tr < 0.2 * e_rn else "does NOT simply truncate")
    return e_tr < 0.2 * e_rn


def ref_recurrence(gx, whh, lengths, B, T, H, n_enc):
    gx, whh = gx.double().cpu(), whh.double().cpu()
    y = torch.zeros(B, T, n_enc * 2 * H, dtype=torch.float64)
    for e in range(n_enc):
        for d in range(2):
            W = whh[e, d]
            for b in range(B):
                n = lengths[b]
                h = torch.zeros(H, dtype=torch.float64)
                c = torch.zeros(H, dtype=torch.float64)
                for s in range(n):
                    t = n - 1 - s if d else s
                    pre = gx[e, b * T + t, d * 4 * H:(d + 1) * 4 * H] + W @ h
                    i, f, g, o = pre[:H].sigmoid(), pre[H:2 * H].sigmoid(), pre[2 * H:3 * H].tanh(), pre[3 * H:].sigmoid()
                    c = f * c + i * g
                    h = o * c.tanh()
                    y[b, t, e * 2 * H + d * H:e * 2 * H + (d + 1) * H] = h
    return y


def run(name, gx, whh, lens, B, T, H, n_enc, save):
    y = torch.full((B, T, n_enc * 2 * H), float("nan"), device=dev)
    gates = torch.full((n_enc, 2, B, T, 5, H), float("nan"), device=dev) if save else None
    ops._call(name, gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), n_enc, B, T, H,
              y.data_ptr(), 0 if gates is None else gates.data_ptr(), *_extra(name), ops._stream())
    torch.cuda.synchronize()
    return y, gates


def compare(B, T, lengths, n_enc=1, save=False, check64=False, mixed_rows=False):
    H = 256
    g = torch.Generator(device=dev).manual_seed(B * 1000 + T)
    gx = torch.randn((n_enc, B * T, 8 * H), device=dev, generator=g)
    whh = torch.randn((n_enc, 2, 4 * H, H), device=dev, generator=g) * 0.06
    if mixed_rows:   # rows of very different magnitude (the fp16-split kernel scales every row by its own power of two)
        whh = whh * (10.0 ** (torch.rand((n_enc, 2, 4 * H, 1), device=dev, generator=g) * 5.0 - 4.0))
        whh[:, :, 5, :] = 0.0
    lens = ops.Lengths(lengths, dev, T)
    y_f, g_f = run("mts_lstm_rec_fwd", gx, whh, lens, B, T, H, n_enc, save)
    y_t, g_t = run(TC_NAME, gx, whh, lens, B, T, H, n_enc, save)
    err = float((y_f - y_t).abs().max())
    msg = f"B={B} T={T} n_enc={n_enc} save={save}{' mixed-rows' if mixed_rows else ''}: max |fma - tc| = {err:.3e}"
    ok = err < 2e-5 and not bool(torch.isnan(y_t).any())
    if save:
        valid = ~torch.isnan(g_f)
        assert bool((torch.isnan(g_t) == torch.isnan(g_f)).all()), "saved-gate coverage differs"
        gerr = float((g_f[valid] - g_t[valid]).abs().max())
        msg += f", gates {gerr:.3e}"
        ok = ok and gerr < 5e-5
    if check64:
        ref = ref_recurrence(gx, whh, lengths, B, T, H, n_enc).to(dev)
        e_f, e_t = float((y_f.double() - ref).abs().max()), float((y_t.double() - ref).abs().max())
        msg += f"; vs float64: fma {e_f:.3e}, tc {e_t:.3e}"
        ok = ok and e_t < 1e-5
    print(("ok   " if ok else "FAIL ") + msg, flush=True)
    return ok


def compare_bwd(B, T, lengths, n_enc=1):
    """backward: tensor-core kernel vs the packed-FMA cluster kernel on the same saved gates and dy."""
    H = 256
    g = torch.Generator(device=dev).manual_seed(B * 77 + T)
    gx = torch.randn((n_enc, B * T, 8 * H), device=dev, generator=g)
    whh = torch.randn((n_enc, 2, 4 * H, H), device=dev, generator=g) * 0.06
    lens = ops.Lengths(lengths, dev, T)
    y, gates = run("mts_lstm_rec_fwd", gx, whh, lens, B, T, H, n_enc, True)
    dy = torch.randn((B, T, n_enc * 2 * H), device=dev, generator=g)
    whh_t = whh.transpose(2, 3).contiguous()
    d_f = torch.full((n_enc, B * T, 8 * H), float("nan"), device=dev)
    d_t = torch.full((n_enc, B * T, 8 * H), float("nan"), device=dev)
    ops._call("mts_lstm_rec_bwd", dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), whh_t.data_ptr(), lens.dev.data_ptr(),
              lens.order.data_ptr(), n_enc, B, T, H, d_f.data_ptr(), ops._stream())
    ops._call("mts_lstm_rec_bwd_tc", dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(),
              lens.order.data_ptr(), n_enc, B, T, H, d_t.data_ptr(), ops._stream())
    torch.cuda.synchronize()
    nan = bool(torch.isnan(d_t).any())
    err = float((d_f - d_t).abs().max())
    scale = float(d_f.abs().max())
    ok = (not nan) and err < 2e-5 * max(scale, 1.0)
    print(("ok   " if ok else "FAIL ") + f"bwd B={B} T={T} n_enc={n_enc}: max |fma - tc| = {err:.3e} (scale {scale:.3e}, nan {nan})", flush=True)
    return ok


def sweep_bwd():
    H, T = 256, 300
    print("B     bwd fma ms (us/step)   bwd tc ms (us/step)")
    for B in (10, 16, 64, 128, 256):
        g = torch.Generator(device=dev).manual_seed(0)
        gx = torch.randn((1, B * T, 8 * H), device=dev, generator=g) * 0.5
        whh = torch.randn((1, 2, 4 * H, H), device=dev, generator=g) * 0.05
        lens = ops.Lengths([T] * B, dev, T)
        y, gates = run("mts_lstm_rec_fwd_tc", gx, whh, lens, B, T, H, 1, True)
        dy = torch.randn_like(y)
        whh_t = whh.transpose(2, 3).contiguous()
        dgx = torch.empty_like(gx)
        a = timeit(lambda: ops._call("mts_lstm_rec_bwd", dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), whh_t.data_ptr(),
                                     lens.dev.data_ptr(), lens.order.data_ptr(), 1, B, T, H, dgx.data_ptr(), ops._stream()))
        b = timeit(lambda: ops._call("mts_lstm_rec_bwd_tc", dy.data_ptr(), gates.data_ptr(), whh.data_ptr(),
                                     lens.dev.data_ptr(), lens.order.data_ptr(), 1, B, T, H, dgx.data_ptr(), ops._stream()))
        print(f"{B:5d} {a:9.3f} ({a * 1e3 / T:6.2f}) {b:9.3f} ({b * 1e3 / T:6.2f})", flush=True)


def timeit(fn, iters=5, warmup=2):
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


def sweep():
    H, T = 256, 300
    print("B     fma ms (us/step)    tc ms (us/step)    tc algorithmic GB/s")
    for B in (8, 16, 32, 64, 128, 224, 256, 1024, 4096):
        g = torch.Generator(device=dev).manual_seed(0)
        gx = torch.randn((1, B * T, 8 * H), device=dev, generator=g) * 0.5
        whh = torch.randn((1, 2, 4 * H, H), device=dev, generator=g) * 0.05
        lens = ops.Lengths([T] * B, dev, T)
        y = torch.empty((B, T, 2 * H), device=dev)
        call = lambda name: ops._call(name, gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), 1,
                                      B, T, H, y.data_ptr(), 0, *_extra(name), ops._stream())
        a = timeit(lambda: call("mts_lstm_rec_fwd")) if B <= 256 else float("nan")
        b = timeit(lambda: call(TC_NAME))
        print(f"{B:5d} {a:9.3f} ({a * 1e3 / T:6.2f}) {b:9.3f} ({b * 1e3 / T:6.2f}) {B * T * 10240 / b / 1e6:10.1f}", flush=True)


def timeline(B=16, T=40):
    """clock64 stamps of 4 steps of CTA 0 (MMA thread slots 0-4, epilogue thread 0 slots 5-11)."""
    H = 256
    g = torch.Generator(device=dev).manual_seed(0)
    gx = torch.randn((1, B * T, 8 * H), device=dev, generator=g) * 0.5
    whh = torch.randn((1, 2, 4 * H, H), device=dev, generator=g) * 0.05
    lens = ops.Lengths([T] * B, dev, T)
    y = torch.empty((B, T, 2 * H), device=dev)
    nsl = 16 if TC_NAME.endswith(("_h3", "_h3p")) else 12
    buf = torch.zeros(4 * nsl, dtype=torch.int64, device=dev)
    call = lambda: ops._call(TC_NAME, gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(),
                             lens.order.data_ptr(), 1, B, T, H, y.data_ptr(), 0, *_extra(TC_NAME), ops._stream())
    call(); call()
    ops._call(PROF_NAME, buf.data_ptr())
    call()
    torch.cuda.synchronize()
    ops._call(PROF_NAME, 0)
    st = buf.cpu().view(4, nsl)
    names = ["mma:start", "mma:h_full", "mma:hi issued", "mma:lo_ready", "mma:commit", "epi:start", "epi:h_full",
             "epi:lo done", "epi:acc_full", "epi:tmem ld", "epi:act+bar", "epi:sent"]
    if TC_NAME.endswith(("_h3", "_h3p")):
        names = ["mma:start", "half0", "half1", "issued0", "issued1", "commit", "-", "-", "-", "-", "epi:acc_full", "tmem ld",
                 "act+bar", "cell", "sent", "-"]
    t0 = int(st[0, 0])
    for i in range(4):
        print(f"step {8 + i}: " + "  ".join(f"{n}={int(st[i, k]) - t0}" for k, n in enumerate(names) if n != "-"))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = True
    if what in ("all", "trunc"):
        ok = trunc_check() and ok
    if what in ("all", "check"):
        ok = compare(3, 5, [5, 3, 1], check64=True) and ok
        ok = compare(16, 40, [40] * 16, check64=True) and ok
        ok = compare(37, 61, [61] + [int(x) for x in torch.randint(1, 62, (36,))], save=True) and ok
        ok = compare(20, 33, [33] + [int(x) for x in torch.randint(1, 34, (19,))], n_enc=2, save=True) and ok
        ok = compare(300, 50, [50] * 300) and ok
        ok = compare(16, 40, [40] * 16, check64=True, mixed_rows=True) and ok
        ok = compare(70, 30, [30] * 70, check64=True) and ok
    if what in ("all", "bwd"):
        ok = compare_bwd(3, 5, [5, 3, 1]) and ok
        ok = compare_bwd(16, 40, [40] * 16) and ok
        ok = compare_bwd(37, 61, [61] + [int(x) for x in torch.randint(1, 62, (36,))]) and ok
        ok = compare_bwd(20, 33, [33] + [int(x) for x in torch.randint(1, 34, (19,))], n_enc=2) and ok
        ok = compare_bwd(300, 50, [50] * 300) and ok
        if ok:
            sweep_bwd()
    if what in ("all", "timeline"):
        if len(sys.argv) > 3:
            timeline(B=int(sys.argv[2]), T=int(sys.argv[3]))
        else:
            timeline()
            timeline(B=10 * 7, T=40)   # cfg1-like: 10 episodes per tile
    if what in ("all", "sweep") and ok:
        sweep()
    sys.exit(0 if ok else 1)
