"""Development probe for the tcgen05 banded-attention kernel (run on the GPU box): compares mts_band_attn_fwd_tc with the
CUDA-core kernel on a ladder of shapes (float64 dense masked attention as the referee for the small ones) and times
both at the configs[2] geometry.  `python tests/attn_tc_check.py [quick]`"""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from multimodaltopicsegmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def dense_ref(qkv, lens, B, S, h, hd, w):
    d = h * hd
    x = qkv.double().view(B, S, 3, h, hd)
    q, k, v = x[:, :, 0].permute(0, 2, 1, 3) / (hd ** 0.5), x[:, :, 1].permute(0, 2, 1, 3), x[:, :, 2].permute(0, 2, 1, 3)
    s = q @ k.transpose(-1, -2)
    i = torch.arange(S, device=qkv.device)
    band = (i[:, None] - i[None, :]).abs() <= w
    out = torch.zeros(B, h, S, hd, dtype=torch.float64, device=qkv.device)
    for b, n in enumerate(lens):
        m = band & (i[None, :] < n)
        sb = s[b].masked_fill(~m, float("-inf"))
        p = torch.softmax(sb[:, :n], dim=-1)
        out[b, :, :n] = p @ v[b]
    return out.permute(0, 2, 1, 3).reshape(B * S, d)


def run(entry, qkv, L, B, S, h, hd, w, split=False):
    d = h * hd
    out = torch.full((B * S, d), float("nan"), device=dev)
    lse = torch.full((B, h, S), float("nan"), device=dev)
    if split:
        hl = torch.full((2, B * S, d), float("nan"), device=dev)
        ops._call(entry, qkv.data_ptr(), 3 * d, L.dev.data_ptr(), 0, B, S, h, hd, w, 0, hl[0].data_ptr(), hl[1].data_ptr(), d,
                  lse.data_ptr(), ops._stream())
        return hl[0], lse, hl[1]
    ops._call(entry, qkv.data_ptr(), 3 * d, L.dev.data_ptr(), 0, B, S, h, hd, w, out.data_ptr(), 0, 0, 0, lse.data_ptr(), ops._stream())
    return out, lse, None


cases = [(1, 64, 1, 16, 4, [64]), (1, 128, 1, 32, 8, [128]), (2, 130, 3, 64, 8, [130, 31]), (3, 200, 2, 112, 48, [200, 77, 1]),
         (2, 96, 2, 32, 100, [96, 50]), (1, 700, 1, 16, 360, [650]), (3, 300, 2, 128, 0, [300, 1, 64]),
         (5, 960, 8, 112, 48, [960, 100, 513, 128, 129])]
ok = True
for (B, S, h, hd, w, lens) in cases:
    g = torch.Generator(device=dev).manual_seed(S + hd + w)
    qkv = torch.randn(B * S, 3 * h * hd, device=dev, generator=g)
    L = ops.Lengths(lens, dev, S)
    try:
        o_t, l_t, _ = run("mts_band_attn_fwd_tc", qkv, L, B, S, h, hd, w)
        torch.cuda.synchronize()
        o_s, l_s, _ = run("mts_band_attn_fwd_simt", qkv, L, B, S, h, hd, w)
        torch.cuda.synchronize()
    except Exception as exc:  # a trap poisons the context: stop
        print(f"case {(B, S, h, hd, w)}: EXCEPTION {type(exc).__name__}: {exc}")
        ok = False
        break
    ref = dense_ref(qkv, lens, B, S, h, hd, w)
    e_t = (o_t.double() - ref).abs().max().item()
    e_s = (o_s.double() - ref).abs().max().item()
    nan_t = int(torch.isnan(o_t).sum())
    e_l = (l_t - l_s).abs().max().item()
    o2, _, lo = run("mts_band_attn_fwd_tc", qkv, L, B, S, h, hd, w, split=True)
    same = bool(torch.equal(o2, o_t))
    good = nan_t == 0 and e_t < 2e-5 and e_l < 1e-4 and same
    ok &= good
    print(f"case B{B} S{S} h{h} hd{hd} w{w}: tc err {e_t:.3e} (simt {e_s:.3e}) lse diff {e_l:.2e} nan {nan_t} split-run identical {same} "
          f"{'OK' if good else 'FAIL'}", flush=True)
    if not good and nan_t == 0:
        bad = ((o_t.double() - ref).abs() > 2e-5).nonzero()
        print("   first bad (row, col):", bad[:5].tolist(), " rows affected:", sorted(set(bad[:, 0].tolist()))[:20])

if ok and len(sys.argv) < 2:
    # timing at the configs[2] geometry: 256 episodes x 960, ragged lengths 100..960, 8 heads x 112, reaches 48 and 8
    from bench import xf_batch

    B, S, h, hd = 256, 960, 8, 112
    lengths = xf_batch(0, B)
    L = ops.Lengths(lengths, dev, S)
    N = int(lengths.sum())
    qkv = torch.randn(N, 3 * h * hd, device=dev)
    out = torch.empty(N, h * hd, device=dev)
    lo = torch.empty(N, h * hd, device=dev)
    for w in (48, 24, 8):
        for entry in ("mts_band_attn_fwd_tc", "mts_band_attn_fwd_simt"):
            f = lambda: ops._call(entry, qkv.data_ptr(), 3 * h * hd, L.dev.data_ptr(), L.offs.data_ptr(), B, S, h, hd, w, 0,  # noqa: E731
                                  out.data_ptr(), lo.data_ptr(), h * hd, 0, ops._stream())
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                f()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            print(f"cfg3 geometry reach {w}: {entry} {ms:.3f} ms  -> {N * 16 * 896 / ms / 1e6:.0f} GB/s algorithmic "
                  f"({N * 16 * 896 / ms / 1e6 / 6551.4:.3f} of HBM peak)", flush=True)
if ok and len(sys.argv) < 2:
    # timeline of CTA 0's first work items (clock64 stamps, cycles relative to the first stamp)
    from multimodaltopicsegmentation_b200 import _lib

    NI, NR, NS = 6, 5, 40
    buf = torch.zeros(NI * NR * NS, dtype=torch.int64, device=dev)
    _lib.call("mts_debug_attn_profile", buf.data_ptr())
    ops._call("mts_band_attn_fwd_tc", qkv.data_ptr(), 3 * h * hd, L.dev.data_ptr(), L.offs.data_ptr(), B, S, h, hd, 48, 0,
              out.data_ptr(), lo.data_ptr(), h * hd, 0, ops._stream())
    torch.cuda.synchronize()
    _lib.call("mts_debug_attn_profile", 0)
    st = buf.cpu().view(NI, NR, NS)
    if not bool((st > 0).any()):
        print("(timeline stamps are compiled out: build csrc/attn_tc.cu with -DMTS_ATTN_TIMELINE)")
        print("ALL OK")
        sys.exit(0)
    t0 = int(st[st > 0].min())
    names = ["softmax", "correct", "K prod ", "V prod ", "MMA    "]
    for it in range(NI):
        for role in range(NR):
            row = [(s_, int(v) - t0) for s_, v in enumerate(st[it, role].tolist()) if v > 0]
            print(f"item {it} {names[role]}: " + " ".join(f"{s_}:{v}" for s_, v in row))
print("ALL OK" if ok else "FAILED")
