"""Why does the recurrence kernel take longer inside the step than alone?  Times ONE recurrence launch (CUDA events
around that launch only) in several contexts.  Not a test; run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodaltopicsegmentation_b200 import ops  # noqa: E402
from bench_kernels import dev, rec_setup  # noqa: E402

B, T, H = 64, 300, 256


def rec(gx, whh, lens, y, ycorr):
    ops._call("mts_lstm_rec_fwd_tc", gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), 1, B, T, H,
              y.data_ptr(), 0, 0 if ycorr is None else ycorr.data_ptr(), ops._stream())


def time_rec(pre, gx, whh, lens, y, ycorr, iters=10):
    tot = 0.0
    for i in range(iters + 3):
        if pre is not None:
            pre()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rec(gx, whh, lens, y, ycorr)
        e.record()
        torch.cuda.synchronize()
        if i >= 3:
            tot += s.elapsed_time(e)
    return tot / iters


gx, whh, lens, y, _ = rec_setup(B, T)
ycorr = torch.empty(B * T, 2 * H, device=dev)
print("alone, random gx              : %.3f ms" % time_rec(None, gx, whh, lens, y, None))
print("alone, + y_corr               : %.3f ms" % time_rec(None, gx, whh, lens, y, ycorr))
M, N, K = B * T, 8 * H, 896
a = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) * 0.03
bias = torch.randn(N, device=dev) * 0.1
a_hi, a_lo = ops.split_tf32(a)
w_hi, w_lo = ops.split_tf32(w, side=ops.B_SIDE)
gx2 = torch.empty(1, M, N, device=dev)
gemm = lambda: ops.gemm_tf32x3(a_hi, a_lo, w_hi, w_lo, bias, gx2[0], M, N, epilogue=1, ldc=N)
print("after the GEMM, random gx     : %.3f ms" % time_rec(gemm, gx, whh, lens, y, None))
print("after the GEMM, its output    : %.3f ms" % time_rec(gemm, gx2, whh, lens, y, None))
gemm()
print("alone, the GEMM's output      : %.3f ms" % time_rec(None, gx2, whh, lens, y, None))
flush = torch.empty(64 * 1024 * 1024, device=dev)
print("after an L2 flush, random gx  : %.3f ms" % time_rec(lambda: flush.zero_(), gx, whh, lens, y, None))
sm = torch.empty(1024, device=dev)
print("after a tiny kernel           : %.3f ms" % time_rec(lambda: sm.zero_(), gx, whh, lens, y, None))
whh2 = torch.nn.init.uniform_(torch.empty_like(whh), -1 / 16, 1 / 16)
print("alone, uniform(-1/16,1/16) whh: %.3f ms" % time_rec(None, gx, whh2, lens, y, None))
# back-to-back launches (what bench_kernels measures)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    rec(gx, whh, lens, y, None)
e.record()
torch.cuda.synchronize()
print("10 back-to-back launches      : %.3f ms each" % (s.elapsed_time(e) / 10))

# the same launch through the model, timed the way bench.py does (events around each ABI call, eager step)
import multimodaltopicsegmentation_b200 as m  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth  # noqa: E402

torch.manual_seed(0)
seg = m.TextSegmenter(2, 896, 256, num_layers=2, architecture="BiLSTM", loss_fn="FocalLoss", threshold=0.5).to(dev)
x1, x2, l = synth(0, B, T, 384, 512)
x1, x2 = x1.to(dev), x2.to(dev)
ln = ops.Lengths(l, dev, T)
for rep in range(2):
    ops.PROFILE = {}
    with torch.no_grad():
        for _ in range(5):
            seg.model.model((x1, x2), ln)
    torch.cuda.synchronize()
    for k, v in ops.PROFILE.items():
        print("model step: %-24s %s" % (k, " ".join("%.3f" % s.elapsed_time(e) for s, e in v)))
    ops.PROFILE = None
print("alone again, random gx        : %.3f ms" % time_rec(None, gx, whh, lens, y, None))

# re-issue the model's own recurrence launches (same pointers, same data) outside the model
captured = []
orig_call = ops._call


def spy(name, *args):
    if name == "mts_lstm_rec_fwd_tc":
        captured.append(args)
    return orig_call(name, *args)


ops._call = spy
with torch.no_grad():
    keep = seg.model.model((x1, x2), ln)
ops._call = orig_call
torch.cuda.synchronize()
for args in captured:
    for rep in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig_call("mts_lstm_rec_fwd_tc", *args)
        e.record()
        torch.cuda.synchronize()
        print("replayed model launch: %.3f ms   args n_enc=%s B=%s T=%s H=%s gates=%s ycorr=%s" % (
            s.elapsed_time(e), args[4], args[5], args[6], args[7], args[9] != 0, args[10] != 0))
# model's whh + random gx, and random whh + zero gx
a0 = list(captured[1])
a0[0] = gx.data_ptr()
for rep in range(2):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig_call("mts_lstm_rec_fwd_tc", *a0); e.record(); torch.cuda.synchronize()
    print("model whh, random gx: %.3f ms" % s.elapsed_time(e))
a1 = list(captured[1])
a1[1] = whh.data_ptr()
for rep in range(2):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig_call("mts_lstm_rec_fwd_tc", *a1); e.record(); torch.cuda.synchronize()
    print("random whh, model gx: %.3f ms" % s.elapsed_time(e))
