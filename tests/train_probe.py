"""Per-entry-point CUDA-event breakdown of one training step of a bench.py training leg, next to the host time of the
step (wall clock with a synchronise on both sides) and the time of the optimiser alone:
    python tests/train_probe.py [train | train_crf | latefusion_train | latefusion_train_b64 | latefusion_crf_train]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodaltopicsegmentation_b200 as m  # noqa: E402
from multimodaltopicsegmentation_b200 import dist as mdist, ops  # noqa: E402
import bench  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "train"
dev = torch.device("cuda:0")
arch, B, _, dims, _ = bench.TRAIN_LEGS[key]
torch.manual_seed(0)
seg = m.TextSegmenter(2, list(dims) if len(dims) > 1 else dims[0], bench.CFG["H"], num_layers=bench.CFG["L"], architecture=arch,
                      loss_fn="FocalLoss", optimizer="Adam", lr=1e-3).to(dev)
opt = seg.configure_optimizers()["optimizer"]
bucket = mdist.GradBucket(seg.parameters())
batch = m.to_device(bench.train_batch(key, 0, 0), dev)
for _ in range(3):
    mdist.train_step(seg, batch, opt, bucket)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
s.record()
for _ in range(10):
    mdist.train_step(seg, batch, opt, bucket)
e.record()
t_issue = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print(f"{key}: step {s.elapsed_time(e) / 10:.3f} ms on the device, host issue time {t_issue * 1e3:.3f} ms per step, "
      f"sentences {int(batch['src_lengths'].sum())}, T {int(batch['src_lengths'].max())}")
s.record()
for _ in range(10):
    opt.step()
e.record()
torch.cuda.synchronize()
print(f"optimizer.step alone: {s.elapsed_time(e) / 10:.3f} ms")
s.record()
for _ in range(10):
    bucket.zero()
e.record()
torch.cuda.synchronize()
print(f"bucket.zero alone: {s.elapsed_time(e) / 10:.3f} ms")
ops.PROFILE = {}
mdist.train_step(seg, batch, opt, bucket)
torch.cuda.synchronize()
tot = 0
for k, v in sorted(ops.PROFILE.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
    t = sum(a.elapsed_time(b) for a, b in v)
    tot += t
    print(f"{k:28s} {len(v):3d} calls {t:8.3f} ms")
print(f"sum of ABI calls: {tot:.3f} ms")
