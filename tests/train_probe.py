"""Per-entry-point CUDA-event breakdown of one training step (bench.py's configs[1] leg)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodaltopicsegmentation_b200 as m
from multimodaltopicsegmentation_b200 import dist as mdist, ops
import bench
dev = torch.device("cuda:0")
c = bench.TRAIN_CFG
torch.manual_seed(0)
seg = m.TextSegmenter(2, c["D1"] + c["D2"], c["H"], num_layers=c["L"], architecture="BiLSTM", loss_fn="FocalLoss", optimizer="Adam", lr=1e-3).to(dev)
opt = seg.configure_optimizers()["optimizer"]
bucket = mdist.GradBucket(seg.parameters())
batch = m.to_device(bench.train_batch(0), dev)
for _ in range(3):
    mdist.train_step(seg, batch, opt, bucket)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    mdist.train_step(seg, batch, opt, bucket)
e.record(); torch.cuda.synchronize()
print(f"step: {s.elapsed_time(e)/5:.3f} ms, sentences {int(batch['src_lengths'].sum())}, T {int(batch['src_lengths'].max())}")
ops.PROFILE = {}
mdist.train_step(seg, batch, opt, bucket)
torch.cuda.synchronize()
tot = 0
for k, v in sorted(ops.PROFILE.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
    t = sum(a.elapsed_time(b) for a, b in v); tot += t
    print(f"{k:28s} {len(v):3d} calls {t:8.3f} ms")
print(f"sum of ABI calls: {tot:.3f} ms")
