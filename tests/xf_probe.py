"""configs[2] windowed-attention leg alone (bench.transformer_bench), for A/B runs of kernel variants."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodaltopicsegmentation_b200 as m
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
out = bench.transformer_bench(m, dev, 0, 1, 3, torch.cuda.synchronize)
print(json.dumps({k: out[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}))
print(out["roofline_attention"]["ms_per_layer"])
