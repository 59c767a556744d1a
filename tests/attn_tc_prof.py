"""ncu target: a few launches of the tcgen05 banded-attention kernel at the configs[2] geometry (reach from argv, default 48)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from bench import xf_batch  # noqa: E402
from multimodaltopicsegmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
w = int(sys.argv[1]) if len(sys.argv) > 1 else 48
entry = sys.argv[2] if len(sys.argv) > 2 else "mts_band_attn_fwd_tc"
B, S, h, hd = 256, 960, 8, 112
lengths = xf_batch(0, B)
L = ops.Lengths(lengths, dev, S)
N = int(lengths.sum())
qkv = torch.randn(N, 3 * h * hd, device=dev)
out = torch.empty(N, h * hd, device=dev)
lo = torch.empty(N, h * hd, device=dev)
for _ in range(3):
    ops._call(entry, qkv.data_ptr(), 3 * h * hd, L.dev.data_ptr(), L.offs.data_ptr(), B, S, h, hd, w, 0, out.data_ptr(), lo.data_ptr(),
              h * hd, 0, ops._stream())
torch.cuda.synchronize()
print("done")
