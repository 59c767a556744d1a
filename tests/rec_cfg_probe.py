"""Time the fp16-split recurrence at one batch size under the MTS_REC_NT / MTS_REC_EPT overrides of this process
(the overrides are read once per process): python tests/rec_cfg_probe.py B [T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodaltopicsegmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
NAME = os.environ.get("REC_TC_NAME", "mts_lstm_rec_fwd_h3")
H = 256
g = torch.Generator(device=dev).manual_seed(0)
gx = torch.randn((1, B * T, 8 * H), device=dev, generator=g) * 0.5
whh = torch.randn((1, 2, 4 * H, H), device=dev, generator=g) * 0.05
lens = ops.Lengths([T] * B, dev, T)
y = torch.empty((B, T, 2 * H), device=dev)
extra = (0, 0) if NAME.endswith(("_h3", "_h3p")) else (0,)
call = lambda: ops._call(NAME, gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), 1, B, T, H,
                         y.data_ptr(), 0, *extra, ops._stream())
for _ in range(3):
    call()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    call()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print(f"{NAME} B={B} T={T} NT={os.environ.get('MTS_REC_NT', 'auto')} EPT={os.environ.get('MTS_REC_EPT', 'auto')}: "
      f"{ms:.3f} ms ({ms * 1e3 / T:.2f} us/step), {B * T * 10240 / ms / 1e6:.0f} GB/s algorithmic")
