"""Time the BPTT kernel at one batch size under the MTS_REC_EPT override of this process: python tests/bwd_ept_probe.py B T [n_enc]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodaltopicsegmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T = int(sys.argv[1]), int(sys.argv[2])
n_enc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H = 256
g = torch.Generator(device=dev).manual_seed(0)
gates = torch.rand((n_enc, 2, B, T, 5, H), device=dev, generator=g)
whh = torch.randn((n_enc, 2, 4 * H, H), device=dev, generator=g) * 0.05
dy = torch.randn((B, T, n_enc * 2 * H), device=dev, generator=g)
dgx = torch.empty((n_enc, B * T, 8 * H), device=dev)
lens = ops.Lengths([T] * B, dev, T)
call = lambda: ops._call("mts_lstm_rec_bwd_h3", dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(),
                         lens.order.data_ptr(), n_enc, B, T, H, dgx.data_ptr(), ops._stream())
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
print(f"bwd_h3 B={B} T={T} n_enc={n_enc} EPT={os.environ.get('MTS_REC_EPT', 'auto')}: {ms:.3f} ms ({ms * 1e3 / T:.2f} us/step)")
