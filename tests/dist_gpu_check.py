"""Multi-GPU parity check, launched under torchrun on N GPUs of one box (gpurun --gpus N):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py

Every rank takes its shard of ONE global batch, runs the CUDA forward/backward with the loss normalised by the
global sentence count, and all-reduces the flat gradient bucket over NCCL.  Rank 0 compares the reduced gradients
and the summed loss with the un-sharded CPU oracle, and the gathered boundary predictions with the oracle's.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import multimodaltopicsegmentation_b200 as m  # noqa: E402
from multimodaltopicsegmentation_b200 import dist as mdist  # noqa: E402
from oracle import ref_torch as rt  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(7)
    B, D, H = 8 * world + 3, 72, 256
    lengths = torch.randint(5, 90, (B,), generator=g)
    T = int(lengths.max())
    x = torch.randn(B, T, D, generator=g)
    y = (torch.rand(B, T, generator=g) < 0.15).float()
    for b, n in enumerate(lengths.tolist()):
        x[b, n:] = 0
        y[b, n:] = -1
    batch = {"src_tokens": x, "src_tokens2": None, "src_lengths": lengths, "tgt_tokens": y, "id": torch.arange(B),
             "domain": None}
    torch.manual_seed(0)
    ref = rt.Segmenter(2, D, H, num_layers=2, loss_fn="FocalLoss", threshold=0.5)
    seg = m.TextSegmenter(2, D, H, num_layers=2, architecture="BiLSTM", loss_fn="FocalLoss", optimizer="SGD", lr=0.0,
                          threshold=0.5)
    seg.model.load_state_dict(ref.state_dict())
    seg = seg.to(dev)
    shard, idx = mdist.shard_batch(batch)
    shard = m.to_device(shard, dev)
    bucket = mdist.GradBucket(seg.parameters())
    opt = torch.optim.SGD(seg.parameters(), lr=0.0)
    loss = mdist.train_step(seg, shard, opt, bucket)
    total = loss.detach().clone().double()
    dist.all_reduce(total)
    # inference: local decode, then the final gather of fixed-size uint8 tag buffers
    seg.model.th = 0.5
    _, tags = seg.model(shard["src_tokens"], shard["src_lengths"])
    buf = torch.full(((B + world - 1) // world, T), 255, dtype=torch.uint8, device=dev)
    for i, t in enumerate(tags):
        buf[i, : len(t)] = torch.tensor(t, dtype=torch.uint8, device=dev)
    gathered = mdist.gather_tags(buf)
    ok = True
    if rank == 0:
        ref_loss = ref.loss(x, lengths, y)
        ref_loss.backward()
        flat_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        got = bucket.flat.cpu()
        scale = float(flat_ref.abs().max())
        err = float((got - flat_ref).abs().max())
        print(f"[dist check] world={world} loss {float(total):.7f} vs oracle {float(ref_loss):.7f}; "
              f"max grad err {err:.3e} (scale {scale:.3e}); bucket {bucket.nbytes / 1e6:.1f} MB")
        ok &= abs(float(total) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
        ok &= err <= 1e-4 * scale
        _, ref_tags = ref(x, lengths)
        p = torch.sigmoid(ref.classification(ref.model(x, lengths)))[:, :, 0]
        mism = 0
        for r in range(world):
            own = mdist.shard_indices(lengths.tolist(), r, world)
            host = gathered[r].cpu().numpy()
            for i, b in enumerate(own):
                for t in range(int(lengths[b])):
                    if abs(float(p[b, t]) - 0.5) > 1e-6 and bool(host[i, t]) != ref_tags[b][t]:
                        mism += 1
        print(f"[dist check] gathered tags: {mism} mismatches against the oracle")
        ok &= mism == 0
        print("[dist check] PASS" if ok else "[dist check] FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
