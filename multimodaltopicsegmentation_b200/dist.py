"""Data parallelism over the GPUs of one box: one process per GPU, episodes sharded across ranks.

The reference has no distributed code of its own (only PL's implicit DDP behind `Trainer(gpus=N)`,
train_fit.py:286,293).  The path shards naturally by episode, so:

  * training : every rank runs forward/backward on its episodes with the loss normalised by the GLOBAL number of
               valid sentences, then ONE NCCL all-reduce (sum) over a flat fp32 gradient bucket that the
               parameters' .grad tensors alias (no staging copies).  Summing -- not averaging -- reproduces the
               un-sharded gradient exactly; plain DDP averaging would weight ranks equally regardless of how many
               sentences they hold and would not match the reference's mean over sentences (models/CRF.py:348-352).
  * inference: no collective on the data path; a final all-gather of the uint8 boundary predictions.

Works with the `gloo` backend on CPU tensors too (used by the host-logic tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world():
    return dist.get_world_size() if is_dist() else 1


def rank():
    return dist.get_rank() if is_dist() else 0


def shard_indices(lengths, rank_, world_):
    """Episodes of one global batch owned by `rank_`: sort by length (descending), deal round-robin.
    Balances the number of sentences -- and the recurrence step count -- across ranks."""
    host = [int(v) for v in lengths]
    order = sorted(range(len(host)), key=lambda i: (-host[i], i))
    return order[rank_::world_]


def shard_batch(batch, rank_=None, world_=None):
    """Slice an EncoderDataset batch dict (src_tokens, src_tokens2, src_lengths, tgt_tokens, id, domain)."""
    rank_ = rank() if rank_ is None else rank_
    world_ = world() if world_ is None else world_
    idx = shard_indices(batch["src_lengths"].tolist(), rank_, world_)
    sel = torch.tensor(idx, dtype=torch.long)
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.index_select(0, sel.to(v.device))
        elif isinstance(v, (list, tuple)) and len(v) == len(batch["src_lengths"]):
            out[k] = [v[i] for i in idx]
        else:
            out[k] = v
    if "src_lengths" in out and "src_tokens" in out and torch.is_tensor(out["src_tokens"]):
        tmax = int(out["src_lengths"].max())  # the collater pads to the batch maximum; keep that contract per shard
        for k in ("src_tokens", "src_tokens2", "tgt_tokens"):
            if torch.is_tensor(out.get(k)):
                out[k] = out[k][:, :tmax].contiguous()
    return out, idx


def all_reduce_sum_scalar(value, device):
    """Sum of a host integer / float over ranks (e.g. the global number of valid sentences)."""
    if not is_dist():
        return value
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


class GradBucket:
    """One flat fp32 buffer aliased by every parameter's .grad; all_reduce() sums it across ranks in one call."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, async_op=False):
        if not is_dist():
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)

    @property
    def nbytes(self):
        return self.flat.numel() * 4


def gather_tags(tags_dev):
    """Final gather of the [B_local, T] uint8 boundary predictions of every rank (fixed-size buffers)."""
    if not is_dist():
        return [tags_dev]
    out = [torch.empty_like(tags_dev) for _ in range(world())]
    dist.all_gather(out, tags_dev)
    return out


def train_step(segmenter, batch, optimizer, bucket):
    """One data-parallel optimisation step on this rank's shard of the global batch.
    `batch` is already this rank's shard (see shard_batch) and already on the device."""
    xs, lengths = segmenter._inputs(batch)
    dev = xs[0][0].device if isinstance(xs[0], (tuple, list)) else xs[0].device
    model = segmenter.model
    crf = hasattr(model, "crf")
    local = len(lengths) if crf else int(torch.as_tensor(lengths).sum())
    global_count = all_reduce_sum_scalar(local, dev)
    bucket.zero()
    loss = model.loss(*xs, lengths, batch["tgt_tokens"], global_count=global_count)
    loss.backward()
    bucket.all_reduce()
    optimizer.step()
    return loss  # this rank's share of the global mean; the sum over ranks is the un-sharded loss
