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


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates
    afterwards (first touch) and its copy-issuing thread sit next to the GPU's PCIe root.  One process per GPU: without
    this, all ranks' staging memory may land on one node and the host->device streams of 8 GPUs share that node's
    memory controllers.  Returns the node, or None when the topology cannot be read (then nothing is changed)."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


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
    """One flat fp32 buffer aliased by every parameter's .grad; all_reduce() sums it across ranks in one call.
    One extra slot behind the gradients carries this rank's normalisation count (valid sentences, or episodes for the
    CRF): reduced in the SAME collective, it gives every rank the global count on the device -- no scalar all-reduce and
    no device->host sync before the step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.buf = torch.zeros(total + 1, dtype=torch.float32, device=dev)
        self.flat = self.buf[:total]          # the gradients
        self.count = self.buf[total:]         # [1]: local count before, global count after all_reduce(with_count=True)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.buf.zero_()

    def all_reduce(self, async_op=False, with_count=False):
        if not is_dist():
            return None
        return dist.all_reduce(self.buf if with_count else self.flat, op=dist.ReduceOp.SUM, async_op=async_op)

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


def local_count(segmenter, lengths):
    """What this rank's loss is normalised by: valid sentences (focal / BCE / CE heads), episodes (CRF NLL mean)."""
    if hasattr(segmenter.model, "crf"):
        return len(lengths)
    return int(torch.as_tensor(lengths).sum()) if torch.is_tensor(lengths) else int(sum(int(v) for v in lengths))


def train_step(segmenter, batch, optimizer, bucket, global_count=None):
    """One data-parallel optimisation step on this rank's shard of the global batch.
    `batch` is already this rank's shard (see shard_batch) and already on the device; `src_lengths` stays on the host.

    global_count given (the sampler knows the whole global batch's lengths): the loss is normalised by it directly and
    the gradients are summed by one all-reduce.  Otherwise the rank back-propagates its UN-normalised loss sum, puts its
    local count into the bucket's last slot, and ONE all-reduce sums gradients and counts together; the division by the
    global count happens on the device.  Either way there is no host synchronisation inside the step.
    Returns this rank's share of the global mean loss (the sum over ranks is the un-sharded loss)."""
    xs, lengths = segmenter._inputs(batch)
    model = segmenter.model
    bucket.zero()
    if global_count is not None or not is_dist():
        count = global_count if global_count is not None else local_count(segmenter, lengths)
        loss = model.loss(*xs, lengths, batch["tgt_tokens"], global_count=count)
        loss.backward()
        bucket.all_reduce()
    else:
        loss_sum = model.loss(*xs, lengths, batch["tgt_tokens"], global_count=1.0)
        loss_sum.backward()
        bucket.count.fill_(float(local_count(segmenter, lengths)))
        bucket.all_reduce(with_count=True)
        bucket.flat.div_(bucket.count)
        loss = loss_sum.detach() / bucket.count[0]
    optimizer.step()
    return loss
