"""Input layout of the hot path: the reference's in-memory episode datasets and their collaters
(EncoderDataset.py:18-152, 154-232).  The batch dict (`src_tokens`, `src_tokens2`, `src_lengths`,
`tgt_tokens`, `id`, `domain`), the zero padding, the tag pad value (0 for CRF architectures, -1 otherwise) and
the truncate semantics are kept verbatim; the PCA projection option is outside the hot path.
"""
from __future__ import annotations

import torch
from torch.utils.data import Dataset


def _pad_episodes(values, truncate, truncate_value, fill=0.0):
    """list of [T_i, D] (or [T_i]) -> [B, T, D] padded with `fill`; T = truncate_value or the batch maximum."""
    T = truncate_value if truncate else max(v.size(0) for v in values)
    shape = (len(values), T) + tuple(values[0].shape[1:])
    out = torch.full(shape, float(fill))
    for i, v in enumerate(values):
        n = min(T, v.size(0))
        out[i, :n] = v[:n]
    return out


class AudioPortionDataset(Dataset):
    def __init__(self, lines, tag_to_ix, encoder="x-vectors", CRF=True, truncate=True, truncate_value=100,
                 umap_project=False, umap_project_value=100, umap_class=None, second_input=None, domain_adapt=False):
        if umap_project:
            raise NotImplementedError("PCA/UMAP projection is outside the B200 hot path")
        self.minus = 0 if CRF else 1
        self.embeddings = [line[0] for line in lines]
        self.tgt_dataset = [line[1] for line in lines]
        self.embeddings2 = [line[0] for line in second_input] if second_input is not None else []
        self.truncate, self.tv = truncate, truncate_value
        self.encoder_name = encoder
        self.da = bool(domain_adapt)
        if self.da:  # RadioNews files start with a digit, NonNews files do not (EncoderDataset.py:36-44)
            self.domain = [1 if str(line[2])[0].isdigit() else 0 for line in lines]
        else:
            self.domain = [None for _ in lines]
        self.reducer = None

    def __getitem__(self, index):
        item = {"id": torch.tensor(index), "target": self.tgt_dataset[index], "embeddings": self.embeddings[index],
                "domain": self.domain[index]}
        if self.embeddings2:
            item["embeddings2"] = self.embeddings2[index]
        return item

    def __len__(self):
        return len(self.embeddings)

    def collater(self, samples):
        if len(samples) == 0:
            return {}
        src = [s["embeddings"] for s in samples]
        if src[0].dim() < 2:
            src_tokens = torch.stack(src)
        else:
            src_tokens = _pad_episodes(src, self.truncate, self.tv)
        src_tokens2 = _pad_episodes([s["embeddings2"] for s in samples], self.truncate, self.tv) if self.embeddings2 else None
        tgt = [torch.as_tensor(s["target"], dtype=torch.float32) for s in samples]
        tgt_tokens = _pad_episodes(tgt, self.truncate, self.tv, fill=-self.minus)
        if self.truncate:
            lengths = [min(self.tv, len(s["embeddings"])) for s in samples]
        else:
            lengths = [len(s["embeddings"]) for s in samples]
        return {"id": torch.tensor([int(s["id"]) for s in samples]), "src_tokens": src_tokens,
                "src_lengths": torch.LongTensor(lengths), "tgt_tokens": tgt_tokens, "src_tokens2": src_tokens2,
                "domain": [s["domain"] for s in samples] if self.da else None}


class AudioPortionDatasetInference(Dataset):
    def __init__(self, lines, encoder="x-vectors", CRF=True, truncate=False, truncate_value=100, umap_project=False,
                 umap_project_value=100, umap_class=None):
        if umap_project:
            raise NotImplementedError("PCA/UMAP projection is outside the B200 hot path")
        self.minus = 0 if CRF else 1
        self.embeddings = lines
        self.truncate, self.tv = truncate, truncate_value
        self.encoder_name = encoder
        self.reducer = None

    def __getitem__(self, index):
        return {"id": torch.tensor(index), "embeddings": self.embeddings[index]}

    def __len__(self):
        return len(self.embeddings)

    def collater(self, samples):
        if len(samples) == 0:
            return {}
        src = [s["embeddings"] for s in samples]
        src_tokens = torch.stack(src) if src[0].dim() < 2 else _pad_episodes(src, self.truncate, self.tv)
        if self.truncate:
            lengths = [self.tv for _ in samples]  # sic: the reference reports tv even for shorter episodes
        else:
            lengths = [len(s["embeddings"]) for s in samples]
        return {"id": torch.tensor([int(s["id"]) for s in samples]), "src_tokens": src_tokens,
                "src_lengths": torch.LongTensor(lengths)}


def to_device(batch, device, non_blocking=True, lengths_on_host=True):
    """What pl.Trainer does with a batch dict: move every tensor, leave the rest.  `src_lengths` is consumed on the
    host by the reference (lengths.data.tolist(), NeuralArchitectures.py:98) and by our launch planning, so by default
    it stays there and no device->host sync is needed per step; the modules accept it on either side."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v) and not (lengths_on_host and k == "src_lengths"):
            out[k] = v.to(device, non_blocking=non_blocking)
        else:
            out[k] = v
    return out


class DevicePrefetcher:
    """Loader -> device pipeline: iterate over the batch dicts of a DataLoader (or any
    iterable of collater outputs) with the host->device copy of batch i+1 running on a side CUDA stream while batch
    i computes.  The reference leaves this to pl.Trainer, which copies on the compute stream; here the copy of the
    next [B,T,D] embedding block (the only large transfer of the path) hides behind the kernels of the current one.

    Two sets of device staging buffers are allocated once and reused (no allocator traffic per step): the copy into
    slot k waits for the event recorded on the compute stream when the consumer asked for the batch after the one
    that last used slot k -- i.e. after all of its work had been issued.
    Host tensors should be pinned (DataLoader(pin_memory=True)); `src_lengths` stays on the host (see to_device).
    `src_tokens` may be a (text, audio) pair of tensors: the early-fusion concat then happens inside the operand
    packing kernel instead of on the host."""

    def __init__(self, loader, device, lengths_on_host=True, depth=2):
        self.loader, self.device, self.lengths_on_host = loader, torch.device(device), lengths_on_host
        self.stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.buffers = [dict() for _ in range(depth)]
        self.free_ev = [None] * depth

    def _copy(self, v, slot, key):
        buf = self.buffers[slot].get(key)
        if buf is None or buf.numel() < v.numel() or buf.dtype != v.dtype:
            buf = torch.empty(max(v.numel(), 1), dtype=v.dtype, device=self.device)
            self.buffers[slot][key] = buf
        dst = buf[: v.numel()].view(v.shape)
        dst.copy_(v, non_blocking=True)
        return dst

    def _stage(self, batch, slot):
        def mv(v, key):
            if torch.is_tensor(v):
                return self._copy(v, slot, key)
            if isinstance(v, (tuple, list)) and v and all(torch.is_tensor(x) for x in v):
                return type(v)(self._copy(x, slot, (key, i)) for i, x in enumerate(v))
            return v

        with torch.cuda.stream(self.stream):
            if self.free_ev[slot] is not None:
                self.stream.wait_event(self.free_ev[slot])
            out = {k: (v if (self.lengths_on_host and k == "src_lengths") else mv(v, k)) for k, v in batch.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev, slot

    def __iter__(self):
        return self.iterate(self.loader)

    def iterate(self, loader):
        """Iterate over another loader with the same staging buffers (e.g. one prefetcher for all epochs)."""
        it = iter(loader)
        try:
            nxt = self._stage(next(it), 0)
        except StopIteration:
            return
        compute = torch.cuda.current_stream(self.device)
        while nxt is not None:
            cur, ev, slot = nxt
            try:
                nxt = self._stage(next(it), (slot + 1) % self.depth)
            except StopIteration:
                nxt = None
            compute.wait_event(ev)
            yield cur
            done = torch.cuda.Event()  # everything the consumer did with `cur` has been issued by now
            done.record(compute)
            self.free_ev[slot] = done

    def __len__(self):
        return len(self.loader)
