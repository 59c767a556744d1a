"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic in multimodaltopicsegmentation_b200/dist.py:
episode sharding, global-N loss weighting, the flat gradient bucket and the final tag gather.  The arithmetic
inside each rank is done by the CPU oracle's torch twin (the CUDA kernels need a GPU); what is under test is that
shard -> local backward -> ONE summed all-reduce reproduces the un-sharded gradient of the reference's loss."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_batch():
    g = torch.Generator().manual_seed(3)
    lengths = torch.tensor([9, 3, 14, 6, 11, 2, 7])
    B, T, D = len(lengths), int(lengths.max()), 10
    x = torch.randn(B, T, D, generator=g)
    y = (torch.rand(B, T, generator=g) < 0.3).float()
    for b, n in enumerate(lengths.tolist()):
        x[b, n:] = 0
        y[b, n:] = -1
    return {"src_tokens": x, "src_tokens2": None, "src_lengths": lengths, "tgt_tokens": y,
            "id": torch.arange(B), "domain": None}


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from multimodaltopicsegmentation_b200 import dist as mdist
    from oracle import ref_torch as rt

    torch.manual_seed(0)
    model = rt.Segmenter(2, 10, 8, num_layers=2, loss_fn="FocalLoss")
    batch = _make_batch()
    shard, idx = mdist.shard_batch(batch)
    bucket = mdist.GradBucket(model.parameters())
    n_local = int(shard["src_lengths"].sum())
    n_global = mdist.all_reduce_sum_scalar(n_local, "cpu")
    bucket.zero()
    # the oracle's loss is a mean over the LOCAL sentences: rescale to the global-N normalisation the kernels use
    loss = model.loss(shard["src_tokens"], shard["src_lengths"], shard["tgt_tokens"]) * (n_local / n_global)
    loss.backward()
    bucket.all_reduce()
    flat1 = bucket.flat.clone()
    # second protocol (dist.train_step without a known global count): un-normalised local loss SUM, the local count rides
    # in the bucket's extra slot, ONE all-reduce for both, division on the device -- no host synchronisation
    bucket.zero()
    loss_sum = model.loss(shard["src_tokens"], shard["src_lengths"], shard["tgt_tokens"]) * n_local
    loss_sum.backward()
    bucket.count.fill_(float(n_local))
    bucket.all_reduce(with_count=True)
    count_after = float(bucket.count[0])
    bucket.flat.div_(bucket.count)
    flat2 = bucket.flat.clone()
    bucket.flat.copy_(flat1)
    tags = torch.full((4, 14), 255, dtype=torch.uint8)
    tags[: len(idx), 0] = torch.tensor(idx, dtype=torch.uint8)
    gathered = mdist.gather_tags(tags)
    total = torch.tensor([float(loss)], dtype=torch.float64)
    dist.all_reduce(total)
    torch.save({"idx": idx, "n_global": n_global, "flat": bucket.flat.clone(), "flat2": flat2, "count_after": count_after, "loss_sum": total.item(),
                "grad_is_view": all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in bucket.params),
                "gathered": [g[:, 0].tolist() for g in gathered], "T": shard["src_tokens"].shape[1]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_backward_plus_one_allreduce_equals_unsharded(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    from oracle import ref_torch as rt

    batch = _make_batch()
    # sharding: disjoint cover, length-sorted round-robin, per-shard crop to the local maximum
    assert sorted(r0["idx"] + r1["idx"]) == list(range(7))
    assert r0["idx"][0] == 2 and r1["idx"][0] == 4  # the two longest episodes go to different ranks
    assert r0["T"] == 14 and r1["T"] == 11
    assert r0["n_global"] == r1["n_global"] == int(batch["src_lengths"].sum())
    # un-sharded reference
    torch.manual_seed(0)
    model = rt.Segmenter(2, 10, 8, num_layers=2, loss_fn="FocalLoss")
    loss = model.loss(batch["src_tokens"], batch["src_lengths"], batch["tgt_tokens"])
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.equal(r0["flat"], r1["flat"])  # every rank holds the same reduced bucket
    np.testing.assert_allclose(r0["flat"].numpy(), flat.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(r0["loss_sum"], float(loss), rtol=1e-5)
    # the fused-count protocol gives the same gradients and every rank learns the global count from the collective
    assert r0["count_after"] == r1["count_after"] == float(batch["src_lengths"].sum())
    assert torch.equal(r0["flat2"], r1["flat2"])
    np.testing.assert_allclose(r0["flat2"].numpy(), flat.numpy(), rtol=1e-4, atol=1e-7)
    assert r0["grad_is_view"] and r1["grad_is_view"]
    # final gather: rank order, fixed-size buffers
    assert r0["gathered"] == r1["gathered"]
    assert r0["gathered"][0][: len(r0["idx"])] == r0["idx"] and r0["gathered"][1][: len(r1["idx"])] == r1["idx"]


def test_shard_indices_balance():
    from multimodaltopicsegmentation_b200.dist import shard_indices

    rng = np.random.default_rng(0)
    lengths = rng.integers(84, 2438, size=37).tolist()
    for world in (2, 4, 8):
        shards = [shard_indices(lengths, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(37))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)  # within one episode of each other
