"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the development container only (it needs /root/reference):

    python tests/golden/make_golden.py

The reference (Ighina/MultimodalTopicSegmentation) has no tests and no golden
vectors of its own, so the pin for the oracle is the output of
the reference's own PyTorch modules executed here on CPU in fp32.  The only
patch is a stub for `models/longformer_noffn.py`, a file missing from the
upstream checkout; nothing on the hot path uses it.

Each fixture is one .npz:  inputs, every parameter under its state-dict name
("p:<name>"), outputs ("o:<name>") and gradients of the loss ("g:<name>").
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MTS_REFERENCE", "/root/reference")


def import_reference():
    sys.path.insert(0, REF)
    stub = types.ModuleType("models.longformer_noffn")

    class LongformerLayer(nn.Module):  # never instantiated on the hot path
        def __init__(self, *a, **k):
            super().__init__()

    stub.LongformerLayer = LongformerLayer
    sys.modules["models.longformer_noffn"] = stub
    import models.CRF as ref_crf  # noqa: E402
    import models.NeuralArchitectures as ref_na  # noqa: E402
    import models.focal_loss as ref_fl  # noqa: E402

    return ref_crf, ref_na, ref_fl


def tags_to_array(tags, T):
    out = -np.ones((len(tags), T), dtype=np.int8)
    for i, t in enumerate(tags):
        out[i, : len(t)] = np.asarray(t, dtype=np.int8)
    return out


def pack(model, inputs, outputs, loss=None):
    d = {}
    for k, v in inputs.items():
        d["i:" + k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    for k, v in model.state_dict().items():
        d["p:" + k] = v.detach().cpu().numpy()
    for k, v in outputs.items():
        d["o:" + k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    if loss is not None:
        model.zero_grad()
        loss.backward()
        d["o:loss"] = loss.detach().cpu().numpy()
        for k, p in model.named_parameters():
            if p.grad is not None:
                d["g:" + k] = p.grad.detach().cpu().numpy().copy()
    return d


def median_threshold(model, scores, lengths):
    """A threshold in the middle of the model's own probabilities, so the golden tags are mixed."""
    with torch.no_grad():
        if model.bce:
            p = torch.sigmoid(scores)[:, :, 0]
        else:
            p = torch.softmax(scores, dim=2)[:, :, 1]
        vals = torch.cat([p[b, :l] for b, l in enumerate(lengths.tolist())])
    return round(float(vals.median()), 3)


def labels(gen, B, T, lengths, pad):
    y = (torch.rand(B, T, generator=gen) < 0.25).float()
    for b, l in enumerate(lengths):
        y[b, l - 1] = 0.0
        y[b, l:] = pad
    return y


def main():
    torch.set_num_threads(1)
    R, NA, FL = import_reference()
    g = torch.Generator().manual_seed(20261018)

    # (i) early-fusion BiLSTM, three loss heads -------------------------------------------
    for loss_fn in ("FocalLoss", "BinaryCrossEntropy", "CrossEntropy"):
        torch.manual_seed(1)
        B, T, D, H, L = 3, 9, 12, 8, 2
        lengths = torch.tensor([9, 4, 6])
        m = R.BiLSTM(2, D, H, num_layers=L, loss_fn=loss_fn, threshold=None, device="cpu")
        with torch.no_grad():  # make the head non-trivial so that tags are mixed
            m.classification.weight.mul_(6.0)
        x = torch.randn(B, T, D, generator=g)
        y = labels(g, B, T, lengths.tolist(), -1.0)
        th = median_threshold(m, m(x, lengths)[0], lengths)
        m.th = th
        scores, tags = m(x, lengths)
        loss = m.loss(x, lengths, y)
        d = pack(m, {"x": x, "lengths": lengths, "y": y, "th": th},
                 {"scores": scores, "tags": tags_to_array(tags, T)}, loss)
        np.savez_compressed(os.path.join(HERE, f"bilstm_{loss_fn.lower()}.npz"), **d)

    # unidirectional / single layer RNN encoder only (covers the non-bidirectional branch)
    torch.manual_seed(2)
    enc = NA.RNN(7, 8, num_layers=1, bidirectional=True)
    x = torch.randn(4, 11, 7, generator=g)
    lengths = torch.tensor([3, 11, 1, 8])
    out = enc(x, lengths)
    np.savez_compressed(os.path.join(HERE, "rnn_encoder.npz"),
                        **pack(enc, {"x": x, "lengths": lengths}, {"out": out}))

    # (iv) late fusion ----------------------------------------------------------------------
    torch.manual_seed(3)
    B, T, H, L = 3, 9, 8, 2
    lengths = torch.tensor([5, 9, 7])
    m = R.BiLSTMLateFusion(2, [5, 7], H, num_layers=L, loss_fn="FocalLoss", device="cpu")
    with torch.no_grad():
        m.classification.weight.mul_(6.0)
    x1 = torch.randn(B, T, 5, generator=g)
    x2 = torch.randn(B, T, 7, generator=g)
    y = labels(g, B, T, lengths.tolist(), -1.0)
    th = median_threshold(m, m(x1, x2, lengths)[0], lengths)
    m.th = th
    scores, tags = m(x1, x2, lengths)
    loss = m.loss(x1, x2, lengths, y)
    np.savez_compressed(os.path.join(HERE, "latefusion_focal.npz"),
                        **pack(m, {"x1": x1, "x2": x2, "lengths": lengths, "y": y, "th": th},
                               {"scores": scores, "tags": tags_to_array(tags, T)}, loss))

    # (ii) CRF stand-alone --------------------------------------------------------------------
    for name, (B, Lq, F, lens) in {"crf_small": (2, 6, 16, [6, 3]),
                                   "crf_ragged": (6, 40, 16, [40, 1, 17, 33, 2, 25])}.items():
        torch.manual_seed(4)
        crf = R.CRF(F, 2)
        with torch.no_grad():  # let the emissions, not the random transitions, drive the path
            crf.fc.weight.mul_(8.0)
        feats = torch.randn(B, Lq, F, generator=g)
        lengths = torch.tensor(lens)
        masks = NA.create_mask(feats, lengths)
        ys = (torch.rand(B, Lq, generator=g) < 0.3).float()
        for b, l in enumerate(lens):
            ys[b, l:] = 0.0  # CRF pad value is 0 (EncoderDataset.py:23,117)
        best, paths = crf(feats, masks)
        emis = crf.fc(feats)
        loss = crf.loss(feats, ys, masks)
        d = pack(crf, {"features": feats, "lengths": lengths, "ys": ys},
                 {"best_score": best, "paths": tags_to_array(paths, Lq), "emissions": emis}, loss)
        # gradient w.r.t. the emissions themselves (what the CUDA backward kernel must produce)
        e2 = emis.detach().clone().requires_grad_(True)
        fwd = crf._CRF__forward_algorithm(e2, masks.float())
        gold = crf._CRF__score_sentence(e2, ys.long(), masks.float())
        (fwd - gold).mean().backward()
        d["o:forward_score"] = fwd.detach().numpy()
        d["o:gold_score"] = gold.detach().numpy()
        d["g:emissions"] = e2.grad.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)

    # (v) the intended BiRnnCrf wiring: reference RNN -> reference CRF -------
    torch.manual_seed(5)
    B, T, D, H, L = 3, 9, 12, 8, 2
    lengths = torch.tensor([9, 4, 6])
    rnn = NA.RNN(D, H, L, 2, True, 0.0, 0.0, batch_first=True, LSTM=True)
    crf = R.CRF(2 * H, 2)
    with torch.no_grad():
        crf.fc.weight.mul_(30.0)
    holder = nn.Module()
    holder.model = rnn
    holder.crf = crf
    x = torch.randn(B, T, D, generator=g)
    ys = (torch.rand(B, T, generator=g) < 0.3).float()
    for b, l in enumerate(lengths.tolist()):
        ys[b, l:] = 0.0
    masks = NA.create_mask(x, lengths)
    feats = rnn(x, lengths)
    best, paths = crf(feats, masks)
    loss = crf.loss(rnn(x, lengths), ys, masks)
    np.savez_compressed(os.path.join(HERE, "bilstm_crf.npz"),
                        **pack(holder, {"x": x, "lengths": lengths, "ys": ys},
                               {"best_score": best, "paths": tags_to_array(paths, T), "features": feats}, loss))

    # (iii) pyramidal windowed-attention segmenter ------------------------------------------
    torch.manual_seed(6)
    B, S, d, F, nl, nh, w = 2, 24, 32, 16, 2, 4, 4
    lengths = torch.tensor([24, 13])
    m = R.Transformer_segmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w)
    m.device = "cpu"
    with torch.no_grad():
        m.classification.weight.mul_(20.0)
    x = torch.randn(B, S, d, generator=g)
    y = labels(g, B, S, lengths.tolist(), -1.0)
    m.th = 0.5
    m.eval()
    scores, tags = m(x, lengths)
    hidden = m.model(x, lengths)
    loss = m.loss(x, lengths, y)
    d_ = pack(m, {"x": x, "lengths": lengths, "y": y, "th": 0.5, "nheads": nh, "window": w},
              {"scores": scores, "tags": tags_to_array(tags, S), "hidden": hidden}, loss)
    # the 30522-row word-embedding table and the *_global / pooler weights never reach the output
    for k in list(d_.keys()):
        if "word_embeddings" in k or "_global" in k or "pooler" in k:
            del d_[k]
    np.savez_compressed(os.path.join(HERE, "transformer_focal.npz"), **d_)

    # (vi) loss functions on raw vectors --------------------------------------------------
    z = torch.randn(257, generator=g) * 3
    yv = (torch.rand(257, generator=g) < 0.2).float()
    z.requires_grad_(True)
    fl = FL.sigmoid_focal_loss(alpha=0.9, gamma=2, reduction="mean")(z, yv)
    fl.backward()
    gfl = z.grad.clone()
    z.grad = None
    bce = nn.BCELoss()(torch.sigmoid(z), yv)
    bce.backward()
    np.savez_compressed(os.path.join(HERE, "losses.npz"), z=z.detach().numpy(), y=yv.numpy(),
                        focal=fl.detach().numpy(), focal_grad=gfl.numpy(),
                        bce=bce.detach().numpy(), bce_grad=z.grad.numpy())
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"  {f:32s} {os.path.getsize(os.path.join(HERE, f)):8d} B")


if __name__ == "__main__":
    main()
