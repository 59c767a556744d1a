"""Golden fixture of the reference's OWN collaters and `TextSegmenter` steps (tests/golden/steps.npz).

Run in the development container only (it needs /root/reference):

    python tests/golden/make_golden_steps.py

`EncoderDataset.py` and `models/lightning_model.py` of the reference import `pytorch_lightning` and `segeval`, neither
of which exists in this image (nor, for segeval, anywhere offline).  They are imported here UNMODIFIED behind two
stubs that carry no arithmetic of the path:

  * `pytorch_lightning.LightningModule` = `torch.nn.Module` + `log` / `log_dict` that remember what was logged;
  * `segeval.pk` / `segeval.window_diff` RECORD the segment masses they are called with (that is the reference's
    whole contribution to Pk / WindowDiff: get_boundaries + the forced last boundary, lightning_model.py:16-55) and
    return the value of the oracle's restatement of segeval's published algorithm (oracle/ref_numpy.py) -- the
    algorithm itself stays "parity unpinned" (DESIGN.md section 2), everything around it is pinned by this fixture:
    which vectors reach it, the end-boundary handling, the AssertionError -> Pk substitution, the sklearn F1, the
    per-batch averaging and the keys of the logged dict.

Keys: "c<k>:<field>" collater outputs of case k; "<arch>:p:<name>" parameters; "<arch>:i:<name>" batch;
"<arch>:o:<name>" step outputs; "<arch>:masses" the recorded segeval arguments as a flat int array
(call count, then per call: kind (0 pk / 1 wd), len(h), h..., len(t), t...).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MTS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import ref_numpy as rn  # noqa: E402  (test infrastructure: the recording stub answers with it)

CALLS = []


def install_stubs():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.logged = {}

        def log(self, name, value, **k):
            self.logged[name] = value

        def log_dict(self, d, **k):
            self.logged.update(d)

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl

    sv = types.ModuleType("segeval")

    def pk(h, t, window_size=None):
        CALLS.append((0, list(h), list(t)))
        return rn.pk_masses(h, t, window_size)

    def window_diff(h, t, window_size=None):
        CALLS.append((1, list(h), list(t)))
        return rn.window_diff_masses(h, t, window_size)

    sv.pk, sv.window_diff = pk, window_diff
    sys.modules["segeval"] = sv

    noffn = types.ModuleType("models.longformer_noffn")

    class LongformerLayer(nn.Module):  # source-less in the upstream checkout; never instantiated here
        def __init__(self, *a, **k):
            super().__init__()

    noffn.LongformerLayer = LongformerLayer
    sys.modules["models.longformer_noffn"] = noffn


def flat_calls():
    out = [len(CALLS)]
    for kind, h, t in CALLS:
        out += [kind, len(h)] + [int(v) for v in h] + [len(t)] + [int(v) for v in t]
    return np.asarray(out, dtype=np.int64)


def episodes(g, sizes, d, d2=None, p=0.25):
    lines, second = [], []
    for i, n in enumerate(sizes):
        labs = (torch.rand(n, generator=g) < p).long().tolist()
        labs[-1] = 0
        name = f"{i}abc.npy" if i % 2 == 0 else f"x{i}.npy"   # leading digit = RadioNews (EncoderDataset.py:36-44)
        lines.append((torch.randn(n, d, generator=g), labs, name))
        if d2:
            second.append((torch.randn(n, d2, generator=g), None, None))
    return lines, second


def put(d, prefix, batch):
    for k, v in batch.items():
        if v is None:
            d[f"{prefix}:{k}:none"] = np.zeros(0)
        elif torch.is_tensor(v):
            d[f"{prefix}:{k}"] = v.numpy()
        else:
            d[f"{prefix}:{k}"] = np.asarray(v)


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import EncoderDataset as ref_ds  # noqa: E402
    import models.lightning_model as ref_lm  # noqa: E402

    d = {}
    g = torch.Generator().manual_seed(20261018)

    # ---- collaters (EncoderDataset.py:91-152, 189-232) ----------------------------------------------------------
    sizes = (5, 9, 2, 7)
    lines, second = episodes(g, sizes, 6, 3)
    for tensors in (lines, second):
        for k, ln in enumerate(tensors):
            d[f"cin:{'x' if tensors is lines else 'x2'}{k}"] = ln[0].numpy()
    for k, ln in enumerate(lines):
        d[f"cin:t{k}"] = np.asarray(ln[1])
    cases = [
        dict(CRF=True, truncate=False, second=True, domain_adapt=True),
        dict(CRF=False, truncate=False, second=True, domain_adapt=False),
        dict(CRF=False, truncate=True, truncate_value=4, second=False, domain_adapt=False),
        dict(CRF=True, truncate=True, truncate_value=12, second=True, domain_adapt=True),
    ]
    for ci, c in enumerate(cases):
        ds = ref_ds.AudioPortionDataset(lines, {"0": 0, "1": 1}, CRF=c["CRF"], truncate=c["truncate"],
                                        truncate_value=c.get("truncate_value", 100),
                                        second_input=second if c["second"] else None, domain_adapt=c["domain_adapt"])
        put(d, f"c{ci}", ds.collater([ds[i] for i in (2, 0, 3, 1)]))
    for ci, c in enumerate([dict(truncate=False), dict(truncate=True, truncate_value=4)], start=len(cases)):
        ds = ref_ds.AudioPortionDatasetInference([ln[0] for ln in lines], **c)
        put(d, f"c{ci}", ds.collater([ds[i] for i in (1, 3, 0, 2)]))

    # ---- TextSegmenter steps (models/lightning_model.py:179-760) ------------------------------------------------
    sizes = (30, 12, 21, 7, 3, 1)
    archs = [
        ("bilstm_pk", dict(architecture="BiLSTM", loss_fn="FocalLoss", metric="Pk", threshold=0.4), False),
        ("bilstm_f1_eb", dict(architecture="BiLSTM", loss_fn="BinaryCrossEntropy", metric="F1", threshold=0.5,
                              end_boundary=True), False),
        ("bilstm_ce_wd", dict(architecture="BiLSTM", loss_fn="CrossEntropy", metric="WD", threshold=None), False),
        # architecture "biLSTMCRF" cannot be run: the reference's BiRnnCrf.loss / forward unpack two values from an
        # encoder that returns one (models/CRF.py:263) -- ValueError on the first step
        ("late_pk", dict(architecture="BiLSTMLateFusion", loss_fn="FocalLoss", metric="Pk", threshold=0.3), True),
    ]
    for name, kw, double in archs:
        lines, second = episodes(g, sizes, 10, 6 if double else None, p=0.3)
        crf = kw["architecture"] == "biLSTMCRF"
        ds = ref_ds.AudioPortionDataset(lines, {"0": 0, "1": 1}, CRF=crf, truncate=False,
                                        second_input=second if double else None)
        batch = ds.collater([ds[i] for i in range(len(sizes))])
        torch.manual_seed(7)
        emb = [10, 6] if double else 10
        seg = ref_lm.TextSegmenter(2, emb, 8, num_layers=2, optimizer="Adam", lr=1e-3, all_results=True, **kw)
        seg.eval()  # dropout_in/out are 0: eval only fixes the mode
        for k, v in seg.state_dict().items():
            d[f"{name}:p:{k}"] = v.numpy()
        put(d, f"{name}:i", batch)
        seg.train()
        loss = seg.training_step(batch, 0)
        d[f"{name}:o:training_loss"] = loss.detach().numpy()
        d[f"{name}:o:logged_training_loss"] = seg.logged["training_loss"].detach().numpy()
        seg.eval()
        with torch.no_grad():
            val = seg.validation_step(batch, 0)
            d[f"{name}:o:val_loss"] = val.numpy()
            d[f"{name}:o:val_threshold"] = np.asarray(seg.logged["threshold"])
            if not double:
                tags = seg.predict_step(batch, 0)
                T = int(batch["src_lengths"].max())
                arr = -np.ones((len(tags), T), dtype=np.int8)
                for i, t in enumerate(tags):
                    arr[i, : len(t)] = np.asarray(t, dtype=np.int8)
                d[f"{name}:o:predict_tags"] = arr
            del CALLS[:]
            seg.logged = {}
            seg.test_step(batch, 0)
            res = seg.results[-1]
            keys = sorted(res)
            d[f"{name}:o:result_keys"] = np.asarray(keys)
            d[f"{name}:o:result_values"] = np.asarray([float(np.asarray(res[k]).reshape(-1)[0]) for k in keys], dtype=np.float64)
            d[f"{name}:masses"] = flat_calls()
        opt = seg.configure_optimizers()
        d[f"{name}:o:optimizer"] = np.asarray([type(opt["optimizer"]).__name__, opt["lr_scheduler"]["monitor"],
                                              opt["lr_scheduler"]["scheduler"].mode])
    out = os.path.join(HERE, "steps.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, len(d), "arrays")


if __name__ == "__main__":
    main()
