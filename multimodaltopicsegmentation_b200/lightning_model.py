"""`TextSegmenter` -- the task module the reference's train_fit.py / predict.py drive
(models/lightning_model.py:178-781), re-hosted on the B200 segmenters of modules.py.

Constructor signature, architecture names, batch-dict keys, logged metric names, threshold handling and
optimiser settings follow the reference.  When pytorch_lightning is importable the class is a real
LightningModule; otherwise a small stand-in base provides `log`, `log_dict` and `load_from_checkpoint` with the
PL checkpoint layout ({'state_dict': ...}) so that reference `.ckpt` files load either way.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .metrics import compute_Pk, compute_window_diff, f1_boundary
from .modules import BiLSTM, BiLSTMLateFusion, BiLSTMLateFusionCrf, BiRnnCrf

try:  # pragma: no cover - depends on the environment
    import pytorch_lightning as pl

    _Base = pl.LightningModule
except Exception:  # pytorch_lightning is not installed in the build image
    pl = None

    class _Base(nn.Module):
        """Minimal LightningModule stand-in (logging sink + checkpoint loader)."""

        def __init__(self):
            super().__init__()
            self.logged = {}

        def log(self, name, value, **kw):
            self.logged[name] = value

        def log_dict(self, d, **kw):
            self.logged.update(d)

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            model = cls(**kwargs)
            state = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
            model.load_state_dict(state, strict=strict)
            return model


class TextSegmenter(_Base):
    def __init__(self, tagset_size, embedding_dim, hidden_dim, num_layers=1, batch_first=True, LSTM=True,
                 bidirectional=True, architecture="biLSTMCRF", lr=0.01, dropout_in=0.0, dropout_out=0.0,
                 optimizer="SGD", positional_encoding=True, nheads=8, end_boundary=False, threshold=None,
                 search_threshold=False, metric="Pk", cosine_loss=False, zero_baseline=False, loss_fn="CrossEntropy",
                 no_validation=False, all_results=False, all_scores=False, alpha=0.9, gamma=2, attention_window=120,
                 switch="dense"):
        super().__init__()
        self.validation = not no_validation
        self.cos = cosine_loss
        self.double_input = False
        self.domain = False
        if architecture == "biLSTMCRF":
            self.cos = False
            self.model = BiRnnCrf(tagset_size, embedding_dim, hidden_dim, num_layers=num_layers,
                                  bidirectional=bidirectional, dropout_in=dropout_in, dropout_out=dropout_out,
                                  batch_first=batch_first, LSTM=LSTM, architecture="rnn")
        elif architecture == "BiLSTM":
            self.model = BiLSTM(tagset_size, embedding_dim, hidden_dim, num_layers=num_layers,
                                bidirectional=bidirectional, dropout_in=dropout_in, dropout_out=dropout_out,
                                batch_first=batch_first, LSTM=LSTM, loss_fn=loss_fn, threshold=threshold, alpha=alpha,
                                gamma=gamma)
        elif architecture == "BiLSTMLateFusion":
            self.model = BiLSTMLateFusion(tagset_size, embedding_dim, hidden_dim, num_layers=num_layers,
                                          bidirectional=bidirectional, dropout_in=dropout_in, dropout_out=dropout_out,
                                          batch_first=batch_first, LSTM=LSTM, loss_fn=loss_fn, threshold=threshold,
                                          alpha=alpha, gamma=gamma)
            self.double_input = True
        elif architecture == "BiLSTMLateFusionCRF":  # additive: BASELINE configs[3]'s dual encoder with the CRF output layer
            self.cos = False
            self.model = BiLSTMLateFusionCrf(tagset_size, embedding_dim, hidden_dim, num_layers=num_layers,
                                             bidirectional=bidirectional, dropout_in=dropout_in, dropout_out=dropout_out,
                                             batch_first=batch_first, LSTM=LSTM)
            self.double_input = True
        elif architecture == "Transformer":
            from .transformer import Transformer_segmenter

            self.model = Transformer_segmenter(tagset_size, embedding_dim, hidden_dim, num_layers=num_layers,
                                               dropout_in=dropout_in, dropout_out=dropout_out, batch_first=batch_first,
                                               loss_fn=loss_fn, positional_encoding=positional_encoding, nheads=nheads,
                                               threshold=threshold, alpha=alpha, gamma=gamma,
                                               window_size=attention_window)
        elif architecture == "BiLSTMRestrictedMHA":   # lightning_model.py:215-216
            from .recurrent_longformer import RecurrentLongformer

            self.model = RecurrentLongformer(tagset_size, embedding_dim, hidden_dim, num_layers=num_layers, dropout_in=dropout_in,
                                             dropout_out=dropout_out, batch_first=batch_first, loss_fn=loss_fn, nheads=nheads,
                                             threshold=threshold, alpha=alpha, gamma=gamma, window_size=attention_window)
        else:
            # SimpleBiLSTM, MLP, Transformer-CRF, RecurrentLongT5, SwitchBiLSTM, SheikhBiLSTM:
            # outside the hot path this package accelerates
            raise ValueError("No other architectures implemented yet")
        self.learning_rate = lr
        self.optimizer = optimizer
        self.eb = end_boundary
        self.threshold = threshold
        self.s_th = search_threshold
        self.metric = metric
        self.best_th, self.losses, self.targets = [], [], []
        self.zero_base = zero_baseline
        self.all = bool(all_results)
        self.results = []
        self.all_scores = bool(all_scores)
        self.scores = []
        # additive switch: evaluation counts on the device; False = the reference's host walk
        self.device_metrics = True

    def forward(self, x):
        return self.model(x)

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Checkpoints written under transformers 4.24 (the reference's pin) carry HF's non-parameter
        `embeddings.position_ids` buffer, which newer HF versions (and this package) do not register: tolerated
        (SURVEY.md section 10).  Everything else stays strict."""
        if any(k.endswith("embeddings.position_ids") for k in state_dict):
            state_dict = {k: v for k, v in state_dict.items() if not k.endswith("embeddings.position_ids")}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # ---- steps --------------------------------------------------------------------------------------------
    def _inputs(self, batch):
        sentence, lengths = batch["src_tokens"], batch["src_lengths"]
        if self.double_input:
            return (sentence, batch["src_tokens2"]), lengths
        return (sentence,), lengths

    def training_step(self, batch, batch_idx):
        if self.cos:
            raise NotImplementedError("the auxiliary cosine loss is outside the B200 hot path")
        xs, lengths = self._inputs(batch)
        self.best_th, self.losses, self.targets = [], [], []
        loss = self.model.loss(*xs, lengths, batch["tgt_tokens"])
        self.log("training_loss", loss, on_step=True, on_epoch=True, prog_bar=True, logger=True)
        return loss

    def validation_step(self, batch, batch_idx):
        xs, lengths = self._inputs(batch)
        if self.s_th:
            scores, _ = self.model(*xs, lengths)
            target = batch["tgt_tokens"]
            for index, score in enumerate(scores):
                n = int(lengths[index])
                self.losses.append(score[:n].detach().cpu().numpy())
                self.targets.append(target[index][:n].detach().cpu().numpy())
            return None
        with torch.no_grad():
            loss = self.model.loss(*xs, lengths, batch["tgt_tokens"])
        self.log_dict({"val_loss": loss, "threshold": 0.5})
        return loss

    def test_step(self, batch, batch_idx):
        xs, lengths = self._inputs(batch)
        target = batch["tgt_tokens"]
        if self.s_th:
            raise NotImplementedError()
        if self.metric.lower() in ("b", "scaiano"):
            raise NotImplementedError("B / WinPR evaluation is outside the hot path (SURVEY.md section 2 row 9)")
        lens_host = [int(v) for v in lengths]
        if self.device_metrics and not self.zero_base and torch.is_tensor(target) and target.is_cuda:
            return self._test_step_on_device(xs, lengths, target, lens_host)
        if self.zero_base:
            threshold = 0.4
            tags = [np.zeros(n) for n in lens_host]
            score = None
        else:
            threshold = self.threshold if self.threshold is not None else 0.4
            if not threshold:
                threshold = 0.5
            self.model.th = threshold
            score, tags = self.model(*xs, lengths)
        target_host = target.detach().cpu().numpy()
        loss_pk = loss_f1 = loss_wd = 0.0
        for i, tag in enumerate(tags):
            tgt = target_host[i][: lens_host[i]]
            if self.eb:
                tag[-1] = 0
                tgt[-1] = 0
            loss_pk += float(compute_Pk(np.array(tag), tgt))
            loss_f1 += f1_boundary(tgt.astype(int), np.array(tag).astype(int))
            try:
                loss_wd += float(compute_window_diff(np.array(tag), tgt))
            except AssertionError:
                loss_wd += float(compute_Pk(np.array(tag), tgt))
        n = len(target)
        results = {"Pk_loss": loss_pk / n, "F1_loss": loss_f1 / n, "WD_loss": loss_wd / n, "threshold": threshold}
        key = {"F1": "F1_loss", "WD": "WD_loss"}.get(self.metric, "Pk_loss")
        results["test_loss"] = results.pop(key)
        if self.all:
            self.results.append(results)
        if self.all_scores and score is not None:
            self.scores.extend([s.detach().cpu().numpy() for s in score])
        self.log_dict(results, on_epoch=True, prog_bar=True)
        return results

    def _test_step_on_device(self, xs, lengths, target, lens_host):
        """Same results as the host walk below, from per-episode integer counts made on the device
        (mts_seg_metrics): one [B, 8] int32 copy instead of a tag matrix + a Python/Decimal loop per episode."""
        from decimal import Decimal

        threshold = self.threshold if self.threshold is not None else 0.4
        if not threshold:
            threshold = 0.5
        self.model.th = threshold
        score, tags_dev, lens = self.model.decode_device(*xs, lengths)
        counts = ops.seg_metrics(tags_dev, target, lens, zero_last=self.eb).cpu().tolist()
        loss_pk = loss_f1 = loss_wd = 0.0
        for pk, wd, windows, _k, tp, fp, fn, _segs in counts:
            pk_val = float(Decimal(pk) / Decimal(windows)) if windows > 0 else 0.0
            loss_pk += pk_val
            loss_wd += float(Decimal(wd) / Decimal(windows)) if windows > 0 else pk_val  # segeval asserts -> Pk (:634-637)
            denom = 2 * tp + fp + fn
            loss_f1 += 2.0 * tp / denom if denom else 0.0
        n = len(counts)
        results = {"Pk_loss": loss_pk / n, "F1_loss": loss_f1 / n, "WD_loss": loss_wd / n, "threshold": threshold}
        key = {"F1": "F1_loss", "WD": "WD_loss"}.get(self.metric, "Pk_loss")
        results["test_loss"] = results.pop(key)
        if self.all:
            self.results.append(results)
        if self.all_scores and score is not None:
            self.scores.extend([s.detach().cpu().numpy() for s in score])
        self.log_dict(results, on_epoch=True, prog_bar=True)
        return results

    # ---- threshold search (lightning_model.py:436-553; dead code in the reference, whose hook is renamed) ------------
    def threshold_search(self, scores, target, lengths, thresholds=None):
        """The reference's sweep over np.arange(0.05, 1, 0.05) as ONE batched device evaluation (SURVEY.md section 8f
        row 2): all 19 thresholded tag matrices are scored by a single mts_seg_metrics launch.  `scores` [B, T, 1|2]
        device tensor as returned by the model (two columns: column 1 is compared, as the reference does at :462; one
        column: its sigmoid), `target` [B, T].  Selection follows the reference: strict improvement, first best wins;
        Pk / WD are minimised from 1, F1 maximised from -1.  Sets and returns (self.best_th, results dict)."""
        from decimal import Decimal

        if self.metric.lower() in ("b", "scaiano"):
            raise NotImplementedError("B / WinPR evaluation is outside the hot path (SURVEY.md section 2 row 9)")
        ths = np.arange(0.05, 1, 0.05) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
        lens = lengths if isinstance(lengths, ops.Lengths) else ops.Lengths(lengths, scores.device, scores.shape[1])
        B, n_th = lens.B, len(ths)
        with torch.no_grad():
            p = scores[:, : lens.T, 1] if scores.shape[2] > 1 else torch.sigmoid(scores[:, : lens.T, 0])
            th_dev = torch.tensor(ths, dtype=torch.float64, device=p.device).view(n_th, 1, 1)
            tags = (p.unsqueeze(0).double() > th_dev).to(torch.uint8).view(n_th * B, lens.T)  # np compares in float64 too
            tgt = target[:, : lens.T].float().unsqueeze(0).expand(n_th, B, lens.T).reshape(n_th * B, lens.T)
            many = ops.Lengths(lens.host * n_th, p.device, lens.T)
            counts = ops.seg_metrics(tags, tgt, many, zero_last=self.eb).view(n_th, B, 8).cpu().tolist()
        minimise = self.metric.lower() in ("pk", "wd")
        best, best_idx, results = (1 if minimise else -1), 0, []
        for index, th in enumerate(ths):
            loss_pk = loss_f1 = loss_wd = 0.0
            for pk, wd, windows, _k, tp, fp, fn, _segs in counts[index]:
                pk_val = float(Decimal(pk) / Decimal(windows)) if windows > 0 else 0.0
                loss_pk += pk_val
                loss_wd += float(Decimal(wd) / Decimal(windows)) if windows > 0 else pk_val
                denom = 2 * tp + fp + fn
                loss_f1 += 2.0 * tp / denom if denom else 0.0
            res = {"Pk_loss": loss_pk / B, "F1_loss": loss_f1 / B, "WD_loss": loss_wd / B}
            key = {"F1": "F1_loss", "WD": "WD_loss"}.get(self.metric, "Pk_loss")
            val = res.pop(key)
            res["valid_loss"] = val
            results.append(res)
            if (val < best) if (key != "F1_loss") else (val > best):
                best, best_idx, self.best_th = val, index, th
        chosen = results[best_idx]
        chosen["threshold"] = self.best_th if not isinstance(self.best_th, list) and self.best_th is not None else 0.4
        self.log_dict(chosen, on_epoch=True, prog_bar=True)
        return chosen["threshold"], chosen

    def predict_step(self, batch, batch_idx):
        xs, lengths = self._inputs(batch)
        _, tags = self.model(*xs, lengths)
        return tags

    def predict_batches(self, batches):
        """The loop `pl.Trainer.predict` runs over `predict_step` (the reference's predict.py), pipelined: the kernels
        of batch i+1 are enqueued before the tags of batch i are awaited, so the device does not idle while the host
        turns a tag matrix into Python lists.  Yields, in order, exactly what `predict_step` returns for each batch.
        Combine with `DevicePrefetcher` for host-resident batches."""
        from .modules import _host_tags_to_lists

        as_bool = not hasattr(self.model, "crf")   # Viterbi paths are int lists, thresholded tags bool lists
        ring, slot, pending = [None, None, None], 0, None   # pinned landing buffers for the tag matrices, reused

        def finish(item):
            host, ev, lens = item
            ev.synchronize()
            return _host_tags_to_lists(host.numpy(), lens, as_bool)

        for batch in batches:
            xs, lengths = self._inputs(batch)
            _, tags_dev, lens = self.model.decode_device(*xs, lengths)
            buf = ring[slot]
            if buf is None or buf.numel() < tags_dev.numel():
                buf = ring[slot] = torch.empty(tags_dev.numel(), dtype=torch.uint8, pin_memory=True)
            host = buf[: tags_dev.numel()].view(tags_dev.shape)
            host.copy_(tags_dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            slot = (slot + 1) % len(ring)
            if pending is not None:
                yield finish(pending)
            pending = (host, ev, lens)
        if pending is not None:
            yield finish(pending)

    # ---- optimisers (lightning_model.py:759-781) ---------------------------------------------------------------
    def configure_optimizers(self):
        if self.optimizer == "SGD":
            optimizer = torch.optim.SGD(self.parameters(), lr=self.learning_rate, weight_decay=1e-4, momentum=0.9)
        else:
            # same optimiser and hyper-parameters as the reference (lightning_model.py:764); on CUDA parameters torch's FUSED
            # implementation of it (one multi-tensor kernel instead of ~12 foreach launches per step: 0.29 -> 0.05 ms of a
            # 3.2 ms configs[3] step); MTS_FUSED_ADAM=0 keeps torch's default choice
            params = list(self.parameters())
            fused = __import__("os").environ.get("MTS_FUSED_ADAM", "1") != "0" and len(params) > 0 and all(p.is_cuda for p in params)
            optimizer = torch.optim.Adam(params, eps=1e-7, lr=self.learning_rate, **({"fused": True} if fused else {}))
        mode = "min" if (self.metric.lower() in ("pk", "wd") or not self.s_th) else "max"
        monitor = "val_loss" if self.validation else "training_loss"
        scheduler = {"scheduler": torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode, factor=0.8, patience=10),
                     "monitor": monitor}
        return {"optimizer": optimizer, "lr_scheduler": scheduler}


def launch_count():
    return ops.launch_count()
