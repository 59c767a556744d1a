"""CPU ORACLE (PyTorch, fp32) -- module-level restatement of the reference segmenters.
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs, never by the product package.

Why a torch twin next to oracle/ref_numpy.py: the reference IS a set of calls into torch
(`nn.LSTM`, `nn.Linear`, HF `LongformerModel`), so the faithful CPU timing baseline and the
autograd gradients come from the same library calls, arranged as the reference arranges them:

    Encoder           <- models/NeuralArchitectures.py:23-145 (RNN: pack -> nn.LSTM -> pad, TF-like init :58-79)
    ChainCRF          <- models/CRF.py:98-240
    Segmenter         <- models/CRF.py:274-369   (BiLSTM)
    LateFusion        <- models/CRF.py:371-479   (BiLSTMLateFusion)
    EncoderCRF        <- models/CRF.py:243-272   (BiRnnCrf, with the obviously intended wiring)
    WindowedSegmenter <- models/CRF.py:508-610 + models/RestrictedTransformerLayer.py:65-133
    RecurrentLongformer <- models/CRF.py:636-684, 764-858 + the FFN-less layer of models/longformer_noffn.py (source-less:
                         restated from the byte code of models/__pycache__/longformer_noffn.cpython-310.pyc; PARITY
                         UNPINNED against the reference class itself, which cannot be imported)

Parameter names and shapes are those of the reference state dict so golden
parameters load with `load_state_dict`.  Pinned by tests/test_oracle_golden.py against
tests/golden/*.npz (outputs of the unmodified reference).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

NEG = -1e4  # models/CRF.py:95


def length_mask(T, lengths):
    """Equivalent of create_mask / create_masks_huggingface (NeuralArchitectures.py:11-21) without the
    Python double loop: True where t < len_b."""
    return torch.arange(T)[None, :] < torch.as_tensor(lengths)[:, None]


def reference_mask_loop(max_len, lengths):
    """What create_masks_huggingface costs (RestrictedTransformerLayer.py:101-116): a Python loop over every (episode,
    position) comparing an int with a 0-d tensor.  Same values as length_mask; used only to TIME the reference's host
    step (bench.py reports the CPU baseline with and without it)."""
    mask = []
    for index in range(len(lengths)):
        mask.append([1 if pos < lengths[index] else 0 for pos in range(max_len)])
    return torch.tensor(mask)


class Encoder(nn.Module):
    def __init__(self, embed_size, hidden_size, num_layers=1, bidirectional=True):
        super().__init__()
        self.rnn = nn.LSTM(embed_size, hidden_size, num_layers=num_layers, batch_first=True,
                           bidirectional=bidirectional)
        for name, p in self.named_parameters():  # NeuralArchitectures.py:58-79
            if "weight_ih" in name:
                nn.init.xavier_uniform_(p.data)
            elif "weight_hh" in name:
                nn.init.orthogonal_(p.data)
            elif "bias_ih" in name:
                p.data.zero_()
                n = p.numel()
                p.data[n // 4: n // 2] = 1.0
            elif "bias_hh" in name:
                p.data.zero_()

    def forward(self, x, lengths):
        packed = pack_padded_sequence(x, torch.as_tensor(lengths).tolist(), batch_first=True, enforce_sorted=False)
        B = x.shape[0]
        nd = 2 if self.rnn.bidirectional else 1
        z = x.new_zeros(nd * self.rnn.num_layers, B, self.rnn.hidden_size)
        out, _ = self.rnn(packed, (z, z.clone()))
        return pad_packed_sequence(out, batch_first=True)[0]


def focal(z, y, alpha=0.9, gamma=2.0):
    """models/focal_loss.py:38-57, mean reduction."""
    p = torch.sigmoid(z)
    ce = F.binary_cross_entropy_with_logits(z, y, reduction="none")
    pt = p * y + (1 - p) * (1 - y)
    loss = ce * (1 - pt) ** gamma
    if alpha >= 0:
        loss = (alpha * y + (1 - alpha) * (1 - y)) * loss
    return loss.mean()


class _Head(nn.Module):
    """Shared loss / decode logic of the three sigmoid-or-softmax segmenters (CRF.py:319-369)."""

    def _init_head(self, in_features, tagset_size, loss_fn, threshold, alpha, gamma):
        if loss_fn not in ("CrossEntropy", "BinaryCrossEntropy", "FocalLoss"):
            raise ValueError("Choose one of CrossEntropy or BinaryCrossEntropy as loss function")
        self.loss_name = loss_fn
        self.bce = loss_fn != "CrossEntropy"
        self.tagset_size = tagset_size
        self.classification = nn.Linear(in_features, 1 if self.bce else tagset_size)
        self.th, self.alpha, self.gamma = threshold, alpha, gamma

    def _loss_from_logits(self, logits, lengths, tags):
        if self.bce:
            L = torch.as_tensor(lengths).tolist()
            z = torch.cat([logits[b, :n, 0] for b, n in enumerate(L)])
            y = torch.cat([tags[b, :n] for b, n in enumerate(L)])
            if self.loss_name == "FocalLoss":
                return focal(z, y, self.alpha, self.gamma)
            return F.binary_cross_entropy(torch.sigmoid(z), y)
        T = logits.shape[1]
        return F.cross_entropy(logits.reshape(-1, self.tagset_size), tags[:, :T].reshape(-1).long(), ignore_index=-1)

    def _decode(self, scores, lengths, threshold):
        if self.th is not None:
            threshold = self.th
        p = torch.sigmoid(scores)[:, :, 0] if self.bce else torch.softmax(scores, dim=2)[:, :, 1]
        tag = p > threshold
        return scores, [tag[b].tolist()[: int(n)] for b, n in enumerate(lengths)]


class Segmenter(_Head):
    def __init__(self, tagset_size, embedding_dim, hidden_dim, num_layers=1, loss_fn="CrossEntropy",
                 threshold=None, alpha=0.9, gamma=2):
        super().__init__()
        self.model = Encoder(embedding_dim, hidden_dim, num_layers)
        self._init_head(2 * hidden_dim, tagset_size, loss_fn, threshold, alpha, gamma)

    def loss(self, xs, lengths, tags, segments=None):
        return self._loss_from_logits(self.classification(self.model(xs, lengths)), lengths, tags)

    def forward(self, xs, lengths, threshold=0.4):
        return self._decode(self.classification(self.model(xs, lengths)), lengths, threshold)


class LateFusion(_Head):
    def __init__(self, tagset_size, embedding_dim, hidden_dim, num_layers=1, loss_fn="CrossEntropy",
                 threshold=None, alpha=0.9, gamma=2):
        super().__init__()
        self.model1 = Encoder(embedding_dim[0], hidden_dim, num_layers)
        self.model2 = Encoder(embedding_dim[1], hidden_dim, num_layers)
        self._init_head(4 * hidden_dim, tagset_size, loss_fn, threshold, alpha, gamma)

    def _features(self, x1, x2, lengths):
        return torch.cat((self.model1(x1, lengths), self.model2(x2, lengths)), dim=2)

    def loss(self, x1, x2, lengths, tags, segments=None):
        return self._loss_from_logits(self.classification(self._features(x1, x2, lengths)), lengths, tags)

    def forward(self, x1, x2, lengths, threshold=0.4):
        return self._decode(self.classification(self._features(x1, x2, lengths)), lengths, threshold)


class ChainCRF(nn.Module):
    """transitions[i, j] scores j -> i; two extra tags START = C-2, STOP = C-1 (CRF.py:105-117)."""

    def __init__(self, in_features, num_tags):
        super().__init__()
        self.num_tags = num_tags + 2
        self.start_idx, self.stop_idx = self.num_tags - 2, self.num_tags - 1
        self.fc = nn.Linear(in_features, self.num_tags)
        self.transitions = nn.Parameter(torch.randn(self.num_tags, self.num_tags))
        self.transitions.data[self.start_idx, :] = NEG
        self.transitions.data[:, self.stop_idx] = NEG

    @staticmethod
    def _lse(x):
        m = x.max(-1)[0]
        return m + (x - m.unsqueeze(-1)).exp().sum(-1).log()

    def partition(self, emis, mask):
        B, L, C = emis.shape
        s = emis.new_full((B, C), NEG)
        s[:, self.start_idx] = 0.0
        for t in range(L):
            nxt = self._lse(s.unsqueeze(1) + self.transitions.unsqueeze(0) + emis[:, t].unsqueeze(2))
            m = mask[:, t].unsqueeze(1)
            s = nxt * m + s * (1 - m)
        return self._lse(s + self.transitions[self.stop_idx])

    def gold(self, emis, tags, mask):
        B = emis.shape[0]
        e = emis.gather(2, tags.unsqueeze(-1)).squeeze(-1)
        seq = torch.cat([tags.new_full((B, 1), self.start_idx), tags], dim=1)
        tr = self.transitions[seq[:, 1:], seq[:, :-1]]
        last = seq.gather(1, mask.sum(1).long().unsqueeze(1)).squeeze(1)
        return ((tr + e) * mask).sum(1) + self.transitions[self.stop_idx, last]

    def loss(self, features, ys, masks):
        emis = self.fc(features)
        L = emis.size(1)
        m = masks[:, :L].float()
        return (self.partition(emis, m) - self.gold(emis, ys[:, :L].long(), m)).mean()

    def forward(self, features, masks):
        emis = self.fc(features)
        m = masks[:, : emis.size(1)].float()
        B, L, C = emis.shape
        bps = torch.zeros(B, L, C, dtype=torch.long)
        s = emis.new_full((B, C), NEG)
        s[:, self.start_idx] = 0.0
        for t in range(L):
            acc, bps[:, t] = (s.unsqueeze(1) + self.transitions).max(dim=-1)
            acc = acc + emis[:, t]
            mt = m[:, t].unsqueeze(1)
            s = acc * mt + s * (1 - mt)
        s = s + self.transitions[self.stop_idx]
        best, tag = s.max(dim=-1)
        bps = bps.numpy()
        paths = []
        for b in range(B):
            cur = int(tag[b])
            n = int(m[b].sum())
            path = [cur]
            for t in range(n - 1, -1, -1):
                cur = int(bps[b, t, cur])
                path.append(cur)
            paths.append(path[-2::-1])
        return best, paths


class EncoderCRF(nn.Module):
    def __init__(self, tagset_size, embedding_dim, hidden_dim, num_layers=1):
        super().__init__()
        self.model = Encoder(embedding_dim, hidden_dim, num_layers)
        self.crf = ChainCRF(2 * hidden_dim, tagset_size)

    def loss(self, xs, lengths, tags):
        feats = self.model(xs, lengths)
        return self.crf.loss(feats, tags, length_mask(xs.shape[1], lengths))

    def forward(self, xs, lengths):
        feats = self.model(xs, lengths)
        return self.crf(feats, length_mask(xs.shape[1], lengths))


class _Windowed(nn.Module):
    """models/RestrictedTransformerLayer.py:65-133: an HF LongformerModel fed `inputs_embeds`."""

    def __init__(self, d_model, nhead, n_layers, dim_feedforward, window_size, dropout, dropout_attention):
        super().__init__()
        from transformers import LongformerConfig, LongformerModel  # third-party, as in the reference

        cfg = LongformerConfig()
        cfg.attention_window = window_size
        cfg.hidden_dropout_prob = dropout
        cfg.num_hidden_layers = n_layers
        cfg.hidden_size = d_model
        cfg.intermediate_size = dim_feedforward
        cfg.num_attention_heads = nhead
        cfg.max_position_embeddings = 4096
        cfg.attention_probs_dropout_prob = dropout_attention
        self.configuration = cfg
        self.model = LongformerModel(cfg)

    def forward(self, x, lengths):
        mask = length_mask(x.shape[1], lengths).long()
        out = self.model(input_ids=None, inputs_embeds=x, attention_mask=mask,
                         global_attention_mask=torch.zeros_like(mask))
        return out.last_hidden_state


class WindowedSegmenter(_Head):
    def __init__(self, tagset_size, embedding_dim, hidden_dim, num_layers=6, nheads=8, loss_fn="CrossEntropy",
                 threshold=None, window_size=127, alpha=0.9, gamma=2, dropout_in=0.0, dropout_out=0.0):
        super().__init__()
        windows = [k * window_size for k in range(num_layers, 0, -1)]  # CRF.py:529
        self.model = _Windowed(embedding_dim, nheads, num_layers, hidden_dim, windows, dropout_in, dropout_out)
        self._init_head(embedding_dim, tagset_size, loss_fn, threshold, alpha, gamma)

    def loss(self, xs, lengths, tags):
        return self._loss_from_logits(self.classification(self.model(xs, lengths)), lengths, tags)

    def forward(self, xs, lengths, threshold=0.4):
        return self._decode(self.classification(self.model(xs, lengths)), lengths, threshold)


class _Redirect(nn.Module):
    """Stands in for `self.key` during one call: ignores the (already transposed) hidden states HF hands it and projects
    the external input instead -- longformer_noffn's `key_vectors = self.key(external_input.transpose(0, 1))`."""

    def __init__(self, linear, external):
        super().__init__()
        self.linear, self.external = linear, external

    def forward(self, _hidden_states):
        return self.linear(self.external.transpose(0, 1))


class NoffnLayer(nn.Module):
    """models/longformer_noffn.py LongformerLayer as its byte code reads: only `attention.self` exists; forward returns the
    sliding-window self-attention output itself -- no output dense, no residual, no LayerNorm, no feed-forward -- with
    query = Q(hidden), key = K(external_input) if given else K(hidden), value = V(hidden).  Everything after the three
    projections is HF's own LongformerSelfAttention.forward (the installed transformers), called unmodified."""

    def __init__(self, d_model, nhead, window):
        super().__init__()
        from transformers import LongformerConfig
        from transformers.models.longformer.modeling_longformer import LongformerSelfAttention

        cfg = LongformerConfig()
        cfg.attention_window = [window]
        cfg.hidden_size = d_model
        cfg.num_attention_heads = nhead
        cfg.attention_probs_dropout_prob = 0.0
        holder = nn.Module()
        holder.self = LongformerSelfAttention(cfg, layer_id=0)
        self.attention = holder

    def forward(self, hidden_states, lengths, external_input=None):
        import inspect

        sa = self.attention.self
        mask = length_mask(hidden_states.shape[1], lengths).long() - 1        # RestrictedTransformerLayer.py:126: 0 valid, -1 padded
        is_index_masked = mask < 0
        kw = dict(attention_mask=mask, is_index_masked=is_index_masked, is_index_global_attn=mask > 0, is_global_attn=False)
        if "layer_head_mask" in inspect.signature(sa.forward).parameters:
            kw["layer_head_mask"] = None
        key = sa._modules["key"]
        if external_input is not None:
            sa._modules["key"] = _Redirect(key, external_input)
        try:
            out = sa(hidden_states, **kw)
        finally:
            sa._modules["key"] = key
        return out[0]


class _RLBlock(nn.Module):
    """RecurrentLongformerBlock (CRF.py:636-684) with separate_forward_backward (forward states = queries / values,
    backward states = keys); the re-padding to 3600 rows is the caller's padded time axis."""

    def __init__(self, embedding_dim, hidden_dim, nheads, window, sep_fb=True):
        super().__init__()
        self.lstm = Encoder(embedding_dim, hidden_dim, 1)
        self.sep_fb = sep_fb
        holder = nn.Module()
        holder.model = NoffnLayer(hidden_dim if sep_fb else 2 * hidden_dim, nheads, window)
        self.transformer = holder

    def forward(self, x, lengths):
        S = x.shape[1]
        y = self.lstm(x, lengths)
        y = F.pad(y, (0, 0, 0, S - y.shape[1]))                       # pad_to_window_multiple (CRF.py:660-668)
        if self.sep_fb:
            bs, sl, _ = y.shape
            y = y.view(bs, sl, 2, -1)
            return self.transformer.model(y[:, :, 0, :], lengths, external_input=y[:, :, 1, :])
        return self.transformer.model(y, lengths)


class RecurrentLongformer(_Head):
    """CRF.py:764-858 (`BiLSTMRestrictedMHA`): num_layers blocks, a last bi-LSTM, the classification head."""

    def __init__(self, tagset_size, embedding_dim, hidden_dim, num_layers=6, nheads=8, loss_fn="CrossEntropy", threshold=None,
                 window_size=16, alpha=0.9, gamma=2, last_bilstm=True):
        super().__init__()
        blocks = [_RLBlock(embedding_dim, hidden_dim, nheads, window_size)]
        blocks += [_RLBlock(hidden_dim, hidden_dim, nheads, window_size) for _ in range(num_layers - 1)]
        self.model = nn.ModuleList(blocks)
        self.last_bilstm = last_bilstm
        if last_bilstm:
            self.model.append(Encoder(hidden_dim, hidden_dim, 1))
        self._init_head(2 * hidden_dim if last_bilstm else hidden_dim, tagset_size, loss_fn, threshold, alpha, gamma)

    def features(self, x, lengths):
        S = x.shape[1]
        for block in self.model:
            x = block(x, lengths)
        return F.pad(x, (0, 0, 0, S - x.shape[1]))

    def loss(self, x, lengths, tags):
        return self._loss_from_logits(self.classification(self.features(x, lengths)), lengths, tags)

    def forward(self, x, lengths, threshold=0.4):
        return self._decode(self.classification(self.features(x, lengths)), lengths, threshold)


def load_golden_params(module, fixture, strict=False):
    """Copy every 'p:<name>' array of a golden .npz into `module` (names = reference state-dict names)."""
    sd = {k[2:]: torch.from_numpy(fixture[k]) for k in fixture.files if k.startswith("p:")}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    if strict and (missing or unexpected):
        raise RuntimeError(f"missing={missing} unexpected={unexpected}")
    return missing, unexpected
