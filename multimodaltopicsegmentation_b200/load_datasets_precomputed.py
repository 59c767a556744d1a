"""Loader of pre-computed per-sentence embeddings -- the step immediately before the hot path (SURVEY.md section 8f
row 1).  Mirrors the behaviour of the reference's `utils/load_datasets_precomputed.py`:

    load_dataset_from_precomputed   <- :103-210   folder(s) of <episode>.npy + labs_dict.pkl [+ split json, timings pkl]
    cross_validation_split          <- :56-100    (the `inverse_augmentation` branch is never reached from the loader,
                                                   which hard-codes False at :203; it is not implemented here)
    load_dataset_for_inference      <- :212-224

Same arguments, same return structure (`[[train, test, validation]]` for a standard split -- note the reference's
order -- or `k_folds` x `[train, test]`), same quirks: early fusion is the concatenation of the modality folders named
in `embedding_directory` joined by '+', the last label of every episode is forced to 0, episodes without labels are
skipped, a fixed list of over-long podcast episodes is dropped, split lists are consumed from their END
(`list.pop()`), one entry per file found in the first folder.

Additive (off by default): `keep_modalities=True` returns each episode's embeddings as a TUPLE of per-modality
tensors instead of their concatenation, so that the early-fusion concat can happen inside the operand-packing kernel
(`BiLSTM((x_text, x_audio), lengths)`); `ResidentDataset` then keeps whole datasets on the GPU and builds padded
batches there (mts_gather_pad), removing the per-batch host padding + H2D copy of the reference's collater.
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np
import torch

from . import ops

_SKIPPED_PODCASTS = ("24580", "25539", "25684", "26071", "26214", "26321", "26427")  # reference :141


def cross_validation_split(dataset, num_folds=5, n_test_folds=1, inverse_augmentation=True):
    """Contiguous test blocks of len(dataset) // num_folds episodes; the rest trains (reference :56-100)."""
    if inverse_augmentation:
        raise NotImplementedError("inverse augmentation is not reachable from load_dataset_from_precomputed "
                                  "(the reference passes False at :203) and is not implemented")
    unit = len(dataset) // num_folds
    test_size = unit * n_test_folds
    folds = []
    for i in range(num_folds):
        lo, hi = i * unit, i * unit + test_size
        test = dataset[lo:hi]
        if i == num_folds + 1 - n_test_folds:  # wrap-around case of the reference (only reachable for n_test_folds >= 2)
            wrap = test_size // n_test_folds
            test = test + dataset[:wrap]
            train = dataset[wrap:-wrap]
        else:
            train = dataset[:lo] + dataset[hi:]
        folds.append([train, test])
    return folds


def _mask_inner_sentences(embs, labels, mask_probability):
    """Reference :171-183: with a fixed numpy seed, drop non-boundary sentences with probability 1 - mask_probability."""
    np.random.seed(1)
    rows = embs.tolist()
    popped = 0
    for index in range(len(embs)):
        if np.random.rand() > mask_probability and not labels[index - popped]:
            rows.pop(index - popped)
            labels.pop(index - popped)
            popped += 1
    return torch.tensor(rows)


def load_dataset_from_precomputed(embedding_directory, lab_file, delete_last_sentence=False,
                                  compute_confidence_intervals=False, inverse_augmentation=False, umap_project=False,
                                  k_folds=5, mask_inner_sentences=False, mask_probability=0.9, split=None,
                                  timing_info=None, keep_modalities=False):
    standard_split = split is not None
    if standard_split:
        with open(split) as f:
            split = json.load(f)
        data = [[], [], []]  # train, TEST, validation -- the reference's order
    else:
        data = []
    original = []
    with open(lab_file, "rb") as f:
        labs = pickle.load(f)
    assert isinstance(labs, dict)
    times = None
    if timing_info is not None:
        with open(timing_info, "rb") as f:
            times = pickle.load(f)
    roots = embedding_directory.split("+")
    if keep_modalities and mask_inner_sentences:
        raise ValueError("keep_modalities cannot be combined with mask_inner_sentences")

    for file in os.listdir(roots[0]):
        if file[-16:] == ":Zone.Identifier" or file[:-4] in _SKIPPED_PODCASTS:
            continue
        part = None
        if standard_split:  # one split entry is consumed per directory entry, train first, then test, then validation
            if len(split["train"]):
                file, part = split["train"].pop(), 0
            elif len(split["test"]):
                file, part = split["test"].pop(), 1
            else:
                file, part = split["validation"].pop(), 2
        mods = [torch.from_numpy(np.load(os.path.join(root, file)).squeeze()) for root in roots]
        name = file[:-4]
        if times is not None:
            mods.append(torch.tensor(times[name]))
        if len(labs[name]) < 1:
            print("Warning: {} has no data".format(name))
            continue
        labs[name][-1] = 0
        if keep_modalities and len(mods) > 2:
            # the device path takes a (first, rest) pair: modalities beyond the second (a third '+' directory, the timing
            # features) are concatenated onto the second one, which keeps the column order of the reference's torch.cat
            mods = [mods[0], torch.cat([m.float() for m in mods[1:]], axis=-1)]
        embs = tuple(mods) if keep_modalities else torch.cat(mods, axis=-1)
        if mask_inner_sentences:
            original.append((embs, labs[name].copy(), file))
            embs = _mask_inner_sentences(embs, labs[name], mask_probability)
        if sum(labs[name]) < 1:
            print("Warning: {} has no positive topic boundaries".format(name))
        (data[part] if standard_split else data).append((embs, labs[name], file))

    if standard_split:
        return [data]
    folds = cross_validation_split(data, num_folds=k_folds, inverse_augmentation=False)
    if mask_inner_sentences:  # the reference restores ONE original episode as the test set of fold i
        for index in range(len(folds)):
            folds[index][1] = [original[index]]
    return folds


def load_dataset_for_inference(embedding_directory):
    files = os.listdir(embedding_directory)
    data = [torch.from_numpy(np.load(os.path.join(embedding_directory, f)).squeeze()) for f in os.listdir(embedding_directory)]
    return data, files


class ResidentDataset:
    """A whole split kept on the GPU: per modality one [sum(len), D_m] tensor plus episode offsets, labels likewise.
    `batch(ids)` builds the reference collater's batch dict (EncoderDataset.py:91-152: zero-padded to the batch
    maximum, tags padded with -1, or 0 for CRF architectures) entirely on the device with one gather kernel per
    tensor -- no host padding loop, no per-step H2D copy of embeddings.  `src_tokens` is a (text, audio, ...) tuple
    when the episodes were loaded with keep_modalities=True and a single tensor otherwise."""

    def __init__(self, episodes, device, CRF=False):
        self.device = torch.device(device)
        self.pad_tag = 0.0 if CRF else -1.0
        first = episodes[0][0]
        self.multi = isinstance(first, (tuple, list))
        n_mod = len(first) if self.multi else 1
        lengths = [len(e[0][0]) if self.multi else len(e[0]) for e in episodes]
        self.lengths = torch.tensor(lengths, dtype=torch.long)
        offs = torch.zeros(len(episodes) + 1, dtype=torch.int64)
        offs[1:] = torch.cumsum(self.lengths, 0)
        self.offsets = offs.to(self.device)
        self.lengths_dev = self.lengths.to(torch.int32).to(self.device)
        self.tokens = []
        for m in range(n_mod):
            cat = torch.cat([(e[0][m] if self.multi else e[0]).float().reshape(lengths[i], -1) for i, e in enumerate(episodes)])
            self.tokens.append(cat.to(self.device).contiguous())
        self.tags = torch.cat([torch.as_tensor(e[1], dtype=torch.float32) for e in episodes]).to(self.device).contiguous()
        self.names = [e[2] if len(e) > 2 else i for i, e in enumerate(episodes)]

    def __len__(self):
        return len(self.lengths)

    def batch(self, ids):
        ids = [int(i) for i in ids]
        idx = torch.tensor(ids, dtype=torch.int32)
        lens = self.lengths[idx.long()]
        T = int(lens.max())
        B = len(ids)
        idx_dev = idx.to(self.device, non_blocking=True)
        outs = []
        for tok in self.tokens:
            D = tok.shape[1]
            out = torch.empty((B, T, D), device=self.device, dtype=torch.float32)
            ops._call("mts_gather_pad", ops._ptr(tok), ops._ptr(self.offsets), ops._ptr(self.lengths_dev), ops._ptr(idx_dev), B,
                      T, D, 0.0, ops._ptr(out), ops._stream())
            outs.append(out)
        tags = torch.empty((B, T), device=self.device, dtype=torch.float32)
        ops._call("mts_gather_pad", ops._ptr(self.tags), ops._ptr(self.offsets), ops._ptr(self.lengths_dev), ops._ptr(idx_dev), B, T,
                  1, self.pad_tag, ops._ptr(tags), ops._stream())
        return {"id": torch.tensor(ids), "src_tokens": tuple(outs) if self.multi else outs[0], "src_lengths": lens,
                "tgt_tokens": tags, "src_tokens2": None, "domain": None}
