"""CPU tests of the host-side mirror: batch layout, metrics, length handling, state-dict compatibility."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_numpy as rn, ref_torch as rt


def test_collater_layout_matches_reference_contract():
    from multimodaltopicsegmentation_b200 import AudioPortionDataset, AudioPortionDatasetInference

    lines = [(torch.randn(n, 6), [0] * (n - 1) + [1], f"{n}.npy") for n in (5, 9, 2)]
    second = [(torch.randn(n, 3), None, None) for n in (5, 9, 2)]
    for crf, pad in ((True, 0.0), (False, -1.0)):
        ds = AudioPortionDataset(lines, {}, CRF=crf, truncate=False, second_input=second, domain_adapt=True)
        batch = ds.collater([ds[i] for i in range(3)])
        assert set(batch) == {"id", "src_tokens", "src_lengths", "tgt_tokens", "src_tokens2", "domain"}
        assert batch["src_tokens"].shape == (3, 9, 6) and batch["src_tokens2"].shape == (3, 9, 3)
        assert batch["src_lengths"].tolist() == [5, 9, 2] and batch["src_lengths"].dtype == torch.int64
        assert batch["tgt_tokens"].dtype == torch.float32
        assert batch["tgt_tokens"][2, 2:].eq(pad).all() and batch["tgt_tokens"][2, 1] == 1
        assert not batch["src_tokens"][0, 5:].any()
        assert batch["domain"] == [1, 1, 1]
    ds = AudioPortionDataset(lines, {}, CRF=False, truncate=True, truncate_value=4)
    batch = ds.collater([ds[i] for i in range(3)])
    assert batch["src_tokens"].shape == (3, 4, 6) and batch["src_lengths"].tolist() == [4, 4, 2]
    assert batch["src_tokens2"] is None and batch["domain"] is None
    assert ds.collater([]) == {}
    inf = AudioPortionDatasetInference([l[0] for l in lines], truncate=False)
    b2 = inf.collater([inf[i] for i in range(3)])
    assert set(b2) == {"id", "src_tokens", "src_lengths"} and b2["src_lengths"].tolist() == [5, 9, 2]


def test_collaters_equal_the_reference_collaters(golden):
    """Outputs of the UNMODIFIED reference collaters (EncoderDataset.py:91-152, 189-232; tests/golden/make_golden_steps.py
    imports them behind a pytorch_lightning / segeval stub) on seeded ragged episodes: every field, dtype and shape."""
    from multimodaltopicsegmentation_b200 import AudioPortionDataset, AudioPortionDatasetInference

    fx = golden("steps")
    n = 4
    lines = [(torch.from_numpy(fx[f"cin:x{k}"]), fx[f"cin:t{k}"].tolist(), f"{k}abc.npy" if k % 2 == 0 else f"x{k}.npy")
             for k in range(n)]
    second = [(torch.from_numpy(fx[f"cin:x2{k}"]), None, None) for k in range(n)]

    def check(ci, batch):
        fields = {k.split(":")[1] for k in fx.files if k.startswith(f"c{ci}:")}
        assert set(batch) == fields
        for k, v in batch.items():
            if f"c{ci}:{k}:none" in fx.files:
                assert v is None, (ci, k)
                continue
            ref = fx[f"c{ci}:{k}"]
            if torch.is_tensor(v):
                assert v.numpy().dtype == ref.dtype and tuple(v.shape) == ref.shape, (ci, k, v.dtype, ref.dtype)
                assert np.array_equal(v.numpy(), ref), (ci, k)
            else:
                assert list(v) == ref.tolist(), (ci, k)

    cases = [dict(CRF=True, truncate=False, second=True, domain_adapt=True),
             dict(CRF=False, truncate=False, second=True, domain_adapt=False),
             dict(CRF=False, truncate=True, truncate_value=4, second=False, domain_adapt=False),
             dict(CRF=True, truncate=True, truncate_value=12, second=True, domain_adapt=True)]
    for ci, c in enumerate(cases):
        ds = AudioPortionDataset(lines, {"0": 0, "1": 1}, CRF=c["CRF"], truncate=c["truncate"],
                                 truncate_value=c.get("truncate_value", 100), second_input=second if c["second"] else None,
                                 domain_adapt=c["domain_adapt"])
        check(ci, ds.collater([ds[i] for i in (2, 0, 3, 1)]))
    for ci, c in enumerate([dict(truncate=False), dict(truncate=True, truncate_value=4)], start=len(cases)):
        ds = AudioPortionDatasetInference([ln[0] for ln in lines], **c)
        check(ci, ds.collater([ds[i] for i in (1, 3, 0, 2)]))


def _recorded_segeval_calls(flat):
    flat = [int(v) for v in flat]
    n, i, calls = flat[0], 1, []
    for _ in range(n):
        kind, nh = flat[i], flat[i + 1]
        h = flat[i + 2: i + 2 + nh]
        nt = flat[i + 2 + nh]
        t = flat[i + 3 + nh: i + 3 + nh + nt]
        calls.append((kind, h, t))
        i += 3 + nh + nt
    assert i == len(flat)
    return calls


def test_metrics_on_the_masses_the_reference_hands_to_segeval(golden):
    """The reference's test_step (run unmodified for the fixture) reaches segeval with these (hypothesis, reference)
    mass lists; the host metrics of the product, fed the equivalent boundary vectors, must return what the oracle's
    segeval restatement returns on the masses -- including the window > length cases where WindowDiff asserts."""
    from multimodaltopicsegmentation_b200 import metrics

    fx = golden("steps")
    n_calls = 0
    for name in ("bilstm_pk", "bilstm_f1_eb", "bilstm_ce_wd", "late_pk"):
        for kind, h, t in _recorded_segeval_calls(fx[name + ":masses"]):
            assert sum(h) == sum(t)
            hb = np.zeros(sum(h), dtype=int)
            tb = np.zeros(sum(t), dtype=int)
            hb[np.cumsum(h) - 1] = 1
            tb[np.cumsum(t) - 1] = 1
            assert metrics.get_boundaries(hb) == h and metrics.get_boundaries(tb) == t
            hb[-1] = tb[-1] = 0   # compute_Pk / compute_window_diff force the last boundary themselves
            if kind == 0:
                assert metrics.compute_Pk(hb, tb) == rn.pk_masses(h, t)
            else:
                try:
                    want = rn.window_diff_masses(h, t)
                except AssertionError:
                    with pytest.raises(AssertionError):
                        metrics.compute_window_diff(hb, tb)
                else:
                    assert metrics.compute_window_diff(hb, tb) == want
            n_calls += 1
    assert n_calls >= 40


def test_metrics_agree_with_oracle():
    from multimodaltopicsegmentation_b200 import compute_Pk, compute_window_diff, get_boundaries

    rng = np.random.default_rng(3)
    for n in (2, 3, 5, 17, 64, 301, 2437):
        for _ in range(25):
            ref = (rng.random(n) < 0.1).astype(np.uint8)
            hyp = (rng.random(n) < 0.1).astype(np.uint8)
            ref[-1] = hyp[-1] = 0
            assert compute_Pk(hyp, ref) == rn.compute_pk(hyp, ref)
            num, den = c_oracle.pk(hyp, ref)
            assert float(compute_Pk(hyp, ref)) == (num / den if den else 0.0)
            try:
                want = rn.compute_window_diff(hyp, ref)
            except AssertionError:
                with pytest.raises(AssertionError):
                    compute_window_diff(hyp, ref)
                continue
            assert compute_window_diff(hyp, ref) == want
            closed = hyp.copy()
            closed[-1] = 1
            assert get_boundaries(closed) == rn.get_boundaries(closed)
    # inputs are not mutated (the reference flips boundaries[-1] in place and restores it)
    a = np.array([0, 1, 0, 0]); b = np.array([0, 0, 1, 0])
    compute_Pk(a, b)
    assert a.tolist() == [0, 1, 0, 0] and b.tolist() == [0, 0, 1, 0]


def test_state_dict_keys_match_reference_names(golden):
    """Reference checkpoints load: same key names and shapes as the golden state dicts."""
    from multimodaltopicsegmentation_b200 import BiLSTM, BiLSTMLateFusion, BiRnnCrf, TextSegmenter

    for fx_name, mod in (("bilstm_focalloss", BiLSTM(2, 12, 8, num_layers=2, loss_fn="FocalLoss")),
                         ("bilstm_crossentropy", BiLSTM(2, 12, 8, num_layers=2, loss_fn="CrossEntropy")),
                         ("latefusion_focal", BiLSTMLateFusion(2, [5, 7], 8, num_layers=2, loss_fn="FocalLoss")),
                         ("bilstm_crf", BiRnnCrf(2, 12, 8, num_layers=2))):
        fx = golden(fx_name)
        want = {k[2:]: fx[k].shape for k in fx.files if k.startswith("p:")}
        got = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
        assert got == want, fx_name
    seg = TextSegmenter(2, 896, 256, num_layers=2, architecture="BiLSTM", loss_fn="FocalLoss")
    keys = set(seg.state_dict())
    assert "model.model.rnn.weight_ih_l0_reverse" in keys and "model.classification.weight" in keys
    assert seg.state_dict()["model.model.rnn.weight_ih_l1"].shape == (1024, 512)


def test_init_matches_reference_initialisers():
    """Same RNG stream and initialisers as NeuralArchitectures.py:58-79 => same weights under a seed as the oracle twin."""
    from multimodaltopicsegmentation_b200 import BiLSTM

    torch.manual_seed(7)
    a = BiLSTM(2, 12, 8, num_layers=2, loss_fn="FocalLoss")
    torch.manual_seed(7)
    b = rt.Segmenter(2, 12, 8, num_layers=2, loss_fn="FocalLoss")
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    bias = a.state_dict()["model.rnn.bias_ih_l0"]
    assert bias[8:16].eq(1).all() and bias[:8].eq(0).all()


def test_lengths_ordering_and_validation():
    from multimodaltopicsegmentation_b200 import ops

    lens = ops.Lengths(torch.tensor([3, 9, 1, 9, 4]), "cpu", 12)
    assert (lens.B, lens.T, lens.N) == (5, 9, 26)
    assert lens.dev.tolist() == [3, 9, 1, 9, 4] and lens.dev.dtype == torch.int32
    assert [lens.host[i] for i in lens.order.tolist()] == [9, 9, 4, 3, 1]
    with pytest.raises(ValueError):
        ops.Lengths([0, 2], "cpu", 4)
    with pytest.raises(ValueError):
        ops.Lengths([5], "cpu", 4)


def test_error_conventions():
    from multimodaltopicsegmentation_b200 import BiLSTM, TextSegmenter

    with pytest.raises(ValueError):
        BiLSTM(2, 4, 4, loss_fn="Hinge")
    with pytest.raises(ValueError):
        TextSegmenter(2, 4, 4, architecture="SheikhBiLSTM")
    seg = TextSegmenter(2, 4, 4, architecture="BiLSTM", loss_fn="FocalLoss", search_threshold=True)
    with pytest.raises(NotImplementedError):
        seg.test_step({"src_tokens": torch.zeros(1, 2, 4), "src_lengths": torch.tensor([2]), "tgt_tokens": torch.zeros(1, 2)}, 0)
    opt = TextSegmenter(2, 4, 4, architecture="BiLSTM", loss_fn="FocalLoss", optimizer="Adam", lr=1e-3).configure_optimizers()
    assert opt["optimizer"].defaults["eps"] == 1e-7 and opt["lr_scheduler"]["monitor"] == "val_loss"


def test_checkpoint_round_trip_like_predict_py(golden, tmp_path):
    """SURVEY.md section 8f row 4 (wire format): a PL-style .ckpt ({'state_dict': ...} with the reference's key names,
    here the golden parameters of the unmodified reference) loads through TextSegmenter.load_from_checkpoint with the
    keyword arguments predict.py:228-241 passes, key for key; a checkpoint saved from our module loads back; the HF
    4.24 `position_ids` buffer of old transformer checkpoints is tolerated."""
    from multimodaltopicsegmentation_b200 import TextSegmenter

    fx = golden("bilstm_binarycrossentropy")
    ref_params = {"model." + k[2:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("p:")}
    path = tmp_path / "best.ckpt"
    torch.save({"state_dict": ref_params, "epoch": 3, "global_step": 120, "optimizer_states": [], "lr_schedulers": []}, path)
    seg = TextSegmenter.load_from_checkpoint(str(path), architecture="BiLSTM", tagset_size=2, embedding_dim=12, hidden_dim=8,
                                             bidirectional=True, lr=1e-3, num_layers=2, loss_fn="BinaryCrossEntropy",
                                             dropout_in=0.0, dropout_out=0.0, threshold=0.5)
    got = seg.state_dict()
    assert set(got) == set(ref_params)
    for k, v in ref_params.items():
        assert torch.equal(got[k], v), k
    # our own save -> load
    path2 = tmp_path / "ours.ckpt"
    torch.save({"state_dict": seg.state_dict()}, path2)
    again = TextSegmenter.load_from_checkpoint(str(path2), architecture="BiLSTM", tagset_size=2, embedding_dim=12,
                                               hidden_dim=8, num_layers=2, loss_fn="BinaryCrossEntropy")
    for k, v in seg.state_dict().items():
        assert torch.equal(again.state_dict()[k], v), k
    # a mismatching head (CrossEntropy checkpoints have a 2-row classifier) must fail loudly, as predict.py expects
    with pytest.raises((RuntimeError, KeyError)):
        TextSegmenter.load_from_checkpoint(str(path), architecture="BiLSTM", tagset_size=2, embedding_dim=12, hidden_dim=8,
                                           num_layers=2, loss_fn="CrossEntropy")
    # transformer checkpoint written under transformers 4.24: extra position_ids buffer
    tr = TextSegmenter(2, 32, 16, num_layers=2, architecture="Transformer", loss_fn="FocalLoss", nheads=4, attention_window=4)
    sd = dict(tr.state_dict())
    sd["model.model.model.embeddings.position_ids"] = torch.arange(4096).unsqueeze(0)
    path3 = tmp_path / "xf.ckpt"
    torch.save({"state_dict": sd}, path3)
    back = TextSegmenter.load_from_checkpoint(str(path3), architecture="Transformer", tagset_size=2, embedding_dim=32,
                                              hidden_dim=16, num_layers=2, loss_fn="FocalLoss", nheads=4, attention_window=4)
    for k, v in tr.state_dict().items():
        assert torch.equal(back.state_dict()[k], v), k


def test_reference_mask_loop_equals_vectorised_mask():
    """oracle/ref_torch.reference_mask_loop restates create_masks_huggingface's Python loop (timed by bench.py); it must
    produce the mask the oracle's vectorised length_mask produces."""
    lengths = torch.tensor([7, 1, 12, 5])
    m = rt.reference_mask_loop(12, lengths)
    assert m.shape == (4, 12) and m.dtype == torch.int64
    assert torch.equal(m.bool(), rt.length_mask(12, lengths))


def test_result_files_round_trip(tmp_path):
    """results.txt / logs / all_results.json / all_scores.json in the reference's formats
    (train_fit.py:399-412, 436-451, 583-624) and the hyper-parameter parse predict.py:168-177 performs on the first."""
    from multimodaltopicsegmentation_b200 import results_io as rio

    best = {"Pk": 0.25, "F1": 0.5, "WD": 0.375}
    lines = rio.summary_lines("exp1", "x-vectors", "BiLSTM", 64, 256, 0.2, 0.1, 2, "Adam", best)
    path = rio.write_results_txt(str(tmp_path), lines)
    text = open(path).read()
    assert text == ("\nResults for experiment exp1 with following parameters:\n\nSentence encoder: x-vectors\n"
                    "\nNeural architecture: BiLSTM\n\nBatch size: 64\n\nHidden units: 256\n\nDropout in: 0.2\n"
                    "\nDropout out: 0.1\n\nNumber of layers: 2\n\nOptimizer: Adam\n\nMean Pk obtained is 0.25\n"
                    "\nMean F1 obtained is 0.5\n\nMean WD obtained is 0.375\n")
    hp = rio.read_hyperparameters(path)
    assert (hp.encoder, hp.architecture, hp.hidden_units, hp.num_layers) == ("x-vectors", "BiLSTM", 256, 2)

    conf = {"Pk": 0.01, "F1": 0.02, "WD": 0.03, "B": 0.04}
    cv = rio.summary_lines("e", "enc", "Transformer", 8, 32, 0.0, 0.0, 1, "SGD", dict(best, B=0.7), confidence=conf, metric="B",
                           zero_shot_labels=["a", "b"])
    assert cv[9] == "Mean Precision obtained is 0.25 with a 95% confidence interval of +- 0.01"
    assert cv[11] == "Mean Recall obtained is 0.375 with a 95% confidence interval of +- 0.03"
    assert cv[12] == "Mean Boundary Similarity obtained is 0.7 with a 95% confidence interval of +- 0.04"
    assert cv[13] == "Labels: ['a', 'b']"
    with pytest.raises(ValueError):
        rio.read_hyperparameters(rio.write_results_txt(str(tmp_path), cv[:3], name="short.txt"))

    # the tuned metric travels as 'test_loss' (train_fit.py:375-397)
    logged_f1 = {"Pk_loss": 0.2, "WD_loss": 0.3, "test_loss": 0.9, "threshold": 0.4}
    logged_pk = {"F1_loss": 0.9, "WD_loss": 0.3, "test_loss": 0.2, "threshold": 0.4}
    assert rio.fold_metrics(logged_f1, "F1") == rio.fold_metrics(logged_pk, "Pk") == {"Pk": 0.2, "WD": 0.3, "F1": 0.9}
    rio.append_fold_log(str(tmp_path), 0, logged_f1, "F1")
    rio.append_fold_log(str(tmp_path), 1, logged_pk, "Pk")
    assert open(tmp_path / "logs").read() == ("Results for fold number 0\nPK score: 0.2\nWD score: 0.3\nF1 score: 0.9\n"
                                              "Results for fold number 1\nPK score: 0.2\nWD score: 0.3\nF1 score: 0.9\n")
    assert rio.read_fold_logs(str(tmp_path / "logs")) == [(0, {"PK": 0.2, "WD": 0.3, "F1": 0.9}), (1, {"PK": 0.2, "WD": 0.3, "F1": 0.9})]
    rio.append_fold_log(str(tmp_path), 0, {"b_precision": 1.0, "b_recall": 0.5, "b_f1": 0.6, "test_loss": 0.4}, "B", name="logs_b")
    assert open(tmp_path / "logs_b").read().splitlines()[1:] == ["B_precision score: 1.0", "B_recall score: 0.5", "B_F1 score: 0.6",
                                                                 "B Similarity score: 0.4"]

    files = [(None, None, "a.npy"), (None, None, "b.npy")]
    per_file = [{"Pk_loss": 0.1, "WD_loss": 0.2, "test_loss": torch.tensor(0.5), "threshold": 0.4},
                {"Pk_loss": 0.3, "WD_loss": 0.4, "test_loss": 0.25, "threshold": 0.4}]
    scores = [np.array([0.5, 0.25], dtype=np.float32), torch.tensor([0.125])]
    res, sc = rio.collect_per_file(files, per_file, scores, metric="F1")
    assert res["a.npy"] == {"Pk_loss": 0.1, "WD_loss": 0.2, "threshold": 0.4, "F1": 0.5} and "test_loss" not in res["b.npy"]
    rio.write_per_file_json(str(tmp_path), res, sc)
    back_res, back_sc = rio.read_per_file_json(str(tmp_path))
    assert back_res == res and back_sc == {"a.npy": [0.5, 0.25], "b.npy": [0.125]}


def test_segment_ranges_and_encoder_table():
    """Sample ranges of predict.py:105-127 (hand-worked cases) and the encoder -> width table of predict.py:184-216."""
    from multimodaltopicsegmentation_b200.predict import encoder_embedding_dim, segment_ranges

    # 5.5 s of audio at sr 10, 1-s units, boundaries after units 1 and 3: remainder appended as the last segment
    assert segment_ranges(55, [0, 1, 0, 1, 0], sr=10) == [(0, 20), (20, 40), (40, 55)]
    # fewer predictions than units: stops at the end of the predictions (the reference's IndexError branch)
    assert segment_ranges(55, [1], sr=10) == [(0, 10), (10, 55)]
    assert segment_ranges(55, [0, 0, 0, 0, 0], sr=10) == [(0, 55)]
    assert segment_ranges(40, [0, 1, 0, 1], sr=10, interval=2) == [(0, 40), (40, 40)]
    # adaptive: 100 chunks of n // 100 samples, no trailing remainder
    seg = [0] * 100
    seg[9] = seg[99] = 1
    assert segment_ranges(1000, seg, adaptive=True) == [(0, 100), (100, 1000)]
    with pytest.raises(IndexError):
        segment_ranges(1000, [0] * 50, adaptive=True)
    for name, dim in (("x-vectors", 512), ("ecapa", 192), ("wav2vec", 768), ("wav2vec_std", 1536), ("openl3_std", 1024),
                      ("openl3", 512), ("crepe_std", 512), ("crepe", 256), ("mfcc", 200), ("prosodic", 167)):
        assert encoder_embedding_dim(name) == dim
    assert encoder_embedding_dim("anything", pca_reduce=True, pca_value=33) == 33
    with pytest.raises(ValueError):
        encoder_embedding_dim("roberta")


def test_recurrent_longformer_parameter_tree_and_constructor_checks():
    """`BiLSTMRestrictedMHA` (models/CRF.py:764-858): state-dict keys of the blocks / the FFN-less attention layer as the
    reference names them, the reference's constructor assertions, and the TextSegmenter dispatch (lightning_model.py:215)."""
    import pytest

    from multimodaltopicsegmentation_b200 import RecurrentLongformer, TextSegmenter

    m = RecurrentLongformer(2, 24, 16, num_layers=2, nheads=4, loss_fn="FocalLoss", window_size=8)
    keys = set(m.state_dict().keys())
    for blk in (0, 1):
        for d in ("", "_reverse"):
            for w in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                assert f"model.{blk}.lstm.rnn.{w}_l0{d}" in keys
        for proj in ("query", "key", "value", "query_global", "key_global", "value_global"):
            assert f"model.{blk}.transformer.model.attention.self.{proj}.weight" in keys
        assert not any(k.startswith(f"model.{blk}.transformer.model.attention.output") for k in keys)   # no dense / LayerNorm
    assert "model.2.rnn.weight_ih_l0" in keys and "classification.weight" in keys
    assert m.model[0].lstm.rnn.input_size == 24 and m.model[1].lstm.rnn.input_size == 16
    assert m.model[0].transformer.model.attention.self.query.weight.shape == (16, 16)   # separate forward / backward: d = H
    assert m.classification.in_features == 32
    with pytest.raises(AssertionError):      # RestrictedTransformerLayer.py:77-80 (the reference's default 127 trips it too)
        RecurrentLongformer(2, 24, 16, num_layers=1, nheads=4, window_size=127)
    seg = TextSegmenter(architecture="BiLSTMRestrictedMHA", tagset_size=2, embedding_dim=24, hidden_dim=16, num_layers=1,
                        loss_fn="FocalLoss", nheads=4, attention_window=8)
    assert isinstance(seg.model, RecurrentLongformer)


def test_f16_pieces_weight_packing():
    """ops.f16_pieces (host-side weight packing of mts_gemm_f16x3, plain torch): exact power-of-two row scales that put the
    row maximum into [2^13, 2^14), pieces that reproduce the scaled row to ~2^-22 of its maximum, zero K padding, zero rows."""
    import torch

    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(3)
    w = torch.randn(37, 100, generator=g) * (10.0 ** (torch.rand(37, 1, generator=g) * 8 - 5))
    w[5] = 0.0
    pieces, scale = ops.f16_pieces(w)
    assert pieces.shape == (37, 2, 128) and pieces.dtype == torch.float16 and scale.shape == (37,)
    lg = torch.log2(scale)
    assert torch.equal(lg, lg.round())
    top = w.abs().amax(dim=1) / scale
    keep = w.abs().amax(dim=1) > 0
    assert float(top[keep].min()) >= 2.0 ** 13 and float(top[keep].max()) < 2.0 ** 14
    rec = (pieces[:, 0, :100].float() + pieces[:, 1, :100].float()) * scale[:, None]
    rel = (rec - w).abs() / w.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    assert float(rel.max()) < 2.0 ** -21
    assert float(pieces[:, :, 100:].abs().max()) == 0.0 and float(pieces[5].abs().max()) == 0.0
    assert torch.isfinite(pieces.float()).all()


def test_weight_pieces_stacked_falls_back_to_the_host_formulation_off_device():
    """ops.f16_pieces_stacked on matrices the packing kernel does not take (here: CPU tensors) is f16_pieces of their row-wise
    concatenation -- the same bits the kernel path is held to on the GPU (test_f16_weight_pieces_kernel_equals_host_formulation)."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(3)
    mats = [torch.randn(12, 40, generator=g), torch.randn(20, 40, generator=g) * 1e-4]
    p, sc = ops.f16_pieces_stacked(mats)
    p_ref, sc_ref = ops.f16_pieces(torch.cat(mats, dim=0))
    assert p.shape == (32, 2, 64) and torch.equal(p.view(torch.int16), p_ref.view(torch.int16)) and torch.equal(sc, sc_ref)
    back = (p[:, 0, :40].double() + p[:, 1, :40].double()) * sc.double()[:, None]
    assert float((back - torch.cat(mats).double()).abs().max() / torch.cat(mats).abs().max()) < 2 ** -21


def test_attention_dropout_seeds_follow_the_torch_generator_or_the_hook(monkeypatch):
    from multimodaltopicsegmentation_b200 import transformer

    torch.manual_seed(123)
    a = [transformer.attn_seed(l) for l in range(3)]
    torch.manual_seed(123)
    assert a == [transformer.attn_seed(l) for l in range(3)] and len(set(a)) == 3 and all(0 <= v < 2 ** 62 for v in a)
    monkeypatch.setattr(transformer, "ATTN_SEED_FN", lambda layer: 40 + layer)
    assert [transformer.attn_seed(l) for l in range(3)] == [40, 41, 42]
