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
