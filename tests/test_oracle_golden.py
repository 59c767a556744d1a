"""Pins the CPU oracle (oracle/ref_numpy.py, oracle/ref_torch.py, oracle/oracle_ref.c) against the
golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import params_of
from oracle import c_oracle, ref_numpy as rn, ref_torch as rt

RTOL, ATOL = 1e-4, 1e-5  # north_star: logits / losses / gradients within rtol 1e-4 in fp32


def tags_equal(tags, arr):
    for b, t in enumerate(tags):
        n = len(t)
        assert (arr[b, n:] == -1).all()
        assert [int(v) for v in t] == arr[b, :n].tolist()


@pytest.mark.parametrize("loss_fn", ["focalloss", "binarycrossentropy", "crossentropy"])
def test_bilstm_numpy_and_c(golden, loss_fn):
    fx = golden("bilstm_" + loss_fn)
    p = params_of(fx)
    x, lengths, th = fx["i:x"], fx["i:lengths"], float(fx["i:th"])
    bce = loss_fn != "crossentropy"
    scores, tags = rn.bilstm_segmenter(x, lengths, p, th=th, bce=bce)
    np.testing.assert_allclose(scores, fx["o:scores"], rtol=RTOL, atol=ATOL)
    tags_equal(tags, fx["o:tags"])
    h_c = c_oracle.bilstm_stack(x, lengths, p, "model.rnn.", 2)
    s_c = rn.linear(h_c, p["classification.weight"], p["classification.bias"])
    np.testing.assert_allclose(s_c, fx["o:scores"], rtol=RTOL, atol=ATOL)
    y = fx["i:y"]
    if loss_fn == "focalloss":
        loss = rn.focal_loss(rn.unpad(scores[:, :, 0], lengths), rn.unpad(y, lengths))
    elif loss_fn == "binarycrossentropy":
        loss = rn.bce_loss(rn.unpad(scores[:, :, 0], lengths), rn.unpad(y, lengths))
    else:
        loss = rn.cross_entropy_ignore(scores, y[:, : scores.shape[1]])
    np.testing.assert_allclose(loss, fx["o:loss"], rtol=RTOL)


@pytest.mark.parametrize("loss_fn", ["FocalLoss", "BinaryCrossEntropy", "CrossEntropy"])
def test_bilstm_torch_twin(golden, loss_fn):
    fx = golden("bilstm_" + loss_fn.lower())
    m = rt.Segmenter(2, 12, 8, num_layers=2, loss_fn=loss_fn)
    assert rt.load_golden_params(m, fx) == ([], [])
    x, lengths, y = (torch.from_numpy(fx[k]) for k in ("i:x", "i:lengths", "i:y"))
    m.th = float(fx["i:th"])
    scores, tags = m(x, lengths)
    np.testing.assert_allclose(scores.detach().numpy(), fx["o:scores"], rtol=RTOL, atol=ATOL)
    tags_equal(tags, fx["o:tags"])
    loss = m.loss(x, lengths, y)
    loss.backward()
    np.testing.assert_allclose(loss.item(), fx["o:loss"], rtol=RTOL)
    for k, prm in m.named_parameters():
        np.testing.assert_allclose(prm.grad.numpy(), fx["g:" + k], rtol=RTOL, atol=1e-6, err_msg=k)


def test_rnn_encoder(golden):
    fx = golden("rnn_encoder")
    out = rn.bilstm_stack(fx["i:x"], fx["i:lengths"], params_of(fx), prefix="rnn.")
    np.testing.assert_allclose(out, fx["o:out"], rtol=RTOL, atol=ATOL)
    assert out.shape[1] == int(fx["i:lengths"].max())
    # padded steps are exact zeros
    for b, n in enumerate(fx["i:lengths"]):
        assert not out[b, n:].any()


def test_late_fusion(golden):
    fx = golden("latefusion_focal")
    p = params_of(fx)
    scores, tags = rn.late_fusion_segmenter(fx["i:x1"], fx["i:x2"], fx["i:lengths"], p, th=float(fx["i:th"]))
    np.testing.assert_allclose(scores, fx["o:scores"], rtol=RTOL, atol=ATOL)
    tags_equal(tags, fx["o:tags"])
    m = rt.LateFusion(2, [5, 7], 8, num_layers=2, loss_fn="FocalLoss")
    assert rt.load_golden_params(m, fx) == ([], [])
    t = lambda k: torch.from_numpy(fx[k])
    loss = m.loss(t("i:x1"), t("i:x2"), t("i:lengths"), t("i:y"))
    loss.backward()
    np.testing.assert_allclose(loss.item(), fx["o:loss"], rtol=RTOL)
    for k, prm in m.named_parameters():
        np.testing.assert_allclose(prm.grad.numpy(), fx["g:" + k], rtol=RTOL, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("name", ["crf_small", "crf_ragged"])
def test_crf(golden, name):
    fx = golden(name)
    emis, lengths, trans, ys = fx["o:emissions"], fx["i:lengths"], fx["p:transitions"], fx["i:ys"]
    # Viterbi: bit-exact scores and identical paths, numpy and C
    best, paths = rn.crf_viterbi(emis, lengths, trans)
    assert np.array_equal(best, fx["o:best_score"])
    tags_equal(paths, fx["o:paths"])
    best_c, paths_c = c_oracle.crf_viterbi(emis, lengths, trans)
    assert np.array_equal(best_c, fx["o:best_score"])
    assert np.array_equal(paths_c.astype(np.int8), fx["o:paths"])
    # partition, gold score, NLL
    np.testing.assert_allclose(rn.crf_forward_algorithm(emis, lengths, trans), fx["o:forward_score"], rtol=1e-5)
    np.testing.assert_allclose(c_oracle.crf_forward(emis, lengths, trans), fx["o:forward_score"], rtol=1e-5)
    np.testing.assert_allclose(rn.crf_gold_score(emis, ys, lengths, trans), fx["o:gold_score"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(c_oracle.crf_gold(emis, ys, lengths, trans), fx["o:gold_score"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rn.crf_nll(emis, ys, lengths, trans), fx["o:loss"], rtol=RTOL)
    # closed-form gradient = autograd through the reference's T-step loop
    ge, gt = rn.crf_marginal_grad(emis, ys, lengths, trans)
    np.testing.assert_allclose(ge, fx["g:emissions"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(gt, fx["g:transitions"], rtol=RTOL, atol=2e-6)
    # the Viterbi path's own score equals best_score (property from SURVEY.md section 4)
    for b, path in enumerate(paths):
        n = int(lengths[b])
        pad = np.zeros((1, emis.shape[1]), dtype=np.int64)
        pad[0, :n] = path
        gold = rn.crf_gold_score(emis[b:b + 1], pad, lengths[b:b + 1], trans)
        np.testing.assert_allclose(gold[0], best[b], rtol=1e-5)


def test_crf_torch_twin(golden):
    fx = golden("crf_ragged")
    crf = rt.ChainCRF(16, 2)
    assert rt.load_golden_params(crf, fx) == ([], [])
    feats, lengths, ys = (torch.from_numpy(fx[k]) for k in ("i:features", "i:lengths", "i:ys"))
    mask = rt.length_mask(feats.shape[1], lengths)
    best, paths = crf(feats, mask)
    assert np.array_equal(best.detach().numpy(), fx["o:best_score"])
    tags_equal(paths, fx["o:paths"])
    loss = crf.loss(feats, ys, mask)
    loss.backward()
    np.testing.assert_allclose(loss.item(), fx["o:loss"], rtol=RTOL)
    for k, prm in crf.named_parameters():
        np.testing.assert_allclose(prm.grad.numpy(), fx["g:" + k], rtol=RTOL, atol=2e-6, err_msg=k)


def test_bilstm_crf_composition(golden):
    fx = golden("bilstm_crf")
    p = params_of(fx)
    feats = rn.bilstm_stack(fx["i:x"], fx["i:lengths"], p, prefix="model.rnn.")
    np.testing.assert_allclose(feats, fx["o:features"], rtol=RTOL, atol=ATOL)
    emis = rn.linear(fx["o:features"], p["crf.fc.weight"], p["crf.fc.bias"])
    best, paths = rn.crf_viterbi(emis, fx["i:lengths"], p["crf.transitions"])
    np.testing.assert_allclose(best, fx["o:best_score"], rtol=1e-5)
    tags_equal(paths, fx["o:paths"])
    m = rt.EncoderCRF(2, 12, 8, num_layers=2)
    assert rt.load_golden_params(m, fx) == ([], [])
    t = lambda k: torch.from_numpy(fx[k])
    best_t, paths_t = m(t("i:x"), t("i:lengths"))
    tags_equal(paths_t, fx["o:paths"])
    loss = m.loss(t("i:x"), t("i:lengths"), t("i:ys"))
    loss.backward()
    np.testing.assert_allclose(loss.item(), fx["o:loss"], rtol=RTOL)
    for k, prm in m.named_parameters():
        np.testing.assert_allclose(prm.grad.numpy(), fx["g:" + k], rtol=2e-4, atol=2e-6, err_msg=k)


def test_transformer(golden):
    fx = golden("transformer_focal")
    p = params_of(fx)
    nh, w = int(fx["i:nheads"]), int(fx["i:window"])
    hidden = rn.longformer_encoder(fx["i:x"], fx["i:lengths"], p, nh, rn.pyramid_windows(2, w))
    lengths = fx["i:lengths"]
    for b, n in enumerate(lengths):  # only valid rows are defined by the contract (padded rows are masked junk)
        np.testing.assert_allclose(hidden[b, :n], fx["o:hidden"][b, :n], rtol=RTOL, atol=2e-5)
    scores = rn.linear(hidden, p["classification.weight"], p["classification.bias"])
    tags = rn.threshold_tags(scores, lengths, float(fx["i:th"]))
    tags_equal(tags, fx["o:tags"])
    loss = rn.focal_loss(rn.unpad(scores[:, :, 0], lengths), rn.unpad(fx["i:y"], lengths))
    np.testing.assert_allclose(loss, fx["o:loss"], rtol=RTOL)


def test_losses(golden):
    fx = golden("losses")
    np.testing.assert_allclose(rn.focal_loss(fx["z"], fx["y"]), fx["focal"], rtol=1e-5)
    np.testing.assert_allclose(rn.focal_loss_grad(fx["z"], fx["y"]), fx["focal_grad"], rtol=RTOL, atol=1e-8)
    np.testing.assert_allclose(rn.bce_loss(fx["z"], fx["y"]), fx["bce"], rtol=1e-5)


def test_pk_windowdiff_hand_cases():
    # UNPINNED against segeval (not installed); hand-computed from the published definitions.
    ref = [0, 0, 1, 0, 0, 0, 1, 0, 0, 0]  # masses 3,4,3 -> k = round(10/3/2) = round(1.67) = 2
    same = rn.compute_pk(ref, ref)
    assert same == 0 and rn.compute_window_diff(ref, ref) == 0
    hyp = [0] * 10  # one segment: every probe spanning a reference boundary disagrees
    # probes (i, i+2), i = 0..7: reference boundary between i and i+2 for i in {1,2,5,6}
    assert rn.compute_pk(hyp, ref) == rn.Decimal(4) / 8
    assert rn.compute_window_diff(hyp, ref) == rn.Decimal(4) / 8
    # C twin agrees on random cases
    rng = np.random.default_rng(0)
    for n in (5, 17, 64, 301):
        for _ in range(20):
            r = (rng.random(n) < 0.15).astype(np.uint8)
            h = (rng.random(n) < 0.15).astype(np.uint8)
            r[-1] = h[-1] = 0
            num, den = c_oracle.pk(h, r)
            assert rn.compute_pk(h, r) == (rn.Decimal(num) / den if den else 0)
            try:
                wd = rn.compute_window_diff(h, r)
                num, den = c_oracle.window_diff(h, r)
                assert wd == rn.Decimal(num) / den
            except AssertionError:
                assert c_oracle.window_diff(h, r)[0] == -1


def test_attention_dropout_keep_mask_restates_the_device_hash():
    """oracle/ref_numpy.py:attn_dropout_keep against the hash written out with Python integers (the definition the device
    function attn_keep_scale in csrc/common.cuh follows): splitmix64 finaliser of seed ^ (bh << 40 | i << 20 | j), top 24 bits
    compared with round(p 2^24).  Plus the statistics a dropout mask needs."""
    M = (1 << 64) - 1

    def keep(seed, bh, i, j, p):
        x = (seed ^ ((bh << 40) | (i << 20) | j)) & M
        x ^= x >> 30
        x = (x * 0xBF58476D1CE4E5B9) & M
        x ^= x >> 27
        x = (x * 0x94D049BB133111EB) & M
        x ^= x >> 31
        return (x >> 40) >= int(float(np.float32(p)) * 16777216.0 + 0.5)

    seed, B, H, S, p = 0x1234_5678_9ABC_DEF0, 2, 3, 40, 0.25
    got = rn.attn_dropout_keep(seed, B, H, S, p)
    assert got.shape == (B, H, S, S) and got.dtype == bool
    for b in range(B):
        for h in range(H):
            for i in (0, 7, 39):
                for j in (0, 1, 38):
                    assert bool(got[b, h, i, j]) == keep(seed, b * H + h, i, j, p), (b, h, i, j)
    big = rn.attn_dropout_keep(seed, 2, 4, 256, 0.1)
    assert abs((1.0 - big.mean()) - 0.1) < 3e-3
    assert abs((1.0 - big[0, 0].mean()) - 0.1) < 1e-2 and abs((1.0 - big[:, :, 5].mean()) - 0.1) < 3e-2   # no head / row is special
    other = rn.attn_dropout_keep(seed + 1, 2, 4, 256, 0.1)
    assert 0.15 < (big != other).mean() < 0.21      # independent masks differ at 2 p (1 - p) = 0.18 of the positions
    assert rn.attn_dropout_keep(seed, 1, 1, 16, 0.0).all()
