"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every check calls the product through the
reference-shaped module API, i.e. through the C ABI of libmts_b200.so, and compares with the CPU oracle
(oracle/ref_numpy.py, oracle/ref_torch.py, oracle/oracle_ref.c) or the committed golden vectors.

Tolerances: integer / index work (Viterbi paths, thresholded tags, Pk, WindowDiff) is bit-exact;
floating point follows north_star: rtol 1e-4 in fp32 (plus an absolute floor for values near zero).
"""
import numpy as np
import pytest
import torch

from conftest import params_of

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 2e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import multimodaltopicsegmentation_b200 as m

    m.ops.device_ok()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def load_params(module, fx, dev):
    sd = {k[2:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("p:")}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return module.to(dev)


def tags_equal(tags, arr):
    for b, t in enumerate(tags):
        n = len(t)
        assert (arr[b, n:] == -1).all()
        assert [int(v) for v in t] == arr[b, :n].tolist()


def check_operand_pair(hi, lo, x, side):
    """(hi, lo) as the tensor-core GEMM wants them (include/mts_b200.h "Operand preparation"): hi = x zero-padded to Kp;
    lo = bf16 [rows, 2 Kp], per 16-wide K block [bf16(x) | bf16(x - trunc_tf32(x))] for an A operand, swapped for B."""
    rows, K = x.shape
    kp = hi.shape[1]
    assert kp % 32 == 0 and kp >= K and tuple(lo.shape) == (rows, kp)
    assert torch.equal(hi[:, :K], x) and float(hi[:, K:].abs().max() if kp > K else 0.0) == 0.0
    xp = torch.zeros(rows, kp, device=x.device)
    xp[:, :K] = x
    rest = xp - (xp.view(torch.int32) & ~0x1FFF).view(torch.float32)
    blocks = lo.contiguous().view(torch.bfloat16).view(rows, kp // 16, 2, 16)
    first, second = (rest, xp) if side else (xp, rest)
    assert torch.equal(blocks[:, :, 0], first.to(torch.bfloat16).view(rows, kp // 16, 16))
    assert torch.equal(blocks[:, :, 1], second.to(torch.bfloat16).view(rows, kp // 16, 16))


def close(a, b, rtol=RTOL, atol=ATOL, msg=""):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


# ----------------------------------------------------------------------------------------------------------
# GEMMs
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(300, 256, 96), (1000, 2048, 896), (129, 136, 40), (4096, 1024, 512), (64, 8, 33)])
def test_gemm_f32_nt(dev, M, N, K):
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(M + N + K)
    a, b, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    ref = (a.double() @ b.double().T + bias.double()).float()
    ad, bd, biasd = a.to(dev), b.to(dev), bias.to(dev)
    c = torch.empty(M, N, device=dev)
    ops.gemm_f32(ad.data_ptr(), K, bd.data_ptr(), K, biasd, c.data_ptr(), N, M, N, K, layout=0, epilogue=1)
    close(c, ref, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("M,N,K", [(300, 256, 96), (1000, 2048, 896), (129, 136, 40), (4096, 1024, 512), (19200, 2048, 896),
                                   (256, 256, 8192), (200, 100, 5000), (1024, 256, 22920)])  # the last three: split-K
def test_gemm_tf32x3(dev, M, N, K):
    """tcgen05 3xTF32 vs float64: error must be fp32-grade (no worse than 4x a plain fp32 matmul's)."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(M * 3 + N + K)
    a, b, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    ad, bd, biasd = a.to(dev), b.to(dev), bias.to(dev)
    ref = (ad.double() @ bd.double().T + biasd.double())
    a_hi, a_lo = ops.split_tf32(ad, side=ops.A_SIDE)
    b_hi, b_lo = ops.split_tf32(bd, side=ops.B_SIDE)
    check_operand_pair(a_hi, a_lo, ad, side=0)
    check_operand_pair(b_hi, b_lo, bd, side=1)
    c = torch.empty(M, N, device=dev)
    ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, biasd, c, M, N, epilogue=1)
    err = (c.double() - ref).abs().max().item()
    err32 = ((ad @ bd.T + biasd).double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"tf32x3 M={M} N={N} K={K}: max err {err:.3e}  (fp32 matmul {err32:.3e}, scale {scale:.1f})")
    # Operands are exact to 2^-22; what remains is the tensor core's own fp32 accumulation, which truncates instead
    # of rounding and so drifts ~K ulps (measured 9e-6 of the output scale at K = 896, 6x a cuBLAS fp32 matmul).
    # Contract: norm-wise 2e-5, i.e. 5x inside north_star's rtol 1e-4.
    assert err <= max(4 * err32, 2e-5 * scale), (err, err32, scale)
    # accumulate and GELU epilogues
    c2 = c.clone()
    ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, None, c2, M, N, epilogue=0, accumulate=True)
    close(c2, (2 * ref - biasd.double()).float(), rtol=1e-5, atol=4e-5 * scale)
    if K <= 3072:  # the GELU epilogue cannot be split over K; the encoder's GELU GEMM has K = d_model
        c3 = torch.empty(M, N, device=dev)
        ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, biasd, c3, M, N, epilogue=2)
        close(c3, torch.nn.functional.gelu(ref).float(), rtol=1e-5, atol=2e-5 * scale)


@pytest.mark.parametrize("M,N,K", [(300, 256, 896), (1000, 64, 96), (135, 896, 256), (4100, 512, 128)])
def test_gemm_gelu_pair_epilogue_equals_gemm_then_gelu_split(dev, M, N, K):
    """mts_gemm_tf32x3_gelu_pair (dense + GELU + operand pair in one epilogue) against the two-kernel form it replaces in
    inference: bit-identical activation and correction operand."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev, generator=g)
    b = torch.randn(N, K, device=dev, generator=g) * 0.1
    bias = torch.randn(N, device=dev, generator=g)
    a_hi, a_lo = ops.split_tf32(a)
    b_hi, b_lo = ops.split_tf32(b, side=ops.B_SIDE)
    zp = torch.empty(M, N, device=dev)
    ops.gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, zp, M, N, epilogue=1)
    two = torch.full((2, M, N), float("nan"), device=dev)
    ops._call("mts_gelu_split", zp.data_ptr(), N, M, N, N, 0, two[0].data_ptr(), two[1].data_ptr(), ops._stream())
    one = torch.full((2, M, N), float("nan"), device=dev)
    ops._call("mts_gemm_tf32x3_gelu_pair", a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), bias.data_ptr(),
              one[0].data_ptr(), one[1].data_ptr(), M, N, a_hi.shape[1], ops._stream())
    assert torch.equal(one[0], two[0])
    assert torch.equal(one[1].view(torch.int32), two[1].view(torch.int32))
    check_operand_pair(one[0], one[1], one[0], side=0)
    from multimodaltopicsegmentation_b200 import _lib
    with pytest.raises(_lib.MtsError):
        ops._call("mts_gemm_tf32x3_gelu_pair", a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), bias.data_ptr(),
                  one[0].data_ptr(), one[1].data_ptr(), M, N - 8, a_hi.shape[1], ops._stream())


@pytest.mark.parametrize("rows,cols,T,shift", [(100, 40, 100, 0), (3 * 37, 65, 37, -1), (3 * 37, 65, 37, 1), (2048, 896, 2048, 0)])
def test_transpose_split(dev, rows, cols, T, shift):
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(rows + cols)
    B = rows // T
    src = torch.randn(B, T + 3, cols + 5, generator=g).to(dev)   # padded batch / row strides
    lengths = torch.randint(1, T + 1, (B,), generator=g)
    lens_dev = lengths.to(torch.int32).to(dev) if shift else None
    hi, lo = ops.transpose_split(src.data_ptr(), src.stride(0), src.stride(1), rows, cols, T, dev, shift=shift,
                                 lengths=lens_dev, side=ops.B_SIDE)
    kp = (rows + 31) // 32 * 32
    assert tuple(hi.shape) == (cols, kp)
    ref = torch.zeros(rows, cols)
    sc = src.cpu()
    for b in range(B):
        n = int(lengths[b]) if shift else T
        for t in range(T):
            if 0 <= t + shift < n:
                ref[b * T + t] = sc[b, t + shift, :cols]
    check_operand_pair(hi, lo, ref.T.contiguous().to(dev), side=1)


def test_gemm_tn_shift(dev):
    """dW_hh-style product: A^T B with the B rows shifted inside episode boundaries."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(5)
    B, T, M, N = 5, 13, 70, 40
    lengths = torch.tensor([13, 4, 9, 1, 7], dtype=torch.int32)
    a = torch.randn(B * T, M, generator=g)
    h = torch.randn(B * T, N, generator=g)
    for shift in (-1, 1):
        hs = torch.zeros_like(h).view(B, T, N)
        hv = h.view(B, T, N)
        for b, n in enumerate(lengths.tolist()):
            for t in range(T):
                if 0 <= t + shift < n:
                    hs[b, t] = hv[b, t + shift]
        ref = a.double().T @ hs.view(B * T, N).double()
        ad, hd, ld = a.to(dev), h.to(dev), lengths.to(dev)  # keep the device tensors alive across the call
        for splits in (1, 4):
            c = torch.empty(M, N, device=dev)
            ops.gemm_f32(ad.data_ptr(), M, hd.data_ptr(), N, None, c.data_ptr(), N, M, N, B * T, layout=3,
                         splits=splits, shift=shift, T=T, lengths=ld)
            close(c, ref.float(), rtol=1e-5, atol=1e-4)


# ----------------------------------------------------------------------------------------------------------
# CRF
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["crf_small", "crf_ragged"])
def test_crf_golden(dev, golden, name):
    from multimodaltopicsegmentation_b200 import CRF
    from oracle import ref_torch as rt

    fx = golden(name)
    crf = load_params(CRF(16, 2), fx, dev)
    feats = torch.from_numpy(fx["i:features"]).to(dev)
    lengths = torch.from_numpy(fx["i:lengths"])
    masks = rt.length_mask(feats.shape[1], lengths).to(dev)
    best, paths = crf(feats, masks)
    tags_equal(paths, fx["o:paths"])
    close(best, fx["o:best_score"], rtol=1e-5)  # emissions come from our own fc GEMM (rounding differs by ulps)
    loss = crf.loss(feats, torch.from_numpy(fx["i:ys"]).to(dev), masks)
    loss.backward()
    close(loss, fx["o:loss"])
    for k, p in crf.named_parameters():
        ref_g = fx["g:" + k]
        close(p.grad, ref_g, atol=1e-5 * float(np.abs(ref_g).max()) + 3e-6, msg=k)


@pytest.mark.parametrize("B,L", [(64, 300), (8, 8192), (300, 77), (3, 1)])
def test_crf_viterbi_bit_exact_vs_c_oracle(dev, B, L):
    """Identical emissions => identical paths and bit-identical best scores (north_star)."""
    from multimodaltopicsegmentation_b200 import ops
    from oracle import c_oracle

    rng = np.random.default_rng(B * 1000 + L)
    emis = rng.standard_normal((B, L, 4)).astype(np.float32) * 2
    # plant exact ties so that first-max-wins is exercised
    emis[:, ::7, 0] = emis[:, ::7, 1]
    lengths = rng.integers(1, L + 1, size=B)
    lengths[0] = L
    trans = rng.standard_normal((4, 4)).astype(np.float32)
    trans[2, :] = -1e4
    trans[:, 3] = -1e4
    trans[0, 1] = trans[0, 0]
    best_ref, paths_ref = c_oracle.crf_viterbi(emis, lengths, trans)
    lens = ops.Lengths(lengths.tolist(), dev, L)
    best, paths = ops.crf_viterbi(torch.from_numpy(emis).to(dev), lens, torch.from_numpy(trans).to(dev))
    assert np.array_equal(best.cpu().numpy(), best_ref)
    assert np.array_equal(paths.cpu().numpy(), paths_ref)
    # idempotence / size-independent property: the decoded path's own score equals best_score
    gold = c_oracle.crf_gold(emis, np.maximum(paths_ref, 0), lengths, trans)
    np.testing.assert_allclose(gold, best_ref, rtol=2e-4, atol=1e-3)  # differs only by fp32 summation order


def test_crf_nll_vs_oracle_large(dev):
    from multimodaltopicsegmentation_b200 import ops
    from oracle import c_oracle, ref_numpy as rn

    rng = np.random.default_rng(11)
    B, L = 10, 700
    emis = rng.standard_normal((B, L, 4)).astype(np.float32)
    lengths = rng.integers(1, L + 1, size=B)
    lengths[3] = L
    tags = (rng.random((B, L)) < 0.3).astype(np.float32)
    trans = rng.standard_normal((4, 4)).astype(np.float32)
    trans[2, :] = -1e4
    trans[:, 3] = -1e4
    lens = ops.Lengths(lengths.tolist(), dev, L)
    e = torch.from_numpy(emis).to(dev).requires_grad_(True)
    tr = torch.from_numpy(trans).to(dev).requires_grad_(True)
    stats = ops.CrfNllFn.apply(e, tr, torch.from_numpy(tags).to(dev), lens)
    close(stats[0], c_oracle.crf_forward(emis, lengths, trans), rtol=1e-5)
    close(stats[1], c_oracle.crf_gold(emis, tags, lengths, trans), rtol=1e-5, atol=1e-4)
    (stats[0] - stats[1]).mean().backward()
    ge, gt = rn.crf_marginal_grad(emis, tags, lengths, trans)
    close(e.grad, ge, atol=1e-6)
    close(tr.grad, gt, rtol=2e-4, atol=1e-4)


# ----------------------------------------------------------------------------------------------------------
# head + losses
# ----------------------------------------------------------------------------------------------------------
def test_losses_golden(dev, golden):
    from multimodaltopicsegmentation_b200 import ops

    fx = golden("losses")
    n = len(fx["z"])
    lens = ops.Lengths([n], dev, n)
    for kind, key in ((0, "focal"), (1, "bce")):
        z = torch.from_numpy(fx["z"]).to(dev).view(1, n, 1).requires_grad_(True)
        y = torch.from_numpy(fx["y"]).to(dev).view(1, n)
        loss = ops.SegLossFn.apply(z, y, lens, kind, 0.9, 2.0, 1.0 / n)
        loss.backward()
        close(loss, fx[key], rtol=1e-5)
        close(z.grad.view(-1), fx[key + "_grad"], atol=1e-8)


# ----------------------------------------------------------------------------------------------------------
# BiLSTM family on the golden fixtures (generic recurrence kernel, H = 8)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("loss_fn", ["FocalLoss", "BinaryCrossEntropy", "CrossEntropy"])
def test_bilstm_golden(dev, golden, loss_fn):
    from multimodaltopicsegmentation_b200 import BiLSTM

    fx = golden("bilstm_" + loss_fn.lower())
    m = load_params(BiLSTM(2, 12, 8, num_layers=2, loss_fn=loss_fn), fx, dev)
    x, y = torch.from_numpy(fx["i:x"]).to(dev), torch.from_numpy(fx["i:y"]).to(dev)
    lengths = torch.from_numpy(fx["i:lengths"])
    m.th = float(fx["i:th"])
    scores, tags = m(x, lengths)
    assert tuple(scores.shape) == fx["o:scores"].shape  # time axis = max(lengths)
    close(scores, fx["o:scores"])
    tags_equal(tags, fx["o:tags"])
    assert all(isinstance(v, bool) for v in tags[0])
    loss = m.loss(x, lengths, y)
    loss.backward()
    close(loss, fx["o:loss"])
    for k, p in m.named_parameters():
        close(p.grad, fx["g:" + k], atol=2e-6, msg=k)


def test_late_fusion_golden(dev, golden):
    from multimodaltopicsegmentation_b200 import BiLSTMLateFusion

    fx = golden("latefusion_focal")
    m = load_params(BiLSTMLateFusion(2, [5, 7], 8, num_layers=2, loss_fn="FocalLoss"), fx, dev)
    t = lambda k: torch.from_numpy(fx[k]).to(dev)
    m.th = float(fx["i:th"])
    scores, tags = m(t("i:x1"), t("i:x2"), torch.from_numpy(fx["i:lengths"]))
    close(scores, fx["o:scores"])
    tags_equal(tags, fx["o:tags"])
    loss = m.loss(t("i:x1"), t("i:x2"), torch.from_numpy(fx["i:lengths"]), t("i:y"))
    loss.backward()
    close(loss, fx["o:loss"])
    for k, p in m.named_parameters():
        close(p.grad, fx["g:" + k], atol=2e-6, msg=k)


def test_bilstm_crf_golden(dev, golden):
    from multimodaltopicsegmentation_b200 import BiRnnCrf

    fx = golden("bilstm_crf")
    m = load_params(BiRnnCrf(2, 12, 8, num_layers=2), fx, dev)
    x = torch.from_numpy(fx["i:x"]).to(dev)
    lengths = torch.from_numpy(fx["i:lengths"])
    best, paths = m(x, lengths)
    tags_equal(paths, fx["o:paths"])
    close(best, fx["o:best_score"], rtol=1e-4)
    loss = m.loss(x, lengths, torch.from_numpy(fx["i:ys"]).to(dev))
    loss.backward()
    close(loss, fx["o:loss"])
    for k, p in m.named_parameters():
        close(p.grad, fx["g:" + k], rtol=2e-4, atol=3e-6, msg=k)


def test_early_fusion_pair_input_equals_concat(dev, golden):
    """Additive API: (text, audio) pair == host-side torch.cat (load_datasets_precomputed.py:158-161)."""
    from multimodaltopicsegmentation_b200 import BiLSTM

    fx = golden("bilstm_focalloss")
    m = load_params(BiLSTM(2, 12, 8, num_layers=2, loss_fn="FocalLoss"), fx, dev)
    x = torch.from_numpy(fx["i:x"]).to(dev)
    lengths = torch.from_numpy(fx["i:lengths"])
    s1, _ = m(x, lengths)
    s2, _ = m((x[:, :, :5].contiguous(), x[:, :, 5:].contiguous()), lengths)
    assert torch.equal(s1, s2)


# ----------------------------------------------------------------------------------------------------------
# H = 256: the cluster recurrence kernels vs the torch CPU oracle
# ----------------------------------------------------------------------------------------------------------
def _oracle_pair(dev, D, H, L, loss_fn="FocalLoss", seed=0):
    from multimodaltopicsegmentation_b200 import BiLSTM
    from oracle import ref_torch as rt

    torch.manual_seed(seed)
    ref = rt.Segmenter(2, D, H, num_layers=L, loss_fn=loss_fn)
    with torch.no_grad():
        ref.classification.weight.mul_(4.0)
    ours = BiLSTM(2, D, H, num_layers=L, loss_fn=loss_fn)
    ours.load_state_dict(ref.state_dict())
    return ref, ours.to(dev)


@pytest.mark.parametrize("B,T,D,lens", [(11, 37, 40, None), (8, 64, 896, "full"), (17, 300, 64, None),
                                        (1, 1, 40, "full"), (1, 2, 33, "full"), (2, 3, 896, None), (65, 9, 16, None)])
def test_bilstm_h256_forward_backward(dev, B, T, D, lens):
    g = torch.Generator().manual_seed(B * T)
    ref, ours = _oracle_pair(dev, D, 256, 2)
    x = torch.randn(B, T, D, generator=g)
    if lens == "full":
        lengths = torch.full((B,), T, dtype=torch.long)
    else:
        lengths = torch.randint(1, T + 1, (B,), generator=g)
        lengths[B // 2] = T
    y = (torch.rand(B, T, generator=g) < 0.1).float()
    for b, n in enumerate(lengths.tolist()):
        y[b, n:] = -1
    ref.th = ours.th = 0.5
    s_ref, tags_ref = ref(x, lengths)
    s, tags = ours(x.to(dev), lengths)
    close(s, s_ref, msg="scores")
    # thresholded tags must agree wherever the oracle's probability is not within 1e-6 of the threshold
    p = torch.sigmoid(s_ref)[:, :, 0]
    for b, n in enumerate(lengths.tolist()):
        for t in range(n):
            if abs(float(p[b, t]) - 0.5) > 1e-6:
                assert tags[b][t] == tags_ref[b][t]
    loss_ref = ref.loss(x, lengths, y)
    loss_ref.backward()
    loss = ours.loss(x.to(dev), lengths, y.to(dev))
    loss.backward()
    close(loss, loss_ref)
    ref_grads = dict(ref.named_parameters())
    for k, prm in ours.named_parameters():
        gr = ref_grads[k].grad
        tol = 1e-4 * float(gr.abs().max()) + 1e-7  # rtol 1e-4 of the tensor's scale
        close(prm.grad, gr, rtol=1e-4, atol=tol, msg=k)


def test_padding_invariance_and_permutation(dev):
    """Changing padded inputs leaves valid outputs bit-identical; permuting the batch permutes the outputs."""
    ref, ours = _oracle_pair(dev, 48, 256, 2, seed=3)
    g = torch.Generator().manual_seed(9)
    B, T = 19, 41
    x = torch.randn(B, T, 48, generator=g)
    lengths = torch.randint(1, T + 1, (B,), generator=g)
    lengths[0] = T
    s1, t1 = ours(x.to(dev), lengths)
    x2 = x.clone()
    for b, n in enumerate(lengths.tolist()):
        x2[b, n:] = 1e3
    s2, t2 = ours(x2.to(dev), lengths)
    assert torch.equal(s1, s2) and t1 == t2
    perm = torch.randperm(B, generator=g)
    s3, t3 = ours(x[perm].to(dev), lengths[perm])
    for i, j in enumerate(perm.tolist()):
        n = int(lengths[j])
        close(s3[i, :n], s1[j, :n], rtol=1e-5, atol=1e-6)
    # padded steps: hidden state 0 => logit == classifier bias
    bias = float(ours.classification.bias)
    for b, n in enumerate(lengths.tolist()):
        assert torch.all(s1[b, n:, 0] == bias)


def test_cfg1_full_size_properties(dev):
    """BASELINE configs[0] shape (64 x 300 x 896, H 256, 2 layers): size-independent checks + sampled oracle rows."""
    from oracle import ref_torch as rt  # noqa: F401

    ref, ours = _oracle_pair(dev, 896, 256, 2, seed=1)
    g = torch.Generator().manual_seed(1235)
    B, T = 64, 300
    x = torch.randn(B, T, 896, generator=g)
    lengths = torch.randint(84, 301, (B,), generator=g)
    lengths[5] = 300
    ours.th = ref.th = 0.5
    s, tags = ours((x[:, :, :384].contiguous().to(dev), x[:, :, 384:].contiguous().to(dev)), lengths)
    sub = [0, 5, 17, 63]
    s_ref, tags_ref = ref(x[sub], lengths[sub])
    for i, b in enumerate(sub):
        n = int(lengths[b])
        close(s[b, :n], s_ref[i, :n], msg=f"episode {b}")
    assert [len(t) for t in tags] == lengths.tolist()


def test_text_segmenter_steps(dev):
    """The reference-facing task module: training / validation / test / predict steps on a collated batch."""
    from multimodaltopicsegmentation_b200 import AudioPortionDataset, TextSegmenter, to_device

    g = torch.Generator().manual_seed(4)
    lines = []
    for n in (30, 12, 21, 7):
        labs = (torch.rand(n, generator=g) < 0.2).long().tolist()
        labs[-1] = 0
        lines.append((torch.randn(n, 20, generator=g), labs, "f"))
    ds = AudioPortionDataset(lines, {"0": 0, "1": 1}, CRF=False, truncate=False)
    batch = to_device(ds.collater([ds[i] for i in range(4)]), dev)
    torch.manual_seed(0)
    seg = TextSegmenter(2, 20, 256, num_layers=2, architecture="BiLSTM", loss_fn="FocalLoss", optimizer="Adam",
                        lr=1e-3, threshold=0.4).to(dev)
    opt = seg.configure_optimizers()["optimizer"]
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = seg.training_step(batch, 0)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    res = seg.test_step(batch, 0)
    assert set(res) == {"F1_loss", "WD_loss", "threshold", "test_loss"}
    tags = seg.predict_step(batch, 0)
    assert [len(t) for t in tags] == [30, 12, 21, 7]
    with pytest.raises(ValueError):
        TextSegmenter(2, 20, 256, architecture="nope")


@pytest.mark.parametrize("name,kw,double", [
    ("bilstm_pk", dict(architecture="BiLSTM", loss_fn="FocalLoss", metric="Pk", threshold=0.4), False),
    ("bilstm_f1_eb", dict(architecture="BiLSTM", loss_fn="BinaryCrossEntropy", metric="F1", threshold=0.5, end_boundary=True), False),
    ("bilstm_ce_wd", dict(architecture="BiLSTM", loss_fn="CrossEntropy", metric="WD", threshold=None), False),
    ("late_pk", dict(architecture="BiLSTMLateFusion", loss_fn="FocalLoss", metric="Pk", threshold=0.3), True),
])
def test_text_segmenter_steps_equal_the_reference_steps(dev, golden, name, kw, double):
    """training / validation / test / predict steps against the outputs of the reference's OWN `TextSegmenter`
    (models/lightning_model.py:179-760, run unmodified behind a pytorch_lightning / segeval stub by
    tests/golden/make_golden_steps.py): losses within rtol 1e-5, tags exact, and the logged result dict of test_step --
    keys, threshold, F1 (sklearn on the reference side), Pk / WindowDiff averages incl. the 1- and 3-sentence episodes
    where WindowDiff asserts and the reference substitutes Pk, and the end_boundary variant -- equal to 1e-12 on both the
    device-count path and the host walk."""
    from multimodaltopicsegmentation_b200 import TextSegmenter

    fx = golden("steps")
    emb = [10, 6] if double else 10
    seg = TextSegmenter(2, emb, 8, num_layers=2, optimizer="Adam", lr=1e-3, all_results=True, **kw)
    sd = {k[len(name) + 3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith(name + ":p:")}
    missing, unexpected = seg.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    seg = seg.to(dev)

    def batch():
        b = {"src_tokens": torch.from_numpy(fx[name + ":i:src_tokens"]).to(dev),
             "tgt_tokens": torch.from_numpy(fx[name + ":i:tgt_tokens"]).to(dev),
             "src_lengths": torch.from_numpy(fx[name + ":i:src_lengths"]), "src_tokens2": None}
        if double:
            b["src_tokens2"] = torch.from_numpy(fx[name + ":i:src_tokens2"]).to(dev)
        return b

    seg.train()
    loss = seg.training_step(batch(), 0)
    close(loss, fx[name + ":o:training_loss"], rtol=1e-5, atol=1e-7)
    close(seg.logged["training_loss"], fx[name + ":o:logged_training_loss"], rtol=1e-5, atol=1e-7)
    seg.eval()
    val = seg.validation_step(batch(), 0)
    close(val, fx[name + ":o:val_loss"], rtol=1e-5, atol=1e-7)
    assert float(seg.logged["threshold"]) == float(fx[name + ":o:val_threshold"])
    if not double:
        tags_equal(seg.predict_step(batch(), 0), fx[name + ":o:predict_tags"])
    keys = fx[name + ":o:result_keys"].tolist()
    want = fx[name + ":o:result_values"]
    for device_metrics in (True, False):
        seg.device_metrics = device_metrics
        seg.results = []
        res = seg.test_step(batch(), 0)
        assert sorted(res) == keys and seg.results[-1] is res
        got = np.asarray([float(res[k]) for k in keys])
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12, err_msg=f"{name} device_metrics={device_metrics} {keys}")
        assert sorted(k for k in seg.logged if k in keys) == keys
    opt = seg.configure_optimizers()
    assert [type(opt["optimizer"]).__name__, opt["lr_scheduler"]["monitor"], opt["lr_scheduler"]["scheduler"].mode] \
        == fx[name + ":o:optimizer"].tolist()


def test_f16_weight_pieces_kernel_equals_host_formulation(dev):
    """ops.f16_pieces_stacked (mts_pack_rows_f16 per matrix, what the LSTM weight cache runs after every optimiser step)
    against ops.f16_pieces (plain torch ops, CPU-tested in test_host_logic.py): identical bits, also for a zero row, rows
    spanning ten decades and a width that needs K padding; unsupported widths fall back to the torch formulation."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(11)
    for K in (384, 896, 72, 1024 + 64):
        mats = [torch.randn(1024, K, generator=g) * (10.0 ** (-10 * torch.rand(1024, 1, generator=g))) for _ in range(2)]
        mats[1][5] = 0.0
        dmats = [m.to(dev) for m in mats]
        p, sc = ops.f16_pieces_stacked(dmats)
        p_ref, sc_ref = ops.f16_pieces(torch.cat(dmats, dim=0))
        assert p.shape == p_ref.shape and torch.equal(p.view(torch.int16), p_ref.view(torch.int16)), K
        assert torch.equal(sc, sc_ref), K
    odd = [torch.randn(8, 10, generator=g).to(dev)]
    p, sc = ops.f16_pieces_stacked(odd)
    p_ref, sc_ref = ops.f16_pieces(odd[0])
    assert torch.equal(p.view(torch.int16), p_ref.view(torch.int16)) and torch.equal(sc, sc_ref)


# ----------------------------------------------------------------------------------------------------------
# pyramidal windowed-attention segmenter (HF LongformerModel in the reference)
# ----------------------------------------------------------------------------------------------------------
def load_params_allow_unused(module, fx, dev):
    """The transformer fixtures omit the tensors HF allocates but never reads (word_embeddings, *_global, pooler)."""
    sd = {k[2:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("p:")}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(("word_embeddings" in k or "_global" in k or "pooler" in k) for k in missing), missing
    return module.to(dev)


def _to_ragged(t, lens, B, S):
    """[B*S, c] padded rows -> [sum(len), c] valid rows."""
    return torch.cat([t.view(B, S, -1)[b, :n] for b, n in enumerate(lens)], dim=0).contiguous()


def _to_padded(t, lens, B, S):
    out = torch.zeros(B, S, t.shape[1], dtype=t.dtype, device=t.device)
    o = 0
    for b, n in enumerate(lens):
        out[b, :n] = t[o:o + n]
        o += n
    return out.view(B * S, -1)


@pytest.mark.parametrize("layout", ["padded", "ragged"])
@pytest.mark.parametrize("entry", ["mts_band_attn_fwd", "mts_band_attn_fwd_tc", "mts_band_attn_fwd_simt", "mts_band_attn_fwd_mma"])
@pytest.mark.parametrize("B,S,h,hd,w,lens", [
    (2, 24, 4, 8, 4, [24, 13]),
    (3, 200, 2, 112, 48, [200, 77, 1]),
    (2, 130, 3, 64, 8, [130, 31]),
    (2, 96, 2, 32, 100, [96, 50]),      # window wider than the episode: dense attention
    (1, 700, 1, 16, 360, [650]),        # default-config reach (window 120 x 6 layers)
    (5, 960, 8, 112, 48, [960, 100, 513, 128, 129]),   # configs[2] geometry: many work items per CTA, block-edge lengths
    (3, 300, 2, 128, 0, [300, 1, 64]),  # reach 0: every token attends to itself only
    (2, 260, 1, 112, 40, [260, 257]),
])
def test_band_attention_forward(dev, B, S, h, hd, w, lens, entry, layout):
    """The tcgen05 kernel (the default), the CUDA-core kernel and the mma.sync kernel against the numpy restatement of
    HF's banded attention, in the reference's padded row layout and in the ragged layout (valid sentences only;
    include/mts_b200.h "Row layouts")."""
    from multimodaltopicsegmentation_b200 import _lib, ops
    from oracle import ref_numpy as rn

    if entry == "mts_band_attn_fwd_tc" and not _lib.load().mts_band_attn_tc_supported(hd):
        pytest.skip("the tcgen05 kernel is instantiated for head dims 16, 32, 64, 112, 128")
    if entry == "mts_band_attn_fwd_mma" and hd not in (8, 16, 32, 64, 112, 128):
        pytest.skip("mma.sync kernel: head dims 8, 16, 32, 64, 112, 128")

    g = torch.Generator().manual_seed(S + hd + w)
    d = h * hd
    qkv = torch.randn(B * S, 3 * d, generator=g)
    L = ops.Lengths(lens, dev, S)
    ragged = layout == "ragged"
    offs = L.offs.data_ptr() if ragged else 0
    qkv_d = (_to_ragged(qkv, lens, B, S) if ragged else qkv).to(dev)
    rows = qkv_d.shape[0]
    out = torch.full((rows + 1, d), 7.0, device=dev)  # one guard row behind the last valid one
    lse = torch.empty(B, h, S, device=dev)
    ops._call(entry, qkv_d.data_ptr(), 3 * d, L.dev.data_ptr(), offs, B, S, h, hd, w, out.data_ptr(), 0, 0, 0,
              lse.data_ptr(), ops._stream())
    assert float((out[rows] - 7.0).abs().max()) == 0.0
    out = out[:rows]
    split = lambda t: t.view(B, S, h, hd).permute(0, 2, 1, 3).numpy()
    q = split(qkv[:, :d]) / np.float32(np.sqrt(hd))
    ref = rn.banded_attention(q.astype(np.float32), split(qkv[:, d:2 * d]), split(qkv[:, 2 * d:]), lens, w)
    full = _to_padded(out, lens, B, S) if ragged else out
    got = full.view(B, S, h, hd).permute(0, 2, 1, 3)
    close(got, ref, rtol=1e-4, atol=1e-5)
    if not ragged:
        for b, n in enumerate(lens):  # padded queries: exact zeros (HF modeling_longformer.py:578)
            assert float(out.view(B, S, d)[b, n:].abs().max() if n < S else 0.0) == 0.0
    if d % 32 == 0:  # fused operand split
        hl = torch.empty(2, rows, d, device=dev)
        ops._call(entry, qkv_d.data_ptr(), 3 * d, L.dev.data_ptr(), offs, B, S, h, hd, w, 0, hl[0].data_ptr(),
                  hl[1].data_ptr(), d, 0, ops._stream())
        check_operand_pair(hl[0], hl[1], out, side=0)


@pytest.fixture(params=["ragged", "padded"])
def xf_layout(request, monkeypatch):
    """Both token layouts of the encoder (transformer.LAYOUT): ragged = valid sentences only (default), padded = the
    reference's [B*S] rows."""
    from multimodaltopicsegmentation_b200 import transformer

    monkeypatch.setattr(transformer, "LAYOUT", request.param)
    return request.param


def test_transformer_golden_forward(dev, golden, xf_layout):
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter

    fx = golden("transformer_focal")
    nh, w = int(fx["i:nheads"]), int(fx["i:window"])
    m = load_params_allow_unused(Transformer_segmenter(2, 32, 16, num_layers=2, nheads=nh, loss_fn="FocalLoss",
                                                        window_size=w), fx, dev).eval()
    x = torch.from_numpy(fx["i:x"]).to(dev)
    lengths = torch.from_numpy(fx["i:lengths"])
    m.th = float(fx["i:th"])
    hidden = m.model(x, lengths)
    scores, tags = m(x, lengths)
    assert tuple(scores.shape) == fx["o:scores"].shape  # time axis = S for the transformer
    if xf_layout == "padded":  # the reference's layout: padded positions carry the reference's values too
        close(hidden, fx["o:hidden"])
        close(scores, fx["o:scores"], atol=1e-4)  # classification weights were scaled x20 in the fixture
    else:  # valid sentences identical; padded sentences are exact zeros (they have no rows inside the encoder)
        for b, n in enumerate(lengths.tolist()):
            close(hidden[b, :n], fx["o:hidden"][b, :n])
            close(scores[b, :n], fx["o:scores"][b, :n], atol=1e-4)
            assert float(hidden[b, n:].abs().max() if n < hidden.shape[1] else 0.0) == 0.0
    tags_equal(tags, fx["o:tags"])
    with torch.no_grad():
        loss = m.loss(x, lengths, torch.from_numpy(fx["i:y"]).to(dev))
    close(loss, fx["o:loss"])


def test_transformer_vs_hf_twin(dev, xf_layout):
    """Medium shapes against the HF LongformerModel twin (oracle/ref_torch.py), ragged lengths."""
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter
    from oracle import ref_torch as rt

    torch.manual_seed(11)
    g = torch.Generator().manual_seed(12)
    B, S, d, F, nl, nh, w = 3, 96, 64, 48, 3, 4, 8   # windows [24, 16, 8]; S is a multiple of all of them
    ref = rt.WindowedSegmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="BinaryCrossEntropy", window_size=w).eval()
    ours = Transformer_segmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="BinaryCrossEntropy", window_size=w)
    sd = {k: v for k, v in ref.state_dict().items()}
    missing, unexpected = ours.load_state_dict(sd, strict=False)
    assert not missing, missing
    assert all("position_ids" in k for k in unexpected), unexpected   # HF buffers, if any
    ours = ours.to(dev).eval()
    x = torch.randn(B, S, d, generator=g)
    lengths = torch.tensor([96, 40, 7])
    y = (torch.rand(B, S, generator=g) < 0.2).float()
    ref.th = ours.th = 0.5
    with torch.no_grad():
        s_ref, t_ref = ref(x, lengths)
        l_ref = ref.loss(x, lengths, y)
        s, t = ours(x.to(dev), lengths)
        l = ours.loss(x.to(dev), lengths, y.to(dev))
    for b, n in enumerate(lengths.tolist()):
        close(s[b, :n], s_ref[b, :n], atol=2e-5)
    assert t == t_ref
    close(l, l_ref)


def _dense_band_attention_torch(qkv, lens, B, S, h, hd, w, prob_scale=None):
    """float64 autograd reference: dense scores with the band / length mask, zero rows for padded queries;
    prob_scale [B, h, S, S]: multiplier of the probabilities after the softmax (dropout keep / (1 - p))."""
    d = h * hd
    q, k, v = [t.view(B, S, h, hd).permute(0, 2, 1, 3) for t in (qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:])]
    s = (q / np.sqrt(hd)) @ k.transpose(-1, -2)
    idx = torch.arange(S)
    band = (idx[:, None] - idx[None, :]).abs() <= w
    outs = []
    for b in range(B):
        n = lens[b]
        allowed = band & (idx[None, :] < n)
        sb = s[b].masked_fill(~allowed, float("-inf"))
        sb = torch.where(allowed.any(-1, keepdim=True), sb, torch.zeros_like(sb))
        p = torch.softmax(sb, dim=-1) * (idx < n)[None, :, None]
        if prob_scale is not None:
            p = p * prob_scale[b]
        outs.append(p @ v[b])
    return torch.stack(outs).permute(0, 2, 1, 3).reshape(B * S, d)


@pytest.mark.parametrize("B,S,h,hd,w,lens", [
    (2, 24, 4, 8, 4, [24, 13]),
    (2, 150, 2, 112, 48, [150, 61]),
    (2, 130, 3, 64, 8, [130, 31]),
    (1, 300, 1, 16, 120, [280]),
])
@pytest.mark.parametrize("layout", ["padded", "ragged"])
def test_band_attention_backward(dev, B, S, h, hd, w, lens, layout):
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator().manual_seed(S + hd + w + 1)
    d = h * hd
    qkv = torch.randn(B * S, 3 * d, generator=g, dtype=torch.float64).requires_grad_(True)
    do = torch.randn(B * S, d, generator=g, dtype=torch.float64)
    ragged = layout == "ragged"
    if ragged:  # the ragged layout has no rows for padded sentences: their upstream gradient does not exist
        for b, n in enumerate(lens):
            do.view(B, S, d)[b, n:] = 0
    ref = _dense_band_attention_torch(qkv, lens, B, S, h, hd, w)
    ref.backward(do)
    L = ops.Lengths(lens, dev, S)
    offs = L.offs.data_ptr() if ragged else 0
    pack = (lambda t: _to_ragged(t, lens, B, S)) if ragged else (lambda t: t)
    qkv_d = pack(qkv.detach().float()).to(dev)
    rows = qkv_d.shape[0]
    out = torch.empty(rows, d, device=dev)
    lse = torch.empty(B, h, S, device=dev)
    ops._call("mts_band_attn_fwd", qkv_d.data_ptr(), 3 * d, L.dev.data_ptr(), offs, B, S, h, hd, w, out.data_ptr(), 0, 0, 0,
              lse.data_ptr(), ops._stream())
    close(out, pack(ref.detach().float()), rtol=1e-4, atol=1e-5)
    dqkv = torch.full((rows + 1, 3 * d), float("nan"), device=dev)
    delta = torch.empty(B, h, S, device=dev)
    do_d = pack(do.float()).to(dev)
    ops._call("mts_band_attn_bwd", qkv_d.data_ptr(), 3 * d, out.data_ptr(), do_d.data_ptr(), lse.data_ptr(),
              L.dev.data_ptr(), offs, B, S, h, hd, w, dqkv.data_ptr(), delta.data_ptr(), ops._stream())
    assert bool(torch.isnan(dqkv[rows]).all())  # nothing written behind the last row
    scale = float(qkv.grad.abs().max())
    close(dqkv[:rows], pack(qkv.grad.float()), rtol=1e-4, atol=2e-5 * scale)


@pytest.mark.parametrize("B,S,h,hd,w,lens,p", [
    (2, 24, 4, 8, 4, [24, 13], 0.5),
    (2, 150, 2, 112, 48, [150, 61], 0.1),
    (3, 130, 3, 64, 8, [130, 31, 1], 0.1),
    (1, 300, 1, 16, 120, [280], 0.25),
])
@pytest.mark.parametrize("layout", ["padded", "ragged"])
def test_band_attention_dropout_forward_backward(dev, B, S, h, hd, w, lens, p, layout):
    """Dropout on the attention probabilities (HF attention_probs_dropout_prob; the reference sets it from dropout_out,
    models/CRF.py:531-536): forward output, saved log-sum-exp and dq / dk / dv of mts_band_attn_fwd_dropout /
    mts_band_attn_bwd_dropout against float64 autograd through the dense masked softmax multiplied by the SAME keep-mask,
    which the oracle regenerates from the seed (oracle/ref_numpy.py:attn_dropout_keep restates the device hash)."""
    from multimodaltopicsegmentation_b200 import ops
    from oracle import ref_numpy as rn

    g = torch.Generator().manual_seed(S + hd + w + 7)
    seed = 0x1234_5678_9ABC_DEF0 ^ (S * 7919 + hd)
    d = h * hd
    keep = rn.attn_dropout_keep(seed, B, h, S, p)
    frac = 1.0 - keep.mean()
    assert abs(frac - p) < 4 * np.sqrt(p * (1 - p) / keep.size) + 1e-3, (frac, p)
    scale_mask = torch.from_numpy(keep.astype(np.float64)) / (1.0 - float(np.float32(p)))
    qkv = torch.randn(B * S, 3 * d, generator=g, dtype=torch.float64).requires_grad_(True)
    do = torch.randn(B * S, d, generator=g, dtype=torch.float64)
    ragged = layout == "ragged"
    if ragged:
        for b, n in enumerate(lens):
            do.view(B, S, d)[b, n:] = 0
    ref = _dense_band_attention_torch(qkv, lens, B, S, h, hd, w, prob_scale=scale_mask)
    ref.backward(do)
    plain = _dense_band_attention_torch(qkv.detach(), lens, B, S, h, hd, w)
    assert float((ref.detach() - plain).abs().max()) > 1e-3     # the mask does something
    L = ops.Lengths(lens, dev, S)
    offs = L.offs.data_ptr() if ragged else 0
    pack = (lambda t: _to_ragged(t, lens, B, S)) if ragged else (lambda t: t)
    qkv_d = pack(qkv.detach().float()).to(dev)
    rows = qkv_d.shape[0]
    out, out0 = torch.empty(rows, d, device=dev), torch.empty(rows, d, device=dev)
    lse, lse0 = torch.empty(B, h, S, device=dev), torch.empty(B, h, S, device=dev)
    ops._call("mts_band_attn_fwd_dropout", qkv_d.data_ptr(), 3 * d, L.dev.data_ptr(), offs, B, S, h, hd, w, out.data_ptr(), 0, 0, 0,
              lse.data_ptr(), p, seed, ops._stream())
    close(out, pack(ref.detach().float()), rtol=1e-4, atol=2e-5)
    ops._call("mts_band_attn_fwd_simt", qkv_d.data_ptr(), 3 * d, L.dev.data_ptr(), offs, B, S, h, hd, w, out0.data_ptr(), 0, 0, 0,
              lse0.data_ptr(), ops._stream())
    assert torch.equal(lse, lse0)                               # the statistics are those of the undropped scores
    out_p0 = torch.empty(rows, d, device=dev)
    ops._call("mts_band_attn_fwd_dropout", qkv_d.data_ptr(), 3 * d, L.dev.data_ptr(), offs, B, S, h, hd, w, out_p0.data_ptr(), 0, 0, 0,
              0, 0.0, seed, ops._stream())
    assert torch.equal(out_p0, out0)                            # p = 0 is the plain kernel
    dqkv = torch.full((rows + 1, 3 * d), float("nan"), device=dev)
    delta = torch.empty(B, h, S, device=dev)
    do_d = pack(do.float()).to(dev)
    ops._call("mts_band_attn_bwd_dropout", qkv_d.data_ptr(), 3 * d, out.data_ptr(), do_d.data_ptr(), lse.data_ptr(),
              L.dev.data_ptr(), offs, B, S, h, hd, w, dqkv.data_ptr(), delta.data_ptr(), p, seed, ops._stream())
    assert bool(torch.isnan(dqkv[rows]).all())
    scale = float(qkv.grad.abs().max())
    close(dqkv[:rows], pack(qkv.grad.float()), rtol=1e-4, atol=2e-5 * scale)


def test_attention_dropout_in_training(dev, monkeypatch):
    """Transformer_segmenter(dropout_out = p) and the BiLSTMRestrictedMHA block (wrapper default 0.1, as in the reference)
    train with dropped attention probabilities: repeatable under a fixed seed function, different from call to call with
    the default generator, absent in eval mode; the gradient is the gradient of the dropped forward pass (directional
    finite difference on the loss under a fixed seed)."""
    from multimodaltopicsegmentation_b200 import RecurrentLongformer, transformer
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter

    g = torch.Generator().manual_seed(5)
    B, S, d, F, nh, w = 3, 48, 32, 16, 4, 8
    x = torch.randn(B, S, d, generator=g).to(dev)
    lengths = torch.tensor([48, 20, 33])
    y = (torch.rand(B, S, generator=g) < 0.2).float()
    for b, n in enumerate(lengths.tolist()):
        y[b, n:] = -1
    y = y.to(dev)
    torch.manual_seed(3)
    models = [Transformer_segmenter(2, d, F, num_layers=2, nheads=nh, loss_fn="FocalLoss", window_size=w, dropout_out=0.3).to(dev),
              RecurrentLongformer(2, d, 16, num_layers=2, nheads=nh, loss_fn="FocalLoss", window_size=w).to(dev)]
    for m in models:
        m.train()
        monkeypatch.setattr(transformer, "ATTN_SEED_FN", None)
        with torch.no_grad():
            a, b_ = float(m.loss(x, lengths, y)), float(m.loss(x, lengths, y))
            assert a != b_
        monkeypatch.setattr(transformer, "ATTN_SEED_FN", lambda layer: 1000 + layer)
        loss = m.loss(x, lengths, y)
        loss.backward()
        with torch.no_grad():
            assert float(m.loss(x, lengths, y)) == float(loss)
            # directional derivative along the gradient, central difference (same seeds -> same masks)
            params = [q for q in m.parameters() if q.grad is not None]
            gnorm2 = sum(float((q.grad.double() ** 2).sum()) for q in params)
            eps = 1e-2 / np.sqrt(gnorm2)
            for sgn in (+1, -1):
                for q in params:
                    q.add_(sgn * eps * q.grad)
                if sgn > 0:
                    lp = float(m.loss(x, lengths, y).double())
                    for q in params:
                        q.sub_(eps * q.grad)
                else:
                    lm = float(m.loss(x, lengths, y).double())
                    for q in params:
                        q.add_(eps * q.grad)
            fd = (lp - lm) / (2 * eps)
            print(f"  {type(m).__name__}: directional derivative {fd:.6e} vs |grad|^2 {gnorm2:.6e}")
            assert abs(fd - gnorm2) <= 2e-2 * gnorm2
            m.eval()
            monkeypatch.setattr(transformer, "ATTN_SEED_FN", None)
            assert float(m.loss(x, lengths, y)) == float(m.loss(x, lengths, y))


@pytest.mark.parametrize("M,d", [(37, 32), (1000, 896), (300, 1024), (64, 100)])
def test_layer_norm_forward_backward(dev, M, d):
    from multimodaltopicsegmentation_b200 import _lib, ops

    g = torch.Generator().manual_seed(M + d)
    a = torch.randn(M, d, generator=g, dtype=torch.float64)
    r = torch.randn(M, d, generator=g, dtype=torch.float64)
    gamma = torch.randn(d, generator=g, dtype=torch.float64).requires_grad_(True)
    beta = torch.randn(d, generator=g, dtype=torch.float64).requires_grad_(True)
    pre = (a + r).requires_grad_(True)
    ref = torch.nn.functional.layer_norm(pre, (d,), gamma, beta, eps=1e-12)
    dy = torch.randn(M, d, generator=g, dtype=torch.float64)
    ref.backward(dy)
    f = lambda t: t.detach().float().to(dev).contiguous()
    kp = (d + 31) // 32 * 32
    y, hl = torch.empty(M, d, device=dev), torch.empty(2, M, kp, device=dev)
    pre_d, stats = torch.empty(M, d, device=dev), torch.empty(M, 2, device=dev)
    a_d, r_d, g_d, b_d, dy_d = f(a), f(r), f(gamma), f(beta), f(dy)  # keep the device copies alive across the calls
    ops._call("mts_add_ln_fwd", a_d.data_ptr(), r_d.data_ptr(), g_d.data_ptr(), b_d.data_ptr(), M, d, 1e-12,
              y.data_ptr(), hl[0].data_ptr(), hl[1].data_ptr(), kp, pre_d.data_ptr(), stats.data_ptr(), ops._stream())
    close(y, ref.detach().float(), rtol=1e-4, atol=1e-5)
    check_operand_pair(hl[0], hl[1], y, side=0)
    dx, dhl = torch.empty(M, d, device=dev), torch.empty(2, M, kp, device=dev)
    dgb = torch.empty(2, d, device=dev)
    ws = torch.empty(_lib.load().mts_ln_bwd_ws_bytes(M, d) // 4, device=dev)
    ops._call("mts_ln_bwd", dy_d.data_ptr(), pre_d.data_ptr(), stats.data_ptr(), g_d.data_ptr(), M, d,
              dx.data_ptr(), dhl[0].data_ptr(), dhl[1].data_ptr(), kp, dgb[0].data_ptr(), dgb[1].data_ptr(),
              ws.data_ptr(), ops._stream())
    close(dx, pre.grad.float(), rtol=1e-4, atol=1e-5 * float(pre.grad.abs().max()))
    close(dgb[0], gamma.grad.float(), rtol=1e-4, atol=1e-5 * float(gamma.grad.abs().max()))
    close(dgb[1], beta.grad.float(), rtol=1e-4, atol=1e-5 * float(beta.grad.abs().max()))
    check_operand_pair(dhl[0], dhl[1], dx, side=0)


def test_transformer_golden_backward(dev, golden, xf_layout):
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter

    fx = golden("transformer_focal")
    nh, w = int(fx["i:nheads"]), int(fx["i:window"])
    m = load_params_allow_unused(Transformer_segmenter(2, 32, 16, num_layers=2, nheads=nh, loss_fn="FocalLoss",
                                                        window_size=w), fx, dev)
    x = torch.from_numpy(fx["i:x"]).to(dev)
    loss = m.loss(x, torch.from_numpy(fx["i:lengths"]), torch.from_numpy(fx["i:y"]).to(dev))
    loss.backward()
    close(loss, fx["o:loss"])
    named = dict(m.named_parameters())
    for k in fx.files:
        if not k.startswith("g:"):
            continue
        ref = fx[k]
        got = named[k[2:]].grad
        assert got is not None, k
        close(got, ref, rtol=2e-4, atol=2e-6 + 1e-4 * float(np.abs(ref).max()), msg=k)
    for k, p in named.items():  # tensors HF never reads receive no gradient
        if "word_embeddings" in k or "_global" in k or "pooler" in k:
            assert p.grad is None, k


def test_transformer_training_vs_hf_twin(dev, xf_layout):
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter
    from oracle import ref_torch as rt

    torch.manual_seed(21)
    g = torch.Generator().manual_seed(22)
    B, S, d, F, nl, nh, w = 3, 96, 64, 48, 3, 4, 8
    ref = rt.WindowedSegmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w).train()
    ours = Transformer_segmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w)
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=False)
    assert not missing, missing
    ours = ours.to(dev).train()
    x = torch.randn(B, S, d, generator=g)
    lengths = torch.tensor([96, 40, 7])
    y = (torch.rand(B, S, generator=g) < 0.2).float()
    l_ref = ref.loss(x, lengths, y)
    l_ref.backward()
    l = ours.loss(x.to(dev), lengths, y.to(dev))
    l.backward()
    close(l, l_ref)
    ref_named = dict(ref.named_parameters())
    checked = 0
    for k, p in ours.named_parameters():
        r = ref_named[k].grad
        if r is None or float(r.abs().max()) == 0.0:
            continue
        if k.endswith("attention.self.key.bias"):
            # softmax is invariant to a per-query constant, so d loss / d b_k is exactly 0 in exact arithmetic:
            # both sides hold rounding noise only
            assert float(p.grad.abs().max()) < 1e-9 and float(r.abs().max()) < 1e-9
            continue
        close(p.grad, r, rtol=2e-4, atol=1e-4 * float(r.abs().max()), msg=k)
        checked += 1
    assert checked >= 4 + 15 * nl + 2


def test_transformer_inference_fusions_are_bit_identical(dev, xf_layout, monkeypatch):
    """Inference puts the GELU + operand split into the FFN1 epilogue and can fold the residual adds into the dense layers'
    epilogues (MTS_XF_FOLD=1; LayerNorm then reads one tensor); training keeps the separate tensors.  Same arithmetic in the same order:
    the hidden states must agree bit for bit with the unfused inference path and with the training-mode forward."""
    from multimodaltopicsegmentation_b200 import ops, transformer
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter

    torch.manual_seed(5)
    B, S, d, F, nl, nh, w = 4, 120, 64, 96, 3, 4, 8
    m = Transformer_segmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w).to(dev).eval()
    x = torch.randn(B, S, d, device=dev)
    lengths = torch.tensor([120, 64, 1, 33])
    with torch.no_grad():
        monkeypatch.setattr(transformer, "FOLD_RESIDUAL", True)
        fused = m.model(x, lengths)
        monkeypatch.setattr(transformer, "FOLD_RESIDUAL", False)
        unfolded = m.model(x, lengths)
    saved = transformer.encoder_forward(x, ops.Lengths(lengths, dev, S), m.model.packed(), nh, m.model.reaches, save=True)[0]
    assert torch.equal(fused, unfolded)
    assert torch.equal(fused, saved)


def test_transformer_hidden_dropout_training_vs_hf_twin(dev, xf_layout, monkeypatch):
    """dropout_in > 0 (HF hidden_dropout_prob: after the embeddings LayerNorm and on both dense outputs that feed a
    residual LayerNorm) under training: loss and every gradient against HF with the SAME keep-masks replayed on both
    sides (RNG parity with torch's CPU generator is not a goal, mask placement and scaling are).  Also: eval ignores
    the probability, and attention-probability dropout stays refused."""
    from multimodaltopicsegmentation_b200 import transformer
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter
    from oracle import ref_torch as rt

    torch.manual_seed(31)
    g = torch.Generator().manual_seed(32)
    B, S, d, F, nl, nh, w, p = 3, 96, 64, 48, 3, 4, 8, 0.25
    ref = rt.WindowedSegmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w, dropout_in=p).train()
    ours = Transformer_segmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w, dropout_in=p)
    missing, _ = ours.load_state_dict(ref.state_dict(), strict=False)
    assert not missing, missing
    ours = ours.to(dev).train()
    x = torch.randn(B, S, d, generator=g)
    lengths = torch.tensor([96, 40, 7])
    y = (torch.rand(B, S, generator=g) < 0.2).float()
    keep = [(torch.rand(B, S, d, generator=g) >= p) for _ in range(1 + 2 * nl)]

    class Replay(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, t):
            assert t.shape == self.m.shape
            return t * self.m / (1.0 - p)

    hf = ref.model.model
    assert isinstance(hf.embeddings.dropout, torch.nn.Dropout) and hf.embeddings.dropout.p == p
    hf.embeddings.dropout = Replay(keep[0])
    for l, lyr in enumerate(hf.encoder.layer):
        assert lyr.attention.output.dropout.p == p and lyr.output.dropout.p == p
        lyr.attention.output.dropout = Replay(keep[1 + 2 * l])
        lyr.output.dropout = Replay(keep[2 + 2 * l])
    used = []

    def replay(site, rows, width, prob, device):
        assert prob == p and width == d
        m = keep[site]
        m = torch.cat([m[b, :n] for b, n in enumerate(lengths.tolist())]) if xf_layout == "ragged" else m.reshape(B * S, d)
        assert m.shape[0] == rows
        used.append(site)
        return m.to(device)

    monkeypatch.setattr(transformer, "DROPOUT_MASK_FN", replay)
    l_ref = ref.loss(x, lengths, y)
    l_ref.backward()
    l = ours.loss(x.to(dev), lengths, y.to(dev))
    l.backward()
    assert used == list(range(1 + 2 * nl))
    close(l, l_ref)
    ref_named = dict(ref.named_parameters())
    checked = 0
    for k, prm in ours.named_parameters():
        r = ref_named[k].grad
        if r is None or float(r.abs().max()) == 0.0:
            continue
        if k.endswith("attention.self.key.bias"):
            assert float(prm.grad.abs().max()) < 1e-9 and float(r.abs().max()) < 1e-9
            continue
        close(prm.grad, r, rtol=2e-4, atol=1e-4 * float(r.abs().max()), msg=k)
        checked += 1
    assert checked >= 4 + 15 * nl + 2

    # the default generator path: a different mask every call, same expectation scale; eval is dropout-free
    monkeypatch.setattr(transformer, "DROPOUT_MASK_FN", None)
    with torch.no_grad():
        a = ours.model(x.to(dev), lengths)
        b = ours.model(x.to(dev), lengths)
        assert not torch.equal(a, b)
        ours.eval()
        c = ours.model(x.to(dev), lengths)
        assert torch.equal(c, ours.model(x.to(dev), lengths))


# ----------------------------------------------------------------------------------------------------------
# tensor-core recurrence (tcgen05, W_hh in tensor memory) against the exact-fp32 FMA kernel and float64
# ----------------------------------------------------------------------------------------------------------
# default = fp16-split kernels (pair kernel from ~113 sequences per direction on); the TF32 + bf16 kernel; the pair kernel at every size
@pytest.mark.parametrize("entry", ["mts_lstm_rec_fwd_tc", "mts_lstm_rec_fwd_tf32", "mts_lstm_rec_fwd_h3p"])
@pytest.mark.parametrize("B,T,n_enc,save", [(3, 5, 1, False), (16, 40, 1, True), (37, 61, 1, True), (20, 33, 2, True),
                                            (300, 50, 1, False), (70, 30, 1, False), (1000, 9, 1, False)])
def test_recurrence_tensor_core_vs_fma(dev, B, T, n_enc, save, entry):
    from multimodaltopicsegmentation_b200 import ops

    H = 256
    g = torch.Generator(device=dev).manual_seed(B * 1000 + T)
    gx = torch.randn((n_enc, B * T, 8 * H), device=dev, generator=g)
    whh = torch.randn((n_enc, 2, 4 * H, H), device=dev, generator=g) * 0.06
    if B == 70:   # rows of very different magnitude, one zero row: the fp16-split kernel scales every row by its own power of two
        whh = whh * (10.0 ** (torch.rand((n_enc, 2, 4 * H, 1), device=dev, generator=g) * 5.0 - 4.0))
        whh[:, :, 5, :] = 0.0
    hg = torch.Generator().manual_seed(B + T)
    lengths = [T] + [int(v) for v in torch.randint(1, T + 1, (B - 1,), generator=hg)]
    lens = ops.Lengths(lengths, dev, T)

    def run(name):
        y = torch.full((B, T, n_enc * 2 * H), float("nan"), device=dev)
        gates = torch.full((n_enc, 2, B, T, 5, H), float("nan"), device=dev) if save else None
        ops._call(name, gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), n_enc, B, T, H,
                  y.data_ptr(), 0 if gates is None else gates.data_ptr(),
                  *(() if name == "mts_lstm_rec_fwd" else (0, 0) if name.endswith("_h3p") else (0,)), ops._stream())
        return y, gates

    y_f, g_f = run("mts_lstm_rec_fwd")
    y_t, g_t = run(entry)
    if n_enc == 1:  # fused operand of the next layer's input projection: (y, y_corr) is a valid A-side operand pair
        y2 = torch.full((B, T, 2 * H), float("nan"), device=dev)
        corr = torch.full((B * T, 2 * H), float("nan"), device=dev)
        ops._call(entry, gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), 1, B, T, H,
                  y2.data_ptr(), 0, corr.data_ptr(), *((0,) if entry.endswith("_h3p") else ()), ops._stream())
        assert torch.equal(y2, y_t)
        check_operand_pair(y2.view(B * T, 2 * H), corr, y2.view(B * T, 2 * H), side=0)
    assert not bool(torch.isnan(y_t).any())          # every position written (zeros beyond len_b)
    for b, n in enumerate(lengths):
        assert float(y_t[b, n:].abs().max() if n < T else 0.0) == 0.0
    close(y_t, y_f, rtol=1e-4, atol=1e-5)
    if save:
        assert bool((torch.isnan(g_t) == torch.isnan(g_f)).all())   # same (valid-step) coverage of the saved gates
        ok = ~torch.isnan(g_f)
        close(g_t[ok], g_f[ok], rtol=1e-4, atol=2e-5)


# ----------------------------------------------------------------------------------------------------------
# full-size configurations: size-independent properties (the CPU oracle would take minutes at these sizes)
# ----------------------------------------------------------------------------------------------------------
def test_cfg3_full_size_properties(dev):
    """BASELINE configs[2] shape at reduced batch (32 of 256 episodes x 960 sentences x 896, 6 layers, window 16):
    episode independence (an episode alone == inside the batch), invariance to whatever sits in the padded rows, and
    a spot check of 2 episodes against the HF twin on the CPU."""
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter
    from oracle import ref_torch as rt

    torch.manual_seed(3)
    g = torch.Generator().manual_seed(33)
    B, S, d, F, nl, nh, w = 32, 960, 896, 256, 6, 8, 16
    ours = Transformer_segmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w).to(dev).eval()
    ours.th = 0.5
    lengths = torch.randint(100, S + 1, (B,), generator=g)
    lengths[0], lengths[1] = S, 333
    x = torch.randn(B, S, d, generator=g).to(dev)
    s_all, t_all = ours(x, lengths)
    assert tuple(s_all.shape) == (B, S, 1) and [len(t) for t in t_all] == lengths.tolist()
    # (1) padded rows of the INPUT never influence valid outputs
    x2 = x.clone()
    for b, n in enumerate(lengths.tolist()):
        x2[b, n:] = 1e3
    s_pad, t_pad = ours(x2, lengths)
    for b, n in enumerate(lengths.tolist()):
        assert torch.equal(s_pad[b, :n], s_all[b, :n])
    assert t_pad == t_all
    # (2) an episode alone gives the same scores as inside the batch (same kernels, same arithmetic per row)
    for b in (1, 7):
        s_one, t_one = ours(x[b:b + 1].contiguous(), lengths[b:b + 1])
        n = int(lengths[b])
        close(s_one[0, :n], s_all[b, :n], rtol=1e-5, atol=1e-6)
    # (3) spot check against the HF LongformerModel twin on the host (2 episodes)
    ref = rt.WindowedSegmenter(2, d, F, num_layers=nl, nheads=nh, loss_fn="FocalLoss", window_size=w).eval()
    missing, unexpected = ref.load_state_dict({k: v.cpu() for k, v in ours.state_dict().items()}, strict=False)
    assert not unexpected, unexpected
    ref.th = 0.5
    with torch.no_grad():
        s_ref, t_ref = ref(x[:2].cpu(), lengths[:2])
    for b in range(2):
        n = int(lengths[b])
        close(s_all[b, :n], s_ref[b, :n], rtol=1e-4, atol=3e-5)
    assert sum(int(a != r) for ta, tr in zip(t_all[:2], t_ref) for a, r in zip(ta, tr)) == 0


def test_cfg5_long_episode_properties(dev):
    """BASELINE configs[4] shape: 8 episodes x 8192 sentences x 1024-d, BiLSTM(+CRF) inference.  Permutation
    equivariance over episodes, padding invariance, Viterbi path score == gold score of the returned path."""
    from multimodaltopicsegmentation_b200 import BiLSTM, BiRnnCrf, ops

    torch.manual_seed(5)
    g = torch.Generator().manual_seed(55)
    B, T, D, H = 8, 8192, 1024, 256
    lengths = torch.randint(3000, T + 1, (B,), generator=g)
    lengths[3] = T
    x = torch.randn(B, T, D, generator=g).to(dev)
    m = BiLSTM(2, D, H, num_layers=2, loss_fn="FocalLoss").to(dev)
    m.th = 0.5
    s, tags = m(x, lengths)
    assert tuple(s.shape) == (B, T, 1) and [len(t) for t in tags] == lengths.tolist()
    perm = torch.tensor([5, 2, 7, 0, 3, 6, 1, 4])
    s_p, tags_p = m(x[perm].contiguous(), lengths[perm])
    for i, b in enumerate(perm.tolist()):
        n = int(lengths[b])
        assert torch.equal(s_p[i, :n], s[b, :n])   # same tile arithmetic whatever the batch order
        assert tags_p[i] == tags[b]
    x2 = x.clone()
    for b, n in enumerate(lengths.tolist()):
        x2[b, n:] = -77.0
    s2, _ = m(x2, lengths)
    for b, n in enumerate(lengths.tolist()):
        assert torch.equal(s2[b, :n], s[b, :n])
    # CRF over the same encoder: the decoded path must score exactly what Viterbi reported
    c = BiRnnCrf(2, D, H, num_layers=2).to(dev)
    c.model.load_state_dict(m.model.state_dict())
    best, paths = c(x, lengths)
    assert [len(p) for p in paths] == lengths.tolist() and all(v in (0, 1) for v in paths[0][:100])
    feats = c.model(x, lengths)
    lens = ops.Lengths(lengths, dev, T)
    emis = c.crf.emissions(feats)
    tag_t = torch.zeros(B, T, device=dev)
    for b, p in enumerate(paths):
        tag_t[b, :len(p)] = torch.tensor(p, dtype=torch.float32)
    stats = ops.CrfNllFn.apply(emis.detach(), c.crf.transitions.detach(), tag_t, lens)
    close(stats[1], best, rtol=3e-4, atol=1e-2)          # gold score of the Viterbi path (8192 fp32 terms summed in another order)
    assert bool((stats[0] >= stats[1] * (1 - 3e-4)).all())     # log Z >= score of any single path


# default = fp16-split kernel (lstm_bwd_h3.cu); the TF32 + bf16 kernel of round 1
@pytest.mark.parametrize("entry", ["mts_lstm_rec_bwd_tc", "mts_lstm_rec_bwd_tf32"])
@pytest.mark.parametrize("B,T,n_enc,spread", [(3, 5, 1, 0), (37, 61, 1, 0), (20, 33, 2, 0), (300, 50, 1, 0), (10, 120, 1, 1), (37, 61, 2, 1)])
def test_recurrence_backward_tensor_core_vs_fma(dev, B, T, n_enc, spread, entry):
    """The tcgen05 BPTT kernels (transposed W_hh slice in tensor memory) against the packed-FMA BPTT kernel.  spread = 1:
    upstream gradients whose magnitude differs by ten decades between episodes and by four along an episode, weight
    columns spanning five decades and a zero column -- the fp16-split kernel rescales every dp column at every step."""
    from multimodaltopicsegmentation_b200 import ops

    H = 256
    g = torch.Generator(device=dev).manual_seed(B * 77 + T)
    gx = torch.randn((n_enc, B * T, 8 * H), device=dev, generator=g)
    whh = torch.randn((n_enc, 2, 4 * H, H), device=dev, generator=g) * 0.06
    if spread:
        colscale = 10.0 ** (-5.0 * torch.rand(H, device=dev, generator=g))
        colscale[7] = 0.0
        whh = whh * colscale.view(1, 1, 1, H)
    hg = torch.Generator().manual_seed(B + T)
    lengths = [T] + [int(v) for v in torch.randint(1, T + 1, (B - 1,), generator=hg)]
    lens = ops.Lengths(lengths, dev, T)
    y = torch.empty((B, T, n_enc * 2 * H), device=dev)
    gates = torch.zeros((n_enc, 2, B, T, 5, H), device=dev)
    ops._call("mts_lstm_rec_fwd_tc", gx.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(), lens.order.data_ptr(), n_enc, B, T,
              H, y.data_ptr(), gates.data_ptr(), 0, ops._stream())
    dy = torch.randn((B, T, n_enc * 2 * H), device=dev, generator=g)
    if spread:
        dy = dy * (10.0 ** (-10.0 * torch.rand(B, 1, 1, device=dev, generator=g))) * (10.0 ** (-4.0 * torch.rand(B, T, 1, device=dev, generator=g)))
    whh_t = whh.transpose(2, 3).contiguous()
    d_f = torch.full((n_enc, B * T, 8 * H), float("nan"), device=dev)
    d_t = torch.full((n_enc, B * T, 8 * H), float("nan"), device=dev)
    ops._call("mts_lstm_rec_bwd", dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), whh_t.data_ptr(), lens.dev.data_ptr(),
              lens.order.data_ptr(), n_enc, B, T, H, d_f.data_ptr(), ops._stream())
    ops._call(entry, dy.data_ptr(), gates.data_ptr(), whh.data_ptr(), lens.dev.data_ptr(),
              lens.order.data_ptr(), n_enc, B, T, H, d_t.data_ptr(), ops._stream())
    assert not bool(torch.isnan(d_t).any())       # every row written, zeros at padded steps
    for b, n in enumerate(lengths):
        assert float(d_t[:, b * T + n:(b + 1) * T].abs().max() if n < T else 0.0) == 0.0
    # per episode: gradients of different episodes may differ by many decades, each is held to ITS OWN scale
    worst = 0.0
    for b in range(B):
        f, t = d_f[:, b * T:(b + 1) * T], d_t[:, b * T:(b + 1) * T]
        sc = float(f.abs().max())
        if sc > 0:
            worst = max(worst, float((f - t).abs().max()) / sc)
    print(f"  {entry} B={B} T={T} n_enc={n_enc} spread={spread}: worst per-episode norm-wise error {worst:.2e} (contract 2e-5)")
    assert worst <= 2e-5


# ----------------------------------------------------------------------------------------------------------
# on-device evaluation counts (Pk / WindowDiff / F1) against the host metrics (integer work: bit-exact)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("zero_last", [False, True])
def test_seg_metrics_counts_bit_exact(dev, zero_last):
    from decimal import Decimal

    from multimodaltopicsegmentation_b200 import ops
    from multimodaltopicsegmentation_b200.metrics import compute_Pk, compute_window_diff, f1_boundary

    g = torch.Generator().manual_seed(17 + int(zero_last))
    lengths = [1, 2, 3, 5, 17, 64, 65, 300, 1000, 8192, 40, 40]
    B, T = len(lengths), max(lengths)
    hyp = (torch.rand(B, T, generator=g) < 0.1).to(torch.uint8)
    ref = (torch.rand(B, T, generator=g) < 0.08).float()
    hyp[10] = 0; ref[10] = 0          # no boundaries at all
    hyp[11] = 1; ref[11] = 1          # every sentence a boundary
    for b, n in enumerate(lengths):
        hyp[b, n:] = 255
        ref[b, n:] = -1.0
    lens = ops.Lengths(lengths, dev, T)
    counts = ops.seg_metrics(hyp.to(dev), ref.to(dev), lens, zero_last=zero_last).cpu().tolist()
    for b, n in enumerate(lengths):
        h = hyp[b, :n].numpy().astype(int).copy()
        r = ref[b, :n].numpy().astype(int).copy()
        if zero_last:
            h[-1] = 0; r[-1] = 0
        pk, wd, windows, k, tp, fp, fn, segs = counts[b]
        pk_host = compute_Pk(h, r)
        assert (Decimal(pk) / Decimal(windows) if windows > 0 else Decimal(0)) == pk_host, (b, n)
        try:
            wd_host = compute_window_diff(h, r)
            assert windows > 0 and Decimal(wd) / Decimal(windows) == wd_host, (b, n)
        except AssertionError as e:
            if windows > 0:
                raise e
        denom = 2 * tp + fp + fn
        assert (2.0 * tp / denom if denom else 0.0) == f1_boundary(r, h)


@pytest.mark.parametrize("arch,eb", [("BiLSTM", False), ("BiLSTM", True), ("biLSTMCRF", False)])
def test_test_step_device_metrics_equal_host_walk(dev, arch, eb):
    from multimodaltopicsegmentation_b200 import AudioPortionDataset, TextSegmenter, to_device

    g = torch.Generator().manual_seed(23)
    lines = []
    for n in (30, 12, 21, 7, 55):
        lab = (torch.rand(n, generator=g) < 0.2).long().tolist()
        lab[-1] = 0
        lines.append((torch.randn(n, 20, generator=g), lab, "ep"))
    ds = AudioPortionDataset(lines, {"0": 0, "1": 1}, CRF=arch.lower().endswith("crf"), truncate=False)
    batch = to_device(ds.collater([ds[i] for i in range(len(lines))]), dev)
    torch.manual_seed(1)
    seg = TextSegmenter(2, 20, 256, num_layers=1, architecture=arch, loss_fn="FocalLoss", threshold=0.45,
                        end_boundary=eb, metric="Pk").to(dev)
    seg.device_metrics = True
    on_device = seg.test_step(batch, 0)
    seg.device_metrics = False
    on_host = seg.test_step(batch, 0)
    assert set(on_device) == set(on_host)
    for k in on_host:
        assert on_device[k] == on_host[k], (k, on_device[k], on_host[k])


@pytest.mark.parametrize("metric", ["Pk", "WD", "F1"])
def test_threshold_search_equals_host_sweep(dev, metric):
    """The batched device sweep picks the same threshold and reports the same numbers as the reference's loop over
    np.arange(0.05, 1, 0.05) (lightning_model.py:436-553) evaluated with the host metrics."""
    from multimodaltopicsegmentation_b200 import TextSegmenter
    from multimodaltopicsegmentation_b200.metrics import compute_Pk, compute_window_diff, f1_boundary

    g = torch.Generator().manual_seed(5)
    lengths = [40, 17, 29, 64, 8]
    B, T = len(lengths), max(lengths)
    scores = torch.randn(B, T, 1, generator=g) * 2.0
    target = (torch.rand(B, T, generator=g) < 0.15).float()
    for b, n in enumerate(lengths):
        target[b, n - 1] = 0
        target[b, n:] = -1
    seg = TextSegmenter(2, 20, 256, num_layers=1, architecture="BiLSTM", loss_fn="FocalLoss", metric=metric).to(dev)
    th, res = seg.threshold_search(scores.to(dev), target.to(dev), lengths)
    # host restatement of the reference loop
    p = torch.sigmoid(scores.to(dev)[:, :, 0]).cpu().numpy()  # the same probabilities the device sweep thresholds
    minimise = metric != "F1"
    best, best_th, best_res = (1 if minimise else -1), None, None
    for t in np.arange(0.05, 1, 0.05):
        pk = wd = f1 = 0.0
        for b, n in enumerate(lengths):
            tag, tgt = p[b, :n] > t, target[b, :n].numpy()
            pk += float(compute_Pk(np.array(tag), tgt))
            f1 += f1_boundary(tgt.astype(int), np.array(tag).astype(int))
            try:
                wd += float(compute_window_diff(np.array(tag), tgt))
            except AssertionError:
                wd += float(compute_Pk(np.array(tag), tgt))
        r = {"Pk_loss": pk / B, "F1_loss": f1 / B, "WD_loss": wd / B}
        val = r.pop({"F1": "F1_loss", "WD": "WD_loss"}.get(metric, "Pk_loss"))
        r["valid_loss"] = val
        if (val < best) if minimise else (val > best):
            best, best_th, best_res = val, t, r
    assert th == best_th
    for k in best_res:
        assert res[k] == best_res[k], (k, res[k], best_res[k])


@pytest.mark.parametrize("arch", ["BiLSTM", "biLSTMCRF", "Transformer", "BiLSTMLateFusion"])
def test_predict_batches_equals_predict_step(dev, arch):
    """The pipelined prediction loop yields, batch by batch, what predict_step returns."""
    from multimodaltopicsegmentation_b200 import DevicePrefetcher, TextSegmenter

    torch.manual_seed(4)
    g = torch.Generator().manual_seed(44)
    late = arch == "BiLSTMLateFusion"
    kw = dict(nheads=4, attention_window=4) if arch == "Transformer" else {}
    seg = TextSegmenter(2, [12, 20] if late else 32, 256 if "LSTM" in arch or "CRF" in arch else 16, num_layers=2,
                        architecture=arch, loss_fn="FocalLoss", threshold=0.5, **kw).to(dev).eval()
    batches = []
    for i in range(5):
        B, T = 3 + i, 24
        lengths = torch.randint(1, T + 1, (B,), generator=g)
        lengths[0] = T
        b = {"src_tokens": torch.randn(B, T, 12 if late else 32, generator=g).pin_memory(), "src_lengths": lengths}
        if late:
            b["src_tokens2"] = torch.randn(B, T, 20, generator=g).pin_memory()
        batches.append(b)
    one_by_one = [seg.predict_step(b, i) for i, b in enumerate(DevicePrefetcher(batches, dev))]
    piped = list(seg.predict_batches(DevicePrefetcher(batches, dev)))
    assert len(piped) == len(one_by_one) == 5
    for a, b in zip(piped, one_by_one):
        assert a == b
        assert all(type(x) is type(y) for ra, rb in zip(a, b) for x, y in zip(ra, rb))


@pytest.mark.parametrize("S,lens", [(1, [1, 1]), (7, [7, 1]), (33, [33, 20, 2])])
def test_transformer_odd_sequence_lengths_vs_numpy(dev, S, lens, xf_layout):
    """Sequence lengths HF itself cannot run (S must be a multiple of every layer's window there):
    the banded encoder against the numpy restatement of the same arithmetic, valid positions."""
    from multimodaltopicsegmentation_b200.transformer import Transformer_segmenter
    from oracle import ref_numpy as rn

    torch.manual_seed(S)
    g = torch.Generator().manual_seed(100 + S)
    tr = Transformer_segmenter(2, 32, 16, num_layers=2, nheads=4, loss_fn="FocalLoss", window_size=4).to(dev).eval()
    x = torch.randn(len(lens), S, 32, generator=g)
    tl = torch.tensor(lens)
    with torch.no_grad():
        hid = tr.model(x.to(dev), tl)
    params = {k: v.detach().cpu().numpy() for k, v in tr.state_dict().items()}
    ref = rn.longformer_encoder(x.numpy(), tl.numpy(), params, 4, rn.pyramid_windows(2, 4))
    for b, n in enumerate(lens):
        close(hid[b, :n], ref[b, :n], rtol=1e-4, atol=2e-5)


# ----------------------------------------------------------------------------------------------------------
# early-fusion input projection with the embeddings read in place (mts_gemm_tf32x3_srcs)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,D1,D2,N", [(7, 50, 384, 512, 2048), (3, 41, 64, 40, 2048), (5, 33, 104, 0, 512), (2, 300, 32, 8, 128),
                                          (4, 20, 12, 0, 64), (3, 9, 5, 7, 64)])
def test_input_projection_reads_sources_in_place(dev, B, T, D1, D2, N):
    """[x1 | x2] W^T + b through two TMA maps split along K == the same product over the concatenated operand pair
    (bit for bit: same tiles, same operand values), and both within the GEMM's contract of a float64 product."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator(device=dev).manual_seed(D1 * 7 + D2)
    x1 = torch.randn((B, T, D1), device=dev, generator=g)
    x2 = torch.randn((B, T, D2), device=dev, generator=g) if D2 else None
    w = torch.randn((N, D1 + D2), device=dev, generator=g) * 0.05
    bias = torch.randn((N,), device=dev, generator=g)
    w_hi, w_lo = ops.split_tf32(w, side=ops.B_SIDE)
    outs = []
    for direct in (True, False):
        ops.PROJ_DIRECT = direct
        try:
            n0 = ops.launch_count()
            out = torch.full((B * T, N), float("nan"), device=dev)
            ops.input_projection(x1, x2, B, T, w_hi, w_lo, bias, out, N)
            outs.append(out)
        finally:
            ops.PROJ_DIRECT = True
    assert torch.equal(outs[0], outs[1])
    x = x1 if x2 is None else torch.cat([x1, x2], dim=2)
    ref = x.reshape(B * T, -1).double() @ w.double().T + bias.double()
    err = float((outs[0].double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 2e-5, err
    # a time-cropped view cannot be read in place: the packing path serves it, same result on the kept rows
    xl = torch.randn((B, T + 3, D1), device=dev, generator=g)
    out_c = torch.empty((B * T, N), device=dev)
    ops.input_projection(xl, None if x2 is None else torch.cat([x2, x2[:, :3]], dim=1), B, T, w_hi, w_lo, bias, out_c, N)
    xc = xl[:, :T] if x2 is None else torch.cat([xl[:, :T], x2], dim=2)
    ref_c = xc.reshape(B * T, -1).double() @ w.double().T + bias.double()
    assert float((out_c.double() - ref_c).abs().max()) / float(ref_c.abs().max()) < 2e-5


# ----------------------------------------------------------------------------------------------------------
# fp16-split dense layers (mts_gemm_f16x3) and the LayerNorms that produce their operand
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(300, 256, 64), (1000, 2688, 896), (513, 256, 896), (4096, 2048, 512), (700, 896, 200)])
def test_gemm_f16x3(dev, M, N, K):
    """(A1 B1^T + A2 B1^T + A1 B2^T) rs cs + bias against a float64 product: rows of A and of B spanning six decades (every row
    carries its own power-of-two scale), K padded to 64 with zeros; the GELU + operand-pair epilogue against torch."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = torch.randn((M, K), device=dev, generator=g) * (10.0 ** (torch.rand((M, 1), device=dev, generator=g) * 6 - 3))
    b = torch.randn((N, K), device=dev, generator=g) * (10.0 ** (torch.rand((N, 1), device=dev, generator=g) * 6 - 4))
    a[3] = 0.0
    bias = torch.randn((N,), device=dev, generator=g)
    ap, asc = ops.f16_pieces(a)
    bp, bsc = ops.f16_pieces(b)
    # the pieces reproduce the scaled operand to ~2^-22 of the row maximum
    rec = (ap[:, 0, :K].float() + ap[:, 1, :K].float()) * asc[:, None]
    assert float(((rec - a).abs() / a.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)).max()) < 2.0 ** -21
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm_f16x3(ap, asc, bp, bsc, bias, out, M, N, epilogue=1)
    ref = a.double() @ b.double().T + bias.double()
    # error relative to |a_m| . |b_n| (what an fp32 product of these rows can promise), worst over the matrix
    bound = (a.double().abs() @ b.double().abs().T) + bias.double().abs()
    err = float(((out.double() - ref).abs() / bound.clamp_min(1e-30)).max())
    print(f"  gemm_f16x3 {M} x {N} x {K}: worst |err| / (|a| . |b|) = {err:.2e}")
    assert err < 1e-5, err
    if N % 32 == 0:
        z = torch.full((M, N), float("nan"), device=dev)
        z_lo = torch.full((M, N), float("nan"), device=dev)
        ops.gemm_f16x3(ap, asc, bp, bsc, bias, z, M, N, epilogue=2, out_lo=z_lo)
        zr = torch.nn.functional.gelu(ref.float())
        close(z, zr, rtol=1e-4, atol=2e-5 * float(zr.abs().max()))
        check_operand_pair(z, z_lo, z, side=0)


@pytest.mark.parametrize("M,d", [(300, 896), (64, 100), (1000, 1024)])
def test_layernorm_writes_fp16_split_operand(dev, M, d):
    """mts_add_ln_fwd_f16: y bit-equal to mts_add_ln_fwd's, pieces + row scale reproduce y to ~2^-22 of the row maximum, zero
    padding up to K64."""
    from multimodaltopicsegmentation_b200 import transformer as tr

    g = torch.Generator(device=dev).manual_seed(M + d)
    a = torch.randn((M, d), device=dev, generator=g) * 3.0
    r = torch.randn((M, d), device=dev, generator=g)
    gamma = torch.randn((d,), device=dev, generator=g)
    beta = torch.randn((d,), device=dev, generator=g) * 0.1
    gamma[::7] *= 1e-3
    y_ref = tr._ln(1, a, r, None, gamma, beta, M, 1, d, False, False)[0]
    y, pieces, scale = tr._ln_f16(1, a, r, None, gamma, beta, M, 1, d)
    assert torch.equal(y, y_ref)
    rec = (pieces[:, 0, :d].float() + pieces[:, 1, :d].float()) * scale[:, None]
    rel = (rec - y).abs() / y.abs().amax(dim=1, keepdim=True)
    assert float(rel.max()) < 2.0 ** -21, float(rel.max())
    assert float(pieces[:, :, d:].abs().max() if pieces.shape[2] > d else 0.0) == 0.0
    lg = torch.log2(scale)
    assert torch.equal(lg, lg.round())                      # exact powers of two
    top = y.abs().amax(dim=1) / scale
    assert float(top.min()) >= 2.0 ** 13 and float(top.max()) < 2.0 ** 14


def test_encoder_fp16_split_path_equals_tf32_path(dev):
    """The inference encoder with the LayerNorm-fed dense layers on fp16-split operands against the same encoder with every
    product on mts_gemm_tf32x3 (both fp32-grade: they must agree far inside the 1e-4 contract), ragged batch, 3 layers."""
    from multimodaltopicsegmentation_b200 import transformer as tr

    torch.manual_seed(5)
    g = torch.Generator().manual_seed(55)
    d, F, nh, w = 256, 256, 8, 8
    seg = tr.Transformer_segmenter(2, d, F, num_layers=3, nheads=nh, loss_fn="FocalLoss", window_size=w).to(dev).eval()
    x = torch.randn(6, 96, d, generator=g).to(dev)
    lengths = torch.tensor([96, 50, 77, 96, 13, 64])
    outs = []
    for flag in (True, False):
        tr.F16X3 = flag
        try:
            n0 = tr.ops.launch_count()
            with torch.no_grad():
                outs.append(seg.model(x, lengths))
        finally:
            tr.F16X3 = True
    err = float((outs[0] - outs[1]).abs().max()) / float(outs[1].abs().max())
    print(f"  encoder fp16-split vs tf32 path: norm-wise difference {err:.2e}")
    assert err < 2e-5
    for b, n in enumerate(lengths.tolist()):
        assert float(outs[0][b, n:].abs().max() if n < 96 else 0.0) == 0.0


@pytest.mark.parametrize("B,T,Tmax,D1,D2,N", [(7, 50, 50, 384, 512, 2048), (3, 100, 117, 64, 40, 2048), (5, 60, 60, 100, 0, 512)])
def test_input_projection_fp16_split(dev, B, T, Tmax, D1, D2, N):
    """mts_pack_rows_f16 + mts_gemm_f16x3 as the layer-0 projection: [x1 | x2] W^T + b from time-cropped sources whose rows span
    five decades, against float64; the packed pieces reproduce the cropped, concatenated rows."""
    from multimodaltopicsegmentation_b200 import ops

    g = torch.Generator(device=dev).manual_seed(D1 * 3 + D2)
    x1 = torch.randn((B, Tmax, D1), device=dev, generator=g) * (10.0 ** (torch.rand((B, Tmax, 1), device=dev, generator=g) * 5 - 3))
    x2 = (torch.randn((B, Tmax, D2), device=dev, generator=g) * 3.0) if D2 else None
    w = torch.randn((N, D1 + D2), device=dev, generator=g) * 0.05
    bias = torch.randn((N,), device=dev, generator=g)
    w_hi, w_lo = ops.split_tf32(w, side=ops.B_SIDE)
    wp = ops.f16_pieces(w)
    out = torch.full((B * T, N), float("nan"), device=dev)
    ops.PROFILE, prof = {}, None
    try:
        ops.input_projection(x1, x2, B, T, w_hi, w_lo, bias, out, N, w_pieces=lambda: wp)
        prof = set(ops.PROFILE)
    finally:
        ops.PROFILE = None
    assert "mts_gemm_f16x3" in prof and "mts_pack_rows_f16" in prof       # the fp16-split path really ran
    x = x1[:, :T] if x2 is None else torch.cat([x1[:, :T], x2[:, :T]], dim=2)
    xr = x.reshape(B * T, -1).double()
    ref = xr @ w.double().T + bias.double()
    bound = xr.abs() @ w.double().abs().T + bias.double().abs()
    err = float(((out.double() - ref).abs() / bound).max())
    print(f"  fp16-split input projection {B * T} x {N} x {D1 + D2}: worst |err| / (|x| . |w|) = {err:.2e}")
    assert err < 1e-5, err
