"""Backward pass of the windowed-attention encoder (transformer.EncoderFn).  The reference obtains these
gradients from autograd through HF LongformerModel; here every step is a kernel of libmts_b200.so:

    dX of a dense layer   tcgen05 3xTF32 GEMM of the (hi, lo) halves of dY against the transposed weight halves
    dW, db                dY^T X contracts over tokens: both operands transposed to K-major (hi, lo) halves
                          (mts_transpose_split), then the same tcgen05 GEMM; column sums for the biases
    LayerNorm, GELU, banded attention, embeddings: csrc/xfmr_bwd.cu

Residual connections are folded in by letting the dX GEMM accumulate into the buffer that already holds the
gradient arriving through the skip path.
"""
from __future__ import annotations

import torch

from . import _lib, ops

_ptr, _call, _stream, _pad32 = ops._ptr, ops._call, ops._stream, ops._pad32


def _ln_bwd(dy, pre, stats, gamma, M, d):
    dev = dy.device
    kp = _pad32(d)
    dx = torch.empty((M, d), device=dev, dtype=torch.float32)
    lo = torch.empty((M, kp), device=dev, dtype=torch.float32)
    hi = None if kp == d else torch.empty((M, kp), device=dev, dtype=torch.float32)  # dx itself is the `hi` operand
    dgb = torch.empty((2, d), device=dev, dtype=torch.float32)
    ws = torch.empty(_lib.load().mts_ln_bwd_ws_bytes(M, d) // 4, device=dev, dtype=torch.float32)
    _call("mts_ln_bwd", _ptr(dy), _ptr(pre), _ptr(stats), _ptr(gamma), M, d, _ptr(dx), _ptr(hi), _ptr(lo), kp,
          _ptr(dgb[0]), _ptr(dgb[1]), _ptr(ws), _stream())
    return dx, (dx if hi is None else hi, lo), dgb[0], dgb[1]


def _dense_param_grads(dy, x, M, n_out, n_in):
    """dW [n_out, n_in] = dY^T X over M tokens, db = column sums of dY."""
    dev = dy.device
    dw = torch.empty((n_out, n_in), device=dev, dtype=torch.float32)
    if ops.GEMM_IMPL != "simt":  # contraction over tokens: transpose both operands to K-major halves, tcgen05 GEMM
        dyT = ops.transpose_split(_ptr(dy), 0, dy.stride(0), M, n_out, M, dev)
        ops.weight_grad(dyT, _ptr(x), 0, x.stride(0), M, n_in, M, dw, n_in, dev)
    else:
        splits = max(1, min(16, M // 2048))
        ops.gemm_f32(_ptr(dy), dy.stride(0), _ptr(x), x.stride(0), None, _ptr(dw), n_in, n_out, n_in, M, layout=3,
                     splits=splits)
    db = torch.empty(n_out, device=dev, dtype=torch.float32)
    ops.colsum(_ptr(dy), dy.stride(0), M, n_out, db)
    return dw, db


def encoder_backward(ctx, dout):
    saved, lens, packed, nheads, reaches = ctx.saved, ctx.lens, ctx.packed, ctx.nheads, ctx.reaches
    B, S, d = ctx.shape
    ragged = saved["ragged"]
    M = lens.N if ragged else B * S
    offs = _ptr(lens.offs) if ragged else 0
    hd = d // nheads
    dev = dout.device
    m = packed.params
    n_layers = len(m.encoder.layer)
    per = packed.PER_LAYER
    grads = [None] * (4 + per * n_layers)
    masks, p_hidden = saved.get("masks", []), saved.get("p_hidden", 0.0)

    def through_dropout(dpre, site):
        """gradient of the dense output that was dropped before the residual add: (d, (hi, lo)) of d * keep / (1 - p);
        the skip path keeps the unmasked dpre."""
        g = (dpre * masks[site]).mul_(1.0 / (1.0 - p_hidden))
        hl = ops.split_tf32(g)
        return g, ((g if _pad32(d) == d else hl[0]), hl[1])

    if ragged:  # gradient rows of the valid sentences only, in the forward pass's ragged order
        dh = torch.empty((M, d), device=dev, dtype=torch.float32)
        _call("mts_ragged_copy", _ptr(dout), _ptr(dh), _ptr(lens.dev), _ptr(lens.offs), B, S, d, 0, 0.0, _stream())
    else:
        dh = dout.reshape(M, d)
    for l in range(n_layers - 1, -1, -1):
        lyr = m.encoder.layer[l]
        sv = saved["layers"][l]
        ent = packed.transposed(l)
        F = lyr.intermediate.dense.out_features
        kf = _pad32(F)
        base = 4 + per * l
        # ---- output LayerNorm:  h_out = LN(u + y) -------------------------------------------------------
        dpre2, dpre2_hl, dg2, db2 = _ln_bwd(dh, sv["pre2"], sv["st2"], lyr.output.LayerNorm.weight.detach(), M, d)
        grads[base + 14], grads[base + 15] = dg2, db2
        # ---- u = z W2^T + b2 -----------------------------------------------------------------------------
        du, du_hl = through_dropout(dpre2, 2 + 2 * l) if p_hidden > 0 else (dpre2, dpre2_hl)
        grads[base + 12], grads[base + 13] = _dense_param_grads(du, sv["z"], M, d, F)
        dz = torch.empty((M, F), device=dev, dtype=torch.float32)
        ops.gemm_tf32x3(du_hl[0], du_hl[1], ent["w2_t"][0], ent["w2_t"][1], None, dz, M, F)
        # ---- z = GELU(zp) ---------------------------------------------------------------------------------
        dzp = torch.empty((M, F), device=dev, dtype=torch.float32)
        dzp_hl = torch.empty((2, M, kf), device=dev, dtype=torch.float32)
        _call("mts_gelu_bwd", _ptr(dz), _ptr(sv["zp"]), M, F, kf, _ptr(dzp), _ptr(dzp_hl[0]), _ptr(dzp_hl[1]), _stream())
        # ---- zp = y W1^T + b1;  dy = dzp W1 + dpre2 (skip path) ----------------------------------------------
        grads[base + 10], grads[base + 11] = _dense_param_grads(dzp, sv["y"], M, F, d)
        ops.gemm_tf32x3(dzp_hl[0], dzp_hl[1], ent["w1_t"][0], ent["w1_t"][1], None, dpre2, M, d, accumulate=True)
        dy = dpre2
        # ---- attention-output LayerNorm:  y = LN(t + h_in) -----------------------------------------------------
        dpre1, dpre1_hl, dg1, db1 = _ln_bwd(dy, sv["pre1"], sv["st1"], lyr.attention.output.LayerNorm.weight.detach(), M, d)
        grads[base + 8], grads[base + 9] = dg1, db1
        # ---- t = a Wo^T + bo ---------------------------------------------------------------------------------
        dt, dt_hl = through_dropout(dpre1, 1 + 2 * l) if p_hidden > 0 else (dpre1, dpre1_hl)
        grads[base + 6], grads[base + 7] = _dense_param_grads(dt, sv["a"], M, d, d)
        da = torch.empty((M, d), device=dev, dtype=torch.float32)
        ops.gemm_tf32x3(dt_hl[0], dt_hl[1], ent["wo_t"][0], ent["wo_t"][1], None, da, M, d)
        # ---- banded attention ----------------------------------------------------------------------------------
        dqkv = torch.empty((M, 3 * d), device=dev, dtype=torch.float32)
        delta = torch.empty((B, nheads, S), device=dev, dtype=torch.float32)
        _call("mts_band_attn_bwd_dropout", _ptr(sv["qkv"]), 3 * d, _ptr(sv["a"]), _ptr(da), _ptr(sv["lse"]), _ptr(lens.dev), offs,
              B, S, nheads, hd, reaches[l], _ptr(dqkv), _ptr(delta), float(saved.get("p_attn", 0.0)), int(sv.get("attn_seed", 0)),
              _stream())
        # ---- qkv = h_in Wqkv^T + b;  dh_in = dqkv Wqkv + dpre1 (skip path) ---------------------------------------
        dwqkv, dbqkv = _dense_param_grads(dqkv, sv["h_in"], M, 3 * d, d)
        for j in range(3):
            grads[base + 2 * j], grads[base + 2 * j + 1] = dwqkv[j * d:(j + 1) * d], dbqkv[j * d:(j + 1) * d]
        dqkv_hl = ops.split_tf32(dqkv)
        ops.gemm_tf32x3(dqkv_hl[0], dqkv_hl[1], ent["wqkv_t"][0], ent["wqkv_t"][1], None, dpre1, M, d, accumulate=True)
        dh = dpre1
    # ---- embeddings: h0 = LN(x + P[2 + t] + E_type[0]) ---------------------------------------------------------
    emb = m.embeddings
    pre0, st0 = saved["emb"]
    if p_hidden > 0:
        dh = (dh * masks[0]).mul_(1.0 / (1.0 - p_hidden))
    dpre0, _, dg0, db0 = _ln_bwd(dh, pre0, st0, emb.LayerNorm.weight.detach(), M, d)
    dpos = torch.zeros_like(emb.position_embeddings.weight)
    _call("mts_embed_bwd", _ptr(dpre0), B, S, d, dpos.data_ptr() + 4 * 2 * d, _ptr(lens.dev) if ragged else 0, offs,
          _stream())
    dtyp = torch.zeros_like(emb.token_type_embeddings.weight)
    ops.colsum(dpos.data_ptr() + 4 * 2 * d, d, S, d, dtyp[0])
    grads[0], grads[1], grads[2], grads[3] = dtyp, dpos, dg0, db0
    return grads
