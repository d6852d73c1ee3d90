"""Training-path MLP on the tcgen05 kernels.

Forward (default, ``train_fwd = "fused"``): ``vanilla_fused`` / ``autodecoder_fused`` run cast_rays + pos_enc + the whole MLP
chain of a level (NeRFMLP.forward, models/vanilla_nerf/model.py:95-120; NeRFMLP_AE.forward, model_autodecoder.py:171-239) in
ONE launch of the fused render kernel (``aon_forward_train``): activations move from layer to layer through shared memory /
TMEM and every layer output is written to HBM once, as the packed fp16 hi+lo planes (+ ReLU bit planes) the backward reads.
``vanilla_mlp`` / ``autodecoder_mlp`` (``train_fwd = "layers"``) run every nn.Linear as an ``aon_gemm_tc`` launch instead.

Backward (both): written by hand, no autograd inside -- the dgrad chain with the ReLU mask in the GEMM epilogue, the weight
gradients as split-K tcgen05 GEMMs that read the saved activation / gradient planes as MN-major operands, the bias gradients
as column sums produced by the dgrad GEMMs (3 MMAs per K step on hi+lo planes, fp32 accumulation: fp32-grade results).

Scaling (exact powers of two, so nothing is lost): activations x8, weights x64 (the fused render kernel's choice, which
keeps the lo planes out of the fp16 subnormal range) and gradients x SG with SG = 2^floor(log2(rays)) -- the loss
gradient is O(1 / rays) (mean over 3 * rays residuals), so scaled gradients stay O(1).
"""
from __future__ import annotations

import math
import os
from typing import List

import torch

from . import lib as L

SA, SW = 8.0, 64.0

# Operand planes: True = fp16 hi + lo (3 MMAs per K step; fp32-grade gradients, the default "tc" path); False = the hi plane
# only (one MMA per K step, half the HBM traffic of every GEMM; ~1e-3 relative gradient noise -- the fast training mode
# "tc16", to "tc" what the one-pass f16 eval mode is to f16x3).  Set per call by vanilla_mlp / autodecoder_mlp.
_X3 = True


def _PK(tiles, feat, dev):
    return L.PK(tiles, feat, dev, _X3)


def _pack_rows(*a, **k):
    return L.pack_rows(*a, x3=_X3, **k)


def _pack_linear(*a, **k):
    return L.pack_linear(*a, x3=_X3, **k)


def _gemm_nt(*a, **k):
    return L.gemm_nt(*a, x3=_X3, **k)


def _gemm_tn(*a, **k):
    return L.gemm_tn(*a, x3=_X3, **k)


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _pad_bias(b: torch.Tensor, n: int) -> torch.Tensor:
    if b.numel() == n:
        return b.detach().contiguous()
    out = torch.zeros(n, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b.detach()
    return out


# row tiles per sub-batch of the backward chain (0 = the whole level at once, the default).  MEASURED AND NOT ADOPTED: with 296
# tiles a 256-wide delta plane of a sub-batch is 39 MB (hi + lo) and stays in L2 between the GEMM that writes it and the two that
# read it, which halves the HBM traffic of the backward -- but 11 sub-batches turn 60 large launches into 660 small ones, and
# each costs ~19 us of fill / drain / split-K reduce: vanilla step 7.75 ms -> 14.0 ms (148: 18.3, 444: 12.5, 592: 11.9 ms).
BWD_TILES = int(os.environ.get("AON_BWD_TILES", "0"))


def _splits(m_tiles: int, a_tiles: int) -> int:
    """split-K factor of a wgrad GEMM: two CTAs per SM on a whole level, one per SM on a sub-batch (every split then still owns
    >= 4 row tiles, and the fp32 partial tiles stay small next to the operand planes)"""
    return max(1, (296 if m_tiles > 2 * SPLIT_TILES else 148) // a_tiles)


SPLIT_TILES = 296


def _wgrad(dY: L.PK, out_valid: int, X: L.PK, x_off: int, n_cols: int, cols_valid: int, dst: torch.Tensor, col_off: int,
           inv: float, accumulate: bool = False) -> None:
    """dst[0:out_valid, col_off : col_off + cols_valid] (+)= inv * dY^T X[:, x_off : x_off + n_cols]."""
    a_tiles = dY.feat // 128
    part = _gemm_tn(dY, 0, a_tiles, X, x_off, n_cols, _splits(dY.m_tiles, a_tiles))
    L.wgrad_reduce(part, inv, dst, col_off, out_valid, cols_valid, accumulate=accumulate)


def _wgrad_head(X: L.PK, in_valid: int, G: L.PK, out_valid: int, dst: torch.Tensor, inv: float, accumulate: bool = False) -> None:
    """Heads (1 / 3 output features, padded to 16): dst[o, i] = inv * sum_m G[m, o] X[m, i] as the transposed product
    (rows = the 128-feature tiles of X, columns = the 16 padded outputs)."""
    a_tiles = X.feat // 128
    part = _gemm_tn(X, 0, a_tiles, G, 0, 16, _splits(X.m_tiles, a_tiles))
    L.wgrad_reduce(part, inv, dst, 0, in_valid, out_valid, transpose=True, accumulate=accumulate)


class _VanillaMLPFn(torch.autograd.Function):
    """(enc [M,63], view_enc [R,27], S, w0, b0, ..., w11, b11) -> raw [M,4] = (rgb, sigma); parameters in
    NeRFMLP.linears() order: pts_linears.0-7, views_linear.0, bottleneck_layer, density_layer, rgb_layer."""

    @staticmethod
    def forward(ctx, enc, view_enc, S, *params):
        ctx.x3 = _X3
        M, dev = enc.shape[0], enc.device
        tiles = (M + 127) // 128
        W = [p.detach().contiguous() for p in params[0::2]]
        B = [p.detach() for p in params[1::2]]
        E = _pack_rows(enc.detach(), M, tiles, 64, SA)
        V = _pack_rows(view_enc.detach(), M, tiles, 32, SA, row_div=S)
        inv = 1.0 / (SA * SW)
        h: List[L.PK] = []
        x, kx = E, 64
        for i in range(8):
            Wp = _pack_linear(W[i], False, 256, _pad(W[i].shape[1], 16) if i != 5 else 320, SW)
            segs = [(x, 0, kx, Wp, 0, 0)]
            if i == 5:
                segs.append((E, 0, 64, Wp, 256, 0))
            out = _PK(tiles, 256, dev)
            _gemm_nt(segs, 256, tiles, dev, bias=B[i].contiguous(), relu=True, inv_scale=inv, out=out, out_scale=SA)
            h.append(out)
            x, kx = out, 256
        raw = torch.empty(tiles * 128, 4, dtype=torch.float32, device=dev)
        Wd = _pack_linear(W[10], False, 16, 256, SW)
        _gemm_nt([(h[7], 0, 256, Wd, 0, 0)], 16, tiles, dev, bias=_pad_bias(B[10], 16), inv_scale=inv, out_f32=raw[:, 3:], n_valid=1)
        Wb = _pack_linear(W[9], False, 256, 256, SW)
        bott = _PK(tiles, 256, dev)
        _gemm_nt([(h[7], 0, 256, Wb, 0, 0)], 256, tiles, dev, bias=B[9].contiguous(), inv_scale=inv, out=bott, out_scale=SA)
        Wv = _pack_linear(W[8], False, 128, 288, SW)
        hv = _PK(tiles, 128, dev)
        _gemm_nt([(bott, 0, 256, Wv, 0, 0), (V, 0, 32, Wv, 256, 0)], 128, tiles, dev, bias=B[8].contiguous(), relu=True,
                  inv_scale=inv, out=hv, out_scale=SA)
        Wr = _pack_linear(W[11], False, 16, 128, SW)
        _gemm_nt([(hv, 0, 128, Wr, 0, 0)], 16, tiles, dev, bias=_pad_bias(B[11], 16), inv_scale=inv, out_f32=raw, n_valid=3)
        ctx.pk = (E, V, h, bott, hv)
        ctx.W = W
        ctx.dims = (M, tiles, S)
        return raw[:M]

    @staticmethod
    def backward(ctx, g_raw):
        return (None, None, None) + _vanilla_backward(ctx, g_raw, None)


def _pack_grad(src: torch.Tensor, M: int, tiles: int, c_pad: int, scale: float, tiled) -> L.PK:
    """gradient rows [M, C] (ray-major) -> packed plane: plain row order, or the tile order of the fused forward (tiled = (R, S))"""
    if tiled is None:
        return _pack_rows(src, M, tiles, c_pad, scale)
    return L.pack_rows_tiled(src, tiled[0], tiled[1], c_pad, scale, x3=_X3)


def _vanilla_backward(ctx, g_raw, tiled):
    """dgrad chain + wgrads + bias gradients of NeRFMLP from the saved planes; returns the 24 parameter gradients."""
    global _X3
    _X3 = ctx.x3
    if ctx.pk is None:
        raise L.AonError("the tcgen05 training path frees its saved operand planes in backward(): a second backward through the "
                         "same graph (retain_graph) is not supported -- re-run the forward, or set train_gemm = 'torch'")
    E, V, h, bott, hv = ctx.pk
    W = ctx.W
    M, tiles, S = ctx.dims
    dev = g_raw.device
    R = max(1, M // S)
    SG = float(2 ** int(math.floor(math.log2(R))))
    g_raw = g_raw.contiguous()
    Gr = _pack_grad(g_raw, M, tiles, 16, SG, tiled)                 # columns 0-2 (+ sigma in column 3, unused here: zero weight rows)
    Gs = _pack_grad(g_raw[:, 3:], M, tiles, 16, SG, tiled)           # sigma gradient alone (C = 1)
    inv_w = 1.0 / (SG * SA)
    gW = [torch.empty_like(w) for w in W]           # every entry is written by a wgrad_reduce launch
    gB = [None] * 12
    gB[11] = g_raw[:, :3].sum(0)
    gB[10] = g_raw[:, 3:].sum(0)
    # transposed weight packs: once per backward pass
    WrT = _pack_linear(W[11], True, 128, 16, SW)
    WvT = _pack_linear(W[8], True, 288, 128, SW)
    WbT = _pack_linear(W[9], True, 256, 256, SW)
    WdT = _pack_linear(W[10], True, 256, 16, SW)
    WT = {i: _pack_linear(W[i], True, 320 if i == 5 else 256, 256, SW) for i in range(1, 8)}

    # The chain can run over sub-batches of row tiles (BWD_TILES, see there: measured, off by default); wgrad partial sums then
    # accumulate over the sub-batches in a fixed order (aon_wgrad_reduce flags & 2): deterministic.
    step = BWD_TILES if BWD_TILES > 0 else tiles
    for t0 in range(0, tiles, step):
        t1 = min(tiles, t0 + step)
        nt, acc = t1 - t0, t0 > 0
        Es, Vs, botts, hvs, Grs, Gss = (x.tiles(t0, t1) for x in (E, V, bott, hv, Gr, Gs))
        hs = [x.tiles(t0, t1) for x in h]

        def bias(i, cs):
            v = L.colsum_finish(cs, 1.0 / SG)
            gB[i] = gB[i] + v if acc else v

        # rgb_layer: the packed gradient has 4 live columns (r, g, b, sigma); W_r^T padded with zero rows ignores sigma
        _wgrad_head(hvs, 128, Grs, 3, gW[11], inv_w, acc)
        d_hv = _PK(nt, 128, dev)
        cs = _gemm_nt([(Grs, 0, 16, WrT, 0, 0)], 128, nt, dev, epi=L.EPI_MASK, mask=(hvs, 0), inv_scale=1.0 / SW, out=d_hv, colsum=True)
        # views_linear.0 : inputs [bottleneck(256), view enc(27)]
        _wgrad(d_hv, 128, botts, 0, 256, 256, gW[8], 0, inv_w, acc)
        _wgrad(d_hv, 128, Vs, 0, 32, 27, gW[8], 256, inv_w, acc)
        bias(8, cs)
        d_bott = _PK(nt, 256, dev)
        cs = _gemm_nt([(d_hv, 0, 128, WvT, 0, 0)], 256, nt, dev, epi=L.EPI_MASK, inv_scale=1.0 / SW, out=d_bott, colsum=True)
        # bottleneck_layer and density_layer both read the last trunk activation h[7]
        _wgrad(d_bott, 256, hs[7], 0, 256, 256, gW[9], 0, inv_w, acc)
        bias(9, cs)
        _wgrad_head(hs[7], 256, Gss, 1, gW[10], inv_w, acc)
        d = _PK(nt, 256, dev)
        cs = _gemm_nt([(d_bott, 0, 256, WbT, 0, 0), (Gss, 0, 16, WdT, 0, 0)], 256, nt, dev, epi=L.EPI_MASK, mask=(hs[7], 0),
                       inv_scale=1.0 / SW, out=d, colsum=True)
        # trunk, last layer first: pts_linears.i maps x_i -> h[i], x_0 = E, x_i = h[i-1] (+ E for i = 5)
        for i in range(7, -1, -1):
            x = Es if i == 0 else hs[i - 1]
            kin = 64 if i == 0 else 256
            _wgrad(d, 256, x, 0, kin, 63 if i == 0 else 256, gW[i], 0, inv_w, acc)
            if i == 5:
                _wgrad(d, 256, Es, 0, 64, 63, gW[5], 256, inv_w, acc)
            bias(i, cs)                                      # column sums of d, produced by the GEMM that wrote d
            if i == 0:
                break
            nd = _PK(nt, 256, dev)
            cs = _gemm_nt([(d, 0, 256, WT[i], 0, 0)], 256, nt, dev, epi=L.EPI_MASK, mask=(hs[i - 1], 0), inv_scale=1.0 / SW, out=nd,
                           colsum=True)
            d = nd
    ctx.pk = None
    out = []
    for w, b in zip(gW, gB):
        out += [w, b]
    return tuple(out)


class _VanillaFusedFn(torch.autograd.Function):
    """(rays_o [R,3], rays_d [R,3], viewdirs [R,3], t_vals [R,S], view_enc [R,27], precision, w0, b0, ...) -> raw [R*S,4]:
    cast_rays + pos_enc + the whole NeRFMLP forward chain of the level in ONE launch of the fused render kernel
    (aon_forward_train, SURVEY.md 8f F1 stage 3): activations move from layer to layer through shared memory / TMEM and each
    layer output is written to HBM once, as the packed planes the backward GEMMs read.  The backward is _vanilla_backward on
    those planes (tile order of the fused kernel)."""

    @staticmethod
    def forward(ctx, o, d, v, t_vals, view_enc, precision, *params):
        ctx.x3 = _X3
        R, S = t_vals.shape
        W = [p.detach().contiguous() for p in params[0::2]]
        B = [p.detach().contiguous() for p in params[1::2]]
        packed = L.pack_weights(L.KIND_VANILLA, precision, W, B)
        acts, enc, raw, _ = L.forward_train(L.KIND_VANILLA, precision, packed, None, o, d, v, t_vals.contiguous(), S)
        V = L.pack_rows_tiled(view_enc.detach().contiguous(), R, S, 32, SA, per_ray=True, x3=_X3)
        ctx.pk = (enc, V, acts[0:8], acts[8], acts[9])
        ctx.W = W
        ctx.dims = (R * S, enc.m_tiles, S)
        ctx.tiled = (R, S)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        return (None,) * 6 + _vanilla_backward(ctx, g_raw, ctx.tiled)


def vanilla_fused(o, d, v, t_vals, view_enc, mlp, x3: bool = True) -> tuple:
    """rays [R,3] x 3, t_vals [R,S], view_enc [R,27] -> (raw_rgb [R,S,3], raw_sigma [R,S,1]) through the fused forward chain."""
    global _X3
    _X3 = bool(x3)
    params = []
    for lin in mlp.linears():
        params += [lin.weight, lin.bias]
    R, S = t_vals.shape
    raw = _VanillaFusedFn.apply(o, d, v, t_vals, view_enc, L.PREC_TC_F16X3 if x3 else L.PREC_TC_F16, *params)
    return raw[:, :3].reshape(R, S, 3), raw[:, 3:].reshape(R, S, 1)


def _fold(b: torch.Tensor, W: torch.Tensor, parts) -> torch.Tensor:
    """bias + sum W[:, c0:c0+len(code)] @ code: the latent columns of a layer contracted once per call (the reference
    broadcasts the codes to every sample and concatenates, model_autodecoder.py:186-198,226-228)."""
    out = b.detach().clone()
    for c0, code in parts:
        out += W[:, c0:c0 + code.numel()] @ code.reshape(-1)
    return out.contiguous()


class _AutoDecoderMLPFn(torch.autograd.Function):
    """(pos [M,3], view_enc [R,27], S, shape [1,128], appearance [1,128], articulation [1,32], w0, b0, ...) -> raw [M,4];
    parameters in NeRFMLP_AE.linears() order: deformations_linear.0-3 (0-3), deformation_layer (4), pts_linears.0-7 (5-12),
    views_linear.0-3 (13-16), bottleneck_layer (17), density_layer (18), rgb_layer (19).  model_autodecoder.py:171-239."""

    @staticmethod
    def forward(ctx, pos, view_enc, S, shape, app, art, *params):
        ctx.x3 = _X3
        M, dev = pos.shape[0], pos.device
        tiles = (M + 127) // 128
        W = [p.detach().contiguous() for p in params[0::2]]
        B = [p.detach() for p in params[1::2]]
        s_, c_, a_ = shape.detach().reshape(-1), app.detach().reshape(-1), art.detach().reshape(-1)
        inv = 1.0 / (SA * SW)
        pos = pos.detach().contiguous()

        def lin(segs, w, k_pad, n, bias, relu):
            Wp = _pack_linear(w, False, n, k_pad, SW)
            out = _PK(tiles, n, dev)
            _gemm_nt([(x, 0, k, Wp, off, 0) for x, k, off in segs], n, tiles, dev, bias=bias, relu=relu, inv_scale=inv, out=out,
                      out_scale=SA)
            return out

        def head(x, k, w, bias, n_valid):
            Wp = _pack_linear(w, False, 16, k, SW)
            o = torch.empty(tiles * 128, n_valid, dtype=torch.float32, device=dev)
            _gemm_nt([(x, 0, k, Wp, 0, 0)], 16, tiles, dev, bias=_pad_bias(bias, 16), inv_scale=inv, out_f32=o, n_valid=n_valid)
            return o

        # articulation warp: deformation MLP on [pos, shape, articulation]
        P = _pack_rows(pos, M, tiles, 16, SA)
        hd = [lin([(P, 16, 0)], W[0], 176, 128, _fold(B[0], W[0], [(3, s_), (131, a_)]), True)]
        for i in (1, 2, 3):
            hd.append(lin([(hd[-1], 128, 0)], W[i], 128, 128, B[i].contiguous(), True))
        delta = head(hd[3], 128, W[4], B[4], 3)
        warped = (delta[:M] + pos).contiguous()                                     # model_autodecoder.py:203
        E = _pack_rows(L.pos_enc(warped, 10), M, tiles, 64, SA)
        V = _pack_rows(view_enc.detach(), M, tiles, 32, SA, row_div=S)
        h = [lin([(E, 64, 0)], W[5], 192, 256, _fold(B[5], W[5], [(63, s_)]), True)]
        for i in range(1, 8):
            if i == 5:
                h.append(lin([(h[-1], 256, 0), (E, 64, 256)], W[10], 448, 256, _fold(B[10], W[10], [(319, s_)]), True))
            else:
                h.append(lin([(h[-1], 256, 0)], W[5 + i], 256, 256, B[5 + i].contiguous(), True))
        raw = torch.empty(tiles * 128, 4, dtype=torch.float32, device=dev)
        Wd = _pack_linear(W[18], False, 16, 256, SW)
        _gemm_nt([(h[7], 0, 256, Wd, 0, 0)], 16, tiles, dev, bias=_pad_bias(B[18], 16), inv_scale=inv, out_f32=raw[:, 3:], n_valid=1)
        bott = lin([(h[7], 256, 0)], W[17], 256, 256, B[17].contiguous(), False)
        hv = [lin([(bott, 256, 0), (V, 32, 256)], W[13], 416, 128, _fold(B[13], W[13], [(283, c_)]), True)]
        for i in (1, 2, 3):
            hv.append(lin([(hv[-1], 128, 0)], W[13 + i], 128, 128, B[13 + i].contiguous(), True))
        Wr = _pack_linear(W[19], False, 16, 128, SW)
        _gemm_nt([(hv[3], 0, 128, Wr, 0, 0)], 16, tiles, dev, bias=_pad_bias(B[19], 16), inv_scale=inv, out_f32=raw, n_valid=3)
        ctx.pk = (P, hd, warped, E, V, h, bott, hv)
        ctx.W, ctx.codes, ctx.dims = W, (s_, c_, a_), (M, tiles, S)
        return raw[:M]

    @staticmethod
    def backward(ctx, g_raw):
        return (None, None, None) + _autodecoder_backward(ctx, g_raw, None)


def _autodecoder_backward(ctx, g_raw, tiled):
    """colour branch, trunk, encoding adjoint, deformation MLP: parameter gradients (40) preceded by the three code gradients.
    tiled = (R, S): the saved planes (and `warped`) are in the tile order of the fused forward."""
    global _X3
    _X3 = ctx.x3
    if ctx.pk is None:
        raise L.AonError("the tcgen05 training path frees its saved operand planes in backward(): a second backward through the "
                         "same graph (retain_graph) is not supported -- re-run the forward, or set train_gemm = 'torch'")
    P, hd, warped, E, V, h, bott, hv = ctx.pk
    W, (s_, c_, a_), (M, tiles, S) = ctx.W, ctx.codes, ctx.dims
    dev = g_raw.device
    SG = float(2 ** int(math.floor(math.log2(max(1, M // S)))))
    inv_w = 1.0 / (SG * SA)
    g_raw = g_raw.contiguous()
    gW = [torch.empty_like(w) for w in W]
    gB = [None] * 20
    g_s, g_c, g_a = torch.zeros_like(s_), torch.zeros_like(c_), torch.zeros_like(a_)

    def dgrad(segs, n, mask, rows_pad, k_pad):
        """segs: [(dY, kext, weight index, first row of W^T)] -> (PK(tiles, n) masked by `mask`, its column sums / SG)."""
        out = _PK(tiles, n, dev)
        gs = [(dY, 0, k, _pack_linear(W[wi], True, rows_pad[j], k_pad[j], SW), 0, r0) for j, (dY, k, wi, r0) in enumerate(segs)]
        cs = _gemm_nt(gs, n, tiles, dev, epi=L.EPI_MASK, mask=None if mask is None else (mask, 0), inv_scale=1.0 / SW, out=out,
                       colsum=True)
        return out, L.colsum_finish(cs, 1.0 / SG)

    def latent(wi, gb, parts):
        """latent columns of layer wi: dW = gb (x) code, d code = W[:, cols]^T gb."""
        for c0, code, acc in parts:
            n = code.numel()
            gW[wi][:, c0:c0 + n] = torch.outer(gb, code)
            acc += W[wi][:, c0:c0 + n].t() @ gb

    # ---- colour branch ----
    Gr = _pack_grad(g_raw, M, tiles, 16, SG, tiled)
    Gs = _pack_grad(g_raw[:, 3:], M, tiles, 16, SG, tiled)
    gB[19], gB[18] = g_raw[:, :3].sum(0), g_raw[:, 3:].sum(0)
    _wgrad_head(hv[3], 128, Gr, 3, gW[19], inv_w)
    d, cs = dgrad([(Gr, 16, 19, 0)], 128, hv[3], [128], [16])
    for i in (3, 2, 1):
        _wgrad(d, 128, hv[i - 1], 0, 128, 128, gW[13 + i], 0, inv_w)
        gB[13 + i] = cs
        d, cs = dgrad([(d, 128, 13 + i, 0)], 128, hv[i - 1], [128], [128])
    _wgrad(d, 128, bott, 0, 256, 256, gW[13], 0, inv_w)
    _wgrad(d, 128, V, 0, 32, 27, gW[13], 256, inv_w)
    gB[13] = cs
    latent(13, gB[13], [(283, c_, g_c)])
    d_bott, cs = dgrad([(d, 128, 13, 0)], 256, None, [416], [128])
    # ---- trunk ----
    _wgrad(d_bott, 256, h[7], 0, 256, 256, gW[17], 0, inv_w)
    gB[17] = cs
    _wgrad_head(h[7], 256, Gs, 1, gW[18], inv_w)
    d, cs = dgrad([(d_bott, 256, 17, 0), (Gs, 16, 18, 0)], 256, h[7], [256, 256], [256, 16])
    d5 = None
    for i in range(7, -1, -1):
        wi = 5 + i
        x = E if i == 0 else h[i - 1]
        _wgrad(d, 256, x, 0, 64 if i == 0 else 256, 63 if i == 0 else 256, gW[wi], 0, inv_w)
        gB[wi] = cs
        if i == 5:
            _wgrad(d, 256, E, 0, 64, 63, gW[wi], 256, inv_w)
            latent(wi, gB[wi], [(319, s_, g_s)])
            d5 = d
        if i == 0:
            latent(wi, gB[wi], [(63, s_, g_s)])
            break
        d, cs = dgrad([(d, 256, wi, 0)], 256, h[i - 1], [448 if i == 5 else 256], [256])
    # ---- encoding -> warped position -> deformation MLP ----
    g_enc = torch.empty(tiles * 128, 63, dtype=torch.float32, device=dev)
    W0T, W5T = _pack_linear(W[5], True, 192, 256, SW), _pack_linear(W[10], True, 448, 256, SW)
    _gemm_nt([(d, 0, 256, W0T, 0, 0), (d5, 0, 256, W5T, 0, 256)], 64, tiles, dev, epi=L.EPI_MASK, inv_scale=1.0 / (SW * SG),
              out_f32=g_enc, n_valid=63)
    # [rows,3]; d warped / d delta = 1.  Fused forward: warped and g_enc are both in tile order (padding rows: zero gradient)
    g_warped = L.pos_enc_backward(warped, g_enc[:warped.shape[0]], 10)
    Gd = _pack_rows(g_warped, g_warped.shape[0], tiles, 16, SG)
    gB[4] = g_warped.sum(0)
    _wgrad_head(hd[3], 128, Gd, 3, gW[4], inv_w)
    d, cs = dgrad([(Gd, 16, 4, 0)], 128, hd[3], [128], [16])
    for i in (3, 2, 1):
        _wgrad(d, 128, hd[i - 1], 0, 128, 128, gW[i], 0, inv_w)
        gB[i] = cs
        d, cs = dgrad([(d, 128, i, 0)], 128, hd[i - 1], [128], [128])
    _wgrad(d, 128, P, 0, 16, 3, gW[0], 0, inv_w)
    gB[0] = cs
    latent(0, gB[0], [(3, s_, g_s), (131, a_, g_a)])
    ctx.pk = None
    out = [g_s.view(1, -1), g_c.view(1, -1), g_a.view(1, -1)]
    for w, b in zip(gW, gB):
        out += [w, b]
    return tuple(out)


class _AutoDecoderFusedFn(torch.autograd.Function):
    """(rays_o, rays_d, viewdirs [R,3], t_vals [R,S], view_enc [R,27], precision, shape, appearance, articulation, w0, b0, ...)
    -> raw [R*S,4]: deformation MLP -> warp -> pos_enc -> trunk -> colour head of the level in ONE launch of the fused render
    kernel (aon_forward_train; latent columns folded into per-call bias stages by aon_fold_latents), every layer output and the
    warped positions written once; the backward is _autodecoder_backward on those planes."""

    @staticmethod
    def forward(ctx, o, d, v, t_vals, view_enc, precision, shape, app, art, *params):
        ctx.x3 = _X3
        R, S = t_vals.shape
        W = [p.detach().contiguous() for p in params[0::2]]
        B = [p.detach().contiguous() for p in params[1::2]]
        s_, c_, a_ = shape.detach().reshape(-1), app.detach().reshape(-1), art.detach().reshape(-1)
        packed = L.pack_weights(L.KIND_AUTODECODER, precision, W, B)
        folded = L.fold_latents(L.KIND_AUTODECODER, precision, packed, s_.float().contiguous(), c_.float().contiguous(),
                                a_.float().contiguous())
        acts, enc, raw, warped = L.forward_train(L.KIND_AUTODECODER, precision, packed, folded, o, d, v, t_vals.contiguous(), S)
        pos = (o[:, None, :] + t_vals[..., None] * d[:, None, :]).reshape(-1, 3)       # helper.py:25-26 (the kernel's cast_rays)
        P = L.pack_rows_tiled(pos, R, S, 16, SA, x3=_X3)
        V = L.pack_rows_tiled(view_enc.detach().contiguous(), R, S, 32, SA, per_ray=True, x3=_X3)
        ctx.pk = (P, acts[0:4], warped, enc, V, acts[4:12], acts[12], acts[13:17])
        ctx.W, ctx.codes, ctx.dims = W, (s_, c_, a_), (R * S, enc.m_tiles, S)
        ctx.tiled = (R, S)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        return (None,) * 6 + _autodecoder_backward(ctx, g_raw, ctx.tiled)


def autodecoder_fused(o, d, v, t_vals, view_enc, latents: dict, mlp, x3: bool = True) -> tuple:
    """rays [R,3] x 3, t_vals [R,S], view_enc [R,27], latents -> (raw_rgb [R,S,3], raw_sigma [R,S,1]) through the fused chain."""
    global _X3
    _X3 = bool(x3)
    params = []
    for lin in mlp.linears():
        params += [lin.weight, lin.bias]
    R, S = t_vals.shape
    raw = _AutoDecoderFusedFn.apply(o, d, v, t_vals, view_enc, L.PREC_TC_F16X3 if x3 else L.PREC_TC_F16, latents["density"],
                                    latents["color"], latents["articulation"], *params)
    return raw[:, :3].reshape(R, S, 3), raw[:, 3:].reshape(R, S, 1)


def autodecoder_mlp(pos: torch.Tensor, view_enc: torch.Tensor, latents: dict, mlp, x3: bool = True) -> tuple:
    """pos [R,S,3] raw sample positions, view_enc [R,27], latents {density, color, articulation} -> (raw_rgb [R,S,3],
    raw_sigma [R,S,1]) through the tcgen05 GEMMs; gradients flow to the parameters and the three codes.
    x3 = False: single fp16 operand planes (fast training mode)."""
    global _X3
    _X3 = bool(x3)
    R, S, _ = pos.shape
    params = []
    for lin in mlp.linears():
        params += [lin.weight, lin.bias]
    raw = _AutoDecoderMLPFn.apply(pos.reshape(-1, 3), view_enc.contiguous(), S, latents["density"], latents["color"],
                                  latents["articulation"], *params)
    return raw[:, :3].reshape(R, S, 3), raw[:, 3:].reshape(R, S, 1)


def vanilla_mlp(enc: torch.Tensor, view_enc: torch.Tensor, S: int, mlp, x3: bool = True) -> tuple:
    """enc [R,S,63] / [M,63], view_enc [R,27] -> (raw_rgb [R,S,3], raw_sigma [R,S,1]) through the tcgen05 GEMMs.
    x3 = False: single fp16 operand planes (fast training mode)."""
    global _X3
    _X3 = bool(x3)
    params = []
    for lin in mlp.linears():
        params += [lin.weight, lin.bias]
    raw = _VanillaMLPFn.apply(enc.reshape(-1, enc.shape[-1]).contiguous(), view_enc.contiguous(), S, *params)
    R = raw.shape[0] // S
    return raw[:, :3].reshape(R, S, 3), raw[:, 3:].reshape(R, S, 1)
