"""Development aid for the Hybrid-Demucs backward: the backward of the time- and frequency-branch encoder
and decoder layers, the framed 2-layer BiLSTM (_BLSTM), the local attention (_LocalState), the spectral front / back end
(_spec / _ispec) and the per-item normalisation, written ONLY in terms of
the primitives the CUDA path has (or will have), checked in fp64 against torch autograd through the torchaudio modules
themselves (TA = torchaudio/models/_hdemucs.py).  Runs on the CPU:

    python tools/hd_bwd_emul.py

Primitives (channel-last activations (B, X, C), like the gemm2 engine's operands):
  conv_taps(A, W, offs, Xout)        out[b,x,n] = sum_tap sum_k A[b, x + offs[tap], k] W[n, tap, k]   (zeros outside A)  = gemm2
  conv_taps_dgrad(G, W, offs, Xin)   = conv_taps(G, W^T, -offs, Xin)                                 = gemm2 on transposed weights
  conv_taps_wgrad(G, A, offs)        dW[n,tap,k] = sum_{b,x} G[b,x,n] A[b, x + offs[tap], k]          = the time-contraction kernel
plus elementwise backward formulas (GELU, GLU, GroupNorm with per-(item, group) sums, LayerScale, residual).
What it pins down for the kernels:
  * strided conv k8 s4 p2 (TA:124) == 3-tap conv on the input VIEWED as (X/4, 4C): the same view of the input gradient makes
    its dgrad a 3-tap launch with offsets (+1, 0, -1) on the transposed repacked weights;
  * transposed conv k8 s4 (TA:243) == 2-tap conv (offsets 0, -1) producing N = 4 Cout viewed as (4X+4, Cout); its dgrad is a
    2-tap launch with offsets (0, +1) reading the output gradient through the same view (crop = zero rows);
  * GroupNorm backward needs exactly two sums per (item, group): sum(dy gamma) and sum(dy gamma xhat);
  * LSTM backward = one GEMM that recomputes every step's gates from the saved h, an elementwise scan for the cell state,
    and a reverse-time chain with W_hh^T that has the forward recurrence kernel's structure; dW_hh / dW_ih are time
    contractions; the 200 / 100 framing's adjoint is a scatter (stitch) and an overlap-add (unfold);
  * local attention backward per (item, head) from q, k, content, decay and the recomputed softmax, diagonal masked;
  * `_spec` backward = adjoint rfft of the kept frames, overlap-add into the once-reflected signal, reflection folded back;
    `_ispec` backward = a forward STFT of (g / envelope) with a per-bin factor c_k / sqrt(N) (c_0 = 1, else 2).
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)


# ----------------------------------------------------------------------------------------------- primitives
def conv_taps(A, W, offs, Xout):
    B, Xin, K = A.shape
    N = W.shape[0]
    out = A.new_zeros(B, Xout, N)
    for tap, off in enumerate(offs):
        lo, hi = max(0, -off), min(Xout, Xin - off)  # output rows whose source row x + off lies inside A
        if hi > lo:
            out[:, lo:hi] += A[:, lo + off:hi + off] @ W[:, tap].T
    return out


def conv_taps_dgrad(G, W, offs, Xin):
    return conv_taps(G, W.permute(2, 1, 0).contiguous(), [-o for o in offs], Xin)


def conv_taps_wgrad(G, A, offs):
    B, Xout, N = G.shape
    Xin, K = A.shape[1], A.shape[2]
    dW = G.new_zeros(N, len(offs), K)
    for tap, off in enumerate(offs):
        lo, hi = max(0, -off), min(Xout, Xin - off)
        if hi > lo:
            dW[:, tap] = torch.einsum("bxn,bxk->nk", G[:, lo:hi], A[:, lo + off:hi + off])
    return dW


def gelu_bwd(x, g):
    return g * (0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi))


def glu_fwd(z):  # channel-last: value = first half, gate = second half (F.glu(dim=channels))
    a, b = z.chunk(2, dim=-1)
    return a * torch.sigmoid(b)


def glu_bwd(z, g):
    a, b = z.chunk(2, dim=-1)
    s = torch.sigmoid(b)
    return torch.cat([g * s, g * a * s * (1 - s)], dim=-1)


def gn_fwd(x, G, gamma, beta, eps=1e-5):
    B, X, C = x.shape
    xg = x.reshape(B, X, G, C // G)
    mu = xg.mean(dim=(1, 3), keepdim=True)
    var = xg.var(dim=(1, 3), unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    xhat = ((xg - mu) * rstd).reshape(B, X, C)
    return xhat * gamma + beta, (xhat, rstd)


def gn_bwd(g, saved, G, gamma):
    xhat, rstd = saved
    B, X, C = g.shape
    dg = g * gamma
    n = X * (C // G)
    s1 = dg.reshape(B, X, G, C // G).sum(dim=(1, 3), keepdim=True) / n              # the two sums per (item, group)
    s2 = (dg * xhat).reshape(B, X, G, C // G).sum(dim=(1, 3), keepdim=True) / n
    dx = (rstd * (dg.reshape(B, X, G, C // G) - s1 - xhat.reshape(B, X, G, C // G) * s2)).reshape(B, X, C)
    return dx, (g * xhat).sum(dim=(0, 1)), g.sum(dim=(0, 1))


# ----------------------------------------------------------------------------------------------- weight repacking
def pack_strided(w):
    """Conv1d weight (Co, Ci, 8), stride 4, pad 2 -> 3-tap weight (Co, 3, 4 Ci) for the (X/4, 4 Ci) view; offsets (-1, 0, +1)."""
    Co, Ci, k = w.shape
    W3 = w.new_zeros(Co, 3, 4 * Ci)
    for tap in range(3):
        for q in range(4):
            j = 4 * (tap - 1) + q + 2
            if 0 <= j < k:
                W3[:, tap, q * Ci:(q + 1) * Ci] = w[:, :, j]
    return W3


def unpack_strided_grad(dW3, Ci):
    Co = dW3.shape[0]
    dw = dW3.new_zeros(Co, Ci, 8)
    for tap in range(3):
        for q in range(4):
            j = 4 * (tap - 1) + q + 2
            if 0 <= j < 8:
                dw[:, :, j] = dW3[:, tap, q * Ci:(q + 1) * Ci]
    return dw


def pack_transposed(w):
    """ConvTranspose1d weight (Ci, Co, 8), stride 4 -> 2-tap weight (4 Co, 2, Ci); offsets (0, -1); row q Co + co."""
    Ci, Co, k = w.shape
    W2 = w.new_zeros(4 * Co, 2, Ci)
    for q in range(4):
        W2[q * Co:(q + 1) * Co, 0] = w[:, :, q].T
        W2[q * Co:(q + 1) * Co, 1] = w[:, :, q + 4].T
    return W2


def unpack_transposed_grad(dW2, Co):
    Ci = dW2.shape[2]
    dw = dW2.new_zeros(Ci, Co, 8)
    for q in range(4):
        dw[:, :, q] = dW2[q * Co:(q + 1) * Co, 0].T
        dw[:, :, q + 4] = dW2[q * Co:(q + 1) * Co, 1].T
    return dw


# ----------------------------------------------------------------------------------------------- DConv branch (no LSTM / attention)
def dconv_fwd(x, layers):
    saved = []
    for d, p in enumerate(layers):
        dil = 2 ** d
        h0 = conv_taps(x, p["w1"], [-dil, 0, dil], x.shape[1]) + p["b1"]
        h1, s1 = gn_fwd(h0, 1, p["g1"], p["be1"])
        h2 = F.gelu(h1)
        h3 = conv_taps(h2, p["w2"], [0], x.shape[1]) + p["b2"]
        h4, s2 = gn_fwd(h3, 1, p["g2"], p["be2"])
        h5 = glu_fwd(h4)
        saved.append((x, h1, s1, h2, h4, s2, h5))
        x = x + p["scale"] * h5
    return x, saved


def dconv_bwd(g, layers, saved):
    grads = []
    for d in reversed(range(len(layers))):
        p, (x, h1, s1, h2, h4, s2, h5) = layers[d], saved[d]
        dil = 2 ** d
        gr = {"scale": (g * h5).sum(dim=(0, 1))}
        g5 = g * p["scale"]
        g4 = glu_bwd(h4, g5)
        g3, gr["g2"], gr["be2"] = gn_bwd(g4, s2, 1, p["g2"])
        gr["b2"] = g3.sum(dim=(0, 1))
        gr["w2"] = conv_taps_wgrad(g3, h2, [0])
        g2 = conv_taps_dgrad(g3, p["w2"], [0], x.shape[1])
        g1 = gelu_bwd(h1, g2)
        g0, gr["g1"], gr["be1"] = gn_bwd(g1, s1, 1, p["g1"])
        gr["b1"] = g0.sum(dim=(0, 1))
        gr["w1"] = conv_taps_wgrad(g0, x, [-dil, 0, dil])
        g = g + conv_taps_dgrad(g0, p["w1"], [-dil, 0, dil], x.shape[1])
        grads.append(gr)
    return g, grads[::-1]


def dconv_params(mod):
    out = []
    for layer in mod.layers:
        c1, n1, _, c2, n2, _, ls = layer
        out.append(dict(w1=c1.weight.detach().permute(0, 2, 1).contiguous(), b1=c1.bias.detach(), g1=n1.weight.detach(), be1=n1.bias.detach(),
                        w2=c2.weight.detach().permute(0, 2, 1).contiguous(), b2=c2.bias.detach(), g2=n2.weight.detach(), be2=n2.bias.detach(),
                        scale=ls.scale.detach()))
    return out


def rel(a, b):
    if float(a.norm()) < 1e-12 and float(b.norm()) < 1e-12:  # e.g. a conv bias in front of a one-channel-per-group GroupNorm: exactly 0
        return 0.0
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


# ----------------------------------------------------------------------------------------------- encoder layer (time branch)
def check_encoder(norm: bool):
    from torchaudio.models._hdemucs import _HEncLayer

    torch.manual_seed(0)
    Ci, Co, B, L = 4, 8, 2, 64
    enc = _HEncLayer(Ci, Co, freq=False, norm_type="group_norm" if norm else "none", norm_groups=4, dconv_kw=dict(depth=2, compress=4, init=0.3))
    for p in enc.parameters():
        p.data.add_(0.1 * torch.randn_like(p))
    x = torch.randn(B, Ci, L, requires_grad=True)
    z = enc(x)
    r = torch.randn_like(z)
    (z * r).sum().backward()

    # ---- forward in primitives (channel-last)
    xl = x.detach().permute(0, 2, 1).contiguous()                      # (B, L, Ci)
    W3 = pack_strided(enc.conv.weight.detach())
    y0 = conv_taps(xl.reshape(B, L // 4, 4 * Ci), W3, [-1, 0, 1], L // 4) + enc.conv.bias.detach()
    if norm:
        y1, sn1 = gn_fwd(y0, 4, enc.norm1.weight.detach(), enc.norm1.bias.detach())
    else:
        y1 = y0
    y2 = F.gelu(y1)
    layers = dconv_params(enc.dconv)
    y3, sdc = dconv_fwd(y2, layers)
    Wr = enc.rewrite.weight.detach().permute(0, 2, 1).contiguous()      # (2Co, 1, Co)
    z0 = conv_taps(y3, Wr, [0], y3.shape[1]) + enc.rewrite.bias.detach()
    if norm:
        z1, sn2 = gn_fwd(z0, 4, enc.norm2.weight.detach(), enc.norm2.bias.detach())
    else:
        z1 = z0
    zz = glu_fwd(z1)
    print(f"encoder(norm={norm}) forward   {rel(zz, z.detach().permute(0, 2, 1)):.1e}")

    # ---- backward in primitives
    g = glu_bwd(z1, r.permute(0, 2, 1))
    errs = {}
    if norm:
        g, dg2, db2 = gn_bwd(g, sn2, 4, enc.norm2.weight.detach())
        errs["norm2.weight"], errs["norm2.bias"] = rel(dg2, enc.norm2.weight.grad), rel(db2, enc.norm2.bias.grad)
    errs["rewrite.bias"] = rel(g.sum(dim=(0, 1)), enc.rewrite.bias.grad)
    errs["rewrite.weight"] = rel(conv_taps_wgrad(g, y3, [0]).permute(0, 2, 1), enc.rewrite.weight.grad)
    g = conv_taps_dgrad(g, Wr, [0], y3.shape[1])
    g, gdc = dconv_bwd(g, layers, sdc)
    for d, gr in enumerate(gdc):
        c1, n1, _, c2, n2, _, ls = enc.dconv.layers[d]
        errs[f"dconv{d}"] = max(rel(gr["w1"].permute(0, 2, 1), c1.weight.grad), rel(gr["b1"], c1.bias.grad), rel(gr["g1"], n1.weight.grad),
                                rel(gr["be1"], n1.bias.grad), rel(gr["w2"].permute(0, 2, 1), c2.weight.grad), rel(gr["b2"], c2.bias.grad),
                                rel(gr["g2"], n2.weight.grad), rel(gr["be2"], n2.bias.grad), rel(gr["scale"], ls.scale.grad))
    g = gelu_bwd(y1, g)
    if norm:
        g, dg1, db1 = gn_bwd(g, sn1, 4, enc.norm1.weight.detach())
        errs["norm1.weight"], errs["norm1.bias"] = rel(dg1, enc.norm1.weight.grad), rel(db1, enc.norm1.bias.grad)
    errs["conv.bias"] = rel(g.sum(dim=(0, 1)), enc.conv.bias.grad)
    errs["conv.weight"] = rel(unpack_strided_grad(conv_taps_wgrad(g, xl.reshape(B, L // 4, 4 * Ci), [-1, 0, 1]), Ci), enc.conv.weight.grad)
    dx = conv_taps_dgrad(g, W3, [-1, 0, 1], L // 4).reshape(B, L, Ci)   # the (X/4, 4 Ci) view is its own adjoint
    errs["input"] = rel(dx, x.grad.permute(0, 2, 1))
    print(f"encoder(norm={norm}) backward  max {max(errs.values()):.1e}   " + ", ".join(f"{k} {v:.0e}" for k, v in errs.items()))
    return max(errs.values())


# ----------------------------------------------------------------------------------------------- decoder layer (time branch)
def check_decoder(norm: bool, last: bool):
    from torchaudio.models._hdemucs import _HDecLayer

    torch.manual_seed(1)
    Ci, Co, B, X = 8, 4, 2, 16
    dec = _HDecLayer(Ci, Co, last=last, freq=False, norm_type="group_norm" if norm else "none", norm_groups=4, context=1, empty=False,
                     dconv_kw=dict(depth=2, compress=4, init=0.3))
    for p in dec.parameters():
        p.data.add_(0.1 * torch.randn_like(p))
    x = torch.randn(B, Ci, X, requires_grad=True)
    skip = torch.randn(B, Ci, X, requires_grad=True)
    length = 4 * X
    z, _ = dec(x, skip, length)
    r = torch.randn_like(z)
    (z * r).sum().backward()

    # TA:252-298 for the time branch: y = GLU(norm1(rewrite_k3(x + skip))); z = conv_tr(y); z = norm2(z); z = z[..., 2:2+length]; GELU unless last
    a = (x + skip).detach().permute(0, 2, 1).contiguous()
    Wr = dec.rewrite.weight.detach().permute(0, 2, 1).contiguous()      # (2Ci, 3, Ci)
    y0 = conv_taps(a, Wr, [-1, 0, 1], X) + dec.rewrite.bias.detach()
    if norm:
        y1, sn1 = gn_fwd(y0, 4, dec.norm1.weight.detach(), dec.norm1.bias.detach())
    else:
        y1 = y0
    y2 = glu_fwd(y1)
    W2 = pack_transposed(dec.conv_tr.weight.detach())
    u = conv_taps(y2, W2, [0, -1], X + 1).reshape(B, 4 * X + 4, Co) + dec.conv_tr.bias.detach()
    if norm:
        u1, sn2 = gn_fwd(u, 4, dec.norm2.weight.detach(), dec.norm2.bias.detach())   # statistics over the UNcropped signal (TA:287-288)
    else:
        u1 = u
    v = u1[:, 2:2 + length]
    zz = v if last else F.gelu(v)
    print(f"decoder(norm={norm}, last={last}) forward   {rel(zz, z.detach().permute(0, 2, 1)):.1e}")

    g = r.permute(0, 2, 1)
    if not last:
        g = gelu_bwd(v, g)
    gfull = g.new_zeros(B, 4 * X + 4, Co)
    gfull[:, 2:2 + length] = g                                           # crop backward = zero rows
    errs = {}
    if norm:
        gfull, dg2, db2 = gn_bwd(gfull, sn2, 4, dec.norm2.weight.detach())
        errs["norm2.weight"], errs["norm2.bias"] = rel(dg2, dec.norm2.weight.grad), rel(db2, dec.norm2.bias.grad)
    errs["conv_tr.bias"] = rel(gfull.sum(dim=(0, 1)), dec.conv_tr.bias.grad)
    gv = gfull.reshape(B, X + 1, 4 * Co)                                 # the (4X+4, Co) <-> (X+1, 4 Co) view is its own adjoint
    errs["conv_tr.weight"] = rel(unpack_transposed_grad(conv_taps_wgrad(gv, y2, [0, -1]), Co), dec.conv_tr.weight.grad)
    g = conv_taps_dgrad(gv, W2, [0, -1], X)
    g = glu_bwd(y1, g)
    if norm:
        g, dg1, db1 = gn_bwd(g, sn1, 4, dec.norm1.weight.detach())
        errs["norm1.weight"], errs["norm1.bias"] = rel(dg1, dec.norm1.weight.grad), rel(db1, dec.norm1.bias.grad)
    errs["rewrite.bias"] = rel(g.sum(dim=(0, 1)), dec.rewrite.bias.grad)
    errs["rewrite.weight"] = rel(conv_taps_wgrad(g, a, [-1, 0, 1]).permute(0, 2, 1), dec.rewrite.weight.grad)
    dx = conv_taps_dgrad(g, Wr, [-1, 0, 1], X)
    errs["input"], errs["skip"] = rel(dx, x.grad.permute(0, 2, 1)), rel(dx, skip.grad.permute(0, 2, 1))
    print(f"decoder(norm={norm}, last={last}) backward  max {max(errs.values()):.1e}   " + ", ".join(f"{k} {v:.0e}" for k, v in errs.items()))
    return max(errs.values())


# ----------------------------------------------------------------------------------------------- encoder layer (frequency branch)
def check_freq_encoder():
    """Frequency-branch encoder layer in the CUDA path's layout (B, T, Fr, C): the strided conv runs along Fr (the (Fr/4, 4C) view
    is free because Fr and C are adjacent), the DConv branch runs along T on (b, fr) rows, GroupNorm(4) takes its statistics over
    (C/4, Fr, T) per item.  Folding (B, T) resp. (B, Fr) into the batch reduces everything to the 1-D primitives above."""
    from torchaudio.models._hdemucs import _HEncLayer

    torch.manual_seed(6)
    Ci, Co, B, Fr, T = 4, 8, 2, 16, 12
    enc = _HEncLayer(Ci, Co, freq=True, norm_type="group_norm", norm_groups=4, dconv_kw=dict(depth=2, compress=4, init=0.3))
    for p in enc.parameters():
        p.data.add_(0.1 * torch.randn_like(p))
    x = torch.randn(B, Ci, Fr, T, requires_grad=True)
    z = enc(x)                                                            # (B, Co, Fr/4, T)
    r = torch.randn_like(z)
    (z * r).sum().backward()
    Fo = Fr // 4

    def gn4_fwd(v, gamma, beta):                                          # v: (B, T, F, C): statistics over (T, F, C/4) per item
        Bv, Tv, Fv, Cv = v.shape
        y, sv = gn_fwd(v.reshape(Bv, Tv * Fv, Cv), 4, gamma, beta)
        return y.reshape(v.shape), sv

    def gn4_bwd(g, sv, gamma):
        Bv, Tv, Fv, Cv = g.shape
        dx, dgam, dbet = gn_bwd(g.reshape(Bv, Tv * Fv, Cv), sv, 4, gamma)
        return dx.reshape(g.shape), dgam, dbet

    xl = x.detach().permute(0, 3, 2, 1).contiguous()                      # (B, T, Fr, Ci)
    W3 = pack_strided(enc.conv.weight.detach()[:, :, :, 0])
    y0 = (conv_taps(xl.reshape(B * T, Fr // 4, 4 * Ci), W3, [-1, 0, 1], Fo) + enc.conv.bias.detach()).reshape(B, T, Fo, Co)
    y1, sn1 = gn4_fwd(y0, enc.norm1.weight.detach(), enc.norm1.bias.detach())
    y2 = F.gelu(y1)
    rows = y2.permute(0, 2, 1, 3).reshape(B * Fo, T, Co)                  # (b, fr) rows along T
    layers = dconv_params(enc.dconv)
    y3r, sdc = dconv_fwd(rows, layers)
    y3 = y3r.reshape(B, Fo, T, Co).permute(0, 2, 1, 3)
    Wr = enc.rewrite.weight.detach()[:, :, 0, 0][:, None, :]             # (2Co, 1, Co)
    z0 = (conv_taps(y3.reshape(B, T * Fo, Co), Wr, [0], T * Fo) + enc.rewrite.bias.detach()).reshape(B, T, Fo, 2 * Co)
    z1, sn2 = gn4_fwd(z0, enc.norm2.weight.detach(), enc.norm2.bias.detach())
    zz = glu_fwd(z1)
    print(f"freq encoder forward   {rel(zz, z.detach().permute(0, 3, 2, 1)):.1e}")

    errs = {}
    g = glu_bwd(z1, r.permute(0, 3, 2, 1))
    g, dg2, db2 = gn4_bwd(g, sn2, enc.norm2.weight.detach())
    errs["norm2"] = max(rel(dg2, enc.norm2.weight.grad), rel(db2, enc.norm2.bias.grad))
    gf = g.reshape(B, T * Fo, 2 * Co)
    errs["rewrite"] = max(rel(conv_taps_wgrad(gf, y3.reshape(B, T * Fo, Co), [0])[:, 0], enc.rewrite.weight.grad[:, :, 0, 0]),
                          rel(gf.sum(dim=(0, 1)), enc.rewrite.bias.grad))
    g = conv_taps_dgrad(gf, Wr, [0], T * Fo).reshape(B, T, Fo, Co)
    gr, gdc = dconv_bwd(g.permute(0, 2, 1, 3).reshape(B * Fo, T, Co), layers, sdc)
    for d, gd in enumerate(gdc):
        c1, n1, _, c2, n2, _, ls = enc.dconv.layers[d]
        errs[f"dconv{d}"] = max(rel(gd["w1"].permute(0, 2, 1), c1.weight.grad), rel(gd["g1"], n1.weight.grad), rel(gd["w2"].permute(0, 2, 1), c2.weight.grad),
                                rel(gd["g2"], n2.weight.grad), rel(gd["scale"], ls.scale.grad))
    g = gr.reshape(B, Fo, T, Co).permute(0, 2, 1, 3)
    g = gelu_bwd(y1, g)
    g, dg1, db1 = gn4_bwd(g, sn1, enc.norm1.weight.detach())
    errs["norm1"] = max(rel(dg1, enc.norm1.weight.grad), rel(db1, enc.norm1.bias.grad))
    gb = g.reshape(B * T, Fo, Co)
    errs["conv"] = max(rel(unpack_strided_grad(conv_taps_wgrad(gb, xl.reshape(B * T, Fr // 4, 4 * Ci), [-1, 0, 1]), Ci), enc.conv.weight.grad[:, :, :, 0]),
                       rel(gb.sum(dim=(0, 1)), enc.conv.bias.grad))
    dx = conv_taps_dgrad(gb, W3, [-1, 0, 1], Fr // 4).reshape(B, T, Fr, Ci)
    errs["input"] = rel(dx, x.grad.permute(0, 3, 2, 1))
    print(f"freq encoder backward  max {max(errs.values()):.1e}   " + ", ".join(f"{k} {v:.0e}" for k, v in errs.items()))
    return max(errs.values())


# ----------------------------------------------------------------------------------------------- decoder layer (frequency branch)
def conv_taps2d(A, W, offs, Yout, Xout):
    """2-D tap form of the engine: A (B, Y, X, K), offs = [(dy, dx)], out[b,y,x,n] = sum_tap A[b, y+dy, x+dx, :] . W[n, tap, :]."""
    B, Yin, Xin, K = A.shape
    out = A.new_zeros(B, Yout, Xout, W.shape[0])
    for tap, (dy, dx) in enumerate(offs):
        y0, y1, x0, x1 = max(0, -dy), min(Yout, Yin - dy), max(0, -dx), min(Xout, Xin - dx)
        if y1 > y0 and x1 > x0:
            out[:, y0:y1, x0:x1] += A[:, y0 + dy:y1 + dy, x0 + dx:x1 + dx] @ W[:, tap].T
    return out


def conv_taps2d_wgrad(G, A, offs):
    B, Yout, Xout, N = G.shape
    Yin, Xin, K = A.shape[1:]
    dW = G.new_zeros(N, len(offs), K)
    for tap, (dy, dx) in enumerate(offs):
        y0, y1, x0, x1 = max(0, -dy), min(Yout, Yin - dy), max(0, -dx), min(Xout, Xin - dx)
        if y1 > y0 and x1 > x0:
            dW[:, tap] = torch.einsum("byxn,byxk->nk", G[:, y0:y1, x0:x1], A[:, y0 + dy:y1 + dy, x0 + dx:x1 + dx])
    return dW


def check_freq_decoder():
    """Frequency-branch decoder layer, layout (B, T, Fr, C): 3x3 `rewrite` = 9 two-dimensional taps, GLU, transposed conv along Fr
    as the 2-tap N = 4 Cout launch, GroupNorm over the UNcropped (Fr 4X+4) tensor, crop [2:-2] along Fr, GELU."""
    from torchaudio.models._hdemucs import _HDecLayer

    torch.manual_seed(7)
    Ci, Co, B, X, T = 8, 4, 2, 4, 10
    dec = _HDecLayer(Ci, Co, last=False, freq=True, norm_type="group_norm", norm_groups=4, context=1)
    for p in dec.parameters():
        p.data.add_(0.1 * torch.randn_like(p))
    x = torch.randn(B, Ci, X, T, requires_grad=True)
    skip = torch.randn(B, Ci, X, T, requires_grad=True)
    z, _ = dec(x, skip, 4 * X)
    r = torch.randn_like(z)                                               # (B, Co, 4X, T)
    (z * r).sum().backward()

    def gn4_fwd(v, gamma, beta):
        Bv, Tv, Fv, Cv = v.shape
        y, sv = gn_fwd(v.reshape(Bv, Tv * Fv, Cv), 4, gamma, beta)
        return y.reshape(v.shape), sv

    def gn4_bwd(g, sv, gamma):
        Bv, Tv, Fv, Cv = g.shape
        dx, dgam, dbet = gn_bwd(g.reshape(Bv, Tv * Fv, Cv), sv, 4, gamma)
        return dx.reshape(g.shape), dgam, dbet

    a = (x + skip).detach().permute(0, 3, 2, 1).contiguous()              # (B, T, Fr, Ci): Y = T, X = Fr
    offs = [(dt, df) for df in (-1, 0, 1) for dt in (-1, 0, 1)]            # Conv2d weight (2Ci, Ci, kF, kT): tap (df, dt)
    wr = dec.rewrite.weight.detach()
    Wr = torch.stack([wr[:, :, df + 1, dt + 1] for (dt, df) in offs], dim=1)   # (2Ci, 9, Ci)
    y0 = conv_taps2d(a, Wr, offs, T, X) + dec.rewrite.bias.detach()
    y1, sn1 = gn4_fwd(y0, dec.norm1.weight.detach(), dec.norm1.bias.detach())
    y2 = glu_fwd(y1)
    W2 = pack_transposed(dec.conv_tr.weight.detach()[:, :, :, 0])
    u = (conv_taps(y2.reshape(B * T, X, Ci), W2, [0, -1], X + 1).reshape(B, T, 4 * X + 4, Co) + dec.conv_tr.bias.detach())
    u1, sn2 = gn4_fwd(u, dec.norm2.weight.detach(), dec.norm2.bias.detach())
    v = u1[:, :, 2:-2]
    zz = F.gelu(v)
    print(f"freq decoder forward   {rel(zz, z.detach().permute(0, 3, 2, 1)):.1e}")

    errs = {}
    g = gelu_bwd(v, r.permute(0, 3, 2, 1))
    gfull = g.new_zeros(B, T, 4 * X + 4, Co)
    gfull[:, :, 2:-2] = g
    gfull, dg2, db2 = gn4_bwd(gfull, sn2, dec.norm2.weight.detach())
    errs["norm2"] = max(rel(dg2, dec.norm2.weight.grad), rel(db2, dec.norm2.bias.grad))
    gv = gfull.reshape(B * T, X + 1, 4 * Co)
    errs["conv_tr"] = rel(unpack_transposed_grad(conv_taps_wgrad(gv, y2.reshape(B * T, X, Ci), [0, -1]), Co), dec.conv_tr.weight.grad[:, :, :, 0])
    g = conv_taps_dgrad(gv, W2, [0, -1], X).reshape(B, T, X, Ci)
    g = glu_bwd(y1, g)
    g, dg1, db1 = gn4_bwd(g, sn1, dec.norm1.weight.detach())
    errs["norm1"] = max(rel(dg1, dec.norm1.weight.grad), rel(db1, dec.norm1.bias.grad))
    dWr = conv_taps2d_wgrad(g, a, offs)
    dwr = torch.zeros_like(wr)
    for tap, (dt, df) in enumerate(offs):
        dwr[:, :, df + 1, dt + 1] = dWr[:, tap]
    errs["rewrite"] = max(rel(dwr, dec.rewrite.weight.grad), rel(g.sum(dim=(0, 1, 2)), dec.rewrite.bias.grad))
    dx = conv_taps2d(g, Wr.permute(2, 1, 0).contiguous(), [(-dt, -df) for (dt, df) in offs], T, X)   # dgrad: transposed weights, negated taps
    errs["input"] = max(rel(dx, x.grad.permute(0, 3, 2, 1)), rel(dx, skip.grad.permute(0, 3, 2, 1)))
    print(f"freq decoder backward  max {max(errs.values()):.1e}   " + ", ".join(f"{k} {v:.0e}" for k, v in errs.items()))
    return max(errs.values())


# ----------------------------------------------------------------------------------------------- _BLSTM (TA:742-788)
def lstm_dir_fwd(Gx, Whh, reverse):
    """One direction of one layer, the recurrence kernel's job: Gx (T, B, 4H) = W_ih x + b_ih + b_hh; gate order i, f, g, o.
    Returns h (T, B, H) and what the backward keeps: h only (gates and c are recomputed)."""
    T, B, H4 = Gx.shape
    H = H4 // 4
    h = Gx.new_zeros(B, H)
    c = Gx.new_zeros(B, H)
    hs = Gx.new_zeros(T, B, H)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = Gx[t] + h @ Whh.T
        i, f, gg, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        hs[t] = h
    return hs


def lstm_dir_bwd(Gx, Whh, hs, dhs, reverse):
    """Backward of lstm_dir_fwd in three kernel-shaped steps:
      1. gates for ALL steps at once: Gfull = Gx + h_prev W_hh^T  (one GEMM over the saved h, no recurrence);
      2. the cell state by an elementwise scan c_t = f c_prev + i g  (no matrix product);
      3. the reverse-time chain: dh_total = dh_t + W_hh^T dG_{t+1}, gate derivatives, dc carry  (matvec with W_hh^T per step:
         the forward recurrence kernel's structure with the transposed matrix).
    Returns dGx (T, B, 4H) -- from which dW_ih, db, dx follow as GEMMs / time contractions -- and dW_hh = sum_t dG_t (x) h_prev."""
    T, B, H4 = Gx.shape
    H = H4 // 4
    step = -1 if reverse else 1
    h_prev = torch.zeros_like(hs)
    if reverse:
        h_prev[:-1] = hs[1:]
    else:
        h_prev[1:] = hs[:-1]
    Gfull = Gx + h_prev @ Whh.T                                                         # step 1
    i, f = torch.sigmoid(Gfull[..., :H]), torch.sigmoid(Gfull[..., H:2 * H])
    gg, o = torch.tanh(Gfull[..., 2 * H:3 * H]), torch.sigmoid(Gfull[..., 3 * H:])
    cs = torch.zeros_like(hs)                                                          # step 2
    c = Gx.new_zeros(B, H)
    order = list(range(T - 1, -1, -1) if reverse else range(T))
    for t in order:
        c = f[t] * c + i[t] * gg[t]
        cs[t] = c
    dGx = torch.zeros_like(Gx)                                                         # step 3
    dh_next = Gx.new_zeros(B, H)
    dc = Gx.new_zeros(B, H)
    for t in reversed(order):
        dh = dhs[t] + dh_next
        tc = torch.tanh(cs[t])
        c_prev = cs[t - step] if 0 <= t - step < T else torch.zeros_like(dc)
        dc = dc + dh * o[t] * (1 - tc * tc)
        dG = torch.cat([dc * gg[t] * i[t] * (1 - i[t]), dc * c_prev * f[t] * (1 - f[t]), dc * i[t] * (1 - gg[t] * gg[t]),
                        dh * tc * o[t] * (1 - o[t])], dim=-1)
        dGx[t] = dG
        dh_next = dG @ Whh
        dc = dc * f[t]
    dWhh = torch.einsum("tbg,tbh->gh", dGx, h_prev)                                    # time contraction
    return dGx, dWhh


def check_blstm():
    from torchaudio.models._hdemucs import _BLSTM

    torch.manual_seed(2)
    H, B, T = 6, 2, 230                                                                # T > 200: the framed path (width 200, stride 100)
    mod = _BLSTM(H, layers=2, skip=True)
    x = torch.randn(B, H, T, requires_grad=True)
    y = mod(x)
    r = torch.randn_like(y)
    (y * r).sum().backward()
    P = {k: v.detach() for k, v in mod.lstm.named_parameters()}

    # ---- forward: frames (B * nframes, width) rows in time-major (T, Bf, H) form
    width, stride = 200, 100
    nframes = math.ceil(T / stride)
    pad_len = (nframes - 1) * stride + width
    xp = F.pad(x.detach(), (0, pad_len - T))
    fr = torch.stack([xp[:, :, k * stride:k * stride + width] for k in range(nframes)], dim=1)      # (B, nf, H, width)
    inp = fr.reshape(B * nframes, H, width).permute(2, 0, 1).contiguous()                           # (width, Bf, H)
    acts = [inp]
    saved = []
    for layer in range(2):
        outs = []
        for d, suf in enumerate(["", "_reverse"]):
            Wih, Whh = P[f"weight_ih_l{layer}{suf}"], P[f"weight_hh_l{layer}{suf}"]
            Gx = acts[-1] @ Wih.T + P[f"bias_ih_l{layer}{suf}"] + P[f"bias_hh_l{layer}{suf}"]
            hs = lstm_dir_fwd(Gx, Whh, reverse=bool(d))
            outs.append(hs)
            saved.append((Gx, hs))
        acts.append(torch.cat(outs, dim=-1))
    lin = acts[-1] @ mod.linear.weight.detach().T + mod.linear.bias.detach()                        # (width, Bf, H)
    frames = lin.permute(1, 2, 0).reshape(B, nframes, H, width)
    limit = stride // 2
    pieces = [frames[:, k, :, (0 if k == 0 else limit):(width if k == nframes - 1 else width - limit)] for k in range(nframes)]
    out = torch.cat(pieces, -1)[..., :T] + x.detach()
    print(f"blstm forward   {rel(out, y.detach()):.1e}")

    # ---- backward
    errs = {}
    g = r                                                                                           # (B, H, T)
    gfr = torch.zeros(B, nframes, H, width)
    pos = 0
    for k in range(nframes):                                                                        # adjoint of the stitching: scatter
        lo, hi = (0 if k == 0 else limit), (width if k == nframes - 1 else width - limit)
        n = min(hi - lo, T - pos)
        if n > 0:
            gfr[:, k, :, lo:lo + n] = g[..., pos:pos + n]
        pos += hi - lo
    glin = gfr.reshape(B * nframes, H, width).permute(2, 0, 1)                                      # (width, Bf, H)
    errs["linear.weight"] = rel(torch.einsum("tbo,tbi->oi", glin, acts[-1]), mod.linear.weight.grad)
    errs["linear.bias"] = rel(glin.sum(dim=(0, 1)), mod.linear.bias.grad)
    gact = glin @ mod.linear.weight.detach()
    for layer in (1, 0):
        gin = torch.zeros_like(acts[layer])
        for d, suf in enumerate(["", "_reverse"]):
            Wih, Whh = P[f"weight_ih_l{layer}{suf}"], P[f"weight_hh_l{layer}{suf}"]
            Gx, hs = saved[2 * layer + d]
            dGx, dWhh = lstm_dir_bwd(Gx, Whh, hs, gact[..., d * H:(d + 1) * H], reverse=bool(d))
            G = dict(mod.lstm.named_parameters())
            errs[f"l{layer}{suf}"] = max(rel(dWhh, G[f"weight_hh_l{layer}{suf}"].grad),
                                         rel(torch.einsum("tbg,tbi->gi", dGx, acts[layer]), G[f"weight_ih_l{layer}{suf}"].grad),
                                         rel(dGx.sum(dim=(0, 1)), G[f"bias_ih_l{layer}{suf}"].grad),
                                         rel(dGx.sum(dim=(0, 1)), G[f"bias_hh_l{layer}{suf}"].grad))
            gin = gin + dGx @ Wih
        gact = gin
    gfr_in = gact.permute(1, 2, 0).reshape(B, nframes, H, width)
    dxp = torch.zeros(B, H, pad_len)
    for k in range(nframes):                                                                        # adjoint of _unfold: overlap-add
        dxp[:, :, k * stride:k * stride + width] += gfr_in[:, k]
    dx = dxp[..., :T] + r                                                                           # + skip
    errs["input"] = rel(dx, x.grad)
    print(f"blstm backward  max {max(errs.values()):.1e}   " + ", ".join(f"{k} {v:.0e}" for k, v in errs.items()))
    return max(errs.values())


# ----------------------------------------------------------------------------------------------- _LocalState (TA:822-857)
def check_local_state():
    from torchaudio.models._hdemucs import _LocalState

    torch.manual_seed(3)
    C, heads, nd, B, T = 8, 4, 4, 2, 24
    mod = _LocalState(C, heads=heads, ndecay=nd)
    for p in mod.parameters():
        p.data.add_(0.2 * torch.randn_like(p))
    x = torch.randn(B, C, T, requires_grad=True)
    y = mod(x)
    r = torch.randn_like(y)
    (y * r).sum().backward()

    xl = x.detach().permute(0, 2, 1)                                         # (B, T, C)
    lin = lambda m: xl @ m.weight.detach()[:, :, 0].T + m.bias.detach()      # 1x1 convs = the projection GEMMs
    q, k, ct, dr = lin(mod.query), lin(mod.key), lin(mod.content), lin(mod.query_decay)
    dh = C // heads
    idx = torch.arange(T, dtype=x.dtype)
    dist = (idx[:, None] - idx[None, :]).abs()                               # [t (key), s (query)]
    kern = -torch.arange(1, nd + 1, dtype=x.dtype).view(-1, 1, 1) * dist / math.sqrt(nd)     # (f, t, s)
    res = torch.zeros(B, T, C)
    saved = {}
    for b in range(B):
        for h in range(heads):                                               # one CTA per (item, head): everything below lives in smem
            qh, kh, ch = q[b, :, h * dh:(h + 1) * dh], k[b, :, h * dh:(h + 1) * dh], ct[b, :, h * dh:(h + 1) * dh]       # (T, dh)
            sg = torch.sigmoid(dr[b, :, h * nd:(h + 1) * nd])                # (s, f)
            dots = kh @ qh.T / math.sqrt(dh) + torch.einsum("fts,sf->ts", kern, sg / 2)
            dots.fill_diagonal_(-100.0)
            W = torch.softmax(dots, dim=0)                                   # over keys t
            res[b, :, h * dh:(h + 1) * dh] = W.T @ ch                        # result[s, c] = sum_t W[t, s] content[t, c]
            saved[b, h] = (qh, kh, ch, sg, W)
    out = xl + res @ mod.proj.weight.detach()[:, :, 0].T + mod.proj.bias.detach()
    print(f"local_state forward   {rel(out, y.detach().permute(0, 2, 1)):.1e}")

    g = r.permute(0, 2, 1)
    errs = {"proj.weight": rel(torch.einsum("bto,bti->oi", g, res), mod.proj.weight.grad[:, :, 0]), "proj.bias": rel(g.sum(dim=(0, 1)), mod.proj.bias.grad)}
    gres = g @ mod.proj.weight.detach()[:, :, 0]
    gq, gk, gc, gd = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(ct), torch.zeros_like(dr)
    for b in range(B):
        for h in range(heads):
            qh, kh, ch, sg, W = saved[b, h]
            gr_ = gres[b, :, h * dh:(h + 1) * dh]                            # (s, c)
            gc[b, :, h * dh:(h + 1) * dh] = W @ gr_                          # dcontent[t, c] = sum_s W[t, s] dres[s, c]
            dW = ch @ gr_.T                                                  # [t, s]
            dd = W * (dW - (W * dW).sum(dim=0, keepdim=True))                # softmax backward over t
            dd.fill_diagonal_(0.0)                                           # masked_fill: no gradient through the diagonal
            gk[b, :, h * dh:(h + 1) * dh] = dd @ qh / math.sqrt(dh)
            gq[b, :, h * dh:(h + 1) * dh] = dd.T @ kh / math.sqrt(dh)
            gd[b, :, h * nd:(h + 1) * nd] = torch.einsum("ts,fts->sf", dd, kern) * sg * (1 - sg) / 2
    gx = g.clone()
    for name, m, gg in (("query", mod.query, gq), ("key", mod.key, gk), ("content", mod.content, gc), ("query_decay", mod.query_decay, gd)):
        errs[name] = max(rel(torch.einsum("bto,bti->oi", gg, xl), m.weight.grad[:, :, 0]), rel(gg.sum(dim=(0, 1)), m.bias.grad))
        gx = gx + gg @ m.weight.detach()[:, :, 0]
    errs["input"] = rel(gx, x.grad.permute(0, 2, 1))
    print(f"local_state backward  max {max(errs.values()):.1e}   " + ", ".join(f"{k} {v:.0e}" for k, v in errs.items()))
    return max(errs.values())


# ----------------------------------------------------------------------------------------------- _spec / _ispec (TA:465-497)
def check_spec_adjoints():
    """Backward of the spectral front / back end in terms of the STFT-family kernels:
      * d/dx of `_spec`  = window / sqrt(N) * adjoint-rfft of the kept frames' gradients, overlap-added into the ONCE-reflected
        signal (the kept frames 2 .. le+1 never touch torch.stft's own centre padding) and folded back over the reflection;
      * d/dZ of `_ispec` = c_k / sqrt(N) * rfft(window * (g / envelope)) on the kept frames and bins, c_0 = 1, c_k = 2:
        an ordinary forward STFT launch with a per-bin factor."""
    from torchaudio.models._hdemucs import _ispectro, _spectro

    torch.manual_seed(4)
    N, hl, B, L = 64, 16, 2, 200                                         # scaled-down nfft 4096 / hop 1024; L not a hop multiple
    le = math.ceil(L / hl)
    pad = hl // 2 * 3
    win = torch.hann_window(N)
    n = torch.arange(N, dtype=torch.float64)
    kk = torch.arange(N // 2, dtype=torch.float64)
    cos, sin = torch.cos(2 * math.pi * kk[:, None] * n[None, :] / N), torch.sin(2 * math.pi * kk[:, None] * n[None, :] / N)   # (k, n)

    # ---- _spec
    x = torch.randn(B, L, requires_grad=True)
    right = pad + le * hl - L
    xp = F.pad(x[:, None], (pad, right), mode="reflect")[:, 0]
    z = _spectro(xp, N, hl)[..., :-1, :][..., 2:2 + le]                  # (B, N/2, le) complex
    zr = torch.view_as_real(z)
    r = torch.randn_like(zr)
    (zr * r).sum().backward()
    Lp = xp.shape[-1]
    gxp = torch.zeros(B, Lp)
    for t in range(le):                                                  # kept frame t covers xp[t hl, t hl + N)
        gfr = (r[:, :, t, 0] @ cos - r[:, :, t, 1] @ sin) * win / math.sqrt(N)
        gxp[:, t * hl:t * hl + N] += gfr
    gx = gxp[:, pad:pad + L].clone()                                     # fold the reflections back: xp[pad - j] = x[j], xp[pad + L - 1 + j] = x[L - 1 - j]
    for j in range(1, pad + 1):
        gx[:, j] += gxp[:, pad - j]
    for j in range(1, right + 1):
        gx[:, L - 1 - j] += gxp[:, pad + L - 1 + j]
    e_spec = rel(gx, x.grad)

    # ---- _ispec
    zin = torch.randn(B, N // 2, le, 2, requires_grad=True)
    zc = torch.view_as_complex(zin)
    zc = F.pad(F.pad(zc, [0, 0, 0, 1]), [2, 2])
    lfull = hl * le + 2 * pad
    y = _ispectro(zc, hl, length=lfull)[..., pad:pad + L]
    ry = torch.randn_like(y)
    (y * ry).sum().backward()
    nfr = le + 4
    env = torch.zeros((nfr - 1) * hl + N)
    for f in range(nfr):
        env[f * hl:f * hl + N] += win * win
    env = env[N // 2:N // 2 + lfull]                                      # centre trim
    gfull = torch.zeros(B, lfull)
    gfull[:, pad:pad + L] = ry
    gq = gfull / env
    gz = torch.zeros(B, N // 2, le, 2)
    ck = torch.full((N // 2,), 2.0)
    ck[0] = 1.0
    for t in range(le):
        f = t + 2                                                        # frame f covers trimmed positions [f hl - N/2, f hl + N/2)
        lo = f * hl - N // 2
        seg = torch.zeros(B, N)
        a, b_ = max(lo, 0), min(lo + N, lfull)
        seg[:, a - lo:b_ - lo] = gq[:, a:b_]
        seg = seg * win
        gz[:, :, t, 0] = ck / math.sqrt(N) * (seg @ cos.T)
        gz[:, :, t, 1] = -ck / math.sqrt(N) * (seg @ sin.T)
    e_ispec = rel(gz, zin.grad)
    print(f"spec backward {e_spec:.1e}   ispec backward {e_ispec:.1e}")
    return max(e_spec, e_ispec)


# ----------------------------------------------------------------------------------------------- per-item normalisation (TA:553-563, 624-631)
def check_item_norm():
    """x_n = (x - mu) / (1e-5 + sigma) on the way in (sigma unbiased, over the whole item) and out = v sigma + mu on the way out:
    mu and sigma are functions of the input, so the input gradient collects three terms -- two fp64 reductions per item:
      d_mu = sum(g_out) - sum(g_n) / (eps + sigma),   d_sigma = sum(g_out v) - sum(g_n (x - mu)) / (eps + sigma)^2,
      dx = g_n / (eps + sigma) + d_mu / n + d_sigma (x - mu) / ((n - 1) sigma)."""
    torch.manual_seed(5)
    B, n = 2, 300
    x = torch.randn(B, n, requires_grad=True)
    v = torch.randn(B, n, requires_grad=True)                                # what the network hands to the de-normalisation
    mu = x.mean(dim=1, keepdim=True)
    sd = x.std(dim=1, keepdim=True)
    xn = (x - mu) / (1e-5 + sd)
    out = v * sd + mu
    gn, go = torch.randn_like(xn), torch.randn_like(out)
    ((xn * gn).sum() + (out * go).sum()).backward()
    xd, vd, mud, sdd = x.detach(), v.detach(), mu.detach(), sd.detach()
    d_mu = go.sum(dim=1, keepdim=True) - gn.sum(dim=1, keepdim=True) / (1e-5 + sdd)
    d_sd = (go * vd).sum(dim=1, keepdim=True) - (gn * (xd - mud)).sum(dim=1, keepdim=True) / (1e-5 + sdd) ** 2
    dx = gn / (1e-5 + sdd) + d_mu / n + d_sd * (xd - mud) / ((n - 1) * sdd)
    e = max(rel(dx, x.grad), rel(go * sdd, v.grad))
    print(f"item normalisation backward {e:.1e}")
    return e


if __name__ == "__main__":
    worst = max(check_encoder(False), check_encoder(True), check_decoder(False, False), check_decoder(True, False), check_decoder(False, True),
                check_freq_encoder(), check_freq_decoder(), check_blstm(), check_local_state(), check_spec_adjoints(), check_item_norm())
    print("worst relative error", f"{worst:.1e}")
    sys.exit(0 if worst < 1e-10 else 1)
