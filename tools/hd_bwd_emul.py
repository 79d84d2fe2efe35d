"""Development aid for the Hybrid-Demucs backward: the backward of a time-branch encoder layer and of a
time-branch decoder layer written ONLY in terms of the primitives the CUDA path has (or will have), checked in fp64 against
torch autograd through the torchaudio modules themselves (TA = torchaudio/models/_hdemucs.py).  Runs on the CPU:

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
  * GroupNorm backward needs exactly two sums per (item, group): sum(dy gamma) and sum(dy gamma xhat).
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


if __name__ == "__main__":
    worst = max(check_encoder(False), check_encoder(True), check_decoder(False, False), check_decoder(True, False), check_decoder(False, True))
    print("worst relative error", f"{worst:.1e}")
    sys.exit(0 if worst < 1e-10 else 1)
