"""Quick check of the tcgen05 recurrence (both instantiations) and the mma.sync kernel against the explicit-loop oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import umx as oumx  # noqa: E402
from remfx_b200 import ops  # noqa: E402


def case(B, F, impl, slots=0):
    H, I = 256, 512
    g = torch.Generator().manual_seed(B * 100 + F)
    k = H ** -0.5
    st = {}
    for sfx in ("", "_reverse"):
        st[f"l.weight_ih_l0{sfx}"] = (torch.rand(4 * H, I, generator=g) * 2 - 1) * k
        st[f"l.weight_hh_l0{sfx}"] = (torch.rand(4 * H, H, generator=g) * 2 - 1) * k
        st[f"l.bias_ih_l0{sfx}"] = (torch.rand(4 * H, generator=g) * 2 - 1) * k
        st[f"l.bias_hh_l0{sfx}"] = (torch.rand(4 * H, generator=g) * 2 - 1) * k
    x = torch.randn(F, B, I, generator=g)
    ref = oumx.lstm_explicit(x, st, "l", layers=1)
    G = torch.cat([x @ st[f"l.weight_ih_l0{s}"].t() + st[f"l.bias_ih_l0{s}"] + st[f"l.bias_hh_l0{s}"] for s in ("", "_reverse")], -1)
    G = G.permute(1, 0, 2).reshape(B * F, 8 * H).contiguous()
    Whh = torch.stack([st["l.weight_hh_l0"], st["l.weight_hh_l0_reverse"]])
    out = ops.lstm_layer(G.cuda(), Whh.cuda(), B, F, impl=impl, slots=slots)
    torch.cuda.synchronize()
    out = out.view(B, F, 2 * H).permute(1, 0, 2).cpu()
    return float((out.double() - ref.double()).norm() / ref.double().norm())


if __name__ == "__main__":
    for B, F in ((3, 4), (16, 20), (21, 33)):
        print(f"B={B} F={F} rel-RMS vs oracle: tc<16> {case(B, F, 'tc'):.3e}  tc<32> {case(B, F, 'tc', slots=32):.3e}  mma {case(B, F, 'mma'):.3e}", flush=True)
