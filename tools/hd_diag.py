"""Layer-by-layer HDemucs parity diagnostics (development aid): GPU taps vs torchaudio hooks."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import hdemucs as ohd, weights  # noqa: E402
from remfx_b200.models import DemucsModel  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
over = dict(dconv_lstm=6, dconv_attn=6) if (len(sys.argv) > 2 and sys.argv[2] == "nolstm") else {}
ref = ohd.build(0, **over)
kw = dict(ohd.KW); kw.update(over)
m = DemucsModel(sample_rate=48000, **kw)
m.model.load_state_dict(ref.state_dict(), strict=True)
m = m.cuda().eval()
x = weights.synth_audio(3, 2, T)


def rr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


mods = dict(ref.named_modules())
names = [n for n in mods if n and (n.count(".") <= 2 or ".dconv.layers." in n and n.count(".") <= 5)]
rt = ohd.taps(x, ref, names)
out = m.sample(x.cuda(), taps=True)

# reference spec normalisation
with torch.no_grad():
    z = ref._spec(x)
    mag = ref._magnitude(z)
    mean = mag.mean(dim=(1, 2, 3), keepdim=True); std = mag.std(dim=(1, 2, 3), keepdim=True)
    xn = (mag - mean) / (1e-5 + std)




def cmp(gname, r, freq):
    try:
        g = m.tap(gname)
    except Exception as e:  # noqa: BLE001
        print(f"{gname:40s} (no tap: {e})")
        return
    gg = g.permute(0, 3, 2, 1) if freq else g[:, 0].permute(0, 2, 1)
    if gg.shape != r.shape:
        if gg.dim() == r.dim() and gg.shape[2] > r.shape[2]:
            d = (gg.shape[2] - r.shape[2]) // 2
            gg = gg[:, :, d:d + r.shape[2]]
        if gg.shape[1] > r.shape[1]:
            gg = gg[:, : r.shape[1]]
    if gg.shape != r.shape:
        print(f"{gname:40s} SHAPE gpu {tuple(gg.shape)} ref {tuple(r.shape)}")
        return
    print(f"{gname:40s} rel-RMS {rr(gg, r):.3e}   |ref| {float(r.abs().mean()):.3e}")


for i in range(6):
    fe = f"freq_encoder.{i}"
    isfreq = rt[fe].dim() == 4
    conv = rt[fe + ".conv"]
    if i < 4:
        cmp(fe + ".act1", F.gelu(conv), True)
    for d in range(2):
        k = f"{fe}.dconv.layers.{d}.2"
        if k in rt:
            r = rt[k]  # (B*Fr, h, T) for freq layers
            if isfreq:
                Bq = x.shape[0]
                r = r.view(Bq, -1, r.shape[1], r.shape[2]).permute(0, 2, 1, 3)  # (B, h, Fr, T)
            cmp(k, r, isfreq)
    cmp(fe, rt[fe], isfreq)
    te = f"time_encoder.{i}"
    if te in rt and i < 4:
        cmp(te, rt[te], False)
for i in range(6):
    fd = f"freq_decoder.{i}"
    if fd in rt and i < 5:
        cmp(fd, rt[fd], rt[fd].dim() == 4)
    td = f"time_decoder.{i}"
    if td in rt and i < 4:
        cmp(td, rt[td], False)
print("output", rr(out, rt["__output__"]))
