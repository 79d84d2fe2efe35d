"""Generate tests/golden/*.npz by running the UNCHANGED reference modules (build container only).

TEST INFRASTRUCTURE.  Usage:  python -m oracle.make_golden
Each fixture stores the seeds/shapes that regenerate inputs and weights (oracle/weights.py),
a checksum of those weights, and the reference module's output on CPU fp32.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import refshim, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TCN_KW = dict(ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256, kernel_size=7,
              stack_size=10, dilation_growth=2, condition=False, latent_dim=2, norm_type="identity",
              causal=False, estimate_loudness=False)


def _save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in arrs.items()})


def main():
    torch.set_flush_denormal(True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    R = refshim.ref_modules()

    # --- STFT / iSTFT known-answer (torch.stft as TorchSTFT calls it, transforms.py:106-116)
    x = weights.synth_audio(11, 2, 8192)[:, 0]
    win = torch.hann_window(2048)
    Z = torch.stft(x, 2048, 512, window=win, center=True, normalized=False, onesided=True, pad_mode="reflect", return_complex=True)
    y = torch.istft(Z, 2048, 512, window=win, center=True, normalized=False, onesided=True, length=8192)
    _save("stft_kat.npz", seed=11, B=2, T=8192, n_fft=2048, hop=512, Z=torch.view_as_real(Z).numpy(), y=y.numpy())

    # --- Open-Unmix wrapper: sample() and forward() (remfx/models.py:294-304), eval mode
    sd = weights.umx_state(0)
    m = R.models.OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
    m.load_state_dict(sd, strict=True)
    m.eval()
    xa, ta = weights.synth_audio(21, 2, 16384), weights.synth_audio(22, 2, 16384)
    with torch.no_grad():
        out = m.sample(xa)
        loss, out2 = m((xa, ta))
    assert torch.equal(out, out2)
    _save("umx_sample.npz", wseed=0, xseed=21, tseed=22, B=2, T=16384, wsum=weights.checksum(sd),
          out=out.numpy(), loss=float(loss))

    # --- TCN wrapper (remfx/models.py:379-390), cfg/model/tcn.yaml hyper-parameters
    sdt = weights.tcn_state(0)
    tm = R.models.TCNModel(sample_rate=48000, num_bins=1025, **TCN_KW)
    tm.load_state_dict(sdt, strict=True)
    tm.eval()
    xt, tt = weights.synth_audio(31, 1, 16384), weights.synth_audio(32, 1, 16384)
    with torch.no_grad():
        loss, out = tm((xt, tt))
    _save("tcn_forward.npz", wseed=0, xseed=31, tseed=32, B=1, T=16384, wsum=weights.checksum(sdt),
          out=out.numpy(), loss=float(loss))

    tcn_backward_golden(R)
    example_wav_golden(R)
    chain_golden(R)
    cnn14_golden(R)
    tcn_full_size_golden(R)


TCN_BWD = dict(wseed=41, xseed=43, rseed=44, B=2, T=3000, nblocks=3, width=64)


def tcn_backward_inputs():
    """Seeded inputs of the TCN gradient fixture: weights with every PReLU slope set to 1 (a kink-free network, so that
    two fp32 implementations agree to rounding -- tests/test_gpu_tcn_backward.py explains), audio, and the weights r of
    the objective sum(out * r)."""
    c = TCN_BWD
    sd = weights.tcn_state(c["wseed"], nblocks=c["nblocks"], width=c["width"])
    for k in sd:
        if k.endswith("relu.weight"):
            sd[k] = torch.ones_like(sd[k])
    x = weights.synth_audio(c["xseed"], c["B"], c["T"])
    Lout = c["T"] - sum(6 * 2 ** (n % 10) for n in range(c["nblocks"]))
    r = torch.randn(c["B"], 1, Lout, generator=torch.Generator().manual_seed(c["rseed"]))
    return sd, x, r


def reference_tcn_gradients(R, sd, x, r, nblocks, width):
    """Parameter gradients of sum(out * r) through the UNCHANGED reference `TCNModel` (remfx/models.py:370-390 ->
    remfx/tcn.py) under torch autograd -- what `loss.backward()` does to the network in the reference's training step."""
    kw = dict(TCN_KW, nblocks=nblocks, channel_width=width)
    tm = R.models.TCNModel(sample_rate=48000, num_bins=1025, **kw)
    tm.load_state_dict(sd, strict=True)
    tm.zero_grad()
    out = tm.sample(x)
    (out * r).sum().backward()
    return out.detach(), {k: p.grad.detach().clone() for k, p in tm.named_parameters()}


def tcn_backward_golden(R):
    c = TCN_BWD
    sd, x, r = tcn_backward_inputs()
    out, grads = reference_tcn_gradients(R, sd, x, r, c["nblocks"], c["width"])
    _save("tcn_backward.npz", wsum=weights.checksum(sd), out=out.numpy(), **{k: int(v) for k, v in c.items()},
          **{"grad/" + k: g.numpy() for k, g in grads.items()})


EXAMPLE_DECIM = 16   # outputs are stored every 16th sample (rel-RMS on the subset); the input is stored whole
EXAMPLE_TCN_T = 65536


def example_input(pcm16) -> torch.Tensor:
    """(1, 1, 262144) fp32 audio from the 16-bit fixture: the same expression in the generator and in every test."""
    return (torch.from_numpy(pcm16.astype("float32")) / 32767.0).reshape(1, 1, -1)


def example_wav_golden(R):
    """Real audio (SURVEY 8d parity gates: "... + example.wav"): the reference repository's one audio file,
    /root/reference/example.wav (48 kHz mono float32, exactly one 262144-sample chunk, RMS 0.101), quantised to 16-bit PCM so the
    fixture stays small, run through the UNCHANGED reference modules with the seeded weights of the other fixtures:
    Open-Unmix sample, TCN sample (first 65536 samples), Cnn14 probabilities / decisions, and torchaudio's HDemucs (oracle
    of the Demucs wrapper).  Outputs are stored decimated by 16."""
    from scipy.io import wavfile

    from oracle import hdemucs as ohd

    sr, wav = wavfile.read(os.path.join(refshim.REF_ROOT, "example.wav"))
    assert sr == 48000 and wav.ndim == 1 and wav.shape[0] == 262144 and wav.dtype == np.float32
    pcm16 = np.clip(np.round(wav * 32767.0), -32768, 32767).astype(np.int16)
    x = example_input(pcm16)
    D = EXAMPLE_DECIM
    with torch.no_grad():
        sdu = weights.umx_state(0)
        um = R.models.OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
        um.load_state_dict(sdu, strict=True)
        um.eval()
        umx_out = um.sample(x)
        sdt = weights.tcn_state(0)
        tm = R.models.TCNModel(sample_rate=48000, num_bins=1025, **TCN_KW)
        tm.load_state_dict(sdt, strict=True)
        tm.eval()
        tcn_out = tm.sample(x[..., :EXAMPLE_TCN_T])
        sdc = weights.cnn14_state(0)
        cm = R.classifier.Cnn14(num_classes=5, n_fft=2048, hop_length=512, n_mels=128, sample_rate=48000, model_sample_rate=48000,
                                specaugment=True)
        cm.load_state_dict(sdc, strict=True)
        cm.eval()
        probs = torch.hstack(cm(x))
        hd_out = ohd.sample(x, ohd.build(0))
    logits = torch.log(probs.double() / (1 - probs.double())).float()
    _save("example_wav.npz", pcm16=pcm16, decim=D, tcn_T=EXAMPLE_TCN_T,
          umx_wsum=weights.checksum(sdu), tcn_wsum=weights.checksum(sdt), cnn14_wsum=weights.checksum(sdc),
          umx_out=umx_out[0, 0, ::D].numpy(), tcn_out=tcn_out[0, 0, ::D].numpy(), tcn_len=tcn_out.shape[-1],
          hdemucs_out=hd_out[0, 0, ::D].numpy(), probs=probs.numpy(), logits=logits.numpy(), decisions=(probs > 0.5).numpy())


def tcn_full_size_golden(R):
    """VERDICT r1 'parity gaps at BASELINE sizes': the UNCHANGED reference TCNModel (20 blocks, cfg/model/tcn.yaml) on ONE FULL
    262144-sample chunk (config 1 of BASELINE.json) -- example.wav whole, where example_wav.npz stops at 65536 samples.
    5.1 TFLOP on the host: about a minute.  Output (1, 1, 249868) stored every 16th sample."""
    g = np.load(os.path.join(OUT, "example_wav.npz"))
    x = example_input(g["pcm16"])
    with torch.no_grad():
        sdt = weights.tcn_state(0)
        tm = R.models.TCNModel(sample_rate=48000, num_bins=1025, **TCN_KW)
        tm.load_state_dict(sdt, strict=True)
        tm.eval()
        out = tm.sample(x)
    assert out.shape[-1] == 262144 - 12276
    _save("tcn_full_size.npz", decim=EXAMPLE_DECIM, tcn_wsum=weights.checksum(sdt), out=out[0, 0, ::EXAMPLE_DECIM].numpy(),
          out_len=out.shape[-1], out_rms=float(out.double().square().mean().sqrt()))


CHAIN_ORDER = ["RandomPedalboardDistortion", "RandomPedalboardCompressor", "RandomPedalboardReverb", "RandomPedalboardChorus",
               "RandomPedalboardDelay"]  # cfg/exp/remfx_detect.yaml:80-85
CHAIN = dict(T=65536, B=4, xseed=77, yseed=78, member_seed0=50, decim=16)


def reference_chain(R, x, y, member_seed0, use_all=False):
    """The UNCHANGED `remfx.models.RemFXChainInference.forward` (remfx/models.py:52-108) with the unchanged Cnn14 as the
    classifier and five unchanged Open-Unmix wrappers as the effect-specific members (member e = ALL_EFFECTS[i] carries
    weights.umx_state(member_seed0 + i)); members are reached as `self.model[effect].model.sample`, like the Lightning
    `RemFX` modules the reference stores."""
    import types

    names = [e.__name__ for e in R.models.ALL_EFFECTS]
    members = {}
    for i, e in enumerate(names):
        m = R.models.OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
        m.load_state_dict(weights.umx_state(member_seed0 + i), strict=True)
        members[e] = types.SimpleNamespace(model=m.eval())
    cm = R.classifier.Cnn14(num_classes=5, n_fft=2048, hop_length=512, n_mels=128, sample_rate=48000, model_sample_rate=48000,
                            specaugment=True)
    cm.load_state_dict(weights.cnn14_state(0), strict=True)
    cm.eval()
    chain = R.models.RemFXChainInference(members, sample_rate=48000, num_bins=1025, effect_order=list(CHAIN_ORDER), classifier=cm,
                                         use_all_effect_models=use_all)
    with torch.no_grad():
        loss, out = chain((x, y, None, None), 0)
        labels = torch.where(torch.hstack(cm(x)) > 0.5, 1.0, 0.0)
    return loss, out, labels, names


def chain_golden(R):
    c = CHAIN
    x, y = weights.synth_diverse(c["xseed"], c["B"], c["T"]), weights.synth_audio(c["yseed"], c["B"], c["T"])
    loss, out, labels, names = reference_chain(R, x, y, c["member_seed0"])
    assert names == ["RandomPedalboardReverb", "RandomPedalboardChorus", "RandomPedalboardDelay", "RandomPedalboardDistortion",
                     "RandomPedalboardCompressor"]   # remfx/effects.py:699-705: the label order the drop-in hard-codes
    _save("chain_forward.npz", loss=float(loss), labels=labels.numpy(), out=out[:, 0, ::c["decim"]].numpy(), **{k: int(v) for k, v in c.items()})


def cnn14_golden(R, n_chunks: int = 1024, T: int = 262144):
    """Cnn14 (remfx/classifier.py:193-233, eval): logits + decisions of the reference on 1024 seeded diverse chunks
    (generated in 64 batches of 16 with seeds 1000..1063) -- the bit-exact per-effect decision gate of BASELINE.json."""
    sd = weights.cnn14_state(0)
    m = R.classifier.Cnn14(num_classes=5, n_fft=2048, hop_length=512, n_mels=128, sample_rate=48000, model_sample_rate=48000,
                           specaugment=True)
    m.load_state_dict(sd, strict=True)
    m.eval()
    probs = []
    with torch.no_grad():
        for i in range(n_chunks // 16):
            x = weights.synth_diverse(1000 + i, 16, T)
            probs.append(torch.hstack(m(x)))
            if i % 8 == 0:
                print("cnn14 batch", i, flush=True)
    probs = torch.cat(probs)
    logits = torch.log(probs.double() / (1 - probs.double())).float()
    _save("cnn14_decisions.npz", wseed=0, first_xseed=1000, batch=16, n_chunks=n_chunks, T=T, wsum=weights.checksum(sd),
          probs=probs.numpy(), logits=logits.numpy(), decisions=(probs > 0.5).numpy())


if __name__ == "__main__":
    import sys

    one = {"tcn_backward": tcn_backward_golden, "example_wav": example_wav_golden, "chain": chain_golden}
    if len(sys.argv) > 1 and sys.argv[1] in one:   # add one fixture without regenerating the others
        torch.set_flush_denormal(True)
        one[sys.argv[1]](refshim.ref_modules())
    else:
        main()
