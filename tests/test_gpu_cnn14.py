"""-m gpu: Cnn14 classifier drop-in vs the oracle and the reference golden decisions.

BASELINE.json gate: the classifier's per-effect decisions (probability > 0.5, remfx/models.py:61-64) must match
the reference bit-exactly.  tests/golden/cnn14_decisions.npz holds the UNCHANGED reference's logits and decisions
on 1024 seeded, spectrally diverse chunks with conditioned weights (logits spread over +-3, several within 0.05
of the threshold); the test also reports how many logits sit within 1e-3 of 0 (where any reassociation could flip)."""
import os

import numpy as np
import pytest
import torch

from oracle import cnn14 as ocnn
from oracle import weights
from tests.util import golden, relrms

pytestmark = pytest.mark.gpu


def _model(sd):
    from remfx_b200.classifier import Cnn14

    m = Cnn14(num_classes=5, sample_rate=48000, model_sample_rate=48000, n_fft=2048, hop_length=512, n_mels=128, specaugment=True)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def test_forward_matches_oracle():
    sd = weights.cnn14_state(0)
    x = weights.synth_diverse(5, 4, 262144)
    m = _model(sd)
    out = m(x.cuda())
    assert isinstance(out, list) and len(out) == 5 and out[0].shape == (4, 1)
    probs, logits = m.probs_and_logits(x.cuda())
    ref = ocnn.logits(x, sd)
    err = (logits.cpu() - ref).abs().max().item()
    print("cnn14 max |logit - oracle| =", err)
    assert err < 2e-3, err
    assert relrms(torch.hstack(out), torch.hstack(ocnn.forward(x, sd))) < 1e-4


def test_decisions_match_reference_golden():
    g = golden("cnn14_decisions.npz")
    n = int(os.environ.get("RFX_CNN_CHUNKS", int(g["n_chunks"])))
    bs, T = int(g["batch"]), int(g["T"])
    sd = weights.cnn14_state(int(g["wseed"]))
    assert abs(weights.checksum(sd) - float(g["wsum"])) < 1e-6 * abs(float(g["wsum"]))
    m = _model(sd)
    lg = []
    for i in range(n // bs):
        x = weights.synth_diverse(int(g["first_xseed"]) + i, bs, T)
        lg.append(m.probs_and_logits(x.cuda())[1].cpu())
    lg = torch.cat(lg).numpy()
    ref_lg, ref_dec = g["logits"][: len(lg)], g["decisions"][: len(lg)]
    dec = lg > 0
    near = int((np.abs(ref_lg) < 1e-3).sum())
    mism = int((dec != ref_dec).sum())
    print(f"cnn14 decisions: {dec.size} logits on {len(lg)} chunks, {mism} mismatches, {near} reference logits within 1e-3 of 0, "
          f"max |dlogit| {np.abs(lg - ref_lg).max():.2e}, min |ref logit| {np.abs(ref_lg).min():.2e}")
    assert mism == 0


def test_bad_inputs():
    sd = weights.cnn14_state(0)
    m = _model(sd)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 65536))
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 1, 65536, device="cuda"), train=True)
