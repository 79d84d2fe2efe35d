"""-m gpu: rendered chunk directory -> native batch ingest (pinned) -> Open-Unmix device pipeline, vs sample() on the same audio
and the loss against the ingested targets (the two ends of SURVEY 8(f) row N2 around the hot path)."""
import numpy as np
import pytest
import torch

from oracle import loss as oloss
from oracle import umx as oumx
from oracle import weights
from tests.test_ingest_cpu import _render_dir
from tests.util import relrms

pytestmark = pytest.mark.gpu


def test_ingest_feeds_pipeline(tmp_path):
    from remfx_b200.ingest import BatchIngest, EffectChunkReader
    from remfx_b200.losses import remfx_loss
    from remfx_b200.models import OpenUnmixModel

    T, B = 16384, 4
    items = _render_dir(tmp_path, 3 * B, T, seed=9)
    sd = weights.umx_state(2)
    m = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    pipe = m.pipeline("cuda:0")
    ing = BatchIngest(EffectChunkReader(str(tmp_path)), batch_size=B, chunk_size=T, threads=4, n_buffers=pipe.depth + 1, sample_rate=48000)
    outs, seqs, targets = [], [], []
    for xb, yb, dry, wet in ing:
        assert xb.is_pinned() and yb.is_pinned()
        out = torch.empty(B, 1, T, device="cuda")
        seqs.append(pipe.push(xb, out))
        outs.append(out)
        targets.append(yb.clone())
    pipe.flush()
    for s in seqs:
        pipe.wait(s)
    assert len(outs) == 3
    for bi, out in enumerate(outs):
        x = torch.from_numpy(np.stack([items[bi * B + j][0] for j in range(B)])).unsqueeze(1)
        ref = oumx.sample(x, sd)
        assert relrms(out.cpu(), ref) < 1e-4
        loss = remfx_loss(out, targets[bi].cuda())
        rloss = oloss.remfx_loss(ref, targets[bi])
        assert abs(float(loss) - float(rloss)) < 1e-3 * abs(float(rloss))
