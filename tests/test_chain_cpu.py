"""`not gpu`: host logic of the cascade drop-in (remfx_b200.chain.RemFXChainInference) with CPU stand-ins for the member networks,
the classifier and the loss kernel.  The drop-in gathers the items that need effect E into one batch per effect model; the
reference walks the batch item by item (remfx/models.py:93-104).  Both must give the same result for every label pattern,
order and option -- checked against the item-by-item oracle always and against the UNCHANGED reference class when
/root/reference is present."""
import itertools
import types

import pytest
import torch

from oracle import chain as ochain
from oracle import loss as oloss
from oracle import refshim
from remfx_b200 import chain as C

ORDER = ["RandomPedalboardDistortion", "RandomPedalboardCompressor", "RandomPedalboardReverb", "RandomPedalboardChorus",
         "RandomPedalboardDelay"]


class _Member:
    """Non-commuting per-item stand-ins: applying them in a different order or to the wrong items changes the result."""

    def __init__(self, k):
        self.k = k

    def sample(self, x):
        return torch.tanh(x * (1.0 + 0.3 * self.k)) + 0.05 * self.k * x.mean(dim=-1, keepdim=True)


def _members():
    return {e: _Member(i + 1) for i, e in enumerate(ochain.ALL_EFFECTS)}


def _probs(labels):
    p = 0.1 + 0.8 * labels
    return lambda x: [p[:, j:j + 1] for j in range(5)]


@pytest.fixture()
def cpu_loss(monkeypatch):
    monkeypatch.setattr(C, "remfx_loss_with_terms", lambda out, y: (oloss.remfx_loss(out, y), None))
    monkeypatch.setattr(C, "sisdr_loss", oloss.sisdr_loss)
    monkeypatch.setattr(C, "mrstft_loss", oloss.mrstft)


def _data(B=6, T=4096):
    g = torch.Generator().manual_seed(3)
    return 0.1 * torch.randn(B, 1, T, generator=g), 0.1 * torch.randn(B, 1, T, generator=g)


LABELS = torch.tensor([[0., 0, 0, 0, 0], [1, 1, 1, 1, 1], [0, 0, 0, 1, 0], [1, 0, 0, 1, 1], [0, 1, 1, 0, 0], [1, 0, 1, 0, 1]])


@pytest.mark.parametrize("use_all,with_classifier", list(itertools.product([False, True], [False, True])))
def test_grouped_cascade_equals_item_by_item(cpu_loss, use_all, with_classifier):
    x, y = _data()
    mem = _members()
    clf = _probs(LABELS) if with_classifier else None
    chain = C.RemFXChainInference(mem, 48000, 1025, ORDER, classifier=clf, use_all_effect_models=use_all)
    given = None if with_classifier else LABELS
    loss, out = chain((x, y, None, given), 0)
    rloss, rout, rlabels = ochain.forward(x, y, given, {e: m.sample for e, m in mem.items()}, ORDER,
                                          classify=(lambda z: torch.hstack(clf(z))) if clf else None, use_all=use_all)
    assert torch.allclose(out, rout, rtol=0, atol=1e-7) and float(loss) == pytest.approx(float(rloss), rel=1e-6)
    assert torch.equal(chain.last_labels, rlabels)
    if not use_all:
        assert torch.equal(out[0], x[0])            # nothing detected: the item passes through untouched
    assert not torch.equal(out[1], x[1])


def test_explicit_order_argument_and_unknown_names(cpu_loss):
    x, y = _data(3)
    mem = _members()
    chain = C.RemFXChainInference(mem, 48000, 1025, ORDER, classifier=None)
    order = ["RandomPedalboardDelay", "NotAnEffect", "RandomPedalboardReverb"]
    labels = torch.tensor([[1., 0, 1, 0, 0], [1, 0, 0, 0, 0], [0, 0, 1, 1, 0]])
    _, out = chain((x, y, None, labels), 0, order=order)
    _, rout, _ = ochain.forward(x, y, labels, {e: m.sample for e, m in mem.items()}, order)
    assert torch.allclose(out, rout, rtol=0, atol=1e-7)
    assert torch.allclose(out[2], mem["RandomPedalboardDelay"].sample(x[2:3])[0])   # Distortion is detected but not in `order`


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("use_all", [False, True])
def test_against_the_unchanged_reference_class(cpu_loss, use_all):
    R = refshim.ref_modules()
    x, y = _data()
    mem = _members()
    ref = R.models.RemFXChainInference({e: types.SimpleNamespace(model=m) for e, m in mem.items()}, sample_rate=48000, num_bins=1025,
                                       effect_order=list(ORDER), classifier=_probs(LABELS), use_all_effect_models=use_all)
    with torch.no_grad():
        rloss, rout = ref((x, y, None, None), 0)
    mine = C.RemFXChainInference(mem, 48000, 1025, ORDER, classifier=_probs(LABELS), use_all_effect_models=use_all)
    loss, out = mine((x, y, None, None), 0)
    assert torch.allclose(out, rout, rtol=0, atol=1e-7) and float(loss) == pytest.approx(float(rloss), rel=1e-6)
    # members stored the reference's way (Lightning modules whose `.model` is the network) are accepted as well
    wrapped = C.RemFXChainInference({e: types.SimpleNamespace(model=m) for e, m in mem.items()}, 48000, 1025, ORDER, classifier=_probs(LABELS),
                                    use_all_effect_models=use_all)
    assert torch.equal(wrapped((x, y, None, None), 0)[1], out)


def test_test_step_metric_names_and_shuffle(cpu_loss):
    x, y = _data(2)
    chain = C.RemFXChainInference(_members(), 48000, 1025, list(ORDER), classifier=None, shuffle_effect_order=True)
    loss, metrics = chain.test_step((x, y, None, LABELS[:2]), 0)
    assert set(metrics) == {"test_loss", "test_SISDR", "Input_SISDR", "test_STFT", "Input_STFT"}
    assert sorted(chain.effect_order) == sorted(ORDER)                              # shuffled in place, like models.py:112-114
    assert float(metrics["Input_SISDR"]) == pytest.approx(-float(oloss.sisdr_loss(x, y)), rel=1e-6)
