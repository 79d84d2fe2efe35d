"""not gpu: the reference's UNCHANGED config tree (cfg/config.yaml, cfg/model/*.yaml, cfg/exp/*.yaml) drives the drop-ins:
compose -> resolve ${...} -> swap `_target_` -> construct (scripts/train.py:9-17, scripts/chain_inference.py:20-37).
Reads the reference's cfg/ from /root/reference or the oracle/_ref copy; skipped when neither is there."""
import os

import pytest

from oracle import refshim
from remfx_b200 import config as rcfg

CFG = os.path.join(refshim.REF_ROOT, "cfg")
pytestmark = pytest.mark.skipif(not os.path.isdir(CFG), reason="no reference cfg tree available")


@pytest.mark.parametrize("name,cls,check", [
    ("umx", "OpenUnmixModel", lambda n: n.n_fft == 2048 and n.hop_length == 512 and n.sample_rate == 48000 and n.alpha == 0.3),
    ("tcn", "TCNModel", lambda n: n.model.nblocks == 20 and n.model.channel_width == 256 and n.model.kernel_size == 7
        and n.model.receptive_field == 12277 and n.num_bins == 1025),
    ("demucs", "DemucsModel", lambda n: n.model.nfft == 4096 and n.model.channels == 48 and n.model.audio_channels == 1
        and n.model.sources == ["mixture"]),
])
def test_model_groups_instantiate_the_dropins(name, cls, check):
    import remfx_b200.models as M
    from remfx_b200.train import RemFX

    cfg = rcfg.compose(CFG, groups={"model": name})
    assert cfg["model"]["_target_"] == "remfx.models.RemFX" and cfg["model"]["sample_rate"] == 48000  # ${sample_rate} resolved
    assert cfg["model"]["lr"] == 1e-4 and isinstance(cfg["model"]["lr"], float)
    mod = rcfg.instantiate(cfg["model"])
    assert isinstance(mod, RemFX) and isinstance(mod.model, getattr(M, cls)) and check(mod.model)
    assert (mod.lr, mod.lr_beta1, mod.lr_beta2, mod.lr_eps, mod.lr_weight_decay) == (1e-4, 0.95, 0.999, 1e-6, 1e-3)


def test_exp_5_5_full_selects_hybrid_demucs_and_the_trainer_settings():
    from remfx_b200.models import DemucsModel

    cfg = rcfg.compose(CFG, exp="5-5_full", overrides={"render_files": False})
    assert cfg["datamodule"]["train_batch_size"] == 16 and cfg["chunk_size"] == 262144 and cfg["render_files"] is False
    tr = cfg["trainer"]
    assert tr["precision"] == 32 and tr["gradient_clip_val"] == 10.0 and tr["max_steps"] == 50000 and tr["devices"] == 1
    mod = rcfg.instantiate(cfg["model"], max_steps=tr["max_steps"], gradient_clip_val=tr["gradient_clip_val"])
    assert isinstance(mod.model, DemucsModel) and mod.max_steps == 50000
    sd_keys = list(mod.state_dict().keys())
    assert sd_keys[0].startswith("model.model.") and len(sd_keys) >= 397  # Lightning checkpoint key layout


def test_exp_remfx_detect_builds_the_chain_members_and_the_classifier():
    from remfx_b200.chain import ALL_EFFECTS, RemFXChainInference
    from remfx_b200.classifier import Cnn14
    from remfx_b200.models import DemucsModel

    cfg = rcfg.compose(CFG, exp="remfx_detect")
    assert cfg["inference_effects_ordering"][0] == "RandomPedalboardDistortion" and set(cfg["ckpts"]) == set(ALL_EFFECTS)
    clf = rcfg.instantiate(cfg["classifier"]["network"])   # scripts/chain_inference.py:30-33 builds FXClassifier(network=Cnn14(...))
    assert isinstance(clf, Cnn14) and clf.num_classes == 5
    models = {}
    for effect, node in cfg["ckpts"].items():             # scripts/chain_inference.py:20-27
        target = node["model"]["network"]["_target_"]
        if target == "remfx.models.DCUNetModel":
            with pytest.raises(NotImplementedError, match="DCUNet"):
                rcfg.instantiate(node["model"])
            continue
        models[effect] = rcfg.instantiate(node["model"])
        assert isinstance(models[effect].model, DemucsModel) and node["ckpt_path"].endswith(".ckpt")
    assert set(models) == {"RandomPedalboardDistortion", "RandomPedalboardCompressor"}   # cfg/exp/remfx_detect.yaml:63-68
    chain = RemFXChainInference(models, sample_rate=cfg["sample_rate"], num_bins=cfg["num_bins"],
                                effect_order=cfg["inference_effects_ordering"], classifier=clf,
                                shuffle_effect_order=cfg["inference_effects_shuffle"],
                                use_all_effect_models=cfg["inference_use_all_effect_models"])
    assert chain.effect_order == cfg["inference_effects_ordering"] and chain.use_all_effect_models is False


def test_interpolation_forms():
    os.environ["RFX_TEST_ROOT"] = "/data"
    cfg = rcfg.resolve({"a": 3, "b": {"c": "${a}", "d": "x-${a}-${b.c}", "e": "${oc.env:RFX_TEST_ROOT}/z", "f": "${oc.env:RFX_NOPE,dflt}"},
                        "g": "${b}", "h": "${now:%Y}"})
    assert cfg["b"] == {"c": 3, "d": "x-3-3", "e": "/data/z", "f": "dflt"} and cfg["g"] == cfg["b"] and len(cfg["h"]) == 4
    with pytest.raises(KeyError):
        rcfg.resolve({"a": "${missing.key}"})
    with pytest.raises(ValueError):
        rcfg.resolve({"a": "${b}", "b": "${a}"})
