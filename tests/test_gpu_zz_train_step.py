"""-m gpu: the training-step mirror (remfx_b200.train.RemFX.fit_step, rows L4 / L5) on the TCN, against the same step
done by torch on the CPU: oracle network under autograd + oracle loss + clip_grad_norm_(10) + torch.optim.AdamW with the
reference's hyper-parameters + MultiStepLR (remfx/models.py:185-256, cfg/config.yaml:110-120).

The file sorts last on purpose: it chains every training component (forward_train, loss forward/backward, TCN backward,
all-reduce-less FusedAdamW, the handle re-sync after the update), each of which has its own tighter test.
lr = 1e-5 (a ctor argument of the reference module) keeps three steps in the regime where the loss falls monotonically.
Tolerances: first loss 1e-4 relative (pure forward), later losses 3e-3 (PReLU-kink sign flips perturb single gradient
elements, and AdamW's normalised update turns a flipped tiny gradient into a 2 lr parameter difference -- see
test_gpu_tcn_backward.py; with 1e-6 .. 1e-5 relative noise injected into the CPU oracle alone the later losses move by
either < 1e-4 or ~4.6e-4, the second when the update of one near-zero-gradient scalar such as output.bias flips), metrics 1e-3 relative / 1e-2 dB.  Direction of the total parameter change: cosine > 0.9 -- AdamW's first updates are
lr * sign(g), so the ~2 % of gradient elements smaller than the kink noise flip their update; emulating 3e-6 relative
noise on the CPU oracle alone gives cosine 0.957 against its own noise-free run, with losses within 5e-4.
"""
import pytest
import torch

from oracle import loss as oloss
from oracle import stft as ostft
from oracle import tcn as otcn
from oracle import weights

pytestmark = pytest.mark.gpu

KW = dict(ninputs=1, noutputs=1, nblocks=4, channel_growth=0, channel_width=256, kernel_size=7, stack_size=10,
          dilation_growth=2, condition=False, latent_dim=2, norm_type="identity", causal=False, estimate_loudness=False)
HP = dict(lr=1e-5, lr_beta1=0.95, lr_beta2=0.999, lr_eps=1e-6, lr_weight_decay=1e-3)
STEPS, MAX_STEPS = 3, 50


def _reference_steps(sd, x, y):
    st = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    params = list(st.values())
    opt = torch.optim.AdamW(params, lr=HP["lr"], betas=(HP["lr_beta1"], HP["lr_beta2"]), eps=HP["lr_eps"], weight_decay=HP["lr_weight_decay"])
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [0.8 * MAX_STEPS, 0.95 * MAX_STEPS], gamma=0.1)
    losses, metrics = [], {}
    for _ in range(STEPS):
        opt.zero_grad()
        loss, out = otcn.forward((x, y), st)
        tgt = ostft.causal_crop(y, out.shape[-1])
        with torch.no_grad():
            metrics = {"train_SISDR": -oloss.sisdr_loss(out, tgt), "Input_SISDR": -oloss.sisdr_loss(x, y),
                       "train_STFT": oloss.mrstft(out, tgt), "Input_STFT": oloss.mrstft(x, y)}
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 10.0)
        opt.step()
        sched.step()
        losses.append(float(loss.detach()))
    return losses, {k: float(v) for k, v in metrics.items()}, {k: v.detach() for k, v in st.items()}


def test_fit_step_matches_torch_training_step():
    from remfx_b200.models import TCNModel
    from remfx_b200.train import RemFX

    sd = weights.tcn_state(21, nblocks=KW["nblocks"])
    x = weights.synth_audio(400, 2, 4000)
    y = weights.synth_audio(401, 2, 4000)
    ref_losses, ref_metrics, ref_params = _reference_steps(sd, x, y)

    net = TCNModel(sample_rate=48000, num_bins=1025, **KW)
    net.load_state_dict(sd, strict=True)
    mod = RemFX(sample_rate=48000, network=net.cuda(), max_steps=MAX_STEPS, **HP)
    batch = (x.cuda(), y.cuda(), None, None)
    losses = [float(mod.fit_step(batch, i)) for i in range(STEPS)]
    assert mod.global_step == STEPS
    assert abs(losses[0] - ref_losses[0]) < 1e-4 * abs(ref_losses[0]), (losses, ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 3e-3 * abs(b), (losses, ref_losses)
    assert losses[-1] < losses[0]
    for k, v in ref_metrics.items():
        got = float(mod.logged[k])
        assert abs(got - v) < max(1e-2 if "SISDR" in k else 0.0, 1e-3 * abs(v)), (k, got, v)
    # total parameter movement: same direction, same size
    num = den_a = den_b = 0.0
    for k, p in net.model.state_dict().items():
        da = (p.detach().cpu() - sd["model." + k]).double().flatten()
        db = (ref_params["model." + k] - sd["model." + k]).double().flatten()
        num += float(da @ db); den_a += float(da @ da); den_b += float(db @ db)
    cos = num / (den_a ** 0.5 * den_b ** 0.5)
    assert cos > 0.9 and 0.9 < (den_a / den_b) ** 0.5 < 1.1, (cos, den_a, den_b)


def test_eval_steps_do_not_build_a_graph_under_no_grad():
    from remfx_b200.models import TCNModel
    from remfx_b200.train import RemFX

    sd = weights.tcn_state(22, nblocks=3)
    net = TCNModel(sample_rate=48000, num_bins=1025, **dict(KW, nblocks=3))
    net.load_state_dict(sd, strict=True)
    mod = RemFX(sample_rate=48000, network=net.cuda(), max_steps=MAX_STEPS, **HP)
    x = weights.synth_audio(410, 1, 3000).cuda()
    with torch.no_grad():
        loss = mod.validation_step((x, x, None, None), 0)
    assert not loss.requires_grad
    assert set(mod.logged) == {"valid_loss", "valid_SISDR", "valid_STFT", "Input_SISDR", "Input_STFT"}
