"""CPU: the Lightning-surface modules keep the reference's checkpoint key names, LR schedule and dispatch
(no kernels run here; the GPU behaviour is in test_gpu_lit.py)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ref_cpu as O


def _hp(**kw):
    return SimpleNamespace(exp_type=kw.pop("exp_type", "vanilla"), run_max_steps=1000, white_back=True,
                           N_max_objs=1, N_obj_code_length=128, **kw)


def test_state_dict_keys_match_reference_layout(built_lib):
    from aon_b200 import lit
    for exp, kind in (("vanilla", "vanilla"), ("vanilla_autodecoder", "autodecoder")):
        sys_ = lit.build_system(_hp(exp_type=exp))
        ours = {k: tuple(v.shape) for k, v in sys_.state_dict().items()}
        ref = {("model." + k if not k.startswith("code_library.") else k): tuple(v.shape)
               for k, v in O.make_state_dict(kind, 0).items()}
        assert ours == ref           # oracle.make_state_dict mirrors the reference's parameter names (gen_golden pins it)


def test_lr_schedule_matches_reference_formula(built_lib):
    from aon_b200 import lit
    s = lit.LitNeRF(_hp())
    # models/vanilla_nerf/model.py:402-416
    for step in (0, 1, 1250, 2500, 5000, 999, 1000, 5000):
        delay = 0.01 + 0.99 * np.sin(0.5 * np.pi * np.clip(step / 2500, 0, 1))
        t = np.clip(step / 1000, 0, 1)
        want = delay * np.exp(np.log(5e-4) * (1 - t) + np.log(5e-6) * t)
        assert s.learning_rate(step) == pytest.approx(want, rel=1e-12)


def test_unknown_exp_type_rejected(built_lib):
    from aon_b200 import lit
    with pytest.raises(ValueError):
        lit.build_system(_hp(exp_type="vanilla_ae_art"))


def test_latent_fold_equals_concat_formulation(built_lib):
    """train_tc._fold: bias + W[:, cols] @ code equals the reference's broadcast-and-concatenate formulation
    (model_autodecoder.py:186-198): W @ [x, shape, art] + b with x = 0."""
    from aon_b200 import train_tc
    torch.manual_seed(0)
    W, b = torch.randn(128, 163), torch.randn(128)
    shape, art = torch.randn(1, 128), torch.randn(1, 32)
    got = train_tc._fold(b, W, [(3, shape.reshape(-1)), (131, art.reshape(-1))])
    want = torch.nn.functional.linear(torch.cat([torch.zeros(1, 3), shape, art], -1), W, b)[0]
    assert torch.allclose(got, want, atol=1e-5)
    assert train_tc._pad(63, 16) == 64 and train_tc._pad(319, 16) == 320 and train_tc._pad(256, 16) == 256


def test_flat_adam_refuses_cpu_parameters(built_lib):
    """No CPU fallback on the training path either: the optimizer fails loudly on CPU parameters."""
    from aon_b200 import lib, lit
    s = lit.LitNeRF(_hp())
    with pytest.raises(lib.AonError):
        s.configure_optimizers()


def test_ssim_matches_an_independent_evaluation(built_lib):
    """lit ssim_each (interface.py:102-112 computes piqa.SSIM(); piqa is absent, the metric is restated from its definition)
    against an independent float64 numpy / scipy evaluation of the same definition: 11-tap Gaussian (sigma 1.5), valid
    windows, k1 = 0.01, k2 = 0.03, mean over channels and pixels; identical images give exactly 1."""
    import numpy as np
    import torch
    from scipy.ndimage import correlate1d
    from aon_b200 import lit
    g = torch.Generator().manual_seed(0)
    H, W = 37, 52
    gt = torch.rand(H, W, 3, generator=g)
    pred = (gt + 0.1 * torch.randn(H, W, 3, generator=g)).clamp(-0.2, 1.2)       # exercises the clip to [0, 1]
    m = lit._LitCommon()
    got = m.ssim_each([pred, gt], [gt, gt])
    assert abs(got[1].item() - 1.0) < 1e-6

    k = np.exp(-(np.arange(11) - 5.0) ** 2 / (2 * 1.5 ** 2)); k /= k.sum()
    blur = lambda a: correlate1d(correlate1d(a, k, axis=0, mode="constant"), k, axis=1, mode="constant")[5:-5, 5:-5]
    x, y = np.clip(pred.double().numpy(), 0, 1), np.clip(gt.double().numpy(), 0, 1)
    vals = []
    for c in range(3):
        mx, my = blur(x[..., c]), blur(y[..., c])
        sxx, syy, sxy = blur(x[..., c] ** 2) - mx ** 2, blur(y[..., c] ** 2) - my ** 2, blur(x[..., c] * y[..., c]) - mx * my
        cs = (2 * sxy + 0.03 ** 2) / (sxx + syy + 0.03 ** 2)
        vals.append(((2 * mx * my + 0.01 ** 2) / (mx ** 2 + my ** 2 + 0.01 ** 2) * cs))
    want = float(np.mean(np.stack(vals)))
    assert abs(got[0].item() - want) < 1e-5, (got[0].item(), want)
    assert 0.0 < want < 0.999
