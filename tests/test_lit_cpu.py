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
