"""GPU (B200): in-kernel random draws of the randomized sampling steps (helper.py:126 stratified jitter, helper.py:227
inverse-cdf draws).  Integer work: the Philox draws must be BIT-EQUAL to the numpy restatement (oracle/philox.py, itself
pinned to the Random123 known-answer vectors), and a sampling kernel that generates its draws must produce exactly what the
same kernel produces from the injected tensor of those draws."""
import numpy as np
import pytest
import torch

from oracle import philox as P
from oracle import ref_cpu as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("seed,advance", [(0, 0), (1234567890123456789, 3), (2 ** 64 - 1, 2 ** 33 + 5)])
def test_draws_bit_equal_oracle(built_lib, seed, advance):
    from aon_b200 import lib as L
    rng = L.Rng(seed, torch.device(DEV))
    if advance:
        rng.advance(advance)
    for stream, (rows, cols) in ((0, (131, 65)), (1, (131, 128)), (1, (1, 1)), (0, (7, 3))):
        got = rng.uniform(stream, rows, cols).cpu().numpy()
        want = P.uniform(seed, advance, stream, rows, cols)
        assert got.dtype == np.float32 and np.array_equal(got, want), (stream, rows, cols)
    assert int(rng.offset_dev.item()) == advance


def test_sampling_kernels_with_in_kernel_draws(built_lib):
    """aon_sample_along_rays_rng / aon_sample_pdf_rng == the tensor-fed kernels on the same draws (torch.equal), and == the
    oracle's sample_along_rays / sample_pdf fed with the numpy draws."""
    from aon_b200 import lib as L
    dev = torch.device(DEV)
    R, nc, nf = 517, 65, 128
    rng = L.Rng(99, dev)
    rng.advance(11)
    t_rand, u = rng.uniform(0, R, nc), rng.uniform(1, R, nf)
    t_a = L.sample_along_rays(2.0, 6.0, nc, R, dev, rng=rng)
    t_b = L.sample_along_rays(2.0, 6.0, nc, R, dev, t_rand=t_rand)
    assert torch.equal(t_a, t_b)
    zero = torch.zeros(R, 3)
    t_ref, _ = O.sample_along_rays(zero, zero, nc - 1, 2.0, 6.0, True, torch.from_numpy(P.uniform(99, 11, 0, R, nc)))
    assert torch.equal(t_a.cpu(), t_ref)
    g = torch.Generator().manual_seed(3)
    w = torch.rand(R, nc, generator=g).to(dev) ** 4
    f_a = L.sample_pdf(t_a, w, nf, rng=rng)
    f_b = L.sample_pdf(t_a, w, nf, u=u)
    assert torch.equal(f_a, f_b)
    assert (f_a[:, 1:] >= f_a[:, :-1]).all() and f_a.min() >= 2.0 and f_a.max() <= 6.0
    rng.advance()
    assert not torch.equal(L.sample_along_rays(2.0, 6.0, nc, R, dev, rng=rng), t_a)      # next step, new draws


def test_randomized_training_render_reproducible(built_lib):
    """The randomized training render draws in-kernel: same seed -> same loss, consecutive steps differ, no torch generator
    is consumed."""
    from aon_b200 import nerf
    dev = torch.device(DEV)
    rays = {k: v[:64].to(dev) for k, v in O.sapien_rays(10, 12, seed=4).items()}
    losses = []
    for rep in range(2):
        torch.manual_seed(5)
        state0 = torch.cuda.get_rng_state(dev)
        net = nerf.NeRF().to(dev).train()
        net.rng_seed = 42
        out = [net(rays, True, True, 2.0, 6.0)[1][0].sum().item() for _ in range(2)]
        losses.append(out)
        assert torch.equal(torch.cuda.get_rng_state(dev), state0)
    assert losses[0] == losses[1] and losses[0][0] != losses[0][1], losses
