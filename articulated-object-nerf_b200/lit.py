"""Lightning-surface modules of the reference, re-hosted on the fused kernels.

* ``LitNeRF``             <- models/vanilla_nerf/model.py:202-507
* ``LitNeRF_AutoDecoder`` <- models/vanilla_nerf/model_autodecoder.py:340-771

Same constructor arguments, hook names and signatures (PL 1.5 ``optimizer_step`` 8-argument form),
same render dict keys (``comp_rgb / acc / depth``; test: ``target / instance_mask / rgb``) and the same
``state_dict`` key names (``model.coarse_mlp.pts_linears.0.weight`` ... / ``code_library.*``), so a
reference checkpoint loads unchanged.  pytorch-lightning is not installed in this image: ``Trainer``
below is the minimal loop that drives those hooks (one process per GPU; gradients averaged with one
flat all-reduce, dist.py); the classes stay plain ``nn.Module`` and also work under real PL 1.5.2.

Differences from the reference, all result-preserving:
* the Python chunk loop of ``render_rays`` / ``render_rays_test`` (model.py:295-348) is gone: the fused
  kernel never materialises a [rays x samples x features] tensor, so a whole image is one call
  (``hparams.chunk`` is accepted and ignored; per-ray arithmetic does not depend on the chunking);
* wandb / image-grid logging is replaced by a ``logged`` dict (no network in the image).
"""
from __future__ import annotations

import json
import math
import os
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import dist as D
from . import lib as L
from .nerf import CodeLibraryArticulated, NeRF, NeRF_AE_Art, img2mse, mse2psnr

Tensor = torch.Tensor


def _hp(hparams) -> SimpleNamespace:
    d = dict(vars(hparams)) if not isinstance(hparams, dict) else dict(hparams)
    d.setdefault("chunk", 16 * 240)          # opt.py:103
    d.setdefault("run_max_steps", 100000)
    d.setdefault("white_back", True)
    d.setdefault("img_wh", (640, 480))
    d.setdefault("N_max_objs", 1)
    d.setdefault("N_obj_code_length", 128)
    return SimpleNamespace(**d)


class _Logged(dict):
    """name -> float, converted lazily: ``log()`` stores the device scalar as it is (no host synchronisation inside
    training_step -- three ``float()`` calls per step would stall the launch queue three times); reading an entry converts."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if torch.is_tensor(v):
            v = float(v)
            dict.__setitem__(self, k, v)
        return v

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in list(self.keys())]

    def values(self):
        return [self[k] for k in list(self.keys())]


class LitModel(nn.Module):
    """models/interface.py:22-62 -- the parts the render path touches (logging + PSNR)."""

    def __init__(self):
        super().__init__()
        self.logged: Dict[str, float] = _Logged()
        self.trainer = SimpleNamespace(global_step=0, is_global_zero=True)

    def log(self, name, value, **_):
        self.logged[name] = value.detach() if torch.is_tensor(value) else float(value)

    @torch.no_grad()
    def psnr_legacy(self, pred: Tensor, gt: Tensor) -> Tensor:
        """interface.py:72-74: -10 log10(mean((pred - gt)^2)), NOT clipped (the validation logs use this one; the
        auto-decoder's padded sigmoid leaves rgb in [-0.001, 1.001])."""
        return -10.0 * torch.log10(torch.mean((pred - gt) ** 2))

    @torch.no_grad()
    def psnr_clipped(self, pred: Tensor, gt: Tensor) -> Tensor:
        """interface.py:54-62 (psnr_each): clip both images to [0,1] first, then -10 log10(mse)."""
        mse = torch.mean((torch.clip(pred, 0, 1) - torch.clip(gt, 0, 1)) ** 2)
        return -10.0 * torch.log(mse) / np.log(10)


def to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def store_image(dirpath, rgbs, name):
    """models/utils.py:21-27: {dirpath}/{name}{NNN}.jpg"""
    from PIL import Image
    for i, rgb in enumerate(rgbs):
        Image.fromarray(to8b(rgb.detach().cpu().numpy())).save(os.path.join(dirpath, "%s%s.jpg" % (name, str(i).zfill(3))))


def write_stats(fpath, *stats):
    """models/utils.py:62-73"""
    d = {st["name"]: {k: float(w) for k, w in st.items() if k not in ("name", "scene_wise")} for st in stats}
    with open(fpath, "w") as fp:
        json.dump(d, fp, indent=4, sort_keys=True)


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(betas, eps, no weight decay) semantics (model.py:386-389) on ONE flat fp32 buffer:
    the parameters and their .grad become views of two contiguous device buffers, so the gradient all-reduce
    (dist.allreduce_mean_) needs no gather/scatter copies and the update is one aon_adam_step launch instead of
    a per-tensor loop.  ``param_groups[0]['lr']`` is honoured every step (optimizer_step writes the schedule there)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        params = [p for p in params if p.requires_grad]
        if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
            raise L.AonError("FlatAdam needs CUDA float32 parameters (move the module to the GPU first): no CPU fallback")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        ps = self.param_groups[0]["params"]
        n, dev = sum(p.numel() for p in ps), ps[0].device
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.steps, self.grad_scale = 0, 1.0
        self.scalars_dev = None           # set by GraphedStep: the Adam scalars come from device memory (graph-capturable step)
        off = 0
        with torch.no_grad():
            for p in ps:
                k = p.numel()
                self.flat[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                off += k

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grad.zero_()
        off = 0
        for p in self.param_groups[0]["params"]:       # re-attach if someone dropped the view (p.grad = None)
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + k].view(p.shape)
            off += k

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        self.steps += 1
        if self.scalars_dev is not None:
            L.adam_step_dev(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.scalars_dev)
        else:
            L.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1],
                        g["eps"], self.steps, self.grad_scale)
        for p in g["params"]:                           # the kernel wrote through raw pointers: tell autograd / the
            torch.autograd.graph.increment_version(p)   # packed-weight cache that the values changed
        return loss

    def state_dict(self):
        return {"steps": self.steps, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])


class GradSync:
    """DDP-style gradient averaging that overlaps with the backward pass (run.py:109-111: PL's DDPPlugin buckets do the same).
    The parameters' .grad are views of ONE flat buffer (FlatAdam); the buffer is cut into contiguous groups (fine MLP, coarse
    MLP, code tables).  A post-accumulate hook counts the gradients of each group as autograd writes them; when a group is
    complete its slice is all-reduced asynchronously (NCCL runs on its own stream), so the fine MLP's 2.4 MB travel while the
    coarse MLP's backward is still computing.  ``finish()`` launches whatever is left, waits, and the optimizer applies the
    1 / world factor (``grad_scale``) -- sums are all-reduced, the mean is taken in the Adam kernel."""

    def __init__(self, named_params, flat_grad: Tensor, group_of=None):
        import torch.distributed as dist
        self.dist, self.flat = dist, flat_grad
        self.world = D.world()[1]
        self.groups, self.works = [], []
        if self.world == 1:
            return
        group_of = group_of or (lambda name: name.split(".")[1] if name.startswith("model.") and name.count(".") > 1 else name.split(".")[0])
        base, esz = flat_grad.data_ptr(), flat_grad.element_size()
        by = {}
        for name, p in named_params:
            if not p.requires_grad or p.grad is None:
                continue
            off = (p.grad.data_ptr() - base) // esz
            if not (0 <= off and off + p.numel() <= flat_grad.numel()):
                raise L.AonError("GradSync: %s.grad is not a view of the flat gradient buffer" % name)
            g = by.setdefault(group_of(name), {"lo": off, "hi": off + p.numel(), "n": 0, "seen": 0, "sent": False})
            g["lo"], g["hi"], g["n"] = min(g["lo"], off), max(g["hi"], off + p.numel()), g["n"] + 1
            p.register_post_accumulate_grad_hook(lambda _p, g=g: self._on_grad(g))
        self.groups = sorted(by.values(), key=lambda g: g["lo"])
        for a, b in zip(self.groups, self.groups[1:]):
            if a["hi"] > b["lo"]:
                raise L.AonError("GradSync: parameter groups are not contiguous in the flat buffer")

    def _send(self, g):
        g["sent"] = True
        self.works.append(self.dist.all_reduce(self.flat[g["lo"]:g["hi"]], op=self.dist.ReduceOp.SUM, async_op=True))

    def _on_grad(self, g):
        g["seen"] += 1
        if g["seen"] == g["n"] and not g["sent"]:
            self._send(g)

    def start(self):
        for g in self.groups:
            g["seen"], g["sent"] = 0, False
        self.works = []

    def finish(self) -> float:
        """-> the factor the optimizer must apply to the summed gradients"""
        if self.world == 1:
            return 1.0
        for g in self.groups:
            if not g["sent"]:          # a group whose hooks did not all fire (a parameter without gradient this step)
                self._send(g)
        for w in self.works:
            w.wait()
        return 1.0 / self.world


class GraphedStep:
    """One training step -- zero_grad, training_step, backward, (gradient all-reduces), optimizer_step -- captured ONCE in a
    CUDA graph and replayed: a vanilla step is ~300 kernel launches whose host-side enqueue (7.4 ms) is almost as long as
    their execution; a replay costs microseconds of host time.  Shapes must not change between steps (the reference's
    fixed-size ray batches, model.py:421-428 / sapien_multi.py:235).  The batch is copied into static buffers; the learning
    rate / bias corrections reach the Adam kernel through a 7-float device buffer refreshed before every replay
    (aon_adam_step_dev), because a graph bakes kernel arguments in.  The stratified and inverse-cdf draws are generated inside
    the sampling kernels from a step counter in device memory that the captured step advances itself (lib.Rng), so they differ
    from replay to replay as they must -- and are the same draws an eager step would have made."""

    RING = 64

    def __init__(self, system, opt, batch, sync: "GradSync" = None, warmup: int = 3):
        self.system, self.opt, self.sync = system, opt, sync
        self.static = {k: (v.clone() if torch.is_tensor(v) and v.is_cuda else v) for k, v in batch.items()}
        dev = opt.flat.device
        # the host runs many replays ahead of the device, and an asynchronous copy reads its pinned source when it EXECUTES:
        # every step gets its own slot of a ring, and a slot is reused only after the copy that read it has completed
        self.sc_ring = torch.zeros(self.RING, 7, dtype=torch.float32).pin_memory()
        self.sc_events = [None] * self.RING
        self.sc_slot = 0
        self.sc_dev = torch.zeros(7, dtype=torch.float32, device=dev)
        opt.scalars_dev = self.sc_dev
        self.graph = None
        self.loss = None
        # DRY RUN: warm-up steps on a side stream (allocator pools, lazily created handles, tensor-map cache) and the capture
        # itself must not count as training -- parameters, Adam moments, step counters and the generator state are restored
        # afterwards, so the first replay is the first real step on this batch
        rng = system.model.rng(dev) if hasattr(system.model, "rng") else None        # in-kernel Philox step counter (device memory)
        saved = (opt.flat.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.steps, system.trainer.global_step,
                 torch.cuda.get_rng_state(dev), None if rng is None else rng.offset_dev.clone())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._refresh_scalars()
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        self._refresh_scalars()
        n0 = L.launch_count()
        try:
            with torch.cuda.graph(g):
                self._body()
            self.launches = L.launch_count() - n0  # kernels of THIS library inside one replay (ATen's come on top)
            self.graph = g
        finally:
            # also after a failed capture (e.g. the caller keeps an autograd graph alive whose AccumulateGrad nodes belong to
            # another stream): the dry run must leave no trace, and the caller falls back to eager steps
            torch.cuda.synchronize()
            with torch.no_grad():
                opt.flat.copy_(saved[0]); opt.exp_avg.copy_(saved[1]); opt.exp_avg_sq.copy_(saved[2])
            opt.steps, system.trainer.global_step = saved[3], saved[4]
            torch.cuda.set_rng_state(saved[5], dev)
            if rng is not None:
                rng.offset_dev.copy_(saved[6])
            for p in opt.param_groups[0]["params"]:
                torch.autograd.graph.increment_version(p)
            if self.graph is None:
                opt.scalars_dev = None
                try:        # a capture that failed half-way leaves torch's CUDA generator in "capturing" state (its next
                    with torch.cuda.graph(torch.cuda.CUDAGraph()):      # eager draw raises); a tiny complete capture resets it
                        torch.rand(1, device=dev)
                    torch.cuda.set_rng_state(saved[5], dev)
                except Exception:
                    pass

    def _refresh_scalars(self):
        system, opt = self.system, self.opt
        lr = system.learning_rate(system.trainer.global_step)
        for pg in opt.param_groups:
            pg["lr"] = lr
        g = opt.param_groups[0]
        scale = 1.0 / self.sync.world if (self.sync is not None and self.sync.world > 1) else 1.0
        i = self.sc_slot
        self.sc_slot = (i + 1) % self.RING
        if self.sc_events[i] is not None:
            self.sc_events[i].synchronize()
        L.adam_scalars(lr, g["betas"][0], g["betas"][1], g["eps"], opt.steps + 1, scale, self.sc_ring[i])
        self.sc_dev.copy_(self.sc_ring[i], non_blocking=True)
        if self.sc_events[i] is None:
            self.sc_events[i] = torch.cuda.Event()
        self.sc_events[i].record()

    def _body(self):
        system, opt = self.system, self.opt
        opt.zero_grad(set_to_none=False)
        if self.sync is not None:
            self.sync.start()
        loss = system.training_step(self.static, 0)
        loss.backward()
        # keep only the VALUE: a live reference to the loss keeps this iteration's autograd graph -- and its AccumulateGrad
        # nodes, which are bound to the stream they were created on -- alive; the capture would then accumulate gradients on
        # the warm-up stream, outside the graph
        self.loss = loss.detach()
        del loss
        if self.sync is not None:
            self.sync.finish()
        opt.step()

    def _advance(self):
        self.system.trainer.global_step += 1

    def __call__(self, batch) -> Tensor:
        for k, v in batch.items():
            if torch.is_tensor(v) and v.is_cuda:
                self.static[k].copy_(v, non_blocking=True)
            elif torch.is_tensor(v) and not torch.equal(v, self.static[k]):
                # a host tensor was read (if at all) while the graph was captured: its value is baked into the replay
                raise L.AonError("GraphedStep: batch['%s'] is a host tensor whose value changed since the capture; pass it as a "
                                 "CUDA tensor (the reference's datasets do) or run eager steps (AON_TRAIN_GRAPH=0)" % k)
        self._refresh_scalars()
        self.opt.steps += 1                   # the captured opt.step() does not run Python on replay
        self.graph.replay()
        for p in self.opt.param_groups[0]["params"]:
            torch.autograd.graph.increment_version(p)
        self._advance()
        return self.loss


class _LitCommon(LitModel):
    near, far, white_bkgd = 2.0, 6.0, True     # datasets/sapien.py:72-73 constants; setup() may override

    def _init(self, hparams, lr_init, lr_final, lr_delay_steps, lr_delay_mult, randomized):
        self.hparams = _hp(hparams)
        self.lr_init, self.lr_final = lr_init, lr_final
        self.lr_delay_steps, self.lr_delay_mult = lr_delay_steps, lr_delay_mult
        self.randomized = randomized
        self.white_bkgd = bool(self.hparams.white_back)

    def setup(self, stage: Optional[str] = None, datasets: Optional[dict] = None) -> None:
        """The reference builds SapienDataset objects here (model.py:220-254).  Dataset IO is out of the
        hot-path scope: callers hand in ``datasets={'train':..,'val':..,'test':..}``
        objects exposing ``near / far / white_back`` like the reference's."""
        for k, v in (datasets or {}).items():
            setattr(self, k + "_dataset", v)
            self.near, self.far = getattr(v, "near", self.near), getattr(v, "far", self.far)
            self.white_bkgd = getattr(v, "white_back", self.white_bkgd)

    # ---- optimisation (model.py:386-419) ----
    def configure_optimizers(self):
        """Adam(lr_init, betas (0.9, 0.999)) like model.py:386-389, as the flat-buffer aon_adam_step optimizer."""
        return FlatAdam(self.parameters(), lr=self.lr_init, betas=(0.9, 0.999))

    def learning_rate(self, step: int) -> float:
        if self.lr_delay_steps > 0:
            delay = self.lr_delay_mult + (1 - self.lr_delay_mult) * np.sin(
                0.5 * np.pi * np.clip(step / self.lr_delay_steps, 0, 1))
        else:
            delay = 1.0
        t = np.clip(step / self.hparams.run_max_steps, 0, 1)
        return float(delay * np.exp(np.log(self.lr_init) * (1 - t) + np.log(self.lr_final) * t))

    def optimizer_step(self, epoch, batch_idx, optimizer, optimizer_idx=0, optimizer_closure=None, on_tpu=False,
                       using_native_amp=False, using_lbfgs=False):
        lr = self.learning_rate(self.trainer.global_step)
        for pg in optimizer.param_groups:
            pg["lr"] = lr
        optimizer.step(closure=optimizer_closure)

    # ---- evaluation artefacts (model.py:459-507; interface.py:64-99,125-138) ----
    @torch.no_grad()
    def psnr_each(self, preds, gts):
        return torch.stack([self.psnr_clipped(p, g) for p, g in zip(preds, gts)])

    @torch.no_grad()
    def psnr(self, preds, gts, i_train=None, i_val=None, i_test=None, name="PSNR"):
        m = self.psnr_each(preds, gts).mean().item()
        return {"name": name, "mean": m, "test": m}

    @torch.no_grad()
    def ssim_each(self, preds, gts):
        """interface.py:102-112: piqa.SSIM() of every (pred, gt) pair of [H,W,3] images clipped to [0,1].  piqa is not in this
        image; the metric is restated here from its definition (Wang et al. 2004 with piqa's defaults: 11-tap Gaussian window,
        sigma 1.5, separable, no padding, k1 = 0.01, k2 = 0.03, value range 1, mean over channels and pixels)."""
        k = torch.arange(11, dtype=torch.float32) - 5.0
        k = torch.exp(-k ** 2 / (2 * 1.5 ** 2))
        k = k / k.sum()
        out = []
        for pred, gt in zip(preds, gts):
            x = torch.clip(pred.permute(2, 0, 1).unsqueeze(0).float(), 0, 1)
            y = torch.clip(gt.permute(2, 0, 1).unsqueeze(0).float(), 0, 1)
            C = x.shape[1]
            wv, wh = k.to(x.device).view(1, 1, 11, 1).repeat(C, 1, 1, 1), k.to(x.device).view(1, 1, 1, 11).repeat(C, 1, 1, 1)
            blur = lambda t: torch.nn.functional.conv2d(torch.nn.functional.conv2d(t, wv, groups=C), wh, groups=C)
            mu_x, mu_y = blur(x), blur(y)
            mu_xx, mu_yy, mu_xy = mu_x ** 2, mu_y ** 2, mu_x * mu_y
            s_xx, s_yy, s_xy = blur(x ** 2) - mu_xx, blur(y ** 2) - mu_yy, blur(x * y) - mu_xy
            c1, c2 = 0.01 ** 2, 0.03 ** 2
            cs = (2 * s_xy + c2) / (s_xx + s_yy + c2)
            ss = (2 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
            out.append(ss.flatten(1).mean(-1).mean())
        return torch.stack(out)

    @torch.no_grad()
    def ssim(self, preds, gts, i_train=None, i_val=None, i_test=None):
        m = self.ssim_each(preds, gts).mean().item()
        return {"name": "SSIM", "mean": m, "test": m}

    @torch.no_grad()
    def test_epoch_end(self, outputs):
        """outputs: list of test_step dicts, one per image ([H*W,3] rgb / target, [H*W] instance_mask).  Regroups them
        per image (every rank renders whole images here, so the reference's per-pixel all_gather interleave,
        interface.py:31-51, is not needed), computes PSNR, SSIM and object-masked PSNR, and on the global-zero rank writes
        ckpts/{exp_name}/{render_name}/imageNNN.jpg and ckpts/{exp_name}/results.json (model.py:459-507).  LPIPS needs the
        pretrained VGG weights piqa downloads (no network here): not computed, and not written."""
        W, H = self.hparams.img_wh
        rgbs = [o["rgb"].reshape(H, W, 3) for o in outputs]
        targets = [o["target"].reshape(H, W, 3) for o in outputs]
        masks = [o["instance_mask"].reshape(H, W).bool() for o in outputs]
        psnr = self.psnr(rgbs, targets)
        ssim = self.ssim(rgbs, targets) if min(H, W) >= 11 else {"name": "SSIM", "mean": float("nan"), "test": float("nan")}
        obj = [(r[m], t[m]) for r, t, m in zip(rgbs, targets, masks) if m.any()]
        psnr_obj = self.psnr([a for a, _ in obj], [b for _, b in obj], name="PSNR_obj") if obj else {"name": "PSNR_obj", "mean": float("nan"), "test": float("nan")}
        self.log("test/psnr", psnr["test"])
        self.log("test/ssim", ssim["test"])
        self.log("test/psnr_obj", psnr_obj["test"])
        if self.trainer.is_global_zero:
            exp, ren = getattr(self.hparams, "exp_name", "exp"), getattr(self.hparams, "render_name", "render")
            image_dir = os.path.join(getattr(self.hparams, "ckpt_root", "ckpts"), exp, ren)
            os.makedirs(image_dir, exist_ok=True)
            store_image(image_dir, rgbs, "image")
            write_stats(os.path.join(os.path.dirname(image_dir), "results.json"), psnr, ssim, psnr_obj)
        return psnr, ssim, psnr_obj

    @staticmethod
    def _squeeze(batch, keep=()):
        return {k: (v if (k in keep or not torch.is_tensor(v)) else v.squeeze(0)) for k, v in batch.items()}

    def _losses(self, rendered, target):
        loss0 = img2mse(rendered[0][0], target)
        loss1 = img2mse(rendered[1][0], target)
        self.log("train/psnr1", mse2psnr(loss1.detach()))
        self.log("train/psnr0", mse2psnr(loss0.detach()))
        return loss0, loss1


class LitNeRF(_LitCommon):
    def __init__(self, hparams, lr_init: float = 5.0e-4, lr_final: float = 5.0e-6, lr_delay_steps: int = 2500,
                 lr_delay_mult: float = 0.01, randomized: bool = True):
        super().__init__()
        self._init(hparams, lr_init, lr_final, lr_delay_steps, lr_delay_mult, randomized)
        self.model = NeRF()

    def training_step(self, batch, batch_idx):
        batch = self._squeeze(batch, keep=("obj_idx",))
        rendered = self.model(batch, self.randomized, self.white_bkgd, self.near, self.far)
        loss0, loss1 = self._losses(rendered, batch["target"])
        loss = loss1 + loss0
        self.log("train/loss", loss.detach())
        return loss

    @torch.no_grad()
    def render_rays(self, batch, batch_idx=0):
        fine = self.model(batch, False, self.white_bkgd, self.near, self.far)[1]
        ret = {"comp_rgb": fine[0], "acc": fine[1], "depth": fine[2]}
        if "target" in batch:
            self.log("val/psnr", self.psnr_legacy(ret["comp_rgb"], batch["target"]).mean())
        return ret

    @torch.no_grad()
    def render_rays_test(self, batch, batch_idx=0):
        fine = self.model(batch, False, self.white_bkgd, self.near, self.far)[1]
        return {"target": batch.get("target"), "instance_mask": batch.get("instance_mask"), "rgb": fine[0]}

    def on_validation_start(self):
        self.random_batch = 0

    def validation_step(self, batch, batch_idx):
        return self.render_rays(self._squeeze(batch, keep=("obj_idx",)), batch_idx)

    def test_step(self, batch, batch_idx):
        return self.render_rays_test({k: (v.squeeze() if torch.is_tensor(v) else v) for k, v in batch.items()}, batch_idx)


class LitNeRF_AutoDecoder(_LitCommon):
    def __init__(self, hparams, lr_init: float = 5.0e-4, lr_final: float = 5.0e-6, lr_delay_steps: int = 2500,
                 lr_delay_mult: float = 0.01, randomized: bool = True):
        super().__init__()
        self._init(hparams, lr_init, lr_final, lr_delay_steps, lr_delay_mult, randomized)
        self.model = NeRF_AE_Art()
        self.code_library = CodeLibraryArticulated(self.hparams)

    def training_step(self, batch, batch_idx):
        batch = self._squeeze(batch, keep=("deg", "instance_id", "articulation_id"))
        latents = self.code_library(batch)
        rendered = self.model(batch, self.randomized, self.white_bkgd, self.near, self.far, latents)
        loss0, loss1 = self._losses(rendered, batch["target"])
        # model_autodecoder.py:456-466: 1e-4 * sum of mean column norms of the three codes
        reg = 1e-4 * sum(torch.mean(torch.norm(latents[k], dim=0)) for k in ("density", "color", "articulation"))
        loss = loss1 + loss0 + reg
        self.log("train/loss", loss.detach())
        self.log("train/loss/reg", reg.detach())
        return loss

    @torch.no_grad()
    def render_rays(self, batch, latents):
        fine = self.model(batch, False, self.white_bkgd, self.near, self.far, latents)[1]
        ret = {"comp_rgb": fine[0], "acc": fine[1], "depth": fine[2]}
        if "target" in batch:
            self.log("val/psnr", self.psnr_legacy(ret["comp_rgb"], batch["target"]).mean())
            if "instance_mask" in batch:
                m = batch["instance_mask"].view(-1, 1).repeat(1, 3).bool()
                if m.any():
                    self.log("val/psnr_obj", self.psnr_legacy(ret["comp_rgb"][m], batch["target"][m]).mean())
        return ret

    @torch.no_grad()
    def render_rays_test(self, batch, latents):
        fine = self.model(batch, False, self.white_bkgd, self.near, self.far, latents)[1]
        return {"target": batch.get("target"), "instance_mask": batch.get("instance_mask"), "rgb": fine[0]}

    def on_validation_start(self):
        self.random_batch = 0

    def validation_step(self, batch, batch_idx):
        batch = self._squeeze(batch, keep=("deg", "instance_id", "articulation_id", "img_wh", "src_imgs"))
        return self.render_rays(batch, self.code_library(batch))

    def test_step(self, batch, batch_idx):
        batch = self._squeeze(batch, keep=("deg", "instance_id", "articulation_id", "img_wh", "src_imgs"))
        return self.render_rays_test(batch, self.code_library(batch, is_test=True))


def build_system(hparams):
    """run.py:21-34 dispatch on --exp_type."""
    exp = getattr(hparams, "exp_type", "vanilla")
    if exp == "vanilla":
        return LitNeRF(hparams)
    if exp == "vanilla_autodecoder":
        return LitNeRF_AutoDecoder(hparams)
    raise ValueError("exp_type %r is outside the hot-path scope (SURVEY.md section 2 rows 11-12)" % exp)


class Trainer:
    """Minimal stand-in for ``pl.Trainer(...).fit/test`` (run.py:135-166): drives the hooks above, one
    process per GPU.  Training gradients are averaged across ranks with ONE flat all-reduce per step
    (DDP semantics of run.py:109; dist.allreduce_mean_)."""

    def __init__(self, max_steps: int = 1000, log_every: int = 0, cuda_graph: Optional[bool] = None):
        self.max_steps, self.log_every = max_steps, log_every
        # replay the step as a CUDA graph (GraphedStep); default: on for the tcgen05 training paths, AON_TRAIN_GRAPH=0/1 overrides
        self.cuda_graph = cuda_graph if cuda_graph is not None else os.environ.get("AON_TRAIN_GRAPH", "1") == "1"
        self.global_step = 0
        self.is_global_zero = D.world()[0] == 0
        self.ckpt_every, self.on_checkpoint = 0, None     # periodic checkpoints (run.py sets both on rank 0)

    def resume(self, system: "_LitCommon", blob: dict) -> None:
        """Continue a run from a checkpoint written by run.save_checkpoint: the step count (LR warm-up / decay position)
        and the Adam moments + bias-correction step.  The weights were loaded by the caller (load_state_dict)."""
        self.global_step = int(blob.get("global_step", 0))
        # the in-kernel sampling draws are a function of (seed, step): continue where the checkpointed run stopped
        system.model.rng_offset0 = self.global_step
        if getattr(system.model, "_rng", None) is not None:
            system.model._rng.offset_dev.fill_(self.global_step)
        states = blob.get("optimizer_states") or []
        if states:
            opt = getattr(system, "_optimizer", None) or system.configure_optimizers()
            opt.load_state_dict(states[0])
            system._optimizer = opt
            system._resumed = True

    def fit(self, system: _LitCommon, batches) -> _LitCommon:
        system.trainer = self
        opt = getattr(system, "_optimizer", None) or system.configure_optimizers()
        if (getattr(system, "_optimizer", None) is None or getattr(system, "_resumed", False)) and D.world()[1] > 1:
            system._resumed = False
            torch.distributed.broadcast(opt.flat, src=0)     # DDP start state: every rank trains rank 0's initial weights
            for p in opt.param_groups[0]["params"]:
                torch.autograd.graph.increment_version(p)    # the packed-weight cache keys on parameter versions
        system._optimizer = opt                          # Adam moments survive repeated fit() calls
        system.train()
        sync = getattr(system, "_grad_sync", None)
        if sync is None:
            sync = system._grad_sync = GradSync(list(system.named_parameters()), opt.flat_grad)
        graphed, shapes = None, None
        # several ranks: the per-MLP all-reduces are captured with the step (NCCL collectives are graph-capturable; every rank
        # captures the same sequence), AON_TRAIN_GRAPH_NCCL=0 keeps multi-rank steps eager
        multi_ok = sync.world == 1 or os.environ.get("AON_TRAIN_GRAPH_NCCL", "1") == "1"
        use_graph = self.cuda_graph and multi_ok and getattr(system.model, "train_gemm", "torch") in ("tc", "tc16")
        for batch_idx, batch in enumerate(batches):
            if self.global_step >= self.max_steps:
                break
            if use_graph:
                sig = tuple((k, tuple(v.shape)) for k, v in sorted(batch.items()) if torch.is_tensor(v))
                if graphed is None or sig != shapes:
                    try:
                        graphed, shapes = GraphedStep(system, opt, batch, sync if sync.world > 1 else None), sig   # dry run + capture; trains nothing
                    except Exception as e:       # capture refused (see GraphedStep): eager steps from here on
                        if self.is_global_zero:
                            print("lit.Trainer: CUDA-graph capture of the training step failed (%s: %s); running eager steps"
                                  % (type(e).__name__, str(e).splitlines()[0]), flush=True)
                        use_graph, graphed = False, None
                if graphed is not None:
                    graphed(batch)
            if not use_graph:
                opt.zero_grad(set_to_none=False)
                sync.start()
                loss = system.training_step(batch, batch_idx)
                loss.backward()                          # the fine MLP's gradient slice is all-reduced while the coarse MLP's backward runs
                opt.grad_scale = sync.finish()
                system.optimizer_step(0, batch_idx, opt, 0, None, False, False, False)
                self.global_step += 1
            if self.log_every and self.global_step % self.log_every == 0:
                # the fp16 hi+lo operand planes of the training GEMMs saturate instead of overflowing; a non-finite gradient
                # (or loss) means the power-of-two operand scaling of train_tc.py does not fit this model / loss: fail loudly
                if not bool(torch.isfinite(opt.flat_grad).all()) or not math.isfinite(system.logged.get("train/loss", 0.0)):
                    raise L.AonError("non-finite training gradient / loss at step %d: operand range of the tcgen05 training GEMMs "
                                     "exceeded (train_tc.py scales: activations x8, weights x64); use train_gemm = 'torch'" % self.global_step)
            if self.log_every and self.is_global_zero and self.global_step % self.log_every == 0:
                print("step %d  %s" % (self.global_step, {k: round(v, 4) for k, v in system.logged.items()}), flush=True)
            if self.ckpt_every and self.on_checkpoint is not None and self.global_step % self.ckpt_every == 0:
                self.on_checkpoint(self.global_step)
        return system

    @torch.no_grad()
    def test(self, system: _LitCommon, batches) -> list:
        system.trainer = self
        system.eval()
        outputs = [system.test_step(b, i) for i, b in enumerate(batches)]
        system.test_results = system.test_epoch_end(outputs) if all(o.get("target") is not None for o in outputs) else None
        return outputs
