"""ctypes binding of libaon_b200.so (C ABI declared in include/aon.h).

The product path FAILS LOUDLY when the library is missing or a call fails -- there is no CPU /
eager fallback anywhere in this package.  torch is used only as the owner of device memory and
streams: every wrapper takes CUDA fp32 contiguous tensors and passes raw pointers + the current
stream to the C entry points.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaon_b200.so")

KIND_VANILLA, KIND_AUTODECODER = 0, 1
PREC_FP32, PREC_TC_F16X3, PREC_TC_F16, PREC_TC_BF16 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "f16x3": PREC_TC_F16X3, "f16": PREC_TC_F16, "bf16": PREC_TC_BF16}

# every symbol include/aon.h declares: name -> (restype, argtypes)
_vp, _fp, _i, _l, _f, _sz, _d = C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_size_t, C.c_double
SYMBOLS = {
    "aon_version": (_i, []),
    "aon_last_error": (C.c_char_p, []),
    "aon_num_layers": (_i, [_i]),
    "aon_layer_shape": (_i, [_i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "aon_packed_bytes": (_sz, [_i, _i]),
    "aon_pack_weights": (_i, [_i, _i, C.POINTER(_vp), C.POINTER(_vp), _vp, _sz, _vp]),
    "aon_folded_floats": (_sz, [_i]),
    "aon_fold_latents": (_i, [_i, _i, _vp, _fp, _fp, _fp, _fp, _vp]),
    "aon_raygen": (_i, [_i, _i, _f, C.POINTER(_f), _fp, _fp, _vp]),
    "aon_sample_along_rays": (_i, [_f, _f, _i, _fp, _i, _fp, _vp]),
    "aon_sample_along_rays_rng": (_i, [_f, _f, _i, _vp, _i, _fp, _vp]),
    "aon_rng_uniform": (_i, [_vp, _i, _i, _i, _fp, _vp]),
    "aon_rng_advance": (_i, [_vp, C.c_ulonglong, _vp]),
    "aon_workspace_bytes": (_sz, [_i, _i]),
    "aon_render_level": (_i, [_i, _i, _vp, _fp, _fp, _fp, _fp, _fp, _l, _i, _i, _i, _fp, _fp, _fp, _fp, _vp, _sz, _vp, _vp]),
    "aon_sample_pdf": (_i, [_fp, _l, _fp, _fp, _l, _i, _i, _i, _fp, _vp]),
    "aon_sample_pdf_rng": (_i, [_fp, _l, _fp, _vp, _i, _i, _i, _fp, _vp]),
    "aon_render_rays": (_i, [_i, _i, _vp, _vp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _f, _f, _i, _fp, _fp, _vp, _sz, _vp, _vp]),
    "aon_render_image": (_i, [_i, _i, _vp, _vp, _fp, _fp, C.POINTER(_f), _f, _i, _i, _l, _i, _f, _f, _i, _fp, _fp, _vp, _sz, _vp, _vp]),
    "aon_render_image_host": (_i, [_i, _i, _vp, _vp, _fp, _fp, _fp, _fp, _fp, _i, _f, _f, _i, _fp, _fp, _vp, _sz, _vp, _vp]),
    "aon_pos_enc": (_i, [_fp, _l, _i, _fp, _vp]),
    "aon_pos_enc_backward": (_i, [_fp, _fp, _l, _i, _fp, _vp]),
    "aon_composite": (_i, [_fp, _fp, _fp, _l, _fp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _vp]),
    "aon_composite_backward": (_i, [_fp, _fp, _fp, _l, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _vp]),
    "aon_adam_step": (_i, [_fp, _fp, _fp, _fp, _l, _d, _d, _d, _d, _l, _d, _vp]),
    "aon_adam_scalars": (_i, [_d, _d, _d, _d, _l, _d, C.POINTER(_f)]),
    "aon_adam_step_dev": (_i, [_fp, _fp, _fp, _fp, _l, _fp, _vp]),
    "aon_gemm_tc": (_i, [_vp, _vp]),
    "aon_gemm_struct_size": (_sz, []),
    "aon_gemm_colsum_rows": (_i, [_vp]),
    "aon_colsum_finish": (_i, [_fp, _i, _i, _f, _fp, _vp]),
    "aon_pack_rows": (_i, [_fp, _l, _i, _l, _i, _i, _i, _f, _vp, _vp, _vp]),
    "aon_pack_linear": (_i, [_fp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "aon_wgrad_reduce": (_i, [_fp, _i, _i, _i, _f, _fp, _l, _i, _i, _i, _i, _vp]),
    "aon_colsum_packed": (_i, [_vp, _vp, _i, _i, _i, _fp, _vp]),
    "aon_train_tiles": (_i, [_i, _i]),
    "aon_forward_train": (_i, [_i, _i, _vp, _fp, _fp, _fp, _fp, _fp, _l, _i, _i, _vp, _vp]),
    "aon_pack_rows_tiled": (_i, [_fp, _l, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "aon_unpack_rows_tiled": (_i, [_fp, _l, _i, _i, _i, _fp, _vp]),
    "aon_launch_count": (_l, [_i]),
}


class AonError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the library (once).  Raises AonError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AonError("libaon_b200.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; "
                       "g.build()'` or `python articulated-object-nerf_b200/build.py`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise AonError("%s failed (%d): %s" % (what, rc, load().aon_last_error().decode()))


def _ptr(t: Optional[torch.Tensor], what: str = "tensor") -> Optional[int]:
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise AonError("%s must be a contiguous CUDA float32 tensor (got %s %s contiguous=%s)"
                       % (what, t.device, t.dtype, t.is_contiguous()))
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _Null()


def _on(device):
    """`with torch.cuda.device(d)` costs ~10 us of host time per call -- more than enqueueing the kernel -- and a training step
    makes ~300 calls: switch only when the tensor's device is not already the current one."""
    d = torch.device(device) if not isinstance(device, torch.device) else device
    idx = d.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(d)


def launch_count(reset: bool = False) -> int:
    return int(load().aon_launch_count(1 if reset else 0))


def layer_shapes(kind: int) -> List[tuple]:
    lib = load()
    out = []
    for i in range(lib.aon_num_layers(kind)):
        o, k = C.c_int(), C.c_int()
        _check(lib.aon_layer_shape(kind, i, C.byref(o), C.byref(k)), "aon_layer_shape")
        out.append((o.value, k.value))
    return out


def pack_weights(kind: int, precision: int, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[out,in] fp32 nn.Linear weights (state_dict order) -> kernel layout.  Returns a uint8 CUDA tensor."""
    lib = load()
    n = lib.aon_num_layers(kind)
    if len(weights) != n or len(biases) != n:
        raise AonError("expected %d layers, got %d/%d" % (n, len(weights), len(biases)))
    shapes = layer_shapes(kind)
    keep = []
    for i, (w, b) in enumerate(zip(weights, biases)):
        if tuple(w.shape) != shapes[i] or tuple(b.shape) != (shapes[i][0],):
            raise AonError("layer %d: expected weight %s, got %s / bias %s" % (i, shapes[i], tuple(w.shape), tuple(b.shape)))
        keep.append((w.detach().contiguous().float(), b.detach().contiguous().float()))
    nbytes = lib.aon_packed_bytes(kind, precision)
    if nbytes == 0:
        raise AonError("aon_packed_bytes: unsupported kind/precision %d/%d" % (kind, precision))
    dev = keep[0][0].device
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    wp = (C.c_void_p * n)(*[_ptr(w, "weight") for w, _ in keep])
    bp = (C.c_void_p * n)(*[_ptr(b, "bias") for _, b in keep])
    with _on(dev):
        _check(lib.aon_pack_weights(kind, precision, wp, bp, out.data_ptr(), out.numel(), _stream()), "aon_pack_weights")
    return out


def fold_latents(kind: int, precision: int, packed: torch.Tensor, shape: torch.Tensor, appearance: torch.Tensor,
                 articulation: torch.Tensor) -> torch.Tensor:
    lib = load()
    folded = torch.empty(lib.aon_folded_floats(kind), dtype=torch.float32, device=packed.device)
    with _on(packed.device):
        _check(lib.aon_fold_latents(kind, precision, packed.data_ptr(), _ptr(shape.reshape(-1), "shape"),
                                    _ptr(appearance.reshape(-1), "appearance"),
                                    _ptr(articulation.reshape(-1), "articulation"), _ptr(folded), _stream()),
               "aon_fold_latents")
    return folded


def raygen(H: int, W: int, focal: float, c2w, device) -> tuple:
    """A1+A2: (rays_o [HW,3], rays_d [HW,3]); rays_d is unit-norm and doubles as viewdirs."""
    lib = load()
    c = torch.as_tensor(c2w, dtype=torch.float32).reshape(-1)[:12].cpu().contiguous()
    arr = (C.c_float * 12)(*c.tolist())
    o = torch.empty(H * W, 3, dtype=torch.float32, device=device)
    d = torch.empty(H * W, 3, dtype=torch.float32, device=device)
    with _on(o.device):
        _check(lib.aon_raygen(H, W, float(focal), arr, _ptr(o), _ptr(d), _stream()), "aon_raygen")
    return o, d


class AonRng(C.Structure):
    """Mirror of `struct AonRng` in include/aon.h."""
    _fields_ = [("seed", C.c_ulonglong), ("offset", C.c_ulonglong), ("offset_dev", _vp)]


class Rng:
    """State of the in-kernel Philox draws: a 64-bit seed and a step offset that lives in DEVICE memory (one int64), so a
    training step captured in a CUDA graph advances it by itself (aon_rng_advance inside the graph)."""

    def __init__(self, seed: int, device):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset_dev = torch.zeros(1, dtype=torch.int64, device=device)

    def struct(self, offset: int = 0) -> AonRng:
        return AonRng(self.seed, int(offset), self.offset_dev.data_ptr())

    def advance(self, by: int = 1) -> None:
        with _on(self.offset_dev.device):
            _check(load().aon_rng_advance(self.offset_dev.data_ptr(), int(by), _stream()), "aon_rng_advance")

    def uniform(self, stream_id: int, rows: int, cols: int) -> torch.Tensor:
        """the draws [rows, cols] of stream 0 (stratified) / 1 (inverse cdf) at the current offset -- tests and debugging"""
        out = torch.empty(rows, cols, dtype=torch.float32, device=self.offset_dev.device)
        with _on(out.device):
            r = self.struct()
            _check(load().aon_rng_uniform(C.byref(r), stream_id, rows, cols, _ptr(out), _stream()), "aon_rng_uniform")
        return out


def sample_along_rays(near: float, far: float, n_points: int, R: int, device, t_rand: Optional[torch.Tensor] = None,
                      rng: Optional[Rng] = None):
    """A3: deterministic -> shared table [n_points]; with t_rand [R,n_points] (or rng: draws generated in the kernel)
    -> [R,n_points]."""
    lib = load()
    shape = (n_points,) if (t_rand is None and rng is None) else (R, n_points)
    t = torch.empty(shape, dtype=torch.float32, device=device)
    with _on(t.device):
        if rng is not None and t_rand is None:
            r = rng.struct()
            _check(lib.aon_sample_along_rays_rng(float(near), float(far), n_points, C.byref(r), R, _ptr(t), _stream()),
                   "aon_sample_along_rays_rng")
        else:
            _check(lib.aon_sample_along_rays(float(near), float(far), n_points, _ptr(t_rand, "t_rand"), R, _ptr(t), _stream()),
                   "aon_sample_along_rays")
    return t


class AonRenderOpts(C.Structure):
    """Mirror of `struct AonRenderOpts` in include/aon.h (per-call debugging / A-B knobs; the library keeps no switches)."""
    _fields_ = [("force_segments", _i), ("no_tail_split", _i), ("no_fuse", _i), ("reserved", _i),
                ("dbg", _vp), ("err_flag", _vp), ("timeline", _vp)]


# Debug knobs live HERE (test / tool convenience), not in the C library: every render call passes them in an AonRenderOpts.
_dbg = {"force_segments": 0, "no_tail_split": 0, "no_fuse": 0, "dbg": None, "err": None, "tl": None}


def _opts():
    if not any((v is not None) if not isinstance(v, int) else (v != 0) for v in _dbg.values()):
        return None
    o = AonRenderOpts()
    o.force_segments, o.no_tail_split, o.no_fuse = int(_dbg["force_segments"]), int(_dbg["no_tail_split"]), int(_dbg["no_fuse"])
    o.dbg = None if _dbg["dbg"] is None else _dbg["dbg"].data_ptr()
    o.err_flag = None if _dbg["err"] is None else _dbg["err"].data_ptr()
    o.timeline = None if _dbg["tl"] is None else _dbg["tl"].data_ptr()
    return C.byref(o)


_ws_cache = {}


def workspace(device, precision: int, R: int) -> torch.Tensor:
    """Caller-owned scratch for the render entry points (aon_workspace_bytes), cached per (device, stream) and grown on
    demand; a uint8 CUDA tensor from torch's caching allocator."""
    need = int(load().aon_workspace_bytes(precision, R))
    if need == 0:
        raise AonError("aon_workspace_bytes: bad precision %d" % precision)
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def render_level(kind: int, precision: int, packed: torch.Tensor, folded: Optional[torch.Tensor], rays_o, rays_d,
                 viewdirs, t_vals: torch.Tensor, white_bkgd: bool, want_weights: bool = True):
    """One level.  t_vals [S] (shared) or [R,S].  Returns (comp_rgb [R,3], acc [R], depth [R], weights [R,S]|None)."""
    lib = load()
    R = rays_o.shape[0]
    S = t_vals.shape[-1]
    stride = 0 if t_vals.dim() == 1 else S
    dev = rays_o.device
    rgb = torch.empty(R, 3, dtype=torch.float32, device=dev)
    acc = torch.empty(R, dtype=torch.float32, device=dev)
    depth = torch.empty(R, dtype=torch.float32, device=dev)
    w = torch.empty(R, S, dtype=torch.float32, device=dev) if want_weights else None
    with _on(dev):
        ws = workspace(dev, precision, R)
        _check(lib.aon_render_level(kind, precision, packed.data_ptr(), _ptr(folded, "folded"), _ptr(rays_o, "rays_o"),
                                    _ptr(rays_d, "rays_d"), _ptr(viewdirs, "viewdirs"), _ptr(t_vals, "t_vals"), stride,
                                    R, S, int(bool(white_bkgd)), _ptr(rgb), _ptr(acc), _ptr(depth), _ptr(w), ws.data_ptr(),
                                    ws.numel(), _opts(), _stream()),
               "aon_render_level")
    return rgb, acc, depth, w


def sample_pdf(t_coarse: torch.Tensor, weights: torch.Tensor, n_fine: int, u: Optional[torch.Tensor] = None,
               rng: Optional[Rng] = None):
    """A7: t_coarse [n_coarse] or [R,n_coarse]; weights [R,n_coarse]; u None | [n_fine] | [R,n_fine]; rng (with u None):
    the inverse-cdf draws are generated in the kernel."""
    lib = load()
    R, nc = weights.shape
    t_stride = 0 if t_coarse.dim() == 1 else nc
    u_stride = 0 if (u is None or u.dim() == 1) else n_fine
    out = torch.empty(R, nc + n_fine, dtype=torch.float32, device=weights.device)
    with _on(weights.device):
        if rng is not None and u is None:
            r = rng.struct()
            _check(lib.aon_sample_pdf_rng(_ptr(t_coarse, "t_coarse"), t_stride, _ptr(weights, "weights"), C.byref(r),
                                          R, nc, n_fine, _ptr(out), _stream()), "aon_sample_pdf_rng")
            return out
        _check(lib.aon_sample_pdf(_ptr(t_coarse, "t_coarse"), t_stride, _ptr(weights, "weights"), _ptr(u, "u"), u_stride,
                                  R, nc, n_fine, _ptr(out), _stream()), "aon_sample_pdf")
    return out


def render_rays(kind: int, precision: int, packed_coarse, packed_fine, folded_coarse, folded_fine, rays_o, rays_d, viewdirs,
                near: float, far: float, white_bkgd: bool, t_coarse: Optional[torch.Tensor] = None,
                u: Optional[torch.Tensor] = None, want_coarse: bool = True):
    """A8/A11: the whole coarse -> sample_pdf -> fine loop of NeRF.forward in one fused kernel (aon_render_rays).
    Returns (fine [R,5], coarse [R,5] | None) with columns (r, g, b, acc, depth)."""
    lib = load()
    R = rays_o.shape[0]
    dev = rays_o.device
    out = torch.empty(R, 5, dtype=torch.float32, device=dev)
    cout = torch.empty(R, 5, dtype=torch.float32, device=dev) if want_coarse else None
    if t_coarse is not None and tuple(t_coarse.shape) != (R, 65):
        raise AonError("render_rays: t_coarse must be [R,65]")
    if u is not None and tuple(u.shape) != (R, 128):
        raise AonError("render_rays: u must be [R,128]")
    with _on(dev):
        ws = workspace(dev, precision, R)
        _check(lib.aon_render_rays(kind, precision, packed_coarse.data_ptr(), packed_fine.data_ptr(), _ptr(folded_coarse, "folded"),
                                   _ptr(folded_fine, "folded"), _ptr(rays_o, "rays_o"), _ptr(rays_d, "rays_d"),
                                   _ptr(viewdirs, "viewdirs"), _ptr(t_coarse, "t_coarse"), _ptr(u, "u"), R, float(near), float(far),
                                   int(bool(white_bkgd)), _ptr(out), _ptr(cout), ws.data_ptr(), ws.numel(), _opts(), _stream()),
               "aon_render_rays")
    return out, cout


def render_image(kind: int, precision: int, packed_coarse, packed_fine, folded_coarse, folded_fine, c2w, focal: float,
                 H: int, W: int, near: float, far: float, white_bkgd: bool, ray0: int = 0, R: Optional[int] = None,
                 want_coarse: bool = False, out: Optional[torch.Tensor] = None):
    """A1+A2 + A8/A11: pixels [ray0, ray0 + R) of the H x W view of camera c2w, ray generation fused into the render
    kernel (aon_render_image).  Returns (fine [R,5], coarse [R,5] | None)."""
    lib = load()
    if R is None:
        R = H * W - ray0
    dev = packed_coarse.device
    c = torch.as_tensor(c2w, dtype=torch.float32).reshape(-1)[:12].cpu().contiguous()
    arr = (C.c_float * 12)(*c.tolist())
    if out is None:
        out = torch.empty(R, 5, dtype=torch.float32, device=dev)
    cout = torch.empty(R, 5, dtype=torch.float32, device=dev) if want_coarse else None
    with _on(dev):
        ws = workspace(dev, precision, R)
        _check(lib.aon_render_image(kind, precision, packed_coarse.data_ptr(), packed_fine.data_ptr(), _ptr(folded_coarse, "folded"),
                                    _ptr(folded_fine, "folded"), arr, float(focal), H, W, int(ray0), R, float(near), float(far),
                                    int(bool(white_bkgd)), _ptr(out), _ptr(cout), ws.data_ptr(), ws.numel(), _opts(), _stream()),
               "aon_render_image")
    return out, cout


def render_image_host(kind: int, precision: int, packed_coarse, packed_fine, folded_coarse, folded_fine,
                      rays_o: torch.Tensor, rays_d: torch.Tensor, viewdirs: torch.Tensor, near: float, far: float,
                      white_bkgd: bool, out: Optional[torch.Tensor] = None, coarse_out: Optional[torch.Tensor] = None):
    """A8/A11 from HOST (ideally pinned) fp32 tensors; returns host [R,5] = rgb, acc, depth (fine level)."""
    lib = load()
    for t in (rays_o, rays_d, viewdirs):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise AonError("render_image_host takes contiguous CPU float32 tensors")
    R = rays_o.shape[0]
    if out is None:
        out = torch.empty(R, 5, dtype=torch.float32).pin_memory()
    dev = packed_coarse.device
    with _on(dev):
        ws = workspace(dev, precision, R)
        _check(lib.aon_render_image_host(kind, precision, packed_coarse.data_ptr(), packed_fine.data_ptr(),
                                         _ptr(folded_coarse, "folded"), _ptr(folded_fine, "folded"),
                                         rays_o.data_ptr(), rays_d.data_ptr(), viewdirs.data_ptr(), R, float(near),
                                         float(far), int(bool(white_bkgd)), out.data_ptr(),
                                         None if coarse_out is None else coarse_out.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _opts(), _stream()),
               "aon_render_image_host")
    return out


# ---- training path, stage 1: per-element stages with hand-written adjoints (csrc/train_ops.cu) ----------
def pos_enc(x: torch.Tensor, max_deg: int) -> torch.Tensor:
    """helper.py:136-140 with min_deg = 0: x [..., 3] -> [..., 3 + 6 max_deg]."""
    lib = load()
    n = x.numel() // 3
    out = torch.empty(tuple(x.shape[:-1]) + (3 + 6 * max_deg,), dtype=torch.float32, device=x.device)
    with _on(x.device):
        _check(lib.aon_pos_enc(_ptr(x, "x"), n, max_deg, _ptr(out), _stream()), "aon_pos_enc")
    return out


def pos_enc_backward(x: torch.Tensor, g_out: torch.Tensor, max_deg: int) -> torch.Tensor:
    lib = load()
    gx = torch.empty_like(x)
    with _on(x.device):
        _check(lib.aon_pos_enc_backward(_ptr(x, "x"), _ptr(g_out, "g_out"), x.numel() // 3, max_deg, _ptr(gx), _stream()),
               "aon_pos_enc_backward")
    return gx


def composite(raw_rgb: torch.Tensor, raw_sigma: torch.Tensor, t_vals: torch.Tensor, dirs: torch.Tensor, white_bkgd: bool,
              act_mode: int):
    """Activations + volumetric_rendering of raw MLP outputs [R,S,3] / [R,S].  Returns (comp_rgb, acc, depth, weights,
    trans); trans [R,S] is the exclusive transmittance the backward needs."""
    lib = load()
    R, S = raw_sigma.shape[0], raw_sigma.shape[1]
    dev = raw_rgb.device
    stride = 0 if t_vals.dim() == 1 else S
    e = lambda *sh: torch.empty(*sh, dtype=torch.float32, device=dev)
    rgb, acc, depth, w, tr = e(R, 3), e(R), e(R), e(R, S), e(R, S)
    with _on(dev):
        _check(lib.aon_composite(_ptr(raw_rgb, "raw_rgb"), _ptr(raw_sigma, "raw_sigma"), _ptr(t_vals, "t_vals"), stride,
                                 _ptr(dirs, "dirs"), R, S, int(bool(white_bkgd)), act_mode, _ptr(rgb), _ptr(acc), _ptr(depth),
                                 _ptr(w), _ptr(tr), _stream()), "aon_composite")
    return rgb, acc, depth, w, tr


def composite_backward(raw_rgb, raw_sigma, t_vals, dirs, weights, trans, g_rgb, g_acc, g_depth, white_bkgd: bool, act_mode: int):
    lib = load()
    R, S = raw_sigma.shape[0], raw_sigma.shape[1]
    stride = 0 if t_vals.dim() == 1 else S
    g_raw_rgb, g_raw_sigma = torch.empty_like(raw_rgb), torch.empty_like(raw_sigma)
    with _on(raw_rgb.device):
        _check(lib.aon_composite_backward(_ptr(raw_rgb, "raw_rgb"), _ptr(raw_sigma, "raw_sigma"), _ptr(t_vals, "t_vals"), stride,
                                          _ptr(dirs, "dirs"), _ptr(weights, "weights"), _ptr(trans, "trans"),
                                          _ptr(g_rgb, "g_rgb"), _ptr(g_acc, "g_acc"), _ptr(g_depth, "g_depth"), R, S,
                                          int(bool(white_bkgd)), act_mode, _ptr(g_raw_rgb), _ptr(g_raw_sigma), _stream()),
               "aon_composite_backward")
    return g_raw_rgb, g_raw_sigma


def adam_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float,
              beta1: float, beta2: float, eps: float, step: int, grad_scale: float = 1.0) -> None:
    """One Adam step over flat fp32 buffers, in place (torch.optim.Adam formulas)."""
    lib = load()
    with _on(params.device):
        _check(lib.aon_adam_step(_ptr(params, "params"), _ptr(grads, "grads"), _ptr(exp_avg, "exp_avg"),
                                 _ptr(exp_avg_sq, "exp_avg_sq"), params.numel(), float(lr), float(beta1), float(beta2),
                                 float(eps), int(step), float(grad_scale), _stream()), "aon_adam_step")


def adam_scalars(lr: float, beta1: float, beta2: float, eps: float, step: int, grad_scale: float, out: torch.Tensor) -> None:
    """the seven step-dependent floats of the Adam kernel into a HOST (pinned) float32 tensor [7]"""
    _check(load().aon_adam_scalars(float(lr), float(beta1), float(beta2), float(eps), int(step), float(grad_scale),
                                   C.cast(out.data_ptr(), C.POINTER(C.c_float))), "aon_adam_scalars")


def adam_step_dev(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, scalars_dev: torch.Tensor) -> None:
    """Adam step whose scalars live in device memory (graph-capturable)."""
    lib = load()
    with _on(params.device):
        _check(lib.aon_adam_step_dev(_ptr(params, "params"), _ptr(grads, "grads"), _ptr(exp_avg, "exp_avg"), _ptr(exp_avg_sq, "exp_avg_sq"),
                                     params.numel(), _ptr(scalars_dev, "scalars"), _stream()), "aon_adam_step_dev")


# ---- training path, stage 2: tcgen05 GEMMs on packed 16-bit hi/lo planes (csrc/gemm_tc.cu) ---------------------
GEMM_NT, GEMM_TN = 0, 1
EPI_LINEAR, EPI_MASK, EPI_PARTIAL = 0, 1, 2
_SEG = 2


class AonGemm(C.Structure):
    """Mirror of `struct AonGemm` in include/aon.h."""
    _fields_ = [("mode", _i), ("epi", _i), ("x3", _i), ("nseg", _i), ("N", _i), ("m_tiles", _i),
                ("a_hi", _vp * _SEG), ("a_lo", _vp * _SEG), ("b_hi", _vp * _SEG), ("b_lo", _vp * _SEG),
                ("a_feat", _i * _SEG), ("a_off", _i * _SEG), ("kext", _i * _SEG),
                ("b_feat", _i * _SEG), ("b_off", _i * _SEG), ("b_row0", _i * _SEG),
                ("a_tiles", _i), ("splits", _i), ("tiles_per_split", _i), ("relu", _i),
                ("partial", _vp), ("inv_scale", _f), ("out_scale", _f),
                ("n_valid", _i), ("mask_feat", _i), ("mask_off", _i), ("out_feat", _i), ("out_off", _i), ("reserved", _i),
                ("bias", _vp), ("mask_hi", _vp), ("out_f32", _vp), ("ldc", _l), ("out_hi", _vp), ("out_lo", _vp), ("colsum", _vp),
                ("relu_bits_out", _vp), ("mask_bits", _vp)]


class PK:
    """Activation / gradient matrix [m_tiles*128, feat] as 16-bit hi (+ lo) planes in the k-group packed layout
    [rows/128][feat/8][128][8] (see csrc/gemm_tc.cu)."""
    __slots__ = ("hi", "lo", "m_tiles", "feat", "bits")

    def __init__(self, m_tiles: int, feat: int, device, x3: bool = True):
        assert feat % 8 == 0
        self.m_tiles, self.feat = m_tiles, feat
        self.bits = None       # [rows, feat/32] int32 ReLU mask bit plane, written by the forward GEMM that produced this tensor
        self.hi = torch.empty(m_tiles * feat * 128, dtype=torch.float16, device=device)
        self.lo = torch.empty_like(self.hi) if x3 else None

    def tiles(self, t0: int, t1: int) -> "PK":
        """row tiles [t0, t1) as a PK that shares this one's storage (the layout is tile-major: a contiguous slice)"""
        v = PK.__new__(PK)
        v.m_tiles, v.feat = t1 - t0, self.feat
        n = self.feat * 128
        v.hi = self.hi[t0 * n:t1 * n]
        v.lo = None if self.lo is None else self.lo[t0 * n:t1 * n]
        v.bits = None if self.bits is None else self.bits[t0 * 128:t1 * 128]
        return v

    def to_dense(self) -> torch.Tensor:
        """fp32 [rows, feat] (hi + lo), for tests."""
        v = self.hi.float() + (self.lo.float() if self.lo is not None else 0)
        return v.view(self.m_tiles, self.feat // 8, 128, 8).permute(0, 2, 1, 3).reshape(self.m_tiles * 128, self.feat)


class PW:
    """Packed weight [k_pad/8][r_pad][8] hi (+ lo) planes: r_pad GEMM output rows, k_pad contraction."""
    __slots__ = ("hi", "lo", "r_pad", "k_pad")

    def __init__(self, r_pad: int, k_pad: int, device, x3: bool = True):
        self.r_pad, self.k_pad = r_pad, k_pad
        self.hi = torch.empty(r_pad * k_pad, dtype=torch.float16, device=device)
        self.lo = torch.empty_like(self.hi) if x3 else None


def _p(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pack_rows(src: torch.Tensor, M: int, m_tiles: int, c_pad: int, scale: float, row_div: int = 1, x3: bool = True) -> PK:
    """src fp32 [rows_in, C] (last dim contiguous) -> PK(m_tiles*128, c_pad) * scale; packed row m reads src[m // row_div]."""
    lib = load()
    if not (src.is_cuda and src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1):
        raise AonError("pack_rows: src must be a CUDA float32 [rows, C] tensor with contiguous rows")
    out = PK(m_tiles, c_pad, src.device, x3)
    with _on(src.device):
        _check(lib.aon_pack_rows(src.data_ptr(), src.stride(0), src.shape[1], M, row_div, m_tiles, c_pad, float(scale),
                                 out.hi.data_ptr(), _p(out.lo), _stream()), "aon_pack_rows")
    return out


def pack_linear(W: torch.Tensor, transpose: bool, r_pad: int, k_pad: int, scale: float, x3: bool = True) -> PW:
    lib = load()
    out = PW(r_pad, k_pad, W.device, x3)
    with _on(W.device):
        _check(lib.aon_pack_linear(_ptr(W, "W"), W.shape[0], W.shape[1], int(transpose), r_pad, k_pad, float(scale),
                                   out.hi.data_ptr(), _p(out.lo), _stream()), "aon_pack_linear")
    return out


def gemm_nt(segs, N: int, m_tiles: int, device, *, epi: int = EPI_LINEAR, bias=None, relu: bool = False, mask=None,
            inv_scale: float = 1.0, out_f32: Optional[torch.Tensor] = None, n_valid: int = 0, out: Optional[PK] = None,
            out_off: int = 0, out_scale: float = 1.0, x3: bool = True, colsum: bool = False) -> Optional[torch.Tensor]:
    """segs: [(A: PK, a_off, kext, B: PW, b_off, b_row0)], all accumulating into D[rows, N].  colsum=True returns partial
    column sums [rows, N] of the epilogue values (per tile, or per CTA of the persistent kernel; sum over dim 0 = bias gradient)."""
    lib = load()
    g = AonGemm()
    g.mode, g.epi, g.x3, g.nseg, g.N, g.m_tiles = GEMM_NT, epi, int(x3), len(segs), N, m_tiles
    for i, (A, a_off, kext, B, b_off, b_row0) in enumerate(segs):
        if A.m_tiles != m_tiles or a_off + kext > A.feat or b_off + kext > B.k_pad or b_row0 + N > B.r_pad:
            raise AonError("gemm_nt: segment %d out of range" % i)
        g.a_hi[i], g.a_lo[i], g.b_hi[i], g.b_lo[i] = A.hi.data_ptr(), _p(A.lo), B.hi.data_ptr(), _p(B.lo)
        g.a_feat[i], g.a_off[i], g.kext[i] = A.feat, a_off, kext
        g.b_feat[i], g.b_off[i], g.b_row0[i] = B.r_pad, b_off, b_row0
    g.relu, g.inv_scale, g.out_scale = int(relu), inv_scale, out_scale
    g.reserved = int(os.environ.get("AON_GEMM_DEBUG", "0"))
    g.bias = _p(bias)
    if mask is not None:
        mk, moff = mask
        if mk.bits is not None and moff == 0 and mk.feat == N:
            g.mask_bits = mk.bits.data_ptr()          # 32 B per row instead of a 512 B activation row
        else:
            g.mask_hi, g.mask_feat, g.mask_off = mk.hi.data_ptr(), mk.feat, moff
    if out_f32 is not None:
        if out_f32.shape[0] < m_tiles * 128 or out_f32.stride(1) != 1:
            raise AonError("gemm_nt: out_f32 must hold m_tiles*128 rows")
        g.out_f32, g.ldc, g.n_valid = out_f32.data_ptr(), out_f32.stride(0), n_valid or N
    if out is not None:
        if out.m_tiles != m_tiles or out_off + N > out.feat:
            raise AonError("gemm_nt: packed output out of range")
        g.out_hi, g.out_lo, g.out_feat, g.out_off = out.hi.data_ptr(), _p(out.lo), out.feat, out_off
        if relu and epi == EPI_LINEAR and out_off == 0 and out.feat == N and N % 32 == 0:
            out.bits = torch.empty(m_tiles * 128, N // 32, dtype=torch.int32, device=device)
            g.relu_bits_out = out.bits.data_ptr()
    cs = None
    if colsum:
        with _on(device):
            rows = lib.aon_gemm_colsum_rows(C.byref(g))        # one partial row per tile, or per CTA of the persistent kernel
        cs = torch.empty(rows, N, dtype=torch.float32, device=device)
        g.colsum = cs.data_ptr()
    with _on(device):
        _check(lib.aon_gemm_tc(C.byref(g), _stream()), "aon_gemm_tc(NT)")
    return cs


def gemm_tn(A: PK, a_off: int, a_tiles: int, B: PK, b_off: int, N: int, splits: int, x3: bool = True) -> torch.Tensor:
    """partial[split][a_tiles*128][N] = sum over the split's rows of A[:, a_off + 128 t + i] * B[:, b_off + j]."""
    lib = load()
    if A.m_tiles != B.m_tiles or a_off + a_tiles * 128 > A.feat or b_off + N > B.feat:
        raise AonError("gemm_tn: operands out of range")
    splits = max(1, min(splits, A.m_tiles))
    tps = (A.m_tiles + splits - 1) // splits
    splits = (A.m_tiles + tps - 1) // tps
    part = torch.empty(splits, a_tiles * 128, N, dtype=torch.float32, device=A.hi.device)
    g = AonGemm()
    g.mode, g.epi, g.x3, g.nseg, g.N, g.m_tiles = GEMM_TN, EPI_PARTIAL, int(x3), 1, N, A.m_tiles
    g.a_hi[0], g.a_lo[0], g.b_hi[0], g.b_lo[0] = A.hi.data_ptr(), _p(A.lo), B.hi.data_ptr(), _p(B.lo)
    g.a_feat[0], g.a_off[0], g.b_feat[0], g.b_off[0] = A.feat, a_off, B.feat, b_off
    g.a_tiles, g.splits, g.tiles_per_split = a_tiles, splits, tps
    g.partial, g.inv_scale = part.data_ptr(), 1.0
    with _on(A.hi.device):
        _check(lib.aon_gemm_tc(C.byref(g), _stream()), "aon_gemm_tc(TN)")
    return part


def pack_rows_tiled(src: torch.Tensor, R: int, S: int, c_pad: int, scale: float, per_ray: bool = False, x3: bool = True) -> PK:
    """src fp32 [R*S, C] (or [R, C] with per_ray) -> PK in the tile order of forward_train (tile rt*S + s, row = ray % 128);
    rows of rays >= R are zero."""
    lib = load()
    if not (src.is_cuda and src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1):
        raise AonError("pack_rows_tiled: src must be a CUDA float32 [rows, C] tensor with contiguous rows")
    if src.shape[0] != (R if per_ray else R * S):
        raise AonError("pack_rows_tiled: src has %d rows, expected %d" % (src.shape[0], R if per_ray else R * S))
    out = PK(lib.aon_train_tiles(R, S), c_pad, src.device, x3)
    with _on(src.device):
        _check(lib.aon_pack_rows_tiled(src.data_ptr(), src.stride(0), src.shape[1], R, S, int(per_ray), c_pad, float(scale),
                                       out.hi.data_ptr(), _p(out.lo), _stream()), "aon_pack_rows_tiled")
    return out


def unpack_rows_tiled(src: torch.Tensor, R: int, S: int) -> torch.Tensor:
    """fp32 [tiles*128, C] in tile order -> [R*S, C] ray-major."""
    lib = load()
    out = torch.empty(R * S, src.shape[1], dtype=torch.float32, device=src.device)
    with _on(src.device):
        _check(lib.aon_unpack_rows_tiled(_ptr(src, "src"), src.stride(0), src.shape[1], R, S, _ptr(out), _stream()), "aon_unpack_rows_tiled")
    return out


class AonTrainDump(C.Structure):
    """Mirror of `struct AonTrainDump` in include/aon.h."""
    _fields_ = [("act_hi", _vp * 28), ("act_lo", _vp * 28), ("relu_bits", _vp * 28), ("enc_hi", _vp), ("enc_lo", _vp),
                ("raw", _vp), ("warped", _vp)]


def forward_train(kind: int, precision: int, packed: torch.Tensor, folded: Optional[torch.Tensor], rays_o, rays_d, viewdirs,
                  t_vals: torch.Tensor, S: int):
    """The forward chain of one level in ONE launch of the fused kernel (aon_forward_train): returns (acts, enc, raw, warped)
    -- acts[i] = PK plane of GEMM unit i's output (+ .bits for the ReLU layers), enc = PK(tiles, 64) encoding operand,
    raw [R*S,4] ray-major, warped [tiles*128,3] (auto-decoder) or None."""
    lib = load()
    R, dev = rays_o.shape[0], rays_o.device
    x3 = precision == PREC_TC_F16X3
    tiles = lib.aon_train_tiles(R, S)
    d = AonTrainDump()
    acts = []
    for i, (n_out, relu) in enumerate(UNIT_OUT[kind]):
        pk = PK(tiles, n_out, dev, x3)
        if relu:
            pk.bits = torch.empty(tiles * 128, n_out // 32, dtype=torch.int32, device=dev)
            d.relu_bits[i] = pk.bits.data_ptr()
        d.act_hi[i], d.act_lo[i] = pk.hi.data_ptr(), _p(pk.lo)
        acts.append(pk)
    enc = PK(tiles, 64, dev, x3)
    raw = torch.empty(R * S, 4, dtype=torch.float32, device=dev)
    warped = torch.empty(tiles * 128, 3, dtype=torch.float32, device=dev) if kind == KIND_AUTODECODER else None
    d.enc_hi, d.enc_lo, d.raw, d.warped = enc.hi.data_ptr(), _p(enc.lo), raw.data_ptr(), _p(warped)
    t_stride = 0 if t_vals.dim() == 1 else t_vals.stride(0)
    with _on(dev):
        _check(lib.aon_forward_train(kind, precision, packed.data_ptr(), _ptr(folded, "folded"), _ptr(rays_o, "rays_o"),
                                     _ptr(rays_d, "rays_d"), _ptr(viewdirs, "viewdirs"), _ptr(t_vals, "t_vals"), t_stride, R, S,
                                     C.byref(d), _stream()), "aon_forward_train")
    return acts, enc, raw, warped


# (out features, relu) of every GEMM unit of the fused kernel, in unit order (csrc/aon_spec.h V_GEMM / A_GEMM)
UNIT_OUT = {KIND_VANILLA: [(256, True)] * 8 + [(256, False), (128, True)],
            KIND_AUTODECODER: [(128, True)] * 4 + [(256, True)] * 8 + [(256, False)] + [(128, True)] * 4}


def wgrad_reduce(partial: torch.Tensor, scale: float, dst: torch.Tensor, col_off: int, rows_valid: int, cols_valid: int,
                 transpose: bool = False, accumulate: bool = False) -> None:
    lib = load()
    splits, rows_pad, N = partial.shape
    with _on(dst.device):
        _check(lib.aon_wgrad_reduce(partial.data_ptr(), splits, rows_pad, N, float(scale), _ptr(dst, "dst"), dst.stride(0), col_off,
                                    rows_valid, cols_valid, int(transpose) | (2 if accumulate else 0), _stream()), "aon_wgrad_reduce")


def colsum_finish(partial: torch.Tensor, scale: float) -> torch.Tensor:
    """partial column sums [rows, N] of a dgrad GEMM (gemm_nt(colsum=True)) -> bias gradient [N] = scale * sum over rows."""
    lib = load()
    rows, N = partial.shape
    out = torch.empty(N, dtype=torch.float32, device=partial.device)
    with _on(partial.device):
        _check(lib.aon_colsum_finish(_ptr(partial, "partial"), rows, N, float(scale), _ptr(out), _stream()), "aon_colsum_finish")
    return out


def colsum_packed(x: PK, splits: int = 16) -> torch.Tensor:
    """[splits, feat] partial column sums (hi + lo) of a PK tensor."""
    lib = load()
    splits = max(1, min(splits, x.m_tiles))
    part = torch.empty(splits, x.feat, dtype=torch.float32, device=x.hi.device)
    with _on(x.hi.device):
        _check(lib.aon_colsum_packed(x.hi.data_ptr(), _p(x.lo), x.feat, x.m_tiles, splits, part.data_ptr(), _stream()),
               "aon_colsum_packed")
    return part


# ---- debug hooks (exported by the library but deliberately not part of include/aon.h) ------------------
def debug_program_info(kind: int, precision: int) -> dict:
    lib = load()
    lib.aon_debug_program_info.restype = C.c_int
    lib.aon_debug_program_info.argtypes = [_i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib.aon_debug_program_info(kind, precision, C.byref(a), C.byref(b), C.byref(c))
    return {"n_units": a.value, "n_stages": b.value, "smem_bytes": c.value}


def debug_set_buffers(dbg: Optional[torch.Tensor], err: Optional[torch.Tensor]) -> None:
    """dbg: CUDA float32 [n_units,128,256] receiving the pre-activation outputs of every unit for ray
    tile 0 / sample 0 of subsequent tensor-core render calls; err: CUDA int32 [1] receiving a code
    if a pipeline barrier times out.  Pass None, None to switch both off."""
    _dbg["dbg"], _dbg["err"] = dbg, err


def debug_set_timeline(tl: Optional[torch.Tensor]) -> None:
    """tl: CUDA int64 [3,4,18,4] receiving SM-clock timestamps of pipeline events of CTA 0 (roles: MMA
    issuer, epilogue warp 0, encoder warp 0; first 4 samples; per unit; 4 events)."""
    _dbg["tl"] = tl


def debug_force_segments(n: int) -> None:
    """n > 0: every tensor-core render_level call cuts each ray's sample range into n segments (one CTA pair per ray tile
    and segment, per-sample (alpha, rgb) composited in order by a second kernel); 0: automatic choice."""
    _dbg["force_segments"] = int(n)


def debug_no_tail_split(on: bool) -> None:
    """True: large ray batches are rendered by ONE unsplit launch (the last, partly filled wave of ray tiles does not get
    its own sample-segmented pass).  For A/B timing and parity tests only."""
    _dbg["no_tail_split"] = 1 if on else 0


def debug_no_fuse(on: bool) -> None:
    """True: render_rays / render_image take the three-launch path (coarse level, sample_pdf, fine level) for every ray."""
    _dbg["no_fuse"] = 1 if on else 0
