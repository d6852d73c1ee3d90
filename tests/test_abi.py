"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/aon.h declares.
No compute calls (no GPU here); argument validation paths that return before touching CUDA are exercised."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aon.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aon_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    lib = built_lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libaon_b200.so does not export %s" % n
    assert set(names) == set(built_lib.SYMBOLS), "lib.py binding table out of sync with include/aon.h"


def test_introspection(built_lib):
    lib = built_lib.load()
    assert lib.aon_version() == 2
    assert lib.aon_num_layers(0) == 12 and lib.aon_num_layers(1) == 20 and lib.aon_num_layers(7) < 0
    from oracle import ref_cpu as O
    assert built_lib.layer_shapes(0) == [(o, i) for _, o, i in O.VANILLA_LAYERS]
    assert built_lib.layer_shapes(1) == [(o, i) for _, o, i in O.AUTODECODER_LAYERS]
    assert lib.aon_packed_bytes(0, 0) > 4 * 593408 and lib.aon_packed_bytes(1, 0) > 4 * 600000
    assert lib.aon_packed_bytes(0, 99) == 0
    assert lib.aon_folded_floats(0) == 0 and lib.aon_folded_floats(1) == 768 + 288 + 6144   # fp32 biases | latents | per-call bias stages (2 ranks)


def test_argument_errors_return_codes(built_lib):
    lib = built_lib.load()
    rc = lib.aon_render_level(0, 0, None, None, None, None, None, None, 0, 1, 65, 1, None, None, None, None, None, 0, None, None)
    assert rc == -1 and b"null" in lib.aon_last_error()
    rc = lib.aon_sample_pdf(None, 0, None, None, 0, 1, 65, 128, None, None)
    assert rc == -1
    rc = lib.aon_raygen(0, 0, ctypes.c_float(1.0), None, None, None, None)
    assert rc == -1
    rc = lib.aon_render_rays(0, 1, None, None, None, None, None, None, None, None, None, 4, 2.0, 6.0, 1, None, None, None, 0, None, None)
    assert rc == -1 and b"null" in lib.aon_last_error()
    rc = lib.aon_render_image(0, 1, None, None, None, None, None, 100.0, 4, 4, 0, 16, 2.0, 6.0, 1, None, None, None, 0, None, None)
    assert rc == -1 and b"camera" in lib.aon_last_error()


def test_workspace_contract(built_lib):
    """The caller owns all scratch: aon_workspace_bytes is monotone in R, covers the fused kernel's scratch slots in the
    tensor-core modes, and rejects bad precisions; the options struct mirror has the documented size."""
    lib = built_lib.load()
    assert lib.aon_workspace_bytes(99, 10) == 0
    a, b, c = lib.aon_workspace_bytes(1, 1), lib.aon_workspace_bytes(1, 3840), lib.aon_workspace_bytes(1, 307200)
    assert 0 < a <= b <= c
    assert a >= 160 * 128 * (65 + 193) * 4                     # scratch slots of the fused kernel
    assert lib.aon_workspace_bytes(0, 3840) >= 3840 * (65 + 193) * 4
    assert ctypes.sizeof(built_lib.AonRenderOpts) == 40


def test_sass_is_sm100a(built_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_sass_has_the_blackwell_instructions(built_lib):
    """The tensor-core paths really are tcgen05 / TMA code: UTCHMMA (tcgen05.mma kind::f16, incl. the 2-CTA form of the fused
    render kernel), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk) in the SASS of the shipped library."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert sass.count(mnemonic) > 0, mnemonic
    assert "UTCHMMA.2CTA" in sass
    for fn in ("render_tc_kernel", "gemm_tc_kernel", "gemm_tc_nt_persistent_kernel"):
        assert fn in sass, fn


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "articulated-object-nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_gemm_descriptor_mirror_and_training_argument_errors(built_lib):
    """The ctypes mirror of `struct AonGemm` has the C struct's size, and the training entry points reject bad arguments
    before touching CUDA."""
    lib = built_lib.load()
    assert ctypes.sizeof(built_lib.AonGemm) == lib.aon_gemm_struct_size()
    assert lib.aon_gemm_tc(None, None) == -1 and b"null" in lib.aon_last_error()
    g = built_lib.AonGemm()
    g.mode, g.N, g.nseg = 0, 24, 1                     # N must be a multiple of 16
    assert lib.aon_gemm_tc(ctypes.byref(g), None) == -1 and b"multiple of 16" in lib.aon_last_error()
    assert lib.aon_composite(None, None, None, 0, None, 1, 65, 1, 0, None, None, None, None, None, None) == -1
    assert lib.aon_pos_enc(None, 1, 10, None, None) == -1
    assert lib.aon_adam_step(None, None, None, None, 1, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None) == -1
    assert lib.aon_pack_rows(None, 1, 1, 1, 1, 1, 8, ctypes.c_float(1.0), None, None, None) == -1


def test_round2_entry_points_argument_errors_and_mirrors(built_lib):
    """The entry points added in round 2 -- in-kernel random draws, the fused training forward, tiled packing, graph-capturable
    Adam -- reject bad arguments before touching CUDA, their ctypes struct mirrors have the C layout the header documents, and
    aon_train_tiles / aon_adam_scalars (pure host functions) return the documented values."""
    import math
    lib = built_lib.load()
    assert ctypes.sizeof(built_lib.AonRng) == 24
    assert ctypes.sizeof(built_lib.AonTrainDump) == (3 * 28 + 4) * 8
    # row tiles of the fused training forward: 2 * ceil(R / 256) * S
    assert lib.aon_train_tiles(2048, 65) == 16 * 65 and lib.aon_train_tiles(96, 65) == 2 * 65
    assert lib.aon_train_tiles(300, 193) == 4 * 193 and lib.aon_train_tiles(0, 65) == 0
    assert lib.aon_sample_along_rays_rng(2.0, 6.0, 65, None, 4, None, None) == -1 and b"rng" in lib.aon_last_error()
    assert lib.aon_sample_pdf_rng(None, 0, None, None, 4, 65, 128, None, None) == -1
    assert lib.aon_rng_uniform(None, 0, 4, 4, None, None) == -1
    assert lib.aon_rng_advance(None, 1, None) == -1 and b"null" in lib.aon_last_error()
    assert lib.aon_forward_train(0, 1, None, None, None, None, None, None, 0, 4, 65, None, None) == -1
    assert lib.aon_forward_train(7, 1, None, None, None, None, None, None, 0, 4, 65, None, None) == -1 and b"kind" in lib.aon_last_error()
    assert lib.aon_pack_rows_tiled(None, 4, 4, 4, 65, 0, 16, 1.0, None, None, None) == -1
    assert lib.aon_unpack_rows_tiled(None, 4, 4, 4, 65, None, None) == -1
    assert lib.aon_adam_step_dev(None, None, None, None, 10, None, None) == -1
    assert lib.aon_gemm_colsum_rows(None) == 0
    # aon_adam_scalars: exactly the scalars aon_adam_step derives (torch.optim.Adam's single-tensor formulas)
    out = (ctypes.c_float * 7)()
    assert lib.aon_adam_scalars(5e-4, 0.9, 0.999, 1e-8, 3, 0.5, out) == 0
    bc1, bc2 = 1 - 0.9 ** 3, 1 - 0.999 ** 3
    want = [1 - 0.9, 0.999, 1 - 0.999, 1e-8, 5e-4 / bc1, math.sqrt(bc2), 0.5]
    assert all(abs(a - b) <= 1e-6 * abs(b) for a, b in zip(out, want)), (list(out), want)
    assert lib.aon_adam_scalars(5e-4, 0.9, 0.999, 1e-8, 0, 1.0, out) == -1


def test_training_dump_table_matches_the_kernel_program(built_lib):
    """lib.UNIT_OUT (the (out features, relu) list forward_train sizes its activation planes from) has one entry per GEMM unit
    of the fused kernel's program, with the widths of the reference's layer table (aon_layer_shape) in unit order."""
    L = built_lib
    for kind, order in ((L.KIND_VANILLA, [0, 1, 2, 3, 4, 5, 6, 7, 9, 8]),                          # csrc/aon_spec.h V_GEMM[].src
                        (L.KIND_AUTODECODER, [0, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 17, 13, 14, 15, 16])):   # A_GEMM[].src
        info = L.debug_program_info(kind, L.PREC_TC_F16X3)
        assert info["n_units"] == len(L.UNIT_OUT[kind]) == len(order) <= 28
        shapes = L.layer_shapes(kind)
        for (n_out, relu), src in zip(L.UNIT_OUT[kind], order):
            assert shapes[src][0] == n_out, (kind, src, shapes[src], n_out)
        # the only unit without a ReLU is the bottleneck layer (model.py:112; model_autodecoder.py:229)
        assert [r for _, r in L.UNIT_OUT[kind]].count(False) == 1
