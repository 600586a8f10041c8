"""CPU: the C-ABI library builds, loads, and exports every symbol include/swinb200.h declares (no compute calls)."""
import ctypes
import os
import re

from swin_v2_weather_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


HEADERS = ("swinb200.h", "swinb200_debug.h")     # the drop-in boundary, and the bring-up hooks kept out of it


def header_text():
    return "\n".join(open(os.path.join(ROOT, "include", h)).read() for h in HEADERS)


def declared_symbols():
    return sorted(set(re.findall(r"\b(swinb200_[a-z0-9_]+)\s*\(", header_text())))


def test_library_loads_and_exports_header_symbols():
    if not os.path.exists(_lib.LIB_PATH):
        from swin_v2_weather_b200.build import build
        build()
    lib = _lib.load()
    assert lib.swinb200_version() == 100
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in swinb200.h but not exported"
    assert "swinb200_adam_step" in syms and "swinb200_patchify_cat" in syms
    # every prototype bound by the Python side is declared in the header, and vice versa
    bound = set(_lib.PROTOTYPES) | {"swinb200_version", "swinb200_last_error"}
    assert bound == set(syms), bound ^ set(syms)


def test_ctypes_prototypes_have_the_arity_of_the_header():
    """A drifting argument list would shift every later pointer by one slot: count them."""
    text = re.sub(r"/\*.*?\*/", "", header_text(), flags=re.S)
    decls = re.findall(r"\bint\s+(swinb200_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(decls) >= 17
    for name, args in decls:
        n = len([a for a in args.replace("\n", " ").split(",") if a.strip() and a.strip() != "void"])
        if name in _lib.PROTOTYPES:
            assert n == len(_lib.PROTOTYPES[name]), (name, n, len(_lib.PROTOTYPES[name]))


def test_new_entry_points_reject_bad_arguments():
    lib = _lib.load()
    assert lib.swinb200_adam_step(1, None, None, None, None, None, None, 1e-3, 0.9, 0.95, 1e-8, 0.0, 1, None, None, None) == 1
    assert b"null table" in lib.swinb200_last_error()
    arr = (ctypes.c_void_p * 1)(16)
    n = (ctypes.c_longlong * 1)(8)
    assert lib.swinb200_adam_step(1, arr, arr, arr, arr, None, n, 1e-3, 0.9, 0.95, 1e-8, 0.0, 0, None, None, None) == 1
    assert b"step counts from 1" in lib.swinb200_last_error()
    assert lib.swinb200_patchify_cat(9, arr, None, None, None, 1, 1, 8, 8, 4, None) == 1
    assert b"sources" in lib.swinb200_last_error()


def test_linear_ln_residual_rejects_what_it_is_not_built_for():
    """The fused GEMM + LayerNorm entry point: null operands are argument errors; other widths / back ends are refused with
    a message that names the two-kernel path (no silent fallback inside the library)."""
    lib = _lib.load()
    p16 = ctypes.c_void_p(16)
    rc = lib.swinb200_linear_ln_residual(1, 256, 768, 768, None, 768, p16, 768, None, p16, 768, p16, p16, p16, None, p16, p16, p16,
                                         256, 1e-5, p16, 2, None)
    assert rc == 1 and b"null pointer" in lib.swinb200_last_error()
    rc = lib.swinb200_linear_ln_residual(1, 256, 512, 768, p16, 768, p16, 768, None, p16, 512, p16, p16, p16, None, p16, p16, p16,
                                         256, 1e-5, p16, 2, None)
    assert rc == 3 and b"768 output channels" in lib.swinb200_last_error()
    rc = lib.swinb200_linear_ln_residual(0, 256, 768, 768, p16, 768, p16, 768, None, p16, 768, p16, p16, p16, None, p16, p16, p16,
                                         256, 1e-5, p16, 2, None)
    assert rc == 3 and b"swinb200_ln_residual_fwd" in lib.swinb200_last_error()
    rc = lib.swinb200_linear_ln_residual(1, 1024, 768, 768, p16, 768, p16, 768, None, p16, 768, p16, p16, p16, None, p16, p16, p16,
                                         256, 1e-5, p16, 2, None)        # 8 row blocks, 2 counters
    assert rc == 1 and b"one counter per 128-row block" in lib.swinb200_last_error()


def test_attention_back_end_is_chosen_from_the_shape():
    from swin_v2_weather_b200 import ops
    from swin_v2_weather_b200._lib import BACKEND_SIMT, BACKEND_TCGEN05
    for heads in (4, 6, 8, 12, 16):      # head_dim 192, 128, 96, 64, 48 at C = 768: tensor cores, any window
        for window in ((9, 18), (6, 12), (12, 24), (18, 36)):
            assert ops.attn_backend_for(ops.MODE_BF16, 768, heads, *window) == BACKEND_TCGEN05
    assert ops.attn_backend_for(ops.MODE_BF16, 160, 2, 9, 18) == BACKEND_SIMT          # head_dim 80: CUDA cores
    assert ops.attn_backend_for(ops.MODE_FP32, 768, 8, 9, 18) == BACKEND_SIMT          # fp32 validation mode


def test_argument_errors_come_back_as_codes_not_crashes():
    lib = _lib.load()
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.swinb200_cast_f32_to_bf16(None, None, 16, None) == 1
    assert b"null" in lib.swinb200_last_error()
    rc = lib.swinb200_gemm(0, 0, 8, 8, None, 0, 8, None, 0, 8, 0, 0, None, None, 8, None, None, 0, 0, 0, 1, None)
    assert rc == 1 and b"bad shape" in lib.swinb200_last_error()
    rc = lib.swinb200_patchify(ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 1, 3, 72, 144, 8, 0, None)
    assert rc == 1 and b"patch_size 4" in lib.swinb200_last_error()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """`tcgen05.mma` -> UTC*MMA, `tcgen05.ld` -> LDTM, TMA -> UTMALDG in the shipped binary."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass
    assert "LDTM" in sass and "UTMALDG" in sass
    assert "sm_100a" in sass
