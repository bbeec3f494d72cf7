"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol that
include/afsl.h declares, the ctypes table matches the header, and the product fails loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "afsl.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long|const char\s*\*)\s+(afsl_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if args == ["void"] else args
    return out


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    from afsl_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    decl = declared_functions()
    assert len(decl) >= 15
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for name in decl:
        assert hasattr(cdll, name), f"{name} declared in include/afsl.h but not exported by libafsl.so"
    assert cdll.afsl_version() == 1


def test_ctypes_table_matches_header(lib):
    decl = declared_functions()
    bound = set(lib.SIGNATURES) | {"afsl_version", "afsl_last_error", "afsl_launch_count",
                                   "afsl_view_fusion_weight_floats", "afsl_view_fusion_param_floats",
                                   "afsl_stage1_channels", "afsl_stage1_acc_slots", "afsl_gbn_nhwc_parts",
                                   "afsl_linear_bwd_workspace_floats", "afsl_cpl_saved_supported"}
    assert bound == set(decl), (bound ^ set(decl))
    for name, argtypes in lib.SIGNATURES.items():
        args = decl[name]
        assert len(argtypes) == len(args), name
        for ct, text in zip(argtypes, args):
            if "*" in text:
                assert ct is ctypes.c_void_p, (name, text)
            elif text.startswith("double"):
                assert ct is ctypes.c_double, (name, text)
            elif text.startswith("float"):
                assert ct is ctypes.c_float, (name, text)
            else:
                assert ct is ctypes.c_int, (name, text)


def test_argument_validation_without_gpu(lib):
    cdll = lib.load()
    # null pointers / bad sizes are rejected before any CUDA call
    rc = cdll.afsl_prototypes_fwd_f32(None, None, None, 1, 5, 5, 64, None)
    assert rc == 1 and b"null" in cdll.afsl_last_error()
    rc = cdll.afsl_eval_vote_i32(None, None, None, None, None, 7, None, None, 1, None)
    assert rc == 1


def test_no_cpu_fallback(lib):
    import afsl_b200.ops as ops
    x = torch.randn(25, 64)
    y = torch.arange(5).repeat_interleave(5)
    with pytest.raises(lib.AfslError):
        ops.prototypes(x, y)


def test_embedding_dimension_mismatch_is_refused(lib):
    """Fused 4-view support features (D=256) against single-view query features (D=64) must raise, not read
    out of bounds (the kernels take one D for both operands)."""
    import afsl_b200.ops as ops
    s, q = torch.randn(2, 25, 256), torch.randn(2, 25, 64)
    y = torch.arange(5).repeat_interleave(5).expand(2, -1)
    with pytest.raises(ValueError, match="embedding dimensions differ"):
        ops.proto_head(s, y, q, y, n_way=5)
    with pytest.raises(ValueError, match="embedding dimensions differ"):
        ops.proto_eval(s, y, q, y, n_way=5)
    with pytest.raises(ValueError, match="embedding dimensions differ"):
        ops.cpl_loss(torch.randn(2, 5, 256), q, y, 1.0)
