"""CPU-side checks: the C-ABI library loads and exports every symbol include/medtok_vq.h declares,
argument validation works without a GPU, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from medtok_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "medtok_vq.h")).read()
    declared = sorted(set(re.findall(r"MTVQ_API[^;(]*?\b(mtvq_\w+)\s*\(", hdr)))
    assert len(declared) >= 15
    from medtok_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in medtok_vq.h but not exported"


def test_host_side_sizes_and_validation(lib):
    assert lib.mtvq_abi_version() == 2
    assert lib.mtvq_stats_len(64, 8192) == 8192 + 8192 * 64 + 2
    small, big = lib.mtvq_workspace_bytes(1000, 64, 512, 0), lib.mtvq_workspace_bytes(1 << 20, 64, 8192, 0)
    assert 0 < small < big < (1 << 30)
    # topk >= 1 adds the soft path's buffers -- for top-1 too (VectorQuantizer(..., k=1) is a legal configuration)
    assert lib.mtvq_workspace_bytes(1000, 64, 512, 5) > small
    assert lib.mtvq_workspace_bytes(1000, 64, 512, 1) > small + (80 << 20)        # record scratch of the top-k pass
    # D not a multiple of 16 is rejected on the host before any CUDA call
    rc = lib.mtvq_hard_assign(None, 10, 20, 0, None, 16, None, None, None, 0, None)
    assert rc == -2 and b"multiple of 16" in lib.mtvq_last_error()
    rc = lib.mtvq_hard_assign(None, 10, 64, 0, None, 16, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.mtvq_last_error()


def test_no_cpu_fallback():
    import torch
    from medtok_b200 import NormEMAVectorQuantizer, VectorQuantizer, _lib
    q = NormEMAVectorQuantizer(64, 32, 0.25)
    with pytest.raises(_lib.MedTokVQError):
        q(torch.randn(4, 32, 1, 1))
    s = VectorQuantizer(96, 32, 0.25, 0.0, True, True, [32, 32])
    with pytest.raises(_lib.MedTokVQError):
        s.specific_embedding(torch.randn(4, 32), types="text")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "medtok_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                # comments may cite the oracle; importing, loading or including it is what is forbidden
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", txt, flags=re.M), f
                assert "libvq_oracle" not in txt and not re.search(r"#include\s+[\"<].*oracle", txt), f


def test_state_dict_keys_match_reference():
    from conftest import load_golden
    from medtok_b200 import NormEMAVectorQuantizer, VectorQuantizer
    meta, _ = load_golden("hard_train_n512_d64_k1024")
    assert sorted(NormEMAVectorQuantizer(1024, 64, 0.25).state_dict().keys()) == meta["state_dict_keys"]
    meta, _ = load_golden("soft_specific_ne3072_d64_n300")
    sq = VectorQuantizer(3072, 64, 0.25, 0.0, True, True, [64, 64])
    assert sorted(sq.state_dict().keys()) == meta["state_dict_keys"]


def test_every_entry_point_validates_its_arguments_on_the_host(lib):
    """Null pointers / out-of-range scalars are rejected before anything is launched: a non-zero return code and a
    message from mtvq_last_error(), never a crash and never a CUDA call (this box has no GPU).  ABI error behaviour:
    include/medtok_vq.h ("functions return 0 or an error code; no exception crosses the boundary")."""
    from medtok_b200 import _lib
    host_only = {"mtvq_abi_version", "mtvq_last_error", "mtvq_workspace_bytes", "mtvq_stats_len", "mtvq_hard_route",
                 "mtvq_debug_set_hard_route", "mtvq_debug_get_hard_route", "mtvq_debug_set_soft_route",
                 "mtvq_debug_set_events", "mtvq_debug_set_trace", "mtvq_debug_coarse_scores"}
    checked = 0
    for name, (_res, args) in _lib._SIGNATURES.items():
        if name in host_only:
            continue
        vals = []
        for a in args:
            if a is ctypes.c_void_p or (hasattr(a, "_type_") and not isinstance(a._type_, str)):
                vals.append(None)                               # every pointer null
            elif a is ctypes.c_double:
                vals.append(0.5)
            else:
                vals.append(64)                                 # d = k = n = topk = 64, norm_mode / flags = 64
        rc = getattr(lib, name)(*vals)
        assert rc != 0, f"{name} accepted null pointers"
        assert len(lib.mtvq_last_error()) > 0, f"{name} failed without a message"
        checked += 1
    assert checked >= 14
