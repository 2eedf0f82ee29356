"""ctypes binding of libmedtok_vq.so (include/medtok_vq.h).

There is exactly one compute backend: the sm_100a CUDA library built in-tree by
`python -m medtok_b200.build`.  If it is missing or cannot be loaded, `lib()` raises -- nothing in
this package falls back to PyTorch or CPU code.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MEDTOK_VQ_LIB") or os.path.join(_HERE, "libmedtok_vq.so")   # override: A/B timing of two builds

NORM_CONTIGUOUS = 0
NORM_STRIDED = 1
NORM_NONE = 2
SCORE_DOT = 4
F_TRAINING = 1
F_UPDATE = 2
F_SKIP_FINALIZE = 4
F_ACCUMULATE = 8

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_sz = ctypes.c_size_t

_SIGNATURES = {
    "mtvq_abi_version": (_int, []),
    "mtvq_last_error": (ctypes.c_char_p, []),
    "mtvq_workspace_bytes": (_sz, [_i64, _int, _int, _int]),
    "mtvq_stats_len": (_sz, [_int, _int]),
    "mtvq_hard_forward": (_int, [_vp, _i64, _int, _int, _vp, _int, _vp, _dbl, _dbl, ctypes.c_uint,
                                 _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mtvq_ema_finalize": (_int, [_vp, _vp, _vp, _int, _int, _dbl, _dbl, _int, _int, _vp, _vp]),
    "mtvq_ema_finalize_peers": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, ctypes.c_uint, _int, _int, _dbl, _dbl, _int,
                                       _int, _vp, _vp]),
    "mtvq_hard_backward": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _dbl, _vp, _vp]),
    "mtvq_kmeans_update": (_int, [_vp, _vp, _int, _int, _vp]),
    "mtvq_hard_assign": (_int, [_vp, _i64, _int, _int, _vp, _int, _vp, _vp, _vp, _sz, _vp]),
    "mtvq_hard_assign_exact": (_int, [_vp, _i64, _int, _int, _vp, _int, _vp, _vp, _vp, _sz, _vp]),
    "mtvq_gather_ste": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "mtvq_code_sums": (_int, [_vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _sz, _vp]),
    "mtvq_lookup": (_int, [_vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mtvq_codebook_usage": (_int, [_vp, _vp, _i64, _int, _int, _vp, _vp, _sz, _vp, _vp]),
    "mtvq_soft_forward": (_int, [_vp, _i64, _int, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mtvq_soft_backward": (_int, [_vp, _i64, _int, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mtvq_soft_backward_fused": (_int, [_vp, _i64, _int, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _dbl, _int, _vp, _vp,
                                        _vp, _sz, _vp]),
    "mtvq_debug_counters": (_int, [_vp, _i64, _int, _int, _int, ctypes.POINTER(_i64), _vp]),
    "mtvq_debug_coarse_scores": (_int, [_vp, _i64, _int, _int, _vp, _int, _vp, _i64, ctypes.POINTER(_int),
                                        _vp, _sz, _vp]),
    "mtvq_debug_set_events": (_int, [_vp, _vp]),
    "mtvq_debug_set_trace": (_int, [_vp]),
    "mtvq_debug_set_soft_route": (_int, [_int]),
    "mtvq_debug_set_hard_route": (_int, [_int]),
    "mtvq_hard_route": (_int, [_i64, _int, _int]),
    "mtvq_debug_get_hard_route": (_int, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
# diagnostics exported only by the `probe` build (medtok_b200/build.py --variant probe -DMTVQ_WITH_MICROBENCH)
_PROBE_SIGNATURES = {"mtvq_microbench": (_int, [_int, _int, _int, ctypes.POINTER(_dbl), _vp])}

_LIB = None


class MedTokVQError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the C-ABI library; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise MedTokVQError(
                f"{SO_PATH} not found: build the CUDA library first (python -m medtok_b200.build). "
                "medtok_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        for name, (res, args) in _PROBE_SIGNATURES.items():
            if hasattr(handle, name):
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
        if handle.mtvq_abi_version() != 2:
            raise MedTokVQError("libmedtok_vq.so ABI version mismatch")
        _LIB = handle
    return _LIB


def on_tensor_device(fn):
    """Decorator: the library launches on the CURRENT device (cudaGetDevice) while pointers and streams come from the
    tensors, so make the device of the first CUDA tensor argument current around the call -- a module moved with
    .to('cuda:1') then works whatever the process's current device is, like the PyTorch reference."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapped


def check_embedding_dim(d: int, who: str) -> None:
    """The tcgen05 operand layout needs D % 16 == 0 and D <= 256 (include/medtok_vq.h); say so when the module is built,
    not at the first forward.  MedTok's shipped configuration uses 64; its unused default `codebook_embed_dim=8` does not fit."""
    if d <= 0 or d % 16 != 0 or d > 256:
        raise MedTokVQError(f"{who}: embedding dim {d} is not supported by libmedtok_vq (needs a multiple of 16, at most 256); "
                            "pad the projection to the next multiple of 16")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mtvq_last_error().decode(errors="replace")
        raise MedTokVQError(f"{what} failed (code {rc}): {msg}")


def ptr(t: torch.Tensor | None):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise MedTokVQError(
            f"{name} lives on {t.device}: the medtok_b200 quantizer runs only on a CUDA (sm_100a) device; "
            "there is no CPU fallback.")


class Workspace:
    """Caller-owned scratch for the library, cached per module and grown on demand."""

    def __init__(self):
        self.buf: torch.Tensor | None = None

    def get(self, n: int, d: int, k: int, topk: int, device) -> torch.Tensor:
        need = int(lib().mtvq_workspace_bytes(n, d, k, topk))
        if self.buf is None or self.buf.device != torch.device(device) or self.buf.numel() < need:
            self.buf = torch.empty(need, dtype=torch.uint8, device=device)
        return self.buf


def stats_len(d: int, k: int) -> int:
    return int(lib().mtvq_stats_len(d, k))


def set_profile_events(before: "torch.cuda.Event | None", after: "torch.cuda.Event | None") -> None:
    """One-shot: the next hard forward on this thread records `before`/`after` around its dominant kernel."""
    b = ctypes.c_void_p(before.cuda_event) if before is not None else None
    a = ctypes.c_void_p(after.cuda_event) if after is not None else None
    check(lib().mtvq_debug_set_events(b, a), "mtvq_debug_set_events")


def debug_counters(ws: torch.Tensor, n: int, d: int, k: int, topk: int = 0) -> dict:
    out = (_i64 * 4)()
    check(lib().mtvq_debug_counters(ptr(ws), n, d, k, topk, out, stream_ptr(ws.device)), "mtvq_debug_counters")
    return {"near_tie_rows": out[0], "full_scan_rows": out[1], "candidates": out[2], "launches": out[3]}
