"""ctypes binding of libinstageo_b200.so (the C ABI of include/instageo_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Build it with ``python instageo-e2e-geospatial-ml_b200/build.py`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# INSTAGEO_B200_LIB: developer override used by the timing-ablation tooling (tools/attn_ablate.sh)
LIB_PATH = os.environ.get("INSTAGEO_B200_LIB") or os.path.join(HERE, "libinstageo_b200.so")

IG_F32, IG_BF16, IG_I16, IG_U16, IG_F64 = 0, 1, 2, 3, 4
IG_I64, IG_I32, IG_U8, IG_I8 = 5, 6, 7, 8
IG_MASK_EACH, IG_MASK_ANY = 0, 1
ERRORS = {-1: "IG_EINVAL", -2: "IG_ESHAPE", -3: "IG_ECUDA", -4: "IG_ENOMEM", -5: "IG_EARCH", -6: "IG_ESTATE"}


class IgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class ModelCfg(C.Structure):
    _fields_ = [("embed_dim", C.c_int), ("depth", C.c_int), ("num_heads", C.c_int), ("temporal", C.c_int),
                ("num_classes", C.c_int), ("img_size", C.c_int), ("patch_size", C.c_int), ("in_chans", C.c_int),
                ("head_dims", C.c_int * 5)]


# name -> (restype, argtypes); every symbol include/instageo_b200.h declares
_P, _I, _I64, _D, _SZ, _U32, _F = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t, C.c_uint32, C.c_float
SIGNATURES = {
    "ig_version": (_I, []),
    "ig_last_error": (C.c_char_p, []),
    "ig_preprocess": (_I, [_P, _I, _I, _I, _I, _I, _I64, _I64, _I64, _P, _I, _I, _P, _I, _I, _D, _P, _P, _I, _D,
                           _P, _U32, _I, _P, _P, _P, _P, _P]),
    "ig_nodata_map": (_I, [_P, _I, _I, _I, _I, _I64, _I64, _P, _I, _I, _D, _I, _D, _P, _U32, _I, _I, _I, _P, _P]),
    "ig_stitch": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P]),
    "ig_model_create": (_I, [C.POINTER(ModelCfg), C.POINTER(_P)]),
    "ig_model_load_weight": (_I, [_P, C.c_char_p, _P, C.POINTER(_I64), _I, _P]),
    "ig_model_finalize": (_I, [_P, _P]),
    "ig_model_workspace_bytes": (_SZ, [_P, _I]),
    "ig_model_forward": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "ig_model_predict_proba": (_I, [_P, _P, _I, _I, _P, _P, _SZ, _P]),
    "ig_model_launches_per_forward": (_I, [_P]),
    "ig_model_debug_tap": (_I, [_P, C.c_char_p, _I, _P, _P, _SZ, _P]),
    "ig_model_set_tap_buffer": (_I, [_P, _P, _SZ]),
    "ig_model_reset_cache": (_I, [_P]),
    "ig_model_graph_status": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.c_char_p, _SZ]),
    "ig_model_destroy": (_I, [_P]),
    "ig_linear": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ig_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "ig_attention": (_I, [_P, _P, _I, _I, _I, _P]),
    "ig_tiff_unpack16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "ig_tiff_predict": (_I, [_P, _P, _I, _I64, _I, _P]),
    "ig_chip_mask": (_I, [_P, _I, _I, _I64, _I64, _P, _I, _U32, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P]),
    "ig_confusion_update": (_I, [_P, _P, _I, _I64, _I, _I, _I64, _P, _P, _P]),
    "ig_seg_metrics_update": (_I, [_P, _I64, _I, _I64, _P, _I, _I, _I64, _P, _P, _I, _F, _F, _P, _P, _P]),
    "ig_auc_update": (_I, [_P, _I, _I64, _I, _P, _I, _I, _D, _D, _P, _P, _P]),
    "ig_regression_update": (_I, [_P, _P, _I64, _I, _F, _F, _F, _P, _P, _P]),
    "ig_profile_enable": (_I, [_I]),
    "ig_profile_report": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_int), _I]),
    "ig_profile_report_launches": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_int), _I]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and declare every prototype.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python instageo-e2e-geospatial-ml_b200/build.py`; there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise IgError(rc, load().ig_last_error().decode("utf-8", "replace"))


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream(device=None) -> int:
    """cudaStream_t of torch's current stream ON ``device`` (a tensor's device, not the current device: launching
    device-1 pointers on device 0's stream is an illegal access)."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def on_device(t):
    """Context manager: make ``t``'s device current for the duration of a C-ABI call (the library launches on the
    current device)."""
    import torch

    return torch.cuda.device(t.device)


def call(name: str, device, *args) -> None:
    """``ig_<name>(*args, stream)`` with ``device`` current and torch's current stream on it."""
    import torch

    with torch.cuda.device(device):
        check(getattr(load(), name)(*args, current_stream(device)))


PROF_FAMILIES = ["preprocess", "stitch", "gemm_linear", "gemm_conv", "attention", "layernorm", "other"]


def profile_enable(on: bool) -> None:
    check(load().ig_profile_enable(int(on)))


def profile_launches(cap: int = 4096) -> list:
    """[(family, milliseconds)] of every launch since the last report, in launch order."""
    ms, cat = (C.c_double * cap)(), (C.c_int * cap)()
    n = load().ig_profile_report_launches(ms, cat, cap)
    if n < 0:
        check(n)
    return [(PROF_FAMILIES[cat[i]], ms[i]) for i in range(min(n, cap))]


def profile_report() -> dict:
    """{family: (milliseconds, launches)} since the last report."""
    n = len(PROF_FAMILIES)
    ms, cnt = (C.c_double * n)(), (C.c_int * n)()
    check(load().ig_profile_report(ms, cnt, n))
    return {f: (ms[i], cnt[i]) for i, f in enumerate(PROF_FAMILIES)}
