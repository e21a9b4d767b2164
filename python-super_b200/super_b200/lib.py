"""ctypes binding of the C-ABI CUDA library (include/super_b200.h -> libsuper_b200.so).

There is NO fallback: if the library is missing or a call returns non-zero, this raises.  PyTorch
is used only to own device memory and to supply the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsuper_b200.so")

c_int, c_double, c_void_p = ctypes.c_int, ctypes.c_double, ctypes.c_void_p

HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "super_b200.h")
_CT = {"p": c_void_p, "i": c_int, "d": c_double, "f": ctypes.c_float, "l": ctypes.c_longlong}


_RESTYPE = {}


def parse_header(path=HEADER_PATH):
    """{function name: argument kinds} from the C header -- the header is the single source of truth
    for the ABI ('p' pointer, 'i' int, 'd' double)."""
    import re
    src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    if os.environ.get("SB_DEBUG_EXPORTS") != "1":          # timing-experiment entry points: only in a debug build
        src = re.sub(r"#ifdef SB_DEBUG_EXPORTS.*?#endif", "", src, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\b(int|long long)\s+(sb_\w+)\s*\(([^)]*)\)\s*;", src):
        args = [a.strip() for a in m.group(3).split(",") if a.strip() and a.strip() != "void"]
        kinds = ""
        for a in args:
            if "*" in a:
                kinds += "p"
            elif a.startswith("double"):
                kinds += "d"
            elif a.startswith("float"):
                kinds += "f"
            elif a.startswith("long long"):
                kinds += "l"
            else:
                kinds += "i"
        sigs[m.group(2)] = kinds
        _RESTYPE[m.group(2)] = ctypes.c_longlong if m.group(1) == "long long" else c_int
    return sigs


_SIGNATURES = parse_header()

_lib = None


class SuperB200Error(RuntimeError):
    pass


def load():
    """Load libsuper_b200.so (built by super_b200.build).  Raises if absent: no CPU / torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SuperB200Error(
            f"{LIB_PATH} not found: build it with `python -m super_b200.build` (nvcc, sm_100a). "
            "super_b200 has no CPU or eager-PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, sig in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.argtypes = [_CT[c] for c in sig]
        fn.restype = _RESTYPE.get(name, c_int)
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES)


_ERR = {1: "invalid argument", 2: "CUDA launch error", 3: "workspace size mismatch"}


def ptr(t):
    """Raw device (or host) pointer of a contiguous tensor; None -> NULL."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise SuperB200Error("non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


_SCOPE_STREAM = None      # inside graph_scope(): the capture stream every call of the scope is issued on


def stream():
    if _SCOPE_STREAM is not None:
        return _SCOPE_STREAM
    return torch.cuda.current_stream().cuda_stream


class graph_scope:
    """with graph_scope(handle): ...C-ABI calls only...  -- the calls are stream-captured and replayed as ONE CUDA graph on
    the current stream (sb_graph_scope_begin/_end; handle = ctypes.c_void_p owned by the caller).  No torch device work, no
    synchronisation inside.  handle None: plain execution."""

    def __init__(self, handle):
        self.handle = handle

    def __enter__(self):
        global _SCOPE_STREAM
        if self.handle is None:
            return self
        self.real = torch.cuda.current_stream().cuda_stream
        use = ctypes.c_void_p()
        rc = load().sb_graph_scope_begin(ctypes.byref(self.handle), self.real, ctypes.byref(use))
        if rc != 0:
            raise SuperB200Error(f"sb_graph_scope_begin failed: {_ERR.get(rc, rc)}")
        _SCOPE_STREAM = use.value or 0
        return self

    def __exit__(self, et, ev, tb):
        global _SCOPE_STREAM
        if self.handle is None:
            return False
        _SCOPE_STREAM = None
        rc = load().sb_graph_scope_end(ctypes.byref(self.handle), self.real, 1 if et is not None else 0)
        if rc != 0 and et is None:
            raise SuperB200Error(f"sb_graph_scope_end failed: {_ERR.get(rc, rc)}")
        return False


# kernels of OURS launched per C-ABI call (library kernels such as the CUB scans are not counted)
KERNELS_PER_CALL = {"sb_knn": 1, "sb_knn_class": 1, "sb_knn_weights": 1, "sb_reweight": 1, "sb_warp_update": 2, "sb_tuple_keys": 1, "sb_tuple_order": 1,
                    "sb_data_term_jtj": 1, "sb_data_term_loss": 1, "sb_data_term_loss_decide": 1, "sb_data_term_rows": 1, "sb_lm_begin": 1,
                    "sb_reg_terms": 1, "sb_lm_damp": 1, "sb_lm_step": 1, "sb_lm_decide": 1, "sb_lm_decide_reg": 1, "sb_preprocess": 2,
                    "sb_fuse": 6, "sb_compact": 2, "sb_band_solve3": 1, "sb_band_from_fixed": 1, "sb_band_solve4": 5, "sb_band_solve4_step": 5, "sb_band_solve4_step_fx": 5, "sb_gather_sorted": 1, "sb_graph_build": 1, "sb_gf_data": 1, "sb_gf_morph": 1, "sb_gf_reg": 1,
                    "sb_gf_step": 2, "sb_gf_global_update": 1, "sb_track_points": 1, "sb_seg_maps": 1,
                    "sb_reweight_semantic": 1}
LAUNCHES = 0            # running count (bench.py resets it around the timed region)
KERNEL_EVENTS = None    # bench.py: list receiving (start, end) CUDA events around each sb_data_term_jtj launch


def call(name, *args):
    global LAUNCHES
    lib = load()
    ev = None
    sink = None
    if KERNEL_EVENTS is not None:
        # bench.py: either a list (events of sb_data_term_jtj) or a dict {entry point: list}
        sink = KERNEL_EVENTS.get(name) if isinstance(KERNEL_EVENTS, dict) else (KERNEL_EVENTS if name == "sb_data_term_jtj" else None)
    if sink is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    rc = getattr(lib, name)(*args)
    if ev is not None:
        ev[1].record()
        sink.append(ev)
    if rc != 0:
        raise SuperB200Error(f"{name} failed: {_ERR.get(rc, rc)}")
    LAUNCHES += KERNELS_PER_CALL.get(name, 0)


def intr_array(fx, fy, cx, cy):
    return (c_double * 4)(float(fx), float(fy), float(cx), float(cy))
