"""ctypes binding of libbmb200.so (include/bmb200.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbmb200.so")

i64 = C.c_int64
dbl = C.c_double
vp = C.c_void_p
ch = C.c_char

# name -> (restype, argtypes); mirrors include/bmb200.h one to one
PROTOTYPES = {
    "bmb200_version": (C.c_int, []),
    "bmb200_create": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
    "bmb200_destroy": (C.c_int, [vp]),
    "bmb200_set_stream": (C.c_int, [vp, vp]),
    "bmb200_sync": (C.c_int, [vp]),
    "bmb200_malloc": (C.c_int, [vp, C.POINTER(vp), C.c_size_t]),
    "bmb200_free": (C.c_int, [vp, vp]),
    "bmb200_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "bmb200_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "bmb200_last_error": (C.c_char_p, [vp]),
    "bmb200_launch_count": (i64, [vp]),
    "bmb200_dgbmv": (C.c_int, [vp, ch, i64, i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]),
    "bmb200_dgbmm_bb": (C.c_int, [vp] + [i64] * 9 + [dbl, vp, i64, vp, i64, dbl, vp, i64]),
    "bmb200_dgbmm_bd": (C.c_int, [vp, ch, i64, i64, i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]),
    "bmb200_dgbmm_db": (C.c_int, [vp, ch, i64, i64, i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]),
    "bmb200_dfill_lmul": (C.c_int, [vp, dbl, vp, i64, i64, i64, i64]),
    "bmb200_dband_widen": (C.c_int, [vp, i64, i64, i64, vp, i64, vp, i64]),
    "bmb200_dgbtrf": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, C.POINTER(C.c_int)]),
    "bmb200_dgbtrf_from": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, C.POINTER(C.c_int)]),
    "bmb200_dgbtrs": (C.c_int, [vp, ch, i64, i64, i64, i64, vp, i64, vp, vp, i64]),
    "bmb200_dtbsv": (C.c_int, [vp, ch, ch, ch, i64, i64, vp, i64, vp, i64]),
    "bmb200_dtbmv": (C.c_int, [vp, ch, ch, ch, i64, i64, vp, i64, vp, i64]),
    "bmb200_dsbmv": (C.c_int, [vp, ch, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]),
    "bmb200_dpbtrf": (C.c_int, [vp, ch, i64, i64, vp, i64, C.POINTER(C.c_int)]),
    "bmb200_dpbtrs": (C.c_int, [vp, ch, i64, i64, i64, vp, i64, vp, i64]),
    **{f"bmb200_{p}gbmv": (C.c_int, [vp, ch, i64, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, i64]) for p in "scz"},
    **{f"bmb200_{p}gbtrf": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, C.POINTER(C.c_int)]) for p in "scz"},
    **{f"bmb200_{p}gbtrs": (C.c_int, [vp, ch, i64, i64, i64, i64, vp, i64, vp, vp, i64]) for p in "scz"},
    **{f"bmb200_{p}": (C.c_int, [vp, ch, i64, i64, vp, vp, i64, vp, i64, vp, vp, i64]) for p in ("ssbmv", "chbmv", "zhbmv")},
    **{f"bmb200_{p}tbsv": (C.c_int, [vp, ch, ch, ch, i64, i64, vp, i64, vp, i64]) for p in "scz"},
    **{f"bmb200_{p}tbmv": (C.c_int, [vp, ch, ch, ch, i64, i64, vp, i64, vp, i64]) for p in "scz"},
    **{f"bmb200_{p}pbtrf": (C.c_int, [vp, ch, i64, i64, vp, i64, C.POINTER(C.c_int)]) for p in "scz"},
    **{f"bmb200_{p}pbtrs": (C.c_int, [vp, ch, i64, i64, i64, vp, i64, vp, i64]) for p in "scz"},
    **{f"bmb200_{p}gbmm_bb": (C.c_int, [vp] + [i64] * 9 + [vp, vp, i64, vp, i64, vp, vp, i64]) for p in "scz"},
    **{f"bmb200_{p}gbmm_bd": (C.c_int, [vp, ch, i64, i64, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, i64]) for p in "scz"},
    "bmb200_dband_axpy": (C.c_int, [vp, i64, i64, dbl, i64, i64, vp, i64, i64, i64, vp, i64, C.POINTER(C.c_int64)]),
    "bmb200_dband_copy": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, i64, i64, vp, i64, C.POINTER(C.c_int64)]),
    "bmb200_dband_lmul_block": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, i64, i64, i64, i64, dbl]),
    "bmb200_dband_transpose": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, i64]),
    "bmb200_dband_nonzero_rows": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, C.POINTER(C.c_int)]),
    "bmb200_dband_axpby": (C.c_int, [vp, i64, i64, dbl, i64, i64, vp, i64, dbl, i64, i64, vp, i64, i64, i64, vp, i64]),
    "bmb200_dgbmv_host": (C.c_int, [vp, ch, i64, i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]),
    "bmb200_dgbsv_host": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, vp, i64, C.POINTER(C.c_int)]),
    "bmb200_dgbmm_bb_host": (C.c_int, [vp] + [i64] * 9 + [dbl, vp, i64, vp, i64, dbl, vp, i64]),
    "bmb200_halo_create": (C.c_int, [vp, i64, vp]),
    "bmb200_halo_connect": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "bmb200_halo_destroy": (C.c_int, [vp]),
    "bmb200_dgbmv_sharded": (C.c_int, [vp, i64, i64, i64, i64, i64, dbl, vp, i64, vp, dbl, vp]),
}

# test / tuning hooks of include/bmb200_internal.h (not part of the drop-in ABI)
INTERNAL_PROTOTYPES = {
    "bmb200_internal_divcheck": (C.c_int, [vp, i64, vp, vp, vp]),
    "bmb200_internal_divcheck2": (C.c_int, [vp, i64, vp, vp, vp]),
    "bmb200_internal_dgbtrf_generic": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, C.POINTER(C.c_int)]),
    "bmb200_internal_dgbtrs_generic": (C.c_int, [vp, ch, i64, i64, i64, i64, vp, i64, vp, vp, i64]),
    "bmb200_internal_last_gbmm_path": (C.c_int, [vp]),
    "bmb200_internal_set_tuning": (C.c_int, [vp, C.c_char_p, C.c_longlong]),
    "bmb200_internal_gbtrs_slot": (C.c_int, [vp] + [C.c_int] * 5 + [i64, i64, i64, i64, vp, i64, vp, vp, i64]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libbmb200.so and attach prototypes.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C bandedmatrices.jl_b200/csrc).  There is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in list(PROTOTYPES.items()) + list(INTERNAL_PROTOTYPES.items()):
            fn = getattr(lib, name)  # AttributeError here == header/library drift
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class BMB200Error(RuntimeError):
    pass


class Handle:
    """One handle per (device, stream); owns scratch and the launch counter."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load()
        self.h = vp()
        rc = self.lib.bmb200_create(C.byref(self.h), int(device), vp(stream) if stream else None)
        if rc != 0:
            raise BMB200Error(f"bmb200_create(device={device}) failed with {rc}: no usable sm_100a GPU (no CPU fallback)")
        self.device = int(device)

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.bmb200_last_error(self.h)
            raise BMB200Error(f"{what} failed: rc={rc} {msg.decode() if msg else ''}")

    def set_stream(self, stream_ptr: int) -> None:
        self.check(self.lib.bmb200_set_stream(self.h, vp(stream_ptr) if stream_ptr else None), "set_stream")

    def tune(self, key: str, value: int) -> None:
        """Development knob of this handle (include/bmb200_internal.h); ``tune("reset", 0)`` restores the defaults."""
        self.check(self.lib.bmb200_internal_set_tuning(self.h, key.encode(), int(value)), f"set_tuning({key})")

    def last_gbmm_path(self) -> int:
        """0 sweep, 1 tile DMMA, 2 ring DMMA, 3 K-blocked DMMA (include/bmb200_internal.h)."""
        return int(self.lib.bmb200_internal_last_gbmm_path(self.h))

    def sync(self) -> None:
        self.check(self.lib.bmb200_sync(self.h), "sync")

    @property
    def launches(self) -> int:
        return int(self.lib.bmb200_launch_count(self.h))

    def close(self) -> None:
        if self.h:
            self.lib.bmb200_destroy(self.h)
            self.h = vp()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_handles: dict = {}


def handle(device: int = 0) -> Handle:
    """Handle for (``device``, torch's CURRENT stream): one handle per stream, because a handle owns scratch memory and
    the device-side status block, which two streams must not share without ordering."""
    import torch

    stream = int(torch.cuda.current_stream(device).cuda_stream)
    hd = _handles.get((device, stream))
    if hd is None:
        hd = _handles[(device, stream)] = Handle(device)
        hd.set_stream(stream)
    return hd
