"""City-dataset scans without ROS: thin ctypes wrappers over the library's host readers.

  read_livox_bin / read_ouster_bin    file_player/src/ROSThread.cpp:776-795, :952-967
  preprocess_livox / preprocess_ouster  MA_LIO/src/preprocess.cpp:59-110, :112-152
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def _read(fn, dtype, path: str, eof_quirk: bool):
    lib = capi.load()
    n = C.c_uint32(0)
    rc = getattr(lib, fn)(path.encode(), None, 0, C.byref(n), 1 if eof_quirk else 0)
    if rc != capi.OK:
        raise capi.MalioError(rc, f"{fn}({path})")
    out = np.zeros(n.value, dtype=dtype)
    if n.value:
        rc = getattr(lib, fn)(path.encode(), capi.ptr(out), n.value, C.byref(n), 1 if eof_quirk else 0)
        if rc != capi.OK:
            raise capi.MalioError(rc, f"{fn}({path})")
    return out


def read_livox_bin(path: str, eof_quirk: bool = True) -> np.ndarray:
    return _read("malio_read_livox_bin", capi.LIVOX_PT, path, eof_quirk)


def read_ouster_bin(path: str, eof_quirk: bool = True) -> np.ndarray:
    return _read("malio_read_ouster_bin", capi.OUSTER_PT, path, eof_quirk)


def preprocess_livox(pts: np.ndarray, n_scans: int = 6, point_filter_num: int = 1, blind: float = 0.5):
    """Preprocess::avia_handler.  Returns (capi.RAW_PT[m], intensity float32[m])."""
    lib = capi.load()
    pts = np.ascontiguousarray(pts)
    assert pts.dtype == capi.LIVOX_PT
    out = np.zeros(max(pts.shape[0], 1), dtype=capi.RAW_PT)
    inten = np.zeros(max(pts.shape[0], 1), np.float32)
    m = C.c_uint32(0)
    rc = lib.malio_preprocess_livox(capi.ptr(pts), pts.shape[0], n_scans, point_filter_num, C.c_double(blind), capi.ptr(out), capi.ptr(inten),
                                    out.shape[0], C.byref(m))
    if rc != capi.OK:
        raise capi.MalioError(rc, "malio_preprocess_livox")
    return out[: m.value].copy(), inten[: m.value].copy()


def preprocess_ouster(pts: np.ndarray, point_filter_num: int = 1, blind: float = 0.5, time_unit_scale: float = 1.0e-3):
    """Preprocess::oust64_handler.  Returns (capi.RAW_PT[m], intensity float32[m])."""
    lib = capi.load()
    pts = np.ascontiguousarray(pts)
    assert pts.dtype == capi.OUSTER_PT
    out = np.zeros(max(pts.shape[0], 1), dtype=capi.RAW_PT)
    inten = np.zeros(max(pts.shape[0], 1), np.float32)
    m = C.c_uint32(0)
    rc = lib.malio_preprocess_ouster(capi.ptr(pts), pts.shape[0], point_filter_num, C.c_double(blind), C.c_float(time_unit_scale),
                                     capi.ptr(out), capi.ptr(inten), out.shape[0], C.byref(m))
    if rc != capi.OK:
        raise capi.MalioError(rc, "malio_preprocess_ouster")
    return out[: m.value].copy(), inten[: m.value].copy()
