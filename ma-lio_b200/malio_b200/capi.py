"""ctypes binding of include/malio_b200.h (libmalio_b200.so, built in-tree by ma-lio_b200/csrc/Makefile).

The library is the product; this module only marshals numpy arrays through the C-ABI.  There is no CPU
fallback: if the shared library is missing, or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAX_LIDAR = 3
K = 5
MAX_DOF = 17 + 6 * MAX_LIDAR
NCCL_UNIQUE_ID_BYTES = 128

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_NCCL, ERR_NO_EFFECTIVE_POINTS, ERR_STATE, ERR_TREE_TOO_DEEP, ERR_CAPACITY = range(1, 8)

LINK_HAS_LEFT = 0x80000000
LINK_HAS_RIGHT = 0x40000000
LINK_POINT_DELETED = 0x20000000
LINK_INDEX_MASK = 0x0FFFFFFF
MAP_SKIP, MAP_ADD, MAP_ADD_NO_DOWNSAMPLE, MAP_DROP = 0, 1, 2, 3

# numpy views of the C structs
MAP_NODE = np.dtype([("xyz", "<f4", 3), ("link", "<u4"), ("lbox", "<f4", 6), ("rbox", "<f4", 6)])
MAP_POINT = np.dtype([("xyz", "<f4", 3), ("link", "<u4")])
POSE = np.dtype([("q", "<f8", 4), ("t", "<f8", 3), ("T", "<f8", (4, 4)), ("cov", "<f8", (6, 6))])   # malio_pose
SCAN_PT = np.dtype([("xyz", "<f4", 3), ("lidar", "<u2"), ("table_idx", "<u2")])
POSE_ENTRY = np.dtype([("T", "<f8", (4, 4)), ("cov", "<f8", (6, 6))])
RIGID = np.dtype([("q", "<f8", 4), ("t", "<f8", 3)])
LIVOX_PT = np.dtype([("xyz", "<f4", 3), ("reflectivity", "u1"), ("tag", "u1"), ("line", "u1"), ("pad", "u1"), ("offset_time", "<u4")])
OUSTER_PT = np.dtype([("xyz", "<f4", 3), ("intensity", "<f4"), ("ring", "<u2"), ("pad", "<u2"), ("t", "<u4")])
RAW_PT = np.dtype([("xyz", "<f4", 3), ("curvature", "<f4")])   # malio_raw_pt
IDX_UNTOUCHED = -(1 << 31)
assert MAP_NODE.itemsize == 64 and SCAN_PT.itemsize == 16 and POSE_ENTRY.itemsize == 416 and RIGID.itemsize == 56


class Rigid(C.Structure):
    _fields_ = [("q", C.c_double * 4), ("t", C.c_double * 3)]


class UndistortArgs(C.Structure):
    _fields_ = [("beg_time", C.c_double), ("extrinsic", Rigid), ("lt_imu_frame", Rigid), ("ctrl_t", C.c_void_p),
                ("ctrl_T", C.c_void_p), ("n_ctrl", C.c_uint32), ("imu_cov_t", C.c_void_p), ("n_cov", C.c_uint32),
                ("cov_pointer", C.c_int32)]


class PassState(C.Structure):
    _fields_ = [("rot", C.c_double * 4), ("pos", C.c_double * 3), ("ext", Rigid * MAX_LIDAR)]


class Params(C.Structure):
    _fields_ = [("n_lidar", C.c_int32), ("extrinsic_est_en", C.c_int32), ("plane_th", C.c_float),
                ("knn_max_sqdist", C.c_float), ("cov_threshold", C.c_double), ("point_cov_max", C.c_double),
                ("point_cov_min", C.c_double), ("plane_cov_max", C.c_double), ("plane_cov_min", C.c_double),
                ("localize_cov_max", C.c_double), ("localize_cov_min", C.c_double),
                ("localize_thresh_max", C.c_double), ("localize_thresh_min", C.c_double),
                ("range_min", C.c_double), ("range_max", C.c_double)]


class Config(C.Structure):
    _fields_ = [("params", Params), ("device", C.c_int32), ("sort_queries", C.c_int32),
                ("max_points", C.c_uint32), ("max_map_nodes", C.c_uint32), ("knn_cell_size", C.c_float)]


class PassStats(C.Structure):
    _fields_ = [("n_points", C.c_uint32), ("n_eff", C.c_uint32), ("valid", C.c_int32), ("searched", C.c_int32),
                ("u_min", C.c_double), ("u_max", C.c_double), ("tau_min", C.c_double), ("tau_max", C.c_double),
                ("sigma", C.c_double * 3), ("loc_weight", C.c_double), ("ms_sort", C.c_float), ("ms_knn", C.c_float),
                ("ms_plane", C.c_float), ("ms_reduce", C.c_float), ("ms_total", C.c_float)]


class State(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("rot", C.c_double * 4), ("ext", Rigid * MAX_LIDAR),
                ("vel", C.c_double * 3), ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("grav", C.c_double * 3)]

    def copy(self) -> "State":
        s = State()
        C.memmove(C.byref(s), C.byref(self), C.sizeof(State))
        return s

    def pass_state(self) -> PassState:
        ps = PassState()
        ps.rot[:] = self.rot[:]
        ps.pos[:] = self.pos[:]
        for l in range(MAX_LIDAR):
            ps.ext[l].q[:] = self.ext[l].q[:]
            ps.ext[l].t[:] = self.ext[l].t[:]
        return ps


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("knn_launches", C.c_uint64), ("knn_queries", C.c_uint64),
                ("knn_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("knn_fallback_queries", C.c_uint64), ("knn_ring2_queries", C.c_uint64),
                ("knn_candidates", C.c_uint64), ("pass_launches", C.c_uint64), ("pass_points", C.c_uint64),
                ("pass_fit_launches", C.c_uint64), ("pass_ms", C.c_double), ("knn_tie_queries", C.c_uint64),
                ("map_slots", C.c_uint64), ("map_live", C.c_uint64), ("map_compactions", C.c_uint64),
                ("exchange_min_wait_ms", C.c_double), ("exchange_sum_wait_ms", C.c_double), ("exchange_passes", C.c_uint64)]


class UpdateReport(C.Structure):
    _fields_ = [("passes", C.c_int32), ("searches", C.c_int32), ("converged_count", C.c_int32),
                ("last_status", C.c_int32), ("n_eff_last", C.c_uint32), ("ms_device_total", C.c_float),
                ("ms_host_solve", C.c_float), ("dx_last", C.c_double * MAX_DOF)]


_LIB_NAME = "libmalio_b200.so"
_lib = None

EXPORTS = [
    "malio_default_params", "malio_create", "malio_destroy", "malio_last_error", "malio_version",
    "malio_get_nccl_unique_id", "malio_comm_init", "malio_upload_map", "malio_upload_scan", "malio_measure",
    "malio_download_rows", "malio_download_aux", "malio_knn", "malio_ieskf_update", "malio_build_static_snapshot",
    "malio_rearm_scan", "malio_get_counters", "malio_set_timing", "malio_upload_map_compact", "malio_download_map_nodes", "malio_map_incremental",
    "malio_map_build", "malio_map_add_points", "malio_map_delete_boxes", "malio_map_sync_voxels", "malio_map_commit", "malio_map_info",
    "malio_map_download", "malio_read_livox_bin", "malio_read_ouster_bin", "malio_preprocess_livox", "malio_preprocess_ouster",
    "malio_undistort", "malio_bspline_get_pose", "malio_voxel_grid", "malio_upload_scan_device",
    "malio_pose_initial", "malio_compound_pose_with_cov", "malio_compound_inv_pose_with_cov", "malio_build_pose_unc",
]


def lib_path() -> str:
    # MALIO_LIB_PATH: another build of the same library (kernel-parameter experiments); the product default is in-tree
    return os.environ.get("MALIO_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load() -> C.CDLL:
    """Load the product library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"or `make -C ma-lio_b200/csrc`")
    lib = C.CDLL(path)
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    lib.malio_default_params.argtypes = [C.POINTER(Params), i32]
    lib.malio_default_params.restype = None
    lib.malio_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    lib.malio_destroy.argtypes = [vp]
    lib.malio_destroy.restype = None
    lib.malio_last_error.argtypes = [vp]
    lib.malio_last_error.restype = C.c_char_p
    lib.malio_version.restype = C.c_char_p
    lib.malio_get_nccl_unique_id.argtypes = [vp]
    lib.malio_comm_init.argtypes = [vp, vp, i32, i32]
    lib.malio_upload_map.argtypes = [vp, vp, vp, u32, u32]
    lib.malio_upload_map_compact.argtypes = [vp, vp, vp, u32, u32, vp]
    lib.malio_download_map_nodes.argtypes = [vp, vp, u32]
    lib.malio_map_incremental.argtypes = [vp, vp, C.c_double, i32, vp, vp]
    lib.malio_pose_initial.restype = None
    lib.malio_pose_initial.argtypes = [vp, vp, vp, vp]
    lib.malio_compound_pose_with_cov.restype = None
    lib.malio_compound_pose_with_cov.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.malio_compound_inv_pose_with_cov.restype = None
    lib.malio_compound_inv_pose_with_cov.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.malio_build_pose_unc.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    lib.malio_upload_scan.argtypes = [vp, vp, u32, vp, vp, vp]
    lib.malio_map_build.argtypes = [vp, vp, vp, vp, u32]
    lib.malio_map_add_points.argtypes = [vp, vp, vp, vp, u32]
    lib.malio_map_delete_boxes.argtypes = [vp, vp, u32, C.POINTER(u32)]
    lib.malio_map_sync_voxels.argtypes = [vp, vp, u32, vp, vp, vp, u32, C.POINTER(u32)]
    lib.malio_map_commit.argtypes = [vp]
    lib.malio_map_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    lib.malio_map_download.argtypes = [vp, vp, vp, vp, vp, u32, C.POINTER(u32)]
    lib.malio_read_livox_bin.argtypes = [C.c_char_p, vp, u32, C.POINTER(u32), i32]
    lib.malio_read_ouster_bin.argtypes = [C.c_char_p, vp, u32, C.POINTER(u32), i32]
    lib.malio_preprocess_livox.argtypes = [vp, u32, i32, i32, C.c_double, vp, vp, u32, C.POINTER(u32)]
    lib.malio_preprocess_ouster.argtypes = [vp, u32, i32, C.c_double, C.c_float, vp, vp, u32, C.POINTER(u32)]
    lib.malio_undistort.argtypes = [vp, i32, vp, u32, C.POINTER(UndistortArgs), vp, vp, vp, vp, C.POINTER(u32), vp]
    lib.malio_bspline_get_pose.argtypes = [vp, vp, u32, C.c_double, vp, vp]
    lib.malio_voxel_grid.argtypes = [vp, i32, vp, u32, C.c_float, vp, u32, C.POINTER(u32)]
    lib.malio_upload_scan_device.argtypes = [vp, vp, vp, vp, C.POINTER(u32)]
    lib.malio_measure.argtypes = [vp, C.POINTER(PassState), i32, vp, vp, C.POINTER(PassStats)]
    lib.malio_download_rows.argtypes = [vp, vp, vp, u32, C.POINTER(u32)]
    lib.malio_download_aux.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.malio_knn.argtypes = [vp, vp, u32, vp, vp, C.POINTER(C.c_float)]
    lib.malio_ieskf_update.argtypes = [vp, C.POINTER(State), vp, i32, C.c_double, C.POINTER(UpdateReport)]
    lib.malio_build_static_snapshot.argtypes = [vp, u32, vp, vp, C.POINTER(u32)]
    lib.malio_rearm_scan.argtypes = [vp]
    lib.malio_get_counters.argtypes = [vp, C.POINTER(Counters)]
    lib.malio_set_timing.argtypes = [vp, i32]
    _lib = lib
    return lib


def ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class MalioError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"malio status {status}: {msg}")
        self.status = status


def default_params(n_lidar: int) -> Params:
    """The reference defaults (City.yaml / mapping_city.launch) — the same values malio_default_params() fills in; kept in
    Python as well so that input generators (synth.py) work without mapping the CUDA library (bench.py --impl reference).
    tests/test_oracle_cpu.py checks the two agree."""
    p = Params()
    p.n_lidar = n_lidar
    p.extrinsic_est_en = 1
    p.plane_th = 0.4
    p.knn_max_sqdist = 5.0
    p.cov_threshold = 0.5
    p.point_cov_max, p.point_cov_min = 0.00125, 0.00075
    p.plane_cov_max, p.plane_cov_min = 1.0, 0.8
    p.localize_cov_max, p.localize_cov_min = 2.0, 0.3
    p.localize_thresh_max, p.localize_thresh_min = 0.7, 0.2
    p.range_min, p.range_max = 0.0, 1.0
    return p


def default_params_from_library(n_lidar: int) -> Params:
    p = Params()
    load().malio_default_params(C.byref(p), n_lidar)
    return p
