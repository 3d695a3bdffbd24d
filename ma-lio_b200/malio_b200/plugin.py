"""Host-side mirror of the reference's measurement plug-in, over the C-ABI.

Names follow the reference (paths relative to /root/reference/MA_LIO):
  MeasurementModel.h_share_model(state, converge)        src/laserMapping.cpp:552-760  (+ esekfom.hpp:622-635)
  MeasurementModel.update_iterated_dyn_share_modified()  include/IKFoM_toolkit/esekfom/esekfom.hpp:495-721
  MeasurementModel.Nearest_Search(points)                include/ikd-Tree/ikd_Tree.cpp:426-461
Everything heavy happens inside libmalio_b200.so on the GPU; this file is marshalling only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class MapSnapshot:
    """Flattened ikd-Tree (include/malio_flatten.hpp or malio_build_static_snapshot)."""
    nodes: np.ndarray        # capi.MAP_NODE[M]
    node_cov: np.ndarray     # float32[M]  map-side weight normal_y (common_lib.h:164)
    node_ids: np.ndarray     # int32/uint32[M] caller's id of the point stored in each node
    max_depth: int

    @property
    def n_nodes(self) -> int:
        return int(self.nodes.shape[0])


def build_static_snapshot(xyz: np.ndarray, normal_y: np.ndarray | float = 0.001) -> MapSnapshot:
    """Balanced k-d tree over xyz directly in snapshot form (median split on the longest axis, like
    KD_TREE::BuildTree, ikd_Tree.cpp:696-735).  Host C++ inside the product library."""
    lib = capi.load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    nodes = np.zeros(n, dtype=capi.MAP_NODE)
    order = np.zeros(n, dtype=np.uint32)
    depth = C.c_uint32(0)
    rc = lib.malio_build_static_snapshot(capi.ptr(xyz), n, capi.ptr(nodes), capi.ptr(order), C.byref(depth))
    if rc != capi.OK:
        raise capi.MalioError(rc, "malio_build_static_snapshot")
    if np.isscalar(normal_y):
        cov = np.full(n, normal_y, dtype=np.float32)
    else:
        cov = np.ascontiguousarray(np.asarray(normal_y, dtype=np.float32)[order])
    return MapSnapshot(nodes, cov, order.astype(np.int64), int(depth.value))


def compact_points(nodes: np.ndarray) -> np.ndarray:
    """capi.MAP_NODE[M] -> capi.MAP_POINT[M]: the first 16 bytes (point + link) of every record."""
    out = np.empty(nodes.shape[0], dtype=capi.MAP_POINT)
    out["xyz"] = nodes["xyz"]
    out["link"] = nodes["link"]
    return out


def build_pose_unc(extrinsic: np.ndarray, temporal_comp: np.ndarray | None, lidar_uncertainty: list[np.ndarray]):
    """pose_unc of one scan (laserMapping.cpp:1028-1048) through the library's host code.  extrinsic: capi.POSE[L],
    temporal_comp: capi.POSE[L-1] or None, lidar_uncertainty: L arrays of capi.POSE.  Returns (table capi.POSE_ENTRY[],
    table_off uint32[L+1]) ready for MeasurementModel.upload_scan."""
    lib = capi.load()
    L = len(lidar_uncertainty)
    lists = [np.ascontiguousarray(a, dtype=capi.POSE) for a in lidar_uncertainty]
    counts = np.array([a.shape[0] for a in lists], np.uint32)
    ptrs = (C.c_void_p * L)(*[a.ctypes.data for a in lists])
    n = int(sum(max(int(c) - 1, 0) for c in counts))
    table = np.zeros(max(n, 1), dtype=capi.POSE_ENTRY)
    off = np.zeros(L + 1, np.uint32)
    ext = np.ascontiguousarray(extrinsic, dtype=capi.POSE)
    tc = None if temporal_comp is None else np.ascontiguousarray(temporal_comp, dtype=capi.POSE)
    rc = lib.malio_build_pose_unc(L, capi.ptr(ext), capi.ptr(tc), C.cast(ptrs, C.c_void_p), capi.ptr(counts), capi.ptr(table),
                                  capi.ptr(off))
    if rc < 0:
        raise capi.MalioError(-rc, "malio_build_pose_unc")
    return table[:rc], off


def bspline_get_pose(ctrl_t: np.ndarray, ctrl_T: np.ndarray, timestamp: float):
    """BsplineSE3::get_pose (BsplineSE3.cpp:84-118) through the library's host code.  Returns (ok, q wxyz, p)."""
    ct = np.ascontiguousarray(ctrl_t, np.float64)
    cT = np.ascontiguousarray(ctrl_T, np.float64).reshape(-1, 16)
    q, p = np.zeros(4), np.zeros(3)
    ok = capi.load().malio_bspline_get_pose(capi.ptr(ct), capi.ptr(cT), ct.shape[0], float(timestamp), capi.ptr(q), capi.ptr(p))
    return bool(ok), q, p


class MeasurementModel:
    """One handle = one GPU.  Mirrors the life cycle of one scan in laserMapping.cpp:935-1082."""

    def __init__(self, n_lidar: int = 3, device: int = 0, sort_queries: bool = True, params: capi.Params | None = None,
                 knn_cell_size: float = 0.0):
        """knn_cell_size: edge of the k-NN fast path's cell-list index (0 automatic, < 0 off = ikd-Tree traversal only)."""
        self.lib = capi.load()
        cfg = capi.Config()
        cfg.params = params if params is not None else capi.default_params(n_lidar)
        cfg.params.n_lidar = n_lidar
        cfg.device = device
        cfg.sort_queries = 1 if sort_queries else 0
        cfg.knn_cell_size = float(knn_cell_size)
        self.n_lidar = n_lidar
        self.n_dof = 17 + 6 * n_lidar
        self.n_cols = 6 * (n_lidar + 1)
        self._h = C.c_void_p()
        rc = self.lib.malio_create(C.byref(self._h), C.byref(cfg))
        if rc != capi.OK:
            raise capi.MalioError(rc, self.lib.malio_last_error(None).decode())
        self.n_points = 0
        self._keep = []   # host buffers must outlive the (synchronous) calls only; kept for clarity

    def close(self):
        if self._h:
            self.lib.malio_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow=()):
        if rc != capi.OK and rc not in allow:
            raise capi.MalioError(rc, self.lib.malio_last_error(self._h).decode())
        return rc

    # ---- multi-GPU plumbing
    @staticmethod
    def nccl_unique_id() -> np.ndarray:
        uid = np.zeros(capi.NCCL_UNIQUE_ID_BYTES, dtype=np.uint8)
        rc = capi.load().malio_get_nccl_unique_id(capi.ptr(uid))
        if rc != capi.OK:
            raise capi.MalioError(rc, "malio_get_nccl_unique_id")
        return uid

    def comm_init(self, uid: np.ndarray, rank: int, world: int):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._check(self.lib.malio_comm_init(self._h, capi.ptr(uid), rank, world))

    # ---- per scan
    def upload_map(self, snap: MapSnapshot):
        nodes = np.ascontiguousarray(snap.nodes)
        cov = np.ascontiguousarray(snap.node_cov, dtype=np.float32)
        self._check(self.lib.malio_upload_map(self._h, capi.ptr(nodes), capi.ptr(cov), nodes.shape[0], snap.max_depth))

    def upload_map_compact(self, snap: MapSnapshot, points: np.ndarray | None = None, root_box: np.ndarray | None = None):
        """Same snapshot, 16 bytes per node (point + link) + the weights: the device rebuilds the children's boxes.
        `points` may hold the pre-extracted capi.MAP_POINT array (e.g. in pinned memory); `root_box` the map's bounding
        box {x_min,x_max,y_min,y_max,z_min,z_max} when the caller knows it (the ikd-Tree root's node_range_*)."""
        pts = compact_points(snap.nodes) if points is None else points
        cov = np.ascontiguousarray(snap.node_cov, dtype=np.float32)
        rb = None if root_box is None else np.ascontiguousarray(root_box, dtype=np.float32)
        self._check(self.lib.malio_upload_map_compact(self._h, capi.ptr(pts), capi.ptr(cov), pts.shape[0], snap.max_depth,
                                                      capi.ptr(rb)))

    def download_map_nodes(self, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=capi.MAP_NODE)
        self._check(self.lib.malio_download_map_nodes(self._h, capi.ptr(out), n))
        return out

    def upload_scan(self, pts: np.ndarray, table: np.ndarray, table_off: np.ndarray, temporal_comp: np.ndarray | None):
        pts = np.ascontiguousarray(pts)
        assert pts.dtype == capi.SCAN_PT
        table = np.ascontiguousarray(table)
        assert table.dtype == capi.POSE_ENTRY
        table_off = np.ascontiguousarray(table_off, dtype=np.uint32)
        tc = None if temporal_comp is None else np.ascontiguousarray(temporal_comp)
        self._check(self.lib.malio_upload_scan(self._h, capi.ptr(pts), pts.shape[0], capi.ptr(table),
                                               capi.ptr(table_off), capi.ptr(tc)))
        self.n_points = int(pts.shape[0])

    # ---- the map as a device-resident point set kept in step by deltas (include/malio_mapsync.hpp)
    @staticmethod
    def _f32(a, cols=None):
        a = np.ascontiguousarray(a, np.float32)
        return a if cols is None else a.reshape(-1, cols)

    def map_build(self, xyz, normal_y, ids=None):
        """KD_TREE::Build: the device-resident map starts as these points (handle switches to device-map mode)."""
        xyz = self._f32(xyz, 3); ny = self._f32(normal_y)
        i = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self.lib.malio_map_build(self._h, capi.ptr(xyz), capi.ptr(ny), capi.ptr(i), xyz.shape[0]))

    def map_add_points(self, xyz, normal_y, ids=None):
        """Mirror of KD_TREE::Add_Points(points, false)."""
        xyz = self._f32(xyz, 3); ny = self._f32(normal_y)
        i = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self.lib.malio_map_add_points(self._h, capi.ptr(xyz), capi.ptr(ny), capi.ptr(i), xyz.shape[0]))

    def map_delete_boxes(self, boxes) -> int:
        """Mirror of KD_TREE::Delete_Point_Boxes; boxes: nb x {min3, max3}, half-open.  Returns the points deleted."""
        b = self._f32(boxes, 6)
        n = C.c_uint32(0)
        self._check(self.lib.malio_map_delete_boxes(self._h, capi.ptr(b), b.shape[0], C.byref(n)))
        return int(n.value)

    def map_sync_voxels(self, sync: dict, want_count: bool = True) -> int:
        """After the host tree's Add_Points(points, true): replace the content of every touched voxel box by what
        malio::collect_voxel_sync read back from the tree (dict with boxes, xyz, normal_y, ids).  want_count=False passes
        n_deleted = NULL: the call then returns without waiting for the device (-1 is returned)."""
        b = self._f32(sync["boxes"], 6); xyz = self._f32(sync["xyz"], 3); ny = self._f32(sync["normal_y"])
        i = None if sync.get("ids") is None else np.ascontiguousarray(sync["ids"], np.int32)
        n = C.c_uint32(0)
        self._check(self.lib.malio_map_sync_voxels(self._h, capi.ptr(b) if b.shape[0] else None, b.shape[0],
                                                   capi.ptr(xyz) if xyz.shape[0] else None, capi.ptr(ny) if xyz.shape[0] else None,
                                                   capi.ptr(i) if (i is not None and xyz.shape[0]) else None, xyz.shape[0],
                                                   C.byref(n) if want_count else None))
        return int(n.value) if want_count else -1

    def map_commit(self):
        self._check(self.lib.malio_map_commit(self._h))

    def map_info(self):
        a, b = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.malio_map_info(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def map_download(self):
        """Live points of the device-resident map in slot order: dict(xyz, normal_y, ids, slots)."""
        live, _ = self.map_info()
        xyz = np.zeros((live, 3), np.float32); ny = np.zeros(live, np.float32); ids = np.zeros(live, np.int32); sl = np.zeros(live, np.uint32)
        n = C.c_uint32(0)
        self._check(self.lib.malio_map_download(self._h, capi.ptr(xyz), capi.ptr(ny), capi.ptr(ids), capi.ptr(sl), live, C.byref(n)))
        return dict(xyz=xyz, normal_y=ny, ids=ids, slots=sl)

    # ---- the two stages before the path
    def undistort(self, lidar: int, pts: np.ndarray, beg_time: float, extrinsic, lt_imu_frame, ctrl_t, ctrl_T, imu_cov_t,
                  cov_pointer: int, want_pose: bool = False):
        """UndistortPcl's per-point loop for one LiDAR (IMU_Processing.hpp:468-508).  pts: capi.RAW_PT[n] sorted by
        curvature; extrinsic / lt_imu_frame: (q wxyz, t).  Returns dict(xyz, idx, ok, pop_point, pose?)."""
        pts = np.ascontiguousarray(pts)
        assert pts.dtype == capi.RAW_PT
        n = pts.shape[0]
        ct = np.ascontiguousarray(ctrl_t, np.float64)
        cT = np.ascontiguousarray(ctrl_T, np.float64).reshape(-1, 16)
        cv = np.ascontiguousarray(imu_cov_t, np.float64)
        a = capi.UndistortArgs()
        a.beg_time = float(beg_time)
        a.extrinsic.q[:] = list(map(float, extrinsic[0])); a.extrinsic.t[:] = list(map(float, extrinsic[1]))
        a.lt_imu_frame.q[:] = list(map(float, lt_imu_frame[0])); a.lt_imu_frame.t[:] = list(map(float, lt_imu_frame[1]))
        a.ctrl_t = ct.ctypes.data; a.ctrl_T = cT.ctypes.data; a.n_ctrl = ct.shape[0]
        a.imu_cov_t = cv.ctypes.data if cv.shape[0] else None; a.n_cov = cv.shape[0]; a.cov_pointer = int(cov_pointer)
        xyz = np.zeros((n, 3), np.float32); idx = np.zeros(n, np.int32); ok = np.zeros(n, np.uint8)
        pop = np.full(max(cv.shape[0], 1), -1, np.int32); npop = C.c_uint32(0)
        pose = np.zeros((n, 7)) if want_pose else None
        self._check(self.lib.malio_undistort(self._h, lidar, capi.ptr(pts), n, C.byref(a), capi.ptr(xyz), capi.ptr(idx), capi.ptr(ok),
                                             capi.ptr(pop), C.byref(npop), capi.ptr(pose)))
        return dict(xyz=xyz, idx=idx, ok=ok, pop_point=pop[: npop.value], pose=pose)

    def voxel_grid(self, lidar: int, pts5: np.ndarray | None, leaf: float):
        """pcl::VoxelGrid (laserMapping.cpp:968-983) on the device.  pts5: float32[n,5] {x,y,z,intensity,curvature}, or None to
        take the device-resident output of the last undistort() of this LiDAR slot.  Returns float32[m,5]."""
        n = 0 if pts5 is None else int(pts5.shape[0])
        inp = None if pts5 is None else np.ascontiguousarray(pts5, np.float32)
        m = C.c_uint32(0)
        self._check(self.lib.malio_voxel_grid(self._h, lidar, capi.ptr(inp), n, C.c_float(leaf), None, 0, C.byref(m)))
        out = np.zeros((m.value, 5), np.float32)
        if m.value:   # second call only fetches (the result is device-resident); kept simple: recompute with the output buffer
            self._check(self.lib.malio_voxel_grid(self._h, lidar, capi.ptr(inp), n, C.c_float(leaf), capi.ptr(out), m.value, C.byref(m)))
        return out

    def upload_scan_device(self, table: np.ndarray, table_off: np.ndarray, temporal_comp: np.ndarray | None) -> int:
        """Merge the device-resident down-sampled clouds of all LiDAR slots into the scan (no host bounce)."""
        table = np.ascontiguousarray(table); table_off = np.ascontiguousarray(table_off, dtype=np.uint32)
        tc = None if temporal_comp is None else np.ascontiguousarray(temporal_comp)
        n = C.c_uint32(0)
        self._check(self.lib.malio_upload_scan_device(self._h, capi.ptr(table), capi.ptr(table_off), capi.ptr(tc), C.byref(n)))
        self.n_points = int(n.value)
        return self.n_points

    def rearm_scan(self):
        """Reset the per-scan state of the scan already resident on the device (no host copy)."""
        self._check(self.lib.malio_rearm_scan(self._h))

    def set_timing(self, enable: bool):
        """Per-pass CUDA-event timing (PassStats.ms_*, Counters.knn_ms); on by default."""
        self._check(self.lib.malio_set_timing(self._h, 1 if enable else 0))

    def counters(self) -> capi.Counters:
        c = capi.Counters()
        self._check(self.lib.malio_get_counters(self._h, C.byref(c)))
        return c

    def h_share_model(self, s: capi.PassState | capi.State, converge: bool):
        """One measurement pass.  Returns (valid, HTH[c,c], HTh[c], stats)."""
        ps = s.pass_state() if isinstance(s, capi.State) else s
        c = self.n_cols
        HTH = np.zeros((c, c))
        HTh = np.zeros(c)
        st = capi.PassStats()
        rc = self._check(self.lib.malio_measure(self._h, C.byref(ps), 1 if converge else 0, capi.ptr(HTH),
                                                capi.ptr(HTh), C.byref(st)), allow=(capi.ERR_NO_EFFECTIVE_POINTS,))
        return rc == capi.OK, HTH, HTh, st

    def rows(self, cap: int | None = None):
        cap = capi.MAX_DOF if cap is None else cap
        hx = np.zeros((cap, self.n_cols))
        hv = np.zeros(cap)
        n = C.c_uint32(0)
        self._check(self.lib.malio_download_rows(self._h, capi.ptr(hx), capi.ptr(hv), cap, C.byref(n)))
        return hx[: n.value], hv[: n.value]

    def aux(self, normal_y=True, nn_idx=True, nn_sqdist=True, selected=True, world=True, out: dict | None = None):
        """Side outputs of the last pass in caller order.  `out` may carry pre-allocated (e.g. pinned) arrays under the
        same keys; missing ones are allocated."""
        n = self.n_points
        if out is not None:
            self._check(self.lib.malio_download_aux(self._h, capi.ptr(out.get("normal_y")), capi.ptr(out.get("nn_idx")),
                                                    capi.ptr(out.get("nn_sqdist")), capi.ptr(out.get("selected")),
                                                    capi.ptr(out.get("world"))))
            return out
        o_ny = np.zeros(n, np.float32) if normal_y else None
        o_idx = np.zeros((n, capi.K), np.uint32) if nn_idx else None
        o_d2 = np.zeros((n, capi.K), np.float32) if nn_sqdist else None
        o_sel = np.zeros(n, np.uint8) if selected else None
        o_w = np.zeros((n, 3), np.float32) if world else None
        self._check(self.lib.malio_download_aux(self._h, capi.ptr(o_ny), capi.ptr(o_idx), capi.ptr(o_d2),
                                                capi.ptr(o_sel), capi.ptr(o_w)))
        return dict(normal_y=o_ny, nn_idx=o_idx, nn_sqdist=o_d2, selected=o_sel, world=o_w)

    def map_incremental(self, s: capi.PassState | capi.State, filter_size_map: float = 0.5, ekf_inited: bool = True,
                        world: bool = True):
        """map_incremental's per-point decision (laserMapping.cpp:398-446) with the state after the update.
        Returns (cls uint8[N] with capi.MAP_* codes, feats_down_world float32[N,3] or None)."""
        ps = s.pass_state() if isinstance(s, capi.State) else s
        cls = np.zeros(self.n_points, np.uint8)
        w = np.zeros((self.n_points, 3), np.float32) if world else None
        self._check(self.lib.malio_map_incremental(self._h, C.byref(ps), C.c_double(filter_size_map), 1 if ekf_inited else 0,
                                                   capi.ptr(cls), capi.ptr(w)))
        return cls, w

    def update_iterated_dyn_share_modified(self, x: capi.State, P: np.ndarray, max_iter: int, R: float = 0.001):
        """IESKF update; x and P are updated in place.  Returns the report (status in report.last_status)."""
        assert P.shape == (self.n_dof, self.n_dof) and P.dtype == np.float64 and P.flags["C_CONTIGUOUS"]
        rep = capi.UpdateReport()
        rc = self.lib.malio_ieskf_update(self._h, C.byref(x), capi.ptr(P), max_iter, R, C.byref(rep))
        self._check(rc, allow=(capi.ERR_NO_EFFECTIVE_POINTS,))
        return rep

    # ---- stand-alone k-NN (BASELINE config C5)
    def Nearest_Search(self, queries: np.ndarray):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        n = q.shape[0]
        idx = np.zeros((n, capi.K), np.uint32)
        d2 = np.zeros((n, capi.K), np.float32)
        ms = C.c_float(0)
        self._check(self.lib.malio_knn(self._h, capi.ptr(q), n, capi.ptr(idx), capi.ptr(d2), C.byref(ms)))
        return idx, d2, float(ms.value)
