"""Seeded synthetic inputs for the measurement hot path: an "urban grid" map, a merged
multi-LiDAR scan expressed in each sensor's frame, pose-uncertainty tables, a perturbed prior state.

Deterministic (numpy PCG64).  Used by tests/, bench.py and __graft_entry__.smoke(); the same buffers go to
the CUDA path and to the CPU oracle.  Extrinsics are the reference's /root/reference/Extrinsic.txt values.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi

# Extrinsic.txt (LiDAR from IMU): (qw,qx,qy,qz), t
EXTRINSICS = [
    ((1.0, 0.0, 0.0, 0.0), (0.215, 0.0, 0.018)),                                   # Ouster
    ((0.6965018, -0.0037329, -0.0038405, 0.717535), (-1.2574, 0.413, 0.0324)),     # Livox Avia
    ((0.0074645, 0.0000044, -0.0005919, -0.999972), (-1.306, -0.361, 0.042)),      # Livox Tele
]
LIDAR_SPLIT = (0.70, 0.13, 0.17)   # City01 ratio implied by paper Table VI (8554, +1618, +2072)
GRAV_LEN = 9.809


# ------------------------------------------------------------------ quaternion helpers (w,x,y,z), float64
def q_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def q_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx])


def q_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def q_exp(v):
    v = np.asarray(v, dtype=np.float64)
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * v / th])


def q_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rand_unit(rng, n=3):
    v = rng.normal(size=n)
    return v / np.linalg.norm(v)


# ------------------------------------------------------------------ world
WALL_H = 20.0
LATTICE = 0.5
BLOCK = 50.0


def _wall_lines(half):
    ks = np.arange(-int(half // BLOCK) - 1, int(half // BLOCK) + 2)
    c = BLOCK / 2 + BLOCK * ks
    return c[np.abs(c) <= half]


def make_world(M: int, seed: int = 42) -> np.ndarray:
    """Ground plane z=0 plus a street grid of 20 m walls every 50 m, sampled on a 0.5 m lattice with in-plane
    jitter U(-0.05,0.05) and normal noise N(0,0.01^2); sub-sampled to exactly M points."""
    rng = np.random.default_rng(seed)
    S = 50.0
    while True:
        half = S / 2
        n_side = int(round(S / LATTICE))
        n_lines = len(_wall_lines(half))
        total = n_side * n_side + 2 * n_lines * n_side * int(WALL_H / LATTICE)
        if total >= M:
            break
        S += 10.0
    half = S / 2
    g = (np.arange(n_side) + 0.5) * LATTICE - half
    gx, gy = np.meshgrid(g, g, indexing="ij")
    ground = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1)
    ground[:, :2] += rng.uniform(-0.05, 0.05, size=(ground.shape[0], 2))
    ground[:, 2] += rng.normal(0, 0.01, size=ground.shape[0])
    parts = [ground]
    hz = (np.arange(int(WALL_H / LATTICE)) + 0.5) * LATTICE
    for c in _wall_lines(half):
        a, z = np.meshgrid(g, hz, indexing="ij")
        n = a.size
        wx = np.stack([a.ravel() + rng.uniform(-0.05, 0.05, n), np.full(n, c) + rng.normal(0, 0.01, n),
                       z.ravel() + rng.uniform(-0.05, 0.05, n)], axis=1)   # wall along x at y=c
        wy = np.stack([np.full(n, c) + rng.normal(0, 0.01, n), a.ravel() + rng.uniform(-0.05, 0.05, n),
                       z.ravel() + rng.uniform(-0.05, 0.05, n)], axis=1)   # wall along y at x=c
        parts += [wx, wy]
    pts = np.concatenate(parts, axis=0)
    keep = rng.choice(pts.shape[0], size=M, replace=False)
    keep.sort()
    return np.ascontiguousarray(pts[keep].astype(np.float32))


def world_half_extent(M: int) -> float:
    S = 50.0
    while True:
        n_side = int(round(S / LATTICE))
        total = n_side * n_side + 2 * len(_wall_lines(S / 2)) * n_side * int(WALL_H / LATTICE)
        if total >= M:
            return S / 2
        S += 10.0


def sample_surfaces(N: int, center_xy, radius: float, half: float, rng) -> np.ndarray:
    """N points on the ground / walls within `radius` (horizontal) of center_xy, area-weighted."""
    out = []
    need = N
    cx, cy = center_xy
    lo_x, hi_x = max(cx - radius, -half), min(cx + radius, half)
    lo_y, hi_y = max(cy - radius, -half), min(cy + radius, half)
    lines = _wall_lines(half)
    xl = lines[(lines >= lo_y) & (lines <= hi_y)]     # walls along x (at y = c)
    yl = lines[(lines >= lo_x) & (lines <= hi_x)]     # walls along y (at x = c)
    a_ground = (hi_x - lo_x) * (hi_y - lo_y)
    a_xw = len(xl) * (hi_x - lo_x) * WALL_H
    a_yw = len(yl) * (hi_y - lo_y) * WALL_H
    w = np.array([a_ground, a_xw, a_yw])
    w = w / w.sum()
    while need > 0:
        n = int(need * 1.5) + 64
        kind = rng.choice(3, size=n, p=w)
        p = np.zeros((n, 3))
        u = rng.uniform(size=(n, 3))
        m = kind == 0
        p[m, 0] = lo_x + u[m, 0] * (hi_x - lo_x); p[m, 1] = lo_y + u[m, 1] * (hi_y - lo_y); p[m, 2] = 0.0
        m = kind == 1
        if len(xl):
            p[m, 0] = lo_x + u[m, 0] * (hi_x - lo_x); p[m, 1] = xl[(u[m, 1] * len(xl)).astype(int) % len(xl)]; p[m, 2] = u[m, 2] * WALL_H
        m = kind == 2
        if len(yl):
            p[m, 1] = lo_y + u[m, 0] * (hi_y - lo_y); p[m, 0] = yl[(u[m, 1] * len(yl)).astype(int) % len(yl)]; p[m, 2] = u[m, 2] * WALL_H
        ok = (p[:, 0] - cx) ** 2 + (p[:, 1] - cy) ** 2 <= radius * radius
        p = p[ok][:need]
        out.append(p)
        need -= p.shape[0]
    return np.concatenate(out, axis=0)


# ------------------------------------------------------------------ case container
@dataclass
class Case:
    name: str
    n_lidar: int
    max_iter: int
    map_xyz: np.ndarray            # float32 [M,3]
    map_normal_y: np.ndarray       # float32 [M]
    pts: np.ndarray                # capi.SCAN_PT [N]
    table: np.ndarray              # capi.POSE_ENTRY [sum T_l]
    table_off: np.ndarray          # uint32 [L+1]
    temporal_comp: np.ndarray | None   # capi.RIGID [L-1]
    x_true: capi.State
    x_prop: capi.State
    P_prop: np.ndarray             # [n,n]
    params: capi.Params
    meta: dict = field(default_factory=dict)

    @property
    def n_dof(self):
        return 17 + 6 * self.n_lidar


def make_state(pos, rot, ext_q, ext_t, vel=(0, 0, 0), bg=(0, 0, 0), ba=(0, 0, 0), grav=(0, 0, -GRAV_LEN)) -> capi.State:
    s = capi.State()
    s.pos[:] = list(map(float, pos))
    s.rot[:] = list(map(float, rot))
    for l in range(capi.MAX_LIDAR):
        q = ext_q[l] if l < len(ext_q) else (1.0, 0.0, 0.0, 0.0)
        t = ext_t[l] if l < len(ext_t) else (0.0, 0.0, 0.0)
        s.ext[l].q[:] = list(map(float, q))
        s.ext[l].t[:] = list(map(float, t))
    s.vel[:] = list(map(float, vel))
    s.bg[:] = list(map(float, bg))
    s.ba[:] = list(map(float, ba))
    s.grav[:] = list(map(float, grav))
    return s


def state_to_vec(s: capi.State, L: int) -> np.ndarray:
    """flat numeric dump (for comparisons): pos, rot(4), ext q/t, vel, bg, ba, grav"""
    out = list(s.pos) + list(s.rot)
    for l in range(L):
        out += list(s.ext[l].q) + list(s.ext[l].t)
    out += list(s.vel) + list(s.bg) + list(s.ba) + list(s.grav)
    return np.array(out)


def init_P(n: int) -> np.ndarray:
    """IMU_Processing.hpp:188-199 pattern with the pose block at 1e-4."""
    P = np.eye(n)
    for i in range(6):
        P[i, i] = 1e-4
    for i in range(6, n):
        if i < n - 8:
            P[i, i] = 1e-6
        elif i < n - 5:
            P[i, i] = 1e-4
        elif i < n - 2:
            P[i, i] = 1e-3
        else:
            P[i, i] = 1e-5
    return P


def make_tables(L: int, T: int, rng):
    """pose_unc[l][j]: small SE(3) (<= 2 cm, <= 0.2 deg) and covariance growing with j (seed 45 stream)."""
    table = np.zeros(L * T, dtype=capi.POSE_ENTRY)
    for l in range(L):
        axis = rand_unit(rng)
        tdir = rand_unit(rng)
        for j in range(T):
            f = (j + 1) / T
            q = q_exp(axis * np.deg2rad(0.2) * f)
            Tm = np.eye(4)
            Tm[:3, :3] = q_to_R(q)
            Tm[:3, 3] = tdir * 0.02 * f
            d = np.linspace(1e-8, 1e-6, 6) * (0.2 + f)
            A = rng.normal(size=(6, 6)) * 1e-4 * np.sqrt(f)
            cov = np.diag(d) + A @ A.T * 0.05
            table[l * T + j]["T"] = Tm
            table[l * T + j]["cov"] = cov
    table_off = np.arange(L + 1, dtype=np.uint32) * T
    return table, table_off


def make_case(name: str, N: int, M: int, n_lidar: int, max_iter: int, *, varied_map_cov: bool = False,
              seed_world: int = 42, seed_scan: int = 43, seed_state: int = 44, seed_tables: int = 45,
              det_range: float = 100.0, table_T: int = 10, map_xyz: np.ndarray | None = None,
              pos_err: float = 0.10, rot_err_deg: float = 0.5) -> Case:
    L = n_lidar
    if map_xyz is None:
        map_xyz = make_world(M, seed_world)
    half = world_half_extent(M)
    rng_w = np.random.default_rng(seed_world + 1000)
    if varied_map_cov:
        map_ny = rng_w.uniform(0.0005, 0.02, size=M).astype(np.float32)
        map_ny[rng_w.uniform(size=M) < 0.02] = 0.0   # exercises the W(0,0) > 1e-5 gate and cov_plane == 0
    else:
        map_ny = np.full(M, 0.001, dtype=np.float32)   # laserMapping.cpp:1004

    rng = np.random.default_rng(seed_scan)
    pos_true = np.array([1.0, 2.0, 1.8])
    rot_true = q_mul(q_exp([0, 0, 0.3]), q_exp([0.02, -0.015, 0.0]))
    extq = [q_normalize(EXTRINSICS[l][0]) for l in range(L)]
    extt = [np.array(EXTRINSICS[l][1]) for l in range(L)]
    rng_t = np.random.default_rng(seed_tables)
    tcomp = np.zeros(max(L - 1, 0), dtype=capi.RIGID)
    for l in range(1, L):
        tcomp[l - 1]["q"] = q_exp(rand_unit(rng_t) * np.deg2rad(0.3))
        tcomp[l - 1]["t"] = rand_unit(rng_t) * 0.05
    table, table_off = make_tables(L, table_T, rng_t)

    pw = sample_surfaces(N, pos_true[:2], min(det_range, half), half, rng)
    pw += rng.normal(0, 0.02, size=pw.shape)
    Rt = q_to_R(rot_true)
    p_imu = (pw - pos_true) @ Rt          # R^T (p - pos), row-vector form
    split = np.array(LIDAR_SPLIT[:L], dtype=np.float64)
    split = split / split.sum()
    counts = np.floor(split * N).astype(int)
    counts[0] += N - counts.sum()
    lidar = np.repeat(np.arange(L), counts)
    pts = np.zeros(N, dtype=capi.SCAN_PT)
    xyz = np.zeros((N, 3))
    for l in range(L):
        m = lidar == l
        v = p_imu[m]
        if l != 0:
            Rc = q_to_R(tcomp[l - 1]["q"])
            v = (v - tcomp[l - 1]["t"]) @ Rc          # qC^T (p - tC)
        Re = q_to_R(extq[l])
        xyz[m] = (v - extt[l]) @ Re                   # qE^T (. - tE)
    pts["xyz"] = xyz.astype(np.float32)
    pts["lidar"] = lidar.astype(np.uint16)
    pts["table_idx"] = rng.integers(0, table_T, size=N).astype(np.uint16)

    rs = np.random.default_rng(seed_state)
    x_true = make_state(pos_true, rot_true, extq, extt, vel=rs.normal(0, 0.5, 3), bg=rs.normal(0, 1e-3, 3),
                        ba=rs.normal(0, 1e-2, 3))
    pos_p = pos_true + rand_unit(rs) * pos_err
    rot_p = q_mul(rot_true, q_exp(rand_unit(rs) * np.deg2rad(rot_err_deg)))
    extq_p = [q_mul(extq[l], q_exp(rand_unit(rs) * np.deg2rad(0.1))) for l in range(L)]
    extt_p = [extt[l] + rand_unit(rs) * 0.01 for l in range(L)]
    x_prop = make_state(pos_p, rot_p, extq_p, extt_p, vel=x_true.vel[:], bg=x_true.bg[:], ba=x_true.ba[:])
    prm = capi.default_params(L)
    return Case(name, L, max_iter, map_xyz, map_ny, pts, table, table_off, tcomp if L > 1 else None, x_true, x_prop,
                init_P(17 + 6 * L), prm, meta=dict(N=N, M=M, half=half))


# BASELINE.json configs
def case_C1():
    return make_case("C1 single-LiDAR 5k scan vs 50k map, 3 iters", 5000, 50000, 1, 3)


def case_C2():
    return make_case("C2 3-LiDAR 100k scan vs 1M map", 100000, 1000000, 3, 3)


def case_C4():
    return make_case("C4 dense urban 300k scan vs 5M map, 5 iters", 300000, 5000000, 3, 5)


def knn_microbench(M: int = 10_000_000, Q: int = 1_000_000, seed_map: int = 42, seed_q: int = 43):
    """C5: uniform points in [-500,500]^2 x [-5,5]; queries = random map points + N(0, 0.1^2)."""
    rng = np.random.default_rng(seed_map)
    xyz = np.empty((M, 3), dtype=np.float32)
    xyz[:, 0] = rng.uniform(-500, 500, M)
    xyz[:, 1] = rng.uniform(-500, 500, M)
    xyz[:, 2] = rng.uniform(-5, 5, M)
    rq = np.random.default_rng(seed_q)
    q = xyz[rq.integers(0, M, Q)] + rq.normal(0, 0.1, size=(Q, 3)).astype(np.float32)
    return xyz, np.ascontiguousarray(q.astype(np.float32))


# ------------------------------------------------------------------ N2: raw scan + spline for the undistortion
def _se3(q, t):
    T = np.eye(4)
    T[:3, :3] = q_to_R(q)
    T[:3, 3] = t
    return T


def undistort_case(n: int = 60000, lidar: int = 0, seed: int = 51, scan_ms: float = 100.0, imu_hz: float = 200.0,
                   sorted_times: bool = True):
    """One LiDAR's raw scan for UndistortPcl's point loop: n points with time offsets (`curvature`, ms) over one scan, a
    cubic-spline control-point set at the reference's fixed 10 ms spacing (BsplineSE3.cpp:34) covering the scan with two
    control points of margin on both sides (a smooth vehicle motion: ~8 m/s, ~0.6 rad/s), the IMU-covariance time list at
    imu_hz, and cov_pointer as IMU_Processing.hpp:455-466 leaves it.  Returns a dict of plain arrays."""
    rng = np.random.default_rng(seed)
    t0 = 1000.0 + 0.0137          # lidar_beg_time
    end_time = t0 + scan_ms * 1e-3
    ctrl_t = t0 - 0.031 + 0.01 * np.arange(int(scan_ms / 10) + 8)
    axis = rand_unit(rng)
    w = axis * 0.6
    v = np.array([8.0, 0.5, -0.1])
    acc = rng.normal(0, 1.0, 3)
    ctrl_T = np.zeros((ctrl_t.shape[0], 4, 4))
    for k, tk in enumerate(ctrl_t):
        dt = tk - t0
        q = q_mul(q_exp([0.1, -0.05, 0.7]), q_exp(w * dt + 0.5 * rng.normal(0, 0.01, 3) * dt * dt))
        ctrl_T[k] = _se3(q, np.array([3.0, -2.0, 1.5]) + v * dt + 0.5 * acc * dt * dt)
    pts = np.zeros(n, dtype=capi.RAW_PT)
    rad = rng.uniform(1.0, 80.0, n)
    az = rng.uniform(-np.pi, np.pi, n)
    el = rng.uniform(-0.4, 0.4, n)
    pts["xyz"] = np.stack([rad * np.cos(el) * np.cos(az), rad * np.cos(el) * np.sin(az), rad * np.sin(el)], axis=1).astype(np.float32)
    curv = rng.uniform(0.0, scan_ms, n).astype(np.float32)
    pts["curvature"] = np.sort(curv) if sorted_times else curv
    # imu_cov: forward-propagation stamps from before the scan start to past its end (IMU_Processing.hpp:327-360)
    cov_t = t0 - 0.004 + np.arange(int((scan_ms * 1e-3 + 0.02) * imu_hz)) / imu_hz
    cp = cov_t.shape[0] - 1
    while cov_t[cp] > end_time:   # :455-466
        cp -= 1
    cp += 1
    cp = min(cp, cov_t.shape[0] - 1)
    ext = (q_normalize(EXTRINSICS[lidar][0]), np.array(EXTRINSICS[lidar][1]))
    return dict(pts=pts, beg_time=t0, end_time=end_time, ctrl_t=ctrl_t, ctrl_T=ctrl_T.reshape(-1, 16), imu_cov_t=cov_t,
                cov_pointer=int(cp), extrinsic=ext)
