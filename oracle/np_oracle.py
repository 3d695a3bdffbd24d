"""Second, independent restatement (numpy / scipy) of the dense algebra of the hot path.  TEST INFRASTRUCTURE ONLY.

Purpose: cross-check the hand-rolled linear algebra of oracle_malio.cpp (5x3 column-pivoted QR, LU inverse,
3x3 eigen-solve, manifold operators) with library routines.  Follows the reference files:
  esti_plane                      MA_LIO/include/common_lib.h:144-190
  evalPointUncertainty            MA_LIO/include/associate_uct.hpp:145-175
  update_iterated_dyn_share_modified   MA_LIO/include/IKFoM_toolkit/esekfom/esekfom.hpp:495-721
  SO3 / S2 / A_matrix             MA_LIO/include/IKFoM_toolkit/mtk/{types/SOn.hpp,types/S2.hpp,src/mtkmath.hpp}
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

GRAV = 98090.0 / 10000.0
TOL = 1e-11


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def q_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx])


def q_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def q_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def so3_exp(v):   # SO3::exp: quaternion (cos(|v|/2), sinc(|v|/2)/2 * v)
    th = np.linalg.norm(v)
    if th < 1e-8:
        return np.array([1 - th * th / 8, *(0.5 * (1 - th * th / 24) * np.asarray(v))])
    return np.array([np.cos(th / 2), *(np.sin(th / 2) / th * np.asarray(v))])


def so3_log(q):   # MTK::log with scale 2, +-periodic
    nv = np.linalg.norm(q[1:])
    if nv < TOL:
        nv = TOL
    return 2.0 / nv * np.arctan(nv / q[0]) * q[1:]


def A_matrix(v):
    n = np.linalg.norm(v)
    if n < TOL:
        return np.eye(3)
    H = hat(v)
    return np.eye(3) + (1 - np.cos(n)) / n ** 2 * H + (1 - np.sin(n) / n) / n ** 2 * H @ H


def S2_Bx(v):
    L = GRAV
    if v[0] + L > TOL:
        d = L + v[0]
        return np.array([[-v[1], -v[2]], [L - v[1] * v[1] / d, -v[2] * v[1] / d], [-v[2] * v[1] / d, L - v[2] * v[2] / d]]) / L
    B = np.zeros((3, 2)); B[1, 1] = -1; B[2, 0] = 1
    return B


def S2_boxplus(v, d):
    Bu = S2_Bx(v) @ d
    return q_R(so3_exp(Bu)) @ v


def S2_boxminus(v, o):
    c = np.cross(v, o)
    v_sin, v_cos = np.linalg.norm(c), float(v @ o)
    theta = np.arctan2(v_sin, v_cos)
    if v_sin < TOL:
        return np.array([3.1415926, 0.0]) if abs(theta) > TOL else np.zeros(2)
    return theta / v_sin * S2_Bx(o).T @ hat(o) @ v


def S2_Nx_yy(v):
    return S2_Bx(v).T @ hat(v) / GRAV / GRAV


def S2_Mx(v, delta):
    Bx = S2_Bx(v)
    if np.linalg.norm(delta) < TOL:
        return -hat(v) @ Bx
    Bu = Bx @ delta
    return -np.eye(3) @ hat(v) @ A_matrix(Bu).T @ Bx    # exp_delta is the identity: scalar(1/2) == 0 (S2.hpp:287)


# ------------------------------------------------------------------ per-point pieces
def esti_plane(near4, threshold, cov_threshold):
    A = near4[:, :3].astype(np.float32)
    W = near4[:, 3].astype(np.float32)
    b = -np.ones(5, dtype=np.float32)
    cov_sum = float(np.sum(np.abs(cov_threshold - W.astype(np.float64))))
    plane_cov = 0.0
    if float(W[0]) > 0.00001:
        w = W.astype(np.float64)
        plane_cov = float(np.sum(((cov_threshold - w) / cov_sum) ** 2 * w))
    Q, R, P = scipy.linalg.qr(A, mode="economic", pivoting=True)
    y = Q.T @ b
    x = np.zeros(3, dtype=np.float32)
    x[P] = scipy.linalg.solve_triangular(R, y)
    n = np.float32(np.linalg.norm(x))
    pabcd = np.array([x[0] / n, x[1] / n, x[2] / n, np.float32(1.0 / float(n))], dtype=np.float32)
    ok = bool(np.all(np.abs(A @ pabcd[:3] + pabcd[3]) <= threshold))
    return ok, pabcd, plane_cov


def eval_point_uncertainty(p, T, cov):
    cov_input = np.zeros((9, 9))
    cov_input[:6, :6] = cov * 10000
    cov_input[6:, 6:] = np.eye(3) * 0.1
    p = np.asarray(p, dtype=np.float64)   # pi.x * distance_weight is float * double in the reference
    pc = np.array([p[0] * 0.05, p[1] * 0.05, p[2] * 0.05, 1.0])
    Tp = T @ pc
    G = np.zeros((4, 9))
    G[:3, :3] = Tp[3] * np.eye(3)
    G[:3, 3:6] = -hat(Tp[:3])
    D = np.zeros((4, 3)); D[:3, :3] = np.eye(3)
    G[:, 6:9] = T @ D
    return (G @ cov_input @ G.T)[:3, :3]


# ------------------------------------------------------------------ state + IESKF
class NpState:
    def __init__(self, s, L):
        self.L = L
        self.pos = np.array(s.pos[:])
        self.rot = np.array(s.rot[:])
        self.eq = [np.array(s.ext[l].q[:]) for l in range(L)]
        self.et = [np.array(s.ext[l].t[:]) for l in range(L)]
        self.vel, self.bg, self.ba = np.array(s.vel[:]), np.array(s.bg[:]), np.array(s.ba[:])
        self.grav = np.array(s.grav[:])

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def vec(self):
        out = list(self.pos) + list(self.rot)
        for l in range(self.L):
            out += list(self.eq[l]) + list(self.et[l])
        return np.array(out + list(self.vel) + list(self.bg) + list(self.ba) + list(self.grav))


def layout(L):
    offR = [6 + 3 * l for l in range(L)]
    offT = [6 + 3 * L + 3 * l for l in range(L)]
    vel = 6 + 6 * L
    return dict(n=17 + 6 * L, c=6 * (L + 1), rot=3, offR=offR, offT=offT, vel=vel, bg=vel + 3, ba=vel + 6, grav=vel + 9)


def boxminus(x, x0):
    ly = layout(x.L)
    d = np.zeros(ly["n"])
    d[0:3] = x.pos - x0.pos
    d[3:6] = so3_log(q_mul(q_conj(x0.rot), x.rot))
    for l in range(x.L):
        d[ly["offR"][l]:ly["offR"][l] + 3] = so3_log(q_mul(q_conj(x0.eq[l]), x.eq[l]))
        d[ly["offT"][l]:ly["offT"][l] + 3] = x.et[l] - x0.et[l]
    d[ly["vel"]:ly["vel"] + 3] = x.vel - x0.vel
    d[ly["bg"]:ly["bg"] + 3] = x.bg - x0.bg
    d[ly["ba"]:ly["ba"] + 3] = x.ba - x0.ba
    d[ly["grav"]:] = S2_boxminus(x.grav, x0.grav)
    return d


def boxplus(x, d):
    ly = layout(x.L)
    x.pos = x.pos + d[0:3]
    x.rot = q_mul(x.rot, so3_exp(d[3:6]))
    for l in range(x.L):
        x.eq[l] = q_mul(x.eq[l], so3_exp(d[ly["offR"][l]:ly["offR"][l] + 3]))
        x.et[l] = x.et[l] + d[ly["offT"][l]:ly["offT"][l] + 3]
    x.vel = x.vel + d[ly["vel"]:ly["vel"] + 3]
    x.bg = x.bg + d[ly["bg"]:ly["bg"] + 3]
    x.ba = x.ba + d[ly["ba"]:ly["ba"] + 3]
    x.grav = S2_boxplus(x.grav, d[ly["grav"]:])


def ieskf_update(measure, x, P, max_iter, R=0.001):
    """measure(x, converge) -> (valid, h_x[N,c], h[N], Rvec[N]).  Appendix B of SURVEY.md / esekfom.hpp:495-721."""
    ly = layout(x.L)
    n, c = ly["n"], ly["c"]
    x0, P0 = x.copy(), P.copy()
    t, converge = 0, True
    so3 = [ly["rot"]] + ly["offR"]
    dx_log = []
    for i in range(-1, max_iter):
        valid, h_x, h, Rv = measure(x, converge)
        if not valid:
            dx_log.append(None)
            continue
        dx = boxminus(x, x0)
        dxn = dx.copy()
        P = P0.copy()
        for idx in so3:
            J = A_matrix(dx[idx:idx + 3]).T
            dxn[idx:idx + 3] = J @ dxn[idx:idx + 3]
            P[idx:idx + 3, :] = J @ P[idx:idx + 3, :]
            P[:, idx:idx + 3] = P[:, idx:idx + 3] @ J.T
        g = ly["grav"]
        J2 = S2_Nx_yy(x.grav) @ S2_Mx(x0.grav, dx[g:g + 2])
        dxn[g:g + 2] = J2 @ dxn[g:g + 2]
        P[g:g + 2, :] = J2 @ P[g:g + 2, :]
        P[:, g:g + 2] = P[:, g:g + 2] @ J2.T
        m = h_x.shape[0]
        if n > m:
            Hc = np.zeros((m, n)); Hc[:, :c] = h_x
            K = P @ Hc.T @ np.linalg.inv(Hc @ P @ Hc.T / R + np.eye(m)) / R
            K_h = K @ h
            K_x = K @ Hc
        else:
            Rv = np.where(Rv < 0.0001, 0.001, Rv)
            HT = h_x.T / Rv
            HTH = HT @ h_x
            P_temp = np.linalg.inv(P)
            P_temp[:c, :c] += HTH
            P_inv = np.linalg.inv(P_temp)
            K_h = P_inv[:, :c] @ (HT @ h)
            K_x = np.zeros((n, n)); K_x[:, :c] = P_inv[:, :c] @ HTH
        dx_ = K_h + (K_x - np.eye(n)) @ dxn
        dx_log.append(dx_)
        boxplus(x, dx_)
        converge = bool(np.all(np.abs(dx_) <= 0.001))
        if converge:
            t += 1
        if t == 0 and i == max_iter - 2:
            converge = True
        if t > 1 or i == max_iter - 1:
            Lm = P.copy()
            for idx in so3:
                J = A_matrix(dx_[idx:idx + 3]).T
                Lm[idx:idx + 3, :] = J @ P[idx:idx + 3, :]
                K_x[idx:idx + 3, :c] = J @ K_x[idx:idx + 3, :c]
                Lm[:, idx:idx + 3] = Lm[:, idx:idx + 3] @ J.T
                P[:, idx:idx + 3] = P[:, idx:idx + 3] @ J.T
            J2 = S2_Nx_yy(x.grav) @ S2_Mx(x0.grav, dx_[g:g + 2])
            Lm[g:g + 2, :] = J2 @ P[g:g + 2, :]
            K_x[g:g + 2, :c] = J2 @ K_x[g:g + 2, :c]
            Lm[:, g:g + 2] = Lm[:, g:g + 2] @ J2.T
            P[:, g:g + 2] = P[:, g:g + 2] @ J2.T
            return x, Lm - K_x[:, :c] @ P[:c, :], dx_log
    return x, P, dx_log


# ------------------------------------------------------------------ pose-uncertainty table
# numpy restatement of MA_LIO/include/associate_uct.hpp:8-142 and the table loop of src/laserMapping.cpp:1028-1048.
class Pose:
    """struct Pose (common_lib.h:57-63)."""

    def __init__(self, q, t, cov):
        self.q = np.asarray(q, np.float64).copy()
        self.t = np.asarray(t, np.float64).copy()
        self.cov = np.asarray(cov, np.float64).reshape(6, 6).copy()
        self.T = np.eye(4)
        self.set_T()

    def set_T(self):
        self.T = np.eye(4)
        self.T[:3, :3] = q_R(self.q)
        self.T[:3, 3] = self.t

    def copy(self):
        p = Pose(self.q, self.t, self.cov)
        p.T = self.T.copy()
        return p


def q_rot(q, v):
    return q_R(q) @ np.asarray(v, np.float64)


def adjoint(T):   # adjointMatrix, associate_uct.hpp:8-15
    Ad = np.zeros((6, 6))
    Ad[:3, :3] = T[:3, :3]
    Ad[:3, 3:] = hat(T[:3, 3]) @ T[:3, :3]
    Ad[3:, 3:] = T[:3, :3]
    return Ad


def covop1(B):
    return -np.trace(B) * np.eye(3) + B


def covop2(B, C):
    return covop1(B) @ covop1(C) + covop1(C @ B)


def _compound_cov(c1p, c2):   # associate_uct.hpp:52-85 / 107-133
    c1rr, c1rp, c1pp = c1p[:3, :3], c1p[:3, 3:], c1p[3:, 3:]
    c2rr, c2rp, c2pp = c2[:3, :3], c2[:3, 3:], c2[3:, 3:]
    A1 = np.zeros((6, 6)); A2 = np.zeros((6, 6)); B = np.zeros((6, 6))
    A1[:3, :3] = covop1(c1pp); A1[:3, 3:] = covop1(c1rp + c1rp.T); A1[3:, 3:] = covop1(c1pp)
    A2[:3, :3] = covop1(c2pp); A2[:3, 3:] = covop1(c2rp + c2rp.T); A2[3:, 3:] = covop1(c2pp)
    Brr = covop2(c1pp, c2rr) + covop2(c1rp.T, c2rp) + covop2(c1rp, c2rp.T) + covop2(c1rr, c2pp)
    Brp = covop2(c1pp, c2rp.T) + covop2(c1rp.T, c2pp)
    Bpp = covop2(c1pp, c2pp)
    B[:3, :3] = Brr; B[:3, 3:] = Brp; B[3:, :3] = Brp.T; B[3:, 3:] = Bpp
    return c1p + c2 + (A1 @ c2 + c2 @ A1.T + A2 @ c1p + c1p @ A2.T) / 12 + B / 4


def compound_pose_with_cov(p1, cov1, p2, cov2, out):
    """compoundPoseWithCov, method 2; `out` may be p2 (aliasing as at laserMapping.cpp:1043): fields are touched in the
    reference's order, so adjointMatrix(pose_2.T_.inverse()) sees the new T_ in that case (associate_uct.hpp:99)."""
    cov1 = np.asarray(cov1, np.float64).reshape(6, 6).copy()
    cov2 = np.asarray(cov2, np.float64).reshape(6, 6).copy()
    q = q_mul(p1.q, p2.q)
    t = q_rot(p1.q, p2.t) + p1.t
    out.q = q
    out.t = t
    out.set_T()
    Ad = adjoint(np.linalg.inv(p2.T))
    c1p = Ad @ cov1 @ Ad.T
    out.cov = _compound_cov(c1p, cov2)
    return out.cov


def compound_inv_pose_with_cov(p1, cov1, p2, cov2, out):
    """compoundInvPoseWithCov, method 2 (associate_uct.hpp:29-86); returns cov_cp (the caller decides where it lives)."""
    cov1 = np.asarray(cov1, np.float64).reshape(6, 6).copy()
    cov2 = np.asarray(cov2, np.float64).reshape(6, 6).copy()
    qc = q_conj(p1.q)
    q = q_mul(qc, p2.q)
    t = q_rot(qc, p2.t - p1.t)
    out.q = q
    out.t = t
    out.set_T()
    Ad = adjoint(np.linalg.inv(out.T))
    c1p = Ad @ cov1 @ Ad.T
    return _compound_cov(c1p, cov2)


def build_pose_unc(extrinsic, temporal_comp, lidar_uncertainty):
    """laserMapping.cpp:1028-1048.  Returns per-LiDAR lists of Pose."""
    out = []
    for num, lst in enumerate(lidar_uncertainty):
        tab = []
        for i in range(len(lst) - 1):
            if num == 0:
                tab.append(lst[i].copy())
                continue
            pp = Pose([1, 0, 0, 0], [0, 0, 0], np.zeros((6, 6)))
            compound_pose_with_cov(extrinsic[num], extrinsic[num].cov, lst[i], lst[i].cov, pp)
            compound_pose_with_cov(temporal_comp[num - 1], temporal_comp[num - 1].cov, pp, pp.cov, pp)
            pp.cov = compound_inv_pose_with_cov(extrinsic[0], extrinsic[0].cov, pp, pp.cov, pp)
            tab.append(pp.copy())
        out.append(tab)
    return out
