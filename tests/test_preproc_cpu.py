"""CPU tier for the rows next to the hot path: the oracle's restatements against independent
implementations (scipy expm/logm, a numpy voxel filter, a numpy replay of the covariance-list walk), and the product's
host-side spline pose against the oracle.  No GPU needed."""
import numpy as np
import pytest
from scipy.linalg import expm, logm

import pyoracle as po
from malio_b200 import capi, plugin, synth


def _quat_to_R(q):
    return synth.q_to_R(q)


def _scipy_get_pose(ct, cT, ts):
    """BsplineSE3::get_pose with generic matrix exp/log (independent of the closed forms of quat_ops.h)."""
    T = cT.reshape(-1, 4, 4)
    i1 = int(np.searchsorted(ct, ts, side="right")) - 1
    i0, i2, i3 = i1 - 1, i1 + 1, i1 + 2
    u = (ts - ct[i1]) / (ct[i2] - ct[i1])
    b0 = (5 + 3 * u - 3 * u * u + u ** 3) / 6
    b1 = (1 + 3 * u + 3 * u * u - 2 * u ** 3) / 6
    b2 = u ** 3 / 6
    A = [expm(b * np.real(logm(np.linalg.inv(T[k]) @ T[k + 1]))) for b, k in ((b0, i0), (b1, i1), (b2, i2))]
    return T[i0] @ A[0] @ A[1] @ A[2]


def test_spline_pose_oracle_vs_scipy_and_product_host_code():
    c = synth.undistort_case(100)
    rng = np.random.default_rng(3)
    for ts in np.concatenate([rng.uniform(c["beg_time"], c["end_time"], 40), [c["ctrl_t"][3], c["ctrl_t"][5] + 1e-9]]):
        ok, q, p = po.bspline_get_pose(c["ctrl_t"], c["ctrl_T"], ts)
        assert ok
        P = _scipy_get_pose(c["ctrl_t"], c["ctrl_T"], ts)
        assert np.abs(_quat_to_R(q) - P[:3, :3]).max() < 1e-9 and np.abs(p - P[:3, 3]).max() < 1e-9
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        ok2, q2, p2 = plugin.bspline_get_pose(c["ctrl_t"], c["ctrl_T"], ts)
        assert ok2 and np.abs(q - q2).max() < 1e-13 and np.abs(p - p2).max() < 1e-13
    # outside the supported span both fail: before the second control point, on/after the second-to-last one
    for ts in (c["ctrl_t"][0] + 1e-4, c["ctrl_t"][0] - 1.0, c["ctrl_t"][-2] + 1e-4, c["ctrl_t"][-1] + 1.0):
        assert not po.bspline_get_pose(c["ctrl_t"], c["ctrl_T"], ts)[0]
        assert not plugin.bspline_get_pose(c["ctrl_t"], c["ctrl_T"], ts)[0]


def test_covariance_list_walk_is_a_min_plus_prefix_scan():
    """IMU_Processing.hpp:476-486 pops at most one entry per point.  The device computes the walk as
    pops_s = s + min(1, min_{j<=s}(need_j - j)) over the reversed point order; replayed here in numpy against the oracle's
    sequential loop, including IMU rates far above the point rate (the pointer lags and catches up one pop per point)."""
    for n, hz in ((500, 200.0), (300, 5000.0), (50, 20000.0), (4000, 1000.0)):
        c = synth.undistort_case(n, imu_hz=hz)
        ok, q, p = po.bspline_get_pose(c["ctrl_t"], c["ctrl_T"], c["end_time"])
        r = po.undistort(c["pts"], c["beg_time"], c["extrinsic"], (q, p), c["ctrl_t"], c["ctrl_T"], c["imu_cov_t"], c["cov_pointer"])
        t = c["pts"]["curvature"].astype(np.float64) / 1000.0 + c["beg_time"]
        top = c["imu_cov_t"][: c["cov_pointer"] + 1]
        need = np.array([np.count_nonzero(top > ti) for ti in t])
        s = np.arange(n - 1)
        v = need[::-1][: n - 1] - s            # reversed order: s = 0 is the last point
        m = np.minimum(1, np.minimum.accumulate(v))
        pops = s + m
        idx = np.full(n, capi.IDX_UNTOUCHED, np.int64)
        idx[(n - 1 - s)] = pops - 1
        touched = r["ok"] > 0
        assert np.array_equal(idx[touched], r["idx"][touched].astype(np.int64)), (n, hz)
        assert r["n_pops"] == pops[-1]
        first = np.flatnonzero(np.diff(np.concatenate([[0], pops])) > 0)
        assert np.array_equal(n - 1 - first[: len(r["pop_point"])], r["pop_point"])


def _numpy_voxel_grid(p, leaf):
    il = np.float32(1.0) / np.float32(leaf)
    fin = np.isfinite(p[:, :3]).all(1)
    q = p[fin]
    mn = np.floor(q[:, :3].min(0) * il).astype(np.int64)
    mx = np.floor(q[:, :3].max(0) * il).astype(np.int64)
    d = mx - mn + 1
    ijk = (np.floor(q[:, :3] * il) - mn.astype(np.float32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * d[0] + ijk[:, 2] * d[0] * d[1]
    order = np.argsort(key, kind="stable")
    keys, starts, counts = np.unique(key[order], return_index=True, return_counts=True)
    out = np.zeros((len(keys), p.shape[1]), np.float32)
    for k, (s, c) in enumerate(zip(starts, counts)):
        acc = np.zeros(p.shape[1], np.float32)
        for row in q[order[s:s + c]]:
            acc = acc + row
        out[k] = acc / np.float32(c)
    return out


def test_voxel_grid_oracle_vs_numpy():
    rng = np.random.default_rng(5)
    p = np.zeros((6000, 8), np.float32)
    p[:, :3] = rng.uniform(-12, 15, (6000, 3))
    p[:, 3] = rng.integers(0, 20, 6000)
    p[:, 7] = rng.uniform(0, 100, 6000)
    p[17, 0] = np.nan
    p[99, 2] = np.inf
    for leaf in (0.5, 0.2, 2.0):
        out, vo = po.voxel_grid(p, leaf)
        ref = _numpy_voxel_grid(p, leaf)
        assert out.shape == ref.shape and np.array_equal(out, ref)
        assert vo[17] == -1 and vo[99] == -1 and vo.max() == out.shape[0] - 1
        # every output point lies in its own voxel, voxels are distinct
        assert len(np.unique(vo[vo >= 0])) == out.shape[0]
    # one point, and all points in one voxel
    out, _ = po.voxel_grid(p[:1], 0.5)
    assert np.array_equal(out, p[:1])
    q = p[:50].copy(); q[:, :3] = 0.1 + 0.01 * rng.uniform(size=(50, 3)).astype(np.float32)
    out, _ = po.voxel_grid(q, 0.5)
    assert out.shape[0] == 1


def test_city_bin_readers_and_preprocess_handlers(tmp_path):
    """N4: the product's host readers / handlers against the numpy restatement on synthetic files written in the player's
    packed layouts (17-byte Livox, 22-byte Ouster records), with and without the eof() extra record, decimation, blind
    zone, tag / line filters and a truncated trailing record."""
    import np_dataset as nd
    from malio_b200 import dataset
    rng = np.random.default_rng(8)
    n = 4000
    lv = np.zeros(n, nd.LIVOX_REC)
    for k in "xyz":
        lv[k] = rng.uniform(-30, 30, n).astype(np.float32)
    lv["x"][:50] = 0.01; lv["y"][:50] = 0.02; lv["z"][:50] = 0.01       # inside the blind zone
    lv["x"][100:104] = lv["x"][99]; lv["y"][100:104] = lv["y"][99]     # repeated returns: only z differs
    lv["reflectivity"] = rng.integers(0, 255, n); lv["tag"] = rng.choice([0x00, 0x10, 0x20, 0x30, 0x11], n); lv["line"] = rng.integers(0, 8, n)
    lv["t16"] = np.sort(rng.integers(0, 65535, n))
    f1 = tmp_path / "1.bin"
    f1.write_bytes(lv.tobytes() + b"\x01\x02\x03")                       # 3 stray bytes: a partial record is dropped
    ou = np.zeros(n, nd.OUSTER_REC)
    for k in "xyz":
        ou[k] = rng.uniform(-60, 60, n).astype(np.float32)
    ou["x"][:30] = 0.1; ou["y"][:30] = 0.1; ou["z"][:30] = 0.1
    ou["intensity"] = rng.uniform(0, 3000, n); ou["ring"] = rng.integers(0, 128, n); ou["t"] = np.sort(rng.integers(0, 100_000_000, n))
    f2 = tmp_path / "2.bin"
    f2.write_bytes(ou.tobytes())
    for quirk in (True, False):
        a = dataset.read_livox_bin(str(f1), quirk)
        b = nd.read_records(str(f1), nd.LIVOX_REC, quirk)
        assert a.shape[0] == b.shape[0] == n + (1 if quirk else 0)
        assert np.array_equal(a["xyz"], np.stack([b["x"], b["y"], b["z"]], 1)) and np.array_equal(a["offset_time"], b["t16"].astype(np.uint32))
        assert np.array_equal(a["tag"], b["tag"]) and np.array_equal(a["line"], b["line"]) and np.array_equal(a["reflectivity"], b["reflectivity"])
        c = dataset.read_ouster_bin(str(f2), quirk)
        d = nd.read_records(str(f2), nd.OUSTER_REC, quirk)
        assert c.shape[0] == d.shape[0] and np.array_equal(c["t"], d["t"]) and np.array_equal(c["ring"], d["ring"])
        assert np.array_equal(c["intensity"], d["intensity"])
        for pf in (1, 3):
            g, gi = dataset.preprocess_livox(a, n_scans=6, point_filter_num=pf, blind=0.5)
            r, ri = nd.avia_handler(b, 6, pf, 0.5)
            assert g.shape[0] == r.shape[0] > 0
            assert np.array_equal(g["xyz"], r[:, :3]) and np.array_equal(g["curvature"], r[:, 3]) and np.array_equal(gi, ri)
            g, gi = dataset.preprocess_ouster(c, point_filter_num=pf, blind=0.5, time_unit_scale=1e-3)
            r, ri = nd.oust64_handler(d, pf, 0.5, 1e-3)
            assert g.shape[0] == r.shape[0] > 0
            assert np.array_equal(g["xyz"], r[:, :3]) and np.array_equal(g["curvature"], r[:, 3]) and np.array_equal(gi, ri)
    with pytest.raises(capi.MalioError):
        dataset.read_livox_bin(str(tmp_path / "missing.bin"))


def test_city_bin_reader_edge_cases(tmp_path):
    """Empty file, a file shorter than one record, capacity below the record count, NULL arguments."""
    import ctypes as C
    import np_dataset as nd
    from malio_b200 import dataset
    empty = tmp_path / "empty.bin"; empty.write_bytes(b"")
    short = tmp_path / "short.bin"; short.write_bytes(b"\x00" * 10)
    for f in (empty, short):
        assert dataset.read_ouster_bin(str(f), True).shape[0] == 1       # only the player's extra default record
        assert dataset.read_ouster_bin(str(f), False).shape[0] == 0
        assert dataset.read_livox_bin(str(f), False).shape[0] == 0
    rec = np.zeros(10, nd.OUSTER_REC); rec["x"] = np.arange(10)
    f = tmp_path / "ten.bin"; f.write_bytes(rec.tobytes())
    lib = capi.load()
    n = C.c_uint32(0)
    out = np.zeros(4, capi.OUSTER_PT)
    assert lib.malio_read_ouster_bin(str(f).encode(), capi.ptr(out), 4, C.byref(n), 0) == capi.ERR_CAPACITY
    assert n.value == 4 and np.array_equal(out["xyz"][:, 0], np.arange(4, dtype=np.float32))    # what fitted was delivered
    out = np.zeros(10, capi.OUSTER_PT)
    assert lib.malio_read_ouster_bin(str(f).encode(), capi.ptr(out), 10, C.byref(n), 1) == capi.ERR_CAPACITY   # no room for the extra record
    assert lib.malio_read_ouster_bin(str(f).encode(), capi.ptr(out), 10, C.byref(n), 0) == capi.OK and n.value == 10
    assert lib.malio_read_ouster_bin(None, capi.ptr(out), 10, C.byref(n), 0) == capi.ERR_INVALID_ARG
    assert lib.malio_read_ouster_bin(str(f).encode(), capi.ptr(out), 10, None, 0) == capi.ERR_INVALID_ARG
    # handlers: nothing in, nothing out; an invalid decimation is rejected
    g, gi = dataset.preprocess_ouster(np.zeros(0, capi.OUSTER_PT))
    assert g.shape[0] == 0 and gi.shape[0] == 0
    m = C.c_uint32(0)
    assert lib.malio_preprocess_ouster(capi.ptr(out), 10, 0, C.c_double(0.5), C.c_float(1e-3), None, None, 0, C.byref(m)) == capi.ERR_INVALID_ARG
