"""Shared helpers for the test-suite (test infrastructure; may use oracle/)."""
import numpy as np

import pyoracle as po
from malio_b200 import capi, synth, plugin


def snapshot_for(case, churn=False, seed=7):
    """Map snapshot of the case.  With the real reference available: ikd-Tree Build (+ optional scripted churn:
    2% Add_Points with down-sampling, one Delete_Point_Boxes) flattened through malio_flatten.hpp.
    Otherwise the product's static builder.  Returns (MapSnapshot, RefTree | None)."""
    if po.ref_available():
        tree = po.RefTree(box_length=0.5)
        tree.build(case.map_xyz, case.map_normal_y)
        if churn:
            rng = np.random.default_rng(seed)
            M = case.map_xyz.shape[0]
            k = max(M // 50, 10)
            add = case.map_xyz[rng.integers(0, M, k)] + rng.normal(0, 0.3, (k, 3)).astype(np.float32)
            ny = rng.uniform(0.0005, 0.01, k).astype(np.float32)
            tree.add_points(add, ny, downsample=True)
            add2 = case.map_xyz[rng.integers(0, M, k // 4)] + rng.normal(0, 0.2, (k // 4, 3)).astype(np.float32)
            tree.add_points(add2, None, downsample=False)
            c = case.map_xyz[rng.integers(0, M)]
            tree.delete_boxes([[c[0] - 6, c[1] - 6, c[2] - 3, c[0] + 6, c[1] + 6, c[2] + 3]])
            tree.wait_rebuild()
        nodes, cov, ids, depth, live = tree.snapshot()
        return plugin.MapSnapshot(nodes, cov, ids, depth), tree
    snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
    return snap, None


def make_oracle(case, snap):
    orc = po.Oracle(case.params)
    orc.set_map_snapshot(snap.nodes, snap.node_cov)
    orc.set_scan(case.pts, case.table, case.table_off, case.temporal_comp)
    return orc


def make_model(case, snap, sort_queries=True, device=0):
    m = plugin.MeasurementModel(case.n_lidar, device=device, sort_queries=sort_queries, params=case.params)
    m.upload_map(snap)
    m.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp)
    return m


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
