"""Ad-hoc: where an e2e step goes (live churned ikd-Tree as the map): flatten / upload_map_compact / upload_scan / update / aux,
wall-clock per stage with a device synchronise after each; both root-box / depth sources."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bench
from malio_b200 import capi, plugin, synth
import pyoracle as po
case = synth.case_C2()
tree = bench.live_tree(case)
nodes, cov, ids, depth, live = tree.snapshot()
snap = plugin.MapSnapshot(nodes, cov, ids, depth)
M = snap.n_nodes
cap = M + M // 8
def pinned(a):
    t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory()
    v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape); v[...] = a
    return t, v
t1, h_mpts = pinned(np.zeros(cap, capi.MAP_POINT)); t2, h_cov = pinned(np.zeros(cap, np.float32)); t3, h_pts = pinned(case.pts)
h_mpts[:M] = plugin.compact_points(snap.nodes); h_cov[:M] = snap.node_cov
n0 = snap.nodes[0]
blo = np.array(n0["xyz"], np.float32); bhi = blo.copy()
for b, bit in ((n0["lbox"], capi.LINK_HAS_LEFT), (n0["rbox"], capi.LINK_HAS_RIGHT)):
    if n0["link"] & bit:
        blo = np.minimum(blo, b[0::2]); bhi = np.maximum(bhi, b[1::2])
root_box = np.stack([blo, bhi], axis=1).reshape(6).astype(np.float32)
m = plugin.MeasurementModel(3, params=case.params)
m.set_timing(False)
lib = po.ref_lib(); fd = C.c_uint32(0); fb = np.zeros(6, np.float32)
print("depth serial", depth, "root_box", root_box)
for mode in ("flatten", "snapshot_only", "flatten", "snapshot_only"):
    acc = {}
    for it in range(6):
        def lap(name, t0):
            torch.cuda.synchronize(); acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        if mode == "flatten":
            t0 = time.perf_counter(); n = lib.ikdref_snapshot_compact_parallel(tree.t, capi.ptr(h_mpts), capi.ptr(h_cov), cap, C.byref(fd), capi.ptr(fb), 16384); lap("flatten", t0)
            sp = plugin.MapSnapshot(None, h_cov[:n], None, int(fd.value)); box = fb; pts = h_mpts[:n]
        else:
            sp = plugin.MapSnapshot(None, h_cov[:M], None, snap.max_depth); box = root_box; pts = h_mpts[:M]
        t0 = time.perf_counter(); m.upload_map_compact(sp, points=pts, root_box=box); lap("upload_map", t0)
        t0 = time.perf_counter(); m.upload_scan(h_pts, case.table, case.table_off, case.temporal_comp); lap("upload_scan", t0)
        x, P = case.x_prop.copy(), case.P_prop.copy()
        t0 = time.perf_counter(); rep = m.update_iterated_dyn_share_modified(x, P, case.max_iter); lap("update", t0)
        t0 = time.perf_counter(); m.aux(nn_idx=False, nn_sqdist=False, world=False); lap("aux", t0)
    print(mode, "depth", sp.max_depth, "box", np.round(box, 2), {k: round(float(np.median(v[1:])), 3) for k, v in acc.items()}, "passes", rep.passes)
m.close(); tree.close()
