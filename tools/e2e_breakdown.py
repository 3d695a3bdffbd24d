"""Ad-hoc: wall time of the per-scan host calls (pinned buffers), C2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
import numpy as np, torch
from malio_b200 import synth, plugin
case = synth.case_C2()
snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
def pinned(a):
    t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory(); v = t.numpy().view(a.dtype).reshape(a.shape); v[...] = a; return t, v
k1, h_nodes = pinned(snap.nodes); k2, h_cov = pinned(snap.node_cov); k3, h_pts = pinned(case.pts); k4, h_mp = pinned(plugin.compact_points(snap.nodes))
sp = plugin.MapSnapshot(h_nodes, h_cov, snap.node_ids, snap.max_depth)
m = plugin.MeasurementModel(case.n_lidar, params=case.params); m.set_timing(False)
def t(f, n=8):
    f(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))
print("upload_map (68 B/node)      ms", t(lambda: m.upload_map(sp)))
print("upload_map_compact (20 B)   ms", t(lambda: m.upload_map_compact(sp, points=h_mp)))
rb = np.array([-1e4, 1e4, -1e4, 1e4, -1e3, 1e3], np.float32)
n0 = snap.nodes[0]; rb = np.stack([np.minimum(np.minimum(n0["lbox"][0::2], n0["rbox"][0::2]), n0["xyz"]), np.maximum(np.maximum(n0["lbox"][1::2], n0["rbox"][1::2]), n0["xyz"])], axis=1).reshape(6).astype(np.float32)
print("upload_map_compact + box    ms", t(lambda: m.upload_map_compact(sp, points=h_mp, root_box=rb)))
print("upload_scan                 ms", t(lambda: m.upload_scan(h_pts, case.table, case.table_off, case.temporal_comp)))
def upd():
    m.rearm_scan(); x, P = case.x_prop.copy(), case.P_prop.copy(); m.update_iterated_dyn_share_modified(x, P, case.max_iter)
print("rearm + update              ms", t(upd))
print("aux (normal_y, selected)    ms", t(lambda: m.aux(normal_y=True, nn_idx=False, nn_sqdist=False, selected=True, world=False)))
