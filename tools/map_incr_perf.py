"""Ad-hoc: wall time of malio_map_incremental on C2 (100k points) after an update."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
import numpy as np
from malio_b200 import synth, plugin
case = synth.case_C2()
snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
m = plugin.MeasurementModel(case.n_lidar, params=case.params)
m.upload_map_compact(snap)
m.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp)
x, P = case.x_prop.copy(), case.P_prop.copy()
m.update_iterated_dyn_share_modified(x, P, case.max_iter)
for world in (True, False):
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); cls, w = m.map_incremental(x, 0.5, True, world=world); ts.append(time.perf_counter() - t0)
    print(f"map_incremental world={world}: median {1e3 * np.median(ts):.3f} ms  classes {np.bincount(cls, minlength=4).tolist()}")
m.close()
