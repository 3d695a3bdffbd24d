"""Ad-hoc A/B of k-NN kernel variants (MALIO_LIB_PATH): device time of the search on C2 (100k queries vs 1M points, through a full
measurement pass) and C5 (1M queries vs 10M points)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
import numpy as np
from malio_b200 import synth, plugin
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("both", "C2"):
    case = synth.case_C2()
    snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
    m = plugin.MeasurementModel(case.n_lidar, sort_queries=True, params=case.params)
    m.upload_map(snap); m.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp)
    ts = []
    for i in range(8):
        ok, HTH, HTh, st = m.h_share_model(case.x_prop, True)
        ts.append(st.ms_knn * 1e3)
    print(f"C2 knn us: {np.median(ts[2:]):.1f}  (all {['%.1f' % t for t in ts]})")
    m.close()
if which in ("both", "C5"):
    xyz, q = synth.knn_microbench()
    snap = plugin.build_static_snapshot(xyz)
    m = plugin.MeasurementModel(1)
    m.upload_map(snap)
    ts = [m.Nearest_Search(q)[2] for _ in range(6)]
    print(f"C5 knn ms: {np.median(ts[2:]):.4f}")
    m.close()
