"""Workload for ncu: C2, search passes only (k-NN fast path + list kernel)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
from malio_b200 import synth, plugin
which = sys.argv[1] if len(sys.argv) > 1 else "C2"
cell = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
case = {"C1": synth.case_C1, "C2": synth.case_C2, "C4": synth.case_C4}[which]()
snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
m = plugin.MeasurementModel(case.n_lidar, sort_queries=True, params=case.params, knn_cell_size=cell)
m.upload_map(snap)
m.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp)
for i in range(3):
    ok, HTH, HTh, st = m.h_share_model(case.x_prop, True)
    print(f"pass {i}: knn {st.ms_knn*1e3:.1f} us plane {st.ms_plane*1e3:.1f} us reduce {st.ms_reduce*1e3:.1f} us")
m.close()
