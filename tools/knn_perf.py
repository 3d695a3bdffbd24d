"""Ad-hoc k-NN timing (not the bench contract): search-pass device times for a case, index on/off, a few cell sizes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
import numpy as np
from malio_b200 import synth, plugin

which = sys.argv[1] if len(sys.argv) > 1 else "C2"
cells = [float(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.0, -1.0]
case = {"C1": synth.case_C1, "C2": synth.case_C2, "C4": synth.case_C4}[which]()
snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
for cell in cells:
    m = plugin.MeasurementModel(case.n_lidar, sort_queries=True, params=case.params, knn_cell_size=cell)
    import time
    t0 = time.time(); m.upload_map(snap); t1 = time.time(); m.upload_map(snap); t2 = time.time()
    m.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp)
    ts = []
    c0 = m.counters()
    for i in range(6):
        ok, HTH, HTh, st = m.h_share_model(case.x_prop, True)
        ts.append(st.ms_knn * 1e3)
    c1 = m.counters()
    print(f"{which} cell={cell}: knn us {['%.1f' % t for t in ts]}  fallback/search {(c1.knn_fallback_queries - c0.knn_fallback_queries) / 6:.1f} "
          f"ring2/search {(c1.knn_ring2_queries - c0.knn_ring2_queries) / 6:.1f}  upload_map {1e3 * (t1 - t0):.2f} / {1e3 * (t2 - t1):.2f} ms  n_eff {st.n_eff}")
    m.close()
