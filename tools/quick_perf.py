"""Ad-hoc device timing (not the bench contract): per-kernel CUDA-event times of the measurement passes."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
import numpy as np
from malio_b200 import synth, plugin, capi

def run(case, sort, reps=5):
    t0 = time.time(); snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y); tb = time.time() - t0
    m = plugin.MeasurementModel(case.n_lidar, sort_queries=sort, params=case.params)
    t0 = time.time(); m.upload_map(snap); tm = time.time() - t0
    print(f"{case.name}: snapshot build {tb:.2f}s depth {snap.max_depth}; upload_map {tm*1e3:.2f} ms; sort={sort}")
    for r in range(reps):
        t0 = time.time(); m.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp); ts = time.time() - t0
        x, P = case.x_prop.copy(), case.P_prop.copy()
        t0 = time.time(); rep = m.update_iterated_dyn_share_modified(x, P, case.max_iter); tu = time.time() - t0
        print(f"  rep {r}: upload_scan {ts*1e3:.3f} ms, update wall {tu*1e3:.3f} ms, passes {rep.passes} searches {rep.searches} "
              f"dev {rep.ms_device_total:.3f} ms host-solve {rep.ms_host_solve:.3f} ms n_eff {rep.n_eff_last}")
    m.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp)
    for conv in (True, False, True):
        ok, HTH, HTh, st = m.h_share_model(case.x_prop, conv)
        print(f"  pass search={conv}: knn {st.ms_knn*1e3:.1f} us plane {st.ms_plane*1e3:.1f} us reduce {st.ms_reduce*1e3:.1f} us total {st.ms_total*1e3:.1f} us n_eff {st.n_eff}")
    m.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "C2"
    case = {"C1": synth.case_C1, "C2": synth.case_C2, "C4": synth.case_C4}[which]()
    for sort in (True, False):
        run(case, sort)
