"""Ad-hoc: k-NN microbench (BASELINE config C5): 1M queries vs 10M-point tree, device time of the search."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
from malio_b200 import synth, plugin
xyz, q = synth.knn_microbench()
snap = plugin.build_static_snapshot(xyz)
for cell in (0.0, -1.0):
    m = plugin.MeasurementModel(1, knn_cell_size=cell)
    m.upload_map(snap)
    ts = []
    c0 = m.counters()
    for _ in range(4):
        idx, d2, ms = m.Nearest_Search(q)
        ts.append(ms)
    c1 = m.counters()
    print(f"C5 cell={cell}: k-NN device ms {['%.3f' % t for t in ts]}  fallback/search {(c1.knn_fallback_queries - c0.knn_fallback_queries) / 4:.0f} ring2/search {(c1.knn_ring2_queries - c0.knn_ring2_queries) / 4:.0f}")
    m.close()
