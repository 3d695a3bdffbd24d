"""Generates tests/golden/ikd_knn_golden.npz from the REAL reference ikd-Tree (oracle/_ref, i.e.
/root/reference/MA_LIO/include/ikd-Tree/ikd_Tree.cpp compiled in place).  Run in the build container only:

    python tests/golden/make_golden.py

Contents: a small map built with KD_TREE::Build, churned with Add_Points (down-sampling on and off) and
Delete_Point_Boxes so that lazy delete flags exist; the snapshot produced by include/malio_flatten.hpp from the
live tree; queries; and the reference's own Nearest_Search answers (ids stashed in the points, distances).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po  # noqa: E402
from malio_b200 import synth  # noqa: E402

rng = np.random.default_rng(2024)
xyz = synth.make_world(6000, seed=11)
ny = rng.uniform(0.0005, 0.01, xyz.shape[0]).astype(np.float32)
tree = po.RefTree(box_length=0.5)
tree.build(xyz, ny)
add = xyz[rng.integers(0, 6000, 600)] + rng.normal(0, 0.3, (600, 3)).astype(np.float32)
n_added = tree.add_points(add, rng.uniform(0.0005, 0.01, 600).astype(np.float32), downsample=True)
add2 = xyz[rng.integers(0, 6000, 200)] + rng.normal(0, 0.2, (200, 3)).astype(np.float32)
tree.add_points(add2, None, downsample=False)
c = xyz[100]
n_del = tree.delete_boxes([[c[0] - 4, c[1] - 4, c[2] - 2, c[0] + 4, c[1] + 4, c[2] + 2]])
tree.wait_rebuild()
nodes, cov, ids, depth, live = tree.snapshot()
q = np.concatenate([xyz[rng.integers(0, 6000, 700)] + rng.normal(0, 0.2, (700, 3)).astype(np.float32),
                    rng.uniform(-30, 30, (200, 3)).astype(np.float32) * np.float32([1, 1, 0.3]),
                    np.tile(c, (100, 1)) + rng.normal(0, 1.5, (100, 3)).astype(np.float32)]).astype(np.float32)
r_ids, r_d2, r_pts, r_found = tree.knn(q, 5)
flat_xyz, flat_ny, flat_ids = tree.flatten_points()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ikd_knn_golden.npz")
np.savez_compressed(out, nodes=nodes.view(np.uint8), node_cov=cov, node_ids=ids, max_depth=depth, n_live=live,
                    queries=q, ref_ids=r_ids, ref_d2=r_d2, ref_pts=r_pts, ref_found=r_found,
                    flat_ids=flat_ids, tree_size=tree.size(), tree_valid=tree.validnum(), n_added=n_added, n_deleted=n_del)
print("wrote", out, os.path.getsize(out), "bytes; nodes", nodes.shape[0], "live", live, "depth", depth,
      "added", n_added, "deleted", n_del, "size/valid", tree.size(), tree.validnum())
