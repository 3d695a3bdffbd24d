"""numpy mirror of the device-resident map's delta protocol (include/malio_mapsync.hpp).  TEST INFRASTRUCTURE ONLY.

A flat table of (xyz, normal_y, id, live) that receives exactly the calls the device receives:
  build / add_points (plain append) / delete_boxes (half-open boxes) / sync_voxels (replace the content of each box).
tests/test_mapops_cpu.py drives the REAL ikd_Tree.cpp (oracle/_ref) and this mirror with the same call stream and checks
after every scan that the mirror's live set equals KD_TREE::flatten() — i.e. the protocol is exact whatever the tree's
topology; the GPU tests then check the device against the same stream."""
import numpy as np

F = np.float32


class MirrorMap:
    def __init__(self):
        self.xyz = np.zeros((0, 3), F); self.ny = np.zeros(0, F); self.ids = np.zeros(0, np.int64); self.live = np.zeros(0, bool)

    def build(self, xyz, normal_y, ids):
        self.xyz = np.array(xyz, F).reshape(-1, 3).copy(); self.ny = np.array(normal_y, F).copy()
        self.ids = np.array(ids, np.int64).copy(); self.live = np.ones(len(self.ids), bool)

    def add_points(self, xyz, normal_y, ids):
        self.xyz = np.concatenate([self.xyz, np.array(xyz, F).reshape(-1, 3)]); self.ny = np.concatenate([self.ny, np.array(normal_y, F)])
        self.ids = np.concatenate([self.ids, np.array(ids, np.int64)]); self.live = np.concatenate([self.live, np.ones(len(ids), bool)])

    def _inside(self, b):
        return self.live & np.all((b[:3] <= self.xyz) & (b[3:] > self.xyz), axis=1)

    def delete_boxes(self, boxes):
        n = 0
        for b in np.array(boxes, F).reshape(-1, 6):
            m = self._inside(b); n += int(m.sum()); self.live[m] = False
        return n

    def sync_voxels(self, sync):
        kill = np.zeros(len(self.live), bool)
        for b in sync["boxes"]:
            kill |= self._inside(b)
        self.live[kill] = False
        self.add_points(sync["xyz"], sync["normal_y"], sync["ids"])

    def live_points(self):
        return self.xyz[self.live], self.ny[self.live], self.ids[self.live]
