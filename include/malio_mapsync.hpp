// malio_mapsync.hpp — host side of the device-resident map: keeping the GPU's copy of the map's
// POINT SET in step with the live ikd-Tree by deltas instead of flattening and re-uploading the tree every scan.
//
// north_star keeps "ikd-Tree incremental insert/delete on CPU".  The device does not need the tree's topology for the
// search (the k-NN fast path works on a cell list of the live points), only the set of live points.  What the
// reference's map-maintenance calls do to that set:
//   KD_TREE::Delete_Point_Boxes (ikd_Tree.cpp:648-676, Delete_by_range :785-857)   every live point with
//        min <= p < max (per axis) dies                      -> mirrored by malio_map_delete_boxes with the same boxes
//   KD_TREE::Add_Points(.., false) (:478-584, Add_by_point :982-1042)              every point is inserted
//                                                            -> mirrored by malio_map_add_points
//   KD_TREE::Add_Points(.., true)                                                   per new point: the live points of its
//        voxel box are searched (Search_by_range, :1257-1296), a down-sample winner is chosen, the box is emptied and the
//        winner re-inserted (:493-533).  WHICH stored point wins depends on the order Search_by_range visits them whenever
//        covariances are equal (the rule at :512-523 is not a total order) — i.e. on the tree's topology, which even depends
//        on the timing of the background re-build thread.  It therefore cannot be replayed without the tree; the host
//        tree stays the authority and the device is RE-SYNCHRONISED per touched voxel: after the reference's own
//        Add_Points call, collect_voxel_sync() asks the tree (public KD_TREE::Box_Search) for the content of every
//        distinct voxel the new points fall into and malio_map_sync_voxels() replaces the device's content of those
//        voxels with it.  Exact whatever the topology; host cost = one Box_Search per touched voxel (the reference's
//        Add_Points itself does one per new point), H2D = the points of the touched voxels.
//
// Works on any tree type with the reference's interface (KD_TREE<PointType>: Box_Search(const BoxPointType&, PointVector&),
// PointType with x, y, z, normal_y).  Nothing here is copied from the reference.
#ifndef MALIO_MAPSYNC_HPP_
#define MALIO_MAPSYNC_HPP_

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace malio {

struct VoxelSync {
  std::vector<float> boxes;        // nb x {min_x, min_y, min_z, max_x, max_y, max_z}   (half-open: min <= p < max)
  std::vector<uint32_t> counts;    // nb: live points of each box after the Add_Points call
  std::vector<float> xyz;          // sum(counts) x 3
  std::vector<float> normal_y;     // sum(counts)
  std::vector<int32_t> ids;        // sum(counts), filled when an id accessor is given
  uint32_t outside_own_box = 0;    // new points not inside the voxel box computed from them (float rounding for a
                                   // downsample_size that is not a power of two): those voxels cannot be re-synchronised
                                   // by box and the caller should fall back to a full snapshot for this scan
};

// Voxel box of a point exactly as KD_TREE::Add_Points computes it (ikd_Tree.cpp:493-498): float arithmetic.
inline void voxel_box_of(float x, float y, float z, float ds, float box[6]) {
  box[0] = std::floor(x / ds) * ds; box[3] = box[0] + ds;
  box[1] = std::floor(y / ds) * ds; box[4] = box[1] + ds;
  box[2] = std::floor(z / ds) * ds; box[5] = box[2] + ds;
}

// Call AFTER tree.Add_Points(added, true).  BoxT = the tree's BoxPointType (vertex_min[3], vertex_max[3]).
// threads > 1 runs the Box_Search calls of the distinct voxels concurrently (OpenMP).  Search_by_range only WRITES to the
// tree through Push_Down, and only on nodes with a pending lazy flag; the Add_Points call that precedes this one walked
// exactly these boxes (Search_by_range, Delete_by_range, Add_by_point all push the flags down on their way), so the
// read-back finds none pending and is read-only in practice — the same situation as the reference's own multi-threaded
// Nearest_Search (laserMapping.cpp:559-563).  Callers that prefer not to rely on that keep threads = 1.
template <class Tree, class BoxT, class PointVector, class IdFn>
void collect_voxel_sync(Tree& tree, const PointVector& added, float downsample_size, VoxelSync& out, IdFn id_of, int threads = 1) {
  out.boxes.clear(); out.counts.clear(); out.xyz.clear(); out.normal_y.clear(); out.ids.clear(); out.outside_own_box = 0;
  // ---- distinct voxels, in order of first appearance: open-addressing table over the packed voxel coordinates
  size_t cap = 64;
  while (cap < added.size() * 2 + 16) cap <<= 1;
  struct Key { int32_t a, b, c; };
  std::vector<Key> slot_key(cap);
  std::vector<uint8_t> slot_used(cap, 0);
  out.boxes.reserve(added.size() * 6);
  for (size_t i = 0; i < added.size(); ++i) {
    float box[6];
    voxel_box_of(added[i].x, added[i].y, added[i].z, downsample_size, box);
    if (!(box[0] <= added[i].x && box[3] > added[i].x && box[1] <= added[i].y && box[4] > added[i].y && box[2] <= added[i].z && box[5] > added[i].z))
      out.outside_own_box++;
    const Key k{(int32_t)std::floor(added[i].x / downsample_size), (int32_t)std::floor(added[i].y / downsample_size),
                (int32_t)std::floor(added[i].z / downsample_size)};
    size_t s = ((size_t)(uint32_t)k.a * 73856093u ^ (size_t)(uint32_t)k.b * 19349663u ^ (size_t)(uint32_t)k.c * 83492791u) & (cap - 1);
    bool dup = false;
    while (slot_used[s]) {
      if (slot_key[s].a == k.a && slot_key[s].b == k.b && slot_key[s].c == k.c) { dup = true; break; }
      s = (s + 1) & (cap - 1);
    }
    if (dup) continue;
    slot_used[s] = 1; slot_key[s] = k;
    out.boxes.insert(out.boxes.end(), box, box + 6);
  }
  const int64_t nb = (int64_t)(out.boxes.size() / 6);
  out.counts.assign((size_t)nb, 0u);
  if (threads < 1) threads = 1;
  if ((int64_t)threads > nb) threads = nb > 0 ? (int)nb : 1;
  // ---- one Box_Search per voxel; thread t owns the contiguous box range [nb t / T, nb (t+1) / T) and appends to buffers of its
  // own, so that concatenating the buffers in thread order is box order (deterministic whatever the thread count)
  struct Local { std::vector<float> xyz, ny; std::vector<int32_t> ids; };
  std::vector<Local> loc((size_t)threads);
#pragma omp parallel num_threads(threads) if (threads > 1)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), T = omp_get_num_threads();
#else
    const int t = 0, T = 1;
#endif
    Local& L = loc[(size_t)t];
    PointVector found;
    const int64_t k0 = nb * t / T, k1 = nb * (t + 1) / T;
    L.xyz.reserve((size_t)(k1 - k0) * 6); L.ny.reserve((size_t)(k1 - k0) * 2); L.ids.reserve((size_t)(k1 - k0) * 2);
    for (int64_t k = k0; k < k1; ++k) {
      BoxT b;
      for (int a = 0; a < 3; ++a) { b.vertex_min[a] = out.boxes[6 * (size_t)k + a]; b.vertex_max[a] = out.boxes[6 * (size_t)k + 3 + a]; }
      tree.Box_Search(b, found);
      out.counts[(size_t)k] = (uint32_t)found.size();
      for (const auto& p : found) {
        L.xyz.push_back(p.x); L.xyz.push_back(p.y); L.xyz.push_back(p.z);
        L.ny.push_back(p.normal_y);
        L.ids.push_back(id_of(p));
      }
    }
  }
  size_t total = 0;
  for (const Local& L : loc) total += L.ny.size();
  out.xyz.reserve(total * 3); out.normal_y.reserve(total); out.ids.reserve(total);
  for (const Local& L : loc) {
    out.xyz.insert(out.xyz.end(), L.xyz.begin(), L.xyz.end());
    out.normal_y.insert(out.normal_y.end(), L.ny.begin(), L.ny.end());
    out.ids.insert(out.ids.end(), L.ids.begin(), L.ids.end());
  }
}

}  // namespace malio

#endif  // MALIO_MAPSYNC_HPP_
