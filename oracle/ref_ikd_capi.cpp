// oracle/_ref — the REAL reference ikd-Tree behind a small C API.  TEST INFRASTRUCTURE ONLY.
//
// This translation unit #includes the reference's own source where it lies
//   /root/reference/MA_LIO/include/ikd-Tree/ikd_Tree.cpp   (unmodified; -I points at it, nothing is copied)
// with the PCL shim of oracle/pcl_shim.  `private` is opened so that the harness can (a) wait for the
// background rebuild thread (Rebuild_Ptr, ikd_Tree.h:264) and (b) nothing else — every query goes
// through the public Nearest_Search / Build / Add_Points / Delete_Point_Boxes / flatten.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

#define private public
#include "ikd_Tree.cpp"   // resolved through -I/root/reference/MA_LIO/include/ikd-Tree
#undef private

#include "malio_flatten.hpp"
#include "malio_mapsync.hpp"

using Tree = KD_TREE<pcl::PointXYZINormal>;
using PV = Tree::PointVector;

static inline float id_to_float(int32_t id) { float f; std::memcpy(&f, &id, 4); return f; }
static inline int32_t float_to_id(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }

static PV make_points(const float* xyz, const float* normal_y, const int32_t* ids, int64_t n) {
  PV v(n);
  for (int64_t i = 0; i < n; ++i) {
    v[i].x = xyz[3 * i]; v[i].y = xyz[3 * i + 1]; v[i].z = xyz[3 * i + 2];
    v[i].normal_y = normal_y ? normal_y[i] : 0.001f;
    v[i].normal_z = id_to_float(ids ? ids[i] : (int32_t)i);   // id stashed bit-exact in an unused field
  }
  return v;
}

extern "C" {

void* ikdref_create(float delete_param, float balance_param, float box_length) {
  return new Tree(delete_param, balance_param, box_length);   // ~80 MB: Rebuild_Logger lives in the object
}
void ikdref_destroy(void* t) { delete (Tree*)t; }

void ikdref_build(void* t, const float* xyz, const float* normal_y, const int32_t* ids, int64_t n) {
  ((Tree*)t)->Build(make_points(xyz, normal_y, ids, n));
}
int ikdref_add_points(void* t, const float* xyz, const float* normal_y, const int32_t* ids, int64_t n,
                      int downsample_on) {
  PV v = make_points(xyz, normal_y, ids, n);
  return ((Tree*)t)->Add_Points(v, downsample_on != 0);
}
int ikdref_delete_boxes(void* t, const float* boxes /* nb x {min3,max3} */, int nb) {
  std::vector<BoxPointType> b(nb);
  for (int i = 0; i < nb; ++i)
    for (int k = 0; k < 3; ++k) { b[i].vertex_min[k] = boxes[6 * i + k]; b[i].vertex_max[k] = boxes[6 * i + 3 + k]; }
  return ((Tree*)t)->Delete_Point_Boxes(b);
}
// block until the background rebuild thread has no pending subtree (deterministic snapshots)
void ikdref_wait_rebuild(void* t) {
  Tree* tr = (Tree*)t;
  for (;;) {
    pthread_mutex_lock(&tr->rebuild_ptr_mutex_lock);
    bool idle = (tr->Rebuild_Ptr == nullptr);
    pthread_mutex_unlock(&tr->rebuild_ptr_mutex_lock);
    if (idle) break;
    usleep(1000);
  }
}
int ikdref_size(void* t) { return ((Tree*)t)->size(); }
int ikdref_validnum(void* t) { return ((Tree*)t)->validnum(); }

// k-NN through the reference's public Nearest_Search.  ids/-1, d2/inf padded; pts = k x {x,y,z,normal_y}
void ikdref_knn(void* t, const float* q, int64_t nq, int k, int32_t* out_ids, float* out_d2,
                float* out_pts, int32_t* out_found, int nthreads) {
  Tree* tr = (Tree*)t;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 256)
  for (int64_t i = 0; i < nq; ++i) {
    pcl::PointXYZINormal p;
    p.x = q[3 * i]; p.y = q[3 * i + 1]; p.z = q[3 * i + 2];
    PV near;
    std::vector<float> d2;
    tr->Nearest_Search(p, k, near, d2);
    int found = (int)near.size();
    if (out_found) out_found[i] = found;
    for (int j = 0; j < k; ++j) {
      bool ok = j < found;
      if (out_ids) out_ids[i * k + j] = ok ? float_to_id(near[j].normal_z) : -1;
      if (out_d2) out_d2[i * k + j] = ok ? d2[j] : INFINITY;
      if (out_pts) {
        out_pts[(i * k + j) * 4 + 0] = ok ? near[j].x : 0.f;
        out_pts[(i * k + j) * 4 + 1] = ok ? near[j].y : 0.f;
        out_pts[(i * k + j) * 4 + 2] = ok ? near[j].z : 0.f;
        out_pts[(i * k + j) * 4 + 3] = ok ? near[j].normal_y : 0.f;
      }
    }
  }
}

// single-query form with the signature the oracle's h_share_model restatement takes as its k-NN hook
int ikdref_knn1(void* t, const float q[3], int k, float* pts4, float* d2, int32_t* ids) {
  int32_t found = 0;
  ikdref_knn(t, q, 1, k, ids, d2, pts4, &found, 1);
  return found;
}

// live points in flatten() order (ikd_Tree.cpp:1638-1648); mutates lazy flags exactly as the reference does
int64_t ikdref_flatten_points(void* t, float* xyz, float* normal_y, int32_t* ids, int64_t cap) {
  Tree* tr = (Tree*)t;
  PV st;
  tr->flatten(tr->Root_Node, st, NOT_RECORD);
  int64_t n = (int64_t)st.size();
  for (int64_t i = 0; i < n && i < cap; ++i) {
    if (xyz) { xyz[3 * i] = st[i].x; xyz[3 * i + 1] = st[i].y; xyz[3 * i + 2] = st[i].z; }
    if (normal_y) normal_y[i] = st[i].normal_y;
    if (ids) ids[i] = float_to_id(st[i].normal_z);
  }
  return n;
}

// snapshot through include/malio_flatten.hpp (the product's host-side flattener, instantiated on the real node type)
int64_t ikdref_snapshot(void* t, malio_map_node* nodes, float* node_cov, int32_t* node_ids, int64_t cap,
                        uint32_t* max_depth, uint32_t* n_live_points) {
  Tree* tr = (Tree*)t;
  auto res = malio::flatten_ikdtree(
      tr->Root_Node, nodes, (uint32_t)cap, [&](const Tree::KD_TREE_NODE* n, uint32_t slot) {
        if (node_cov) node_cov[slot] = n->point.normal_y;
        if (node_ids) node_ids[slot] = float_to_id(n->point.normal_z);
      });
  if (max_depth) *max_depth = res.max_depth;
  if (n_live_points) *n_live_points = res.n_points;
  return res.overflow ? -1 : (int64_t)res.n_nodes;
}

// the compact flattener (16 B per node + the root's node_range_*) on the real node type
int64_t ikdref_snapshot_compact(void* t, malio_map_point* pts, float* node_cov, int64_t cap, uint32_t* max_depth, float* root_box) {
  Tree* tr = (Tree*)t;
  auto res = malio::flatten_ikdtree_compact(
      tr->Root_Node, pts, (uint32_t)cap, [&](const Tree::KD_TREE_NODE* n, uint32_t slot) { if (node_cov) node_cov[slot] = n->point.normal_y; },
      root_box);
  if (max_depth) *max_depth = res.max_depth;
  return res.overflow ? -1 : (int64_t)res.n_nodes;
}

// the parallel flatteners (same output, OpenMP)
int64_t ikdref_snapshot_parallel(void* t, malio_map_node* nodes, float* node_cov, int32_t* node_ids, int64_t cap,
                                 uint32_t* max_depth, uint32_t* n_live_points, uint32_t grain) {
  Tree* tr = (Tree*)t;
  auto res = malio::flatten_ikdtree_parallel(
      tr->Root_Node, nodes, (uint32_t)cap, [&](const Tree::KD_TREE_NODE* n, uint32_t slot) {
        if (node_cov) node_cov[slot] = n->point.normal_y;
        if (node_ids) node_ids[slot] = float_to_id(n->point.normal_z);
      }, grain);
  if (max_depth) *max_depth = res.max_depth;
  if (n_live_points) *n_live_points = res.n_points;
  return res.overflow ? -1 : (int64_t)res.n_nodes;
}
int64_t ikdref_snapshot_compact_parallel(void* t, malio_map_point* pts, float* node_cov, int64_t cap, uint32_t* max_depth,
                                         float* root_box, uint32_t grain) {
  Tree* tr = (Tree*)t;
  auto res = malio::flatten_ikdtree_compact_parallel(
      tr->Root_Node, pts, (uint32_t)cap, [&](const Tree::KD_TREE_NODE* n, uint32_t slot) { if (node_cov) node_cov[slot] = n->point.normal_y; },
      root_box, grain);
  if (max_depth) *max_depth = res.max_depth;
  return res.overflow ? -1 : (int64_t)res.n_nodes;
}

// Add_Points(.., true) followed by the product's collect_voxel_sync (include/malio_mapsync.hpp) on the real tree.
// Returns tmp_counter of Add_Points; the sync record is copied into caller arrays (capacities given), sizes in out_n[3] =
// {n_boxes, n_points, outside_own_box}.
static malio::VoxelSync g_sync;
int ikdref_add_points_synced(void* t, const float* xyz, const float* normal_y, const int32_t* ids, int64_t n, float ds, int64_t* out_n, int threads) {
  Tree* tr = (Tree*)t;
  PV v = make_points(xyz, normal_y, ids, n);
  const int c = tr->Add_Points(v, true);
  malio::collect_voxel_sync<Tree, BoxPointType>(*tr, v, ds, g_sync, [](const pcl::PointXYZINormal& p) { return float_to_id(p.normal_z); }, threads);
  out_n[0] = (int64_t)g_sync.counts.size(); out_n[1] = (int64_t)g_sync.normal_y.size(); out_n[2] = g_sync.outside_own_box;
  return c;
}
// collect only (after the caller ran Add_Points(.., true) with these points): what the product adds per scan on the host
void ikdref_collect_sync(void* t, const float* xyz, int64_t n, float ds, int64_t* out_n, int threads) {
  Tree* tr = (Tree*)t;
  PV v = make_points(xyz, nullptr, nullptr, n);
  malio::collect_voxel_sync<Tree, BoxPointType>(*tr, v, ds, g_sync, [](const pcl::PointXYZINormal& p) { return float_to_id(p.normal_z); }, threads);
  out_n[0] = (int64_t)g_sync.counts.size(); out_n[1] = (int64_t)g_sync.normal_y.size(); out_n[2] = g_sync.outside_own_box;
}
void ikdref_fetch_sync(float* boxes, uint32_t* counts, float* pxyz, float* pny, int32_t* pids) {
  if (boxes && !g_sync.boxes.empty()) std::memcpy(boxes, g_sync.boxes.data(), g_sync.boxes.size() * 4);
  if (counts && !g_sync.counts.empty()) std::memcpy(counts, g_sync.counts.data(), g_sync.counts.size() * 4);
  if (pxyz && !g_sync.xyz.empty()) std::memcpy(pxyz, g_sync.xyz.data(), g_sync.xyz.size() * 4);
  if (pny && !g_sync.normal_y.empty()) std::memcpy(pny, g_sync.normal_y.data(), g_sync.normal_y.size() * 4);
  if (pids && !g_sync.ids.empty()) std::memcpy(pids, g_sync.ids.data(), g_sync.ids.size() * 4);
}

int ikdref_node_bytes() { return (int)sizeof(Tree::KD_TREE_NODE); }

}  // extern "C"
