// malio_internal.h — shared between the CUDA translation unit and the host C++ of libmalio_b200.so.
#ifndef MALIO_INTERNAL_H_
#define MALIO_INTERNAL_H_

#include <cstdint>
#include <string>
#include <vector>

#include "malio_b200.h"

// Reduced system as the device produces it: fixed, L=3-shaped and padded so that one kernel serves L=1..3.
//   rows  : 24  (Jacobian columns, always at their L=3 positions: 0-5 | 6+3l | 15+3l)
//   cols  : 28  = 24 Jacobian columns | 1 residual column (-> H^T R^-1 h) | 3 "rho*h_0..2" columns
//           (-> the un-weighted 3x3 normal scatter that JacobiSVD's singular values come from)
// Only the 27 upper-triangular 4x4 blocks (row group i <= col group j) are computed and stored.
#define MALIO_RED_ROWS 24
#define MALIO_RED_COLS 28
#define MALIO_RED_BLOCKS 27
#define MALIO_RED_DOUBLES (MALIO_RED_BLOCKS * 16 + 2)   // + n_eff + spare

struct malio_handle {
  malio_config cfg{};
  std::string err;
  void* dev = nullptr;   // DeviceState*, owned by the CUDA TU
  void* pre = nullptr;   // PreState*, owned by malio_preproc.cu (undistortion / voxel grid buffers)
  void* mapst = nullptr; // MapOpsState*, owned by malio_mapops.cu (device-resident map replay)
  int want_prelaunch = 0; // set by malio_ieskf_update around malio_dev::measure: another pass may follow this one
};

// CUDA TU entry points used by the C-ABI wrappers
namespace malio_dev {
int create(malio_handle* h);
void destroy(malio_handle* h);
int upload_map(malio_handle* h, const malio_map_node* nodes, const float* cov, uint32_t n, uint32_t depth);
int upload_map_compact(malio_handle* h, const malio_map_point* pts, const float* cov, uint32_t n, uint32_t depth,
                       const float* root_box);
int download_map_nodes(malio_handle* h, malio_map_node* out, uint32_t cap);
int upload_scan(malio_handle* h, const malio_scan_pt* pts, uint32_t n, const malio_pose_entry* table,
                const uint32_t* table_off, const malio_rigid* tcomp);
int measure(malio_handle* h, const malio_pass_state* s, int redo_knn, double* HtRinvH, double* HtRinvh,
            malio_pass_stats* st);
int download_rows(malio_handle* h, double* h_x, double* hvec, uint32_t cap, uint32_t* n_rows);
int download_aux(malio_handle* h, float* normal_y, uint32_t* nn_idx, float* nn_d2, uint8_t* sel, float* world);
int knn(malio_handle* h, const float* q, uint32_t nq, uint32_t* idx, float* d2, float* ms);
int map_incremental(malio_handle* h, const malio_pass_state* s, double fs, int ekf_inited, uint8_t* cls, float* world);
int rearm_scan(malio_handle* h);
int reserve_scan(malio_handle* h, uint32_t n);
int cancel_prelaunch(malio_handle* h);   // a pass enqueued ahead of its state is told not to run (no-op when none is pending)
int update_on_device(malio_handle* h, malio_state* x, double* P, int max_iter, malio_update_report* rep, int* handled);
int grow_slots(malio_handle* h, uint32_t n, uint32_t keep);
int index_from_slots(malio_handle* h, uint32_t n_slots, const float box[6]);
int get_counters(malio_handle* h, malio_counters* out);
int set_timing(malio_handle* h, int enable);
int comm_init(malio_handle* h, const uint8_t* id, int rank, int world);
int get_unique_id(uint8_t* id);
}  // namespace malio_dev

// malio_solve.cu — the IESKF step on the device; ScanCtl is defined in malio_device.cuh (CUDA translation units only)
#ifdef __CUDACC__
namespace malio_devstate { struct ScanCtl; }
namespace malio_solve {
int setup(malio_handle* h);
int launch_init(malio_handle* h, cudaStream_t st, malio_devstate::ScanCtl* d_ctl, uint32_t seq0, int parity, uint32_t* d_bar);
int launch_solve(malio_handle* h, cudaStream_t st, malio_devstate::ScanCtl* d_ctl, const double* d_res, uint32_t* d_bar, double* h_out_dev,
                 uint32_t* h_flag_dev);
}  // namespace malio_solve
#endif

// malio_mapops.cu — the device-resident map
namespace malio_map {
void destroy(malio_handle* h);
int commit(malio_handle* h);   // pending deltas -> compaction (when worth it) + cell-list index; no-op when clean
int build(malio_handle* h, const float* xyz, const float* normal_y, const int32_t* ids, uint32_t n);
int add_points(malio_handle* h, const float* xyz, const float* normal_y, const int32_t* ids, uint32_t n);
int delete_boxes(malio_handle* h, const float* boxes, uint32_t nb, uint32_t* n_deleted);
int sync_voxels(malio_handle* h, const float* boxes, uint32_t nb, const float* xyz, const float* normal_y, const int32_t* ids, uint32_t m,
                uint32_t* n_deleted);
int info(malio_handle* h, uint32_t* n_live, uint32_t* n_slots);
int download(malio_handle* h, float* xyz, float* normal_y, int32_t* ids, uint32_t* slots, uint32_t cap, uint32_t* n);
}  // namespace malio_map

// malio_preproc.cu
namespace malio_pre {
void destroy(malio_handle* h);
int undistort(malio_handle* h, int lidar, const malio_raw_pt* pts, uint32_t n, const malio_undistort_args* a, float* xyz,
              int32_t* idx, uint8_t* ok, int32_t* pop_point, uint32_t* n_pops, double* pose);
int voxel_grid(malio_handle* h, int lidar, const float* in, uint32_t n, float leaf, float* out, uint32_t out_cap, uint32_t* n_out);
int upload_scan_device(malio_handle* h, const malio_pose_entry* table, const uint32_t* table_off, const malio_rigid* tcomp,
                       uint32_t* n_total);
}  // namespace malio_pre

// host math shared by measure() (localization weight) and the IESKF
namespace malio_host {
void sym3_singular_values(const double S[6] /* xx,xy,xz,yy,yz,zz */, double sv[3]);
}

#endif
