/*
 * malio_b200.h — C-ABI of the B200-native MA-LIO measurement hot path.
 *
 * Drop-in boundary.  The reference's plug-in point is the IKFoM
 * measurement-model callback
 *     typedef void measurementModel_dyn_share(state&, dyn_share_datastruct<double>&)
 *         MA_LIO/include/IKFoM_toolkit/esekfom/esekfom.hpp:130   (installed :152-168, invoked :512)
 * implemented by
 *     void h_share_model(state_ikfom&, esekfom::dyn_share_datastruct<double>&)
 *         MA_LIO/src/laserMapping.cpp:552-760
 * which in turn calls
 *     KD_TREE::Nearest_Search(PointType, int k, PointVector&, vector<float>&, float)
 *         MA_LIO/include/ikd-Tree/ikd_Tree.h:328 / ikd_Tree.cpp:426-461  (call site laserMapping.cpp:586)
 * and whose N_eff x 24 output is reduced by
 *     esekf::update_iterated_dyn_share_modified        esekfom.hpp:495-721  (O(N) part :621-637).
 *
 * Because the device path fuses the H^T R^-1 H / H^T R^-1 h reduction, this ABI sits one
 * level above the callback: malio_measure() replaces h_share_model *and* esekfom.hpp:622-635
 * and returns the reduced c x c system.  malio_ieskf_update() is the host-side iterated
 * update (replaces esekfom.hpp:495-721 as a whole).
 *
 * Conventions: plain C types only; caller owns every host buffer (pinned memory recommended);
 * the handle owns device memory, streams, events and the NCCL communicator.  Every call returns
 * a malio_status; calls on one handle must come from one thread at a time; all calls are
 * synchronous at return unless stated otherwise.
 */
#ifndef MALIO_B200_H_
#define MALIO_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MALIO_MAX_LIDAR 3   /* use-ikfom.hpp:14-27 builds the manifold for 3 LiDARs; L is run-time here (1..3) */
#define MALIO_K 5           /* NUM_MATCH_POINTS, common_lib.h:22 */
#define MALIO_MAX_COLS (6 * (MALIO_MAX_LIDAR + 1))            /* c = 6(L+1), laserMapping.cpp:642 */
#define MALIO_MAX_DOF (17 + 6 * MALIO_MAX_LIDAR)              /* n = 17+6L, esekfom.hpp:158 */
#define MALIO_MAX_TREE_DEPTH 96                              /* traversal stack bound; deeper snapshots are rejected */
#define MALIO_NCCL_UNIQUE_ID_BYTES 128

typedef enum malio_status {
  MALIO_OK = 0,
  MALIO_ERR_INVALID_ARG = 1,
  MALIO_ERR_CUDA = 2,          /* CUDA runtime / driver failure (no GPU, OOM, launch error) */
  MALIO_ERR_NCCL = 3,
  MALIO_ERR_NO_EFFECTIVE_POINTS = 4, /* N_eff == 0: reference sets valid=false (laserMapping.cpp:635-639) */
  MALIO_ERR_STATE = 5,         /* call order violated (e.g. measure before upload) */
  MALIO_ERR_TREE_TOO_DEEP = 6,
  MALIO_ERR_CAPACITY = 7
} malio_status;

/* ---- map snapshot: one record per live ikd-Tree node, DFS pre-order ---------------------------
 * Mirrors what KD_TREE::Search reads per visit (ikd_Tree.cpp:1073-1255): the node's point, its
 * point_deleted flag, and both children's node_range_* boxes (KD_TREE_NODE, ikd_Tree.h:59-82),
 * after Push_Down (ikd_Tree.cpp:1371-1466).  tree_deleted subtrees are not exported (Search
 * returns immediately on them, :1075).  The left child of node i, when present, is node i+1. */
typedef struct malio_map_node {
  float x, y, z;
  uint32_t link;      /* bits 0..27 index of the right child; bit 31 has_left; bit 30 has_right;
                         bit 29 point_deleted */
  float lbox[6];      /* left child's {x_min,x_max,y_min,y_max,z_min,z_max}; undefined if !has_left */
  float rbox[6];      /* right child's */
} malio_map_node;     /* 64 bytes */

#define MALIO_LINK_HAS_LEFT  0x80000000u
#define MALIO_LINK_HAS_RIGHT 0x40000000u
#define MALIO_LINK_POINT_DELETED 0x20000000u
#define MALIO_LINK_INDEX_MASK 0x0FFFFFFFu

/* Compact snapshot record = the first 16 bytes of malio_map_node.  With malio_upload_map_compact() the host ships only
 * these (plus the map-side weights): 20 bytes per node instead of 68, and the device rebuilds both children's boxes
 * of every node bottom-up as the tight bounding box of the live (not point_deleted) points below — which is what
 * KD_TREE::Update maintains in node_range_* (ikd_Tree.cpp:1469-1635). */
typedef struct malio_map_point {
  float x, y, z;
  uint32_t link;      /* as in malio_map_node */
} malio_map_point;    /* 16 bytes */

/* ---- scan point: feats_down_body entry as h_share_model reads it (laserMapping.cpp:565-570,694) */
typedef struct malio_scan_pt {
  float x, y, z;        /* point in its own LiDAR frame */
  uint16_t lidar;       /* PointType::intensity after laserMapping.cpp:975 */
  uint16_t table_idx;   /* int(PointType::normal_x), laserMapping.cpp:694,737 */
} malio_scan_pt;        /* 16 bytes */

/* ---- one entry of pose_unc[l][j] (struct Pose, common_lib.h:57-63): only T_ and cov_ are read by
 * evalPointUncertainty (associate_uct.hpp:153-175) */
typedef struct malio_pose_entry {
  double T[16];        /* row-major 4x4 */
  double cov[36];      /* row-major 6x6 */
} malio_pose_entry;

typedef struct malio_rigid {
  double q[4];         /* w, x, y, z */
  double t[3];
} malio_rigid;

/* ---- pose part of state_ikfom read by h_share_model (use-ikfom.hpp:14-27; extrinsic_update,
 * laserMapping.cpp:291-308) */
typedef struct malio_pass_state {
  double rot[4];       /* s.rot, (w,x,y,z) */
  double pos[3];       /* s.pos */
  malio_rigid ext[MALIO_MAX_LIDAR];   /* offset_R_l / offset_T_l */
} malio_pass_state;

/* ---- run-time parameters (parameters.cpp:17-66; launch/mapping_city.launch:9-15; City.yaml:41-50) */
typedef struct malio_params {
  int32_t n_lidar;            /* L, 1..3 */
  int32_t extrinsic_est_en;   /* laserMapping.cpp:681 */
  float plane_th;             /* esti_plane threshold, common_lib.h:184 */
  float knn_max_sqdist;       /* 5.0f, laserMapping.cpp:587 */
  double cov_threshold;
  double point_cov_max, point_cov_min;
  double plane_cov_max, plane_cov_min;
  double localize_cov_max, localize_cov_min;
  double localize_thresh_max, localize_thresh_min;
  double range_min, range_max;
} malio_params;

typedef struct malio_config {
  malio_params params;
  int32_t device;             /* CUDA device ordinal */
  int32_t sort_queries;       /* 1: Morton-order the scan on the device (internal; outputs stay in caller order) */
  uint32_t max_points;        /* capacity hints; buffers grow on demand when 0 */
  uint32_t max_map_nodes;
  float knn_cell_size;        /* k-NN fast path: edge of the cell-list index built over the snapshot at upload
                                 (metres).  0 = automatic (starts at 1 m = 2 x filter_size_map), < 0 = index off:
                                 every query walks the flattened ikd-Tree.  Results are identical either way. */
} malio_config;

typedef struct malio_pass_stats {
  uint32_t n_points;          /* N of this rank's shard */
  uint32_t n_eff;             /* effct_feat_num, summed over ranks */
  int32_t valid;              /* dyn_share.valid */
  int32_t searched;           /* 1 if this pass ran the k-NN */
  double u_min, u_max;        /* min/max_unit_cov, laserMapping.cpp:615-628 */
  double tau_min, tau_max;    /* min/max_cov, laserMapping.cpp:646-703 */
  double sigma[3];            /* singular values of h_x[:,0:3] before the localization weight */
  double loc_weight;          /* weight, laserMapping.cpp:749-756 */
  float ms_sort;              /* Morton keys + sort of the scan (first pass of a scan only, else 0) */
  float ms_knn, ms_plane, ms_reduce, ms_total;   /* device time of this pass (CUDA events); ms_knn = the k-NN kernel alone */
} malio_pass_stats;

typedef struct malio_handle malio_handle;

/* fill p with the reference defaults for L LiDARs (City.yaml / mapping_city.launch values) */
void malio_default_params(malio_params* p, int n_lidar);

int malio_create(malio_handle** out, const malio_config* cfg);
void malio_destroy(malio_handle* h);
const char* malio_last_error(const malio_handle* h);   /* static storage when h == NULL */
const char* malio_version(void);

/* multi-GPU: one process per GPU.  Rank 0 calls malio_get_nccl_unique_id and ships the bytes to the
 * other ranks (torch.distributed / MPI / a socket — plumbing); every rank then calls malio_comm_init.
 * With a communicator attached, malio_measure all-reduces {min,max} and the reduced system over the
 * ranks; each rank uploads only its own block of scan points. */
int malio_get_nccl_unique_id(uint8_t id[MALIO_NCCL_UNIQUE_ID_BYTES]);
int malio_comm_init(malio_handle* h, const uint8_t id[MALIO_NCCL_UNIQUE_ID_BYTES], int rank, int world);

/* once per scan: flattened ikd-Tree snapshot (include/malio_flatten.hpp produces it) and the map-side
 * weight normal_y of every node (quirk: an input field).  root is node 0. */
int malio_upload_map(malio_handle* h, const malio_map_node* nodes, const float* node_cov,
                     uint32_t n_nodes, uint32_t max_depth);

/* Same snapshot in compact form (see malio_map_point): 3.4x less host->device traffic per scan.
 * root_box (optional, may be NULL): {x_min,x_max,y_min,y_max,z_min,z_max} bounding every live point — the ikd-Tree
 * root's node_range_* — saves the read-back of the rebuilt root record. */
int malio_upload_map_compact(malio_handle* h, const malio_map_point* pts, const float* node_cov,
                             uint32_t n_nodes, uint32_t max_depth, const float* root_box);
/* test / debug: read back the device-resident 64-byte node records (boxes included). */
int malio_download_map_nodes(malio_handle* h, malio_map_node* out, uint32_t capacity);

/* once per scan: this rank's block of the down-sampled merged scan + the pose tables.
 * table holds pose_unc[0], pose_unc[1], ... back to back; pose_unc[l] = table[table_off[l] .. table_off[l+1]).
 * temporal_comp[l-1] is kf.temporal_comp[l-1] (IMU_Processing.hpp:517-519), l = 1..L-1. */
int malio_upload_scan(malio_handle* h, const malio_scan_pt* pts, uint32_t n_pts,
                      const malio_pose_entry* table, const uint32_t* table_off /* L+1 */,
                      const malio_rigid* temporal_comp /* L-1, may be NULL when L==1 */);

/* re-arm the scan that is already resident on the device (same points / tables as the last malio_upload_scan):
 * resets the per-scan state (Nearest_Points, point_selected_surf, internal ordering) without any host copy.
 * Lets a caller time the device path with inputs resident in HBM. */
int malio_rearm_scan(malio_handle* h);

/* one pass of h_share_model + esekfom.hpp:622-635.
 *   HtRinvH : c x c row-major, = h_x^T diag(1/R) h_x   (HTH, esekfom.hpp:629)
 *   HtRinvh : c,              = h_x^T diag(1/R) h     (HT * dyn_share.h, esekfom.hpp:635)
 * redo_knn = dyn_share.converge on entry (laserMapping.cpp:583).
 * Returns MALIO_ERR_NO_EFFECTIVE_POINTS (stats->valid = 0) when N_eff == 0. */
int malio_measure(malio_handle* h, const malio_pass_state* s, int redo_knn,
                  double* HtRinvH, double* HtRinvh, malio_pass_stats* stats);

/* rows of the last pass for the degenerate branch n > N_eff (esekfom.hpp:574-582): up to cap rows of
 * h_x (row-major, c columns) and h, in the library's internal (cell-sorted) point order — with a communicator attached the
 * rows of all ranks, concatenated in rank order — scaled by the plane weight (laserMapping.cpp:714-715) but NOT yet by
 * the localization weight: multiply by stats->loc_weight (laserMapping.cpp:758-759), as malio_ieskf_update does. */
int malio_download_rows(malio_handle* h, double* h_x, double* hvec, uint32_t cap, uint32_t* n_rows);

/* side outputs of the last pass, caller order; any pointer may be NULL.
 *   normal_y : trace of the point covariance for all N points (laserMapping.cpp:699,730,741)
 *   nn_idx   : N x 5 snapshot node indices, ascending by (distance, x) as Nearest_Search returns them;
 *              0xFFFFFFFF where fewer than 5 were found
 *   nn_sqdist: N x 5
 *   selected : point_selected_surf after the pass
 *   world    : N x 3 feats_down_world (laserMapping.cpp:576-578) */
int malio_download_aux(malio_handle* h, float* normal_y, uint32_t* nn_idx, float* nn_sqdist,
                       uint8_t* selected, float* world);

/* ---- next to the path: map_incremental's per-point decision ------------------------
 * laserMapping.cpp:398-446, evaluated on the device from what the last pass left there (Nearest_Points of the last
 * search, normal_y) and the state AFTER the update (pointBodyToWorld, laserMapping.cpp:134-147).  Caller order.
 *   cls[i] = MALIO_MAP_SKIP   normal_y > cov_threshold (:406)
 *            MALIO_MAP_ADD    goes to PointToAdd            -> ikdtree.Add_Points(.., true)   (:434, :437, :440)
 *            MALIO_MAP_ADD_NO_DOWNSAMPLE  PointNoNeedDownsample -> ikdtree.Add_Points(.., false)  (:421-425)
 *            MALIO_MAP_DROP   a neighbour is closer to the voxel centre (:428-433)
 *   world : N x 3 feats_down_world for the points that were not skipped (zeros for skipped ones); may be NULL
 * filter_size_map = filter_size_map_min, ekf_inited = flg_EKF_inited (laserMapping.cpp:989). */
#define MALIO_MAP_SKIP 0
#define MALIO_MAP_ADD 1
#define MALIO_MAP_ADD_NO_DOWNSAMPLE 2
#define MALIO_MAP_DROP 3
int malio_map_incremental(malio_handle* h, const malio_pass_state* s, double filter_size_map, int ekf_inited,
                          uint8_t* cls, float* world);

/* ---- next to the path: the pose-uncertainty table that feeds malio_upload_scan -------------
 * struct Pose (common_lib.h:57-63) and Barfoot's 4th-order SE(3) covariance compounding
 * (associate_uct.hpp:29-142), host C++.  Output arguments MAY alias the second input pose exactly as the reference's
 * call sites do (laserMapping.cpp:1042-1044): the functions touch the fields in the reference's order, so the
 * aliasing behaviour (adjointMatrix sees the already overwritten T_, associate_uct.hpp:99) is reproduced. */
typedef struct malio_pose {
  double q[4];        /* q_, (w,x,y,z) */
  double t[3];        /* t_ */
  double T[16];       /* T_, row-major 4x4 */
  double cov[36];     /* cov_, row-major 6x6 */
} malio_pose;

/* PoseInitial (common_lib.h:129-142) */
void malio_pose_initial(malio_pose* pose, const double t[3], const double q[4], const double cov[36]);
/* compoundPoseWithCov, method 2 (associate_uct.hpp:88-142); also sets pose_cp->cov (:134) */
void malio_compound_pose_with_cov(const malio_pose* pose_1, const double cov_1[36], const malio_pose* pose_2,
                                  const double cov_2[36], malio_pose* pose_cp, double cov_cp[36]);
/* compoundInvPoseWithCov, method 2 (associate_uct.hpp:29-86); does NOT touch pose_cp->cov unless cov_cp points to it */
void malio_compound_inv_pose_with_cov(const malio_pose* pose_1, const double cov_1[36], const malio_pose* pose_2,
                                      const double cov_2[36], malio_pose* pose_cp, double cov_cp[36]);
/* pose_unc of one scan (laserMapping.cpp:1028-1048) in the layout malio_upload_scan takes.
 *   extrinsic[l]            struct Pose of LiDAR l's extrinsic, l < n_lidar
 *   temporal_comp[l-1]      kf.temporal_comp[l-1], l = 1..n_lidar-1 (may be NULL when n_lidar == 1)
 *   lidar_uncertainty[l]    kf.lidar_uncertainty[l][0 .. counts[l]); the LAST entry of each list is dropped (:1035,:1040)
 * table must hold sum(counts[l] - 1) entries; table_off gets n_lidar + 1 offsets.  Returns the number of entries,
 * or a negative malio_status. */
int malio_build_pose_unc(int n_lidar, const malio_pose* extrinsic, const malio_pose* temporal_comp,
                         const malio_pose* const* lidar_uncertainty, const uint32_t* counts,
                         malio_pose_entry* table, uint32_t* table_off);

/* ---- next to the path: per-raw-point B-spline undistortion ------------------------------------
 * ImuProcess::UndistortPcl's point loop (IMU_Processing.hpp:468-508) for ONE LiDAR: per raw point the spline pose
 * BsplineSE3::get_pose (BsplineSE3.cpp:84-118: 3 x exp_se3(b * log_se3(.)), quat_ops.h:190-243), the motion compensation
 * into the LiDAR's scan-end frame (:492) and the walk of the IMU-covariance list that yields the point's uncertainty-table
 * index (:476-486, written to `intensity` :496).  On the device the log_se3 of every control-point pair is taken once,
 * each point evaluates three exp_se3, and the sequential one-pop-per-point walk is computed as a min-plus prefix scan.
 * Points arrive in buffer order (ascending time); like the reference the loop runs from the last point down to the
 * SECOND one — point 0 is never touched.  Where get_pose fails the point keeps its coordinates and `intensity`. */
typedef struct malio_raw_pt {
  float x, y, z;
  float curvature;      /* time offset from the scan start in ms (preprocess.cpp:91,139,200) */
} malio_raw_pt;

typedef struct malio_undistort_args {
  double beg_time;            /* meas.lidar_beg_time[lid_num - num - 1] (:474) */
  malio_rigid extrinsic;      /* extrinsic_quat[num], extrinsic_trans[num] */
  malio_rigid lt_imu_frame;   /* lt_imu_frame_quat[num], lt_imu_frame_trans[num]: IMU pose at this LiDAR's scan end */
  const double* ctrl_t;       /* spline control points: timestamps ascending (BsplineSE3::control_points keys) */
  const double* ctrl_T;       /* ... and their 4x4 poses, row-major, n_ctrl x 16 */
  uint32_t n_ctrl;            /* <= MALIO_MAX_CTRL */
  const double* imu_cov_t;    /* imu_cov[k].first.first, ascending */
  uint32_t n_cov;             /* <= MALIO_MAX_COV */
  int32_t cov_pointer;        /* its value when the point loop starts (:455-466) */
} malio_undistort_args;
#define MALIO_MAX_CTRL 512
#define MALIO_MAX_COV 512
#define MALIO_IDX_UNTOUCHED INT32_MIN

/* lidar = slot (0..L-1) whose device-resident result malio_voxel_grid(.., NULL input) picks up.  Host outputs (any may be
 * NULL): xyz n x 3 float; idx n (the value the reference writes to `intensity`, MALIO_IDX_UNTOUCHED where it writes nothing);
 * ok n (spline_flag); pop_point[k] = index of the point at which the k-th table entry is pushed (:487-494), n_pops their
 * number (the host builds those <= n_cov entries with malio_bspline_get_pose + malio_compound_*); pose n x 7 (q wxyz, p). */
int malio_undistort(malio_handle* h, int lidar, const malio_raw_pt* pts, uint32_t n, const malio_undistort_args* a,
                    float* xyz, int32_t* idx, uint8_t* ok, int32_t* pop_point, uint32_t* n_pops, double* pose);

/* BsplineSE3::get_pose on the host (lt_imu_frame poses, the <= n_cov table-entry poses).  q = (w,x,y,z).  1 on success. */
int malio_bspline_get_pose(const double* ctrl_t, const double* ctrl_T, uint32_t n_ctrl, double timestamp, double q[4], double p[3]);

/* ---- next to the path (second half): pcl::VoxelGrid down-sampling (laserMapping.cpp:968-983) --------
 * PCL's filter with setLeafSize(leaf, leaf, leaf) and default settings restated on the device: voxel index
 * floor(x / leaf) relative to the cloud's minimum, one output point per occupied voxel in ascending voxel index, every
 * field the centroid of the voxel's points (float sums taken in ascending input index).  What the hot path reads of the
 * result is x, y, z and the averaged `intensity` (the table index, moved to normal_x at :975).
 * in == NULL: take the device-resident output of the last malio_undistort(h, lidar, ...) (x, y, z, intensity = idx,
 * curvature; points the undistortion left untouched keep intensity 0).  in != NULL: n x 5 floats
 * {x, y, z, intensity, curvature} from the host.  Outputs (may be NULL): out n_out x 5 floats, same layout; *n_out.
 * The down-sampled cloud also stays on the device as LiDAR `lidar`'s part of the next scan (malio_upload_scan_device). */
int malio_voxel_grid(malio_handle* h, int lidar, const float* in, uint32_t n, float leaf, float* out, uint32_t out_cap,
                     uint32_t* n_out);

/* Merge the device-resident down-sampled clouds of LiDARs 0..L-1 (feats_down_body = sum of feats_down_vec[num],
 * laserMapping.cpp:983; lidar id -> `intensity`, int(averaged intensity) -> table_idx) into the scan of the handle without
 * a host bounce; then as malio_upload_scan.  n_total (optional) receives the merged size. */
int malio_upload_scan_device(malio_handle* h, const malio_pose_entry* table, const uint32_t* table_off,
                             const malio_rigid* temporal_comp, uint32_t* n_total);

/* ---- next to the path: City dataset .bin scans without ROS -------------------------------------
 * One file = one scan of packed records as the reference's file player reads them (file_player/src/ROSThread.cpp:776-795
 * Livox, :952-967 Ouster), followed by the per-sensor conversion of MA_LIO/src/preprocess.cpp to the time-stamped raw
 * cloud (x, y, z, curvature = time offset in ms) that UndistortPcl / malio_undistort consume.  Host code. */
typedef struct malio_livox_pt {   /* livox_ros_driver::CustomPoint fields the player fills */
  float x, y, z;
  uint8_t reflectivity, tag, line, pad;
  uint32_t offset_time;           /* only 2 bytes per record are stored in the file (ROSThread.cpp:789) */
} malio_livox_pt;
typedef struct malio_ouster_pt {  /* OusterPointXYZIRT */
  float x, y, z, intensity;
  uint16_t ring, pad;
  uint32_t t;
} malio_ouster_pt;
/* out may be NULL to count.  eof_quirk != 0 appends the zero record the player's `while(!file.eof())` loop produces. */
int malio_read_livox_bin(const char* path, malio_livox_pt* out, uint32_t cap, uint32_t* n_out, int eof_quirk);
int malio_read_ouster_bin(const char* path, malio_ouster_pt* out, uint32_t cap, uint32_t* n_out, int eof_quirk);
/* Preprocess::avia_handler (preprocess.cpp:59-110): n_scans = N_SCANS[lidar], point_filter_num, blind as in the YAML */
int malio_preprocess_livox(const malio_livox_pt* pts, uint32_t n, int n_scans, int point_filter_num, double blind, malio_raw_pt* out,
                           float* intensity, uint32_t cap, uint32_t* n_out);
/* Preprocess::oust64_handler (preprocess.cpp:112-152) */
int malio_preprocess_ouster(const malio_ouster_pt* pts, uint32_t n, int point_filter_num, double blind, float time_unit_scale,
                            malio_raw_pt* out, float* intensity, uint32_t cap, uint32_t* n_out);

/* ---- next to the path: the map as a device-resident point set, kept in step by deltas -------------
 * Instead of flattening the live ikd-Tree and re-uploading ~20 B per node every scan, the device keeps the map's live
 * points in slots {x, y, z, normal_y, id} and receives what the reference's map-maintenance calls change
 * (include/malio_mapsync.hpp has the host side and the reasoning):
 *   malio_map_build          <- KD_TREE::Build(points)                         ikd_Tree.cpp:370-424  (laserMapping.cpp:1007)
 *   malio_map_delete_boxes   <- KD_TREE::Delete_Point_Boxes(boxes)             :648-676              (laserMapping.cpp:223)
 *   malio_map_add_points     <- KD_TREE::Add_Points(points, false)             :478-584              (laserMapping.cpp:444)
 *   malio_map_sync_voxels    <- KD_TREE::Add_Points(points, true) followed by malio::collect_voxel_sync on the tree: the
 *                               content of every touched voxel box replaces the device's                (laserMapping.cpp:443)
 * Boxes are half-open, min <= p < max per axis, exactly Search_by_range's / Delete_by_range's test (:1270,:807).
 * The cell-list index is rebuilt on the device (~70 us at 1M points) by the next malio_measure / malio_knn / malio_map_commit;
 * dead slots are compacted away when they exceed half of the slots.  In this mode the search never walks a tree: every
 * neighbour list and every distance is the reference's, except for exact distance ties, which the reference breaks by
 * traversal order — those queries are counted (malio_counters.knn_tie_queries) and broken by slot index.
 * nn_idx of malio_download_aux / malio_knn are SLOT indices in this mode; malio_map_download gives slot -> id.
 * A later malio_upload_map* call switches the handle back to snapshot mode.
 * Host arrays are consumed when a call returns.  Per-scan-sized batches (<= 65536 points, <= 16384 boxes) pass through a pinned
 * bounce buffer and the call returns WITHOUT waiting for the device; n_deleted (may be NULL) costs a device round trip when asked. */
int malio_map_build(malio_handle* h, const float* xyz, const float* normal_y, const int32_t* ids /* may be NULL: 0..n-1 */, uint32_t n);
int malio_map_add_points(malio_handle* h, const float* xyz, const float* normal_y, const int32_t* ids /* may be NULL */, uint32_t n);
int malio_map_delete_boxes(malio_handle* h, const float* boxes /* nb x {min3, max3} */, uint32_t nb, uint32_t* n_deleted);
int malio_map_sync_voxels(malio_handle* h, const float* boxes, uint32_t nb, const float* xyz, const float* normal_y,
                          const int32_t* ids, uint32_t n_points, uint32_t* n_deleted);
int malio_map_commit(malio_handle* h);
int malio_map_info(malio_handle* h, uint32_t* n_live, uint32_t* n_slots);
/* live points in slot order (test / debug): any of xyz (n x 3), normal_y, ids, slots may be NULL */
int malio_map_download(malio_handle* h, float* xyz, float* normal_y, int32_t* ids, uint32_t* slots, uint32_t cap, uint32_t* n);

/* cumulative counters since malio_create: kernels launched by this library, k-NN kernel launches, queries they
 * processed and their summed device time (CUDA events) — what bench.py's roofline is computed from. */
typedef struct malio_counters {
  uint64_t kernel_launches;
  uint64_t knn_launches;
  uint64_t knn_queries;
  double knn_ms;
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t knn_fallback_queries;  /* queries the cell-list fast path handed to the exact ikd-Tree-order traversal */
  uint64_t knn_ring2_queries;     /* queries that needed the 5x5x5 cell block */
  uint64_t knn_candidates;        /* map points the 3x3x3 scans looked at (summed only while timing is enabled) */
  uint64_t pass_launches;         /* measurement passes timed (fused pass kernel, or gate+reduce+fold on the NCCL path) */
  uint64_t pass_points;           /* scan points those passes processed */
  uint64_t pass_fit_launches;     /* ... of which ran the plane fit (search passes) */
  double pass_ms;                 /* their summed device time (CUDA events; only while timing is enabled) */
  uint64_t knn_tie_queries;       /* device-resident map mode: queries with two of the six best distances within PointType_CMP's
                                     1e-10 window, broken by slot index instead of the reference's traversal order (exempted ties) */
  uint64_t map_slots, map_live;   /* device-resident map: slots in use / live points after the last commit */
  uint64_t map_compactions;
  /* multi-GPU, in-kernel exchange: device time (globaltimer) block 0 spent pushing its min/max keys to the peers and waiting for
   * all of theirs (this includes waiting for the slowest rank to LAUNCH its pass), the same for the system sum, and the number
   * of passes that exchanged */
  double exchange_min_wait_ms, exchange_sum_wait_ms;
  uint64_t exchange_passes;
} malio_counters;
int malio_get_counters(malio_handle* h, malio_counters* out);

/* per-pass CUDA-event timing (the ms_* fields of malio_pass_stats and knn_ms of the counters); on by default,
 * costs ~7 event records per pass on the host when enabled. */
int malio_set_timing(malio_handle* h, int enable);

/* stand-alone k-NN (BASELINE config C5, the microbench): queries are world-frame points. */
int malio_knn(malio_handle* h, const float* queries_xyz, uint32_t n_queries,
              uint32_t* nn_idx, float* nn_sqdist, float* ms_device);

/* ---- host IESKF -----------------------------------------------------------------------------
 * state_ikfom for L LiDARs (use-ikfom.hpp:14-27): DOF n = 17+6L laid out as
 *   pos 0-2 | rot 3-5 | offset_R_l 6+3l | offset_T_l 6+3L+3l | vel | bg | ba | grav (S2, 2 DOF). */
typedef struct malio_state {
  double pos[3];
  double rot[4];                       /* w,x,y,z */
  malio_rigid ext[MALIO_MAX_LIDAR];    /* offset_R_l, offset_T_l */
  double vel[3], bg[3], ba[3];
  double grav[3];                      /* S2, |grav| = 9.809 */
} malio_state;

typedef struct malio_update_report {
  int32_t passes;               /* passes run (<= max_iter+1) */
  int32_t searches;             /* passes that ran the k-NN */
  int32_t converged_count;      /* t in esekfom.hpp:658 */
  int32_t last_status;
  uint32_t n_eff_last;
  float ms_device_total;        /* sum of device time over the passes */
  float ms_host_solve;          /* host 35x35 algebra */
  double dx_last[MALIO_MAX_DOF];  /* last state increment dx_ (esekfom.hpp:642) */
} malio_update_report;

/* esekf::update_iterated_dyn_share_modified (esekfom.hpp:495-721): iterates malio_measure and the host
 * algebra; x and P (n x n row-major) are updated in place.  R is LASER_POINT_COV (laserMapping.cpp:38),
 * only used by the degenerate branch. */
int malio_ieskf_update(malio_handle* h, malio_state* x, double* P, int max_iter, double R,
                       malio_update_report* report);

/* ---- host utility: static snapshot builder --------------------------------------------------
 * Builds a balanced k-d tree over points (median split on the longest axis, as KD_TREE::BuildTree,
 * ikd_Tree.cpp:696-735) directly in snapshot form, for callers without a live ikd-Tree (benchmarks,
 * the k-NN microbench).  order_out[i] = index into xyz of the point stored in node i. */
int malio_build_static_snapshot(const float* xyz, uint32_t n, malio_map_node* nodes_out,
                                uint32_t* order_out, uint32_t* max_depth_out);

#ifdef __cplusplus
}
#endif
#endif /* MALIO_B200_H_ */
