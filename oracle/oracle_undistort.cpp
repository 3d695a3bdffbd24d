// oracle_undistort.cpp — CPU ORACLE for the rows next to the hot path: N2 per-raw-point B-spline
// undistortion and N3's PCL VoxelGrid.  TEST INFRASTRUCTURE ONLY (same rules as oracle_malio.cpp: only tests/, smoke()
// and bench.py's CPU legs may load it).
//
// Restated function by function from the reference (paths relative to /root/reference/MA_LIO):
//   ov_core::log_so3 / exp_se3 / log_se3 / Inv_se3             include/quat_ops.h:150-243
//   BsplineSE3::find_bounding_poses / _control_points          src/BsplineSE3.cpp:120-231
//   BsplineSE3::get_pose                                        src/BsplineSE3.cpp:84-118
//   the per-point loop of ImuProcess::UndistortPcl              src/IMU_Processing.hpp:452-508
//   Eigen::Quaterniond = Matrix3d (QuaternionBase::operator=)   Eigen 3.3 Geometry/Quaternion.h (unpinned third party,
//                                                               restated from the published algorithm)
//   pcl::VoxelGrid<PointXYZINormal>::applyFilter                 PCL 1.10 filters/impl/voxel_grid.hpp (absent third party;
//                                                               call site src/laserMapping.cpp:968-983)
// PARITY STATUS: unpinned by the reference (no tests / fixtures there; Eigen and PCL absent here).  Cross-checked in
// tests/test_oracle_cpu.py against scipy.linalg expm/logm and an independent numpy voxel filter.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace {

// 3x3 / 4x4 row-major helpers in Eigen's coefficient order (sum over k ascending)
inline void mul3(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
inline void mul4(const double* A, const double* B, double* C) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = A[4 * i] * B[j];
      for (int k = 1; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
      C[4 * i + j] = s;
    }
}

// quat_ops.h:150-187
void log_so3(const double R[9], double omega[3]) {
  const double R11 = R[0], R12 = R[1], R13 = R[2], R21 = R[3], R22 = R[4], R23 = R[5], R31 = R[6], R32 = R[7], R33 = R[8];
  const double tr = R11 + R22 + R33;
  if (tr + 1.0 < 1e-10) {
    if (std::fabs(R33 + 1.0) > 1e-5) {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R33);
      omega[0] = f * R13; omega[1] = f * R23; omega[2] = f * (1.0 + R33);
    } else if (std::fabs(R22 + 1.0) > 1e-5) {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R22);
      omega[0] = f * R12; omega[1] = f * (1.0 + R22); omega[2] = f * R32;
    } else {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R11);
      omega[0] = f * (1.0 + R11); omega[1] = f * R21; omega[2] = f * R31;
    }
  } else {
    double magnitude;
    const double tr_3 = tr - 3.0;
    if (tr_3 < -1e-7) {
      const double theta = std::acos((tr - 1.0) / 2.0);
      magnitude = theta / (2.0 * std::sin(theta));
    } else {
      magnitude = 0.5 - tr_3 / 12.0;
    }
    omega[0] = magnitude * (R32 - R23); omega[1] = magnitude * (R13 - R31); omega[2] = magnitude * (R21 - R12);
  }
}

// quat_ops.h:190-220
void exp_se3(const double vec[6], double mat[16]) {
  const double w[3] = {vec[0], vec[1], vec[2]}, u[3] = {vec[3], vec[4], vec[5]};
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double A, B, Cc;
  if (theta < 1e-7) { A = 1; B = 0.5; Cc = 1.0 / 6.0; }
  else { A = std::sin(theta) / theta; B = (1 - std::cos(theta)) / (theta * theta); Cc = (1 - A) / (theta * theta); }
  double K2[9];
  mul3(K, K, K2);
  double V[9], Rm[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    V[i] = I + B * K[i] + Cc * K2[i];      // C * wskew * wskew parses as (C * wskew) * wskew: same value up to rounding
    Rm[i] = I + A * K[i] + B * K2[i];
  }
  std::memset(mat, 0, 16 * sizeof(double));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) mat[4 * i + j] = Rm[3 * i + j];
    mat[4 * i + 3] = V[3 * i] * u[0] + V[3 * i + 1] * u[1] + V[3 * i + 2] * u[2];
  }
  mat[15] = 1;
}

// quat_ops.h:223-243
void log_se3(const double mat[16], double out[6]) {
  const double R[9] = {mat[0], mat[1], mat[2], mat[4], mat[5], mat[6], mat[8], mat[9], mat[10]};
  double w[3];
  log_so3(R, w);
  const double T[3] = {mat[3], mat[7], mat[11]};
  const double t = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
  if (t < 1e-10) { out[3] = T[0]; out[4] = T[1]; out[5] = T[2]; return; }
  const double a[3] = {w[0] / t, w[1] / t, w[2] / t};
  const double W[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
  const double Tan = std::tan(0.5 * t);
  double WT[3], WWT[3];
  for (int i = 0; i < 3; ++i) WT[i] = W[3 * i] * T[0] + W[3 * i + 1] * T[1] + W[3 * i + 2] * T[2];
  for (int i = 0; i < 3; ++i) WWT[i] = W[3 * i] * WT[0] + W[3 * i + 1] * WT[1] + W[3 * i + 2] * WT[2];
  for (int i = 0; i < 3; ++i) out[3 + i] = T[i] - (0.5 * t) * WT[i] + (1 - t / (2. * Tan)) * WWT[i];
}

// quat_ops.h:252-257
void inv_se3(const double T[16], double Ti[16]) {
  std::memset(Ti, 0, 16 * sizeof(double));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ti[4 * i + j] = T[4 * j + i];
  for (int i = 0; i < 3; ++i) Ti[4 * i + 3] = -(Ti[4 * i] * T[3] + Ti[4 * i + 1] * T[7] + Ti[4 * i + 2] * T[11]);
  Ti[15] = 1;
}

// Eigen 3.3 QuaternionBase::operator=(rotation matrix)  -> (w, x, y, z)
void quat_from_R(const double m[9], double q[4]) {
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[7] - m[5]) * t;
    q[2] = (m[2] - m[6]) * t;
    q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
// Eigen Quaternion * Vector3 (_transformVector), q = (w,x,y,z)
inline void q_rot(const double q[4], const double v[3], double o[3]) {
  const double uv[3] = {2 * (q[2] * v[2] - q[3] * v[1]), 2 * (q[3] * v[0] - q[1] * v[2]), 2 * (q[1] * v[1] - q[2] * v[0])};
  const double c[3] = {q[2] * uv[2] - q[3] * uv[1], q[3] * uv[0] - q[1] * uv[2], q[1] * uv[1] - q[2] * uv[0]};
  for (int i = 0; i < 3; ++i) o[i] = v[i] + q[0] * uv[i] + c[i];
}

// BsplineSE3::find_bounding_control_points (BsplineSE3.cpp:172-231) on the std::map's keys as an ascending array
bool find_bounding(const double* ct, int n, double ts, int idx[4]) {
  const int lb = (int)(std::lower_bound(ct, ct + n, ts) - ct);
  const int ub = (int)(std::upper_bound(ct, ct + n, ts) - ct);
  bool found_older = false, found_newer = false;
  int i1 = -1, i2 = -1;
  if (lb != n) {
    if (ct[lb] == ts) { i1 = lb; found_older = true; }
    else if (lb != 0) { i1 = lb - 1; found_older = true; }
  }
  if (ub != n) { i2 = ub; found_newer = true; }
  if (!(found_older && found_newer)) return false;
  if (i1 == 0) return false;
  const int i0 = i1 - 1, i3 = i2 + 1;
  if (i3 == n) return false;
  idx[0] = i0; idx[1] = i1; idx[2] = i2; idx[3] = i3;
  return true;
}

}  // namespace

extern "C" {

void orc_log_se3(const double T[16], double v[6]) { log_se3(T, v); }
void orc_exp_se3(const double v[6], double T[16]) { exp_se3(v, T); }
void orc_quat_from_R(const double R[9], double q[4]) { quat_from_R(R, q); }

// BsplineSE3::get_pose (BsplineSE3.cpp:84-118).  ctrl_T: n x 16 row-major.  q = (w,x,y,z).  Returns 1 on success.
int orc_bspline_get_pose(const double* ct, const double* cT, int n, double timestamp, double q[4], double p[3]) {
  int id[4];
  if (!find_bounding(ct, n, timestamp, id)) { p[0] = p[1] = p[2] = 0; return 0; }
  const double* pose0 = cT + 16 * (size_t)id[0];
  const double* pose1 = cT + 16 * (size_t)id[1];
  const double* pose2 = cT + 16 * (size_t)id[2];
  const double* pose3 = cT + 16 * (size_t)id[3];
  const double t1 = ct[id[1]], t2 = ct[id[2]];
  const double DT = (t2 - t1);
  const double u = (timestamp - t1) / DT;
  const double b0 = 1.0 / 6.0 * (5 + 3 * u - 3 * u * u + u * u * u);
  const double b1 = 1.0 / 6.0 * (1 + 3 * u + 3 * u * u - 2 * u * u * u);
  const double b2 = 1.0 / 6.0 * (u * u * u);
  auto seg = [](const double* a, const double* b, double bw, double A[16]) {
    double ai[16], rel[16], l[6];
    inv_se3(a, ai);
    mul4(ai, b, rel);
    log_se3(rel, l);
    for (int k = 0; k < 6; ++k) l[k] = bw * l[k];
    exp_se3(l, A);
  };
  double A0[16], A1[16], A2[16], m1[16], m2[16], pose[16];
  seg(pose0, pose1, b0, A0);
  seg(pose1, pose2, b1, A1);
  seg(pose2, pose3, b2, A2);
  mul4(pose0, A0, m1);
  mul4(m1, A1, m2);
  mul4(m2, A2, pose);
  const double R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
  quat_from_R(R, q);
  p[0] = pose[3]; p[1] = pose[7]; p[2] = pose[11];
  return 1;
}

// The per-point loop of UndistortPcl for ONE LiDAR (IMU_Processing.hpp:468-508), points given in buffer order
// (ascending time); the loop walks them from the last to the second (`it_pcl != begin()`: the first point is never touched).
//   pts        n x {x, y, z, curvature[ms]}   (float)
//   ext_q/t    extrinsic_quat[num] (w,x,y,z), extrinsic_trans[num];  lt_q/t  lt_imu_frame_quat[num], lt_imu_frame_trans[num]
//   cov_t      imu_cov[k].first.first, k < n_cov;  cov_pointer = its value when the loop starts (:455-466)
// outputs: xyz_out (n x 3 float: compensated where spline_flag != 0, else the input), idx_out (n: the value written to
// `intensity`, INT32_MIN where untouched), ok_out (spline_flag), pop_point[k] = index of the point at which the k-th
// table entry was pushed (:487-494), *n_pops; pose_out (optional, n x 7: q (w,x,y,z), p of get_pose).
void orc_undistort(const float* pts, int64_t n, double beg_time, const double ext_q[4], const double ext_t[3],
                   const double lt_q[4], const double lt_t[3], const double* ct, const double* cT, int n_ctrl,
                   const double* cov_t, int n_cov, int cov_pointer, float* xyz_out, int32_t* idx_out, uint8_t* ok_out,
                   int32_t* pop_point, int32_t* n_pops, double* pose_out) {
  int idx = -1, pops = 0;
  for (int64_t i = 0; i < n; ++i) {
    xyz_out[3 * i] = pts[4 * i]; xyz_out[3 * i + 1] = pts[4 * i + 1]; xyz_out[3 * i + 2] = pts[4 * i + 2];
    idx_out[i] = std::numeric_limits<int32_t>::min();
    ok_out[i] = 0;
    if (pose_out) for (int k = 0; k < 7; ++k) pose_out[7 * i + k] = 0.0;
  }
  const double ext_qc[4] = {ext_q[0], -ext_q[1], -ext_q[2], -ext_q[3]};
  const double lt_qc[4] = {lt_q[0], -lt_q[1], -lt_q[2], -lt_q[3]};
  for (int64_t i = n - 1; i > 0; --i) {
    const double point_t = (double)pts[4 * i + 3] / double(1000) + beg_time;                    // :474
    double q[4] = {1, 0, 0, 0}, p[3];
    const int flag = orc_bspline_get_pose(ct, cT, n_ctrl, point_t, q, p);                        // :475
    if (cov_pointer >= 0 && cov_pointer < n_cov && cov_t[cov_pointer] > point_t) {              // :476
      cov_pointer = cov_pointer - 1;
      if (pop_point && pops < n_cov) pop_point[pops] = (int32_t)i;
      pops++;
      idx += 1;                                                                                  // :485
    }
    if (pose_out && flag) { for (int k = 0; k < 4; ++k) pose_out[7 * i + k] = q[k]; for (int k = 0; k < 3; ++k) pose_out[7 * i + 4 + k] = p[k]; }
    if (flag != 0) {                                                                             // :488-496
      const double P_i[3] = {pts[4 * i], pts[4 * i + 1], pts[4 * i + 2]};
      const double T_ei[3] = {p[0] - lt_t[0], p[1] - lt_t[1], p[2] - lt_t[2]};
      double a[3], b[3], c[3], d[3];
      q_rot(ext_q, P_i, a);
      for (int k = 0; k < 3; ++k) a[k] += ext_t[k];
      q_rot(q, a, b);
      for (int k = 0; k < 3; ++k) b[k] += T_ei[k];
      q_rot(lt_qc, b, c);
      for (int k = 0; k < 3; ++k) c[k] -= ext_t[k];
      q_rot(ext_qc, c, d);
      xyz_out[3 * i] = (float)d[0]; xyz_out[3 * i + 1] = (float)d[1]; xyz_out[3 * i + 2] = (float)d[2];
      idx_out[i] = idx;
      ok_out[i] = 1;
    }
  }
  if (n_pops) *n_pops = pops;
}

// ---------------------------------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZINormal>::applyFilter with setLeafSize(l, l, l), default settings otherwise
// (downsample_all_data = true, min_points_per_voxel = 0, no field filter, keep_organized = false), call site
// laserMapping.cpp:968-983.  PCL 1.10 voxel_grid.hpp:
//   bounds  = getMinMax3D over finite points                         (:62-64, common.hpp)
//   min_b[k] = floor(min[k] * inverse_leaf_size[k]), max_b likewise; div_b = max_b - min_b + 1; divb_mul = (1, dx, dx*dy)
//   idx(p)  = (floor(x*il) - min_b0) * 1 + (floor(y*il) - min_b1) * dx + (floor(z*il) - min_b2) * dx*dy
//   std::sort of (idx, cloud index) pairs by idx (operator< on idx only: NOT stable), one output point per distinct
//   idx in ascending idx order = centroid of ALL fields (CentroidPoint: float accumulation in sorted order, divided by the
//   count).  Because std::sort is unstable the accumulation ORDER inside a voxel is unspecified in PCL itself; this
//   restatement accumulates in ascending cloud index (what a stable sort would give).  Fields here: x, y, z, intensity,
//   normal_x, normal_y, normal_z, curvature (8 floats), averaged like PCL's centroid of PointXYZINormal does.
// pts_in: n x 8 float {x,y,z,intensity,normal_x,normal_y,normal_z,curvature}; out: same layout, returns the count;
// voxel_of (optional, n): output slot of each input point (-1 for non-finite points).
int64_t orc_voxel_grid(const float* pts_in, int64_t n, float leaf, float* out, int64_t cap, int32_t* voxel_of) {
  const float il = 1.0f / leaf;   // inverse_leaf_size_ = Eigen::Array4f::Ones() / leaf_size_.array()
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  for (int64_t i = 0; i < n; ++i) {
    const float* p = pts_in + 8 * i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
  }
  int minb[3], maxb[3];
  for (int k = 0; k < 3; ++k) { minb[k] = (int)std::floor(mn[k] * il); maxb[k] = (int)std::floor(mx[k] * il); }
  const int64_t dx = (int64_t)maxb[0] - minb[0] + 1, dy = (int64_t)maxb[1] - minb[1] + 1;
  std::vector<std::pair<int64_t, int64_t>> order;
  order.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    const float* p = pts_in + 8 * i;
    if (voxel_of) voxel_of[i] = -1;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    const int ijk0 = (int)(std::floor(p[0] * il) - (float)minb[0]);
    const int ijk1 = (int)(std::floor(p[1] * il) - (float)minb[1]);
    const int ijk2 = (int)(std::floor(p[2] * il) - (float)minb[2]);
    order.emplace_back((int64_t)ijk0 + (int64_t)ijk1 * dx + (int64_t)ijk2 * dx * dy, i);
  }
  std::stable_sort(order.begin(), order.end(), [](const std::pair<int64_t, int64_t>& a, const std::pair<int64_t, int64_t>& b) { return a.first < b.first; });
  int64_t m = 0;
  size_t k = 0;
  while (k < order.size()) {
    size_t e = k;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (e < order.size() && order[e].first == order[k].first) {
      const float* p = pts_in + 8 * order[e].second;
      for (int f = 0; f < 8; ++f) acc[f] += p[f];
      if (voxel_of) voxel_of[order[e].second] = (int32_t)m;
      ++e;
    }
    if (m < cap) for (int f = 0; f < 8; ++f) out[8 * m + f] = acc[f] / (float)(e - k);
    ++m;
    k = e;
  }
  return m;
}

}  // extern "C"
