// malio_preproc.cu — the two stages immediately before the hot path, sm_100a:
//   N2  ImuProcess::UndistortPcl's per-raw-point loop        src/IMU_Processing.hpp:468-508
//       BsplineSE3::get_pose                                  src/BsplineSE3.cpp:84-118, quat_ops.h:150-257
//   N3  pcl::VoxelGrid down-sampling + merge into the scan    src/laserMapping.cpp:968-983
// Compiled with -fmad=false like the rest of the library: every double expression is evaluated with the IEEE operations
// the reference's x86-64 build uses, in its order.
//
// N2 on the device.  (1) log_se3(Inv(P_k) P_{k+1}) depends on the control points only: one thread per control-point pair
// computes it once per call (the reference recomputes three of them per point).  (2) One thread per raw point: bounding
// control points by binary search in shared memory, three exp_se3, the pose product, Eigen's matrix->quaternion
// conversion, the compensation of :492, and need_i = how many entries of the IMU-covariance list lie above the point's
// time.  (3) The reference walks that list with AT MOST ONE pop per point (:476-486), last point first:
//        pops_s = min(need_s, pops_{s-1} + 1)   =>   pops_s = s + min(1, min_{j<=s}(need_j - j))
// a min-plus prefix scan over the reversed point order, done by one block (chunked Hillis-Steele in shared memory);
// intensity = pops - 1, and every increase of pops marks the point whose pose seeds a table entry.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "malio_device.cuh"

using namespace malio_devstate;

namespace {

struct PreLidar {
  uint32_t n = 0, cap = 0;            // raw / undistorted points
  malio_raw_pt* d_raw = nullptr;
  float* d_ud = nullptr;              // n x 5: x, y, z, intensity (= idx as float, 0 where untouched), curvature
  int32_t* d_need = nullptr;
  int32_t* d_idx = nullptr;
  uint8_t* d_ok = nullptr;
  double* d_pose = nullptr;           // n x 7, only when asked for
  // voxel grid
  uint32_t n_ds = 0, cap_ds = 0;
  float* d_ds = nullptr;              // n_ds x 5
  bool ds_valid = false;
};
struct PreState {
  cudaStream_t stream = nullptr;
  PreLidar lid[MALIO_MAX_LIDAR];
  double *d_ct = nullptr, *d_cT = nullptr, *d_seglog = nullptr, *d_covt = nullptr;
  int32_t* d_pop = nullptr;           // MALIO_MAX_COV + 1 (last = count)
  // voxel scratch
  float* d_vin = nullptr; uint32_t cap_vin = 0;       // host-provided input staged on the device
  uint32_t *d_vidx = nullptr, *d_vtmp = nullptr, *d_vorder = nullptr; uint32_t cap_vpts = 0;
  uint32_t *d_vcnt = nullptr, *d_voff = nullptr, *d_vocc = nullptr, *d_vcur = nullptr; uint64_t cap_cells = 0;
  uint32_t *d_ctot = nullptr, *d_cbase = nullptr;     // chunk totals / bases of the two cell scans (2 x 65536)
  uint32_t* d_bounds = nullptr;       // 6 order-preserving keys + out count
  uint32_t* h_small = nullptr;        // pinned: 8 words
};

template <class T>
int grow(malio_handle* h, T*& p, size_t count) {
  if (p) { cudaFree(p); p = nullptr; }
  CUDA_TRY(cudaMalloc((void**)&p, count * sizeof(T)));
  return MALIO_OK;
}

int get_state(malio_handle* h, PreState*& S) {
  S = (PreState*)h->pre;
  if (S) return MALIO_OK;
  DeviceState* D = (DeviceState*)h->dev;
  CUDA_TRY(cudaSetDevice(D->device));
  S = new PreState;
  h->pre = S;
  S->stream = D->stream;   // same stream as the scan path: malio_upload_scan_device is ordered after the voxel grid
  CUDA_TRY(cudaMalloc((void**)&S->d_ct, MALIO_MAX_CTRL * sizeof(double)));
  CUDA_TRY(cudaMalloc((void**)&S->d_cT, (size_t)MALIO_MAX_CTRL * 16 * sizeof(double)));
  CUDA_TRY(cudaMalloc((void**)&S->d_seglog, (size_t)MALIO_MAX_CTRL * 6 * sizeof(double)));
  CUDA_TRY(cudaMalloc((void**)&S->d_covt, MALIO_MAX_COV * sizeof(double)));
  CUDA_TRY(cudaMalloc((void**)&S->d_pop, (MALIO_MAX_COV + 1) * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc((void**)&S->d_bounds, 8 * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc((void**)&S->d_ctot, 2 * 65536 * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc((void**)&S->d_cbase, 2 * 65536 * sizeof(uint32_t)));
  CUDA_TRY(cudaHostAlloc((void**)&S->h_small, 8 * sizeof(uint32_t), cudaHostAllocDefault));
  return MALIO_OK;
}

// ------------------------------------------------------------------ quat_ops.h restated for the device
__host__ __device__ inline void mul3(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// 4x4 homogeneous product in Eigen's coefficient order (k ascending, the bottom row 0 0 0 1 included)
__host__ __device__ inline void mul4(const double* A, const double* B, double* C) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = A[4 * i] * B[j];
      for (int k = 1; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
      C[4 * i + j] = s;
    }
}
// quat_ops.h:150-187
__host__ __device__ inline void log_so3(const double R[9], double omega[3]) {
  const double R11 = R[0], R12 = R[1], R13 = R[2], R21 = R[3], R22 = R[4], R23 = R[5], R31 = R[6], R32 = R[7], R33 = R[8];
  const double tr = R11 + R22 + R33;
  const double PI = 3.14159265358979323846;
  if (tr + 1.0 < 1e-10) {
    if (fabs(R33 + 1.0) > 1e-5) {
      const double f = PI / sqrt(2.0 + 2.0 * R33);
      omega[0] = f * R13; omega[1] = f * R23; omega[2] = f * (1.0 + R33);
    } else if (fabs(R22 + 1.0) > 1e-5) {
      const double f = PI / sqrt(2.0 + 2.0 * R22);
      omega[0] = f * R12; omega[1] = f * (1.0 + R22); omega[2] = f * R32;
    } else {
      const double f = PI / sqrt(2.0 + 2.0 * R11);
      omega[0] = f * (1.0 + R11); omega[1] = f * R21; omega[2] = f * R31;
    }
  } else {
    double magnitude;
    const double tr_3 = tr - 3.0;
    if (tr_3 < -1e-7) {
      const double theta = acos((tr - 1.0) / 2.0);
      magnitude = theta / (2.0 * sin(theta));
    } else {
      magnitude = 0.5 - tr_3 / 12.0;
    }
    omega[0] = magnitude * (R32 - R23); omega[1] = magnitude * (R13 - R31); omega[2] = magnitude * (R21 - R12);
  }
}
// quat_ops.h:190-220
__host__ __device__ inline void exp_se3(const double vec[6], double mat[16]) {
  const double w0 = vec[0], w1 = vec[1], w2 = vec[2];
  const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
  const double K[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
  double A, B, Cc;
  if (theta < 1e-7) { A = 1; B = 0.5; Cc = 1.0 / 6.0; }
  else { A = sin(theta) / theta; B = (1 - cos(theta)) / (theta * theta); Cc = (1 - A) / (theta * theta); }
  double K2[9];
  mul3(K, K, K2);
  double V[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    V[i] = I + B * K[i] + Cc * K2[i];
    const double r = I + A * K[i] + B * K2[i];
    mat[4 * (i / 3) + (i % 3)] = r;
  }
  for (int i = 0; i < 3; ++i) mat[4 * i + 3] = V[3 * i] * vec[3] + V[3 * i + 1] * vec[4] + V[3 * i + 2] * vec[5];
  mat[12] = 0; mat[13] = 0; mat[14] = 0; mat[15] = 1;
}
// quat_ops.h:223-243
__host__ __device__ inline void log_se3(const double mat[16], double out[6]) {
  const double R[9] = {mat[0], mat[1], mat[2], mat[4], mat[5], mat[6], mat[8], mat[9], mat[10]};
  double w[3];
  log_so3(R, w);
  const double T[3] = {mat[3], mat[7], mat[11]};
  const double t = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
  if (t < 1e-10) { out[3] = T[0]; out[4] = T[1]; out[5] = T[2]; return; }
  const double a[3] = {w[0] / t, w[1] / t, w[2] / t};
  const double W[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
  const double Tan = tan(0.5 * t);
  double WT[3], WWT[3];
  for (int i = 0; i < 3; ++i) WT[i] = W[3 * i] * T[0] + W[3 * i + 1] * T[1] + W[3 * i + 2] * T[2];
  for (int i = 0; i < 3; ++i) WWT[i] = W[3 * i] * WT[0] + W[3 * i + 1] * WT[1] + W[3 * i + 2] * WT[2];
  for (int i = 0; i < 3; ++i) out[3 + i] = T[i] - (0.5 * t) * WT[i] + (1 - t / (2. * Tan)) * WWT[i];
}
// quat_ops.h:252-257
__host__ __device__ inline void inv_se3(const double T[16], double Ti[16]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ti[4 * i + j] = T[4 * j + i];
  for (int i = 0; i < 3; ++i) Ti[4 * i + 3] = -(Ti[4 * i] * T[3] + Ti[4 * i + 1] * T[7] + Ti[4 * i + 2] * T[11]);
  Ti[12] = 0; Ti[13] = 0; Ti[14] = 0; Ti[15] = 1;
}
// Eigen 3.3 QuaternionBase::operator=(rotation matrix) -> (w, x, y, z); m row-major 3x3
__host__ __device__ inline void quat_from_R(const double m[9], double q[4]) {
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
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
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double qv[3];
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
    qv[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    qv[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q[1] = qv[0]; q[2] = qv[1]; q[3] = qv[2];
  }
}
// Eigen Quaternion * Vector3 (_transformVector), q = (w,x,y,z)
__host__ __device__ inline void q_rot(const double q[4], const double v[3], double o[3]) {
  const double uv0 = 2 * (q[2] * v[2] - q[3] * v[1]), uv1 = 2 * (q[3] * v[0] - q[1] * v[2]), uv2 = 2 * (q[1] * v[1] - q[2] * v[0]);
  const double c0 = q[2] * uv2 - q[3] * uv1, c1 = q[3] * uv0 - q[1] * uv2, c2 = q[1] * uv1 - q[2] * uv0;
  o[0] = v[0] + q[0] * uv0 + c0;
  o[1] = v[1] + q[0] * uv1 + c1;
  o[2] = v[2] + q[0] * uv2 + c2;
}
// BsplineSE3::find_bounding_control_points (BsplineSE3.cpp:120-231) on the ascending key array; i0..i3 on success
__host__ __device__ inline bool find_bounding(const double* ct, int n, double ts, int& i0, int& i1, int& i2, int& i3) {
  int lo = 0, hi = n;                 // lower_bound: first key >= ts
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (ct[mid] < ts) lo = mid + 1; else hi = mid; }
  const int lb = lo;
  lo = lb; hi = n;                    // upper_bound: first key > ts
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (ct[mid] <= ts) lo = mid + 1; else hi = mid; }
  const int ub = lo;
  bool older = false;
  i1 = -1;
  if (lb != n) {
    if (ct[lb] == ts) { i1 = lb; older = true; }
    else if (lb != 0) { i1 = lb - 1; older = true; }
  }
  if (!older || ub == n) return false;
  i2 = ub;
  if (i1 == 0) return false;
  i0 = i1 - 1; i3 = i2 + 1;
  return i3 != n;
}
// the spline pose from the per-pair logarithms (seglog[k] = log_se3(Inv(P_k) P_{k+1}))
__host__ __device__ inline bool spline_pose(const double* ct, const double* cT, const double* seglog, int n, double ts, double q[4], double p[3]) {
  int i0, i1, i2, i3;
  if (!find_bounding(ct, n, ts, i0, i1, i2, i3)) return false;
  const double t1 = ct[i1], t2 = ct[i2];
  const double DT = (t2 - t1);
  const double u = (ts - t1) / DT;
  const double b0 = 1.0 / 6.0 * (5 + 3 * u - 3 * u * u + u * u * u);
  const double b1 = 1.0 / 6.0 * (1 + 3 * u + 3 * u * u - 2 * u * u * u);
  const double b2 = 1.0 / 6.0 * (u * u * u);
  double l[6], A[16], m1[16], m2[16];
  for (int k = 0; k < 6; ++k) l[k] = b0 * seglog[6 * i0 + k];
  exp_se3(l, A);
  mul4(cT + 16 * i0, A, m1);
  for (int k = 0; k < 6; ++k) l[k] = b1 * seglog[6 * i1 + k];
  exp_se3(l, A);
  mul4(m1, A, m2);
  for (int k = 0; k < 6; ++k) l[k] = b2 * seglog[6 * i2 + k];
  exp_se3(l, A);
  mul4(m2, A, m1);
  const double R[9] = {m1[0], m1[1], m1[2], m1[4], m1[5], m1[6], m1[8], m1[9], m1[10]};
  quat_from_R(R, q);
  p[0] = m1[3]; p[1] = m1[7]; p[2] = m1[11];
  return true;
}

__global__ void seglog_kernel(const double* __restrict__ cT, int n, double* __restrict__ seglog) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k + 1 >= n) return;
  double ai[16], rel[16], l[6];
  inv_se3(cT + 16 * k, ai);
  mul4(ai, cT + 16 * (k + 1), rel);
  log_se3(rel, l);
  for (int j = 0; j < 6; ++j) seglog[6 * k + j] = l[j];
}

struct UdConst {
  double beg_time;
  double eq[4], et[3], lq[4], lt[3];
  int n_ctrl, n_cov, cov_pointer;
};
constexpr int UD_THREADS = 128;
__global__ void __launch_bounds__(UD_THREADS)
undistort_kernel(const malio_raw_pt* __restrict__ raw, uint32_t n, UdConst c, const double* __restrict__ ct,
                 const double* __restrict__ cT, const double* __restrict__ seglog, const double* __restrict__ covt,
                 float* __restrict__ ud, int32_t* __restrict__ need, uint8_t* __restrict__ ok, double* __restrict__ pose) {
  extern __shared__ double s_mem[];
  double* s_ct = s_mem;                       // n_ctrl
  double* s_cov = s_ct + c.n_ctrl;            // n_cov
  for (int k = threadIdx.x; k < c.n_ctrl; k += UD_THREADS) s_ct[k] = ct[k];
  for (int k = threadIdx.x; k < c.n_cov; k += UD_THREADS) s_cov[k] = covt[k];
  __syncthreads();
  const uint32_t i = blockIdx.x * UD_THREADS + threadIdx.x;
  if (i >= n) return;
  const malio_raw_pt pt = raw[i];
  float ox = pt.x, oy = pt.y, oz = pt.z;
  uint8_t flag = 0;
  int32_t nd = 0;
  if (i > 0) {                                // the reference's loop stops before begin(): point 0 is never touched
    const double point_t = (double)pt.curvature / double(1000) + c.beg_time;                   // :474
    double q[4], p[3];
    if (spline_pose(s_ct, cT, seglog, c.n_ctrl, point_t, q, p)) {                               // :475
      flag = 1;
      const double P_i[3] = {(double)pt.x, (double)pt.y, (double)pt.z};
      const double T_ei[3] = {p[0] - c.lt[0], p[1] - c.lt[1], p[2] - c.lt[2]};
      const double eqc[4] = {c.eq[0], -c.eq[1], -c.eq[2], -c.eq[3]}, lqc[4] = {c.lq[0], -c.lq[1], -c.lq[2], -c.lq[3]};
      double a[3], b[3], cc[3], d[3];
      q_rot(c.eq, P_i, a);                                                                      // :492
      a[0] += c.et[0]; a[1] += c.et[1]; a[2] += c.et[2];
      q_rot(q, a, b);
      b[0] += T_ei[0]; b[1] += T_ei[1]; b[2] += T_ei[2];
      q_rot(lqc, b, cc);
      cc[0] -= c.et[0]; cc[1] -= c.et[1]; cc[2] -= c.et[2];
      q_rot(eqc, cc, d);
      ox = (float)d[0]; oy = (float)d[1]; oz = (float)d[2];
      if (pose) { double* o = pose + 7 * (size_t)i; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; o[4] = p[0]; o[5] = p[1]; o[6] = p[2]; }
    } else if (pose) {
      double* o = pose + 7 * (size_t)i;
      for (int k = 0; k < 7; ++k) o[k] = 0.0;
    }
    // entries k <= cov_pointer of the (ascending) list with time > point_t: what the walk may still pop at this point
    int lo = 0, hi = c.cov_pointer + 1;
    if (hi > c.n_cov) hi = c.n_cov;
    const int top = hi;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cov[mid] <= point_t) lo = mid + 1; else hi = mid; }
    nd = top - lo;
  } else if (pose) {
    double* o = pose;
    for (int k = 0; k < 7; ++k) o[k] = 0.0;
  }
  float* o5 = ud + 5 * (size_t)i;
  o5[0] = ox; o5[1] = oy; o5[2] = oz; o5[4] = pt.curvature;   // [3] = intensity, written by the scan kernel
  need[i] = nd;
  ok[i] = flag;
}

// reverse-order min-plus scan (see the file header): one block, chunks of SCAN_T points, carry = running prefix minimum
constexpr int SCAN_T = 1024;
__global__ void __launch_bounds__(SCAN_T)
idx_scan_kernel(const int32_t* __restrict__ need, const uint8_t* __restrict__ ok, uint32_t n, float* __restrict__ ud,
                int32_t* __restrict__ idx_out, int32_t* __restrict__ pop_point, int n_cov) {
  __shared__ int32_t s_v[SCAN_T];
  __shared__ int32_t s_carry;
  if (threadIdx.x == 0) s_carry = 1;          // min(1, ...): pops_s <= s + 1
  for (int k = threadIdx.x; k <= n_cov; k += SCAN_T) pop_point[k] = (k == n_cov) ? 0 : -1;
  __syncthreads();
  const uint32_t steps = n > 0 ? n - 1 : 0;   // s = 0 .. n-2  <->  i = n-1 .. 1
  for (uint32_t base = 0; base < steps; base += SCAN_T) {
    const uint32_t s = base + threadIdx.x;
    const bool valid = s < steps;
    const uint32_t i = valid ? (n - 1 - s) : 0;
    int32_t v = valid ? (need[i] - (int32_t)s) : INT32_MAX;
    s_v[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_T; o <<= 1) {    // inclusive prefix minimum
      const int32_t t = (threadIdx.x >= (unsigned)o) ? s_v[threadIdx.x - o] : INT32_MAX;
      __syncthreads();
      v = min(v, t);
      s_v[threadIdx.x] = v;
      __syncthreads();
    }
    const int32_t carry = s_carry;
    const int32_t m = min(carry, v);
    const int32_t m_prev = (threadIdx.x == 0) ? carry : min(carry, s_v[threadIdx.x - 1]);
    if (valid) {
      const int32_t pops = (int32_t)s + m;                                  // after processing this point
      const int32_t pops_prev = (s == 0) ? 0 : (int32_t)(s - 1) + m_prev;   // before it
      if (pops > pops_prev && pops - 1 < n_cov) pop_point[pops - 1] = (int32_t)i;
      const bool touched = ok[i] != 0;
      idx_out[i] = touched ? pops - 1 : INT32_MIN;                          // :496 only where spline_flag != 0
      ud[5 * (size_t)i + 3] = touched ? (float)(pops - 1) : 0.f;
    }
    __syncthreads();
    if (threadIdx.x == SCAN_T - 1) s_carry = m;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (n > 0) { idx_out[0] = INT32_MIN; ud[3] = 0.f; }
    const int32_t total = steps ? (int32_t)(steps - 1) + s_carry : 0;
    pop_point[n_cov] = total < 0 ? 0 : total;
  }
}

// ------------------------------------------------------------------ N3: voxel grid
__device__ __forceinline__ uint32_t fkey(float f) {   // order-preserving float -> uint32
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float fkey_inv(uint32_t k) {
  const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  float f;
  memcpy(&f, &b, 4);
  return f;
}
__global__ void __launch_bounds__(256) vg_bounds_kernel(const float* __restrict__ in, uint32_t n, uint32_t* __restrict__ keys) {
  // getMinMax3D over the finite points: grid-stride, one set of atomics per block
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
    const float x = in[5 * (size_t)i], y = in[5 * (size_t)i + 1], z = in[5 * (size_t)i + 2];
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x); lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y); lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
  __shared__ float s_lo[8][3], s_hi[8][3];
  if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; ++k) { s_lo[threadIdx.x >> 5][k] = lo[k]; s_hi[threadIdx.x >> 5][k] = hi[k]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], s_lo[w][k]); hi[k] = fmaxf(hi[k], s_hi[w][k]); }
    if (lo[0] <= hi[0]) for (int k = 0; k < 3; ++k) { atomicMin(keys + k, fkey(lo[k])); atomicMax(keys + 3 + k, fkey(hi[k])); }
  }
}
struct VgConst { float il; int minb[3]; long long dx, dxy; };
__global__ void vg_count_kernel(const float* __restrict__ in, uint32_t n, VgConst c, uint32_t* __restrict__ vidx,
                                uint32_t* __restrict__ cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[5 * (size_t)i], y = in[5 * (size_t)i + 1], z = in[5 * (size_t)i + 2];
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) { vidx[i] = 0xFFFFFFFFu; return; }
  // static_cast<int>(std::floor(p.x * inverse_leaf_size_[0]) - static_cast<float>(min_b_[0]))   (voxel_grid.hpp)
  const int a = (int)(floorf(x * c.il) - (float)c.minb[0]);
  const int b = (int)(floorf(y * c.il) - (float)c.minb[1]);
  const int d = (int)(floorf(z * c.il) - (float)c.minb[2]);
  const uint32_t v = (uint32_t)((long long)a + (long long)b * c.dx + (long long)d * c.dxy);
  vidx[i] = v;
  atomicAdd(cnt + v, 1u);
}
// exclusive scan of count[] (-> off) and of (count != 0) (-> occ) over `cells` entries: chunk-local pass, chunk totals,
// add-back (two quantities side by side)
constexpr int VS_T = 1024, VS_PER = 4, VS_CHUNK = VS_T * VS_PER;
__global__ void __launch_bounds__(VS_T) vg_scan_local_kernel(const uint32_t* __restrict__ cnt, uint64_t cells, uint32_t* __restrict__ off,
                                                             uint32_t* __restrict__ occ, uint32_t* __restrict__ ctot) {
  __shared__ uint32_t s_a[32], s_b[32];
  const uint64_t base = ((uint64_t)blockIdx.x * VS_T + threadIdx.x) * VS_PER;
  uint32_t v[VS_PER], sa = 0, sb = 0;
#pragma unroll
  for (int k = 0; k < VS_PER; ++k) { v[k] = (base + k < cells) ? cnt[base + k] : 0u; sa += v[k]; sb += v[k] != 0; }
  uint32_t ia = sa, ib = sb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
    if ((threadIdx.x & 31) >= (unsigned)o) { ia += ta; ib += tb; }
  }
  if ((threadIdx.x & 31) == 31) { s_a[threadIdx.x >> 5] = ia; s_b[threadIdx.x >> 5] = ib; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t wa = s_a[threadIdx.x], wb = s_b[threadIdx.x];
    uint32_t xa = wa, xb = wb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
      if (threadIdx.x >= (unsigned)o) { xa += ta; xb += tb; }
    }
    s_a[threadIdx.x] = xa - wa; s_b[threadIdx.x] = xb - wb;
    if (threadIdx.x == 31) { ctot[2 * blockIdx.x] = xa; ctot[2 * blockIdx.x + 1] = xb; }
  }
  __syncthreads();
  uint32_t ra = s_a[threadIdx.x >> 5] + ia - sa, rb = s_b[threadIdx.x >> 5] + ib - sb;
#pragma unroll
  for (int k = 0; k < VS_PER; ++k) {
    if (base + k < cells) { off[base + k] = ra; occ[base + k] = rb; }
    ra += v[k]; rb += v[k] != 0;
  }
}
__global__ void __launch_bounds__(1024) vg_scan_tot_kernel(const uint32_t* __restrict__ ctot, uint32_t nchunk, uint32_t* __restrict__ cbase,
                                                           uint32_t* __restrict__ totals) {
  // nchunk <= 65536: each of 1024 threads owns 64 consecutive chunks
  __shared__ uint32_t s_a[32], s_b[32];
  uint32_t sa = 0, sb = 0;
  const uint32_t b0 = threadIdx.x * 64;
  for (uint32_t k = 0; k < 64; ++k) { const uint32_t c = b0 + k; if (c < nchunk) { sa += ctot[2 * c]; sb += ctot[2 * c + 1]; } }
  uint32_t ia = sa, ib = sb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
    if ((threadIdx.x & 31) >= (unsigned)o) { ia += ta; ib += tb; }
  }
  if ((threadIdx.x & 31) == 31) { s_a[threadIdx.x >> 5] = ia; s_b[threadIdx.x >> 5] = ib; }
  __syncthreads();
  uint32_t wa = 0, wb = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) { wa += s_a[w]; wb += s_b[w]; }
  uint32_t ra = wa + ia - sa, rb = wb + ib - sb;
  for (uint32_t k = 0; k < 64; ++k) {
    const uint32_t c = b0 + k;
    if (c < nchunk) { cbase[2 * c] = ra; cbase[2 * c + 1] = rb; ra += ctot[2 * c]; rb += ctot[2 * c + 1]; }
  }
  if (threadIdx.x == 1023) { totals[0] = ra; totals[1] = rb; }   // points binned, occupied voxels
}
__global__ void __launch_bounds__(VS_T) vg_scan_add_kernel(uint32_t* __restrict__ off, uint32_t* __restrict__ occ, uint64_t cells,
                                                           const uint32_t* __restrict__ cbase) {
  const uint64_t base = ((uint64_t)blockIdx.x * VS_T + threadIdx.x) * VS_PER;
  const uint32_t a = cbase[2 * blockIdx.x], b = cbase[2 * blockIdx.x + 1];
#pragma unroll
  for (int k = 0; k < VS_PER; ++k) if (base + k < cells) { off[base + k] += a; occ[base + k] += b; }
}
__global__ void vg_scatter_kernel(const uint32_t* __restrict__ vidx, uint32_t n, const uint32_t* __restrict__ off,
                                  uint32_t* __restrict__ cursor, uint32_t* __restrict__ tmp) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = vidx[i];
  if (v == 0xFFFFFFFFu) return;
  tmp[off[v] + atomicAdd(cursor + v, 1u)] = i;
}
// each point finds its rank among its voxel-mates by input index (deterministic order inside a voxel) ...
__global__ void vg_rank_kernel(const uint32_t* __restrict__ vidx, uint32_t n, const uint32_t* __restrict__ off,
                               const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ tmp, uint32_t* __restrict__ order) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = vidx[i];
  if (v == 0xFFFFFFFFu) return;
  const uint32_t o = off[v], c = cnt[v];
  uint32_t r = 0;
  for (uint32_t j = 0; j < c; ++j) r += (tmp[o + j] < i) ? 1u : 0u;
  order[o + r] = i;
}
// ... and the first point of every voxel sums the voxel's run sequentially (float, ascending input index: the order a
// stable sort would give PCL's centroid accumulator) and writes the centroid at the voxel's rank among the occupied ones
__global__ void vg_centroid_kernel(const float* __restrict__ in, const uint32_t* __restrict__ vidx, uint32_t n,
                                   const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                                   const uint32_t* __restrict__ occ, const uint32_t* __restrict__ order,
                                   float* __restrict__ out, uint32_t out_cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = vidx[i];
  if (v == 0xFFFFFFFFu) return;
  const uint32_t o = off[v];
  if (order[o] != i) return;
  const uint32_t c = cnt[v], slot = occ[v];
  if (slot >= out_cap) return;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (uint32_t j = 0; j < c; ++j) {
    const float* p = in + 5 * (size_t)order[o + j];
#pragma unroll
    for (int f = 0; f < 5; ++f) acc[f] += p[f];
  }
  const float cf = (float)c;
#pragma unroll
  for (int f = 0; f < 5; ++f) out[5 * (size_t)slot + f] = acc[f] / cf;
}
__global__ void merge_scan_kernel(const float* __restrict__ ds, uint32_t n, uint16_t lidar, malio_scan_pt* __restrict__ dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = ds + 5 * (size_t)i;
  malio_scan_pt o;
  o.x = p[0]; o.y = p[1]; o.z = p[2];
  o.lidar = lidar;                                   // laserMapping.cpp:976
  const int t = (int)p[3];                           // normal_x = averaged intensity (:975), int() at :694/:737
  o.table_idx = (uint16_t)(t < 0 ? 0 : (t > 65535 ? 65535 : t));
  dst[i] = o;
}

}  // namespace

namespace malio_pre {

void destroy(malio_handle* h) {
  PreState* S = (PreState*)h->pre;
  if (!S) return;
  for (auto& L : S->lid) {
    void* p[] = {L.d_raw, L.d_ud, L.d_need, L.d_idx, L.d_ok, L.d_pose, L.d_ds};
    for (void* q : p) if (q) cudaFree(q);
  }
  void* p[] = {S->d_ct, S->d_cT, S->d_seglog, S->d_covt, S->d_pop, S->d_vin, S->d_vidx, S->d_vtmp, S->d_vorder, S->d_vcnt, S->d_voff,
               S->d_vocc, S->d_vcur, S->d_ctot, S->d_cbase, S->d_bounds};
  for (void* q : p) if (q) cudaFree(q);
  if (S->h_small) cudaFreeHost(S->h_small);
  delete S;
  h->pre = nullptr;
}

int undistort(malio_handle* h, int lidar, const malio_raw_pt* pts, uint32_t n, const malio_undistort_args* a, float* xyz,
              int32_t* idx, uint8_t* ok, int32_t* pop_point, uint32_t* n_pops, double* pose) {
  PreState* S;
  if (int rc = get_state(h, S)) return rc;
  DeviceState* D = (DeviceState*)h->dev;
  CUDA_TRY(cudaSetDevice(D->device));
  if (lidar < 0 || lidar >= MALIO_MAX_LIDAR) { h->err = "lidar slot out of range"; return MALIO_ERR_INVALID_ARG; }
  if (a->n_ctrl > MALIO_MAX_CTRL || a->n_cov > MALIO_MAX_COV) { h->err = "too many control points / covariance entries"; return MALIO_ERR_CAPACITY; }
  for (uint32_t k = 1; k < a->n_ctrl; ++k) if (!(a->ctrl_t[k] > a->ctrl_t[k - 1])) { h->err = "control-point times must ascend"; return MALIO_ERR_INVALID_ARG; }
  for (uint32_t k = 1; k < a->n_cov; ++k) if (!(a->imu_cov_t[k] >= a->imu_cov_t[k - 1])) { h->err = "imu_cov times must ascend"; return MALIO_ERR_INVALID_ARG; }
  if (a->cov_pointer < -1 || a->cov_pointer >= (int32_t)a->n_cov) { h->err = "cov_pointer outside the list"; return MALIO_ERR_INVALID_ARG; }
  PreLidar& L = S->lid[lidar];
  cudaStream_t st = S->stream;
  if (n > L.cap) {
    const uint32_t cap = n + n / 8 + 1024;
    if (int rc = grow(h, L.d_raw, cap)) return rc;
    if (int rc = grow(h, L.d_ud, (size_t)cap * 5)) return rc;
    if (int rc = grow(h, L.d_need, cap)) return rc;
    if (int rc = grow(h, L.d_idx, cap)) return rc;
    if (int rc = grow(h, L.d_ok, cap)) return rc;
    if (L.d_pose) { cudaFree(L.d_pose); L.d_pose = nullptr; }
    L.cap = cap;
  }
  if (pose && !L.d_pose) { if (int rc = grow(h, L.d_pose, (size_t)L.cap * 7)) return rc; }
  L.n = n;
  L.ds_valid = false;
  if (n == 0) { if (n_pops) *n_pops = 0; return MALIO_OK; }
  CUDA_TRY(cudaMemcpyAsync(L.d_raw, pts, (size_t)n * sizeof(malio_raw_pt), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(S->d_ct, a->ctrl_t, a->n_ctrl * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(S->d_cT, a->ctrl_T, (size_t)a->n_ctrl * 16 * sizeof(double), cudaMemcpyHostToDevice, st));
  if (a->n_cov) CUDA_TRY(cudaMemcpyAsync(S->d_covt, a->imu_cov_t, a->n_cov * sizeof(double), cudaMemcpyHostToDevice, st));
  UdConst c{};
  c.beg_time = a->beg_time;
  std::memcpy(c.eq, a->extrinsic.q, sizeof(c.eq)); std::memcpy(c.et, a->extrinsic.t, sizeof(c.et));
  std::memcpy(c.lq, a->lt_imu_frame.q, sizeof(c.lq)); std::memcpy(c.lt, a->lt_imu_frame.t, sizeof(c.lt));
  c.n_ctrl = (int)a->n_ctrl; c.n_cov = (int)a->n_cov; c.cov_pointer = a->cov_pointer;
  if (a->n_ctrl > 1) seglog_kernel<<<(a->n_ctrl + 63) / 64, 64, 0, st>>>(S->d_cT, (int)a->n_ctrl, S->d_seglog);
  const size_t smem = (size_t)(a->n_ctrl + a->n_cov) * sizeof(double);
  undistort_kernel<<<(n + UD_THREADS - 1) / UD_THREADS, UD_THREADS, smem, st>>>(L.d_raw, n, c, S->d_ct, S->d_cT, S->d_seglog, S->d_covt,
                                                                                  L.d_ud, L.d_need, L.d_ok, pose ? L.d_pose : nullptr);
  idx_scan_kernel<<<1, SCAN_T, 0, st>>>(L.d_need, L.d_ok, n, L.d_ud, L.d_idx, S->d_pop, (int)a->n_cov);
  CUDA_TRY(cudaGetLastError());
  D->ctr.kernel_launches += 3;
  D->ctr.h2d_bytes += (uint64_t)n * sizeof(malio_raw_pt) + (uint64_t)a->n_ctrl * 17 * 8 + (uint64_t)a->n_cov * 8;
  std::vector<float> ud_host;
  if (xyz) ud_host.resize((size_t)n * 5);
  std::vector<int32_t> pops(a->n_cov + 1);
  if (xyz) CUDA_TRY(cudaMemcpyAsync(ud_host.data(), L.d_ud, (size_t)n * 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (idx) CUDA_TRY(cudaMemcpyAsync(idx, L.d_idx, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (ok) CUDA_TRY(cudaMemcpyAsync(ok, L.d_ok, (size_t)n, cudaMemcpyDeviceToHost, st));
  if (pose) CUDA_TRY(cudaMemcpyAsync(pose, L.d_pose, (size_t)n * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(pops.data(), S->d_pop, (a->n_cov + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (xyz) for (uint32_t i = 0; i < n; ++i) { xyz[3 * (size_t)i] = ud_host[5 * (size_t)i]; xyz[3 * (size_t)i + 1] = ud_host[5 * (size_t)i + 1]; xyz[3 * (size_t)i + 2] = ud_host[5 * (size_t)i + 2]; }
  if (pop_point) std::memcpy(pop_point, pops.data(), a->n_cov * sizeof(int32_t));
  if (n_pops) *n_pops = (uint32_t)pops[a->n_cov];
  D->ctr.d2h_bytes += (uint64_t)n * ((xyz ? 20 : 0) + (idx ? 4 : 0) + (ok ? 1 : 0) + (pose ? 56 : 0));
  return MALIO_OK;
}

int voxel_grid(malio_handle* h, int lidar, const float* in, uint32_t n, float leaf, float* out, uint32_t out_cap, uint32_t* n_out) {
  PreState* S;
  if (int rc = get_state(h, S)) return rc;
  DeviceState* D = (DeviceState*)h->dev;
  CUDA_TRY(cudaSetDevice(D->device));
  if (lidar < 0 || lidar >= MALIO_MAX_LIDAR || !(leaf > 0.f)) { h->err = "voxel_grid: bad lidar slot or leaf size"; return MALIO_ERR_INVALID_ARG; }
  PreLidar& L = S->lid[lidar];
  cudaStream_t st = S->stream;
  const float* d_in;
  if (in) {
    if (n > S->cap_vin) { if (int rc = grow(h, S->d_vin, (size_t)(n + 1024) * 5)) return rc; S->cap_vin = n + 1024; }
    if (n) CUDA_TRY(cudaMemcpyAsync(S->d_vin, in, (size_t)n * 5 * sizeof(float), cudaMemcpyHostToDevice, st));
    D->ctr.h2d_bytes += (uint64_t)n * 20;
    d_in = S->d_vin;
  } else {
    n = L.n;
    d_in = L.d_ud;
  }
  L.ds_valid = false;
  L.n_ds = 0;
  if (n_out) *n_out = 0;
  if (n == 0) { L.ds_valid = true; return MALIO_OK; }
  if (n > S->cap_vpts) {
    const uint32_t cap = n + n / 8 + 1024;
    if (int rc = grow(h, S->d_vidx, cap)) return rc;
    if (int rc = grow(h, S->d_vtmp, cap)) return rc;
    if (int rc = grow(h, S->d_vorder, cap)) return rc;
    S->cap_vpts = cap;
  }
  // ---- bounds (getMinMax3D), then the grid geometry on the host exactly as PCL derives it
  const uint32_t init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
  CUDA_TRY(cudaMemcpyAsync(S->d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
  vg_bounds_kernel<<<(n + 255) / 256 < 296u ? (n + 255) / 256 : 296u, 256, 0, st>>>(d_in, n, S->d_bounds);
  CUDA_TRY(cudaMemcpyAsync(S->h_small, S->d_bounds, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  D->ctr.kernel_launches += 1;
  if (S->h_small[0] == 0xFFFFFFFFu) { L.ds_valid = true; return MALIO_OK; }   // no finite point
  VgConst c{};
  c.il = 1.0f / leaf;
  int maxb[3];
  for (int k = 0; k < 3; ++k) {
    c.minb[k] = (int)std::floor(fkey_inv(S->h_small[k]) * c.il);
    maxb[k] = (int)std::floor(fkey_inv(S->h_small[3 + k]) * c.il);
  }
  const long long dx = (long long)maxb[0] - c.minb[0] + 1, dy = (long long)maxb[1] - c.minb[1] + 1, dz = (long long)maxb[2] - c.minb[2] + 1;
  const long long cells = dx * dy * dz;
  if (cells <= 0 || cells > (1ll << 28)) {   // PCL gives up (returns the input) above INT_MAX voxels; the dense index here stops at 2^28
    h->err = "voxel_grid: leaf size too small for the cloud's extent (more than 2^28 voxels)";
    return MALIO_ERR_CAPACITY;
  }
  c.dx = dx; c.dxy = dx * dy;
  if ((uint64_t)cells > S->cap_cells) {
    const uint64_t cap = (uint64_t)cells + (uint64_t)cells / 4 + 4096;
    if (int rc = grow(h, S->d_vcnt, cap)) return rc;
    if (int rc = grow(h, S->d_voff, cap)) return rc;
    if (int rc = grow(h, S->d_vocc, cap)) return rc;
    if (int rc = grow(h, S->d_vcur, cap)) return rc;
    S->cap_cells = cap;
  }
  const uint32_t nchunk = (uint32_t)((cells + VS_CHUNK - 1) / VS_CHUNK);   // <= 65536
  CUDA_TRY(cudaMemsetAsync(S->d_vcnt, 0, (size_t)cells * sizeof(uint32_t), st));
  CUDA_TRY(cudaMemsetAsync(S->d_vcur, 0, (size_t)cells * sizeof(uint32_t), st));
  vg_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_in, n, c, S->d_vidx, S->d_vcnt);
  vg_scan_local_kernel<<<nchunk, VS_T, 0, st>>>(S->d_vcnt, (uint64_t)cells, S->d_voff, S->d_vocc, S->d_ctot);
  vg_scan_tot_kernel<<<1, 1024, 0, st>>>(S->d_ctot, nchunk, S->d_cbase, S->d_bounds + 6);
  vg_scan_add_kernel<<<nchunk, VS_T, 0, st>>>(S->d_voff, S->d_vocc, (uint64_t)cells, S->d_cbase);
  vg_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(S->d_vidx, n, S->d_voff, S->d_vcur, S->d_vtmp);
  vg_rank_kernel<<<(n + 255) / 256, 256, 0, st>>>(S->d_vidx, n, S->d_voff, S->d_vcnt, S->d_vtmp, S->d_vorder);
  CUDA_TRY(cudaMemcpyAsync(S->h_small, S->d_bounds + 6, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const uint32_t m = S->h_small[1];
  if (m > L.cap_ds) { if (int rc = grow(h, L.d_ds, (size_t)(m + m / 8 + 1024) * 5)) return rc; L.cap_ds = m + m / 8 + 1024; }
  vg_centroid_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_in, S->d_vidx, n, S->d_voff, S->d_vcnt, S->d_vocc, S->d_vorder, L.d_ds, m);
  CUDA_TRY(cudaGetLastError());
  D->ctr.kernel_launches += 7;
  L.n_ds = m;
  L.ds_valid = true;
  if (n_out) *n_out = m;
  if (out) {
    const uint32_t k = m < out_cap ? m : out_cap;
    CUDA_TRY(cudaMemcpyAsync(out, L.d_ds, (size_t)k * 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
    D->ctr.d2h_bytes += (uint64_t)k * 20;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return MALIO_OK;
}

int upload_scan_device(malio_handle* h, const malio_pose_entry* table, const uint32_t* table_off, const malio_rigid* tcomp,
                       uint32_t* n_total) {
  PreState* S;
  if (int rc = get_state(h, S)) return rc;
  DeviceState* D = (DeviceState*)h->dev;
  CUDA_TRY(cudaSetDevice(D->device));
  const int Ln = h->cfg.params.n_lidar;
  uint32_t total = 0;
  for (int l = 0; l < Ln; ++l) {
    if (!S->lid[l].ds_valid) { h->err = "upload_scan_device: no down-sampled cloud for a LiDAR (call malio_voxel_grid first)"; return MALIO_ERR_STATE; }
    total += S->lid[l].n_ds;
  }
  // size the scan buffers and tables through the regular entry point (no points copied), then fill d_pts on the device
  if (int rc = malio_dev::upload_scan(h, nullptr, 0, table, table_off, tcomp)) return rc;
  if (int rc = malio_dev::reserve_scan(h, total)) return rc;
  uint32_t at = 0;
  for (int l = 0; l < Ln; ++l) {
    const uint32_t m = S->lid[l].n_ds;
    if (m) merge_scan_kernel<<<(m + 255) / 256, 256, 0, S->stream>>>(S->lid[l].d_ds, m, (uint16_t)l, D->d_pts + at);
    at += m;
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  D->N = total;
  D->ctr.kernel_launches += Ln;
  if (n_total) *n_total = total;
  return MALIO_OK;
}

}  // namespace malio_pre

// host-side spline pose (same code as the device's, compiled for the host)
extern "C" int malio_bspline_get_pose(const double* ctrl_t, const double* ctrl_T, uint32_t n_ctrl, double timestamp, double q[4], double p[3]) {
  if (!ctrl_t || !ctrl_T || !q || !p || n_ctrl < 4 || n_ctrl > MALIO_MAX_CTRL) { if (p) { p[0] = p[1] = p[2] = 0; } return 0; }
  std::vector<double> seglog((size_t)n_ctrl * 6, 0.0);
  int i0, i1, i2, i3;
  if (!find_bounding(ctrl_t, (int)n_ctrl, timestamp, i0, i1, i2, i3)) { p[0] = p[1] = p[2] = 0; return 0; }
  for (int k = i0; k <= i2; ++k) {
    double ai[16], rel[16];
    inv_se3(ctrl_T + 16 * (size_t)k, ai);
    mul4(ai, ctrl_T + 16 * (size_t)(k + 1), rel);
    log_se3(rel, seglog.data() + 6 * (size_t)k);
  }
  return spline_pose(ctrl_t, ctrl_T, seglog.data(), (int)n_ctrl, timestamp, q, p) ? 1 : 0;
}
