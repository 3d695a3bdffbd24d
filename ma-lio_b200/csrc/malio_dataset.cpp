// malio_dataset.cpp — SURVEY.md §8f N4: ROS-free reading of the City dataset's per-scan .bin files and the per-sensor
// conversion to the time-stamped raw cloud UndistortPcl consumes.  Host C++, no CUDA.
//
// Reference interface mirrored (paths relative to /root/reference):
//   file_player/src/ROSThread.cpp:776-795 (Livox Avia / Tele), :952-967 (Ouster)   packed records of one scan
//   MA_LIO/src/preprocess.cpp:59-110  Preprocess::avia_handler      livox_ros_driver::CustomMsg -> pl_surf
//   MA_LIO/src/preprocess.cpp:112-152 Preprocess::oust64_handler    ouster PointCloud2           -> pl_surf
// The player reads with `while(!file.eof())`, so every scan ends with one extra default-constructed record (the read
// past the end fails and the freshly constructed point is pushed anyway): eof_quirk = 1 reproduces that (all-zero
// record appended), 0 returns the file's records only.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "malio_b200.h"

extern "C" {

}  // extern "C"

namespace {
// whole records of the file (a trailing partial record — what the player's failed last read leaves behind — is dropped, see eof_quirk).
// count_only: the size comes from the file length, nothing is read.
int slurp_records(const char* path, size_t rec_size, bool count_only, std::vector<unsigned char>& buf, uint32_t* n_rec) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return MALIO_ERR_INVALID_ARG;
  int rc = MALIO_OK;
  if (std::fseek(f, 0, SEEK_END) != 0) rc = MALIO_ERR_INVALID_ARG;
  const long sz = rc == MALIO_OK ? std::ftell(f) : -1;
  if (sz < 0) rc = MALIO_ERR_INVALID_ARG;
  if (rc == MALIO_OK) {
    const size_t n = (size_t)sz / rec_size;
    if (n > 0xFFFFFFF0u) rc = MALIO_ERR_CAPACITY;
    else {
      *n_rec = (uint32_t)n;
      if (!count_only && n) {
        buf.resize(n * rec_size);
        std::rewind(f);
        if (std::fread(buf.data(), 1, buf.size(), f) != buf.size()) rc = MALIO_ERR_INVALID_ARG;   // one read for the whole scan
      }
    }
  }
  std::fclose(f);
  return rc;
}
}  // namespace

extern "C" {

int malio_read_livox_bin(const char* path, malio_livox_pt* out, uint32_t cap, uint32_t* n_out, int eof_quirk) {
  if (!path || !n_out) return MALIO_ERR_INVALID_ARG;
  // x y z f32 | reflectivity u8 | tag u8 | line u8 | offset_time: sizeof(uint16_t) bytes (ROSThread.cpp:789)
  constexpr size_t REC = 17;
  std::vector<unsigned char> buf;
  uint32_t n = 0;
  if (int rc = slurp_records(path, REC, out == nullptr, buf, &n)) return rc;
  int rc = MALIO_OK;
  if (out) {
    const uint32_t m = n < cap ? n : cap;
    if (n > cap) rc = MALIO_ERR_CAPACITY;
    for (uint32_t i = 0; i < m; ++i) {
      const unsigned char* rec = buf.data() + (size_t)i * REC;
      malio_livox_pt p;
      std::memcpy(&p.x, rec, 12);
      p.reflectivity = rec[12]; p.tag = rec[13]; p.line = rec[14];
      uint16_t t16;
      std::memcpy(&t16, rec + 15, 2);
      p.offset_time = t16;           // the upper half of CustomPoint::offset_time stays 0
      p.pad = 0;
      out[i] = p;
    }
    if (rc != MALIO_OK) n = m;
  }
  if (rc == MALIO_OK && eof_quirk) {
    if (out) {
      if (n >= cap) rc = MALIO_ERR_CAPACITY;
      else std::memset(&out[n], 0, sizeof(malio_livox_pt));
    }
    ++n;
  }
  *n_out = n;
  return rc;
}

int malio_read_ouster_bin(const char* path, malio_ouster_pt* out, uint32_t cap, uint32_t* n_out, int eof_quirk) {
  if (!path || !n_out) return MALIO_ERR_INVALID_ARG;
  constexpr size_t REC = 22;   // x y z intensity f32 | ring u16 | t u32 (ROSThread.cpp:960-965)
  std::vector<unsigned char> buf;
  uint32_t n = 0;
  if (int rc = slurp_records(path, REC, out == nullptr, buf, &n)) return rc;
  int rc = MALIO_OK;
  if (out) {
    const uint32_t m = n < cap ? n : cap;
    if (n > cap) rc = MALIO_ERR_CAPACITY;
    for (uint32_t i = 0; i < m; ++i) {
      const unsigned char* rec = buf.data() + (size_t)i * REC;
      malio_ouster_pt p;
      std::memcpy(&p.x, rec, 16);
      std::memcpy(&p.ring, rec + 16, 2);
      std::memcpy(&p.t, rec + 18, 4);
      p.pad = 0;
      out[i] = p;
    }
    if (rc != MALIO_OK) n = m;
  }
  if (rc == MALIO_OK && eof_quirk) {
    if (out) {
      if (n >= cap) rc = MALIO_ERR_CAPACITY;
      else std::memset(&out[n], 0, sizeof(malio_ouster_pt));
    }
    ++n;
  }
  *n_out = n;
  return rc;
}

// Preprocess::avia_handler (preprocess.cpp:59-110).  out / intensity may be NULL (count only).
int malio_preprocess_livox(const malio_livox_pt* pts, uint32_t n, int n_scans, int point_filter_num, double blind, malio_raw_pt* out,
                           float* intensity, uint32_t cap, uint32_t* n_out) {
  if ((n && !pts) || !n_out || point_filter_num < 1) return MALIO_ERR_INVALID_ARG;
  uint32_t m = 0;
  unsigned valid_num = 0;
  // pl_full: value-initialised points; only the entries that pass the decimation are filled in (:84-88)
  float px = 0.f, py = 0.f, pz = 0.f;      // pl_full[i-1].{x,y,z}
  bool prev_filled = false;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (uint32_t i = 1; i < n; ++i) {
    // state of pl_full[i-1] as seen by this iteration
    px = prev_filled ? fx : 0.f; py = prev_filled ? fy : 0.f; pz = prev_filled ? fz : 0.f;
    prev_filled = false;
    const malio_livox_pt& p = pts[i];
    if ((p.line < n_scans) && ((p.tag & 0x30) == 0x10 || (p.tag & 0x30) == 0x00)) {           // :81
      valid_num++;
      if (valid_num % (unsigned)point_filter_num == 0) {                                       // :84
        fx = p.x; fy = p.y; fz = p.z;
        prev_filled = true;
        const float curv = (float)p.offset_time / float(1000000);                              // :90
        if (curv > 100) continue;                                                               // :91-92
        // :95  a || b || (c && range): && binds tighter than ||
        const bool keep = (std::fabs(fx - px) > 1e-7) || (std::fabs(fy - py) > 1e-7) ||
                          ((std::fabs(fz - pz) > 1e-7) && ((double)(fx * fx + fy * fy + fz * fz) > (blind * blind)));
        if (keep) {
          if (out) {
            if (m >= cap) { *n_out = m; return MALIO_ERR_CAPACITY; }
            out[m].x = fx; out[m].y = fy; out[m].z = fz; out[m].curvature = curv;
            if (intensity) intensity[m] = (float)p.reflectivity;
          }
          ++m;
        }
      }
    }
  }
  *n_out = m;
  return MALIO_OK;
}

// Preprocess::oust64_handler (preprocess.cpp:112-152); time_unit_scale as Preprocess::process sets it (preprocess.cpp:36-52)
int malio_preprocess_ouster(const malio_ouster_pt* pts, uint32_t n, int point_filter_num, double blind, float time_unit_scale,
                            malio_raw_pt* out, float* intensity, uint32_t cap, uint32_t* n_out) {
  if ((n && !pts) || !n_out || point_filter_num < 1) return MALIO_ERR_INVALID_ARG;
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (i % (uint32_t)point_filter_num != 0) continue;                                          // :129
    const malio_ouster_pt& p = pts[i];
    const double range = p.x * p.x + p.y * p.y + p.z * p.z;                                     // :132 (float products, widened on assignment)
    if (range < (blind * blind)) continue;
    if (out) {
      if (m >= cap) { *n_out = m; return MALIO_ERR_CAPACITY; }
      out[m].x = p.x; out[m].y = p.y; out[m].z = p.z;
      out[m].curvature = p.t * time_unit_scale * 1.e-9f;                                        // :146
      if (intensity) intensity[m] = p.intensity;
    }
    ++m;
  }
  *n_out = m;
  return MALIO_OK;
}

}  // extern "C"
