// On-disk and wire formats either side of the filter path, host code:
//   * EuRoC / ASL csv inputs: loadImuFile, loadImageList (include/utils/DataReader.hpp:30-140) and the ground-truth
//     file with its nearest-stamp lookup (DatasetReader::load_gt_file / get_gt_state, include/orcvio/dataset_reader.h:64-140);
//   * the pose log processFeatures writes (state_est_geo_feat.txt, src/orcvio.cpp:422, 640-645) -- reader; the writer
//     is orcvio_set_pose_log;
//   * orcvio_ros_msgs/ObjectLM (ros_wrapper/src/orcvio_ros_msgs/msg/ObjectLM.msg): the message the object front-end
//     sends to the filter (residual, Jacobians, camera poses, time stamps), in ROS 1 wire serialisation with the
//     matrices laid out the way the reference fills them (tf::matrixEigenToMsg: two dimensions, row-major data).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/orcvio_b200.h"

namespace {

// field k (0-based) of a comma separated line; empty when the line has fewer fields
std::string field(const std::string& line, int k) {
  size_t a = 0;
  for (int i = 0; i < k; ++i) {
    a = line.find(',', a);
    if (a == std::string::npos) return std::string();
    ++a;
  }
  const size_t b = line.find(',', a);
  return line.substr(a, b == std::string::npos ? std::string::npos : b - a);
}

struct Writer {
  std::vector<unsigned char> b;
  void raw(const void* p, size_t n) { const unsigned char* c = (const unsigned char*)p; b.insert(b.end(), c, c + n); }
  void u32(uint32_t v) { raw(&v, 4); }
  void matrix(const double* m, int rows, int cols) {          // std_msgs/Float64MultiArray as tf::matrixEigenToMsg fills it
    u32(2);
    u32(0); u32((uint32_t)rows); u32((uint32_t)(rows * cols)); // dim[0]: label "", size, stride
    u32(0); u32((uint32_t)cols); u32((uint32_t)cols);          // dim[1]
    u32(0);                                                    // data_offset
    u32((uint32_t)(rows * cols));
    raw(m, sizeof(double) * (size_t)rows * cols);
  }
};

struct Reader {
  const unsigned char* p; size_t n, o = 0; bool ok = true;
  bool raw(void* d, size_t k) {
    if (o + k > n) { ok = false; return false; }
    std::memcpy(d, p + o, k);
    o += k;
    return true;
  }
  uint32_t u32() { uint32_t v = 0; raw(&v, 4); return v; }
  // returns rows, cols (0, 0 for an empty layout); data copied row-major into out (cap doubles)
  bool matrix(double* out, int cap, int* rows, int* cols) {
    const uint32_t nd = u32();
    uint32_t size[2] = {0, 0};
    for (uint32_t d = 0; d < nd && ok; ++d) {
      const uint32_t len = u32();
      if (o + len > n) { ok = false; return false; }
      o += len;                                                // label
      const uint32_t sz = u32();
      u32();                                                   // stride
      if (d < 2) size[d] = sz;
    }
    u32();                                                     // data_offset
    const uint32_t cnt = u32();
    if (!ok || (int)cnt > cap || (nd >= 2 && (uint64_t)size[0] * size[1] != cnt)) { ok = false; return false; }
    *rows = nd >= 2 ? (int)size[0] : (int)cnt;
    *cols = nd >= 2 ? (int)size[1] : (cnt ? 1 : 0);
    return raw(out, sizeof(double) * cnt);
  }
};

}  // namespace

extern "C" {

int orcvio_read_imu_csv(const char* path, OrcvioImu* out, int cap) {
  std::ifstream f(path ? path : "");
  if (!f) return ORCVIO_ERR_ARG;
  std::string line;
  std::getline(f, line);                                       // header
  int n = 0;
  while (std::getline(f, line)) {
    if (line.empty() || line == "\r") continue;                // (the reference pushes a duplicate of the last row at eof)
    if (out && n < cap) {
      out[n].t = 1e-9 * (double)std::atol(field(line, 0).c_str());
      for (int k = 0; k < 3; ++k) {
        out[n].gyro[k] = std::atof(field(line, 1 + k).c_str());
        out[n].acc[k] = std::atof(field(line, 4 + k).c_str());
      }
    }
    ++n;
  }
  return n;
}

int orcvio_read_image_list_csv(const char* path, double* t_out, char* names, int name_stride, int cap) {
  std::ifstream f(path ? path : "");
  if (!f) return ORCVIO_ERR_ARG;
  std::string line;
  std::getline(f, line);
  int n = 0;
  while (std::getline(f, line)) {
    if (line.empty() || line == "\r") continue;
    if (n < cap) {
      if (t_out) t_out[n] = 1e-9 * (double)std::atol(field(line, 0).c_str());
      if (names && name_stride > 0) {
        std::string nm = field(line, 1);
        while (!nm.empty() && (nm.back() == '\r' || nm.back() == '\n')) nm.pop_back();
        std::snprintf(names + (size_t)n * name_stride, (size_t)name_stride, "%s", nm.c_str());
      }
    }
    ++n;
  }
  return n;
}

int orcvio_read_gt_csv(const char* path, double* out17, int cap) {
  std::ifstream f(path ? path : "");
  if (!f) return ORCVIO_ERR_ARG;
  std::string line;
  std::getline(f, line);
  int n = 0;
  while (std::getline(f, line)) {
    if (line.empty() || line == "\r") continue;
    std::istringstream s(line);
    std::string fld;
    double row[17] = {0};
    int i = 0;
    while (std::getline(s, fld, ',')) {
      if (i > 16) return ORCVIO_ERR_ARG;                       // "Invalid groundtruth line, too long"
      row[i++] = std::atof(fld.c_str());
    }
    if (out17 && n < cap) {
      std::memcpy(out17 + 17 * (size_t)n, row, sizeof(row));
      out17[17 * (size_t)n] = 1e-9 * row[0];                    // the map key: seconds
    }
    ++n;
  }
  return n;
}

int orcvio_gt_lookup(const double* gt17, int n, double t, double out17[17]) {
  if (!gt17 || n < 1 || !out17) return 0;
  double closest = INFINITY;
  int best = -1;
  for (int i = 0; i < n; ++i)
    if (std::fabs(gt17[17 * (size_t)i] - t) < std::fabs(closest - t)) { closest = gt17[17 * (size_t)i]; best = i; }
  if (std::fabs(closest - t) < 0.005) t = closest;              // close enough: use it
  if (best < 0 || gt17[17 * (size_t)best] != t) return 0;       // otherwise the stamp itself has to be in the file
  std::memcpy(out17, gt17 + 17 * (size_t)best, 17 * sizeof(double));
  return 1;
}

int orcvio_read_pose_log(const char* path, double* out8, int cap) {
  std::ifstream f(path ? path : "");
  if (!f) return ORCVIO_ERR_ARG;
  int n = 0;
  double r[8];
  while (f >> r[0] >> r[1] >> r[2] >> r[3] >> r[4] >> r[5] >> r[6] >> r[7]) {
    if (out8 && n < cap) std::memcpy(out8 + 8 * (size_t)n, r, sizeof(r));
    ++n;
  }
  return n;
}

int orcvio_objectlm_pack(long long object_id, const double* residual, int rows, const double* jac_object, int odim,
                         const double* jac_sensor, const double* cam_pose_se3, int n_poses, const double* timestamps,
                         int n_ts, const int* zs_num, int n_zs, unsigned char* buf, int cap) {
  if (rows < 0 || odim < 0 || n_poses < 0 || n_ts < 0 || n_zs < 0) return ORCVIO_ERR_ARG;
  Writer w;
  const int64_t id = object_id;
  w.raw(&id, 8);
  w.matrix(residual, rows, 1);
  w.matrix(jac_object, rows, odim);
  w.matrix(jac_sensor, rows, 6);
  w.matrix(cam_pose_se3, 6, n_poses);
  w.u32((uint32_t)n_ts);
  w.raw(timestamps, sizeof(double) * (size_t)n_ts);
  w.u32((uint32_t)n_zs);
  w.raw(zs_num, sizeof(int32_t) * (size_t)n_zs);
  if (buf && (int)w.b.size() <= cap) std::memcpy(buf, w.b.data(), w.b.size());
  return (int)w.b.size();
}

int orcvio_objectlm_unpack(const unsigned char* buf, int len, long long* object_id, double* residual, int* rows,
                           double* jac_object, int* odim, double* jac_sensor, double* cam_pose_se3, int* n_poses,
                           double* timestamps, int* n_ts, int* zs_num, int* n_zs, int cap_rows, int cap_odim, int cap_n) {
  if (!buf || len < 8) return ORCVIO_ERR_ARG;
  Reader r{buf, (size_t)len};
  int64_t id = 0;
  r.raw(&id, 8);
  if (object_id) *object_id = id;
  int rr = 0, cc = 0, r2 = 0, c2 = 0, r3 = 0, c3 = 0, r4 = 0, c4 = 0;
  if (!r.matrix(residual, cap_rows, &rr, &cc)) return ORCVIO_ERR_ARG;
  if (!r.matrix(jac_object, cap_rows * cap_odim, &r2, &c2)) return ORCVIO_ERR_ARG;
  if (!r.matrix(jac_sensor, cap_rows * 6, &r3, &c3)) return ORCVIO_ERR_ARG;
  if (!r.matrix(cam_pose_se3, 6 * cap_n, &r4, &c4)) return ORCVIO_ERR_ARG;
  if (r2 != rr || r3 != rr || (rr && c3 != 6) || (c4 && r4 != 6)) return ORCVIO_ERR_ARG;
  const uint32_t nt = r.u32();
  if (!r.ok || (int)nt > cap_n || !r.raw(timestamps, sizeof(double) * nt)) return ORCVIO_ERR_ARG;
  const uint32_t nz = r.u32();
  if (!r.ok || (int)nz > cap_n || !r.raw(zs_num, sizeof(int32_t) * nz)) return ORCVIO_ERR_ARG;
  if (rows) *rows = rr;
  if (odim) *odim = c2;
  if (n_poses) *n_poses = c4;
  if (n_ts) *n_ts = (int)nt;
  if (n_zs) *n_zs = (int)nz;
  return r.o == (size_t)len ? ORCVIO_OK : ORCVIO_ERR_ARG;
}

}  // extern "C"
