// yaml reader + chi-square quantile (host only).
#include "config.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace ob {

namespace {

struct Yaml {
  std::map<std::string, std::string> scalars;                 // "key" or "parent.key"
  std::map<std::string, std::vector<double>> matrices;
};

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

std::string strip_comment(const std::string& s) {
  bool in_q = false;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '"') in_q = !in_q;
    if (s[i] == '#' && !in_q) return s.substr(0, i);
  }
  return s;
}

bool parse_yaml(const std::string& path, Yaml& y) {
  std::ifstream in(path);
  if (!in.good()) return false;
  std::string line;
  std::string parent;            // current one-level map / matrix key
  bool in_matrix = false;
  std::string data_accum;
  bool in_data = false;
  auto finish_data = [&]() {
    std::string d = data_accum;
    for (auto& ch : d)
      if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
    std::istringstream ss(d);
    std::vector<double> v;
    std::string tok;
    while (ss >> tok) v.push_back(std::strtod(tok.c_str(), nullptr));
    y.matrices[parent] = v;
    in_data = false;
    data_accum.clear();
  };
  while (std::getline(in, line)) {
    std::string raw = strip_comment(line);
    if (trim(raw).empty()) continue;
    if (raw[0] == '%' || trim(raw) == "---") continue;
    if (in_data) {
      data_accum += " " + raw;
      if (raw.find(']') != std::string::npos) finish_data();
      continue;
    }
    const bool indented = (raw[0] == ' ' || raw[0] == '\t');
    std::string t = trim(raw);
    size_t colon = t.find(':');
    if (colon == std::string::npos) continue;
    std::string key = trim(t.substr(0, colon));
    std::string val = trim(t.substr(colon + 1));
    if (!indented) {
      in_matrix = false;
      if (val.empty()) { parent = key; continue; }
      if (val.find("!!opencv-matrix") != std::string::npos) { parent = key; in_matrix = true; continue; }
      if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
      y.scalars[key] = val;
      parent.clear();
    } else {
      if (in_matrix && key == "data") {
        data_accum = val;
        in_data = true;
        if (val.find(']') != std::string::npos) finish_data();
      } else if (!in_matrix && !parent.empty()) {
        y.scalars[parent + "." + key] = val;
      }
    }
  }
  return true;
}

double num(const Yaml& y, const std::string& k, double dflt = 0.0) {
  auto it = y.scalars.find(k);
  if (it == y.scalars.end()) return dflt;     // cv::FileNode of a missing key converts to 0
  return std::strtod(it->second.c_str(), nullptr);
}

}  // namespace

bool load_params(const std::string& path, Params& p, std::string& err) {
  Yaml y;
  if (!parse_yaml(path, y)) {
    err = "config_file error: cannot open " + path;
    return false;
  }
  p.use_left_perturbation_flag = (int)num(y, "use_left_perturbation_flag");
  p.use_closed_form_cov_prop_flag = (int)num(y, "use_closed_form_cov_prop_flag");
  p.use_larvio_flag = (int)num(y, "use_larvio_flag");
  p.discard_large_update_flag = (int)num(y, "discard_large_update_flag");
  p.features_rate = num(y, "pub_frequency");
  p.imu_rate = num(y, "imu_rate");
  p.imu_img_timeTh = 1 / (2 * p.imu_rate);
  p.rotation_threshold = num(y, "rotation_threshold");
  p.translation_threshold = num(y, "translation_threshold");
  p.tracking_rate_threshold = num(y, "tracking_rate_threshold");
  p.max_track_len = (int)num(y, "max_track_len");
  p.feature_translation_threshold = num(y, "feature_translation_threshold");
  p.feature_cost_threshold = num(y, "feature_cost_threshold");
  p.init_final_dist_threshold = num(y, "init_final_dist_threshold");
  p.td = num(y, "td");
  p.estimate_td = (int)num(y, "estimate_td") != 0;
  p.estimate_extrin = (int)num(y, "estimate_extrin") != 0;
  p.calib_imu = (int)num(y, "calib_imu_instrinsic") != 0;
  p.if_FEJ = (int)num(y, "if_FEJ") != 0;
  auto sq = [](double v) { return v * v; };
  p.imu_gyro_noise = sq(num(y, "noise_gyro"));
  p.imu_acc_noise = sq(num(y, "noise_acc"));
  p.imu_gyro_bias_noise = sq(num(y, "noise_gyro_bias"));
  p.imu_acc_bias_noise = sq(num(y, "noise_acc_bias"));
  p.feature_observation_noise = sq(num(y, "noise_feature"));
  p.zupt_noise_v = sq(num(y, "zupt_noise_v"));
  p.zupt_noise_p = sq(num(y, "zupt_noise_p"));
  p.zupt_noise_q = sq(num(y, "zupt_noise_q"));
  p.initial_use_gt = (int)num(y, "initial_use_gt") != 0;
  if (p.initial_use_gt) {
    p.initial_state_time = num(y, "initial_state_time");
    auto vec = [&](const char* k, double* out, int n) {
      auto it = y.matrices.find(k);
      if (it == y.matrices.end() || (int)it->second.size() < n) return false;
      for (int i = 0; i < n; ++i) out[i] = it->second[i];
      return true;
    };
    if (!vec("initial_bg", p.init_bg, 3) || !vec("initial_ba", p.init_ba, 3) ||
        !vec("initial_pos", p.init_pos, 3) || !vec("initial_vel", p.init_vel, 3) ||
        !vec("initial_quat", p.init_quat, 4)) {
      err = "initial_use_gt: 1 but initial_* matrices missing";
      return false;
    }
  }
  p.prediction_only_flag = (int)num(y, "prediction_only_flag") != 0;
  p.cov_orientation = num(y, "initial_covariance_orientation");
  p.cov_position = num(y, "initial_covariance_position");
  p.cov_velocity = num(y, "initial_covariance_velocity");
  p.cov_gyro_bias = num(y, "initial_covariance_gyro_bias");
  p.cov_acc_bias = num(y, "initial_covariance_acc_bias");
  p.cov_extrin_rot = num(y, "initial_covariance_extrin_rot");
  p.cov_extrin_trans = num(y, "initial_covariance_extrin_trans");
  auto itT = y.matrices.find("T_cam_imu");
  if (itT == y.matrices.end() || itT->second.size() < 16) {
    err = "T_cam_imu missing";
    return false;
  }
  {
    // src/orcvio.cpp:232-246: T_imu_cam0 = yaml matrix; T_cam0_imu = its inverse;
    // R_imu_cam0 = T_cam0_imu.linear()^T (= yaml rotation), t_cam0_imu = T_cam0_imu.translation()
    const std::vector<double>& T = itT->second;
    double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    double t[3] = {T[3], T[7], T[11]};
    // inverse: Rinv = R^T, tinv = -R^T t ; R_imu_cam0 = Rinv^T
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) p.R_imu_cam0[3 * i + j] = R[3 * i + j];
    for (int i = 0; i < 3; ++i) p.t_cam0_imu[i] = -((R[i] * t[0] + R[3 + i] * t[1]) + R[6 + i] * t[2]);
  }
  p.sw_size = (int)num(y, "sw_size");
  p.least_Obs_Num = (int)num(y, "least_observation_number");
  p.if_ZUPT_valid = (int)num(y, "if_ZUPT_valid") != 0;
  p.if_use_feature_zupt_flag = (int)num(y, "if_use_feature_zupt_flag") != 0;
  p.zupt_max_feature_dis = num(y, "zupt_max_feature_dis");
  p.use_object_residual_update_cam_pose_flag = (int)num(y, "use_object_residual_update_cam_pose_flag");
  p.grid_rows = (int)num(y, "aug_grid_rows");
  p.grid_cols = (int)num(y, "aug_grid_cols");
  p.max_features = (int)num(y, "max_features_in_one_grid");
  if (p.max_features < 0) p.max_features = 0;
  {
    const double fx = num(y, "intrinsics.fx"), fy = num(y, "intrinsics.fy"), cx = num(y, "intrinsics.cx"),
                 cy = num(y, "intrinsics.cy");
    const int U = (int)num(y, "resolution_width"), V = (int)num(y, "resolution_height");
    if (fx != 0 && fy != 0) {
      p.x_min = -cx / fx;
      p.y_min = -cy / fy;
      const double x_max = (U - cx) / fx, y_max = (V - cy) / fy;
      if (p.grid_rows * p.grid_cols != 0) {
        p.grid_width = (x_max - p.x_min) / p.grid_cols;
        p.grid_height = (y_max - p.y_min) / p.grid_rows;
      } else {
        p.grid_width = x_max - p.x_min;
        p.grid_height = y_max - p.y_min;
      }
    }
  }
  p.feature_idp_dim = (int)num(y, "feature_idp_dim");
  if (p.feature_idp_dim != 1 && p.feature_idp_dim != 3) p.feature_idp_dim = 3;
  p.use_schmidt = (int)num(y, "use_schmidt") != 0;
  p.chi_square_threshold_feat = num(y, "chi_square_threshold_feat");
  auto od = y.scalars.find("output_dir");
  if (od != y.scalars.end()) p.output_dir = od->second;
  return true;
}

// ---- chi-square quantile -------------------------------------------------------------
namespace {

// regularised lower incomplete gamma P(a, x)
double gamma_p(double a, double x) {
  if (x <= 0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double sum = 1.0 / a, term = sum, ap = a;
    for (int n = 0; n < 2000; ++n) {
      ap += 1.0;
      term *= x / ap;
      sum += term;
      if (std::fabs(term) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  // continued fraction for Q (modified Lentz)
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 2000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  const double q = std::exp(-x + a * std::log(x) - lg) * h;
  return 1.0 - q;
}

}  // namespace

double chi2_quantile(double p, int dof) {
  const double a = 0.5 * dof;
  // Wilson-Hilferty start
  const double z = [&]() {
    // inverse normal CDF (Acklam), enough for a starting point
    const double q = p - 0.5;
    if (std::fabs(q) <= 0.425) {
      const double r = 0.180625 - q * q;
      return q * (((((((2509.0809287301226727 * r + 33430.575583588128105) * r + 67265.770927008700853) * r +
                       45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                    133.14166789178437745) * r + 3.387132872796366608) /
             (((((((5226.495278852545925 * r + 28729.085735721942674) * r + 39307.89580009271061) * r +
                  21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
               42.313330701600911252) * r + 1.0);
    }
    double r = q < 0 ? p : 1 - p;
    r = std::sqrt(-std::log(r));
    double v;
    if (r <= 5) {
      r -= 1.6;
      v = (((((((7.7454501427834140764e-4 * r + 0.0227238449892691845833) * r + 0.24178072517745061177) * r +
               1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
            4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((1.05075007164441684324e-9 * r + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r +
               0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r +
            2.05319162663775882187) * r + 1.0);
    } else {
      r -= 5;
      v = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r +
               0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r +
            5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((2.04426310338993978564e-15 * r + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
               7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r +
            0.59983226555248595313) * r + 1.0);
    }
    return q < 0 ? -v : v;
  }();
  const double k = dof;
  double wh = 1.0 - 2.0 / (9.0 * k) + z * std::sqrt(2.0 / (9.0 * k));
  double x = 0.5 * k * wh * wh * wh;     // in gamma units (chi2 / 2)
  if (!(x > 0)) x = 0.5 * k;
  double lo = 0.0, hi = std::max(4.0 * a + 50.0, 4.0 * x);
  const double lg = std::lgamma(a);
  for (int it = 0; it < 200; ++it) {
    const double f = gamma_p(a, x) - p;
    if (f > 0) hi = x; else lo = x;
    const double pdf = std::exp(-x + (a - 1.0) * std::log(x) - lg);
    double xn = x - f / pdf;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (std::fabs(xn - x) <= 1e-16 * std::fabs(x)) { x = xn; break; }
    x = xn;
  }
  return 2.0 * x;
}

}  // namespace ob
