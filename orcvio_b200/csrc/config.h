// Filter parameters with the reference's config/*.yaml semantics
// (OrcVIO::loadParameters, reference src/orcvio.cpp:62-329).
#pragma once
#include <map>
#include <string>
#include <vector>

namespace ob {

struct Params {
  int use_left_perturbation_flag = 0, use_closed_form_cov_prop_flag = 1, use_larvio_flag = 0;
  int discard_large_update_flag = 0;
  double features_rate = 10, imu_rate = 200, imu_img_timeTh = 0.0025;
  double rotation_threshold = 0.2618, translation_threshold = 0.4, tracking_rate_threshold = 0.5;
  int max_track_len = 6;
  double feature_translation_threshold = -1, feature_cost_threshold = 4.7673e-4,
         init_final_dist_threshold = 5;
  double td = 0;
  bool estimate_td = false, estimate_extrin = false, calib_imu = false, if_FEJ = false;
  double imu_gyro_noise = 0, imu_acc_noise = 0, imu_gyro_bias_noise = 0, imu_acc_bias_noise = 0,
         feature_observation_noise = 0;   // variances
  double zupt_noise_v = 0, zupt_noise_p = 0, zupt_noise_q = 0;
  bool initial_use_gt = false;
  double initial_state_time = 0;
  double init_bg[3] = {0, 0, 0}, init_ba[3] = {0, 0, 0}, init_pos[3] = {0, 0, 0}, init_vel[3] = {0, 0, 0};
  double init_quat[4] = {0, 0, 0, 1};
  bool prediction_only_flag = false;
  double cov_orientation = 0, cov_position = 0, cov_velocity = 0, cov_gyro_bias = 0, cov_acc_bias = 0;
  double cov_extrin_rot = 0, cov_extrin_trans = 0;
  double R_imu_cam0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // R_b2c
  double t_cam0_imu[3] = {0, 0, 0};                     // camera origin in the body frame
  int sw_size = 20, least_Obs_Num = 3;
  bool if_ZUPT_valid = false, if_use_feature_zupt_flag = false;
  double zupt_max_feature_dis = 2e-3;
  int use_object_residual_update_cam_pose_flag = 0;
  int grid_rows = 0, grid_cols = 0, max_features = 0, feature_idp_dim = 1;
  // boundary of the normalised image plane and the cell size of the EKF-feature grid (src/orcvio.cpp:293-311)
  double x_min = 0, y_min = 0, grid_width = 1, grid_height = 1;
  bool use_schmidt = false;
  double chi_square_threshold_feat = 0.95;
  std::string output_dir;
};

// Parses an OpenCV-style yaml (`%YAML:1.0`, scalars, one-level maps, !!opencv-matrix).
// Returns false (and fills err) when the file cannot be read or a required key is missing.
bool load_params(const std::string& path, Params& p, std::string& err);

// chi-square quantile (inverse regularised lower incomplete gamma), replaces
// boost::math::quantile(chi_squared(dof), p) at reference src/orcvio.cpp:486-494.
double chi2_quantile(double p, int dof);

}  // namespace ob
