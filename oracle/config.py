"""CPU oracle -- config loading (TEST INFRASTRUCTURE, not product code).

Follows OrcVIO::loadParameters, src/orcvio.cpp:62-329: same keys, same squaring of
the noise standard deviations, same T_cam_imu handling.  Parsed with
cv2.FileStorage (the reference uses cv::FileStorage), which is independent of the
product's own C++ yaml reader.
"""
from types import SimpleNamespace
import numpy as np


def load_params(path):
    import cv2
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    if not fs.isOpened():
        raise FileNotFoundError(path)

    def num(key, default=0.0):
        n = fs.getNode(key)
        return default if n.empty() else n.real()

    def mat(key):
        n = fs.getNode(key)
        return None if n.empty() else np.asarray(n.mat(), dtype=float)

    p = SimpleNamespace()
    p.use_left_perturbation_flag = int(num("use_left_perturbation_flag"))
    p.use_closed_form_cov_prop_flag = int(num("use_closed_form_cov_prop_flag"))
    p.use_larvio_flag = int(num("use_larvio_flag"))
    p.discard_large_update_flag = int(num("discard_large_update_flag"))
    p.features_rate = num("pub_frequency")
    p.imu_rate = num("imu_rate")
    p.imu_img_timeTh = 1 / (2 * p.imu_rate)
    p.rotation_threshold = num("rotation_threshold")
    p.translation_threshold = num("translation_threshold")
    p.tracking_rate_threshold = num("tracking_rate_threshold")
    p.max_track_len = int(num("max_track_len"))
    p.feature_translation_threshold = num("feature_translation_threshold")
    p.feature_cost_threshold = num("feature_cost_threshold")
    p.init_final_dist_threshold = num("init_final_dist_threshold")
    p.td = num("td")
    p.estimate_td = int(num("estimate_td")) != 0
    p.estimate_extrin = int(num("estimate_extrin")) != 0
    # variances (orcvio.cpp:108-121)
    p.imu_gyro_noise = num("noise_gyro") ** 2
    p.imu_acc_noise = num("noise_acc") ** 2
    p.imu_gyro_bias_noise = num("noise_gyro_bias") ** 2
    p.imu_acc_bias_noise = num("noise_acc_bias") ** 2
    p.feature_observation_noise = num("noise_feature") ** 2
    p.zupt_noise_v = num("zupt_noise_v") ** 2
    p.zupt_noise_p = num("zupt_noise_p") ** 2
    p.zupt_noise_q = num("zupt_noise_q") ** 2
    p.initial_use_gt = int(num("initial_use_gt")) != 0
    if p.initial_use_gt:
        p.initial_state_time = num("initial_state_time")
        p.init_gyro_bias = mat("initial_bg").ravel()[:3]
        p.init_acc_bias = mat("initial_ba").ravel()[:3]
        p.init_position = mat("initial_pos").ravel()[:3]
        p.init_velocity = mat("initial_vel").ravel()[:3]
        p.init_orientation = mat("initial_quat").ravel()[:4]  # [x,y,z,w]
    p.prediction_only_flag = int(num("prediction_only_flag")) != 0
    p.orientation_cov = num("initial_covariance_orientation")
    p.position_cov = num("initial_covariance_position")
    p.velocity_cov = num("initial_covariance_velocity")
    p.gyro_bias_cov = num("initial_covariance_gyro_bias")
    p.acc_bias_cov = num("initial_covariance_acc_bias")
    p.extrinsic_rotation_cov = num("initial_covariance_extrin_rot")
    p.extrinsic_translation_cov = num("initial_covariance_extrin_trans")
    p.calib_imu = int(num("calib_imu_instrinsic")) != 0
    p.LEG_DIM = 46 if p.calib_imu else 22
    # extrinsics (orcvio.cpp:232-246): R_imu_cam0 = R_b2c, t_cam0_imu = cam origin in body
    T = mat("T_cam_imu")
    T_imu_cam0_R = T[:3, :3]
    T_imu_cam0_t = T[:3, 3]
    inv_R = T_imu_cam0_R.T
    inv_t = -inv_R @ T_imu_cam0_t
    p.R_imu_cam0 = inv_R.T.copy()
    p.t_cam0_imu = inv_t.copy()
    p.sw_size = int(num("sw_size"))
    p.if_FEJ_config = int(num("if_FEJ")) != 0
    p.least_Obs_Num = int(num("least_observation_number"))
    p.if_ZUPT_valid = int(num("if_ZUPT_valid")) != 0
    p.if_use_feature_zupt_flag = int(num("if_use_feature_zupt_flag")) != 0
    p.zupt_max_feature_dis = num("zupt_max_feature_dis")
    p.use_object_residual_update_cam_pose_flag = int(num("use_object_residual_update_cam_pose_flag"))
    p.grid_rows = int(num("aug_grid_rows"))
    p.grid_cols = int(num("aug_grid_cols"))
    p.max_features = max(0, int(num("max_features_in_one_grid")))
    p.cam_resolution = [int(num("resolution_width")), int(num("resolution_height"))]
    n = fs.getNode("intrinsics")
    p.cam_intrinsics = [n.getNode(k).real() for k in ("fx", "fy", "cx", "cy")]
    fx, fy, cx, cy = p.cam_intrinsics
    U, V = p.cam_resolution
    p.x_min, p.y_min = -cx / fx, -cy / fy
    p.x_max, p.y_max = (U - cx) / fx, (V - cy) / fy
    if p.grid_rows * p.grid_cols != 0:
        p.grid_width = (p.x_max - p.x_min) / p.grid_cols
        p.grid_height = (p.y_max - p.y_min) / p.grid_rows
    else:
        p.grid_width = p.x_max - p.x_min
        p.grid_height = p.y_max - p.y_min
    p.feature_idp_dim = int(num("feature_idp_dim"))
    if p.feature_idp_dim not in (1, 3):
        p.feature_idp_dim = 3
    p.use_schmidt = int(num("use_schmidt")) != 0
    p.chi_square_threshold_feat = num("chi_square_threshold_feat")
    fs.release()
    return p
