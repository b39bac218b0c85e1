"""CPU oracle -- the OrcVIO filter update (TEST INFRASTRUCTURE, not product code).

NumPy restatement of the reference's `orcvio::OrcVIO` hot path, one method per
reference function with the same name, dense like the reference (dense H_j, dense
H P H^T gate, Householder QR compression, P <- (I-KH)P + symmetrise).  Citations
are src/orcvio.cpp:line unless stated otherwise.

Scope of this restatement: LEG_DIM = 22 (calib_imu_instrinsic = 0, as in every shipped
config), closed-form covariance propagation, both ZUPT variants, and both feature modes:
pure MSCKF (max_features_in_one_grid = 0) and the hybrid MSCKF / EKF-SLAM mode of
euroc.yaml / kitti_odom.yaml (max_features_in_one_grid > 0, feature_idp_dim = 1,
use_schmidt = 0: grid map, initializeInvParamPosition, featureJacobian_ekf[_new],
new-feature sparsification, measurementUpdate_hybrid with delayed initialisation,
re-anchoring, rmLostFeaturesCov).  The arithmetic of the hybrid rows lives in hybrid.py.

parity: the reference cannot be built in this image (needs Eigen, SuiteSparse, Boost,
Sophus, OpenCV C++, Ceres -- none present), so this file is pinned only where the
reference's own tests pin the path: constructObjectResidualJacobians
(src/tests/test_state_update.cpp:16-103) and the nullspace projection property
(:106-212).  Everything else here is "parity UNPINNED by the reference" and is pinned
by agreement between this oracle, the C restatement and the CUDA path.
"""
import math
from types import SimpleNamespace

import numpy as np

from . import mathutils as mu
from . import feature as feat
from . import hybrid as hyb
from .config import load_params


class Feature:
    """struct Feature, include/orcvio/feat/feature.hpp:34-265 (fields used by the path)."""

    def __init__(self, fid):
        self.id = fid
        self.observations = {}      # state_id -> np.array([u, v])   (std::map: ascending id)
        self.observations_vel = {}
        self.position = np.zeros(3)
        self.is_initialized = False
        self.id_anchor = -1
        self.in_state = False
        self.ekf_feature = False
        self.totalObsNum = 0
        self.tri_log = None
        self.invDepth = 0.0              # 1-D inverse depth in the anchor camera (feature.hpp:243)
        self.obs_anchor = np.array([0.0, 0.0, 1.0])   # bearing (x, y, 1) in the anchor camera (:246)

    def obs_ids(self):
        return sorted(self.observations.keys())


class Clone:
    """IMUState_Aug, include/orcvio/imu_state.h:100-153."""

    def __init__(self, sid):
        self.id = sid
        self.time = 0.0
        self.dt = 0.0
        self.orientation = np.eye(3)
        self.position = np.zeros(3)
        self.R_imu_cam0 = np.eye(3)
        self.t_cam0_imu = np.zeros(3)
        self.orientation_cam = np.eye(3)
        self.position_cam = np.zeros(3)


class OracleVIO:
    def __init__(self, config_file, overrides=None):
        self.config_file = config_file
        self.overrides = overrides or {}
        self.is_gravity_set = False
        self.log = []          # per-update diagnostics used by parity tests

    # ------------------------------------------------------------------ initialize
    def initialize(self):
        """OrcVIO::initialize :418-497 (+ loadParameters :62-415)."""
        try:
            p = load_params(self.config_file)
        except FileNotFoundError:
            return False
        for k, v in self.overrides.items():
            setattr(p, k, v)
        if p.calib_imu:
            raise NotImplementedError("calib_imu_instrinsic=1 (LEG_DIM 46) not restated")
        self.p = p
        self.LEG_DIM = p.LEG_DIM
        self.opt = feat.default_opt_config()
        self.opt.translation_threshold = p.feature_translation_threshold
        self.opt.cost_threshold = p.feature_cost_threshold
        self.opt.init_final_dist_threshold = p.init_final_dist_threshold
        L = self.LEG_DIM
        P = np.zeros((L, L))
        P[0:3, 0:3] = np.eye(3) * p.orientation_cov
        P[3:6, 3:6] = np.eye(3) * p.velocity_cov
        P[6:9, 6:9] = np.eye(3) * p.position_cov
        P[9:12, 9:12] = np.eye(3) * p.gyro_bias_cov
        P[12:15, 12:15] = np.eye(3) * p.acc_bias_cov
        if p.estimate_extrin:
            P[15:18, 15:18] = np.eye(3) * p.extrinsic_rotation_cov
            P[18:21, 18:21] = np.eye(3) * p.extrinsic_translation_cov
        if p.estimate_td:
            P[21, 21] = 4e-6
        self.state_cov = P
        s = SimpleNamespace()
        s.id = 0
        s.time = 0.0
        s.dt = 0.0
        s.orientation = np.eye(3)
        s.position = np.zeros(3)
        s.velocity = np.zeros(3)
        s.gyro_bias = np.zeros(3)
        s.acc_bias = np.zeros(3)
        s.R_imu_cam0 = p.R_imu_cam0.copy()
        s.t_cam0_imu = p.t_cam0_imu.copy()
        self.imu_state = s
        self.imu_state_old = self._copy_imu(s)
        self.td = p.td
        self.gravity = np.array([0.0, 0.0, -mu.GRAVITY_ACCELERATION])
        self.next_id = 0
        self.Qc = np.zeros((12, 12))
        self.Qc[0:3, 0:3] = np.eye(3) * p.imu_gyro_noise
        self.Qc[3:6, 3:6] = np.eye(3) * p.imu_acc_noise
        self.Qc[6:9, 6:9] = np.eye(3) * p.imu_gyro_bias_noise
        self.Qc[9:12, 9:12] = np.eye(3) * p.imu_acc_bias_noise
        self.if_ZUPT = False
        self.bFirstFeatures = False
        self.chi_table_feat = mu.chi2_table(p.chi_square_threshold_feat)
        self.chi_table_zupt = mu.chi2_table(0.95)
        self.map_server = {}
        self.clones = {}               # imu_states_augment
        self.feature_states = []
        self.cur_window_timestamps = []
        self.coarse_feature_dis = []
        self.imu_recent_zupt = []
        self.tracking_rate = 0.0
        self.m_gyro_old = np.zeros(3)
        self.m_acc_old = np.zeros(3)
        self.dcampose_dimupose_fixed_flag = False
        self.dcampose_dimupose_mat = np.eye(6)
        self.pose_log = []             # (t - take_off, p, q[x,y,z,w])  -- state_est_geo_feat.txt
        self.grid_map = {}             # code -> [feature ids]  (std::map<int, vector>: any code is a valid key)
        if p.max_features * p.grid_rows * p.grid_cols != 0:
            if p.feature_idp_dim != 1:
                raise NotImplementedError("feature_idp_dim = 3 not restated (every shipped yaml uses 1)")
            if p.use_schmidt:
                raise NotImplementedError("use_schmidt = 1 not restated (every shipped yaml uses 0)")
        return True

    @staticmethod
    def _copy_imu(s):
        c = SimpleNamespace(**s.__dict__)
        for k, v in c.__dict__.items():
            if isinstance(v, np.ndarray):
                setattr(c, k, v.copy())
        return c

    # test hooks, include/orcvio/orcvio.h:101-119
    def setStateCov(self, imu_dim, num_clone):
        self.LEG_DIM = imu_dim
        d = imu_dim + 6 * num_clone
        self.state_cov = np.eye(d)

    def setWinPoseTimestamps(self, ts):
        self.cur_window_timestamps = list(ts)

    def fixDcamposeDimuposeToI(self):
        self.dcampose_dimupose_mat = np.eye(6)
        self.dcampose_dimupose_fixed_flag = True

    # ------------------------------------------------------------------ frame
    def processFeatures(self, msg, imu_buf):
        """OrcVIO::processFeatures :500-661.  msg = (t, [(id,u,v,u_init,v_init,u_vel,
        v_vel,u_init_vel,v_init_vel), ...]); imu_buf = list of (t, w(3), a(3)), mutated."""
        p = self.p
        t_img, feats = msg
        if not self.bFirstFeatures:
            if len(imu_buf) and (imu_buf[0][0] - t_img - self.td <= 0.0):
                self.bFirstFeatures = True
            else:
                return False
        if not self.is_gravity_set:
            if not p.initial_use_gt:
                return False           # static/dynamic initialisers are out of scope
            s = self.imu_state
            s.time = p.initial_state_time
            s.gyro_bias = np.array(p.init_gyro_bias, dtype=float)
            s.acc_bias = np.array(p.init_acc_bias, dtype=float)
            s.position = np.array(p.init_position, dtype=float)
            s.velocity = np.array(p.init_velocity, dtype=float)
            s.orientation = mu.quaternion_to_rotation(np.array(p.init_orientation, dtype=float))
            useful = 0
            for m in imu_buf:
                if m[0] > p.initial_state_time:
                    break
                useful += 1
            if useful >= len(imu_buf):
                useful -= 1
            self.m_gyro_old = np.array(imu_buf[useful][1], dtype=float)
            self.m_acc_old = np.array(imu_buf[useful][2], dtype=float)
            del imu_buf[:useful]
            self.is_gravity_set = True
            self.take_off_stamp = s.time
            self.last_ZUPT_time = s.time
            self.last_update_time = s.time

        self.batchImuProcessing(t_img + self.td, imu_buf)
        if not p.prediction_only_flag:
            self.addFeatureObservations(feats)
        self.stateAugmentation()
        if p.if_ZUPT_valid:
            if p.if_use_feature_zupt_flag:
                self.if_ZUPT = self.checkZUPTFeat()
            else:
                self.if_ZUPT = self.checkZUPTIMU()
        self.removeLostFeatures()
        self.pruneImuStateBuffer()
        s = self.imu_state
        self.pose_log.append((s.time - self.take_off_stamp, s.position.copy(),
                              mu.rotation_to_quaternion(s.orientation)))
        return True

    # ------------------------------------------------------------------ stage 6
    def batchImuProcessing(self, time_bound, imu_buf):
        """:664-724."""
        used = 0
        dt = 0.0
        self.imu_recent_zupt = []
        for m in imu_buf:
            t = m[0]
            if t <= self.imu_state.time:
                used += 1
                continue
            if t - time_bound > self.p.imu_img_timeTh:
                break
            self.imu_recent_zupt.append(m)
            dt = t - time_bound
            g = np.array(m[1], dtype=float)
            a = np.array(m[2], dtype=float)
            self.processModel(t, g, a)
            used += 1
            self.m_gyro_old = g
            self.m_acc_old = a
        self.imu_state.id = self.next_id
        self.next_id += 1
        self.imu_state.dt = dt
        del imu_buf[:used]

    def processModel(self, time, m_gyro, m_acc):
        """:727-823 (Ma = Tg = I, As = 0 as hard-coded at :167-169)."""
        p = self.p
        s = self.imu_state
        f = m_acc - s.acc_bias
        acc = f
        w = m_gyro - s.gyro_bias
        gyro = w
        f_old = self.m_acc_old - s.acc_bias
        acc_old = f_old
        w_old = self.m_gyro_old - s.gyro_bias
        gyro_old = w_old
        dtime = time - s.time
        if not p.use_larvio_flag:
            self.predictNewStateOrcVIO(dtime, gyro, acc)
        else:
            self.predictNewStateLARVIO(dtime, gyro, acc)
        Phi = self.calPhiClosedForm(dtime, acc, gyro, acc_old, gyro_old)
        L = self.LEG_DIM
        C = self.imu_state_old.orientation
        G = np.zeros((L, 12))
        if p.use_larvio_flag or p.use_left_perturbation_flag:
            G[0:3, 0:3] = -C
        else:
            G[0:3, 0:3] = -np.eye(3)
        G[3:6, 3:6] = -C
        G[9:12, 6:9] = np.eye(3)
        G[12:15, 9:12] = np.eye(3)
        Q = Phi @ G @ self.Qc @ G.T @ Phi.T * dtime
        P = self.state_cov
        P[0:L, 0:L] = Phi @ P[0:L, 0:L] @ Phi.T + Q
        if P.shape[1] > L:
            P[0:L, L:] = Phi @ P[0:L, L:]
            P[L:, 0:L] = P[L:, 0:L] @ Phi.T
        self.state_cov = (P + P.T) / 2.0
        s.time = time

    def predictNewStateLARVIO(self, dt, gyro, acc):
        """:825-897."""
        s = self.imu_state
        gn = float(np.linalg.norm(gyro))
        Om = np.zeros((4, 4))
        Om[0:3, 0:3] = -mu.skew(gyro)
        Om[0:3, 3] = gyro
        Om[3, 0:3] = -gyro
        self.imu_state_old = self._copy_imu(s)
        q = mu.rotation_to_quaternion(s.orientation)
        v = s.velocity
        pp = s.position
        I4 = np.eye(4)
        if gn > 1e-5:
            dq_dt = (math.cos(gn * dt * 0.5) * I4 + 1 / gn * math.sin(gn * dt * 0.5) * Om) @ q
            dq_dt2 = (math.cos(gn * dt * 0.25) * I4 + 1 / gn * math.sin(gn * dt * 0.25) * Om) @ q
        else:
            dq_dt = (I4 + 0.5 * dt * Om) @ q * math.cos(gn * dt * 0.5)
            dq_dt2 = (I4 + 0.25 * dt * Om) @ q * math.cos(gn * dt * 0.25)
        dR_dt = mu.eigen_quat_to_rotation(dq_dt[3], dq_dt[0], dq_dt[1], dq_dt[2])
        dR_dt2 = mu.eigen_quat_to_rotation(dq_dt2[3], dq_dt2[0], dq_dt2[1], dq_dt2[2])
        g = self.gravity
        k1_v_dot = mu.eigen_quat_to_rotation(q[3], q[0], q[1], q[2]) @ acc + g
        k1_p_dot = v
        k1_v = v + k1_v_dot * dt / 2
        k2_v_dot = dR_dt2 @ acc + g
        k2_p_dot = k1_v
        k2_v = v + k2_v_dot * dt / 2
        k3_v_dot = dR_dt2 @ acc + g
        k3_p_dot = k2_v
        k3_v = v + k3_v_dot * dt
        k4_v_dot = dR_dt @ acc + g
        k4_p_dot = k3_v
        q = mu.quaternion_normalize(dq_dt)
        s.velocity = v + dt / 6 * (k1_v_dot + 2 * k2_v_dot + 2 * k3_v_dot + k4_v_dot)
        s.position = pp + dt / 6 * (k1_p_dot + 2 * k2_p_dot + 2 * k3_p_dot + k4_p_dot)
        s.orientation = mu.quaternion_to_rotation(q)

    def predictNewStateOrcVIO(self, dt, gyro, acc):
        """:899-928."""
        s = self.imu_state
        self.imu_state_old = self._copy_imu(s)
        R = s.orientation
        v = s.velocity
        g = self.gravity
        Hl = mu.Hl_operator(dt * gyro)
        s.position = s.position + dt * v + g * (dt ** 2 / 2) + R @ Hl @ acc * dt ** 2
        Jl = mu.Jl_operator(dt * gyro)
        s.velocity = v + g * dt + R @ Jl @ acc * dt
        s.orientation = R @ mu.so3_exp(dt * gyro)

    def calPhiClosedForm(self, dtime, acc, gyro, acc_old, gyro_old):
        """:3980-4370 with Ma = Tg = I, As = 0 (so TA = 0), if_FEJ = false."""
        p = self.p
        L = self.LEG_DIM
        Phi = np.eye(L)
        I3 = np.eye(3)
        C = self.imu_state_old.orientation
        if p.use_larvio_flag or p.use_left_perturbation_flag:
            Axis_Angle = dtime * (gyro_old + gyro) / 2 + dtime * dtime * np.cross(gyro_old, gyro) / 12
            AA = mu.skew(Axis_Angle)
            vk = self.imu_state_old.velocity
            pk = self.imu_state_old.position
            vkp1 = self.imu_state.velocity
            pkp1 = self.imu_state.position
            g = self.gravity
            Phi[0:3, 9:12] = -0.5 * C @ (2 * I3 + AA) * dtime
            Phi[0:3, 12:15] = np.zeros((3, 3))      # 0.5*C*(2I+AA)*dt*TA*Ma with TA = 0
            Phi[3:6, 0:3] = -mu.skew(vkp1 - vk - g * dtime)
            Phi[3:6, 9:12] = (mu.skew(-pkp1 + pk + vkp1 * dtime - 0.5 * g * dtime * dtime) @ C
                              + mu.skew(-0.5 * pkp1 + 0.5 * pk + 0.5 * vkp1 * dtime
                                        - g * dtime * dtime / 6) @ C @ AA)
            Phi[3:6, 12:15] = -0.5 * C @ (2 * I3 + AA) * dtime
            Phi[6:9, 0:3] = -mu.skew(pkp1 - pk - vk * dtime - 0.5 * g * dtime * dtime)
            Phi[6:9, 3:6] = I3 * dtime
            Phi[6:9, 9:12] = (-dtime * dtime * dtime * mu.skew(g) @ C / 6
                              + dtime * mu.skew(pkp1 - pk - g * dtime * dtime / 6) @ C @ AA / 4)
            Phi[6:9, 12:15] = -C @ (3 * I3 + AA) * dtime * dtime / 6
        else:
            wRi = C
            gh = gyro
            ah = acc
            a_skew = mu.skew(ah)
            g_skew = mu.skew(gh)
            gn = float(np.linalg.norm(gh))
            gn2 = gn ** 2
            tt = mu.so3_exp(-dtime * gh)
            JLp = mu.Jl_operator(dtime * gh)
            JLm = mu.Jl_operator(-dtime * gh)
            Delta = -(g_skew / gn2) @ (tt.T @ (dtime * g_skew - I3) + I3)
            HLp = mu.Hl_operator(dtime * gh)
            HLm = mu.Hl_operator(-dtime * gh)
            ga = np.outer(gh, ah)
            adg = float(ah @ gh)
            Phi[0:3, 0:3] = tt
            Phi[0:3, 9:12] = -dtime * JLm
            Phi[3:6, 0:3] = -dtime * wRi @ mu.skew(JLp @ ah)
            Phi[3:6, 9:12] = (wRi @ Delta @ a_skew @ (I3 + (g_skew @ g_skew / gn2))
                              + dtime * wRi @ JLp @ (a_skew @ g_skew / gn2)
                              + dtime * wRi @ (ga / gn2) @ JLm
                              - dtime * (adg / gn2) * I3)
            Phi[3:6, 12:15] = -dtime * wRi @ JLp
            Phi[6:9, 0:3] = -dtime ** 2 * wRi @ mu.skew(HLp @ ah)
            Phi[6:9, 3:6] = dtime * I3
            Phi[6:9, 9:12] = (wRi @ (-g_skew @ Delta - dtime * JLp + dtime * I3) @ a_skew
                              @ (I3 + (g_skew @ g_skew / gn2)) @ (g_skew / gn2)
                              + dtime ** 2 * wRi @ HLp @ (a_skew @ g_skew / gn2)
                              + dtime ** 2 * wRi @ (ga / gn2) @ HLm
                              - dtime ** 2 * (adg / (2 * gn2)) * wRi)
            Phi[6:9, 12:15] = -dtime ** 2 * wRi @ HLp
        return Phi

    def stateAugmentation(self):
        """:930-1013 (no nuisance blocks: use_schmidt = 0)."""
        s = self.imu_state
        self.cur_window_timestamps.append(s.time)
        c = Clone(s.id)
        c.time = s.time
        c.dt = s.dt
        c.orientation = s.orientation.copy()
        c.position = s.position.copy()
        c.R_imu_cam0 = s.R_imu_cam0.copy()
        c.t_cam0_imu = s.t_cam0_imu.copy()
        R_b2w = s.orientation
        R_w2c = s.R_imu_cam0 @ R_b2w.T
        c.orientation_cam = R_w2c.T.copy()
        c.position_cam = s.position + R_b2w @ s.t_cam0_imu
        n_before = len(self.clones)
        self.clones[s.id] = c
        # :963-1010: the new 6 x 6 block goes in front of the EKF-feature block
        self.state_cov = hyb.state_augmentation_cov(self.state_cov, n_before)

    # ------------------------------------------------------------------ bookkeeping
    def addFeatureObservations(self, feats):
        """:1016-1068."""
        p = self.p
        sid = self.imu_state.id
        curr_feature_num = len(self.map_server)
        tracked = 0
        dt = self.imu_state.dt
        for f in feats:
            fid, u, v, u_init, v_init, u_vel, v_vel, u_init_vel, v_init_vel = f
            if fid not in self.map_server:
                ft = Feature(fid)
                self.map_server[fid] = ft
                ft.observations[sid] = np.array([u + u_vel * dt, v + v_vel * dt])
                ft.observations_vel[sid] = np.array([u_vel, v_vel])
                ft.totalObsNum += 1
                if not (u_init == -1 and v_init == -1) and (sid - 1) in self.clones:
                    dt_ = self.clones[sid - 1].dt
                    ft.observations[sid - 1] = np.array([u_init + u_init_vel * dt_, v_init + v_init_vel * dt_])
                    ft.observations_vel[sid - 1] = np.array([u_init_vel, v_init_vel])
                    ft.totalObsNum += 1
            else:
                ft = self.map_server[fid]
                ft.observations[sid] = np.array([u + u_vel * dt, v + v_vel * dt])
                ft.observations_vel[sid] = np.array([u_vel, v_vel])
                ft.totalObsNum += 1
                tracked += 1
                if p.if_ZUPT_valid and p.if_use_feature_zupt_flag and (sid - 1) in ft.observations:
                    d = np.array([u, v]) - ft.observations[sid - 1]
                    self.coarse_feature_dis.append(float(np.linalg.norm(d)))
        # :1063-1065 (0/0 -> nan on the first frame, like the reference)
        self.tracking_rate = (tracked / curr_feature_num) if curr_feature_num else float("nan")

    def _cam_poses(self, ft, exclude_id=None):
        Rs, ts, zs, ids = [], [], [], []
        for sid in ft.obs_ids():
            if sid not in self.clones:
                continue
            if exclude_id is not None and sid == exclude_id:
                continue
            c = self.clones[sid]
            Rs.append([float(x) for x in c.orientation_cam.ravel()])
            ts.append([float(x) for x in c.position_cam])
            zs.append((float(ft.observations[sid][0]), float(ft.observations[sid][1])))
            ids.append(sid)
        return Rs, ts, zs, ids

    def checkMotion(self, ft, if_tracked):
        """Feature::checkMotion, feature.hpp:353-396."""
        ids = ft.obs_ids()
        first = ids[0]
        last = ids[-2] if if_tracked else ids[-1]
        c0 = self.clones[first]
        c1 = self.clones[last]
        z0 = ft.observations[first]
        return feat.check_motion([float(x) for x in c0.orientation_cam.ravel()],
                                 [float(x) for x in c0.position_cam],
                                 [float(x) for x in c1.position_cam],
                                 (float(z0[0]), float(z0[1])), self.opt.translation_threshold)

    def _initialize(self, ft, exclude_id):
        """Feature::initializePosition (:398-449, exclude curr_id) and
        initializePosition_AssignAnchor (:451-500, exclude nothing)."""
        Rs, ts, zs, ids = self._cam_poses(ft, exclude_id)
        res = feat.triangulate(Rs, ts, zs, ft.is_initialized, [float(x) for x in ft.position], self.opt)
        ft.tri_log = res
        if res.valid:
            ft.is_initialized = True
            ft.position = np.array(res.position)
            ft.id_anchor = ids[-1]
        return res.valid

    def _initialize_inv_param(self, ft, exclude_id):
        """Feature::initializeInvParamPosition, feature.hpp:502-554: the same triangulation as
        initializePosition (observations of `exclude_id` skipped), keeping the 1-D inverse-depth
        parametrisation in the anchor camera = the last camera used."""
        Rs, ts, zs, ids = self._cam_poses(ft, exclude_id)
        res = feat.triangulate(Rs, ts, zs, ft.is_initialized, [float(x) for x in ft.position], self.opt)
        ft.tri_log = res
        if res.valid:
            ft.ekf_feature = True
            ft.is_initialized = True
            ft.position = np.array(res.position)
            ft.id_anchor = ids[-1]
            fp = res.final_position
            ft.invDepth = 1 / fp[2]
            ft.obs_anchor = np.array([fp[0] * ft.invDepth, fp[1] * ft.invDepth, 1.0])
        return res.valid

    def _grid_code(self, ft):
        """:2286-2289 / :3841-3845 (int casts truncate toward zero)."""
        p = self.p
        xy = ft.observations[self.imu_state.id]
        row = int((xy[1] - p.y_min) / p.grid_height)
        col = int((xy[0] - p.x_min) / p.grid_width)
        return row * p.grid_cols + col

    def updateGridMap(self):
        """:3831-3850."""
        p = self.p
        if p.grid_rows * p.grid_cols == 0:
            return
        self.grid_map = {i: [] for i in range(p.grid_rows * p.grid_cols)}
        for fid in self.feature_states:
            self.grid_map.setdefault(self._grid_code(self.map_server[fid]), []).append(fid)

    def rmLostFeaturesCov(self, lost_ids):
        """:3776-3828 (use_schmidt = 0)."""
        for fid in lost_ids:
            seq = self.feature_states.index(fid)
            a = self.LEG_DIM + 6 * len(self.clones) + seq
            keep = [i for i in range(self.state_cov.shape[0]) if i != a]
            self.state_cov = self.state_cov[np.ix_(keep, keep)]
            self.feature_states.pop(seq)
            del self.map_server[fid]

    def _drop_all_feature_states(self):
        """checkZUPTFeat :3104-3115 / checkZUPTIMU :3304-3315: a stationary frame removes every EKF feature."""
        if self.feature_states:
            E = len(self.feature_states)
            n = self.state_cov.shape[0] - E
            self.state_cov = self.state_cov[:n, :n].copy()
            for fid in self.feature_states:
                ft = self.map_server[fid]
                ft.is_initialized = False
                ft.ekf_feature = False
                ft.in_state = False
            self.feature_states = []

    def _update_feature_states(self, delta_x):
        """The 1-D inverse-depth branch of :1700-1737 / :1843-1882 / :3391-3430: invDepth += its component of
        delta_x, world position recomputed from the (already updated) anchor clone."""
        base = self.LEG_DIM + 6 * len(self.clones)
        for i, fid in enumerate(self.feature_states):
            ft = self.map_server[fid]
            c = self.clones[ft.id_anchor]
            ft.invDepth = ft.invDepth + delta_x[base + i]
            p_c = np.array([ft.obs_anchor[0] / ft.invDepth, ft.obs_anchor[1] / ft.invDepth, 1 / ft.invDepth])
            ft.position = c.orientation_cam @ p_c + c.position_cam

    def _ekf_clone_arrays(self):
        ids = sorted(self.clones.keys())
        return ids, [self.clones[i].orientation for i in ids], [self.clones[i].position for i in ids]

    def featureJacobian_ekf(self, ft):
        """:1575-1651 (1-D inverse depth, anchor in the window)."""
        ids, cR, cp = self._ekf_clone_arrays()
        s = self.imu_state
        cur = s.id
        H, r = hyb.feature_jacobian_ekf(cR, cp, s.R_imu_cam0, s.t_cam0_imu, ids.index(cur), ids.index(ft.id_anchor),
                                        self.feature_states.index(ft.id), self.state_cov.shape[1] - self.LEG_DIM - 6 * len(ids),
                                        ft.obs_anchor, ft.invDepth, ft.position, ft.observations[cur])
        return H, r

    def featureJacobian_ekf_new(self, ft, n_cols):
        """:1481-1572: rows of every observing clone but the anchor; H_f at LEG + 6N + index in feature_states."""
        ids, cR, cp = self._ekf_clone_arrays()
        s = self.imu_state
        sids = [i for i in ft.obs_ids() if i in self.clones]
        col = self.LEG_DIM + 6 * len(ids) + self.feature_states.index(ft.id)
        return hyb.feature_jacobian_ekf_new(cR, cp, s.R_imu_cam0, s.t_cam0_imu, [ids.index(i) for i in sids],
                                            [ft.observations[i] for i in sids], ids.index(ft.id_anchor), col, n_cols,
                                            ft.obs_anchor, ft.invDepth, ft.position)

    # ------------------------------------------------------------------ stage 2
    def measurementJacobian_msckf(self, state_id, ft):
        """:1071-1168 -> H_x (2x6), H_e (2x6), H_f (2x3), r (2)."""
        p = self.p
        c = self.clones[state_id]
        R_b2c = c.R_imu_cam0
        t_c_b = c.t_cam0_imu
        R_b2w = c.orientation
        R_w2b = R_b2w.T
        t_b_w = c.position
        R_w2c = R_b2c @ R_w2b
        t_c_w = t_b_w + R_b2w @ t_c_b
        p_w = ft.position
        z = ft.observations[state_id]
        p_c = R_w2c @ (p_w - t_c_w)
        p_bf_w = p_w - t_b_w
        dz = np.zeros((2, 3))
        dz[0, 0] = 1 / p_c[2]
        dz[1, 1] = 1 / p_c[2]
        dz[0, 2] = -p_c[0] / (p_c[2] * p_c[2])
        dz[1, 2] = -p_c[1] / (p_c[2] * p_c[2])
        dpc_dxb = np.zeros((3, 6))
        if not p.use_larvio_flag:
            temp = np.zeros((3, 4))
            temp[:, :3] = np.eye(3)
            wTc = np.eye(4)
            wTc[:3, :3] = R_w2c.T
            wTc[:3, 3] = t_c_w
            ul = np.array([p_w[0], p_w[1], p_w[2], 1.0])
            dcd = mu.cam_wrt_imu_se3_jacobian(R_b2c, t_c_b, R_w2c, t_b_w, p.use_left_perturbation_flag)
            cTw = np.linalg.inv(wTc)
            if p.use_left_perturbation_flag:
                dpc_dxb = temp @ cTw @ mu.odot(ul) @ dcd
            else:
                dpc_dxb = temp @ mu.odot(cTw @ ul) @ dcd
            H_x = -dz @ dpc_dxb
        else:
            dpc_dxb[:, :3] = R_w2c @ mu.skew(p_bf_w)
            dpc_dxb[:, 3:] = -R_w2c
            H_x = dz @ dpc_dxb
        dpc_dxe = np.zeros((3, 6))
        dpc_dxe[:, :3] = (R_w2c @ mu.skew(p_bf_w) @ R_b2w) - (R_b2c @ mu.skew(t_c_b))
        dpc_dxe[:, 3:] = -R_b2c
        H_e = dz @ dpc_dxe
        H_f = dz @ R_w2c
        r = z - np.array([p_c[0] / p_c[2], p_c[1] / p_c[2]])
        return H_x, H_e, H_f, r

    def featureJacobian_msckf(self, ft, state_ids, project=True):
        """:1171-1226 -> (H_x (2m-3 x D), r).  With project=False returns the
        unprojected (H_xj, H_fj, r_j) for invariant-based parity tests."""
        L = self.LEG_DIM
        valid = [s for s in state_ids if s in ft.observations]
        rows = 2 * len(valid)
        D = self.state_cov.shape[1]
        H_xj = np.zeros((rows, D))
        H_fj = np.zeros((rows, 3))
        r_j = np.zeros(rows)
        order = sorted(self.clones.keys())
        k = 0
        for sid in valid:
            H_xi, H_ei, H_fi, r_i = self.measurementJacobian_msckf(sid, ft)
            cntr = order.index(sid)
            H_xj[k:k + 2, L + 6 * cntr:L + 6 * cntr + 6] = H_xi
            H_xj[k:k + 2, 15:21] = H_ei
            if self.p.estimate_td:
                H_xj[k:k + 2, 21] = ft.observations_vel[sid]
            H_fj[k:k + 2, :] = H_fi
            r_j[k:k + 2] = r_i
            k += 2
        if not project:
            return H_xj, H_fj, r_j
        ok, H_x, r = nullspace_project_inplace_svd(H_fj, H_xj, r_j)
        return H_x, r

    def gatingTestFeature(self, H, r, dof, log=None):
        """:1953-1976."""
        P1 = H @ self.state_cov @ H.T
        P2 = self.p.feature_observation_noise * np.eye(H.shape[0])
        gamma = float(r @ np.linalg.solve(P1 + P2, r))
        if dof < 500:
            chi = self.chi_table_feat[dof]
        else:
            from scipy.stats import chi2
            chi = float(chi2.ppf(self.p.chi_square_threshold_feat, dof))
        if log is not None:
            log["gamma"] = gamma
            log["chi2"] = chi
        return gamma < chi

    # ------------------------------------------------------------------ stages 4+5
    def _compress(self, H, r, keep_rows):
        """:1664-1683 / :2532-2552.  SPQR (SuiteSparse, third party, absent) is a
        Householder QR; any orthogonal Q gives the same posterior, so LAPACK's QR
        restates it.  Returns (H_thin, r_thin)."""
        if H.shape[0] > H.shape[1]:
            Q, R = np.linalg.qr(H, mode="reduced")
            return R[:keep_rows, :], (Q.T @ r)[:keep_rows]
        return H, r

    def measurementUpdate_msckf(self, H, r):
        """:1654-1763 (use_schmidt = 0)."""
        if H.shape[0] == 0 or r.shape[0] == 0:
            return
        keep = self.LEG_DIM + 6 * len(self.clones)
        H_thin, r_thin = self._compress(H, r, keep)
        P = self.state_cov
        S = H_thin @ P @ H_thin.T + self.p.feature_observation_noise * np.eye(H_thin.shape[0])
        K_T = np.linalg.solve(S, H_thin @ P)
        K = K_T.T
        delta_x = K @ r_thin
        applied = self.incrementState_IMUCam(delta_x)
        self._update_feature_states(delta_x)
        I_KH = np.eye(K.shape[0], H_thin.shape[1]) - K @ H_thin
        P = I_KH @ P
        self.state_cov = (P + P.T) / 2.0
        self.last_update_time = self.imu_state.time
        self.log.append(dict(kind="update", rows=int(H.shape[0]), delta_x=delta_x.copy(), applied=applied))

    def measurementUpdate_hybrid(self, H_ekf_new, r_ekf_new, H_ekf, r_ekf, H_msckf, r_msckf):
        """:1766-1950 (feature_idp_dim = 1, use_schmidt = 0).  H_o = [H_msckf; H_ekf; top rows of H_ekf_new] is
        NOT compressed again here; the last sz_new rows of the sparsified H_ekf_new are (H_1 | H_2), r_1."""
        D = self.state_cov.shape[1]
        sz_new = H_ekf_new.shape[1] - D
        sz_r = r_ekf_new.shape[0] + r_ekf.shape[0] + r_msckf.shape[0]
        if sz_r == 0:
            return
        n_top = H_ekf_new.shape[0] - sz_new
        H_o = np.concatenate([H_msckf, H_ekf, H_ekf_new[:n_top, :D]], axis=0)
        r_o = np.concatenate([r_msckf, r_ekf, r_ekf_new[:n_top]])
        H_1 = H_ekf_new[n_top:, :D]
        H_2 = H_ekf_new[n_top:, D:]
        r_1 = r_ekf_new[n_top:]
        P = self.state_cov
        S = H_o @ P @ H_o.T + self.p.feature_observation_noise * np.eye(H_o.shape[0])
        K_T = np.linalg.solve(S, H_o @ P)
        K = K_T.T
        dx_leg = K @ r_o
        if sz_new > 0:
            HH = np.linalg.solve(H_2, H_1)
            dx_new = -HH @ dx_leg + np.linalg.solve(H_2, r_1)
            delta_x = np.concatenate([dx_leg, dx_new])
        else:
            delta_x = dx_leg
        applied = self.incrementState_IMUCam(delta_x)
        self._update_feature_states(delta_x)
        I_KH = np.eye(K.shape[0], H_o.shape[1]) - K @ H_o
        P = I_KH @ P
        P = (P + P.T) / 2.0
        if sz_new > 0:
            nHHP = -HH @ P
            P22 = -nHHP @ HH.T + self.p.feature_observation_noise * np.linalg.inv(H_2.T @ H_2)
            Pn = np.zeros((D + sz_new, D + sz_new))
            Pn[:D, :D] = P
            Pn[D:, :D] = nHHP
            Pn[:D, D:] = nHHP.T
            Pn[D:, D:] = P22
            P = (Pn + Pn.T) / 2.0
        self.state_cov = P
        self.last_update_time = self.imu_state.time
        self.log.append(dict(kind="update", rows=int(H_o.shape[0]), delta_x=delta_x.copy(), applied=applied,
                             sz_new=int(sz_new)))

    def incrementState_IMUCam(self, delta_x):
        """:4468-4567.  Returns False when the large-update guard discarded the step."""
        p = self.p
        L = self.LEG_DIM
        d = delta_x[:L]
        if (np.linalg.norm(d[3:6]) > 1.0 or np.linalg.norm(d[6:9]) > 1.5) and p.discard_large_update_flag:
            return False
        s = self.imu_state
        Rt = mu.so3_exp(d[0:3])
        left = p.use_larvio_flag or p.use_left_perturbation_flag
        s.orientation = Rt @ s.orientation if left else s.orientation @ Rt
        s.velocity = s.velocity + d[3:6]
        s.position = s.position + d[6:9]
        s.gyro_bias = s.gyro_bias + d[9:12]
        s.acc_bias = s.acc_bias + d[12:15]
        dq = mu.small_angle_quaternion(d[15:18])
        s.R_imu_cam0 = s.R_imu_cam0 @ mu.eigen_quat_to_rotation(dq[3], dq[0], dq[1], dq[2]).T
        s.t_cam0_imu = s.t_cam0_imu + d[18:21]
        self.td += d[21]
        for i, sid in enumerate(sorted(self.clones.keys())):
            c = self.clones[sid]
            da = delta_x[L + 6 * i:L + 6 * i + 6]
            Rt = mu.so3_exp(da[0:3])
            c.orientation = Rt @ c.orientation if left else c.orientation @ Rt
            c.position = c.position + da[3:6]
            R_b2w = c.orientation
            c.orientation_cam = R_b2w @ s.R_imu_cam0.T
            c.position_cam = c.position + R_b2w @ s.t_cam0_imu
        return True

    # ------------------------------------------------------------------ orchestrators
    def removeLostFeatures(self):
        """:2196-2579 (pure-MSCKF and hybrid branches; use_schmidt = 0, feature_idp_dim = 1)."""
        p = self.p
        cur = self.imu_state.id
        rows = 0
        rows_new = 0
        invalid, msckf_ids, lost_ids = [], [], []
        ekf_new_ids, ekf_lost_ids, ekf_ids = [], [], []
        # features of the state: tracked now -> a 2-row update, lost -> dropped from the state (:2210-2232)
        for fid in sorted(self.map_server.keys()):
            ft = self.map_server[fid]
            if ft.in_state:
                if cur in ft.observations:
                    ekf_ids.append(fid)
                else:
                    ekf_lost_ids.append(fid)
        self.rmLostFeaturesCov(ekf_lost_ids)
        self.updateGridMap()
        cap = p.max_features * p.grid_rows * p.grid_cols
        for fid in sorted(self.map_server.keys()):
            ft = self.map_server[fid]
            if ft.in_state:
                continue
            tracked_now = cur in ft.observations
            if not tracked_now:
                if len(ft.observations) < p.least_Obs_Num:
                    invalid.append(fid)
                    continue
                if not ft.is_initialized:
                    if not self.checkMotion(ft, tracked_now):
                        invalid.append(fid)
                        continue
                    if not self._initialize(ft, cur):
                        invalid.append(fid)
                        continue
                rows += 2 * len(ft.observations) - 3
                msckf_ids.append(fid)
                lost_ids.append(fid)
            else:
                if not (len(ft.observations) >= p.max_track_len):
                    continue
                # :2283-2323: EKF-SLAM feature if its grid cell (and the state) has room, else MSCKF feature
                code = self._grid_code(ft)
                cell = self.grid_map.setdefault(code, [])
                if (len(cell) < p.max_features and self.imu_state.time - self.last_ZUPT_time > 5
                        and (len(self.feature_states) + len(ekf_new_ids)) < cap):
                    if not ft.ekf_feature:
                        ft.is_initialized = False
                        if self.checkMotion(ft, tracked_now):
                            self._initialize_inv_param(ft, cur)
                    if not ft.is_initialized:
                        continue
                    rows_new += 2 * (len(ft.observations) - 1)
                    ekf_new_ids.append(fid)
                    cell.append(fid)
                else:
                    if not ft.is_initialized:
                        if self.checkMotion(ft, tracked_now):
                            self._initialize(ft, cur)
                    if not ft.is_initialized:
                        continue
                    rows += 2 * len(ft.observations) - 3
                    msckf_ids.append(fid)
                    lost_ids.append(fid)
        for fid in invalid:
            del self.map_server[fid]
        frame_log = dict(kind="removeLostFeatures", state_id=cur, invalid=list(invalid),
                         candidates=list(msckf_ids), gate={}, zupt=self.if_ZUPT,
                         ekf_lost=list(ekf_lost_ids), ekf=list(ekf_ids), ekf_new=list(ekf_new_ids),
                         gate_ekf={}, gate_ekf_new={})
        self.log.append(frame_log)
        if not msckf_ids and not ekf_new_ids and not ekf_ids:
            return
        if not self.if_ZUPT:
            L = self.LEG_DIM
            D = self.state_cov.shape[1]
            for fid in ekf_new_ids:
                self.map_server[fid].in_state = True
                self.feature_states.append(fid)
            # ---- new EKF-SLAM features (:2343-2446)
            blocks, invalid_new = [], []
            for fid in list(ekf_new_ids):
                ft = self.map_server[fid]
                sids = ft.obs_ids()
                H_m, r_m = self.featureJacobian_msckf(ft, sids)        # gate with the MSCKF rows (:2365-2370)
                g = {}
                ok = self.gatingTestFeature(H_m, r_m, 2 * len(sids) - 3, g)
                g["pass"] = ok
                g["position"] = ft.position.copy()
                g["inv_depth"] = ft.invDepth
                frame_log["gate_ekf_new"][fid] = g
                if not ok:
                    invalid_new.append(fid)
            for fid in invalid_new:          # :2384-2411
                self.map_server[fid].in_state = False
                self.feature_states.remove(fid)
                ekf_new_ids.remove(fid)
            n_new = len(ekf_new_ids)
            H_ekf_new = np.zeros((0, D + n_new))
            r_ekf_new = np.zeros(0)
            if n_new > 0:
                # the columns of the surviving new features are contiguous behind the old state (the reference
                # builds the rows first and deletes the columns of the rejected ones afterwards: same matrix)
                for fid in ekf_new_ids:
                    H_j, r_j = self.featureJacobian_ekf_new(self.map_server[fid], D + n_new)
                    blocks.append((H_j, r_j))
                H_ekf_new = np.concatenate([b[0] for b in blocks], axis=0)
                r_ekf_new = np.concatenate([b[1] for b in blocks])
                H_ekf_new, r_ekf_new = hyb.sparsify_new_features(H_ekf_new, r_ekf_new, n_new)
            # ---- features of the state (:2449-2495)
            He, re_ = [], []
            for fid in ekf_ids:
                ft = self.map_server[fid]
                H_j, r_j = self.featureJacobian_ekf(ft)
                g = {}
                ok = self.gatingTestFeature(H_j, r_j, 2, g)
                g["pass"] = ok
                frame_log["gate_ekf"][fid] = g
                if ok:
                    He.append(H_j)
                    re_.append(r_j)
            H_ekf = np.concatenate(He, axis=0) if He else np.zeros((0, D))
            r_ekf = np.concatenate(re_) if re_ else np.zeros(0)
            if not (H_ekf.shape[0] == 0 or H_ekf.shape[0] <= H_ekf.shape[1]):
                H_ekf, r_ekf = self._compress(H_ekf, r_ekf, D)
            # ---- MSCKF features (:2498-2560)
            cols = L + 6 * len(self.clones)
            H = np.zeros((rows, cols))
            r = np.zeros(rows)
            k = 0
            for fid in msckf_ids:
                ft = self.map_server[fid]
                sids = ft.obs_ids()
                H_xj, r_j = self.featureJacobian_msckf(ft, sids)
                g = {}
                ok = self.gatingTestFeature(H_xj, r_j, 2 * len(sids) - 3, g)
                g["pass"] = ok
                g["position"] = ft.position.copy()
                frame_log["gate"][fid] = g
                if ok:
                    H[k:k + H_xj.shape[0], :] = H_xj[:, :cols]
                    r[k:k + H_xj.shape[0]] = r_j
                    k += H_xj.shape[0]
            H = H[:k]
            r = r[:k]
            if not (H.shape[0] == 0 or H.shape[0] <= H.shape[1]):
                H, r = self._compress(H, r, cols)
            Hfull = np.zeros((H.shape[0], D))
            Hfull[:, :H.shape[1]] = H
            self.measurementUpdate_hybrid(H_ekf_new, r_ekf_new, H_ekf, r_ekf, Hfull, r)
        else:
            for fid in msckf_ids:
                self.map_server[fid].is_initialized = False
        for fid in lost_ids:
            del self.map_server[fid]

    def findRedundantImuStates(self):
        """:2582-2626 (literal iterator arithmetic, including the two decrements)."""
        p = self.p
        ids = sorted(self.clones.keys())
        key_i = len(ids) - 4
        st_i = key_i + 1
        first_i = 0
        key = self.clones[ids[key_i]]
        rm = []
        for _ in range(2):
            c = self.clones[ids[st_i]]
            distance = float(np.linalg.norm(c.position_cam - key.position_cam))
            angle = mu.angle_axis_angle(c.orientation_cam.T @ key.orientation_cam)
            if (angle < p.rotation_threshold and distance < p.translation_threshold
                    and self.tracking_rate > p.tracking_rate_threshold):
                rm.append(ids[st_i])
                st_i += 1
            else:
                rm.append(ids[first_i])
                first_i += 1
                st_i -= 2
        return sorted(rm)

    def pruneImuStateBuffer(self):
        """:2629-2959 (use_schmidt = 0, feature_idp_dim = 1)."""
        p = self.p
        cur = self.imu_state.id
        if not self.if_ZUPT:
            if len(self.clones) < p.sw_size:
                return
            rm_ids = self.findRedundantImuStates()
        else:
            rm_ids = [cur - 1]
        rows = 0
        used = []
        reanchored = {}
        for fid in sorted(self.map_server.keys()):
            ft = self.map_server[fid]
            involved = [s for s in rm_ids if s in ft.observations]
            if not involved:
                continue
            if ft.in_state:
                # :2665-2722: a feature of the state whose anchor leaves the window moves to a new anchor
                if ft.id_anchor in involved:
                    new_id = self.getNewAnchorId(ft, involved)
                    c = self.clones[new_id]
                    p_new = np.linalg.solve(c.orientation_cam, ft.position - c.position_cam)
                    ft.invDepth = 1 / p_new[2]
                    ft.obs_anchor = np.array([p_new[0] / p_new[2], p_new[1] / p_new[2], ft.obs_anchor[2]])
                    self.updateFeatureCov_1didp(ft, ft.id_anchor, new_id)
                    reanchored[fid] = (ft.id_anchor, new_id)
                    ft.id_anchor = new_id
                continue
            if ft.is_initialized and ft.id_anchor in involved:
                # :2724-2773: initialised feature outside the state: anchor moved, observation NOT corrected
                new_id = self.getNewAnchorId(ft, involved)
                c = self.clones[new_id]
                p_new = np.linalg.solve(c.orientation_cam, ft.position - c.position_cam)
                ft.invDepth = 1 / p_new[2]
                if new_id not in ft.observations:
                    ft.observations[new_id] = np.zeros(2)       # std::map::operator[] default-constructs (:2763)
                ft.obs_anchor = np.array([ft.observations[new_id][0], ft.observations[new_id][1], ft.obs_anchor[2]])
                ft.id_anchor = new_id
            if (not self.if_ZUPT) and (not ft.ekf_feature) and len(involved) > 1:
                tracked = cur in ft.observations
                if not ft.is_initialized:
                    if not self.checkMotion(ft, tracked):
                        continue
                    if not self._initialize(ft, None):
                        continue
                used.append(fid)
                rows += 2 * len(involved) - 3
        frame_log = dict(kind="prune", state_id=cur, rm_ids=list(rm_ids), candidates=list(used), gate={},
                         reanchored=reanchored)
        self.log.append(frame_log)
        if (not self.if_ZUPT) and used:
            D = self.state_cov.shape[1]
            H = np.zeros((rows, D))
            r = np.zeros(rows)
            k = 0
            used_set = set(used)
            for fid in sorted(self.map_server.keys()):
                ft = self.map_server[fid]
                involved = [s for s in rm_ids if s in ft.observations]
                if fid in used_set:
                    H_xj, r_j = self.featureJacobian_msckf(ft, involved)
                    g = {}
                    ok = self.gatingTestFeature(H_xj, r_j, 2 * len(involved) - 3, g)
                    g["pass"] = ok
                    g["position"] = ft.position.copy()
                    frame_log["gate"][fid] = g
                    if ok:
                        H[k:k + H_xj.shape[0], :] = H_xj
                        r[k:k + H_xj.shape[0]] = r_j
                        k += H_xj.shape[0]
                for s in involved:
                    del ft.observations[s]
            self.measurementUpdate_msckf(H[:k], r[:k])
        else:
            for ft in self.map_server.values():
                for s in [s for s in rm_ids if s in ft.observations]:
                    del ft.observations[s]
        L = self.LEG_DIM
        for sid in rm_ids:
            seq = sorted(self.clones.keys()).index(sid)
            a = L + 6 * seq
            keep = [i for i in range(self.state_cov.shape[0]) if not (a <= i < a + 6)]
            self.state_cov = self.state_cov[np.ix_(keep, keep)]
            t_erase = self.clones[sid].time
            self.cur_window_timestamps = [t for t in self.cur_window_timestamps if t != t_erase]
            del self.clones[sid]

    def getNewAnchorId(self, ft, rm_ids):
        """:3892-3950: among the clones observing the feature (the two newest excluded, the removed ones
        excluded) the one whose stored observation is closest to the reprojection of feature.position;
        the newest clone if there is none."""
        ids = sorted(self.clones.keys())
        size = len(ids)
        if size <= 2:
            return ids[-1]
        best, min_dis = None, 99999.0
        for sid in ids[:size - 2]:
            if sid not in ft.observations or sid in rm_ids:
                continue
            c = self.clones[sid]
            p_new = np.linalg.solve(c.orientation_cam, ft.position - c.position_cam)
            z = ft.observations[sid]
            dis = math.sqrt((p_new[0] / p_new[2] - z[0]) ** 2 + (p_new[1] / p_new[2] - z[1]) ** 2)
            if min_dis > dis:
                min_dis = dis
                best = sid
        return best if best is not None else ids[-1]

    def updateFeatureCov_1didp(self, ft, old_id, new_id):
        """:3611-3773; ft.invDepth already holds the inverse depth in the new anchor."""
        ids, cR, cp = self._ekf_clone_arrays()
        s = self.imu_state
        self.state_cov, _ = hyb.update_feature_cov_1didp(self.state_cov, len(ids), self.feature_states.index(ft.id),
                                                         ids.index(old_id), ids.index(new_id), cR, cp, s.R_imu_cam0,
                                                         s.t_cam0_imu, ft.position, ft.invDepth)

    # ------------------------------------------------------------------ ZUPT (Z1)
    def checkZUPTFeat(self):
        """:3081-3125."""
        if len(self.coarse_feature_dis) < 20:
            self.coarse_feature_dis = []
            return False
        d = sorted(self.coarse_feature_dis)
        maxDis = d[-9]
        self.coarse_feature_dis = []
        if maxDis < self.p.zupt_max_feature_dis:
            self._drop_all_feature_states()
            self.measurementUpdate_ZUPT_vpq()
            return True
        return False

    def checkZUPTIMU(self):
        """:3129-3323."""
        self.zupt_info = None
        if len(self.imu_recent_zupt) < 2:
            return False
        if (self.imu_state.id - 1) not in self.clones:
            # the reference would default-construct a clone through std::map::operator[] here (:3345-3346)
            # and corrupt its window; a one-clone window never takes a ZUPT in this restatement
            return False
        sigma_w_2 = 1.6968e-04 ** 2
        sigma_a_2 = 2.0000e-3 ** 2
        sigma_wb = 1.9393e-05
        sigma_ab = 3.0000e-03
        n = len(self.imu_recent_zupt) - 1
        H = np.zeros((6 * n, 9))
        res = np.zeros(6 * n)
        Rm = np.eye(6 * n)
        wRi = self.imu_state.orientation
        dt_summed = 0.0
        for i in range(n):
            dt = self.imu_recent_zupt[i + 1][0] - self.imu_recent_zupt[i][0]
            acc = np.array(self.imu_recent_zupt[i][2], dtype=float) - self.imu_state.acc_bias
            res[6 * i:6 * i + 3] = 0.0
            res[6 * i + 3:6 * i + 6] = -wRi @ acc - self.gravity
            H[6 * i:6 * i + 3, 3:6] = np.eye(3)
            if self.p.use_left_perturbation_flag:
                H[6 * i + 3:6 * i + 6, 0:3] = mu.skew(wRi @ acc)
            else:
                H[6 * i + 3:6 * i + 6, 0:3] = wRi @ mu.skew(acc)
            H[6 * i + 3:6 * i + 6, 6:9] = wRi
            Rm[6 * i:6 * i + 3, 6 * i:6 * i + 3] *= sigma_w_2 / dt
            Rm[6 * i + 3:6 * i + 6, 6 * i + 3:6 * i + 6] *= sigma_a_2 / dt
            dt_summed += dt
        Q_bias = np.eye(6)
        Q_bias[0:3, 0:3] *= dt_summed * sigma_wb
        Q_bias[3:6, 3:6] *= dt_summed * sigma_ab
        P = self.state_cov
        idx = [0, 1, 2, 9, 10, 11, 12, 13, 14]
        P_marg = P[np.ix_(idx, idx)].copy()
        P_marg[3:9, 3:9] += Q_bias
        S = H @ P_marg @ H.T + Rm
        chi2 = float(res @ np.linalg.solve(S, res))
        dof = res.shape[0]
        if dof < 500:
            chk = self.chi_table_zupt[dof]
        else:
            from scipy.stats import chi2 as c2
            chk = float(c2.ppf(0.95, dof))
        self.zupt_info = (chi2, float(np.linalg.norm(self.imu_state.velocity)))
        if chi2 > chk or np.linalg.norm(self.imu_state.velocity) > 0.25:
            return False
        self._drop_all_feature_states()
        self.measurementUpdate_ZUPT_vpq()
        return True

    def measurementUpdate_ZUPT_vpq(self):
        """:3326-3454."""
        p = self.p
        L = self.LEG_DIM
        N = len(self.clones)
        D = self.state_cov.shape[1]
        H = np.zeros((9, D))
        H[0:3, 3:6] = np.eye(3)
        H[3:6, L + 6 * N - 3:L + 6 * N] = np.eye(3)
        H[3:6, L + 6 * N - 9:L + 6 * N - 6] = -np.eye(3)
        H[6:9, L + 6 * N - 6:L + 6 * N - 3] = -0.5 * np.eye(3)
        H[6:9, L + 6 * N - 12:L + 6 * N - 9] = 0.5 * np.eye(3)
        r = np.zeros(9)
        r[0:3] = -self.imu_state.velocity
        cur = self.imu_state.id
        r[3:6] = -(self.clones[cur].position - self.clones[cur - 1].position)
        qc = mu.rotation_to_quaternion(self.clones[cur].orientation)
        qp = mu.rotation_to_quaternion(self.clones[cur - 1].orientation)
        # Eigen quaternion product q_curr * conj(q_prev), (w,x,y,z)
        aw, ax, ay, az = qc[3], qc[0], qc[1], qc[2]
        bw, bx, by, bz = qp[3], -qp[0], -qp[1], -qp[2]
        r[6] = aw * bx + ax * bw + ay * bz - az * by
        r[7] = aw * by + ay * bw + az * bx - ax * bz
        r[8] = aw * bz + az * bw + ax * by - ay * bx
        Rz = np.zeros((9, 9))
        Rz[0:3, 0:3] = p.zupt_noise_v * np.eye(3)
        Rz[3:6, 3:6] = p.zupt_noise_p * np.eye(3)
        Rz[6:9, 6:9] = p.zupt_noise_q * np.eye(3)
        P = self.state_cov
        S = H @ P @ H.T + Rz
        K = np.linalg.solve(S, H @ P).T
        delta_x = K @ r
        applied = self.incrementState_IMUCam(delta_x)
        P = (np.eye(D) - K @ H) @ P
        self.state_cov = (P + P.T) / 2.0
        self.last_update_time = self.imu_state.time
        self.last_ZUPT_time = self.imu_state.time
        self.log.append(dict(kind="zupt", delta_x=delta_x.copy(), applied=applied))

    # ------------------------------------------------------------------ stage 3 (filter side)
    def constructObjectResidualJacobians(self, jac_sensor, object_timestamps, Hf, res,
                                         zs_num_wrt_timestamps, valid_camera_pose_mat):
        """:2017-2151.  Returns (flag, Hx, Hf, res)."""
        p = self.p
        if not p.use_object_residual_update_cam_pose_flag:
            return False, None, Hf, res
        total = len(object_timestamps)
        D = self.state_cov.shape[1]
        odim = Hf.shape[1]
        sum_zs = sum(2 * n for n in zs_num_wrt_timestamps)
        Hx = np.zeros((jac_sensor.shape[0], D))
        Hf_t = Hf.copy()
        res_t = res.copy()
        row = 0
        frow = 0
        flag = False
        L = self.LEG_DIM
        for k in range(total):
            ts = object_timestamps[k]
            nz = 2 * zs_num_wrt_timestamps[k]
            if ts in self.cur_window_timestamps:
                if not self.dcampose_dimupose_fixed_flag:
                    wTc = mu.se3_exp(valid_camera_pose_mat[:, k])
                    R_b2c = self.imu_state.R_imu_cam0
                    t_c_b = self.imu_state.t_cam0_imu
                    R_w2c = np.linalg.inv(wTc)[:3, :3]
                    t_b_w = wTc[:3, :3] @ (-R_b2c @ t_c_b) + wTc[:3, 3]
                    self.dcampose_dimupose_mat = mu.cam_wrt_imu_se3_jacobian(
                        R_b2c, t_c_b, R_w2c, t_b_w, p.use_left_perturbation_flag)
                pi = self.cur_window_timestamps.index(ts)
                Jm = self.dcampose_dimupose_mat
                Hx[row:row + nz, L + 6 * pi:L + 6 * pi + 6] = jac_sensor[frow:frow + nz, :] @ Jm
                Hf_t[row:row + nz, :] = Hf[frow:frow + nz, :]
                res_t[row:row + nz] = res[frow:frow + nz]
                row += nz
                b0 = sum_zs + k * 4
                Hx[row:row + 4, L + 6 * pi:L + 6 * pi + 6] = jac_sensor[b0:b0 + 4, :] @ Jm
                Hf_t[row:row + 4, :] = Hf[b0:b0 + 4, :]
                res_t[row:row + 4] = res[b0:b0 + 4]
                row += 4
                frow += nz
                flag = True
            else:
                frow += nz
        return flag, Hx[:row], Hf_t[:row], res_t[:row]

    def removeLostObjects(self, H_x, H_f, res):
        """:2154-2193.  Returns a status string for the parity tests."""
        if H_x.shape[0] == 0 or res.shape[0] == 0:
            return "empty"
        if not self.p.use_object_residual_update_cam_pose_flag:
            return "disabled"
        ok, H_x, res = nullspace_project_inplace_svd(H_f, H_x, res)
        if not ok:
            return "nullspace_fail"
        g = {}
        if not self.gatingTestFeature(H_x, res, res.shape[0], g):
            self.log.append(dict(kind="object_gate_fail", **g))
            return "gate_fail"
        if np.isnan(H_x).any() or np.isnan(res).any():
            return "nan"
        self.log.append(dict(kind="object_gate_pass", **g))
        self.measurementUpdate_msckf(H_x, res)
        return "updated"

    # getters, :2962-3026 (T_imu_body = I)
    def getPpose(self):
        P = self.state_cov
        out = np.zeros((6, 6))
        out[0:3, 0:3] = P[6:9, 6:9]
        out[0:3, 3:6] = P[6:9, 0:3]
        out[3:6, 0:3] = P[0:3, 6:9]
        out[3:6, 3:6] = P[0:3, 0:3]
        return out

    def getPvel(self):
        return self.state_cov[3:6, 3:6].copy()


def nullspace_project_inplace_svd(H_f, H_x, res):
    """math_utils.hpp:287-312: A = U[:, cols:] of the full SVD of H_f."""
    if H_f.shape[0] <= H_f.shape[1]:
        return False, H_x, res
    U, _, _ = np.linalg.svd(H_f, full_matrices=True)
    A = U[:, H_f.shape[1]:]
    return True, A.T @ H_x, A.T @ res


def nullspace_project_inplace_qr(H_f, H_x, res):
    """math_utils.hpp:315-344 (ColPivHouseholderQR; the pivoting does not change the
    column space, so plain Householder QR spans the same Q2)."""
    if H_f.shape[0] - H_f.shape[1] <= 0:
        return False, H_x, res
    Q, _ = np.linalg.qr(H_f, mode="complete")
    Q2 = Q[:, H_f.shape[1]:]
    return True, Q2.T @ H_x, Q2.T @ res
