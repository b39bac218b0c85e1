"""CPU oracle -- math utilities (TEST INFRASTRUCTURE, not product code).

NumPy restatement of the small closed-form operators the OrcVIO filter update is
built from.  Every function cites the reference file:line it follows
(paths relative to the upstream repo root).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import math
import numpy as np

GRAVITY_ACCELERATION = 9.81  # include/orcvio/imu_state.h:20


def skew(w):
    """skewSymmetric, include/orcvio/utils/math_utils.hpp:27-39."""
    return np.array([[0.0, -w[2], w[1]],
                     [w[2], 0.0, -w[0]],
                     [-w[1], w[0], 0.0]])


def quaternion_normalize(q):
    """math_utils.hpp:68-72."""
    return q / np.linalg.norm(q)


def small_angle_quaternion(dtheta):
    """smallAngleQuaternion, math_utils.hpp:104-121 ([x,y,z,w])."""
    dq = np.asarray(dtheta, dtype=float) / 2.0
    n2 = float(dq @ dq)
    q = np.zeros(4)
    q[:3] = dq
    if n2 <= 1:
        q[3] = math.sqrt(1 - n2)
    else:
        q[3] = 1
        q = q / math.sqrt(1 + n2)
    return q


def quaternion_to_rotation(q):
    """quaternionToRotation (Hamilton, [x,y,z,w]), math_utils.hpp:164-177."""
    qx, qy, qz, qw = q
    return np.array([
        [1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy)],
        [2 * (qx * qy + qw * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qw * qx)],
        [2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), 1 - 2 * (qx * qx + qy * qy)]])


def eigen_quat_to_rotation(w, x, y, z):
    """Eigen::Quaterniond(w,x,y,z).toRotationMatrix() (no normalisation), used at
    src/orcvio.cpp:858-862,4516.  Same polynomial as Eigen's implementation."""
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def rotation_to_quaternion(R):
    """rotationToQuaternion, math_utils.hpp:188-227 ([x,y,z,w], w>=0)."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    score = [R[0, 0], R[1, 1], R[2, 2], tr]
    k = int(np.argmax(score))  # first maximal coefficient, like Eigen maxCoeff
    q = np.zeros(4)
    if k == 0:
        q[0] = math.sqrt(1 + 2 * R[0, 0] - tr) / 2.0
        q[1] = (R[0, 1] + R[1, 0]) / (4 * q[0])
        q[2] = (R[0, 2] + R[2, 0]) / (4 * q[0])
        q[3] = (R[2, 1] - R[1, 2]) / (4 * q[0])
    elif k == 1:
        q[1] = math.sqrt(1 + 2 * R[1, 1] - tr) / 2.0
        q[0] = (R[0, 1] + R[1, 0]) / (4 * q[1])
        q[2] = (R[1, 2] + R[2, 1]) / (4 * q[1])
        q[3] = (R[0, 2] - R[2, 0]) / (4 * q[1])
    elif k == 2:
        q[2] = math.sqrt(1 + 2 * R[2, 2] - tr) / 2.0
        q[0] = (R[0, 2] + R[2, 0]) / (4 * q[2])
        q[1] = (R[1, 2] + R[2, 1]) / (4 * q[2])
        q[3] = (R[1, 0] - R[0, 1]) / (4 * q[2])
    else:
        q[3] = math.sqrt(1 + tr) / 2.0
        q[0] = (R[2, 1] - R[1, 2]) / (4 * q[3])
        q[1] = (R[0, 2] - R[2, 0]) / (4 * q[3])
        q[2] = (R[1, 0] - R[0, 1]) / (4 * q[3])
    if q[3] < 0:
        q = -q
    return quaternion_normalize(q)


def Hl_operator(g):
    """Hl_operator, math_utils.hpp:230-249."""
    n = math.sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2])
    term1 = 0.5 * np.eye(3)
    if n < 1.0e-5:
        return term1
    S = skew(g)
    term2 = ((n - math.sin(n)) / n ** 3) * S
    term3 = ((2 * (math.cos(n) - 1) + n ** 2) / (2 * n ** 4)) * (S @ S)
    return term1 + term2 + term3


def Jl_operator(g):
    """Jl_operator, math_utils.hpp:251-270."""
    n = math.sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2])
    term1 = np.eye(3)
    if n < 1.0e-5:
        return term1
    S = skew(g)
    term2 = ((1 - math.cos(n)) / n ** 2) * S
    term3 = ((n - math.sin(n)) / n ** 3) * (S @ S)
    return term1 + term2 + term3


def so3_exp(omega):
    """Sophus::SO3d::exp(omega).matrix() (third party, Sophus v1.0.0 pinned by
    ros_wrapper/install-deps/install-sophus.mk; published algorithm: unit quaternion
    (cos(t/2), sin(t/2)/t * omega) with a Taylor branch below 1e-10, then
    Eigen toRotationMatrix).  Call sites: src/orcvio.cpp:919,4331,4497,4542."""
    omega = np.asarray(omega, dtype=float)
    theta_sq = float(omega @ omega)
    theta = math.sqrt(theta_sq)
    half = 0.5 * theta
    if theta < 1e-10:
        po4 = theta_sq * theta_sq
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * po4
        real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * po4
    else:
        imag = math.sin(half) / theta
        real = math.cos(half)
    return eigen_quat_to_rotation(real, imag * omega[0], imag * omega[1], imag * omega[2])


def so3_log(R):
    """Rotation matrix -> rotation vector (used by se3_log; Sophus SO3::log through
    the unit quaternion)."""
    q = rotation_to_quaternion(R)  # [x,y,z,w], w >= 0
    v = q[:3]
    n2 = float(v @ v)
    n = math.sqrt(n2)
    w = q[3]
    if n2 < 1e-20:
        two_atan = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w)
    else:
        two_atan = 2.0 * math.atan2(n, w) / n
    return two_atan * v


def se3_exp(xi):
    """Sophus::SE3d::exp([upsilon(3), omega(3)]).matrix(); call site
    src/orcvio.cpp:2083.  Published algorithm: R = exp(omega), t = V(omega) upsilon,
    V = I + (1-cos t)/t^2 W + (t - sin t)/t^3 W^2."""
    xi = np.asarray(xi, dtype=float)
    ups, om = xi[:3], xi[3:]
    R = so3_exp(om)
    theta = math.sqrt(float(om @ om))
    W = skew(om)
    if theta < 1e-10:
        V = R.copy()
    else:
        V = (np.eye(3) + (1 - math.cos(theta)) / theta ** 2 * W
             + (theta - math.sin(theta)) / theta ** 3 * (W @ W))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ ups
    return T


def se3_log(T):
    """Sophus::SE3d(T).log() -> [upsilon, omega]; call site
    src/obj/ObjectResJacCam.cpp:594."""
    R = T[:3, :3]
    t = T[:3, 3]
    om = so3_log(R)
    theta = math.sqrt(float(om @ om))
    W = skew(om)
    if theta < 1e-10:
        Vinv = np.eye(3) - 0.5 * W + (1.0 / 12.0) * (W @ W)
    else:
        half = 0.5 * theta
        Vinv = (np.eye(3) - 0.5 * W
                + (1 - theta * math.cos(half) / (2 * math.sin(half))) / theta ** 2 * (W @ W))
    return np.concatenate([Vinv @ t, om])


def inverse_pose(T):
    """inversePose, include/orcvio/utils/se3_ops.hpp:137-168."""
    iT = np.eye(4)
    iT[:3, :3] = T[:3, :3].T
    iT[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return iT


def odot(x):
    """odotOperator, se3_ops.hpp:510-519: 4x6 [x4 I, -skew(x123); 0]."""
    out = np.zeros((4, 6))
    out[:3, 3:] = -skew(x[:3])
    out[0, 0] = out[1, 1] = out[2, 2] = x[3]
    return out


def circled_circ(x):
    """circledCirc, se3_ops.hpp:229-240: 6x4."""
    out = np.zeros((6, 4))
    out[3:, :3] = -skew(x[:3])
    out[:3, 3] = x[:3]
    return out


def project_image_df(x):
    """project_image_df, se3_ops.hpp:325-339."""
    z = x[2]
    zsq = z * z
    return np.array([[1 / z, 0.0, -x[0] / zsq],
                     [0.0, 1 / z, -x[1] / zsq]])


def cam_wrt_imu_se3_jacobian(R_b2c, t_c_b, R_w2c, t_b_w, left):
    """get_cam_wrt_imu_se3_jacobian, se3_ops.hpp:531-552."""
    J = np.zeros((6, 6))
    if left:
        J[0:3, 0:3] = skew(t_b_w)
        J[3:6, 0:3] = np.eye(3)
        J[0:3, 3:6] = np.eye(3)
    else:
        J[0:3, 0:3] = -R_b2c @ skew(t_c_b)
        J[3:6, 0:3] = R_b2c
        J[0:3, 3:6] = R_w2c
    return J


def angle_axis_angle(R):
    """Eigen::AngleAxisd(R).angle() (via quaternion: 2*atan2(|vec|, |w|));
    call site src/orcvio.cpp:2606-2607."""
    q = rotation_to_quaternion_eigen(R)
    n = math.sqrt(q[1] ** 2 + q[2] ** 2 + q[3] ** 2)
    return 2.0 * math.atan2(n, abs(q[0]))


def rotation_to_quaternion_eigen(R):
    """Eigen's Quaternion(Matrix3) (Shoemake), returns [w,x,y,z]."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = [0.0] * 4
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (R[2, 1] - R[1, 2]) * t
        q[2] = (R[0, 2] - R[2, 0]) * t
        q[3] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[k, j] - R[j, k]) * t
        q[1 + j] = (R[j, i] + R[i, j]) * t
        q[1 + k] = (R[k, i] + R[i, k]) * t
    return q


_CHI2_CACHE = {}


def chi2_table(p, n=500):
    """Chi-square quantile table, src/orcvio.cpp:481-494 (boost::math::quantile of
    chi_squared(i) at probability p, i = 1..n-1).  Boost 1.65 (Ubuntu 18.04 package)
    is third party and absent; scipy's chi2.ppf computes the same inverse regularised
    incomplete gamma.  Cross-checked against mpmath in tests/test_oracle_math.py."""
    key = (float(p), int(n))
    if key not in _CHI2_CACHE:
        from scipy.stats import chi2
        tab = np.zeros(n)
        tab[1:] = chi2.ppf(p, np.arange(1, n))
        _CHI2_CACHE[key] = tab
    return _CHI2_CACHE[key]
