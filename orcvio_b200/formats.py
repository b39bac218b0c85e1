"""Host-side mirrors of the data formats either side of the filter path: thin ctypes wrappers over
csrc/formats.cpp -- EuRoC / ASL csv readers (include/utils/DataReader.hpp:30-140), the ground-truth file and its
lookup (include/orcvio/dataset_reader.h:64-140), the pose log (src/orcvio.cpp:640-645) and the ObjectLM message
(ros_wrapper/src/orcvio_ros_msgs/msg/ObjectLM.msg)."""
import ctypes as C

import numpy as np

from . import api


def _lib():
    L = api.lib()
    if not getattr(L, "_formats_ready", False):
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orcvio_read_imu_csv.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.orcvio_read_image_list_csv.argtypes = [C.c_char_p, dp, C.c_char_p, C.c_int, C.c_int]
        L.orcvio_read_gt_csv.argtypes = [C.c_char_p, dp, C.c_int]
        L.orcvio_gt_lookup.argtypes = [dp, C.c_int, C.c_double, dp]
        L.orcvio_read_pose_log.argtypes = [C.c_char_p, dp, C.c_int]
        L.orcvio_objectlm_pack.argtypes = [C.c_longlong, dp, C.c_int, dp, C.c_int, dp, dp, C.c_int, dp, C.c_int, ip, C.c_int,
                                           C.c_void_p, C.c_int]
        L.orcvio_objectlm_unpack.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), dp, ip, dp, ip, dp, dp, ip, dp, ip,
                                             ip, ip, C.c_int, C.c_int, C.c_int]
        L._formats_ready = True
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def read_imu_csv(path):
    """loadImuFile -> struct array (api.IMU_DTYPE)."""
    L = _lib()
    n = L.orcvio_read_imu_csv(str(path).encode(), None, 0)
    if n < 0:
        raise FileNotFoundError(path)
    out = np.zeros(n, dtype=api.IMU_DTYPE)
    L.orcvio_read_imu_csv(str(path).encode(), out.ctypes.data, n)
    return out


def read_image_list_csv(path):
    """loadImageList -> (stamps [s], file names)."""
    L = _lib()
    n = L.orcvio_read_image_list_csv(str(path).encode(), None, None, 0, 0)
    if n < 0:
        raise FileNotFoundError(path)
    t = np.zeros(n)
    names = C.create_string_buffer(n * 128)
    L.orcvio_read_image_list_csv(str(path).encode(), _dp(t), names, 128, n)
    return t, [names.raw[i * 128:(i + 1) * 128].split(b"\0", 1)[0].decode() for i in range(n)]


def read_gt_csv(path):
    """load_gt_file -> (n, 17): [t (s), q (4), p (3), v (3), bg (3), ba (3)]."""
    L = _lib()
    n = L.orcvio_read_gt_csv(str(path).encode(), None, 0)
    if n < 0:
        raise FileNotFoundError(path)
    out = np.zeros((n, 17))
    L.orcvio_read_gt_csv(str(path).encode(), _dp(out), n)
    return out


def gt_lookup(gt, t):
    """get_gt_state -> the 17-vector, or None."""
    gt = np.ascontiguousarray(gt, dtype=np.float64)
    out = np.zeros(17)
    return out if _lib().orcvio_gt_lookup(_dp(gt), len(gt), float(t), _dp(out)) else None


def read_pose_log(path):
    L = _lib()
    n = L.orcvio_read_pose_log(str(path).encode(), None, 0)
    if n < 0:
        raise FileNotFoundError(path)
    out = np.zeros((n, 8))
    L.orcvio_read_pose_log(str(path).encode(), _dp(out), n)
    return out


def objectlm_pack(object_id, residual, jac_object, jac_sensor, cam_pose_se3, timestamps, zs_num):
    """ObjectLM message bytes (ROS 1 serialisation)."""
    L = _lib()
    r = np.ascontiguousarray(residual, dtype=np.float64).ravel()
    jo = np.ascontiguousarray(jac_object, dtype=np.float64).reshape(len(r), -1)
    js = np.ascontiguousarray(jac_sensor, dtype=np.float64).reshape(len(r), 6)
    cp = np.ascontiguousarray(cam_pose_se3, dtype=np.float64).reshape(6, -1)
    ts = np.ascontiguousarray(timestamps, dtype=np.float64).ravel()
    zs = np.ascontiguousarray(zs_num, dtype=np.int32).ravel()
    args = (int(object_id), _dp(r), len(r), _dp(jo), jo.shape[1], _dp(js), _dp(cp), cp.shape[1], _dp(ts), len(ts),
            zs.ctypes.data_as(C.POINTER(C.c_int)), len(zs))
    n = L.orcvio_objectlm_pack(*args, None, 0)
    buf = np.zeros(n, dtype=np.uint8)
    L.orcvio_objectlm_pack(*args, buf.ctypes.data, n)
    return buf.tobytes()


def objectlm_unpack(data, cap_rows=1024, cap_odim=64, cap_n=256):
    L = _lib()
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    oid = C.c_longlong(0)
    rows, odim, npos, nts, nzs = (C.c_int(0) for _ in range(5))
    r = np.zeros(cap_rows)
    jo = np.zeros(cap_rows * cap_odim)
    js = np.zeros(cap_rows * 6)
    cp = np.zeros(6 * cap_n)
    ts = np.zeros(cap_n)
    zs = np.zeros(cap_n, dtype=np.int32)
    rc = L.orcvio_objectlm_unpack(buf.ctypes.data, len(buf), C.byref(oid), _dp(r), C.byref(rows), _dp(jo), C.byref(odim),
                                  _dp(js), _dp(cp), C.byref(npos), _dp(ts), C.byref(nts), zs.ctypes.data_as(C.POINTER(C.c_int)),
                                  C.byref(nzs), cap_rows, cap_odim, cap_n)
    if rc != 0:
        raise ValueError("malformed ObjectLM message")
    R, O, Np = rows.value, odim.value, npos.value
    return dict(object_id=oid.value, residual=r[:R].copy(), jacobian_wrt_object_state=jo[:R * O].reshape(R, O).copy(),
                jacobian_wrt_sensor_state=js[:R * 6].reshape(R, 6).copy(), valid_camera_pose_mat=cp[:6 * Np].reshape(6, Np).copy(),
                timestamps=ts[:nts.value].copy(), zs_num_wrt_timestamps=zs[:nzs.value].copy())
