"""CPU: the data formats either side of the path (csrc/formats.cpp) against independent parsers written here."""
import struct

import numpy as np
import pytest

from orcvio_b200 import formats


def test_euroc_imu_and_image_csv(tmp_path):
    rng = np.random.default_rng(0)
    t_ns = 1403636579758555392 + np.arange(50, dtype=np.int64) * 5_000_000
    vals = rng.normal(0, 1, (50, 6))
    p = tmp_path / "imu.csv"
    with open(p, "w") as f:
        f.write("#timestamp [ns],w_RS_S_x [rad s^-1],w_y,w_z,a_RS_S_x [m s^-2],a_y,a_z\n")
        for t, v in zip(t_ns, vals):
            f.write(f"{t}," + ",".join(repr(float(x)) for x in v) + "\r\n")
    imu = formats.read_imu_csv(p)
    assert len(imu) == 50
    np.testing.assert_array_equal(imu["t"], 1e-9 * t_ns.astype(np.float64))
    np.testing.assert_array_equal(imu["gyro"], vals[:, :3])
    np.testing.assert_array_equal(imu["acc"], vals[:, 3:])
    q = tmp_path / "data.csv"
    with open(q, "w") as f:
        f.write("#timestamp [ns],filename\n")
        for t in t_ns[:7]:
            f.write(f"{t},{t}.png\n")
    t, names = formats.read_image_list_csv(q)
    np.testing.assert_array_equal(t, 1e-9 * t_ns[:7].astype(np.float64))
    assert names == [f"{x}.png" for x in t_ns[:7]]
    with pytest.raises(FileNotFoundError):
        formats.read_imu_csv(tmp_path / "missing.csv")


def test_groundtruth_csv_and_lookup(tmp_path):
    rng = np.random.default_rng(1)
    t_ns = 1403636580838555648 + np.arange(40, dtype=np.int64) * 5_000_000
    rows = rng.normal(0, 1, (40, 16))
    p = tmp_path / "gt.csv"
    with open(p, "w") as f:
        f.write("#timestamp, p_RS_R_x [m], ...\n")
        for t, r in zip(t_ns, rows):
            f.write(f"{t}," + ",".join(repr(float(x)) for x in r) + "\n")
    gt = formats.read_gt_csv(p)
    assert gt.shape == (40, 17)
    np.testing.assert_array_equal(gt[:, 0], 1e-9 * t_ns.astype(np.float64))
    np.testing.assert_array_equal(gt[:, 1:], rows)
    # get_gt_state: the closest stamp when it is within 5 ms, otherwise only an exact hit
    hit = formats.gt_lookup(gt, gt[10, 0] + 0.002)
    np.testing.assert_array_equal(hit, gt[10])
    assert formats.gt_lookup(gt, gt[-1, 0] + 0.2) is None
    np.testing.assert_array_equal(formats.gt_lookup(gt, gt[3, 0]), gt[3])


def _ros_matrix(m):
    m = np.asarray(m, dtype=np.float64)
    rows, cols = m.shape
    out = struct.pack("<I", 2)
    out += struct.pack("<III", 0, rows, rows * cols) + struct.pack("<III", 0, cols, cols)
    out += struct.pack("<I", 0) + struct.pack("<I", rows * cols) + m.tobytes(order="C")
    return out


def test_objectlm_wire_format():
    """Byte-for-byte against an encoder written from the ROS 1 serialisation rules (little endian, uint32 length
    prefixes; Float64MultiArray = layout {dim[] {label, size, stride}, data_offset} + data) and the layout
    tf::matrixEigenToMsg produces; then the round trip."""
    rng = np.random.default_rng(2)
    rows, odim, n = 56, 45, 3
    res = rng.normal(0, 1, rows)
    jo = rng.normal(0, 1, (rows, odim))
    js = rng.normal(0, 1, (rows, 6))
    cp = rng.normal(0, 1, (6, n))
    ts = np.array([10.5, 10.6, 10.7])
    zs = np.array([12, 11, 12], dtype=np.int32)
    msg = formats.objectlm_pack(7, res, jo, js, cp, ts, zs)
    want = struct.pack("<q", 7) + _ros_matrix(res.reshape(-1, 1)) + _ros_matrix(jo) + _ros_matrix(js) + _ros_matrix(cp)
    want += struct.pack("<I", n) + ts.tobytes() + struct.pack("<I", n) + zs.tobytes()
    assert msg == want
    back = formats.objectlm_unpack(msg)
    assert back["object_id"] == 7
    np.testing.assert_array_equal(back["residual"], res)
    np.testing.assert_array_equal(back["jacobian_wrt_object_state"], jo)
    np.testing.assert_array_equal(back["jacobian_wrt_sensor_state"], js)
    np.testing.assert_array_equal(back["valid_camera_pose_mat"], cp)
    np.testing.assert_array_equal(back["timestamps"], ts)
    np.testing.assert_array_equal(back["zs_num_wrt_timestamps"], zs)
    with pytest.raises(ValueError):
        formats.objectlm_unpack(msg[:-3])


def test_pose_log_reader(tmp_path):
    p = tmp_path / "state_est_geo_feat.txt"
    rows = np.array([[0.1, 1, 2, 3, 0, 0, 0, 1], [0.2, 1.5, 2.5, 3.5, 0, 0, 0.70710678, 0.70710678]])
    with open(p, "w") as f:
        for r in rows:
            f.write(" ".join(f"{x:.8g}" for x in r) + "\n")
    np.testing.assert_allclose(formats.read_pose_log(p), rows, rtol=1e-8)
