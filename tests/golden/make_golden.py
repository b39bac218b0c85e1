"""Regenerates the committed golden fixtures from the reference's own test data.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Sources: /root/reference/src/tests/data/test_error_feature_quadric.h5 and
test_error_bbox_quadric.h5, test_error_deform_reg.h5, test_error_mean_shape_reg.h5 (golden `error` / `jacobian` used
by the reference's src/tests/test_object_lm.cpp:90-295, 482-584) and one_car/frame_*.h5, one_car_no_zb/frame_*.h5 (multi-frame object
observations, src/tests/test_object_lm_multiframe.cpp, test_object_init_multiframe.cpp).  Data files only -- no reference
source code is copied.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.hdf5_min import read_h5  # noqa: E402

SRC = "/root/reference/src/tests/data"


def main():
    for name in ("test_error_feature_quadric", "test_error_bbox_quadric", "test_error_deform_reg",
                 "test_error_mean_shape_reg"):
        d = read_h5(os.path.join(SRC, name + ".h5"))
        np.savez(os.path.join(HERE, name + ".npz"), **{k: v.astype(np.float64) for k, v in d.items()})
        print(name, {k: v.shape for k, v in d.items()})
    for seq in ("one_car", "one_car_no_zb"):
        frames = {}
        i = 0
        while os.path.exists(os.path.join(SRC, seq, f"frame_{i}.h5")):
            d = read_h5(os.path.join(SRC, seq, f"frame_{i}.h5"))
            for k, v in d.items():
                frames.setdefault(k, []).append(v.astype(np.float64))
            i += 1
        np.savez(os.path.join(HERE, seq + ".npz"), **{k: np.stack(v) for k, v in frames.items()})
        print(seq, "frames:", i, {k: np.stack(v).shape for k, v in frames.items()})


if __name__ == "__main__":
    main()
