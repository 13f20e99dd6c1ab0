#!/usr/bin/env python
"""Generates tests/golden/cv_convert.npz with OpenCV's real cv::Mat::convertTo -- the routine the
reference calls at src/urdf_filter.cpp:288 (16UC1 -> 32FC1, alpha 0.001) and :311
(32FC1 -> 16UC1, alpha 1000.0).

The Python binding does not expose Mat::convertTo, but cv::normalize(src, dst, alpha, 0, NORM_INF,
rtype) is literally `scale = alpha / max|src|; src.convertTo(dst, rtype, scale, 0)` (OpenCV
modules/core/src/norm.cpp), so choosing alpha = wanted_scale * max|src| reaches convertTo with the
wanted scale.  For the float destination convertTo narrows the scale to float, so the tiny double
error of alpha / max does not matter (asserted below); for the 16U destination max|src| is a power
of two, which makes the double scale exactly 1000.

Run once where cv2 is importable; the .npz is committed:
    python tests/golden/make_cv_convert.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- 16U -> 32F, scale 0.001: all 65536 values ----
u16 = np.arange(65536, dtype=np.uint16).reshape(256, 256)
alpha = 0.001 * 65535.0
assert np.float32(alpha / 65535.0) == np.float32(0.001)
f32 = cv2.normalize(u16, None, alpha=alpha, beta=0, norm_type=cv2.NORM_INF, dtype=cv2.CV_32F)
assert f32.dtype == np.float32

# ---- 32F -> 16U, scale 1000: ties, saturation, negatives, denormals, infinities ----
rng = np.random.default_rng(20261017)
vals = [rng.uniform(0.0, 70.0, 40000), rng.uniform(-5.0, 0.0, 2000),
        (np.arange(0, 20000) + 0.5) / 1000.0,            # x.5 mm: round-half-even
        np.arange(0, 65536, 7) * np.float32(0.001),       # values that came from 16U frames
        [0.0, -0.0, 65.535, 65.5354, 65.5355, 65.536, 66.0, 100.0, 127.0, 1e-30, -1e-30,
         0.0005, 0.0015, 0.0025, 5.0, 7.92]]
src = np.concatenate([np.asarray(v, np.float64) for v in vals]).astype(np.float32)
src = np.concatenate([src, np.float32([128.0])])          # max|src| = 2^7 exactly
pad = (-src.size) % 256
src = np.concatenate([src, np.zeros(pad, np.float32)]).reshape(-1, 256)
assert float(np.abs(src).max()) == 128.0
out_u16 = cv2.normalize(src, None, alpha=128.0 * 1000.0, beta=0, norm_type=cv2.NORM_INF, dtype=cv2.CV_16U)
assert out_u16.dtype == np.uint16

np.savez_compressed(os.path.join(HERE, "cv_convert.npz"), u16=u16, f32_from_u16=f32, f32=src, u16_from_f32=out_u16,
                    cv_version=np.array(cv2.__version__))
print("wrote cv_convert.npz", cv2.__version__, f32.shape, src.shape)
