"""Generates tests/golden/bitplanes_ref.npz: the BitPlanes descriptor channels of a small image as the
reference's computeBitPlanes (src/imgproc.cc:222-245) produces them, with the two cv::GaussianBlur calls
executed by OpenCV itself (cv2, available in the build container only) and the census transform
(src/imgproc.cc:140-220: bit k = neighbour k >= centre, zero border) in numpy.

    python tests/golden/make_bitplanes_golden.py
"""
import os
import cv2
import numpy as np

rng = np.random.default_rng(20161201)
rows, cols = 45, 67
yy, xx = np.mgrid[0:rows, 0:cols]
img = 127.5 + 50 * np.cos(0.31 * xx + 0.17 * yy) + 35 * np.sin(0.23 * yy - 0.11 * xx) + 18 * rng.standard_normal((rows, cols))
img = np.clip(np.rint(img), 0, 255).astype(np.uint8)

blur = cv2.GaussianBlur(img, (3, 3), 1.0)                       # computeBitPlanes: sigma_ct = 1.0
census = np.zeros_like(blur)
c = blur[1:-1, 1:-1]
offs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
acc = np.zeros(c.shape, dtype=np.uint8)
for k, (dy, dx) in enumerate(offs):
    nb = blur[1 + dy:rows - 1 + dy, 1 + dx:cols - 1 + dx]
    acc |= ((nb >= c).astype(np.uint8) << k)
census[1:-1, 1:-1] = acc
planes = np.zeros((8, rows, cols), dtype=np.float32)
for b in range(8):
    bit = ((census >> b) & 1).astype(np.float32)
    planes[b] = cv2.GaussianBlur(bit, (5, 5), 1.5, sigmaY=1.5)  # ExtractBitPlanesChannel: sigma_bp = 1.5

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bitplanes_ref.npz")
np.savez_compressed(out, image=img, blur=blur, census=census, planes=planes, cv2_version=cv2.__version__)
print("wrote", out, os.path.getsize(out), "bytes; cv2", cv2.__version__)
