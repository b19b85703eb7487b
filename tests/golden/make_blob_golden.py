"""Generates tests/golden/blobs_v1.npz: the blob path's golden vectors.

The reference's find_blobs.cc cannot be built in this image (no OpenCV C++ headers), so the fixtures
come from the same OpenCV class through its Python binding: cv2.SimpleBlobDetector with exactly the
four parameters find_blobs.cc:19-22 sets, then the float32 -> PointInt conversion of
find_blobs.cc:40-41. OpenCV version used: see 'cv2_version' inside the .npz (the reference does not
pin one). Run from the repo root:  python tests/golden/make_blob_golden.py
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import blob_golden_images  # noqa: E402


def reference_blobs(image):
    p = cv2.SimpleBlobDetector_Params()
    p.minArea = 20
    p.maxArea = 80000
    p.minDistBetweenBlobs = 5
    p.blobColor = 0
    kps = cv2.SimpleBlobDetector_create(p).detect(image)
    pts = np.array([kp.pt for kp in kps], dtype=np.float32).reshape(-1, 2)
    # (int)(pt.x * FIND_GRID_SCALE + 0.5): float32 product, double add, truncation
    scaled = (pts * np.float32(1000)).astype(np.float64) + 0.5
    return np.trunc(scaled).astype(np.int32)


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, img in blob_golden_images().items():
        out["img/" + name] = img
        out["pts/" + name] = reference_blobs(img)
        print(name, img.shape, len(out["pts/" + name]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "blobs_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
