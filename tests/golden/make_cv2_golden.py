"""Generates the cv2 golden vectors that pin the oracle's restatement of the THIRD-PARTY image ops the
reference calls (cv::pyrDown at bpvo/image_pyramid.cc:49, cv::GaussianBlur 5x5 f32 at
bpvo/bitplanes_descriptor.cc:56, cv::GaussianBlur 3x3 u8 at bpvo/census.cc:65).  Run once in the build container (cv2 4.13.0); the GPU box never
needs cv2.  Output: tests/golden/cv2_golden.npz (small)."""
import os
import numpy as np
import cv2

rng = np.random.RandomState(0xB200)
out = {}
for i, (r, c) in enumerate([(47, 61), (48, 64), (33, 18), (94, 311)]):
    img = rng.randint(0, 256, size=(r, c)).astype(np.uint8)
    out[f"pyr_in_{i}"] = img
    out[f"pyr_out_{i}"] = cv2.pyrDown(img)
for i, ((r, c), sigma) in enumerate([((37, 53), 0.5), ((40, 64), 1.618), ((21, 19), 0.75), ((47, 156), 0.5)]):
    bits = (rng.randint(0, 2, size=(r, c))).astype(np.float32)
    out[f"blur_in_{i}"] = bits
    out[f"blur_sigma_{i}"] = np.float32(sigma)
    out[f"blur_out_{i}"] = cv2.GaussianBlur(bits, (5, 5), sigma, sigmaY=sigma)
f = rng.rand(29, 43).astype(np.float32) * 255
out["blur_in_f"] = f
out["blur_sigma_f"] = np.float32(1.2)
out["blur_out_f"] = cv2.GaussianBlur(f, (5, 5), 1.2, sigmaY=1.2)
# cv::GaussianBlur 3x3 on CV_8U (bpvo/census.cc:65, sigmaPriorToCensusTransform > 0); drawn AFTER the arrays above so that those stay unchanged
for i, ((r, c), sigma) in enumerate([((37, 53), 0.75), ((48, 64), 0.5), ((19, 23), 1.3), ((94, 311), 0.75), ((16, 24), 2.0)]):
    img = rng.randint(0, 256, size=(r, c)).astype(np.uint8)
    if i == 1:
        img[rng.rand(r, c) < 0.3] = 255
    out[f"cblur_in_{i}"] = img
    out[f"cblur_sigma_{i}"] = np.float32(sigma)
    out[f"cblur_out_{i}"] = cv2.GaussianBlur(img, (3, 3), sigma, sigmaY=sigma)
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv2_golden.npz"), **out)
print("wrote cv2_golden.npz with", len(out), "arrays")
