"""Independent pin for the two kornia calls of the path (models/raytracer.py:557 `kornia.morphology.closing(depth, ones(3,3))`
and :569 `kornia.filters.sobel(depth)`): kornia cannot be installed in the build container, so the oracle restates the two
functions from kornia's documentation -- and this script pins that restatement (and through it the CUDA kernels
ironb_depth_closing / ironb_sobel_depth) against an INDEPENDENT implementation, OpenCV 4.13, which is present here:

    closing  = cv2.morphologyEx(depth, MORPH_CLOSE, ones(3,3))     default border = the morphology default value, i.e. pixels
               outside the image never win (kornia's border_type='geodesic', its default)
    sobel    = sqrt(gx^2 + gy^2 + 1e-6), gx/gy = cv2.Sobel(depth, CV_32F, 1/0, 0/1, ksize=3, scale=1/8, BORDER_REPLICATE)
               (kornia.filters.sobel: normalized=True divides the 3x3 kernels by 8, replicate padding, eps=1e-6)

    erosion  = cv2.erode(mask, ones(11,11))   (kornia.morphology.erosion of the SSIM mask, models/image_losses.py:154)

    python oracle/make_golden_cv2.py        -> tests/golden/morph_cv2.npz   (TEST INFRASTRUCTURE ONLY)
"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def depth_images():
    """Depth maps the way raytrace_camera produces them: a smooth surface (depth ~1.5..2.5) times a hit mask with holes,
    silhouettes, isolated pixels, and the degenerate all-zero / all-hit / 1-pixel-wide cases."""
    rng = np.random.default_rng(7)
    out = []
    for (h, w) in [(32, 32), (37, 53), (64, 64), (5, 3), (1, 9), (128, 96)]:
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        surf = (2.0 - 0.4 * np.cos(xx / max(w, 2) * 3.0) * np.sin(yy / max(h, 2) * 2.0)).astype(np.float32)
        surf += rng.normal(0, 0.01, size=(h, w)).astype(np.float32)
        disk = ((xx - w / 2) ** 2 / max(w * w / 5.0, 1) + (yy - h / 2) ** 2 / max(h * h / 5.0, 1)) < 1.0
        holes = rng.random((h, w)) < 0.08
        specks = rng.random((h, w)) < 0.02
        mask = (disk & ~holes) | specks
        out.append((surf * mask).astype(np.float32))
    out.append(np.zeros((16, 16), np.float32))
    out.append(np.full((16, 16), 1.75, np.float32))
    return out


def main():
    k3 = np.ones((3, 3), np.uint8)
    data = {}
    for i, d in enumerate(depth_images()):
        closed = cv2.morphologyEx(d, cv2.MORPH_CLOSE, k3)
        gx = cv2.Sobel(d, cv2.CV_32F, 1, 0, ksize=3, scale=1.0 / 8.0, borderType=cv2.BORDER_REPLICATE)
        gy = cv2.Sobel(d, cv2.CV_32F, 0, 1, ksize=3, scale=1.0 / 8.0, borderType=cv2.BORDER_REPLICATE)
        data[f"depth{i}"] = d
        data[f"closing{i}"] = closed.reshape(d.shape).astype(np.float32)
        data[f"sobel{i}"] = np.sqrt(gx.astype(np.float32) ** 2 + gy.astype(np.float32) ** 2 + np.float32(1e-6)).reshape(d.shape)
    # kornia.morphology.erosion(mask, ones(11, 11)) of ssim_loss_fn (models/image_losses.py:154): cv2.erode, default border
    k11 = np.ones((11, 11), np.uint8)
    rng = np.random.default_rng(9)
    masks = []
    for (h, w) in [(64, 64), (37, 53), (16, 16), (9, 30), (128, 128)]:
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        m = ((xx - 0.55 * w) ** 2 / (0.42 * w) ** 2 + (yy - 0.5 * h) ** 2 / (0.46 * h) ** 2) < 1.0
        m &= ~(rng.random((h, w)) < 0.004)
        masks.append(m)
    masks.append(np.ones((32, 32), bool))
    for i, m in enumerate(masks):
        data[f"mask{i}"] = m
        data[f"erode11_{i}"] = cv2.erode(m.astype(np.uint8), k11).reshape(m.shape).astype(bool)
    data["n_masks"] = np.int64(len(masks))
    data["n"] = np.int64(len(depth_images()))
    data["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(ROOT, "tests", "golden", "morph_cv2.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
