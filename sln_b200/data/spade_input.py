"""SPADE shading input preparation: depth EXR + per-class masks -> the [1, 41, 256, 256] tensor SPADEGenerator4 eats.

Reference: testing/test_SPADE_shade.py:50-76.  Host-side numpy (it runs once per room, before the hot path):
  depth:  channel 0 of the EXR, shifted to min 0, clipped at the largest value below 20, scaled to [0, 1], then to [-1, 1]   (:54-60)
  masks:  40 NYU classes at 1024 x 1024; values < 120 -> 0, > 120 -> 1 (a pixel equal to 120 keeps 120, as in the reference)   (:61-73)
  resize: skimage.transform.resize(total, [256, 256], preserve_range=True, order=3, anti_aliasing=True)                         (:75)

`skimage` is not installed in this image, so the resize is a restatement of skimage >= 0.19's algorithm on scipy.ndimage (Gaussian
pre-filter with sigma = (factor - 1) / 2 in 'mirror' mode, cubic-spline zoom with grid_mode=True, clip to the input range) — PARITY
UNPINNED for that one call; everything around it is pinned by executing the reference's own statements (tests/test_spade_input.py).
"""
import numpy as np

NYU_CLASS = ['wall', 'floor', 'cabinet', 'bed', 'chair', 'sofa', 'table', 'door', 'window', 'bookshelf', 'picture', 'counter', 'blinds',
             'desk', 'shelves', 'curtain', 'dresser', 'pillow', 'mirror', 'floor_mat', 'clothes', 'ceiling', 'books', 'refridgerator',
             'television', 'paper', 'towel', 'shower_curtain', 'box', 'whiteboard', 'person', 'night_stand', 'toilet', 'sink', 'lamp',
             'bathtub', 'bag', 'otherstructure', 'otherfurniture', 'otherprop']      # test_SPADE_shade.py:31-36 (underscore spelling)


def resize_bicubic_antialiased(image, out_hw):
    """skimage.transform.resize(image [H, W, C], out_hw, preserve_range=True, order=3, anti_aliasing=True, mode='reflect', clip=True)
    as skimage >= 0.19 computes it, on scipy.ndimage."""
    from scipy import ndimage as ndi
    image = np.asarray(image, dtype=np.float64)
    out_shape = (int(out_hw[0]), int(out_hw[1])) + tuple(image.shape[2:])
    factors = np.divide(image.shape, out_shape)
    sigma = np.maximum(0, (factors - 1) / 2)
    filtered = ndi.gaussian_filter(image, sigma, cval=0, mode='mirror') if np.any(sigma > 0) else image
    out = ndi.zoom(filtered, [1 / f for f in factors], order=3, mode='mirror', cval=0, grid_mode=True)
    return np.clip(out, image.min(), image.max())


def class_of_mask_file(basename):
    """'<room>_<x>_<y>_<class>[_<class2>].png' -> NYU class name (test_SPADE_shade.py:64-71)."""
    parts = basename.split(".")[0].split("_")
    return parts[3] + "_" + parts[4] if len(parts) == 5 else parts[3]


def prepare_spade_input(depth, masks, out_size=256, resize=resize_bicubic_antialiased):
    """depth: [H, W] float array (EXR channel 0); masks: {NYU class name: [H, W] array of 0..255} -> float32 [1, 41, out, out]
    (channel 0 depth in [-1, 1], channels 1..40 the class masks in NYU order), ready for torch.from_numpy(...).cuda()."""
    depth = np.asarray(depth)
    depth = depth - np.min(depth)
    depth_max = np.max(depth[depth < 20])
    depth = np.clip(depth, 0, depth_max)
    depth = depth / depth_max
    depth = ((depth - 0.5) * 2).astype("float32")[None, :]
    H, W = depth.shape[1:]
    buffer = np.zeros((40, H, W))
    for name, m in masks.items():
        buffer[NYU_CLASS.index(name)] = np.asarray(m)
    buffer = buffer.astype("float32")
    buffer[buffer < 120] = 0.0
    buffer[buffer > 120] = 1.0
    total = np.vstack([depth, buffer])
    total = np.moveaxis(total, 0, 2)
    total = resize(total, [out_size, out_size])
    total = np.moveaxis(total, 2, 0)[None, :]
    return np.ascontiguousarray(total, dtype=np.float32)
