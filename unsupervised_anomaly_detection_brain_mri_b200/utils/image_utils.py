"""Slice helpers with the names of reference utils/image_utils.py (crop, crop_center,
augment_prediction_and_groundtruth_to_image)."""
import numpy as np


def crop(img, y, x, height, width):
    """The [y, y+height) x [x, x+width) window of an image."""
    return img[y:y + height, x:x + width]


def crop_center(img, cropx, cropy):
    """A cropx-wide, cropy-high window around the image centre (any trailing channel axis is kept)."""
    top = img.shape[0] // 2 - cropy // 2
    left = img.shape[1] // 2 - cropx // 2
    return img[top:top + cropy, left:left + cropx, ...]


def augment_prediction_and_groundtruth_to_image(image, p, g):
    """RGB overlay: true positives green, false positives orange, false negatives red, elsewhere the grey image."""
    grey = np.asarray(image, np.float64)
    if grey.ndim < 3:
        grey = grey[..., None]
    rgb = np.repeat(grey, 3, axis=2)
    rgb[rgb < 0] = 0
    pred, gt = np.squeeze(np.asarray(p).astype(bool)), np.squeeze(np.asarray(g).astype(bool))
    for mask, colour in ((pred & gt, (0.0, 1.0, 0.0)), (pred & ~gt, (1.0, 0.5, 0.0)), (~pred & gt, (1.0, 0.0, 0.0))):
        rgb[mask] = colour
    return rgb
