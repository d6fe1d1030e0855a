"""Context-encoder patch masking (mirror of reference trainers/CE.py:123-139)."""
import random

import numpy as np


def retrieve_masked_batch(batch, brainmasks):
    """1-3 random 20x20 patches inside each sample's brain bounding box are zeroed.

    Reference quirk preserved (CE.py:130-138): the loop variable shadows the mask array, so the mask that is
    finally multiplied is the LAST sample's [H,W,C] mask, broadcast over the whole batch."""
    def retrieve_brain_range(brainmask):
        pixels = np.argwhere(brainmask).T
        return (min(pixels[0]), max(pixels[0])), (min(pixels[1]), max(pixels[1]))

    brain_ranges = [retrieve_brain_range(bm) for bm in brainmasks]
    masks = np.ones(batch.shape)
    last = masks[0]
    for sample_mask, brain_range in zip(masks, brain_ranges):
        last = sample_mask
        for _ in range(random.randint(1, 3)):
            size_w, size_h = 20, 20
            if brain_range[0][0] < brain_range[0][1] - size_w and brain_range[1][0] < brain_range[1][1] - size_h:
                px = random.randint(brain_range[0][0], brain_range[0][1] - size_w)
                py = random.randint(brain_range[1][0], brain_range[1][1] - size_h)
                sample_mask[px:px + size_w, py:py + size_h] = 0
    return (batch * last).astype(batch.dtype)
