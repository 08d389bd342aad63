"""``sige.utils`` -> ``sige_b200.masks`` (reference sige/utils.py:8,40,74,88)."""
from sige_b200.masks import compute_difference_mask, dilate_mask, downsample_mask, reduce_mask  # noqa: F401
