"""Host-side mirror of the array arithmetic in ``instageo.data`` that sits next to the hot path."""
from .data_pipeline import (  # noqa: F401
    MASK_DECODING_POS, apply_mask, create_chip, decode_fmask_value, mask_segmentation_map,
)
from .geotiff import read_geotiff, read_geotiff_device, write_geotiff, write_geotiffs_device  # noqa: F401
