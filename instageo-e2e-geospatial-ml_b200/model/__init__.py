"""Host-side mirror of ``instageo.model`` for the chip-inference path."""
from .model import PrithviSeg, PrithviViT  # noqa: F401
from .dataloader import (  # noqa: F401
    crop_array, normalize_and_convert_to_tensor, process_and_augment, process_raw_chips, process_test,
)
from .infer_utils import (  # noqa: F401
    chip_inference, sliding_window_inference, sliding_window_inference_sharded,
)
