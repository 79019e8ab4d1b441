"""instageo_b200 -- B200-native (sm_100a) implementation of InstaGeo's chip-inference hot path.

Import as ``import instageo_b200`` (the repo-root shim maps that name onto this directory,
whose on-disk name ``instageo-e2e-geospatial-ml_b200`` is not a valid Python identifier).

Layout mirrors the reference modules that sit on the path:
  instageo_b200.model.model        <- instageo/model/model.py        (PrithviSeg)
  instageo_b200.model.dataloader   <- instageo/model/dataloader.py   (normalise / mask / crop grid)
  instageo_b200.model.infer_utils  <- instageo/model/infer_utils.py  (chip_inference, sliding window)
  instageo_b200.model.metrics      <- instageo/model/metrics.py      (streaming eval metrics, device-resident)
  instageo_b200.data.data_pipeline <- instageo/data/data_pipeline.py (apply_mask, mask_segmentation_map)
  instageo_b200.ops                   tensor wrappers over the C ABI (include/instageo_b200.h)
  instageo_b200.csrc                  hand-written CUDA kernels + the C ABI
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
__all__ = ["_lib", "ops", "model", "data"]


def __getattr__(name):
    if name in ("ops", "model", "data"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
