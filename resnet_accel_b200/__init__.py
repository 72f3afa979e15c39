"""resnet_accel_b200 - B200-native drop-in for the ACCEL-v1 BSR-INT8 hot path.

Layers (each mirrors a reference surface):
  * ``_lib`` / ``ops``   C ABI binding + torch-facing wrappers (device memory and streams only)
  * ``golden``           sw/golden layer calls  (gemm_bsr_int8_golden, load_bsr_layer, gemm_bsr_int8)
  * ``exporters``        sw/training + sw/exporters packing format, GPU packer / pruner behind it
  * ``host``             sw/host driver entry points (AccelDriver, BSRMatrix, pack_activations)
  * ``layers``           conv / linear layer objects and the MNIST / ResNet layer tables

There is no CPU fallback anywhere in this package: compute calls raise if libaccel_b200.so or a
CUDA device is missing.
"""
from ._lib import AcceleratorError, LIB_PATH  # noqa: F401

__all__ = ["AcceleratorError", "LIB_PATH"]
__version__ = "0.1.0"
