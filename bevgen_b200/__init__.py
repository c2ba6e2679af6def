"""bevgen_b200 — B200-native (sm_100a) implementation of BEVGen's two hot paths:
stage-1 VQGAN encode -> nearest-code quantise -> decode, and the stage-2 autoregressive multi-view
transformer with camera-bias attention (teacher-forced forward + KV-cache sampling).

All device arithmetic runs in hand-written CUDA kernels reached through the C-ABI in
``include/bevgen_b200.h`` (``libbevgen_b200.so``).  There is no CPU or eager-PyTorch fallback: the
package raises at first use when the library is missing.
"""
__version__ = "0.1.0"
