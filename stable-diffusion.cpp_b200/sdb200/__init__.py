"""sdb200 -- host-side Python plumbing for the B200-native ggml backend (libggml-b200.so).

The product is the C-ABI plugin in ../csrc (CUDA, sm_100a) plus the C++ harness in ../harness;
this package only loads them (ctypes) for tests and bench.py.
"""
from .harness import Harness, Model, FLAG_FLASH_ATTN, FLAG_CONV_DIRECT, FLAG_NO_WEIGHT_VALUES, B200_SO, HARNESS_SO, REPO  # noqa: F401
