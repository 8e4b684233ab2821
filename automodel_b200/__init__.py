"""B200-native sharded-data-parallel training step behind the `automodel` recipe surface.

Host orchestration in Python over torch tensors; all device math in hand-written sm_100a CUDA kernels
bound through the C ABI in include/b200_train.h (automodel_b200/csrc -> libb200_train.so).
There is no CPU or library fallback: importing ops on a machine without the built extension raises.
"""
__version__ = "0.1.0"
