"""B200-native sharded-data-parallel training step behind the `automodel` recipe surface.

Host orchestration in Python over torch tensors; all device math in hand-written sm_100a CUDA kernels
bound through the C ABI in include/b200_train.h (automodel_b200/csrc -> libb200_train.so).
There is no CPU or library fallback: importing ops on a machine without the built extension raises.
"""
__version__ = "0.1.0"

import os as _os

# One hardware launch queue per stream (the driver's default is 8 queues shared by all streams of the process).  The step keeps several
# streams busy (compute, communication, optimizer, weight-gradient) and, at N > 1, runs kernels that wait for a peer GPU inside the kernel
# (csrc/comm.cu): such a kernel must never sit in front of unrelated work in a shared queue.  Only effective when set before the CUDA
# context is created (import this package before the first CUDA call); a user-provided value wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
