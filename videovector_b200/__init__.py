"""videovector_b200 -- B200 (sm_100a) implementation of the temporal-context embedding
training hot path of eevignesh/videovector, behind the C-ABI in include/vv_b200.h.

Only what the path needs lives here:
  csrc/      CUDA kernels (tcgen05/TMA GEMMs, fused rank-loss, gather, SGD update) + C-ABI
  csrc/host/ host C++: index sampler, data-parallel trainer, Caffe-interface mirror
  _lib.py    ctypes binding (no fallback: raises if libvv_b200.so is missing)
  ops.py     torch-tensor wrappers used by tests and bench.py
"""
from ._lib import VVError, PREC, load, LIB_PATH  # noqa: F401
