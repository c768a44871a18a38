"""pion_b200 -- B200 (sm_100a) implementation of PION's finite-volume hydro/MHD
dynamics update.

The product is ``libpion_b200.so`` (hand-written CUDA behind a C ABI, see
``include/pion_b200.h``) plus the C++ host mirror in ``pion_b200/host``.  This
Python package is only the thin ctypes binding used by the tests, the benchmark
and multi-GPU launch plumbing (``torch.distributed`` rendezvous for the NCCL
unique id); it contains no numerics and no fallback path: importing
``pion_b200.capi`` fails loudly when the CUDA library has not been built.
"""
from .capi import Context, GpuConfig, LIB_PATH, load_library  # noqa: F401
