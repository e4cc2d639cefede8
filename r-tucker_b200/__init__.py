"""rtucker_b200 -- B200-native (sm_100a) hot path of R-TuckER behind the reference's Python API.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every arithmetic step of
the path runs in hand-written CUDA reached through the C ABI of ``librtucker_b200.so``
(include/rtucker.h).  There is no CPU fallback, no Triton and no multi-backend dispatch.
"""
from ._lib import LIB_PATH, RTuckerError, lib  # noqa: F401

__all__ = ["LIB_PATH", "RTuckerError", "lib"]
