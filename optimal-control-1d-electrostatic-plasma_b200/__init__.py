"""B200-native 1D electrostatic PIC step behind the reference's `PIC` interface.

    from pic_b200 import PIC, BatchedPIC, Engine

The CUDA library (lib/libpic_b200.so, sources under csrc/) is loaded on first use; there is no CPU fallback.
"""
from . import _lib
from ._lib import PicError, PicDeviceError
from .engine import Engine, DeviceArray
from .pic import PIC
from .batched import BatchedPIC, shard_range
from .sharded import ShardedPIC
from .dist import BumpOnTail, TwoStream
from .actuator import E_field

__all__ = ["PIC", "BatchedPIC", "ShardedPIC", "shard_range", "Engine", "DeviceArray", "PicError", "PicDeviceError", "BumpOnTail", "TwoStream", "E_field"]
