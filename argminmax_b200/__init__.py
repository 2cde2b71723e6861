"""argminmax_b200 -- B200-native (sm_100a) fused argmin+argmax behind the reference's
ArgMinMax / NaNArgMinMax surface.  See include/argminmax_b200.h for the C ABI that is the
actual drop-in boundary; this package is the thin host-side mirror used by the tests and
bench.py."""
from ._ffi import (AmmError, AmmRecord, EmptyInputError, DTYPES, DTYPE_SIZE, FLOAT_DTYPES,
                   IGNORE_NAN, RETURN_NAN, lib, lib_path)
from .device_slice import (DeviceSlice, HostSlice, ShardedDeviceSlice, Workspace,
                           combine_records, default_workspace)

from .distributed import (DistributedShardedSlice, combine_across_ranks, exchange_records,
                          shard_bounds)

__all__ = [
    "DistributedShardedSlice", "combine_across_ranks", "exchange_records", "shard_bounds",
    "AmmError", "AmmRecord", "EmptyInputError", "DTYPES", "DTYPE_SIZE", "FLOAT_DTYPES",
    "IGNORE_NAN", "RETURN_NAN", "lib", "lib_path", "DeviceSlice", "HostSlice",
    "ShardedDeviceSlice", "Workspace", "combine_records", "default_workspace",
]
