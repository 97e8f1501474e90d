"""B200-native replacement for the per-pixel sample job of
renaudbedard/raytracing-in-one-weekend (Runtime/Jobs/SampleBatchJob.cs).

  host      stand-in for the C# host: scenes, BVH build, View (librtb_host.so, CPU only)
  plugin    ctypes binding of the C ABI (include/rtb.h) of librtb.so — the sm_100a kernels
  job       SampleBatchJob mirror with the reference's field names + the ScheduleSample batch loop
  sharding  row-tile partition of a frame across GPUs + the one NCCL gather per frame (imported lazily: needs torch)

The directory name is not a Python identifier; import it with
    rtb = importlib.import_module("raytracing-in-one-weekend_b200")
"""
from . import _abi as abi  # noqa: F401
from . import build  # noqa: F401
from . import host  # noqa: F401

__all__ = ["abi", "build", "host"]
try:  # plugin/job need nothing but ctypes; kept in a try so a half-built tree still imports `host`
    from . import plugin  # noqa: F401
    from . import job  # noqa: F401

    __all__ += ["plugin", "job"]
except ImportError:  # pragma: no cover
    pass
