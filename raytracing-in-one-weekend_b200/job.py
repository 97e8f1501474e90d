"""Host-side mirrors of the reference's job structs for this path, with the reference's own
field names, so tests and the stand-in host read like Raytracer.ScheduleSample
(Unity/Raytracer.cs:671-736).

    SampleBatchJob   Runtime/Jobs/SampleBatchJob.cs:17-51   -> rtb_sample_batch (the hot path)
    run_progressive  Raytracer.ScheduleSample batch loop (Raytracer.cs:602-816, ping-pong :798-802)

In the reference a job is a blittable struct whose public fields are its signature and
`.Schedule(W*H, 1, dep).Complete()` runs it; here `.Run(ctx)` makes the one blocking plugin
call that replaces it.  Nothing in this module computes pixels on the CPU.
"""
from dataclasses import dataclass, field

import numpy as np

from . import _abi as abi
from . import host as _host
from . import plugin as _plugin


@dataclass
class SampleBatchJob:
    # uniforms (SampleBatchJob.cs:25-39); PerlinNoise/BlueNoise/StbNoise are outside the
    # supported path (NoiseColor is White -> Philox, BASELINE.json north_star)
    Size: tuple = (0.0, 0.0)
    SliceOffset: int = 0
    SliceDivider: int = 1
    Seed: int = 1
    View: abi.View = field(default_factory=abi.View)
    Environment: abi.Environment = field(default_factory=abi.Environment)
    SampleCountRange: tuple = (1, 1)
    TraceDepth: int = 50
    SubPixelJitter: bool = True
    SampleCountWeightExtrema: tuple = (0.0, 0.0)
    # buffers (SampleBatchJob.cs:41-51)
    Buffers: _plugin.HostBuffers = None
    CancellationToken: np.ndarray = None  # np.uint8[1], the NativeReference<bool>
    # rtb extension: row-tile sharding
    RowBegin: int = 0
    RowEnd: int = 0

    def params(self):
        p = abi.BatchParams()
        p.size[0], p.size[1] = float(self.Size[0]), float(self.Size[1])
        p.slice_offset, p.slice_divider = int(self.SliceOffset), int(self.SliceDivider)
        p.seed = int(self.Seed)
        p.view = self.View
        p.environment = self.Environment
        p.sample_count_range[0], p.sample_count_range[1] = int(self.SampleCountRange[0]), int(self.SampleCountRange[1])
        p.trace_depth = int(self.TraceDepth)
        p.sub_pixel_jitter = 1 if self.SubPixelJitter else 0
        p.sample_count_weight_extrema[0] = float(self.SampleCountWeightExtrema[0])
        p.sample_count_weight_extrema[1] = float(self.SampleCountWeightExtrema[1])
        p.row_begin, p.row_end = int(self.RowBegin), int(self.RowEnd)
        return p

    def Run(self, ctx):
        """sampleBatchJob.Schedule(W*H, 1, dep).Complete() (Raytracer.cs:730-736)."""
        ctx.sample_batch(self.params(), self.Buffers, cancel=self.CancellationToken)
        return self.Buffers


def reduce_metrics_host(buffers):
    """ReduceMetricsJob (ReduceMetricsJob.cs:22-45) on HOST arrays — O(pixels) host plumbing the
    reference also runs on the CPU; the device version is Context.reduce_metrics_device."""
    sc = buffers.out_color[:, 3].astype(np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = buffers.out_weight / sc.astype(np.float32)
    finite = w[~np.isnan(w)]
    rays = int(buffers.diagnostics["ray_count"].astype(np.int64).sum()) if buffers.diagnostics is not None else 0
    return {
        "TotalRayCount": rays,
        "TotalSamples": int(sc.astype(np.int64).sum()),
        "SampleCountWeightExtrema": (float(finite.min()) if len(finite) else float("inf"),
                                     float(finite.max()) if len(finite) else float("-inf")),
        "SampleCountExtrema": (int(sc.min()), int(sc.max())),
    }


def run_progressive(ctx, scene, width, height, batches, spp_per_batch, trace_depth, aperture=None, interlacing=1,
                    spp_max=None, seed0=1):
    """The batch loop of Raytracer.ScheduleSample for `batches` batches: builds the View, zeroes
    the accumulators on the first batch, runs one SampleBatchJob per batch with
    frameSeed = batch + 1 (:660), feeds SampleCountWeightExtrema back (:529,706), interlaces
    rows through Tools.SpaceFillingSeries (:650-661) and ping-pongs the buffers (:798-802).
    Returns the HostBuffers whose out_* hold the final accumulation."""
    buffers = _plugin.HostBuffers(width, height)
    view, _ = _host.view_for(scene, width, height, aperture)
    series = _host.space_filling_series(interlacing) if interlacing > 1 else [0]
    extrema = (0.0, 0.0)
    for b in range(batches):
        job = SampleBatchJob(
            Size=(width, height), SliceOffset=series[b % len(series)], SliceDivider=interlacing, Seed=seed0 + b,
            View=view, Environment=scene.environment,
            SampleCountRange=(spp_per_batch, spp_per_batch if spp_max is None else spp_max),
            TraceDepth=trace_depth, SubPixelJitter=True, SampleCountWeightExtrema=extrema, Buffers=buffers,
        )
        if interlacing > 1:  # unprocessed rows carry over (Raytracer.cs:717-726)
            buffers.out_color[:] = buffers.in_color
            buffers.out_weight[:] = buffers.in_weight
            buffers.out_normal[:] = buffers.in_normal
            buffers.out_albedo[:] = buffers.in_albedo
        job.Run(ctx)
        m = reduce_metrics_host(buffers)
        extrema = m["SampleCountWeightExtrema"]
        if b != batches - 1:
            buffers.swap()
    return buffers
