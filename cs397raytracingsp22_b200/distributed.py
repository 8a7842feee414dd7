"""One process per GPU: shard the frame, render shards into exact integer accumulators, sum them once.

The reference's only parallelism is a rayon parallel-for over image rows (tracing.rs:228).  Here every
(pixel, sample) is independent and the RNG is keyed on (pixel, sample, bounce), so the frame can be cut by
SAMPLE RANGE (every rank renders all pixels for spp/N sample indices; perfectly balanced) or by interleaved
TILES.  Each rank adds its samples into a W*H*4 int64 accumulator (fixed point, 2^-30); integer sums are exact
and order-free, so one `all_reduce(SUM)` / `reduce(SUM)` over NCCL gives bit-for-bit the accumulator a single GPU
would have produced.  That collective is the only communication per frame; there is no data-path exchange.

torch is plumbing here (device memory, streams, torch.distributed); the rendering is librt_b200.so.
"""
from __future__ import annotations

import math

from . import _ffi


def sample_range(rank: int, world: int, begin: int, end: int) -> tuple[int, int]:
    """Contiguous, as-even-as-possible split of [begin, end) — mirrors plan_shard() in csrc/rt_api.cu."""
    n = end - begin
    base, rem = divmod(n, world)
    b = begin + rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


def tile_stride(world: int) -> int:
    """The smallest odd number >= 3 coprime to `world` - mirrors plan_shard() in csrc/rt_api.cu."""
    stride = 3
    while math.gcd(stride, world) != 1:
        stride += 2
    return stride


def tiles_of_rank(rank: int, world: int, width: int, height: int, tile: int = 64) -> list[tuple[int, int]]:
    """Tile (tx, ty) belongs to rank (tx + stride * ty) % world, stride = tile_stride(world): diagonals, so that
    vertical structures of the scene do not line up with one rank's tiles.  Row-major order - mirrors plan_shard()
    in csrc/rt_api.cu."""
    tx, ty = (width + tile - 1) // tile, (height + tile - 1) // tile
    stride = tile_stride(world)
    return [(x, y) for y in range(ty) for x in range(tx) if (x + stride * y) % world == rank]


def shard_opts(rank: int, world: int, seed: int, mode: str = "samples", tile: int = 64, sample_begin: int = 0,
               sample_end: int = 0, wavefront: int = 0, flags: int = 0, engine: int = 0, ray_sort: int = 0,
               work_order: int = 0, blocks_per_sm: int = 0) -> _ffi.rt_render_opts:
    o = _ffi.rt_render_opts()
    o.seed = seed
    o.shard_mode = {"all": _ffi.RT_SHARD_ALL, "samples": _ffi.RT_SHARD_SAMPLES, "tiles": _ffi.RT_SHARD_TILES}[mode]
    o.shard_rank, o.shard_count = rank, world
    o.tile_size = tile
    o.sample_begin, o.sample_end = sample_begin, sample_end
    o.wavefront = wavefront
    o.flags = flags
    o.engine, o.ray_sort, o.work_order, o.blocks_per_sm = engine, ray_sort, work_order, blocks_per_sm
    return o


def new_accum(width: int, height: int, device):
    import torch
    return torch.zeros(height * width * 4, dtype=torch.int64, device=device)


def reduce_accum(accum, dst: int | None = 0):
    """Sum the per-rank accumulators.  dst=None -> all_reduce (every rank gets the frame)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return accum
    if dst is None:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def render_shard(backend: "_ffi.GpuBackend", cam: "_ffi.rt_camera", opts: "_ffi.rt_render_opts", accum) -> "_ffi.rt_stats":
    """Render this rank's shard into `accum` (a CUDA int64 tensor) on torch's current stream."""
    import torch
    assert accum.is_cuda and accum.dtype == torch.int64
    stream = torch.cuda.current_stream(accum.device).cuda_stream
    return backend.render_accum(cam, opts, accum.data_ptr(), stream)


def resolve(backend: "_ffi.GpuBackend", cam: "_ffi.rt_camera", accum, total_spp: int, want_linear=True, want_rgb8=True):
    """mean + output transform (tracing.rs:241-256) on the device -> (linear f32 HxWx3, rgb8 HxWx3) CUDA tensors."""
    import torch
    h, w = cam.screen_height, cam.screen_width
    lin = torch.empty((h, w, 3), dtype=torch.float32, device=accum.device) if want_linear else None
    rgb = torch.empty((h, w, 3), dtype=torch.uint8, device=accum.device) if want_rgb8 else None
    stream = torch.cuda.current_stream(accum.device).cuda_stream
    backend.resolve(cam, accum.data_ptr(), total_spp, lin.data_ptr() if lin is not None else 0,
                    rgb.data_ptr() if rgb is not None else 0, stream)
    return lin, rgb
