"""Multi-GPU rendering: image-tile partition across ranks + ONE gather of the final framebuffer.

The path shards naturally: every pixel-sample is independent and its RNG/Sobol
state depends only on (px, py, frame), so any partition gives bit-identical pixels.  Each rank
holds a full scene replica, renders the 16x16 tiles (tx+ty) % world == rank for all spp, and the
compact per-rank buffers are gathered to rank 0 (NCCL over NVLink on GPUs, gloo on CPU) and
de-interleaved there.  There is no per-bounce or per-sample communication.
"""
import numpy as np

from . import api


def part_sizes(width, height, world):
    return [api.partition_pixels(width, height, r, world) for r in range(world)]


def gather_framebuffer(local, width, height, channels, rank, world, group=None, dst=0):
    """Gather the compact per-rank buffers (torch tensors, CPU or CUDA, float32, n_local*channels)
    to rank `dst` and scatter them into a full [height, width, channels] image there.
    Returns the image on `dst`, None elsewhere.  One collective: torch.distributed.gather."""
    import torch
    import torch.distributed as dist

    if world == 1:  # a single part is already the row-major image
        return local.reshape(-1)[: height * width * channels].reshape(height, width, channels)
    sizes = part_sizes(width, height, world)
    n_max = max(sizes) * channels
    buf = local.new_zeros(n_max)
    buf[: sizes[rank] * channels] = local.reshape(-1)[: sizes[rank] * channels]
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst, group=group)
    if rank != dst:
        return None
    full = local.new_zeros(height * width * channels)
    for r in range(world):
        if local.is_cuda:
            api.check(api.lib.ezrt_partition_scatter(parts[r].data_ptr(), full.data_ptr(), width, height, channels, r, world,
                                                     torch.cuda.current_stream().cuda_stream))
        else:
            src = parts[r].numpy()[: sizes[r] * channels]
            api.partition_scatter_host(np.ascontiguousarray(src), full.numpy(), width, height, channels, r, world)
    return full.reshape(height, width, channels)


def render_partitioned(scene, cfg, rank, world, group=None, d_local=None):
    """display() x spp on this rank's tiles (device), then the single framebuffer gather to rank 0."""
    import torch

    cfg.part_rank, cfg.part_count = rank, world
    n_local = api.partition_pixels(cfg.width, cfg.height, rank, world)
    if d_local is None:
        d_local = torch.zeros(max(1, n_local) * cfg.out_channels, dtype=torch.float32, device="cuda")
    scene.render_device(cfg, d_local, torch.cuda.current_stream())
    return gather_framebuffer(d_local, cfg.width, cfg.height, cfg.out_channels, rank, world, group)
