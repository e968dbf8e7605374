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


class FramebufferGather:
    """Preallocated buffers for the per-step gather: a padded send buffer per rank, the receive list
    and the assembled image on rank `dst`.  One collective per call: torch.distributed.gather."""

    def __init__(self, width, height, channels, rank, world, device, group=None, dst=0):
        import torch
        self.w, self.h, self.c, self.rank, self.world, self.group, self.dst = width, height, channels, rank, world, group, dst
        self.sizes = part_sizes(width, height, world)
        n_max = max(self.sizes) * channels
        self.send = torch.zeros(n_max, dtype=torch.float32, device=device)
        self.parts = [torch.empty_like(self.send) for _ in range(world)] if (rank == dst and world > 1) else None
        self.full = torch.zeros(height * width * channels, dtype=torch.float32, device=device) if rank == dst else None

    def __call__(self, local):
        import torch
        import torch.distributed as dist
        if self.world == 1:  # a single part is already the row-major image
            return local.reshape(-1)[: self.h * self.w * self.c].reshape(self.h, self.w, self.c)
        n = self.sizes[self.rank] * self.c
        self.send[:n].copy_(local.reshape(-1)[:n])
        dist.gather(self.send, self.parts, dst=self.dst, group=self.group)
        if self.rank != self.dst:
            return None
        for r in range(self.world):
            if self.send.is_cuda:
                api.check(api.lib.ezrt_partition_scatter(self.parts[r].data_ptr(), self.full.data_ptr(), self.w, self.h, self.c, r, self.world,
                                                         torch.cuda.current_stream().cuda_stream))
            else:
                src = self.parts[r].numpy()[: self.sizes[r] * self.c]
                api.partition_scatter_host(np.ascontiguousarray(src), self.full.numpy(), self.w, self.h, self.c, r, self.world)
        return self.full.reshape(self.h, self.w, self.c)


def gather_framebuffer(local, width, height, channels, rank, world, group=None, dst=0):
    """One-shot form of FramebufferGather (allocates its buffers)."""
    return FramebufferGather(width, height, channels, rank, world, local.device, group, dst)(local)


def render_partitioned(scene, cfg, rank, world, group=None, d_local=None):
    """display() x spp on this rank's tiles (device), then the single framebuffer gather to rank 0."""
    import torch

    cfg.part_rank, cfg.part_count = rank, world
    n_local = api.partition_pixels(cfg.width, cfg.height, rank, world)
    if d_local is None:
        d_local = torch.zeros(max(1, n_local) * cfg.out_channels, dtype=torch.float32, device="cuda")
    scene.render_device(cfg, d_local, torch.cuda.current_stream())
    return gather_framebuffer(d_local, cfg.width, cfg.height, cfg.out_channels, rank, world, group)
