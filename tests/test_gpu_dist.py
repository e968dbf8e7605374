"""Multi-GPU: the tile-partitioned render over NCCL (one gather) is bit-identical to the 1-GPU image."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ezrt_b200 import api, scenes
    from ezrt_b200 import dist as ezdist
    tris, nodes, eye, cam = scenes.s_bunny()
    scene = api.Scene(tris, nodes, device=rank)
    cfg = api.RenderConfig(width=200, height=120, spp=3, max_bounce=2, mode=api.MODE_DISNEY_SOBOL_P5, eye=tuple(eye), camera_rotate=tuple(cam),
                           env_color=(0.35, 0.45, 0.6))
    img = ezdist.render_partitioned(scene, cfg, rank, world)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out_path, img.cpu().numpy())
    dist.barrier()
    scene.close()
    dist.destroy_process_group()


def test_two_gpu_render_equals_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from ezrt_b200 import api, scenes
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    tris, nodes, eye, cam = scenes.s_bunny()
    scene = api.Scene(tris, nodes)
    cfg = api.RenderConfig(width=200, height=120, spp=3, max_bounce=2, mode=api.MODE_DISNEY_SOBOL_P5, eye=tuple(eye), camera_rotate=tuple(cam),
                           env_color=(0.35, 0.45, 0.6))
    ref = scene.render(cfg)
    scene.close()
    assert got.tobytes() == ref.tobytes()
