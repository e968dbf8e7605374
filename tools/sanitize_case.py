#!/usr/bin/env python
"""One small render per integrator mode x traversal policy x pipeline through the C ABI (host buffers), for
compute-sanitizer (tools/sanitize.sh).  No torch: numpy + ctypes only, so the sanitizer only sees this repo's kernels.
Every image is also compared with the first policy's (they must be bit-identical), so a run doubles as a sanity check."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezrt_b200 import api, scenes  # noqa: E402


def main():
    w, h = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "96x54").split("x"))
    tris, nodes, eye, cam = scenes.s_bunny()
    hdr = scenes.synth_hdr(64, 32)
    cache = api.hdr_cache(hdr)
    scene = api.Scene(tris, nodes, hdr, cache, device=0)
    n = 0
    for mode, bounces in ((0, 2), (1, 3), (2, 2), (3, 2)):
        first = None
        for traverse in (api.TRAVERSE_ACCEL, api.TRAVERSE_PRUNED, api.TRAVERSE_REFERENCE):
            for pipeline in (api.PIPELINE_WAVEFRONT, api.PIPELINE_MEGAKERNEL):
                if pipeline == api.PIPELINE_MEGAKERNEL and traverse == api.TRAVERSE_ACCEL:
                    continue
                cfg = api.RenderConfig(width=w, height=h, spp=2, max_bounce=bounces, mode=mode, eye=tuple(eye), camera_rotate=tuple(cam),
                                       traverse=traverse, pipeline=pipeline)
                img = scene.render(cfg).copy()
                if first is None:
                    first = img
                assert img.tobytes() == first.tobytes(), "mode %d traverse %d pipeline %d differs" % (mode, traverse, pipeline)
                n += 1
    # partitioned render (compact output + scatter path) and a continued accumulation
    cfg = api.RenderConfig(width=w, height=h, spp=1, first_frame=2, max_bounce=2, mode=2, eye=tuple(eye), camera_rotate=tuple(cam),
                           part_rank=1, part_count=3)
    scene.render(cfg)
    o = np.random.default_rng(1).normal(size=(4096, 3)).astype(np.float32)
    d = np.random.default_rng(2).normal(size=(4096, 3)).astype(np.float32)
    d[::97, 0] = 0.0
    for traverse in (api.TRAVERSE_ACCEL, api.TRAVERSE_PRUNED, api.TRAVERSE_REFERENCE):
        scene.trace_rays(o, d, traverse=traverse)
        scene.trace_rays(o, d, traverse=traverse, any_hit=True)
    scene.close()
    # the accel policy's deferred lane (side stream): 30 coincident triangles facing the camera defer every ray that hits them
    wall = np.zeros((30, 36), np.float32)
    e = np.asarray(eye, np.float64)
    nrm = e / np.linalg.norm(e)
    u = np.cross([0.0, 1.0, 0.0], nrm)
    u /= np.linalg.norm(u)
    v = np.cross(nrm, u)
    c, hh = 0.55 * e, 0.25 * np.linalg.norm(e)
    wall[:, :9] = np.concatenate([c - 2 * hh * u - hh * v, c + 2 * hh * u - hh * v, c + 3 * hh * v]).astype(np.float32)
    wall[:, 9:18] = np.tile(nrm.astype(np.float32), 3)
    wall[:, 21:24] = 0.7
    wall[:, 28] = 0.5
    tl = api.TriangleList()
    tl.append_encoded(np.concatenate([np.asarray(tris, np.float32).reshape(-1, 36), wall]))
    t2, n2 = tl.build_bvh(8, api.BVH_SAH_FAST)
    scene = api.Scene(t2, n2, hdr, cache, device=0)
    for mode in (2, 3):
        first = None
        for traverse in (api.TRAVERSE_ACCEL, api.TRAVERSE_PRUNED):
            cfg = api.RenderConfig(width=w, height=h, spp=2, max_bounce=2, mode=mode, eye=tuple(eye), camera_rotate=tuple(cam), traverse=traverse)
            img = scene.render(cfg).copy()
            if first is None:
                first = img
                assert scene.counters().deferred_rays > 0
            assert img.tobytes() == first.tobytes(), "deferred lane: mode %d traverse %d differs" % (mode, traverse)
            n += 1
    scene.close()
    print("sanitize_case: %d renders + trace_rays done" % n)


if __name__ == "__main__":
    main()
