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
    print("sanitize_case: %d renders + trace_rays done" % n)


if __name__ == "__main__":
    main()
