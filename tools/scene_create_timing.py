"""Stage timing of ezrt_scene_create (env EZRT_VERBOSE=1) on the bench scenes.  Usage: python tools/scene_create_timing.py [c3 c2 ...]"""
import os, sys, time
os.environ.setdefault("EZRT_VERBOSE", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from ezrt_b200 import api

for name in (sys.argv[1:] or ["c3"]):
    wl = bench.build_workload(name)
    for rep in range(2):
        t0 = time.time()
        sc = api.Scene(wl["tris"], wl["nodes"], wl.get("hdr"), wl.get("cache"), device=0)
        print(f"== {name} rep {rep}: {len(wl['tris'])} triangles, ezrt_scene_create {1e3 * (time.time() - t0):.1f} ms", file=sys.stderr, flush=True)
        del sc

    import numpy as np
    for where in ("device", "host"):
        for rep in range(2):
            links, boxes, order, ms = api.accel_build(wl["tris"], 4, where)
            print(f"== {name} binary SAH tree on the {where}: {len(links)} nodes, {ms:.1f} ms (upload + build + read-back)", file=sys.stderr, flush=True)
