#!/usr/bin/env python
"""Randomised comparison of the oracle with the reference's own (transpiled) shaders: scenes incl. a degenerate triangle
soup, all four modes, 0-4 bounces, both filters, random image shapes / cameras / environment sizes / frame offsets.
Authoring-container tool (needs /root/reference).  Round 1: 71,342 cases in 150 s, 0 mismatches.  usage: fuzz_oracle_vs_refshader.py [seconds]"""
import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezrt_b200 import api, scenes
from tests import oracle_binding as oracle, refshader_binding as refshader
from tests.test_gpu_parity import _soup
rng = np.random.default_rng(2026)
scn = {"bunny": scenes.s_bunny()[:2], "grid": scenes.s_grid(2,2,1)[:2]}
tl = api.TriangleList(); tl.append_encoded(_soup(1500, 9)); scn["soup"] = tl.build_bvh(5)
t0=time.time(); n=0; bad=0
while time.time()-t0 < (float(sys.argv[1]) if len(sys.argv) > 1 else 150):
    name = rng.choice(list(scn)); tris,nodes = scn[name]
    mode = int(rng.integers(0,4)); mb = int(rng.integers(0,5)); lin = bool(rng.integers(0,2))
    w,h = int(rng.integers(3,40)), int(rng.integers(3,30)); spp=int(rng.integers(1,4)); ff=int(rng.integers(0,2000)) if rng.uniform()<0.5 else 0
    hw = int(rng.choice([8,32,64])); hdr = scenes.synth_hdr(hw, hw//2, seed=int(rng.integers(1,9))); cache = api.hdr_cache(hdr)
    eye,cam = api.camera_orbit(float(rng.uniform(-180,180)), float(rng.uniform(-89,89)), float(rng.uniform(0.3,9)))
    cfg = api.RenderConfig(width=w,height=h,spp=spp,max_bounce=mb,mode=mode,eye=tuple(eye),camera_rotate=tuple(cam),first_frame=ff,traverse=int(rng.choice([1,2])))
    fb0 = rng.uniform(0,3,(h,w,3)).astype(np.float32) if ff else None
    a,_ = oracle.render(tris,nodes,cfg,hdr=hdr,hdr_cache=cache,hdr_linear=lin,framebuffer=None if fb0 is None else fb0.copy())
    b = refshader.render(tris,nodes,cfg,hdr,cache,hdr_linear=lin,framebuffer=None if fb0 is None else fb0.copy())
    same = ((a.view(np.uint32)==b.view(np.uint32)) | (np.isnan(a)&np.isnan(b))).all()
    n+=1
    if not same:
        bad+=1; print("MISMATCH", name, mode, mb, lin, w,h,spp,ff, float(np.nanmax(np.abs(a-b))))
print("cases", n, "mismatches", bad)
