#!/usr/bin/env python
"""Generate tests/golden/refhost.npz in the authoring container (needs /root/reference).

Everything in it was computed by THE REFERENCE'S OWN HOST CODE (P5 main.cpp compiled from where it lies,
oracle/ref_host_shim.cpp + oracle/ref_stubs/):
  cache_128x64 / hdr_crc_128x64 : calculateHdrCache (P5/main.cpp:591-683) of scenes.synth_hdr(128, 64)
  <scene>_crc_tris/_crc_nodes   : crc32 of the texture buffers readObj + buildBVHwithSAH (+ main()'s encode loops)
                                  produce for the synthetic scenes' OBJ text (ezrt_b200/scenes.py)
  main_*                        : shapes and crc32s of what the reference's main() uploads for its shipped scene
tests/test_ref_host.py checks the product's host pipeline against them (and live against the reference here)."""
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ezrt_b200 import scenes  # noqa: E402
from tests import refhost_binding as refhost  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def reference_built(meshes, tmp):
    files = {}
    calls = []
    for text, m, trans, smooth in meshes:
        if text not in files:
            files[text] = os.path.join(tmp, "mesh%d.obj" % len(files))
            with open(files[text], "w") as f:
                f.write(text)
        calls.append((files[text], m.as_array(), trans, smooth))
    return refhost.build_scene(calls, 8, True)


def main():
    assert refhost.available(), "needs /root/reference"
    d = {}
    hdr = scenes.synth_hdr(128, 64)
    d["hdr_crc_128x64"] = np.uint32(crc(hdr))
    d["cache_128x64"] = refhost.hdr_cache(hdr)
    with tempfile.TemporaryDirectory() as tmp:
        for name, meshes in (("bunny", scenes.bunny_meshes()), ("grid", scenes.grid_meshes(3, 2, 2))):
            tris, nodes = reference_built(meshes, tmp)
            d[name + "_crc_tris"] = np.uint32(crc(tris)); d[name + "_crc_nodes"] = np.uint32(crc(nodes))
            d[name + "_shape"] = np.array([tris.shape[0], nodes.shape[0]])
            print(name, tris.shape, nodes.shape)
    tris, nodes, hdr2k, cache2k = refhost.run_main()
    d["main_shape"] = np.array([tris.shape[0], nodes.shape[0], hdr2k.shape[1], hdr2k.shape[0]])
    d["main_crc"] = np.array([crc(tris), crc(nodes), crc(hdr2k), crc(cache2k)], np.uint32)
    np.savez_compressed(os.path.join(HERE, "refhost.npz"), **d)
    print("main()", d["main_shape"], [hex(int(c)) for c in d["main_crc"]])


if __name__ == "__main__":
    main()
