#!/usr/bin/env python
"""Generate tests/golden/refshader.npz in the authoring container (needs /root/reference).

Every image in it is the output of THE REFERENCE'S OWN SHADER SOURCE: P3/P4/P5 shaders/fshader.fsh,
transpiled from where they lie to C++ (oracle/ref_shader/transpile.py: literal suffixes, swizzle
accessors, qualifiers; statements and expression order untouched) and run per fragment on the CPU with
GLSL's built-ins bound to include/ezrt_math.h (oracle/ref_shader/glsl_emul.h).  The reference cannot
travel to the GPU box, these frames can:
  * tests/test_ref_shader.py (CPU): the hand-written oracle reproduces them bit for bit (and, where
    /root/reference exists, is compared with the transpiled shaders live on more inputs);
  * tests/test_gpu_parity.py::test_reference_shader_golden_frames (GPU): the CUDA path reproduces
    them bit for bit through the C ABI.

Scenes: the reference's own P3 scene (arrays committed in p3_scene.npz), S-bunny and the blob grid
(ezrt_b200/scenes.py, crc-pinned in synth.npz); environment = scenes.synth_hdr(128, 64) + its cache."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refshader_binding as refshader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
from tests import refshader_cases as cases  # noqa: E402


def main():
    assert refshader.available(), "needs /root/reference"
    hdr, cache = cases.environment()
    d = {}
    for name in cases.SCENES:
        tris, nodes, eye, cam = cases.scene(name)
        for case in cases.CASES:
            key, mode, mb, lin, first, spp = case
            fb = d["%s_m3" % name].copy() if first else None
            d["%s_%s" % (name, key)] = refshader.render(tris, nodes, cases.config(case, eye, cam), hdr, cache, hdr_linear=lin, framebuffer=fb)
        print(name, "done")
    from tests.test_ref_shader import _hdr_frame
    d["pass3_out"] = refshader.pass3(_hdr_frame())  # shaders/pass3.fsh on a fixed HDR frame
    np.savez_compressed(os.path.join(HERE, "refshader.npz"), **d)


if __name__ == "__main__":
    main()
