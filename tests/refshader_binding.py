"""ctypes binding of oracle/_ref/libezrt_refshader.so: the reference's own fragment shaders (P3/P4/P5
shaders/fshader.fsh), transpiled to C++ from where they lie under /root/reference and run on the CPU
(oracle/ref_shader/).  TEST INFRASTRUCTURE; exists only where /root/reference does (the authoring
container) -- `available()` is False on the GPU box, where the committed tests/golden/refshader.npz
frames stand in for it."""
import ctypes as C
import os

import numpy as np

from ezrt_b200 import build as _build
from ezrt_b200._lib import RenderParams

_fp = C.POINTER(C.c_float)
_lib = None


def _load():
    global _lib
    if _lib is None:
        so = _build.build_reference_shaders()
        if so is None or not os.path.exists(so):
            return None
        _lib = C.CDLL(so)
        _lib.refshader_render.restype = C.c_int
        _lib.refshader_render.argtypes = [_fp, C.c_int, _fp, C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, C.POINTER(RenderParams), _fp, C.c_int]
    return _lib


def available():
    return _load() is not None


def render(tris, nodes, cfg, hdr, hdr_cache=None, hdr_linear=True, framebuffer=None, threads=0):
    """cfg.spp frames of the shader that implements cfg.mode (the hdr map is mandatory: the shaders always
    sample it).  Same framebuffer convention as oracle_binding.render."""
    lib = _load()
    f = lambda a: None if a is None else a.ctypes.data_as(_fp)
    tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 36)
    nodes = np.ascontiguousarray(nodes, np.float32).reshape(-1, 12)
    hdr = np.ascontiguousarray(hdr, np.float32)
    hdr_cache = None if hdr_cache is None else np.ascontiguousarray(hdr_cache, np.float32)
    fb = np.zeros((cfg.height, cfg.width, cfg.out_channels), np.float32) if framebuffer is None else framebuffer
    p = cfg.to_struct()
    rc = lib.refshader_render(f(tris), tris.shape[0], f(nodes), nodes.shape[0], f(hdr), f(hdr_cache), hdr.shape[1], hdr.shape[0],
                              int(bool(hdr_linear)), C.byref(p), f(fb), int(threads))
    if rc != 0:
        raise RuntimeError("refshader_render failed (%d)" % rc)
    return fb


def pass3(image):
    """shaders/pass3.fsh (tone mapping + gamma) on an [H,W,C>=3] float image -> [H,W,3]"""
    lib = _load()
    lib.refshader_pass3.restype = C.c_int
    lib.refshader_pass3.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp]
    image = np.ascontiguousarray(image, np.float32)
    h, w, c = image.shape
    out = np.zeros((h, w, 3), np.float32)
    assert lib.refshader_pass3(image.ctypes.data_as(_fp), c, w, h, out.ctypes.data_as(_fp)) == 0
    return out
