"""The oracle against THE REFERENCE'S OWN SHADER SOURCE.

oracle/ref_shader/ transpiles P3/P4/P5 shaders/fshader.fsh (read where they lie under /root/reference,
never copied) to C++ and runs them per fragment on the CPU; GLSL's built-ins -- which the GLSL
specification leaves bit-unspecified -- are bound to include/ezrt_math.h.  Everything else (statements,
expression order, control flow, constants, the Sobol table, the RNG seeding) is the reference's text.

  * frames rendered that way are committed in tests/golden/refshader.npz and the hand-written oracle
    must reproduce them bit for bit (runs everywhere, GPU box included);
  * where /root/reference exists the oracle is also compared live, on more inputs."""
import numpy as np
import pytest

from ezrt_b200 import api, scenes
from tests import refshader_binding as refshader
from tests import refshader_cases as cases

needs_reference = pytest.mark.skipif(not refshader.available(), reason="/root/reference (shader sources) not present")


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.parametrize("name", cases.SCENES)
def test_oracle_reproduces_reference_shader_frames(oracle, name):
    g = cases.load()
    hdr, cache = cases.environment()
    tris, nodes, eye, cam = cases.scene(name)
    for case in cases.CASES:
        key, mode, mb, lin, first, spp = case
        want = g["%s_%s" % (name, key)]
        assert np.isfinite(want).all() and float(want.mean()) > 0.05, "golden frame is not a trivial image"
        for traverse in (api.TRAVERSE_REFERENCE, api.TRAVERSE_PRUNED):
            fb = g["%s_m3" % name].copy() if first else None
            got, _ = oracle.render(tris, nodes, cases.config(case, eye, cam, traverse=traverse), hdr=hdr, hdr_cache=cache, hdr_linear=lin,
                                   framebuffer=fb)
            assert same_bits(got, want), "%s %s traverse %d" % (name, key, traverse)


def test_golden_inputs_are_the_committed_ones():
    """refshader.npz was rendered from exactly these arrays (synth.npz pins their crc32)."""
    import os
    import zlib
    s = np.load(os.path.join(cases.GOLDEN, "synth.npz"))
    hdr, cache = cases.environment()
    crc = lambda a: zlib.crc32(np.ascontiguousarray(a).tobytes())
    assert crc(hdr) == int(s["hdr_crc"]) and crc(cache) == int(s["cache_crc"])
    tris, nodes, _, _ = cases.scene("bunny")
    assert crc(tris) == int(s["bunny_crc_tris"]) and crc(nodes) == int(s["bunny_crc_nodes"])
    tris, nodes, _, _ = cases.scene("grid")
    assert crc(tris) == int(s["grid_crc_tris"]) and crc(nodes) == int(s["grid_crc_nodes"])


@needs_reference
def test_golden_frames_are_current():
    """the committed frames are what the transpiled shaders produce today"""
    g = cases.load()
    hdr, cache = cases.environment()
    tris, nodes, eye, cam = cases.scene("bunny")
    for case in cases.CASES[:4]:
        key, mode, mb, lin, first, spp = case
        got = refshader.render(tris, nodes, cases.config(case, eye, cam), hdr, cache, hdr_linear=lin)
        assert same_bits(got, g["bunny_" + key]), key


@needs_reference
@pytest.mark.parametrize("mode,bounces", [(0, 2), (0, 3), (1, 4), (1, 1), (2, 2), (2, 4), (3, 2), (3, 3)])
@pytest.mark.parametrize("linear", [False, True])
def test_oracle_equals_transpiled_shader_live(oracle, grid_scene, mode, bounces, linear):
    """other image shape, camera, environment size, filter and bounce counts than the committed frames"""
    tris, nodes, _, _ = grid_scene
    hdr = scenes.synth_hdr(64, 32)
    cache = api.hdr_cache(hdr)
    eye, cam = api.camera_orbit(37.0, 12.0, 5.5)
    cfg = api.RenderConfig(width=37, height=23, spp=4, max_bounce=bounces, mode=mode, eye=tuple(eye), camera_rotate=tuple(cam),
                           traverse=api.TRAVERSE_REFERENCE)
    want = refshader.render(tris, nodes, cfg, hdr, cache, hdr_linear=linear)
    got, _ = oracle.render(tris, nodes, cfg, hdr=hdr, hdr_cache=cache, hdr_linear=linear)
    assert float(want.mean()) > 0.01
    assert same_bits(got, want)


@needs_reference
def test_oracle_equals_transpiled_shader_inside_a_box(oracle):
    """camera inside closed geometry (isInside hits, every path terminates on geometry or an emitter)"""
    tl = api.TriangleList()
    tl.read_obj_text(scenes.box_obj(), api.Material(baseColor=(0.7, 0.6, 0.5), roughness=0.4, metallic=0.3), api.transform_matrix((0, 0, 0), (0, 0, 0), (3, 3, 3)), False)
    tl.read_obj_text(scenes.sphere_obj(2), api.Material(baseColor=(1, 1, 1), emissive=(9, 8, 7)), api.transform_matrix((0, 0, 0), (0, 0.6, 0), (0.5, 0.5, 0.5)), True)
    tris, nodes = tl.build_bvh(4, api.BVH_SAH_LITERAL)
    hdr, cache = cases.environment()
    eye, cam = api.camera_orbit(20.0, -10.0, 1.2)
    for mode, mb, lin in ((0, 3, False), (1, 4, False), (2, 3, True), (3, 3, True)):
        cfg = api.RenderConfig(width=32, height=24, spp=3, max_bounce=mb, mode=mode, eye=tuple(eye), camera_rotate=tuple(cam),
                               traverse=api.TRAVERSE_REFERENCE)
        want = refshader.render(tris, nodes, cfg, hdr, cache, hdr_linear=lin)
        got, _ = oracle.render(tris, nodes, cfg, hdr=hdr, hdr_cache=cache, hdr_linear=lin)
        assert same_bits(got, want), mode


def _hdr_frame(seed=5, h=24, w=40, c=3):
    rng = np.random.default_rng(seed)
    fb = (rng.uniform(0, 1, (h, w, c)) ** 4 * 40.0).astype(np.float32)  # HDR range, many dark texels
    fb[0, 0, :3] = 0.0
    fb[0, 1, :3] = (1e-30, 1.0, 3e4)
    return fb


@needs_reference
@pytest.mark.parametrize("channels", [3, 4])
def test_tonemap_equals_transpiled_pass3_shader(oracle, channels):
    """shaders/pass3.fsh (identical in parts 3, 4, 5): toneMapping(c, 1.5) then pow(c, 1/2.2)"""
    fb = _hdr_frame(c=channels)
    want = refshader.pass3(fb)
    assert same_bits(oracle.tonemap(fb), want)


def test_tonemap_reproduces_reference_pass3_golden(oracle):
    """runs everywhere: pass3 output of the reference shader for a fixed frame, committed in refshader.npz"""
    g = cases.load()
    assert same_bits(oracle.tonemap(_hdr_frame()), g["pass3_out"])


@needs_reference
def test_oracle_equals_transpiled_shader_on_degenerate_soup(oracle):
    """zero-area triangles (NaN normals), slivers, duplicated vertices, axis-aligned coordinates: the NaN / tie
    behaviour of the oracle is the shader's, statement by statement"""
    from tests.test_gpu_parity import _soup
    tl = api.TriangleList()
    tl.append_encoded(_soup(3000, 4))
    tris, nodes = tl.build_bvh(8)
    hdr, cache = cases.environment()
    eye, cam = api.camera_orbit(30.0, 20.0, 6.0)
    for mode, mb, lin in ((0, 3, False), (1, 3, False), (2, 2, True), (3, 2, True)):
        cfg = api.RenderConfig(width=64, height=48, spp=2, max_bounce=mb, mode=mode, eye=tuple(eye), camera_rotate=tuple(cam),
                               traverse=api.TRAVERSE_REFERENCE)
        want = refshader.render(tris, nodes, cfg, hdr, cache, hdr_linear=lin)
        got, _ = oracle.render(tris, nodes, cfg, hdr=hdr, hdr_cache=cache, hdr_linear=lin)
        assert same_bits(got, want), mode


@needs_reference
def test_oracle_equals_transpiled_shader_on_the_reference_p5_scene(oracle):
    """the reference's own shipped P5 set-up, end to end: its main() (compiled here, tests/refhost_binding.py) supplies
    the texture buffers, the HDR map and the sampling cache; its shader (transpiled) renders them with its own camera
    (P5/main.cpp:796-798) at a reduced size; the oracle must agree bit for bit"""
    from tests import refhost_binding as refhost
    if not refhost.available():
        pytest.skip("reference host sources not present")
    tris, nodes, hdr, cache = refhost.run_main()
    eye, cam = api.camera_orbit(90.0, 10.0, 2.0)
    cfg = api.RenderConfig(width=64, height=64, spp=2, max_bounce=2, mode=api.MODE_DISNEY_IS_MIS_P5, eye=tuple(eye), camera_rotate=tuple(cam),
                           traverse=api.TRAVERSE_REFERENCE)
    want = refshader.render(tris, nodes, cfg, hdr, cache, hdr_linear=True)
    got, c = oracle.render(tris, nodes, cfg, hdr=hdr, hdr_cache=cache, hdr_linear=True)
    assert c["rays_shadow"] > 0 and float(want.mean()) > 0.01
    assert same_bits(got, want)
