"""The product's host pipeline against THE REFERENCE'S OWN HOST CODE.

oracle/ref_host_shim.cpp compiles P5/main.cpp and lib/hdrloader.cpp from where they lie under /root/reference
(GL/GLUT: no-op stand-ins that capture uploads; glm: the subset main.cpp uses, oracle/ref_stubs/) and
  * runs the reference's main() -- its shipped scene: teapot.obj (78k triangles) + a 13000-unit floor quad +
    chinese_garden_2k.hdr -- up to glutMainLoop() and captures what it would upload to the GPU;
  * exposes readObj / buildBVH / buildBVHwithSAH / calculateHdrCache for other inputs.
The product (ezrt_b200/csrc/host_scene.cpp) must produce the same BYTES.

Needs /root/reference (skipped on the GPU box); tests/golden/refhost.npz carries what can travel: the
reference-computed sampling cache of the synthetic environment and crc32s of reference-built scenes."""
import os
import zlib

import numpy as np
import pytest

from ezrt_b200 import api, scenes
from tests import refhost_binding as refhost

needs_reference = pytest.mark.skipif(not refhost.available(), reason="/root/reference (host sources) not present")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refhost.npz")


def same_bytes(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


@needs_reference
def test_reference_main_uploads_are_reproduced_byte_for_byte():
    r_tris, r_nodes, r_hdr, r_cache = refhost.run_main()
    assert r_tris.shape == (78188, 36) and r_nodes.shape[1] == 12 and r_hdr.shape == (1024, 2048, 3)
    g = np.load(GOLDEN)  # the committed fingerprints of this run
    assert [crc(r_tris), crc(r_nodes), crc(r_hdr), crc(r_cache)] == [int(c) for c in g["main_crc"]]
    src = refhost.SOURCE_DIR
    for builder in (api.BVH_SAH_LITERAL, api.BVH_SAH_FAST):
        tl = api.TriangleList()  # P5/main.cpp:773-788
        m = api.Material(roughness=0.5, specular=1.0, metallic=1.0, clearcoat=1.0, clearcoatGloss=0.0, baseColor=(1, 0.73, 0.25))
        tl.read_obj(src + "/models/teapot.obj", m, api.transform_matrix((0, 0, 0), (0, -0.5, 0), (0.75, 0.75, 0.75)), True)
        m = api.Material(roughness=0.01, metallic=0.1, specular=1.0, clearcoat=1.0, clearcoatGloss=0.0, baseColor=(1, 1, 1))
        tl.read_obj(src + "/models/quad.obj", m, api.transform_matrix((0, 0, 0), (0, -0.5, 0), (13000.0, 0.01, 13000.0)), False)
        tris, nodes = tl.build_bvh(8, builder)
        assert same_bytes(tris, r_tris), "triangle texture buffer, builder %d" % builder
        assert same_bytes(nodes, r_nodes), "BVH texture buffer, builder %d" % builder
    hdr = api.hdr_load(src + "/HDR/chinese_garden_2k.hdr")
    assert same_bytes(hdr, r_hdr)
    assert same_bytes(api.hdr_cache(hdr), r_cache), "calculateHdrCache"


@needs_reference
def test_reference_p4_main_uploads_are_reproduced_byte_for_byte():
    """P4/main.cpp:689-743: one golden teapot, SAH tree, 1024x512 map"""
    r_tris, r_nodes, r_hdr = refhost.run_main(4)
    src = refhost.source_dir(4)
    tl = api.TriangleList()
    m = api.Material(baseColor=(0.75, 0.7, 0.15), roughness=0.15, metallic=1.0, clearcoat=1.0, subsurface=1.0)
    tl.read_obj(src + "/models/teapot.obj", m, api.transform_matrix((0, 0, 0), (0, -0.4, 0), (1.75, 1.75, 1.75)), True)
    tris, nodes = tl.build_bvh(8, api.BVH_SAH_FAST)
    assert same_bytes(tris, r_tris) and same_bytes(nodes, r_nodes)
    assert same_bytes(api.hdr_load(src + "/HDR/peppermint_powerplant_4k.hdr"), r_hdr)


@needs_reference
def test_reference_p3_main_uploads_equal_the_committed_p3_scene(tmp_path):
    """P3/main.cpp:688-715 (Stanford bunny + floor + emissive sphere).  Its main() opens ./HDR/sunset.hdr, which the
    reference does not ship: it is run in a directory whose HDR/sunset.hdr points at the map P3 does ship.
    tests/golden/p3_scene.npz (the scene the golden frames are rendered from) holds the same geometry and tree;
    only the Disney parameters differ, which P3's Material defaults to 0 (P3/main.cpp:28-43) and its shader never reads."""
    src = refhost.source_dir(3)
    os.symlink(src + "/models", tmp_path / "models")
    os.symlink(src + "/shaders", tmp_path / "shaders")
    os.mkdir(tmp_path / "HDR")
    os.symlink(src + "/HDR/circus_arena_4k.hdr", tmp_path / "HDR" / "sunset.hdr")
    r_tris, r_nodes, r_hdr = refhost.run_main(3, str(tmp_path))
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "p3_scene.npz"))
    assert same_bytes(g["nodes"], r_nodes)
    assert same_bytes(g["tris"][:, :24], r_tris[:, :24])  # positions, normals, emissive, baseColor
    p3_defaults = dict(specular=0.0, roughness=0.0, sheenTint=0.0, clearcoatGloss=0.0)
    tl = api.TriangleList()
    tl.read_obj(src + "/models/Stanford Bunny.obj", api.Material(baseColor=(1, 1, 1), **p3_defaults),
                api.transform_matrix((0, 0, 0), (0.3, -1.6, 0), (1.5, 1.5, 1.5)), True)
    tl.read_obj(src + "/models/quad.obj", api.Material(baseColor=(0.725, 0.71, 0.68), **p3_defaults),
                api.transform_matrix((0, 0, 0), (0, -1.4, 0), (18.83, 0.01, 18.83)), False)
    tl.read_obj(src + "/models/sphere.obj", api.Material(baseColor=(1, 1, 1), emissive=(30, 20, 10), **p3_defaults),
                api.transform_matrix((0, 0, 0), (0.0, 0.9, 0.0), (1, 1, 1)), False)
    tris, nodes = tl.build_bvh(8, api.BVH_SAH_LITERAL)
    assert same_bytes(tris, r_tris) and same_bytes(nodes, r_nodes)
    assert same_bytes(api.hdr_load(src + "/HDR/circus_arena_4k.hdr"), r_hdr)


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


@needs_reference
@pytest.mark.parametrize("leaf_n", [1, 4, 8, 13])
def test_builders_equal_reference_functions_on_synthetic_meshes(tmp_path, leaf_n):
    blob = _write(tmp_path, "blob.obj", scenes.blob_obj(3, 11))
    sphere = _write(tmp_path, "sphere.obj", scenes.sphere_obj(2))
    box = _write(tmp_path, "box.obj", scenes.box_obj())
    rng = np.random.default_rng(leaf_n)
    meshes = []
    for path, smooth in ((blob, True), (sphere, False), (box, False), (blob, False), (sphere, True)):
        rot, tr, sc = rng.uniform(-180, 180, 3), rng.uniform(-2, 2, 3), rng.uniform(0.2, 3, 3)
        mat = api.Material(baseColor=tuple(rng.uniform(0, 1, 3)), emissive=tuple(rng.uniform(0, 5, 3)), roughness=float(rng.uniform()),
                           metallic=float(rng.uniform()), sheen=float(rng.uniform()), clearcoat=float(rng.uniform()))
        trans = api.transform_matrix(tuple(rot), tuple(tr), tuple(sc))
        assert same_bytes(trans, refhost.transform_matrix(rot, tr, sc))  # same restatement of glm on both sides (see glm.hpp)
        meshes.append((path, mat, trans, smooth))
    for sah, builders in ((True, (api.BVH_SAH_LITERAL, api.BVH_SAH_FAST)), (False, (api.BVH_MEDIAN,))):
        r_tris, r_nodes = refhost.build_scene([(p, m.as_array(), t, s) for p, m, t, s in meshes], leaf_n, sah)
        for b in builders:
            tl = api.TriangleList()
            for p, m, t, s in meshes:
                tl.read_obj(p, m, t, s)
            tris, nodes = tl.build_bvh(leaf_n, b)
            assert same_bytes(tris, r_tris), (sah, b)
            assert same_bytes(nodes, r_nodes), (sah, b)


@needs_reference
@pytest.mark.parametrize("w,h", [(128, 64), (64, 32), (96, 40), (16, 8)])
def test_hdr_cache_equals_reference_function(w, h):
    hdr = scenes.synth_hdr(w, h)
    assert same_bytes(api.hdr_cache(hdr), refhost.hdr_cache(hdr))


@needs_reference
def test_golden_is_current():
    g = np.load(GOLDEN)
    hdr = scenes.synth_hdr(128, 64)
    assert same_bytes(refhost.hdr_cache(hdr), g["cache_128x64"])


def test_hdr_cache_equals_reference_computed_golden():
    """runs everywhere: the cache the REFERENCE's calculateHdrCache computed for the synthetic environment"""
    g = np.load(GOLDEN)
    hdr = scenes.synth_hdr(128, 64)
    assert crc(hdr) == int(g["hdr_crc_128x64"])
    assert same_bytes(api.hdr_cache(hdr), g["cache_128x64"])


def test_synthetic_scenes_equal_reference_built_golden(bunny_scene, grid_scene):
    """runs everywhere: crc32 of the arrays the REFERENCE's readObj + buildBVHwithSAH built from the same OBJ text"""
    g = np.load(GOLDEN)
    for name, sc in (("bunny", bunny_scene), ("grid", grid_scene)):
        tris, nodes = sc[0], sc[1]
        assert (crc(tris), crc(nodes)) == (int(g[name + "_crc_tris"]), int(g[name + "_crc_nodes"])), name


@needs_reference
@pytest.mark.parametrize("form", ["v", "v/vt", "v/vt/vn"])
def test_read_obj_face_forms_equal_reference(tmp_path, form):
    """the three `f` forms the reference's parser distinguishes by counting slashes (P5/main.cpp:306-333), among vt / vn /
    comment / blank lines and trailing spaces -- same triangles and tree as the reference's own parser.  (`v//vn` makes the
    reference read uninitialised indices -- it crashes here; the product reads the leading integer of every token, see
    test_read_obj_tolerates_the_form_the_reference_cannot_parse.)"""
    base = scenes.blob_obj(2, 5).splitlines()
    rng = np.random.default_rng(len(form))
    out = ["# a comment", "", "vt 0.5 0.5", "vn 0 1 0"]
    for ln in base:
        if ln.startswith("f "):
            ids = ln.split()[1:]
            if form == "v":
                tok = ids
            elif form == "v/vt":
                tok = ["%s/1" % i for i in ids]
            elif form == "v/vt/vn":
                tok = ["%s/1/1" % i for i in ids]
            else:
                tok = ["%s//1" % i for i in ids]
            out.append("f " + " ".join(tok) + ("  " if rng.uniform() < 0.2 else ""))
        else:
            out.append(ln)
    path = _write(tmp_path, "forms.obj", "\n".join(out) + "\n")
    mat = api.Material(baseColor=(0.3, 0.5, 0.7))
    trans = api.transform_matrix((10, 20, 30), (0.1, -0.2, 0.3), (1.5, 0.5, 1.0))
    for smooth in (False, True):
        r_tris, r_nodes = refhost.build_scene([(path, mat.as_array(), trans, smooth)], 8, True)
        tl = api.TriangleList()
        tl.read_obj(path, mat, trans, smooth)
        tris, nodes = tl.build_bvh(8, api.BVH_SAH_FAST)
        assert same_bytes(tris, r_tris) and same_bytes(nodes, r_nodes), (form, smooth)


def test_read_obj_tolerates_the_form_the_reference_cannot_parse():
    plain = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 2 3\nf 1 3 4\n"
    vn = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nf 1//1 2//1 3//1\nf 1//1 3//1 4//1\n"
    out = []
    for text in (plain, vn):
        tl = api.TriangleList()
        tl.read_obj_text(text, api.Material(), api.transform_matrix(), False)
        out.append(tl.build_bvh(8, api.BVH_SAH_LITERAL))
    assert same_bytes(out[0][0], out[1][0]) and same_bytes(out[0][1], out[1][1])
