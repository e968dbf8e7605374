"""BASELINE.json's GPU configurations on their OWN pixel grid and scene (C2: Stanford-bunny scene, 1024x1024, diffuse;
C3: S-1M = 201 bunnies, 999,860 triangles, 1920x1080, Disney + Sobol; C4: C3 + 2k environment importance sampling + MIS):
the GPU renders the FULL frame, the CPU oracle renders crop windows of the same pixel grid (oracle_render_window:
pixel (px, py) depends only on (px, py, frame), P5/fsh:315-318, :379-382), bits must be identical -- in the windows and,
through ray counts, over the whole frame.  Plus the API features added in round 2."""
import numpy as np
import pytest

from ezrt_b200 import api, scenes
from tests.test_gpu_parity import assert_same_bits

pytestmark = pytest.mark.gpu

ENV = (0.35, 0.45, 0.6)


def _cfg(eye, cam, **kw):
    return api.RenderConfig(eye=tuple(eye), camera_rotate=tuple(cam), env_color=ENV, **kw)


def _windows(W, H, w=64, h=40):
    """Crop windows spread over the frame: corners (clipped-tile edge included), centre, and two off-centre ones."""
    pts = [(0, 0), (W - w, H - h), (W // 2 - w // 2, H // 2 - h // 2), (W // 4, (2 * H) // 3), ((3 * W) // 4 - w, H // 5)]
    return [(x, y, x + w, y + h) for x, y in pts]


@pytest.fixture(scope="module")
def s1m():
    tris, nodes, eye, cam = scenes.s_1m_bunny()
    hdr = scenes.synth_hdr(2048, 1024)
    cache = api.hdr_cache_device(hdr)[0]
    sc = api.Scene(tris, nodes, hdr, cache)
    yield dict(tris=tris, nodes=nodes, eye=eye, cam=cam, hdr=hdr, cache=cache, scene=sc)
    sc.close()


def test_s1m_is_the_scene_of_survey_8d(s1m):
    assert s1m["tris"].shape[0] == 201 * 4968 + 4 * 320 + 12 == 999860


@pytest.mark.parametrize("mode,spp", [(api.MODE_DISNEY_SOBOL_P5, 2), (api.MODE_DISNEY_IS_MIS_P5, 2)])
def test_c3_c4_full_frame_on_their_own_grid(oracle, s1m, mode, spp):
    W, H = 1920, 1080
    cfg = _cfg(s1m["eye"], s1m["cam"], width=W, height=H, spp=spp, max_bounce=2, mode=mode)
    img = s1m["scene"].render(cfg)
    c = s1m["scene"].counters()
    assert np.isfinite(img).all() and c.samples == W * H * spp
    assert c.deferred_rays < 0.01 * c.rays
    for win in _windows(W, H):
        ref, _ = oracle.render(s1m["tris"], s1m["nodes"], cfg, hdr=s1m["hdr"], hdr_cache=s1m["cache"], window=win)
        x0, y0, x1, y1 = win
        assert_same_bits(img[y0:y1, x0:x1], ref, "mode %d window %s of the 1920x1080 grid" % (mode, win))
    # the three traversal policies give the same frame and the same number of rays
    for policy in (api.TRAVERSE_PRUNED,):
        cfg_p = _cfg(s1m["eye"], s1m["cam"], width=W, height=H, spp=spp, max_bounce=2, mode=mode, traverse=policy)
        other = s1m["scene"].render(cfg_p)
        assert other.tobytes() == img.tobytes()
        assert s1m["scene"].counters().rays == c.rays


def test_c2_full_frame_on_its_own_grid(oracle):
    tris, nodes, eye, cam = scenes.s_p3_bunny()
    assert tris.shape[0] == 5300
    sc = api.Scene(tris, nodes)
    try:
        W = H = 1024
        cfg = _cfg(eye, cam, width=W, height=H, spp=2, max_bounce=2, mode=api.MODE_DIFFUSE_P3)
        img = sc.render(cfg)
        ref, rc = oracle.render(tris, nodes, cfg)        # the whole 1024x1024 frame: ~1 s of CPU
        assert_same_bits(img, ref, "C2 full frame")
        assert sc.counters().rays == rc["rays"]
    finally:
        sc.close()


def test_counting_instantiation_and_accumulating_counters(s1m):
    sc = s1m["scene"]
    base = _cfg(s1m["eye"], s1m["cam"], width=320, height=180, spp=2, max_bounce=2, mode=api.MODE_DISNEY_IS_MIS_P5)
    a = sc.render(base).copy()
    ca = sc.counters()
    assert ca.node_visits == 0 and ca.tri_tests == 0
    cnt = _cfg(s1m["eye"], s1m["cam"], width=320, height=180, spp=2, max_bounce=2, mode=api.MODE_DISNEY_IS_MIS_P5, profile=2)
    b = sc.render(cnt)
    cb = sc.counters()
    assert b.tobytes() == a.tobytes() and cb.rays == ca.rays
    assert cb.node_visits > 3 * cb.rays and cb.tri_tests > cb.rays          # several node and triangle records per ray
    assert cb.tri_bytes == 64 * cb.tri_tests
    assert cb.node_bytes == 96 * cb.node_visits_96 + 128 * (cb.node_visits - cb.node_visits_96)
    # EZRT_PARAM_ACCUMULATE: the second render's counters continue from the first
    acc = _cfg(s1m["eye"], s1m["cam"], width=320, height=180, spp=2, first_frame=2, max_bounce=2, mode=api.MODE_DISNEY_IS_MIS_P5, accumulate=True)
    fb = a.reshape(-1, 3).copy()
    sc.render(base)
    sc.render(acc, framebuffer=fb)
    cc = sc.counters()
    assert cc.samples == 2 * ca.samples and cc.rays > ca.rays


def test_deep_caller_tree_and_degenerate_geometry(oracle):
    """(i) a caller tree deeper than 64 levels (a chain; the shader's stack holds 256) is accepted and rendered exactly;
    (ii) 300 coincident triangles: the acceleration tree must not fail the scene, ties go to the exact kernel."""
    rng = np.random.default_rng(3)
    n = 90
    tl = api.TriangleList()
    tris = np.zeros((n, 36), np.float32)
    for i in range(n):  # a row of small triangles along x: a median/SAH builder gives a shallow tree; we hand-build a chain
        x = 0.1 * i
        tris[i, :9] = [x, 0, 0, x + 0.08, 0, 0, x, 0.08, 0]
        tris[i, 9:18] = [0, 0, 1] * 3
        tris[i, 21:24] = 0.8
        tris[i, 28] = 0.5
    nodes = [np.zeros(12, np.float32)]   # dummy node 0
    # chain: node k (inner) -> left = leaf with triangle k, right = node for the rest
    def box(a, b):
        v = tris[a:b, :9].reshape(-1, 3)
        return v.min(0), v.max(0)
    idx = 1
    recs = []
    for k in range(n - 1):
        inner, leaf = idx, idx + 1
        idx += 2
        aa, bb = box(k, n)
        la, lb = box(k, k + 1)
        recs.append((inner, [leaf, idx if k < n - 2 else idx, 0, 0, 0, 0, *aa, *bb]))
        recs.append((leaf, [0, 0, 0, 1, k, 0, *la, *lb]))
    la, lb = box(n - 1, n)
    recs.append((idx, [0, 0, 0, 1, n - 1, 0, *la, *lb]))
    arr = np.zeros((idx + 1, 12), np.float32)
    for i, r in recs:
        arr[i] = r
    eye, cam = api.camera_orbit(10.0, 20.0, 9.0)
    sc = api.Scene(tris, arr)
    try:
        cfg = _cfg(eye, cam, width=64, height=48, spp=2, max_bounce=2, mode=api.MODE_DISNEY_SOBOL_P5)
        ref, rc = oracle.render(tris, arr, cfg)
        assert_same_bits(sc.render(cfg), ref, "90-deep chain tree")
        assert sc.counters().rays == rc["rays"]
    finally:
        sc.close()
    # (ii)
    t2 = np.repeat(tris[:1], 300, axis=0).copy()
    t2[:, :9] *= 20.0
    tl.append_encoded(t2)
    tt, nn = tl.build_bvh(8, api.BVH_SAH_FAST) if False else (None, None)
    tl2 = api.TriangleList()
    tl2.append_encoded(t2)
    tt, nn = tl2.build_bvh(8, api.BVH_MEDIAN)
    sc = api.Scene(tt, nn)
    try:
        cfg = _cfg(eye, cam, width=48, height=32, spp=1, max_bounce=1, mode=api.MODE_DISNEY_SOBOL_P5)
        ref, rc = oracle.render(tt, nn, cfg)
        assert_same_bits(sc.render(cfg), ref, "coincident triangles")
    finally:
        sc.close()


@pytest.mark.parametrize("lane", ["1", "0"])
def test_deferred_lane_few_and_many_deferred_rays(oracle, lane, monkeypatch):
    """The accel policy's deferred rays are traced exactly and shaded on a side stream beside k_shade (capi.cu "deferred lane") when
    there are at most EZRT_SIDE_CAP = 65536 of them in a pass, in line otherwise; EZRT_DEFERRED_LANE=0 is round 1's in-line pass.
    A wall of 40 coincident triangles in front of the bunny defers every ray that hits it (ties): at 64x48 a few thousand rays per
    pass, at 640x480x2 more than the cap in the camera pass and fewer in the others.  All bit-identical to the oracle,
    with and without shadow rays (IS/MIS shades through the same lane), and equal ray counts."""
    monkeypatch.setenv("EZRT_DEFERRED_LANE", lane)
    tris, nodes, eye, cam = scenes.s_p3_bunny()
    wall = np.zeros((40, 36), np.float32)
    e = np.asarray(eye, np.float64)
    n = e / np.linalg.norm(e)
    u = np.cross([0.0, 1.0, 0.0], n)
    u /= np.linalg.norm(u)
    v = np.cross(n, u)
    c, h = 0.55 * e, 0.6 * np.linalg.norm(e)      # between the camera and the bunny, facing the camera, wider than the view
    wall[:, :9] = np.concatenate([c - 2 * h * u - h * v, c + 2 * h * u - h * v, c + 3 * h * v]).astype(np.float32)
    wall[:, 9:18] = np.tile(n.astype(np.float32), 3)
    wall[:, 18:21] = 0.0          # emissive 0
    wall[:, 21:24] = (0.7, 0.6, 0.5)
    wall[:, 28] = 0.4
    tl = api.TriangleList()
    tl.append_encoded(np.concatenate([tris, wall]))
    tt, nn = tl.build_bvh(8, api.BVH_SAH_FAST)
    hdr = scenes.synth_hdr(64, 32)
    cache = api.hdr_cache(hdr)
    sc = api.Scene(tt, nn, hdr, cache)
    try:
        for mode in (api.MODE_DISNEY_SOBOL_P5, api.MODE_DISNEY_IS_MIS_P5):
            cfg = _cfg(eye, cam, width=64, height=48, spp=3, max_bounce=2, mode=mode)
            ref, rc = oracle.render(tt, nn, cfg, hdr=hdr, hdr_cache=cache)
            assert_same_bits(sc.render(cfg), ref, "few deferred rays, mode %d" % mode)
            c = sc.counters()
            assert c.rays == rc["rays"] and 0 < c.deferred_rays < 65536
            cfg = _cfg(eye, cam, width=640, height=480, spp=2, max_bounce=2, mode=mode)
            img = sc.render(cfg)
            c = sc.counters()
            assert c.deferred_rays > 5 * 65536, c.deferred_rays     # 3 extend + 2 shadow passes: at least one pass is over the cap
            for win in _windows(640, 480, 48, 32):
                ref, _ = oracle.render(tt, nn, cfg, hdr=hdr, hdr_cache=cache, window=win)
                x0, y0, x1, y1 = win
                assert_same_bits(img[y0:y1, x0:x1], ref, "many deferred rays, mode %d, window %s" % (mode, win))
    finally:
        sc.close()


@pytest.mark.parametrize("env", [{}, {"EZRT_ACCEL_Q16": "0"}, {"EZRT_ACCEL": "8"}, {"EZRT_CAMERA_ORDER": "frame"}, {"EZRT_W4_COLLAPSE": "greedy"},
                                 {"EZRT_SORT_RAYS": "1"}])
def test_every_form_of_the_accel_policy_gives_the_oracle_image(oracle, grid_scene, small_hdr, env, monkeypatch):
    """The acceleration tree's form (4-wide exact / 4-wide + 16-bit planes for the incoherent launches / W8), the collapse rule, the
    order in which the camera pass takes its work and the optional ray sort are read from the environment at scene creation; none of
    them may change a bit of the image or the ray counts."""
    tris, nodes, eye, cam = grid_scene
    hdr, cache = small_hdr
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sc = api.Scene(tris, nodes, hdr, cache)
    try:
        for mode in (api.MODE_DISNEY_SOBOL_P5, api.MODE_DISNEY_IS_MIS_P5):
            cfg = _cfg(eye, cam, width=200, height=120, spp=3, max_bounce=2, mode=mode)
            ref, rc = oracle.render(tris, nodes, cfg, hdr=hdr, hdr_cache=cache)
            assert_same_bits(sc.render(cfg), ref, "env %s mode %d" % (env, mode))
            assert sc.counters().rays == rc["rays"]
    finally:
        sc.close()
