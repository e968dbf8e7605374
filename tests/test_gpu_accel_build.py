"""The GPU builder of the acceleration tree's binary SAH tree (ezrt_b200/csrc/accel_build.cu) produces the host builder's
tree (host_scene.cpp ezrt_build_accel) node for node: links, leaf ranges, boxes and the triangle order -- on the bunny, on
S-1M, on tiny inputs, on coincident triangles (equal costs everywhere: the tie rule and the median rule below depth 32) and
on grids of equal centroids (stable sort order)."""
import numpy as np
import pytest

from ezrt_b200 import api, scenes

pytestmark = pytest.mark.gpu


def _same_tree(tris, leaf_n=4):
    ld, bd, od, ms_d = api.accel_build(tris, leaf_n, "device")
    lh, bh, oh, ms_h = api.accel_build(tris, leaf_n, "host")
    assert ld.shape == lh.shape, (ld.shape, lh.shape)
    assert np.array_equal(od, oh), "triangle order differs at %d positions" % int((od != oh).sum())
    assert np.array_equal(ld, lh), "links differ first at node %d" % int(np.argwhere((ld != lh).any(1))[0, 0])
    assert np.array_equal(bd, bh), "boxes differ first at node %d" % int(np.argwhere((bd != bh).any(1))[0, 0])   # == : -0.0 equals +0.0
    return len(ld), ms_d, ms_h


def _random_tris(rng, n, spread=1.0):
    t = np.zeros((n, 36), np.float32)
    c = rng.uniform(-spread, spread, (n, 1, 3))
    t[:, :9] = (c + rng.uniform(-0.05, 0.05, (n, 3, 3))).reshape(n, 9).astype(np.float32)
    return t


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 33, 257, 1000])
def test_small_inputs(n):
    rng = np.random.default_rng(100 + n)
    _same_tree(_random_tris(rng, n))
    _same_tree(_random_tris(rng, n), leaf_n=1)


def test_bunny_and_leaf_sizes():
    tris = scenes.s_p3_bunny()[0]
    for leaf_n in (1, 2, 4, 8):
        _same_tree(tris, leaf_n)


def test_equal_costs_and_equal_keys():
    rng = np.random.default_rng(5)
    one = _random_tris(rng, 1)
    _same_tree(np.repeat(one, 300, axis=0))                       # coincident: every cost ties, 300 > 2^(depth limit) never reached
    _same_tree(np.repeat(_random_tris(rng, 40), 25, axis=0))      # 40 clusters of 25 coincident triangles
    g = np.zeros((16 * 16, 36), np.float32)                       # a regular grid in the plane z = 0: equal centroid coordinates
    for j in range(16):
        for i in range(16):
            g[j * 16 + i, :9] = [i, j, 0, i + 0.5, j, 0, i, j + 0.5, 0]
    _same_tree(g)
    _same_tree(g[rng.permutation(len(g))])
    flat = _random_tris(rng, 500)
    flat[:, [1, 4, 7]] = 0.0                                      # all in the plane y = 0 (zero areas, +-0 extents)
    _same_tree(flat)
    neg = flat.copy()
    neg[:, [1, 4, 7]] = -0.0
    _same_tree(np.concatenate([flat, neg]))


def test_deep_tree_uses_the_median_rule():
    """Triangles whose SAH split peels one off per level (sizes growing geometrically along x): deeper than 32 levels,
    where both builders switch to the median split."""
    n = 4000
    t = np.zeros((n, 36), np.float32)
    x = np.float32(1.0)
    for i in range(n):
        w = np.float32(1e-3) * np.float32(1.004) ** i
        t[i, :9] = [x, 0, 0, x + w, 0, 0, x, w, 0]
        x = np.float32(x + w * np.float32(1.5))
    nn, _, _ = _same_tree(t)
    assert nn > n // 4


def test_s1m_same_tree_and_faster():
    tris = scenes.s_1m_bunny()[0]
    nn, ms_d, ms_h = _same_tree(tris)
    _, _, _, ms_d2 = api.accel_build(tris, 4, "device")   # second call: no first-use costs
    print("S-1M binary SAH tree: %d nodes; device %.1f ms (first call %.1f), host %.1f ms" % (nn, ms_d2, ms_d, ms_h))
    assert ms_d2 < ms_h
