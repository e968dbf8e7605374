"""The 8-wide quantised acceleration tree (ezrt_b200/csrc/accel_w8.cpp, w8_node.h) is conservative: walking it with the
device's decode arithmetic and visit rule (CPU model tools/w8_model.cpp, built over the PRODUCT builders) finds, for every
ray, exactly the closest-hit distance brute force over all triangles finds.  No GPU needed."""
import os
import subprocess

import numpy as np

from ezrt_b200 import build, scenes


def _run_model(tmp_path, tris, rays, brute):
    exe = build.build_w8_model()
    tf, rf = os.path.join(tmp_path, "tris.f32"), os.path.join(tmp_path, "rays.f32")
    np.ascontiguousarray(tris, np.float32).tofile(tf)
    np.ascontiguousarray(rays, np.float32).tofile(rf)
    r = subprocess.run([exe, tf, str(tris.shape[0]), rf] + (["brute"] if brute else []), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert "differing from" in r.stdout and ": 0 of" in r.stdout, r.stdout
    assert "violations 0" in r.stdout, r.stdout        # the 4-wide collapse covers every triangle exactly once, leaves <= 4
    return r.stdout


def _rays(tris, n, seed):
    """Half start on surfaces (bounce-like, random directions incl. grazing and axis-parallel ones), half outside looking in."""
    rng = np.random.default_rng(seed)
    v = tris[:, :9].reshape(-1, 3, 3)
    pick = rng.integers(0, v.shape[0], n)
    w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    o = (v[pick] * w[:, :, None]).sum(1)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    lo, hi = v.reshape(-1, 3).min(0), v.reshape(-1, 3).max(0)
    far = rng.uniform(lo - 3, hi + 3, (n // 2, 3)).astype(np.float32)
    o[: n // 2] = far
    d[: n // 2] = (v[pick[: n // 2]].mean(1) - far)
    d[: n // 2] /= np.linalg.norm(d[: n // 2], axis=1, keepdims=True)
    d[::101, 0] = 0.0            # exactly axis-parallel components: left to the exact kernel, must not miscount
    d[::103, 1] = 1e-30
    rays = np.zeros((n, 7), np.float32)
    rays[:, :3], rays[:, 3:6], rays[:, 6] = o, d, 1
    return rays


def test_w8_traversal_equals_brute_force_on_the_bunny_scene(tmp_path):
    tris, _, _, _ = scenes.s_p3_bunny()
    out = _run_model(str(tmp_path), tris, _rays(tris, 20000, 1), brute=True)
    assert "8-wide nodes" in out


def test_w8_traversal_on_a_degenerate_soup(tmp_path):
    """Coincident, needle and zero-area triangles: the builder must neither fail nor lose a hit."""
    rng = np.random.default_rng(5)
    n = 3000
    tris = np.zeros((n, 36), np.float32)
    p = rng.uniform(-1, 1, (n, 3, 3)).astype(np.float32) * rng.choice([1e-3, 0.1, 1.0], (n, 1, 1)).astype(np.float32)
    p += rng.uniform(-2, 2, (n, 1, 3)).astype(np.float32)
    p[:200] = p[0]                      # 200 coincident triangles
    p[200:260, 2] = p[200:260, 1]       # zero-area
    tris[:, :9] = p.reshape(n, 9)
    tris[:, 21:24] = 1.0
    _run_model(str(tmp_path), tris, _rays(tris, 6000, 2), brute=True)


def test_w8_traversal_large_grid_against_exact_boxes(tmp_path):
    tris, _, _, _ = scenes.s_grid(4, 3, 2, mesh="bunny")
    out = _run_model(str(tmp_path), tris, _rays(tris, 40000, 3), brute=False)
    print(out)


def _run_w4_check(tmp_path, tris):
    exe = build.build_w4_check()
    tf = os.path.join(tmp_path, "tris_w4.f32")
    np.ascontiguousarray(tris, np.float32).tofile(tf)
    r = subprocess.run([exe, tf, str(tris.shape[0])], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    return r.stdout


def test_w4_builder_is_thread_count_independent_and_well_formed(tmp_path):
    """ezrt_build_w4 (the default 4-wide form + its 96-byte Q16 twin): the same arrays for 1, 2, 5 and 16 threads and both collapse
    rules; every triangle in exactly one leaf, child boxes (exact and quantised) contain their triangles."""
    tris, _, _, _ = scenes.s_p3_bunny()
    out = _run_w4_check(str(tmp_path), tris)
    assert "violations 0" in out
    rng = np.random.default_rng(11)
    soup = np.zeros((70000, 36), np.float32)     # > 65536 binary nodes: the threaded path without the override, too
    c = rng.uniform(-5, 5, (70000, 1, 3))
    soup[:, :9] = (c + rng.uniform(-0.05, 0.05, (70000, 3, 3))).reshape(-1, 9).astype(np.float32)
    soup[:300, :9] = soup[0, :9]                 # coincident triangles
    _run_w4_check(str(tmp_path), soup)
