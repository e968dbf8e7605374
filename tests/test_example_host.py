"""examples/ezrt_main.cpp: the reference's main() + display() as a C++ host program over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from ezrt_b200 import api, build, scenes

SCENE = """# P3-style scene (P3/main.cpp:688-701) with synthetic meshes
set baseColor 1 1 1
mesh blob.obj smooth rotate 0 0 0 translate 0.3 -0.65 0 scale 1.5 1.5 1.5
set baseColor 0.725 0.71 0.68
mesh box.obj flat rotate 0 0 0 translate 0 -1.4 0 scale 18.83 0.01 18.83
set emissive 30 20 10
set baseColor 1 1 1
mesh sphere.obj flat rotate 0 0 0 translate 0 0.9 0 scale 1 1 1
camera 20 15 4
"""


def _scene_dir(tmp_path):
    (tmp_path / "blob.obj").write_text(scenes.blob_obj())
    (tmp_path / "box.obj").write_text(scenes.box_obj())
    (tmp_path / "sphere.obj").write_text(scenes.sphere_obj())
    (tmp_path / "scene.txt").write_text(SCENE)
    return str(tmp_path / "scene.txt")


def test_example_host_builds_and_fails_loudly_without_a_gpu(tmp_path):
    exe = build.build_example()
    assert os.access(exe, os.X_OK)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    import torch
    if not torch.cuda.is_available():  # no CPU fallback: the C ABI reports the missing device, the host exits non-zero
        r = subprocess.run([exe, "--scene", _scene_dir(tmp_path), str(tmp_path / "out.png"), "--size", "32", "24", "--spp", "1", "--mode", "2"],
                           capture_output=True, text=True)
        assert r.returncode == 1 and "CUDA" in r.stderr and not (tmp_path / "out.png").exists()


@pytest.mark.gpu
def test_example_host_renders_the_same_png_as_the_python_mirror(tmp_path):
    exe = build.build_example()
    scene_txt = _scene_dir(tmp_path)
    out = tmp_path / "cxx.png"
    r = subprocess.run([exe, "--scene", scene_txt, str(out), "--size", "96", "64", "--spp", "3", "--mode", "2", "--bounces", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Mrays/s" in r.stdout
    tl, cam, _ = api.load_scene_file(scene_txt)
    tris, nodes = tl.build_bvh(8, api.BVH_SAH_FAST)
    eye, rot = api.camera_orbit(*cam)
    sc = api.Scene(tris, nodes)
    try:
        img = sc.render(api.RenderConfig(width=96, height=64, spp=3, max_bounce=2, mode=api.MODE_DISNEY_SOBOL_P5, eye=tuple(eye),
                                         camera_rotate=tuple(rot)))
    finally:
        sc.close()
    ref = tmp_path / "py.png"
    api.write_png(str(ref), img, tonemap=True)
    assert out.read_bytes() == ref.read_bytes()
