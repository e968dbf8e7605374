"""The C-ABI library loads and exports every symbol include/ezrt.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ezrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ezrt_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from ezrt_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libezrt_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "python binding lacks %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_struct_layout_matches_header():
    from ezrt_b200 import _lib
    # ezrt_render_params: 6 int32 + 3+16+3 floats + 6 int32 + 4 reserved
    assert ctypes.sizeof(_lib.RenderParams) == 4 * (6 + 22 + 6 + 4)
    # ezrt_counters: 6 uint64 + double + 6 uint64
    assert ctypes.sizeof(_lib.Counters) == 8 * 13
    assert _lib.lib.ezrt_version() == 200


def test_product_does_not_link_the_oracle():
    """The product path must never route through oracle/: no symbol in the library, no import in the package
    (ezrt_b200/build.py only knows how to BUILD the checker, which is not using it)."""
    from ezrt_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    assert not hasattr(raw, "oracle_render") and not hasattr(raw, "oracle_trace_rays")
    pkg = os.path.join(ROOT, "ezrt_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or fn == "build.py":
                continue
            src = open(os.path.join(dirpath, fn)).read()
            assert "oracle_binding" not in src and "libezrt_oracle" not in src and "ezrt_oracle" not in src, fn


def test_device_entry_points_fail_loudly_without_gpu():
    import torch
    from ezrt_b200 import api
    if torch.cuda.is_available():
        return
    tris = np.zeros((1, 36), np.float32)
    nodes = np.zeros((2, 12), np.float32)
    nodes[1, 3] = 1
    try:
        api.Scene(tris, nodes)
    except api.EzrtError as e:
        assert e.code == -3
    else:
        raise AssertionError("scene_create must fail without a CUDA device (no CPU fallback)")


def test_partition_covers_image_exactly_once():
    from ezrt_b200 import api
    W, H = 203, 77
    for count in (1, 2, 3, 8):
        full = np.zeros((H, W, 1), np.float32)
        total = 0
        for rank in range(count):
            n = api.partition_pixels(W, H, rank, count)
            total += n
            api.partition_scatter_host(np.full((n, 1), rank + 1, np.float32), full, W, H, 1, rank, count)
            # scatter adds nothing twice: every written pixel is this rank's
        assert total == W * H
        assert (full > 0).all()
        ty, tx = np.mgrid[0:H, 0:W] // 16
        np.testing.assert_array_equal(full[..., 0], ((tx + ty) % count + 1).astype(np.float32))
