"""Post pass: pass3 tone map + gamma, PNG output."""
import struct
import zlib

import numpy as np
import pytest

from ezrt_b200 import api


def _read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert crc == (zlib.crc32(tag + body) & 0xffffffff)
        chunks.append((tag, body))
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[0][1][:10])
    assert (depth, ctype) == (8, 2) and chunks[-1][0] == b"IEND"
    raw = zlib.decompress(b"".join(b for t, b in chunks if t == b"IDAT"))
    img = np.frombuffer(raw, np.uint8).reshape(h, 1 + 3 * w)
    assert (img[:, 0] == 0).all()
    return img[:, 1:].reshape(h, w, 3)


def test_tonemap_definition(oracle):
    """toneMapping(c, 1.5) = c / (1 + lum/1.5), then pow(c, 1/2.2) (P5/shaders/pass3.fsh:14-25)."""
    rng = np.random.default_rng(1)
    fb = rng.uniform(0, 8, (5, 7, 3)).astype(np.float32)
    fb[0, 0] = 0.0
    out = oracle.tonemap(fb)
    lum = 0.3 * fb[..., 0] + 0.6 * fb[..., 1] + 0.1 * fb[..., 2]
    expect = (fb.astype(np.float64) / (1.0 + lum.astype(np.float64) / 1.5)[..., None]) ** (1 / 2.2)
    np.testing.assert_allclose(out, expect, rtol=2e-5, atol=1e-7)
    assert (out[0, 0] == 0).all()


def test_write_png_roundtrip(tmp_path, oracle):
    rng = np.random.default_rng(2)
    fb = rng.uniform(0, 3, (6, 9, 4)).astype(np.float32)
    fb[..., 3] = 1.0
    p = str(tmp_path / "a.png")
    api.write_png(p, fb, tonemap=True)
    img = _read_png(p)
    expect = np.clip(oracle.tonemap(fb) * np.float32(255.0), 0, 255).astype(np.uint8)[::-1]  # PNG is top row first
    np.testing.assert_array_equal(img, expect)
    api.write_png(p, fb[..., :3], tonemap=False)
    img = _read_png(p)
    np.testing.assert_array_equal(img, np.clip(fb[..., :3] * np.float32(255.0), 0, 255).astype(np.uint8)[::-1])


@pytest.mark.gpu
def test_gpu_tonemap_matches_oracle(oracle):
    import torch
    rng = np.random.default_rng(3)
    for ch in (3, 4):
        fb = rng.uniform(0, 20, (64, 48, ch)).astype(np.float32)
        fb[:4] = 0.0
        got = api.post_tonemap(torch.from_numpy(fb).cuda()).cpu().numpy()
        assert got.tobytes() == oracle.tonemap(fb).tobytes()


@pytest.mark.gpu
def test_gpu_tonemap_reproduces_reference_pass3_golden():
    """tests/golden/refshader.npz: output of the reference's own shaders/pass3.fsh (transpiled) for a fixed HDR frame"""
    import torch
    from tests import refshader_cases as cases
    from tests.test_ref_shader import _hdr_frame
    got = api.post_tonemap(torch.from_numpy(_hdr_frame()).cuda()).cpu().numpy()
    assert got.tobytes() == cases.load()["pass3_out"].tobytes()


@pytest.mark.gpu
def test_gpu_hdr_cache_reproduces_reference_computed_golden():
    """tests/golden/refhost.npz: the cache the reference's own calculateHdrCache computed (P5/main.cpp compiled in the
    authoring container) for the synthetic environment; the GPU kernels must produce the same bytes."""
    import os
    from ezrt_b200 import scenes
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refhost.npz"))
    got, _ = api.hdr_cache_device(scenes.synth_hdr(128, 64))
    assert np.ascontiguousarray(got, np.float32).tobytes() == g["cache_128x64"].tobytes()


@pytest.mark.gpu
def test_gpu_hdr_cache_equals_host():
    """calculateHdrCache on the GPU keeps every fp32 sum in the reference's order -> identical bits."""
    import time
    from ezrt_b200 import scenes
    for w, h in ((128, 64), (2048, 1024)):
        hdr = scenes.synth_hdr(w, h)
        t0 = time.perf_counter()
        host = api.hdr_cache(hdr)
        host_ms = 1e3 * (time.perf_counter() - t0)
        dev, dev_ms = api.hdr_cache_device(hdr)
        assert dev.tobytes() == host.tobytes()
        print("hdr cache %dx%d: host %.1f ms, device kernels %.2f ms" % (w, h, host_ms, dev_ms))
