"""Host-side mirror of the reference's main()/display() for the path-tracing hot path.

Names follow the reference (P5/main.cpp): Material, readObj -> TriangleList.read_obj,
getTransformMatrix -> transform_matrix, buildBVHwithSAH -> TriangleList.build_bvh,
calculateHdrCache -> hdr_cache, display() -> Scene.render.  Everything calls the C ABI of
include/ezrt.h through ctypes; numpy carries host arrays, torch (optional) carries device
framebuffers and streams.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Counters, EzrtError, RenderParams, check, lib

MODE_DIFFUSE_P3 = 0
MODE_DISNEY_ANISO_P4 = 1
MODE_DISNEY_SOBOL_P5 = 2
MODE_DISNEY_IS_MIS_P5 = 3
MODES = {"diffuse_p3": 0, "disney_aniso_p4": 1, "disney_sobol_p5": 2, "disney_is_mis_p5": 3}

TRAVERSE_ACCEL = 0
TRAVERSE_REFERENCE = 1
TRAVERSE_PRUNED = 2
PIPELINE_WAVEFRONT = 0
PIPELINE_MEGAKERNEL = 1

BVH_SAH_FAST = 0
BVH_SAH_LITERAL = 1
BVH_MEDIAN = 2

TRIANGLE_FLOATS = 36
BVHNODE_FLOATS = 12


def _fp(a):
    return a.ctypes.data_as(_lib.c_float_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


@dataclass
class Material:
    """struct Material, P5/main.cpp:27-42 (same defaults)."""
    emissive: tuple = (0.0, 0.0, 0.0)
    baseColor: tuple = (1.0, 1.0, 1.0)
    subsurface: float = 0.0
    metallic: float = 0.0
    specular: float = 0.5
    specularTint: float = 0.0
    roughness: float = 0.5
    anisotropic: float = 0.0
    sheen: float = 0.0
    sheenTint: float = 0.5
    clearcoat: float = 0.0
    clearcoatGloss: float = 1.0
    IOR: float = 1.0
    transmission: float = 0.0

    def as_array(self):
        return np.array(list(self.emissive) + list(self.baseColor) + [
            self.subsurface, self.metallic, self.specular, self.specularTint, self.roughness, self.anisotropic,
            self.sheen, self.sheenTint, self.clearcoat, self.clearcoatGloss, self.IOR, self.transmission], dtype=np.float32)


def transform_matrix(rotate=(0, 0, 0), translate=(0, 0, 0), scale=(1, 1, 1)):
    """getTransformMatrix(rotateCtrl, translateCtrl, scaleCtrl), P5/main.cpp:255-271 -> 16 floats, column-major."""
    out = np.zeros(16, dtype=np.float32)
    lib.ezrt_transform_matrix(_fp(_f32(rotate)), _fp(_f32(translate)), _fp(_f32(scale)), _fp(out))
    return out


def camera_orbit(rotate_angle=0.0, up_angle=0.0, r=4.0):
    """eye / cameraRotate of display(), P5/main.cpp:710-713."""
    eye = np.zeros(3, dtype=np.float32)
    cam = np.zeros(16, dtype=np.float32)
    lib.ezrt_camera_orbit(float(rotate_angle), float(up_angle), float(r), _fp(eye), _fp(cam))
    return eye, cam


class TriangleList:
    """std::vector<Triangle> triangles of main() (P5/main.cpp:801) plus its BVH."""

    def __init__(self):
        self._h = lib.ezrt_trilist_create()
        if not self._h:
            raise MemoryError("ezrt_trilist_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.ezrt_trilist_destroy(h)

    def __len__(self):
        return check(lib.ezrt_trilist_size(self._h))

    def read_obj(self, path, material, trans, smooth_normal):
        """readObj(filepath, triangles, material, trans, smoothNormal), P5/main.cpp:274-392."""
        check(lib.ezrt_trilist_read_obj(self._h, str(path).encode(), _fp(material.as_array()), _fp(_f32(trans)), int(smooth_normal)))
        return self

    def read_obj_text(self, text, material, trans, smooth_normal):
        data = text.encode() if isinstance(text, str) else bytes(text)
        check(lib.ezrt_trilist_read_obj_text(self._h, data, len(data), _fp(material.as_array()), _fp(_f32(trans)), int(smooth_normal)))
        return self

    def append_encoded(self, tris):
        tris = _f32(tris, (-1, TRIANGLE_FLOATS))
        check(lib.ezrt_trilist_append_encoded(self._h, _fp(tris), tris.shape[0]))
        return self

    def build_bvh(self, leaf_n=8, builder=BVH_SAH_FAST):
        """nodes{testNode}; buildBVHwithSAH(triangles, nodes, 0, N-1, 8) + encode (P5/main.cpp:830-871).
        Returns (triangles_encoded [N,36], nodes_encoded [M,12])."""
        n_nodes = check(lib.ezrt_trilist_build_bvh(self._h, int(leaf_n), int(builder)))
        return self.encode_triangles(), self.encode_nodes(n_nodes)

    def encode_triangles(self):
        out = np.zeros((len(self), TRIANGLE_FLOATS), dtype=np.float32)
        check(lib.ezrt_trilist_encode_triangles(self._h, _fp(out)))
        return out

    def encode_nodes(self, n_nodes=None):
        if n_nodes is None:
            n_nodes = check(lib.ezrt_trilist_node_count(self._h))
        out = np.zeros((n_nodes, BVHNODE_FLOATS), dtype=np.float32)
        check(lib.ezrt_trilist_encode_nodes(self._h, _fp(out)))
        return out


OBJ_HARDENED = 2


def host_sort_is_reference():
    """True if this host's std::sort reproduces libstdc++'s order of equal keys, i.e. build_bvh yields the reference's
    triangle order (ezrt_host_sort_is_reference, include/ezrt.h)."""
    return bool(lib.ezrt_host_sort_is_reference())


def load_scene_file(path):
    """Scene description file -> (TriangleList, (rotatAngle, upAngle, r), hdr path or None); see include/ezrt.h."""
    tl = TriangleList()
    cam = np.zeros(3, dtype=np.float32)
    buf = C.create_string_buffer(4096)
    check(lib.ezrt_scene_file_load(str(path).encode(), tl._h, _fp(cam), buf, 4096))
    hdr = buf.value.decode() or None
    return tl, tuple(float(x) for x in cam), hdr


def hdr_load(path):
    """HDRLoader::load, P5/lib/hdrloader.cpp:29-97 -> float32 [h, w, 3], row 0 = first scanline."""
    w, h = C.c_int(0), C.c_int(0)
    check(lib.ezrt_hdr_load(str(path).encode(), C.byref(w), C.byref(h), None))
    cols = np.zeros((h.value, w.value, 3), dtype=np.float32)
    check(lib.ezrt_hdr_load(str(path).encode(), C.byref(w), C.byref(h), _fp(cols)))
    return cols


def hdr_cache(hdr):
    """calculateHdrCache(HDR, width, height), P5/main.cpp:592-689."""
    hdr = _f32(hdr)
    h, w = hdr.shape[0], hdr.shape[1]
    out = np.zeros((h, w, 3), dtype=np.float32)
    check(lib.ezrt_hdr_cache(_fp(hdr), w, h, _fp(out)))
    return out


def hdr_cache_device(hdr, device=0):
    """calculateHdrCache on the GPU (bit-identical to hdr_cache).  Returns (cache, kernel milliseconds)."""
    hdr = _f32(hdr)
    h, w = hdr.shape[0], hdr.shape[1]
    out = np.zeros((h, w, 3), dtype=np.float32)
    ms = C.c_double(0.0)
    check(lib.ezrt_hdr_cache_device(int(device), _fp(hdr), w, h, _fp(out), C.byref(ms)))
    return out, ms.value


def partition_pixels(width, height, rank, count):
    return int(check(lib.ezrt_partition_pixels(width, height, rank, count)))


def partition_scatter_host(compact, full, width, height, channels, rank, count):
    compact = _f32(compact)
    assert full.dtype == np.float32 and full.flags.c_contiguous
    check(lib.ezrt_partition_scatter_host(_fp(compact), _fp(full), width, height, channels, rank, count))
    return full


@dataclass
class RenderConfig:
    """The uniforms display() sets plus the shader literals (ezrt_render_params)."""
    width: int = 512
    height: int = 512
    spp: int = 1
    first_frame: int = 0
    max_bounce: int = 2
    mode: int = MODE_DISNEY_SOBOL_P5
    eye: tuple = (0.0, 0.0, 4.0)
    camera_rotate: tuple = tuple(np.eye(4, dtype=np.float32).reshape(-1))
    env_color: tuple = (0.0, 0.0, 0.0)
    traverse: int = TRAVERSE_ACCEL
    pipeline: int = PIPELINE_WAVEFRONT
    out_channels: int = 3
    part_rank: int = 0
    part_count: int = 1
    frames_per_batch: int = 0
    profile: int = 0
    accumulate: bool = False   # EZRT_PARAM_ACCUMULATE: counters / kernel times continue from the previous render

    def to_struct(self):
        p = RenderParams()
        p.width, p.height, p.spp, p.first_frame = int(self.width), int(self.height), int(self.spp), int(self.first_frame)
        p.max_bounce, p.mode = int(self.max_bounce), int(self.mode)
        p.eye[:] = [float(x) for x in self.eye]
        p.camera_rotate[:] = [float(x) for x in np.asarray(self.camera_rotate, dtype=np.float32).reshape(-1)]
        p.env_color[:] = [float(x) for x in self.env_color]
        p.traverse, p.pipeline, p.out_channels = int(self.traverse), int(self.pipeline), int(self.out_channels)
        p.part_rank, p.part_count, p.frames_per_batch = int(self.part_rank), int(self.part_count), int(self.frames_per_batch)
        p.profile = int(self.profile)
        p.reserved[0] = 1 if self.accumulate else 0
        return p


class Scene:
    """Device-resident scene: the two texture buffers + two HDR textures of P5/main.cpp:878-906."""

    def __init__(self, tris, nodes, hdr=None, hdr_cache_=None, device=0, hdr_filter_linear=True):
        self.tris = _f32(tris, (-1, TRIANGLE_FLOATS))
        self.nodes = _f32(nodes, (-1, BVHNODE_FLOATS))
        self.hdr = None if hdr is None else _f32(hdr)
        self.hdr_cache = None if hdr_cache_ is None else _f32(hdr_cache_)
        self.device = int(device)
        hw = hh = 0
        if self.hdr is not None:
            hh, hw = self.hdr.shape[0], self.hdr.shape[1]
        elif self.hdr_cache is not None:
            hh, hw = self.hdr_cache.shape[0], self.hdr_cache.shape[1]
        self._h = C.c_void_p()
        check(lib.ezrt_scene_create(self.device, _fp(self.tris), self.tris.shape[0], _fp(self.nodes), self.nodes.shape[0],
                                    None if self.hdr is None else _fp(self.hdr),
                                    None if self.hdr_cache is None else _fp(self.hdr_cache), hw, hh,
                                    int(bool(hdr_filter_linear)), C.byref(self._h)))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.ezrt_scene_destroy(h)

    __del__ = close

    def render(self, cfg, framebuffer=None):
        """render(width, height, spp) -> framebuffer: `spp` display() calls through HOST buffers
        (H2D of lastFrame when first_frame > 0, D2H of the result, synchronous)."""
        n = partition_pixels(cfg.width, cfg.height, cfg.part_rank, cfg.part_count)
        if framebuffer is None:
            framebuffer = np.zeros((n, cfg.out_channels), dtype=np.float32)
        assert framebuffer.dtype == np.float32 and framebuffer.size == n * cfg.out_channels and framebuffer.flags.c_contiguous
        p = cfg.to_struct()
        check(lib.ezrt_render(self._h, C.byref(p), _fp(framebuffer)))
        if cfg.part_count == 1:
            return framebuffer.reshape(cfg.height, cfg.width, cfg.out_channels)
        return framebuffer

    def render_device(self, cfg, d_framebuffer, stream=None):
        """Enqueue the render on a CUDA stream into a device buffer (torch tensor or raw pointer)."""
        ptr = d_framebuffer.data_ptr() if hasattr(d_framebuffer, "data_ptr") else int(d_framebuffer)
        st = 0 if stream is None else (stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
        p = cfg.to_struct()
        check(lib.ezrt_render_device(self._h, C.byref(p), C.c_void_p(ptr), C.c_void_p(st)))
        return d_framebuffer

    def counters(self):
        c = Counters()
        check(lib.ezrt_get_counters(self._h, C.byref(c)))
        return c

    def kernel_times(self):
        """{class: (ms, launches)} of the last render with cfg.profile = 1."""
        ms = (C.c_double * 4)()
        n = (C.c_uint64 * 4)()
        check(lib.ezrt_get_kernel_times(self._h, ms, n))
        return {k: (ms[i], int(n[i])) for i, k in enumerate(("extend", "shade", "shadow", "other"))}

    def trace_rays(self, origins, dirs, traverse=TRAVERSE_ACCEL, any_hit=False, p3_normal_fudge=False):
        """hitBVH for n rays on the device (P5/fsh:254-306)."""
        o = _f32(origins, (-1, 3))
        d = _f32(dirs, (-1, 3))
        n = o.shape[0]
        hit = np.zeros(n, dtype=np.int32); tri = np.zeros(n, dtype=np.int32); inside = np.zeros(n, dtype=np.int32)
        dist = np.zeros(n, dtype=np.float32); point = np.zeros((n, 3), dtype=np.float32); normal = np.zeros((n, 3), dtype=np.float32)
        ip = lambda a: a.ctypes.data_as(_lib.c_int32_p)
        check(lib.ezrt_trace_rays(self._h, n, _fp(o), _fp(d), int(traverse), int(bool(any_hit)), int(bool(p3_normal_fudge)),
                                  ip(hit), _fp(dist), ip(tri), ip(inside), _fp(point), _fp(normal)))
        return dict(hit=hit, distance=dist, triangle=tri, inside=inside, point=point, normal=normal)


def accel_build(tris, leaf_n=4, where="device", device=0):
    """The binary SAH tree ezrt_scene_create derives its acceleration tree from (ezrt_accel_build, include/ezrt.h):
    returns (links int32 [n,4], boxes float32 [n,6], order uint32 [n_triangles], ms).  where: "device" | "host"."""
    tris = _f32(tris).reshape(-1, 36)
    n = tris.shape[0]
    cap = 2 * n
    links = np.zeros((cap, 4), np.int32)
    boxes = np.zeros((cap, 6), np.float32)
    order = np.zeros(n, np.uint32)
    ms = C.c_double(0.0)
    rc = lib.ezrt_accel_build(device, _fp(tris), n, leaf_n, 1 if where == "host" else 0, links.ctypes.data_as(_lib.c_int32_p), _fp(boxes), cap,
                              order.ctypes.data_as(_lib.c_uint32_p), C.byref(ms))
    if rc < 0:
        check(rc)
    return links[:rc].copy(), boxes[:rc].copy(), order, ms.value


def post_tonemap(d_in, d_out=None, limit=1.5, stream=None):
    """pass3 (P5/shaders/pass3.fsh:14-25) on a device framebuffer (torch CUDA tensor [..., 3|4]) -> [..., 3]."""
    import torch
    channels = d_in.shape[-1]
    n = d_in.numel() // channels
    if d_out is None:
        d_out = torch.empty(tuple(d_in.shape[:-1]) + (3,), dtype=torch.float32, device=d_in.device)
    st = torch.cuda.current_stream() if stream is None else stream
    check(lib.ezrt_post_tonemap(C.c_void_p(d_in.data_ptr()), channels, C.c_void_p(d_out.data_ptr()), n, float(limit), C.c_void_p(st.cuda_stream)))
    return d_out


def write_png(path, framebuffer, tonemap=True):
    """Write a linear framebuffer [H, W, 3|4] (row 0 = bottom) as an 8-bit PNG (pass3 tone map + gamma when tonemap)."""
    fb = _f32(framebuffer)
    h, w, c = fb.shape
    check(lib.ezrt_write_png(str(path).encode(), _fp(fb), w, h, c, int(bool(tonemap))))


def eval_brdf(which, V, N, L, xi, materials, device=0):
    V = _f32(V, (-1, 3)); N = _f32(N, (-1, 3))
    L = None if L is None else _f32(L, (-1, 3))
    xi = None if xi is None else _f32(xi, (-1, 3))
    materials = _f32(materials, (-1, 18))
    out = np.zeros_like(V)
    check(lib.ezrt_eval_brdf(device, which, V.shape[0], _fp(V), _fp(N), None if L is None else _fp(L),
                             None if xi is None else _fp(xi), _fp(materials), _fp(out)))
    return out


def eval_math(which, a, b=None, device=0):
    a = _f32(a).reshape(-1)
    b = None if b is None else _f32(b).reshape(-1)
    out = np.zeros_like(a)
    check(lib.ezrt_eval_math(device, which, a.size, _fp(a), None if b is None else _fp(b), _fp(out)))
    return out
