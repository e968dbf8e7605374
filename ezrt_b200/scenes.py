"""Synthetic scenes for the benchmark configurations of BASELINE.json.

The named scenes use the Stanford bunny.  /root/reference (and its OBJ files) does not exist on the GPU box, so the
bunny travels as an indexed mesh asset (ezrt_b200/data/bunny.npz, 2503 vertices / 4968 faces, recovered from the
committed P3 scene arrays by tests/golden/make_bunny_asset.py); it is emitted as OBJ text and pushed through the same
readObj() parser, unit-box normalisation, transform and smooth-normal code as any mesh file.  sphere.obj (320
triangles) and quad.obj (a 12-triangle box) are regenerated procedurally with the same triangle counts.

  s_p3_bunny()   P3/main.cpp:690-701: bunny (4968) + floor (12) + emissive sphere (320) = 5300 tris     [C1, C2]
  s_1m_bunny()   SURVEY 8(d) "S-1M": 201 bunnies in the first 201 slots of a 15x14 grid + 4 emissive spheres + floor
                 = 201*4968 + 1280 + 12 = 999,860 triangles                                              [C3, C4, C5]

Second workload family kept from round 1 (a procedural stand-in mesh, "blob" = bumpy icosphere, 5120 triangles;
geometry uses only + - * / sqrt and integer hashing, so the OBJ text is identical on every machine):

  s_bunny()  blob (5120) + floor box (12) + emissive sphere (320) = 5452 tris
  s_1m()     195 blobs on a 15x13 grid + 4 emissive spheres + floor = 999,692 tris
"""
import os
import numpy as np

from . import api
from .api import Material, TriangleList, transform_matrix


def _wang(seed):
    seed = ((seed ^ 61) ^ (seed >> 16)) & 0xFFFFFFFF
    seed = (seed * 9) & 0xFFFFFFFF
    seed = (seed ^ (seed >> 4)) & 0xFFFFFFFF
    seed = (seed * 0x27d4eb2d) & 0xFFFFFFFF
    seed = (seed ^ (seed >> 15)) & 0xFFFFFFFF
    return seed


def _unit_randoms(seed, n):
    out = []
    s = seed | 1
    for _ in range(n):
        s = _wang(s)
        out.append(s / 4294967296.0)
    return out


def icosphere(subdiv):
    """Vertices (unit sphere, float64) and faces of an icosahedron subdivided `subdiv` times: 20*4^subdiv faces."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.sqrt(1.0 + t * t) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
             (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
             (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        cache = {}
        new_faces = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                m = m / np.sqrt(np.dot(m, m))
                verts.append(m)
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    return np.array(verts), faces


def _obj_text(verts, faces):
    lines = ["v %.6f %.6f %.6f" % (p[0], p[1], p[2]) for p in verts]
    lines += ["f %d %d %d" % (a + 1, b + 1, c + 1) for a, b, c in faces]
    return "\n".join(lines) + "\n"


_CACHE = {}


def blob_obj(subdiv=4, seed=7):
    """Bumpy icosphere: r = 1 + sum_k a_k * max(0, 1 - |v-c_k|^2 / r_k^2)^2 (polynomial bumps)."""
    key = ("blob", subdiv, seed)
    if key not in _CACHE:
        verts, faces = icosphere(subdiv)
        u = _unit_randoms(seed * 7919 + 13, 5 * 12)
        r = np.ones(len(verts))
        for k in range(12):
            c = np.array([u[5 * k] * 2 - 1, u[5 * k + 1] * 2 - 1, u[5 * k + 2] * 2 - 1])
            c = c / np.sqrt(np.dot(c, c) + 1e-9)
            amp = 0.15 + 0.3 * u[5 * k + 3]
            if k % 3 == 2:
                amp = -0.5 * amp
            rad2 = (0.35 + 0.5 * u[5 * k + 4]) ** 2
            d2 = ((verts - c) ** 2).sum(axis=1)
            w = np.maximum(0.0, 1.0 - d2 / rad2)
            r = r + amp * w * w
        _CACHE[key] = _obj_text(verts * r[:, None] * np.array([1.0, 1.15, 0.85]), faces)
    return _CACHE[key]


def bunny_obj():
    """The Stanford bunny asset as OBJ text ("%.9g" round-trips fp32)."""
    key = ("bunny",)
    if key not in _CACHE:
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "bunny.npz"))
        lines = ["v %.9g %.9g %.9g" % (p[0], p[1], p[2]) for p in z["verts"]]
        lines += ["f %d %d %d" % (a + 1, b + 1, c + 1) for a, b, c in z["faces"]]
        _CACHE[key] = "\n".join(lines) + "\n"
    return _CACHE[key]


def _unit_ymin(text):
    """Lowest y of a mesh after readObj's unit-box normalisation (identity transform): instances are set on the floor with it."""
    key = ("ymin", hash(text))
    if key not in _CACHE:
        tl = TriangleList()
        tl.read_obj_text(text, Material(), transform_matrix(), False)
        _CACHE[key] = float(tl.encode_triangles()[:, :9].reshape(-1, 3)[:, 1].min())
    return _CACHE[key]


def sphere_obj(subdiv=2):
    """320-triangle sphere (same count as the reference's sphere.obj)."""
    key = ("sphere", subdiv)
    if key not in _CACHE:
        verts, faces = icosphere(subdiv)
        _CACHE[key] = _obj_text(verts, faces)
    return _CACHE[key]


def box_obj():
    """Unit cube centred at the origin, 8 v / 12 f (stands for the reference's quad.obj box)."""
    v = [(-.5, -.5, -.5), (.5, -.5, -.5), (.5, .5, -.5), (-.5, .5, -.5), (-.5, -.5, .5), (.5, -.5, .5), (.5, .5, .5), (-.5, .5, .5)]
    f = [(0, 2, 1), (0, 3, 2), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (3, 7, 6), (3, 6, 2), (0, 4, 7), (0, 7, 3), (1, 2, 6), (1, 6, 5)]
    return _obj_text(np.array(v, dtype=np.float64), f)


# Eight Disney parameter sets spanning the reference's scene set-ups (P4/main.cpp:690-727, P5/main.cpp:803-817)
MATERIAL_PRESETS = [
    Material(baseColor=(1.0, 0.73, 0.25), roughness=0.5, specular=1.0, metallic=1.0, clearcoat=1.0, clearcoatGloss=0.0),
    Material(baseColor=(1.0, 0.5, 0.5), roughness=0.1, metallic=0.0, clearcoat=1.0, subsurface=1.0),
    Material(baseColor=(0.75, 0.7, 0.15), roughness=0.15, metallic=1.0, clearcoat=1.0),
    Material(baseColor=(0.5, 0.5, 1.0), roughness=0.1, metallic=0.0, clearcoat=1.0),
    Material(baseColor=(0.725, 0.71, 0.68), roughness=0.5),
    Material(baseColor=(0.2, 0.8, 0.3), roughness=0.3, metallic=0.7, sheen=0.5, sheenTint=0.8),
    Material(baseColor=(0.9, 0.9, 0.9), roughness=0.05, metallic=0.9, specularTint=0.5, anisotropic=0.6),
    Material(baseColor=(0.8, 0.3, 0.1), roughness=0.7, subsurface=0.5, clearcoat=0.5, clearcoatGloss=0.3),
]


def bunny_meshes():
    """P3's scene with the blob in place of the bunny (P3/main.cpp:688-701) as readObj calls:
    [(obj text, Material, trans, smooth)]"""
    return [
        (blob_obj(), Material(baseColor=(1, 1, 1)), transform_matrix((0, 0, 0), (0.3, -0.65, 0.0), (1.5, 1.5, 1.5)), True),
        (box_obj(), Material(baseColor=(0.725, 0.71, 0.68)), transform_matrix((0, 0, 0), (0, -1.4, 0), (18.83, 0.01, 18.83)), False),
        (sphere_obj(), Material(baseColor=(1, 1, 1), emissive=(30, 20, 10)), transform_matrix((0, 0, 0), (0.0, 0.9, 0.0), (1, 1, 1)), False),
    ]


def s_bunny(builder=api.BVH_SAH_FAST):
    """bunny_meshes() built with the default leaf size.  Returns (tris, nodes, eye, cam)."""
    tl = TriangleList()
    for text, m, trans, smooth in bunny_meshes():
        tl.read_obj_text(text, m, trans, smooth)
    tris, nodes = tl.build_bvh(8, builder)
    eye, cam = api.camera_orbit(0.0, 0.0, 4.0)  # P3/main.cpp:148-150
    return tris, nodes, eye, cam


def grid_meshes(nx, nz, n_lights=4, pitch=1.2, mesh="blob", count=None):
    """Instances of `mesh` ("blob" | "bunny") in the first `count` (default all) row-major slots of an nx x nz grid
    (y-rotation 360*u0 and scale 0.8+0.4*u1 from successive wang_hash outputs of seed (id*9781+1)|1, material preset
    id%8), `n_lights` emissive spheres above it and a floor box, as readObj calls: [(obj text, Material, trans, smooth)]"""
    out = []
    text = blob_obj() if mesh == "blob" else bunny_obj()
    ymin = None if mesh == "blob" else _unit_ymin(text)
    count = nx * nz if count is None else count
    inst = 0
    for iz in range(nz):
        for ix in range(nx):
            if inst >= count:
                break
            u = _unit_randoms(inst * 9781 + 1, 2)
            rot = 360.0 * u[0]
            sc = 0.8 + 0.4 * u[1]
            x = (ix - (nx - 1) / 2.0) * pitch
            z = (iz - (nz - 1) / 2.0) * pitch
            m = MATERIAL_PRESETS[inst % 8]
            y = (-1.4 + 0.6 * sc) if ymin is None else (-1.395 - sc * ymin)  # feet on the floor (top face at y = -1.395)
            out.append((text, m, transform_matrix((0, rot, 0), (x, y, z), (sc, sc, sc)), True))
            inst += 1
    light = Material(baseColor=(1, 1, 1), emissive=(20, 20, 20))
    sph = sphere_obj()
    span_x, span_z = nx * pitch, nz * pitch
    for k in range(n_lights):
        lx = (((k % 2) * 2 - 1) * 0.25) * span_x
        lz = (((k // 2) * 2 - 1) * 0.25) * span_z
        out.append((sph, light, transform_matrix((0, 0, 0), (lx, 2.5, lz), (1.5, 1.5, 1.5)), False))
    floor = Material(baseColor=(0.725, 0.71, 0.68), roughness=0.3, metallic=0.1)
    ext = max(span_x, span_z) * 1.5 + 4.0
    out.append((box_obj(), floor, transform_matrix((0, 0, 0), (0, -1.4, 0), (ext, 0.01, ext)), False))
    return out


def s_grid(nx, nz, n_lights=4, builder=api.BVH_SAH_FAST, pitch=1.2, mesh="blob", count=None):
    """grid_meshes() built with the default leaf size.  Returns (tris, nodes, eye, cam)."""
    tl = TriangleList()
    for text, m, trans, smooth in grid_meshes(nx, nz, n_lights, pitch, mesh, count):
        tl.read_obj_text(text, m, trans, smooth)
    tris, nodes = tl.build_bvh(8, builder)
    r = 0.62 * max(nx * pitch, nz * pitch) + 3.0
    eye, cam = api.camera_orbit(30.0, 25.0, r)
    return tris, nodes, eye, cam


def s_1m(builder=api.BVH_SAH_FAST):
    """195 blobs (15 x 13) + 4 spheres + floor = 195*5120 + 1280 + 12 = 999,692 triangles."""
    return s_grid(15, 13, 4, builder)


def s_1m_bunny(builder=api.BVH_SAH_FAST):
    """S-1M of SURVEY.md 8(d): 201 Stanford bunnies (15 x 14 grid, first 201 slots) + 4 spheres + floor = 999,860 triangles,
    camera orbit rotatAngle 30, upAngle 25."""
    return s_grid(15, 14, 4, builder, mesh="bunny", count=201)


def p3_bunny_meshes():
    """P3's scene block (P3/main.cpp:690-701) with the real bunny."""
    return [
        (bunny_obj(), Material(baseColor=(1, 1, 1)), transform_matrix((0, 0, 0), (0.3, -1.6, 0.0), (1.5, 1.5, 1.5)), True),
        (box_obj(), Material(baseColor=(0.725, 0.71, 0.68)), transform_matrix((0, 0, 0), (0, -1.4, 0), (18.83, 0.01, 18.83)), False),
        (sphere_obj(), Material(baseColor=(1, 1, 1), emissive=(30, 20, 10)), transform_matrix((0, 0, 0), (0.0, 0.9, 0.0), (1, 1, 1)), False),
    ]


def s_p3_bunny(builder=api.BVH_SAH_FAST):
    """Stanford bunny 4968 + floor 12 + emissive sphere 320 = 5300 triangles, eye (0,0,4) (P3/main.cpp:148-150)."""
    tl = TriangleList()
    for text, m, trans, smooth in p3_bunny_meshes():
        tl.read_obj_text(text, m, trans, smooth)
    tris, nodes = tl.build_bvh(8, builder)
    eye, cam = api.camera_orbit(0.0, 0.0, 4.0)
    return tris, nodes, eye, cam


def synth_hdr(width=512, height=256, seed=3, n_lamps=16):
    """Procedural environment map: vertical gradient 0.2 -> 1.0 plus polynomial-falloff 'lamps' (peak 500)."""
    ys = (np.arange(height, dtype=np.float64) + 0.5) / height
    xs = (np.arange(width, dtype=np.float64) + 0.5) / width
    base = 1.0 - 0.8 * ys  # row 0 = top (zenith) bright
    img = np.zeros((height, width, 3), dtype=np.float64)
    img[:, :, 0] = base[:, None] * 0.9
    img[:, :, 1] = base[:, None] * 0.95
    img[:, :, 2] = base[:, None] * 1.0
    u = _unit_randoms(seed * 104729 + 7, 4 * n_lamps)
    for k in range(n_lamps):
        cx, cy = u[4 * k], 0.05 + 0.45 * u[4 * k + 1]
        rad = 0.01 + 0.03 * u[4 * k + 2]
        peak = 500.0 * (0.3 + 0.7 * u[4 * k + 3])
        dx = np.abs(xs - cx)
        dx = np.minimum(dx, 1.0 - dx)
        d2 = (dx[None, :] ** 2) * 4.0 + (ys[:, None] - cy) ** 2
        w = np.maximum(0.0, 1.0 - d2 / (rad * rad))
        lamp = peak * w * w
        img[:, :, 0] += lamp
        img[:, :, 1] += lamp * 0.9
        img[:, :, 2] += lamp * 0.7
    return img.astype(np.float32)
