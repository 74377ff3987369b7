"""Scene / camera-path descriptions shared by the bench driver and the tests.

A scene is a plain dict of float32 data (no classes), mirroring what the reference builds through
``Scene::addSphere/addTriangle/addLight/addTexture/setSkyboxTexture`` (reference Scene.h:29-39):

    {"ambient": (rgb[3], power),
     "skybox": None | HxW uint32 ARGB array,            # None = texture failed to load -> checker fallback
     "textures": [None | HxW uint32 ARGB array, ...],   # Scene::addTexture order
     "lights": [(origin[3], radius, rgb[3], power), ...],
     "objects": [("sphere", center[3], radius, mtype, rgb[3], refl, transp) |
                 ("tri", v[9], mtype, rgb[3], refl, transp, tex_index_or_-1, uv[6]) |
                 ("plane", pos[3], norm[3], mtype, rgb[3], refl, transp), ...]}   # insertion order matters (ties)

Material types follow reference Material.h:8: 0 = mtMetal, 1 = mtDielectric.
Nothing here touches the GPU; :mod:`reflaxman_b200.capi` uploads a scene through the C ABI.
"""
from __future__ import annotations

import math
import struct

import numpy as np

MT_METAL = 0
MT_DIELECTRIC = 1

DEFAULT_EYE = (7.427, 3.494, -3.773)       # reference Render.cpp:30
DEFAULT_AT = (6.5981, 3.127, -3.352)
DEFAULT_FOV = 1.05
SCREENSHOT_REFLECTIONS = 20                # reference defaults.h:9
STATIC_REFLECTIONS = 15                    # reference defaults.h:7
MOTION_REFLECTIONS = 4                     # reference defaults.h:8


def f32(x):
    return np.asarray(x, dtype=np.float32)


def default_scene(skybox=None, floor=None):
    """The reference's built-in demo scene (reference Render.cpp:25-55).

    ``skybox`` / ``floor``: optional uint32 ARGB arrays standing in for textures/skybox.tga and
    textures/himiya.tga; ``None`` reproduces this checkout's behaviour (files missing -> checker fallback).
    """
    M, D = MT_METAL, MT_DIELECTRIC
    objs = [
        ("sphere", (-1.25, 1.5, -0.25), 1.5, M, (1.0, 1.0, 1.0), 1.0, 0.0),
        ("sphere", (0.15, 1.0, 1.75), 1.0, M, (1.0, 1.0, 1.0), 0.95, 0.0),
        ("sphere", (-3.0, 0.6, -3.0), 0.6, D, (1.0, 1.0, 1.0), 0.0, 0.0),
        ("sphere", (-0.5, 0.5, -2.5), 0.5, D, (0.5, 1.0, 0.15), 0.75, 0.0),
        ("sphere", (1.0, 0.4, -1.5), 0.4, D, (0.0, 0.5, 1.0), 1.0, 0.0),
        ("sphere", (1.8, 0.4, 0.1), 0.4, M, (1.0, 0.65, 0.45), 1.0, 0.0),
        ("sphere", (1.7, 0.5, 1.9), 0.5, M, (1.0, 0.90, 0.60), 0.75, 0.0),
        ("sphere", (0.6, 0.6, 4.2), 0.6, M, (0.9, 0.9, 0.9), 0.0, 0.0),
        ("tri", (-14.0, 0.0, -10.0, -14.0, 0.0, 10.0, 14.0, 0.0, -10.0), D, (1.0, 1.0, 1.0), 0.95, 0.0, 0,
         (0.0, 0.0, 0.0, 1.0, 1.0, 0.0)),
        ("tri", (-14.0, 0.0, 10.0, 14.0, 0.0, 10.0, 14.0, 0.0, -10.0), D, (1.0, 1.0, 1.0), 0.95, 0.0, 0,
         (0.0, 1.0, 1.0, 1.0, 1.0, 0.0)),
    ]
    return {
        "ambient": ((0.95, 0.95, 1.0), 0.15),
        "skybox": skybox,
        "textures": [floor],
        "lights": [((11.8e9, 4.26e9, 3.08e9), 3.48e8, (1.0, 1.0, 0.95), 0.85)],
        "objects": objs,
    }


class _Lcg:
    """Generator fixed by SURVEY §8(d) config 4: s = s*1664525 + 1013904223, u = (s >> 8) / 2^24."""

    def __init__(self, seed=0x9E3779B9):
        self.s = seed & 0xFFFFFFFF

    def u(self):
        self.s = (self.s * 1664525 + 1013904223) & 0xFFFFFFFF
        return (self.s >> 8) / float(1 << 24)


def synthetic_scene(n_side=32, floor=None, skybox=None, seed=0x9E3779B9, back_wall=True):
    """Config 4: n_side^2 random reflective spheres on a jittered grid over the 28x20 textured floor.

    x in [-13,13], z in [-9,9], r in [0.1,0.3], y = r; alternating metal/dielectric; colour in [0.3,1]^3;
    reflectivity 1.0 w.p. 1/2 else U(0,1).  Floor (and optional back wall) are textured triangle pairs
    (the reference's Plane is untextured and unreachable through Scene).
    """
    g = _Lcg(seed)
    objs = []
    for iz in range(n_side):
        for ix in range(n_side):
            cellx, cellz = 26.0 / n_side, 18.0 / n_side
            r = 0.1 + 0.2 * g.u()
            r = min(r, 0.45 * min(cellx, cellz))
            x = -13.0 + (ix + 0.5) * cellx + (g.u() - 0.5) * (cellx - 2 * r)
            z = -9.0 + (iz + 0.5) * cellz + (g.u() - 0.5) * (cellz - 2 * r)
            col = (0.3 + 0.7 * g.u(), 0.3 + 0.7 * g.u(), 0.3 + 0.7 * g.u())
            refl = 1.0 if g.u() < 0.5 else g.u()
            mt = MT_METAL if (ix + iz) % 2 == 0 else MT_DIELECTRIC
            objs.append(("sphere", (x, r, z), r, mt, col, refl, 0.0))
    D = MT_DIELECTRIC
    objs.append(("tri", (-14.0, 0.0, -10.0, -14.0, 0.0, 10.0, 14.0, 0.0, -10.0), D, (1.0, 1.0, 1.0), 0.95, 0.0, 0,
                 (0.0, 0.0, 0.0, 1.0, 1.0, 0.0)))
    objs.append(("tri", (-14.0, 0.0, 10.0, 14.0, 0.0, 10.0, 14.0, 0.0, -10.0), D, (1.0, 1.0, 1.0), 0.95, 0.0, 0,
                 (0.0, 1.0, 1.0, 1.0, 1.0, 0.0)))
    if back_wall:
        # wall at x = -14 facing +x (normal = (v1-v0) x (v2-v0))
        objs.append(("tri", (-14.0, 0.0, -10.0, -14.0, 8.0, -10.0, -14.0, 0.0, 10.0), D, (1.0, 1.0, 1.0), 0.5, 0.0, 0,
                     (0.0, 0.0, 0.0, 1.0, 1.0, 0.0)))
        objs.append(("tri", (-14.0, 8.0, -10.0, -14.0, 8.0, 10.0, -14.0, 0.0, 10.0), D, (1.0, 1.0, 1.0), 0.5, 0.0, 0,
                     (0.0, 1.0, 1.0, 1.0, 1.0, 0.0)))
    return {
        "ambient": ((0.95, 0.95, 1.0), 0.15),
        "skybox": skybox,
        "textures": [floor],
        "lights": [((11.8e9, 4.26e9, 3.08e9), 3.48e8, (1.0, 1.0, 0.95), 0.85)],
        "objects": objs,
    }


def synthetic_texture(w, h, seed, alpha=True):
    """Seeded synthetic ARGB texture (smooth gradients + hashed detail) — stands in for the missing TGA blobs."""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.uint32)
    v = (xs * np.uint32(2654435761) ^ (ys * np.uint32(40503) + np.uint32(seed))) * np.uint32(2246822519)
    v ^= v >> np.uint32(13)
    r = ((xs * 255 // max(w - 1, 1)) + (v & np.uint32(31))) & np.uint32(255)
    g = ((ys * 255 // max(h - 1, 1)) + ((v >> np.uint32(8)) & np.uint32(31))) & np.uint32(255)
    b = (((xs // 16 + ys // 16) % 2) * 128 + ((v >> np.uint32(16)) & np.uint32(127))) & np.uint32(255)
    a = np.uint32(0xFF000000) if alpha else np.uint32(0)
    return (a | (r << np.uint32(16)) | (g << np.uint32(8)) | b).astype(np.uint32)


def write_tga(path, argb, bpp=32):
    """Type-2 uncompressed TGA as the reference loader expects (reference Texture.cpp:34-108, image_headers.h:4-18).

    Row 0 of the file is row 0 of the texture (the loader ignores the origin flag)."""
    h, w = argb.shape
    header = struct.pack("<bbbhhbhhhhbb", 0, 0, 2, 0, 0, 0, 0, 0, w, h, bpp, 0)
    a = np.ascontiguousarray(argb, dtype="<u4").view(np.uint8).reshape(h, w, 4)  # B, G, R, A in memory
    body = a.tobytes() if bpp == 32 else np.ascontiguousarray(a[:, :, :3]).tobytes()
    with open(path, "wb") as f:
        f.write(header)
        f.write(body)


def loaded_texture(argb, bpp=32):
    """What the reference holds in memory after loading ``write_tga(argb, bpp)``: 24-bpp files get alpha 0xFF."""
    if bpp == 24:
        return (argb | np.uint32(0xFF000000)).astype(np.uint32)
    return argb.astype(np.uint32)


def camera_lookat(eye, at, fov=DEFAULT_FOV):
    """``Camera(eye, at, fov)`` (reference Camera.cpp:24-36) in float32, op for op: returns (eye[3], view[9] row-major, fov).

    view columns are ox, oy, oz.  All arithmetic is float32 numpy scalars (IEEE RN, no contraction) so the result is
    bit-identical to the reference's constructor compiled without -ffast-math.
    """
    f = np.float32
    eye = [f(v) for v in eye]
    at = [f(v) for v in at]

    def sub(a, b):
        return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]

    def cross(a, b):
        return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]

    def norm(a):
        ln = np.sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2], dtype=np.float32)
        if ln > f(2.0 ** -63):
            return [a[0] / ln, a[1] / ln, a[2] / ln]
        return a

    up = [f(0.0), f(1.0), f(0.0)]
    oz = norm(sub(at, eye))
    ox = norm(cross(up, oz))
    oy = norm(cross(oz, ox))
    view = [ox[0], oy[0], oz[0], ox[1], oy[1], oz[1], ox[2], oy[2], oz[2]]
    return f32(eye), f32(view), np.float32(fov)


def default_camera():
    return camera_lookat(DEFAULT_EYE, DEFAULT_AT, DEFAULT_FOV)


def orbit_cameras(n_frames=240, fov=DEFAULT_FOV):
    """Config 5: frame k = default eye and at rotated about the world y axis by 2*pi*k/n (frame 0 == default camera)."""
    cams = []
    for k in range(n_frames):
        if k == 0:
            cams.append(default_camera())
            continue
        a = 2.0 * math.pi * k / n_frames
        c, s = math.cos(a), math.sin(a)

        def rot(p):
            return (np.float32(c * p[0] + s * p[2]), np.float32(p[1]), np.float32(-s * p[0] + c * p[2]))

        cams.append(camera_lookat(rot(DEFAULT_EYE), rot(DEFAULT_AT), fov))
    return cams


def _fmt(x):
    return "%.9g" % float(np.float32(x))


def scene_to_text(scene, tex_paths=None, sky_path=None):
    """Serialise for oracle/ref_harness.cpp (test infrastructure).  ``tex_paths[i]`` / ``sky_path``: TGA file
    paths for textures that are present; missing ones are written as '-' (load fails -> checker fallback)."""
    out = []
    rgb, p = scene["ambient"]
    out.append("ambient " + " ".join(_fmt(v) for v in rgb) + " " + _fmt(p))
    out.append("skybox " + (sky_path if (scene.get("skybox") is not None and sky_path) else "-"))
    for o, r, c, p in scene["lights"]:
        out.append("light " + " ".join(_fmt(v) for v in (*o, r, *c, p)))
    for i, t in enumerate(scene["textures"]):
        out.append("texture " + (tex_paths[i] if (t is not None and tex_paths) else "-"))
    for ob in scene["objects"]:
        if ob[0] == "sphere":
            _, c, r, mt, col, refl, tr = ob
            out.append("sphere " + " ".join(_fmt(v) for v in (*c, r)) + " %d " % mt + " ".join(_fmt(v) for v in (*col, refl, tr)))
        elif ob[0] == "tri":
            _, v, mt, col, refl, tr, tex, uv = ob
            out.append("tri " + " ".join(_fmt(x) for x in v) + " %d " % mt + " ".join(_fmt(x) for x in (*col, refl, tr)) +
                       " %d " % tex + " ".join(_fmt(x) for x in uv))
        else:
            raise ValueError("the reference Scene cannot hold a %s (no addPlane)" % ob[0])
    return "\n".join(out) + "\n"


def cameras_to_text(cams):
    lines = []
    for eye, view, fov in cams:
        lines.append("view " + " ".join(_fmt(v) for v in (*eye, *view, fov)))
    return "\n".join(lines) + "\n"
