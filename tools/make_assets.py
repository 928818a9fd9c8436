#!/usr/bin/env python
"""Synthetic, seeded assets for the textured parity / stand-in scenes ("placeholder images
generated with a fixed seed").  Own geometry and images - nothing here comes from the reference tree.

    python tools/make_assets.py                # small set -> scenes/assets/{tex,texroom}   (committed)
    python tools/make_assets.py --size 2048    # 2K set   -> scenes/assets/tex2k            (git-ignored, built by build())

The PNG flavours are deliberately mixed (grey, RGB, RGBA, palette, 16-bit grey) so the loader's decoder is
exercised against the decoder the reference links on every storage type it has to convert.
"""
import argparse
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def value_noise(rng, size, octaves=5, base=4, channels=1):
    """Tileable multi-octave value noise in [0, 1], shape (size, size, channels)."""
    out = np.zeros((size, size, channels), np.float64)
    amp, total = 1.0, 0.0
    ys = np.arange(size)
    for o in range(octaves):
        cells = base << o
        if cells > size:
            break
        grid = rng.random((cells, cells, channels))
        t = ys * cells / size
        i0 = np.floor(t).astype(int) % cells
        i1 = (i0 + 1) % cells
        f = t - np.floor(t)
        f = f * f * (3 - 2 * f)
        rows = grid[i0] * (1 - f)[:, None, None] + grid[i1] * f[:, None, None]          # (size, cells, c)
        full = rows[:, i0] * (1 - f)[None, :, None] + rows[:, i1] * f[None, :, None]     # (size, size, c)
        out += amp * full
        total += amp
        amp *= 0.5
    return out / total


def to8(a):
    return np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8)


def make_textures(dst, size, seed=0):
    os.makedirs(dst, exist_ok=True)
    rng = np.random.default_rng(seed)
    # albedo: two-tone checker modulated by colour noise (RGB)
    n = value_noise(rng, size, channels=3)
    yy, xx = np.mgrid[0:size, 0:size]
    checker = (((xx * 8 // size) + (yy * 8 // size)) % 2)[..., None]
    albedo = np.clip(0.15 + 0.75 * n * (0.55 + 0.45 * checker), 0, 1)
    Image.fromarray(to8(albedo), "RGB").save(os.path.join(dst, "albedo.png"))
    # second albedo (warmer), stored RGBA with a varying alpha channel (alpha must be ignored by the renderer)
    n2 = value_noise(rng, size, channels=3)
    warm = np.clip(n2 * np.array([0.9, 0.6, 0.35]) + 0.08, 0, 1)
    alpha = value_noise(rng, size, channels=1)
    Image.fromarray(to8(np.concatenate([warm, alpha], -1)), "RGBA").save(os.path.join(dst, "albedo_warm.png"))
    # specular tint: palette PNG (quantised to 64 colours)
    spec = np.clip(0.55 + 0.45 * value_noise(rng, size, octaves=3, channels=3), 0, 1)
    Image.fromarray(to8(spec), "RGB").quantize(64, method=Image.Quantize.MEDIANCUT, dither=Image.Dither.NONE).save(
        os.path.join(dst, "specular.png"))
    # glossy / sigma_a map (RGB)
    glos = np.clip(0.35 + 0.65 * value_noise(rng, size, octaves=4, channels=3), 0, 1)
    Image.fromarray(to8(glos), "RGB").save(os.path.join(dst, "glossy.png"))
    # height field -> tangent-space normal map (RGB, OpenGL convention)
    hgt = value_noise(rng, size, octaves=6, channels=1)[..., 0]
    dx = (np.roll(hgt, -1, 1) - np.roll(hgt, 1, 1)) * size / 32.0
    dy = (np.roll(hgt, -1, 0) - np.roll(hgt, 1, 0)) * size / 32.0
    nrm = np.stack([-dx, -dy, np.ones_like(hgt)], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    Image.fromarray(to8(nrm * 0.5 + 0.5), "RGB").save(os.path.join(dst, "normal.png"))
    # roughness maps: 8-bit grey and an RGB one (converted to luma on load)
    r1 = 0.02 + 0.45 * value_noise(rng, size, octaves=4, channels=1)[..., 0]
    Image.fromarray(to8(r1), "L").save(os.path.join(dst, "rough_x.png"))
    r2 = 0.02 + 0.35 * value_noise(rng, size, octaves=4, channels=3)
    Image.fromarray(to8(r2), "RGB").save(os.path.join(dst, "rough_y.png"))
    # ior map: 16-bit grey
    ior = 0.2 + 0.3 * value_noise(rng, size, octaves=3, channels=1)[..., 0]
    Image.fromarray(np.clip(np.rint(ior * 65535.0), 0, 65535).astype(np.uint16)).save(os.path.join(dst, "ior16.png"))
    # emission pattern for the area lights (RGB)
    es = max(size // 4, 16)
    ey, ex = np.mgrid[0:es, 0:es] / float(es)
    em = np.stack([0.6 + 0.4 * np.sin(6.28318 * 3 * ex), 0.6 + 0.4 * np.cos(6.28318 * 2 * ey), 0.7 + 0.3 * np.sin(6.28318 * (ex + ey))], -1)
    Image.fromarray(to8(np.clip(em, 0, 1)), "RGB").save(os.path.join(dst, "emission.png"))
    # lat-long environment: sky gradient + sun + ground + cloud noise (RGB, 2:1)
    ew, eh = 2 * size, size
    v, u = np.mgrid[0:eh, 0:ew]
    theta = (v + 0.5) / eh * np.pi
    phi = (u + 0.5) / ew * 2 * np.pi
    up = np.cos(theta)
    sky = np.stack([0.35 + 0.25 * (1 - up), 0.5 + 0.25 * (1 - up), 0.85 + 0.1 * up], -1) * np.clip(up * 2 + 0.6, 0.2, 1)[..., None]
    ground = np.array([0.22, 0.2, 0.17]) * (0.6 + 0.4 * value_noise(rng, ew, octaves=4, channels=1)[:eh, :, :])
    env = np.where((up > 0)[..., None], sky, ground)
    sun_dir = np.array([np.sin(0.9) * np.cos(1.1), np.sin(0.9) * np.sin(1.1), np.cos(0.9)])
    d = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], -1)
    sun = np.clip((d @ sun_dir - 0.985) / 0.015, 0, 1)[..., None]
    clouds = value_noise(rng, ew, octaves=5, channels=1)[:eh, :, :]
    env = np.clip(env * (0.8 + 0.4 * clouds) + sun * np.array([1.0, 0.95, 0.8]), 0, 1)
    Image.fromarray(to8(env), "RGB").save(os.path.join(dst, "envmap.png"))


# ---- geometry: an open room with UV-mapped walls, two boxes, a UV sphere and two light quads -------------------
def write_obj(path, verts, uvs, norms, faces):
    """faces: list of ((v, vt, vn) x 3), 0-based indices."""
    with open(path, "w") as f:
        f.write("# synthetic mesh written by tools/make_assets.py\n")
        for v in verts:
            f.write("v %.6f %.6f %.6f\n" % tuple(v))
        for t in uvs:
            f.write("vt %.6f %.6f\n" % tuple(t))
        for n in norms:
            f.write("vn %.6f %.6f %.6f\n" % tuple(n))
        for tri in faces:
            f.write("f " + " ".join("%d/%d/%d" % (a + 1, b + 1, c + 1) for a, b, c in tri) + "\n")


def quad(p0, p1, p2, p3, normal, uv_scale=1.0, uv_offset=(0.0, 0.0)):
    verts = [p0, p1, p2, p3]
    uvs = [(uv_offset[0], uv_offset[1]), (uv_offset[0] + uv_scale, uv_offset[1]),
           (uv_offset[0] + uv_scale, uv_offset[1] + uv_scale), (uv_offset[0], uv_offset[1] + uv_scale)]
    faces = [((0, 0, 0), (1, 1, 0), (2, 2, 0)), ((0, 0, 0), (2, 2, 0), (3, 3, 0))]
    return verts, uvs, [normal], faces


def box(center, half, yaw, uv_scale=1.0):
    c, s = np.cos(yaw), np.sin(yaw)
    rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    verts, uvs, norms, faces = [], [], [], []
    axes = [(0, 1, 2), (1, 2, 0), (2, 0, 1)]
    for axis, (a, b, n_axis) in enumerate(axes):
        for sign in (-1.0, 1.0):
            nrm = np.zeros(3)
            nrm[n_axis] = sign
            corners = []
            for da, db in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
                p = np.zeros(3)
                p[a], p[b], p[n_axis] = da * half[a], db * half[b], sign * half[n_axis]
                corners.append(p)
            if np.dot(np.cross(corners[1] - corners[0], corners[2] - corners[0]), nrm) < 0:
                corners = corners[::-1]
            base, nb, tb = len(verts), len(norms), len(uvs)
            verts += [tuple(rot @ p + center) for p in corners]
            norms.append(tuple(rot @ nrm))
            uvs += [(0, 0), (uv_scale, 0), (uv_scale, uv_scale), (0, uv_scale)]
            faces += [((base, tb, nb), (base + 1, tb + 1, nb), (base + 2, tb + 2, nb)),
                      ((base, tb, nb), (base + 2, tb + 2, nb), (base + 3, tb + 3, nb))]
    return verts, uvs, norms, faces


def uv_sphere(center, radius, nu=48, nv=24, uv_scale=2.0):
    verts, uvs, norms, faces = [], [], [], []
    for j in range(nv + 1):
        th = np.pi * j / nv
        for i in range(nu + 1):
            ph = 2 * np.pi * i / nu
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            verts.append(tuple(center + radius * n))
            norms.append(tuple(n))
            uvs.append((uv_scale * i / nu, uv_scale * j / nv))
    for j in range(nv):
        for i in range(nu):
            a, b = j * (nu + 1) + i, j * (nu + 1) + i + 1
            c, d = a + nu + 1, b + nu + 1
            if j > 0:
                faces.append(((a, a, a), (b, b, b), (c, c, c)))
            if j < nv - 1:
                faces.append(((b, b, b), (d, d, d), (c, c, c)))
    return verts, uvs, norms, faces


def make_texroom(dst):
    os.makedirs(dst, exist_ok=True)
    S = 5.5
    write_obj(os.path.join(dst, "floor.obj"), *quad((0, 0, 0), (0, 0, S), (S, 0, S), (S, 0, 0), (0, 1, 0), 2.0))
    write_obj(os.path.join(dst, "ceiling.obj"), *quad((0, S, 0), (S, S, 0), (S, S, S), (0, S, S), (0, -1, 0), 1.0))
    write_obj(os.path.join(dst, "back.obj"), *quad((0, 0, S), (0, S, S), (S, S, S), (S, 0, S), (0, 0, -1), 1.5, (0.25, 0.1)))
    write_obj(os.path.join(dst, "left.obj"), *quad((S, 0, 0), (S, 0, S), (S, S, S), (S, S, 0), (-1, 0, 0), 1.0))
    write_obj(os.path.join(dst, "right.obj"), *quad((0, 0, 0), (0, S, 0), (0, S, S), (0, 0, S), (1, 0, 0), 3.0))
    write_obj(os.path.join(dst, "light.obj"), *quad((2.1, S - 0.02, 2.2), (3.4, S - 0.02, 2.2), (3.4, S - 0.02, 3.3), (2.1, S - 0.02, 3.3), (0, -1, 0), 1.0))
    write_obj(os.path.join(dst, "spot.obj"), *quad((0.6, S - 0.03, 0.8), (1.4, S - 0.03, 0.8), (1.4, S - 0.03, 1.6), (0.6, S - 0.03, 1.6), (0, -1, 0), 1.0))
    write_obj(os.path.join(dst, "tall_box.obj"), *box(np.array([3.7, 1.65, 3.5]), (0.8, 1.65, 0.8), 0.3, 1.0))
    write_obj(os.path.join(dst, "short_box.obj"), *box(np.array([1.6, 0.8, 1.7]), (0.8, 0.8, 0.8), -0.3, 1.0))
    write_obj(os.path.join(dst, "ball.obj"), *uv_sphere(np.array([1.7, 2.4, 1.7]), 0.8))
    write_obj(os.path.join(dst, "ball2.obj"), *uv_sphere(np.array([4.3, 0.7, 1.2]), 0.7, 32, 16, 1.0))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    assets = os.path.join(ROOT, "scenes", "assets")
    if args.size <= 256:
        make_textures(args.out or os.path.join(assets, "tex"), args.size)
        make_texroom(os.path.join(assets, "texroom"))
    else:
        make_textures(args.out or os.path.join(assets, "tex2k"), args.size)
