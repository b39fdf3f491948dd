"""Synthetic scenes (our own generators; the reference ships only OBJ models).

* ``box_scene``  — a small closed room with two blocks and a ceiling light, in the
  spirit of the Cornell models the reference is run on (scripts/run.sh:28-30).
* ``soup_scene`` — BASELINE.json config 4/5: n triangles in G geometries, L quad
  lights under the ceiling, optional analytic spheres (extension).
* ``from_npz``   — a flat scene stored by tests/golden/make_golden.py: the reference's own
  ``model::loadobj`` output for one of its shipped OBJ models (the models themselves live in
  the reference tree and do not travel to the GPU box).
"""
from __future__ import annotations

import numpy as np

from .api import Scene


def _mat(ka=(0, 0, 0), kd=(0.5, 0.5, 0.5), ks=(0, 0, 0), ke=(0, 0, 0), ns=10.0):
    return np.array([*ka, *kd, *ks, *ke, ns], np.float32)


def _quad(a, b, c, d):
    """two triangles (a,b,c), (a,c,d)"""
    return [[a, b, c], [a, c, d]]


def _block(lo, hi):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    p = lambda x, y, z: (x, y, z)
    t = []
    t += _quad(p(x0, y1, z0), p(x0, y1, z1), p(x1, y1, z1), p(x1, y1, z0))  # top
    t += _quad(p(x0, y0, z1), p(x1, y0, z1), p(x1, y1, z1), p(x0, y1, z1))  # front
    t += _quad(p(x1, y0, z0), p(x0, y0, z0), p(x0, y1, z0), p(x1, y1, z0))  # back
    t += _quad(p(x0, y0, z0), p(x0, y0, z1), p(x0, y1, z1), p(x0, y1, z0))  # left
    t += _quad(p(x1, y0, z1), p(x1, y0, z0), p(x1, y1, z0), p(x1, y1, z1))  # right
    return t


def box_scene(specular=False) -> Scene:
    geoms = []
    geoms.append((_quad((-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1)), _mat(kd=(0.72, 0.72, 0.72), ka=(0.1, 0.1, 0.1))))
    geoms.append((_quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1)), _mat(kd=(0.72, 0.72, 0.72))))
    geoms.append((_quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1)), _mat(kd=(0.72, 0.72, 0.72))))
    geoms.append((_quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1)), _mat(kd=(0.14, 0.45, 0.09))))
    geoms.append((_quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1)), _mat(kd=(0.63, 0.065, 0.05))))
    ks = (0.6, 0.6, 0.6) if specular else (0, 0, 0)
    geoms.append((_block((-0.7, 0.0, -0.6), (-0.1, 1.2, -0.05)), _mat(kd=(0.7, 0.7, 0.7), ks=ks, ns=20.0)))
    geoms.append((_block((0.1, 0.0, 0.0), (0.7, 0.6, 0.6)), _mat(kd=(0.7, 0.7, 0.7), ka=(0.05, 0.05, 0.05))))
    geoms.append((_quad((-0.24, 1.98, -0.22), (0.23, 1.98, -0.22), (0.23, 1.98, 0.16), (-0.24, 1.98, 0.16)),
                  _mat(kd=(0.78, 0.78, 0.78), ke=(17, 12, 4))))
    off, verts, mats = [0], [], []
    for tris, m in geoms:
        verts += tris
        off.append(off[-1] + len(tris))
        mats.append(m)
    return Scene(np.array(off), np.array(verts, np.float32), np.array(mats), np.array([len(geoms) - 1]))


def coplanar_scene(grid=12) -> Scene:
    """Stress case for the filter's sign-ambiguous rows: box_scene plus (a) a grid of triangles lying
    exactly in the light's plane (y = 1.98), some sharing the light's own vertices, and (b) a thin fin
    whose plane passes exactly through the default eye (0, 1, 2.9)."""
    base = box_scene()
    tris = []
    xs = np.linspace(-0.9, 0.9, grid + 1)
    zs = np.linspace(-0.9, 0.9, grid + 1)
    for i in range(grid):
        for j in range(grid):
            x0, x1, z0, z1 = xs[i], xs[i + 1], zs[j], zs[j + 1]
            if x1 <= -0.3 or x0 >= 0.3 or z1 <= -0.3 or z0 >= 0.25:
                tris += _quad((x0, 1.98, z0), (x1, 1.98, z0), (x1, 1.98, z1), (x0, 1.98, z1))
    # triangles touching the light's corners, in its plane
    tris.append([(-0.24, 1.98, -0.22), (-0.3, 1.98, -0.22), (-0.24, 1.98, -0.3)])
    tris.append([(0.23, 1.98, 0.16), (0.3, 1.98, 0.16), (0.23, 1.98, 0.25)])
    fin = [[(0.0, 0.2, 0.5), (0.0, 1.6, 0.5), (0.0, 0.9, -0.4)], [(0.0, 0.2, 0.5), (0.0, 0.9, -0.4), (0.0, 0.1, -0.4)]]
    verts = np.concatenate([base.tri_verts[: base.geom_tri_offset[-2]], np.array(tris, np.float32), np.array(fin, np.float32),
                            base.tri_verts[base.geom_tri_offset[-2]:]])
    n0 = int(base.geom_tri_offset[-2])
    off = list(base.geom_tri_offset[:-1]) + [n0 + len(tris), n0 + len(tris) + 2, n0 + len(tris) + 2 + 2]
    mats = np.concatenate([base.geom_material[:-1], _mat(kd=(0.3, 0.3, 0.8))[None], _mat(kd=(0.8, 0.8, 0.2))[None],
                           base.geom_material[-1:]])
    return Scene(np.array(off), verts, mats, np.array([len(off) - 2]))


def soup_scene(n_tris=1_000_000, n_geoms=1000, n_lights=4, n_spheres=0, seed=42, edge=(0.01, 0.02), specular=False,
               with_normals=False) -> Scene:
    """Triangle soup in the box [-1,1] x [0,2] x [-1,1] seen from (0,1,3) looking at (0,1,0).

    The first geometry is a floor quad and a back wall (4 big triangles), the next
    ``n_lights`` geometries are 2-face quad lights just under y=2, and the remaining
    triangles are small (edge length in ``edge``), uniformly placed, split evenly over
    the remaining geometries.  Materials are random (ka, kd in [0,1]^3), ks = 0 unless
    ``specular``.  Everything is float32 and deterministic in ``seed``."""
    rng = np.random.default_rng(seed)
    n_fixed = 4 + 2 * n_lights
    assert n_tris > n_fixed and n_geoms > 1 + n_lights
    verts = np.zeros((n_tris, 3, 3), np.float32)
    verts[0:2] = np.array(_quad((-1.5, 0, -1.5), (-1.5, 0, 1.5), (1.5, 0, 1.5), (1.5, 0, -1.5)), np.float32)
    verts[2:4] = np.array(_quad((-1.5, 0, -1.5), (1.5, 0, -1.5), (1.5, 2.2, -1.5), (-1.5, 2.2, -1.5)), np.float32)
    off = [0, 4]
    for l in range(n_lights):
        cx = -0.75 + 1.5 * (l + 0.5) / n_lights
        cz = 0.3 * (-1) ** l
        q = _quad((cx - 0.1, 1.99, cz - 0.1), (cx + 0.1, 1.99, cz - 0.1), (cx + 0.1, 1.99, cz + 0.1), (cx - 0.1, 1.99, cz + 0.1))
        verts[4 + 2 * l: 6 + 2 * l] = np.array(q, np.float32)
        off.append(off[-1] + 2)
    n_small = n_tris - n_fixed
    c = rng.uniform([-1, 0.02, -1], [1, 1.9, 1], size=(n_small, 3)).astype(np.float32)
    d1 = rng.normal(size=(n_small, 3)).astype(np.float32)
    d2 = rng.normal(size=(n_small, 3)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    l1 = rng.uniform(edge[0], edge[1], size=(n_small, 1)).astype(np.float32)
    l2 = rng.uniform(edge[0], edge[1], size=(n_small, 1)).astype(np.float32)
    verts[n_fixed:, 0] = c
    verts[n_fixed:, 1] = c + d1 * l1
    verts[n_fixed:, 2] = c + d2 * l2
    g_small = n_geoms - 1 - n_lights
    cuts = n_fixed + (np.arange(1, g_small + 1, dtype=np.int64) * n_small) // g_small
    off += [int(x) for x in cuts]
    mats = np.zeros((n_geoms, 13), np.float32)
    mats[:, 0:3] = rng.uniform(0, 0.3, size=(n_geoms, 3))
    mats[:, 3:6] = rng.uniform(0.1, 1.0, size=(n_geoms, 3))
    if specular:
        mats[:, 6:9] = rng.uniform(0, 0.5, size=(n_geoms, 3))
    mats[:, 12] = 10.0
    mats[0] = _mat(kd=(0.7, 0.7, 0.7), ka=(0.1, 0.1, 0.1))
    for l in range(n_lights):
        mats[1 + l] = _mat(kd=(0.8, 0.8, 0.8), ke=(10, 10, 10))
    normals = None
    has_n = np.zeros(n_geoms, np.int32)
    if with_normals:
        nrm = np.cross(verts[:, 1] - verts[:, 0], verts[:, 2] - verts[:, 0])
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
        jit = rng.normal(scale=0.2, size=(n_tris, 3, 3)).astype(np.float32)
        normals = nrm[:, None, :] + jit
        normals /= np.linalg.norm(normals, axis=2, keepdims=True)
        normals = normals.astype(np.float32)
        has_n[1 + n_lights::2] = 1
    sph = np.zeros((n_spheres, 4), np.float32)
    smat = np.zeros((n_spheres, 13), np.float32)
    if n_spheres:
        sph[:, 0:3] = rng.uniform([-1, 0.05, -1], [1, 1.8, 1], size=(n_spheres, 3))
        sph[:, 3] = rng.uniform(0.01, 0.03, size=n_spheres)
        smat[:, 0:3] = rng.uniform(0, 0.3, size=(n_spheres, 3))
        smat[:, 3:6] = rng.uniform(0.1, 1.0, size=(n_spheres, 3))
        smat[:, 12] = 10.0
    return Scene(np.array(off), verts, mats, np.arange(1, 1 + n_lights), tri_normals=normals, geom_has_normals=has_n,
                 sphere_cr=sph, sphere_material=smat)


def from_npz(path):
    """Scene + (eye, look) from a fixture written by tests/golden/make_golden.py."""
    d = np.load(path)
    sc = Scene(d["geom_tri_offset"], d["tri_verts"], d["geom_material"], d["light_geom"],
               tri_normals=d["tri_normals"] if "tri_normals" in d.files else None,
               geom_has_normals=d["geom_has_normals"] if "geom_has_normals" in d.files else None)
    return sc, tuple(float(x) for x in d["eye"]), tuple(float(x) for x in d["look"])
