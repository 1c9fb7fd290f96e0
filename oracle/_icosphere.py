"""Geodesic icosphere vertices -- oracle restatement (TEST INFRASTRUCTURE).

Follows ns_reni/reni/model_components/illumination_samplers.py:100-326 (IcosahedronSampler
.icosphere / .icosahedron / .subdivide_mesh / .inside_points) for the VERTICES only; faces
are never used on the hot path (illumination_samplers.py:99-100 keeps ``vertices`` and
drops ``faces``).  Vertex order and the float64 arithmetic order are preserved so that the
float32 directions -- and therefore the sign of d_z for near-equator vertices, which the
upper-hemisphere mask at neusky/models/neusky_model.py:1653-1657 depends on -- are
bit-identical to the reference (checked in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np


def _base_icosahedron():
    # illumination_samplers.py:139-155
    phi = (1 + np.sqrt(5)) / 2
    half = np.array([[0, 1, phi], [0, -1, phi], [1, phi, 0], [-1, phi, 0], [phi, 0, 1], [-phi, 0, 1]]) / np.sqrt(1 + phi**2)
    verts = np.r_[half, -half]
    faces = np.array(
        [[0, 5, 1], [0, 3, 5], [0, 2, 3], [0, 4, 2], [0, 1, 4], [1, 5, 8], [5, 3, 10], [3, 2, 7], [2, 4, 11], [4, 1, 9],
         [7, 11, 6], [11, 9, 6], [9, 8, 6], [8, 10, 6], [10, 7, 6], [2, 11, 7], [4, 9, 11], [1, 8, 9], [5, 10, 8], [3, 7, 10]],
        dtype=int,
    )
    return verts, faces


def subdivision_frequency(nr_verts: int) -> int:
    # illumination_samplers.py:125-130: smallest nu with 12 + 10 (nu+1)(nu-1) >= nr_verts
    return int(max(1, np.ceil(np.sqrt(max(1 + (nr_verts - 12) / 10, 1)))))


def icosphere(nu: int = 1, nr_verts=None):
    """Returns (vertices [V,3] float64 on the unit sphere, None)."""
    base, faces = _base_icosahedron()
    if nr_verts is not None:
        nu = max(nu, subdivision_frequency(nr_verts))
    if nu <= 1:
        return base, None

    # unique undirected edges, lexicographically sorted (illumination_samplers.py:181-182)
    e = np.r_[faces[:, :-1], faces[:, 1:], faces[:, [0, 2]]]
    edges = np.unique(np.sort(e, axis=1), axis=0)
    V, E, F = base.shape[0], edges.shape[0], faces.shape[0]
    n_edge, n_face = nu - 1, (nu - 1) * (nu - 2) // 2
    out = np.empty((V + E * n_edge + F * n_face, 3))
    out[:V] = base

    # on-edge vertices (illumination_samplers.py:204-209)
    w = np.arange(1, nu) / nu
    lookup = {}
    for i, (a, b) in enumerate(edges):
        lookup[(a, b)] = (i, False)
        lookup[(b, a)] = (i, i != 0)  # the reference stores -i for the reversed edge; -0 == 0 is "not reversed"
        for k in range(n_edge):
            out[V + i * n_edge + k] = w[-1 - k] * base[a] + w[k] * base[b]

    def edge_rows(a, b):
        i, rev = lookup[(a, b)]
        rows = V + i * n_edge + np.arange(n_edge)
        return rows[::-1] if rev else rows

    # on-face vertices, interpolated between the A-B and A-C edge vertices (:213-228, :303-321)
    for f in range(F):
        A, B, C = faces[f]
        vab, vac = out[edge_rows(A, B)], out[edge_rows(A, C)]
        pts = []
        for i in range(1, n_edge):
            wi = np.arange(1, i + 1) / (i + 1)
            for k in range(i):
                pts.append(wi[-1 - k] * vab[i] + wi[k] * vac[i])
        start = V + E * n_edge + f * n_face
        if n_face:
            out[start : start + n_face] = np.array(pts).reshape(-1, 3)

    out = out / np.sqrt(np.sum(out**2, axis=1, keepdims=True))  # :135
    return out, None
