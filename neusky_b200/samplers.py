"""Illumination direction sets (host side, numpy float64 -> float32 like the reference).

Mirrors the reference sampler interface: ``sampler()`` returns an object whose ``.frustums.directions`` is the
[D,3] direction tensor (ns_reni/reni/model_components/illumination_samplers.py:59-69, used at
neusky/models/neusky_model.py:452-479).  `IcosahedronSampler(num_directions)` builds the geodesic icosphere with the
smallest subdivision frequency that has at least `num_directions` vertices (:125-130 -- 512 -> 642 directions) in the
reference's vertex order and float64 operation order, because the upper-hemisphere mask d_z > 0
(neusky_model.py:1653-1657) is sign-sensitive for the near-equator vertices;
tests/test_samplers.py checks the vertices bit-for-bit against fixtures produced by the reference's own class.
`EquirectangularSampler(width)` is the width x width/2 lat-long grid (:373-432) that BASELINE.json config 2 uses.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

_PHI = (1.0 + np.sqrt(5.0)) / 2.0
# unit icosahedron: 6 vertices and their antipodes; 20 faces (illumination_samplers.py:139-155 fixes this labelling,
# which in turn fixes the vertex order of every subdivision level)
_HALF = np.array([[0, 1, _PHI], [0, -1, _PHI], [1, _PHI, 0], [-1, _PHI, 0], [_PHI, 0, 1], [-_PHI, 0, 1]], dtype=np.float64) / np.sqrt(1 + _PHI**2)
_FACES = np.array([[0, 5, 1], [0, 3, 5], [0, 2, 3], [0, 4, 2], [0, 1, 4], [1, 5, 8], [5, 3, 10], [3, 2, 7], [2, 4, 11], [4, 1, 9],
                   [7, 11, 6], [11, 9, 6], [9, 8, 6], [8, 10, 6], [10, 7, 6], [2, 11, 7], [4, 9, 11], [1, 8, 9], [5, 10, 8], [3, 7, 10]])


def subdivision_frequency(num_directions: int) -> int:
    """Smallest nu with 12 + 10 (nu+1)(nu-1) >= num_directions."""
    return int(max(1, np.ceil(np.sqrt(max(1 + (num_directions - 12) / 10, 1)))))


def icosphere_vertices(num_directions: int) -> np.ndarray:
    """[V,3] float64 unit vectors: 12 corners, then (nu-1) points per edge (edges in sorted order), then the interior
    points of each face, row by row."""
    nu = subdivision_frequency(num_directions)
    corners = np.concatenate([_HALF, -_HALF], 0)
    if nu == 1:
        return corners
    pairs = np.sort(np.concatenate([_FACES[:, [0, 1]], _FACES[:, [1, 2]], _FACES[:, [0, 2]]], 0), axis=1)
    edges = np.unique(pairs, axis=0)                                   # [30,2], lexicographic
    m = nu - 1
    t = np.arange(1, nu, dtype=np.float64) / nu                        # 1/nu .. (nu-1)/nu
    # point k of edge (a,b): t[m-1-k] * a + t[k] * b   (this exact expression: the result is rounded as in the reference)
    on_edge = t[::-1][None, :, None] * corners[edges[:, 0]][:, None, :] + t[None, :, None] * corners[edges[:, 1]][:, None, :]   # [30,m,3]
    edge_id = {(int(a), int(b)): i for i, (a, b) in enumerate(edges)}

    def along(a: int, b: int) -> np.ndarray:
        """the m points of edge {a,b} walking from a to b.  The reference marks a reversed edge by a negated index, so
        edge 0 can never be marked reversed (-0 == 0) and is always walked in stored order."""
        if (a, b) in edge_id:
            return on_edge[edge_id[(a, b)]]
        i = edge_id[(b, a)]
        return on_edge[i] if i == 0 else on_edge[i][::-1]

    inner = []
    for A, B, C in _FACES:
        ab, ac = along(int(A), int(B)), along(int(A), int(C))
        for i in range(1, m):                                           # row i has i interior points
            u = np.arange(1, i + 1, dtype=np.float64) / (i + 1)
            inner.append(u[::-1][:, None] * ab[i][None, :] + u[:, None] * ac[i][None, :])
    parts = [corners, on_edge.reshape(-1, 3)] + inner
    v = np.concatenate(parts, 0)
    return v / np.sqrt(np.sum(v**2, axis=1, keepdims=True))


class _Samples(SimpleNamespace):
    """Stand-in for the RaySamples the reference samplers return: only `.frustums.directions` and an assignable
    `.camera_indices` are used downstream (neusky_model.py:452-479)."""


class IcosahedronSampler:
    def __init__(self, num_directions: int = 512, apply_random_rotation: bool = False, remove_lower_hemisphere: bool = False, seed: Optional[int] = None):
        self.apply_random_rotation = apply_random_rotation
        self.remove_lower_hemisphere = remove_lower_hemisphere
        self.directions = torch.from_numpy(icosphere_vertices(num_directions)).float()     # [D,3], z up
        self._rng = np.random.default_rng(seed)

    def generate_direction_samples(self, apply_random_rotation: Optional[bool] = None) -> _Samples:
        """illumination_samplers.py:331-355: optional random SO(3) rotation of the whole set (training), optional
        removal of the lower hemisphere."""
        d = self.directions
        rot = self.apply_random_rotation if apply_random_rotation is None else apply_random_rotation
        if rot:
            from scipy.spatial.transform import Rotation

            R = torch.from_numpy(Rotation.random(1, random_state=self._rng).as_matrix()[0]).float()
            d = d @ R
        if self.remove_lower_hemisphere:
            d = d[d[:, 2] > 0]
        return _Samples(frustums=SimpleNamespace(directions=d), camera_indices=None)

    __call__ = forward = generate_direction_samples


class EquirectangularSampler:
    """width x width/2 lat-long grid through nerfstudio's equirectangular ray generation [SURVEY A.8]:
    fx = fy = H, cx = W/2, cy = H/2, pixel centres at +0.5, y/z swapped to z-up (illumination_samplers.py:386-395)."""

    def __init__(self, width: int = 64, remove_lower_hemisphere: bool = False):
        H, W = width // 2, width
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")
        u = (xs - float(W // 2)) / float(H)
        v = -(ys - float(H // 2)) / float(H)
        theta, phi = -torch.pi * u, torch.pi * (0.5 - v)
        d_cam = torch.stack([-torch.sin(theta) * torch.sin(phi), torch.cos(phi), -torch.cos(theta) * torch.sin(phi)], -1).reshape(-1, 3)
        d = d_cam @ torch.tensor([[1.0, 0, 0], [0, 0, 1.0], [0, 1.0, 0]]).T
        self.directions = d / d.norm(dim=-1, keepdim=True)
        self.remove_lower_hemisphere = remove_lower_hemisphere

    def generate_direction_samples(self, apply_random_rotation: Optional[bool] = None) -> _Samples:
        d = self.directions
        if self.remove_lower_hemisphere:
            d = d[d[:, 2] > 0]
        return _Samples(frustums=SimpleNamespace(directions=d), camera_indices=None)

    __call__ = forward = generate_direction_samples
