"""Minimal ray containers with nerfstudio's attribute names (nerfstudio.cameras.rays.{Frustums, RaySamples, RayBundle},
SURVEY Appendix A.1), so the drop-in modules can be driven without nerfstudio installed.  The modules themselves are
duck-typed: nerfstudio's own objects work unchanged (only ``.frustums.origins/.directions/.starts/.ends``, ``.deltas``,
``.camera_indices``, ``.metadata["directions_norm"]``, ``.nears/.fars`` are read)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import torch

Tensor = torch.Tensor


@dataclass
class Frustums:
    origins: Tensor = None
    directions: Tensor = None
    starts: Tensor = None
    ends: Tensor = None
    pixel_area: Optional[Tensor] = None
    offsets: Optional[Tensor] = None

    def get_positions(self) -> Tensor:
        return self.origins + self.directions * (self.starts + self.ends) / 2

    def get_start_positions(self) -> Tensor:
        return self.origins + self.directions * self.starts

    @property
    def shape(self):
        return self.directions.shape[:-1]


@dataclass
class RaySamples:
    frustums: Frustums = None
    camera_indices: Optional[Tensor] = None
    deltas: Optional[Tensor] = None
    spacing_starts: Optional[Tensor] = None
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Any = None
    metadata: Optional[Dict[str, Tensor]] = None
    times: Optional[Tensor] = None

    @property
    def shape(self):
        return self.frustums.shape

    def to(self, device):
        mv = lambda t: t.to(device) if isinstance(t, torch.Tensor) else t
        fr = Frustums(**{k: mv(getattr(self.frustums, k)) for k in ("origins", "directions", "starts", "ends", "pixel_area", "offsets")})
        return RaySamples(frustums=fr, camera_indices=mv(self.camera_indices), deltas=mv(self.deltas), spacing_starts=mv(self.spacing_starts),
                          spacing_ends=mv(self.spacing_ends), spacing_to_euclidean_fn=self.spacing_to_euclidean_fn, metadata=self.metadata, times=mv(self.times))


@dataclass
class RayBundle:
    origins: Tensor = None
    directions: Tensor = None
    pixel_area: Optional[Tensor] = None
    camera_indices: Optional[Tensor] = None
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None

    def __len__(self) -> int:
        return int(self.origins.shape[0])

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        sl = lambda t: t.reshape(-1, *t.shape[len(self.shape):])[start_idx:end_idx] if isinstance(t, torch.Tensor) else t
        return RayBundle(origins=sl(self.origins), directions=sl(self.directions), pixel_area=sl(self.pixel_area), camera_indices=sl(self.camera_indices),
                         nears=sl(self.nears), fars=sl(self.fars), metadata={k: sl(v) for k, v in (self.metadata or {}).items()}, times=sl(self.times))
