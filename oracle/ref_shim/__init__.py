"""Import shim so the reference's own modules (/root/reference/neusky, /root/reference/ns_reni/reni)
can be imported and run on CPU in THIS container, where nerfstudio / tinycudann / nerfacc /
torchmetrics / matplotlib are absent.  TEST INFRASTRUCTURE -- used only by
``tests/golden/make_golden.py`` to generate fixtures; nothing here ships or travels.

``install()`` registers:
  * real (restated-from-memory, [NS-mem] SURVEY Appendix A) implementations of the few
    nerfstudio symbols whose arithmetic is on the hot path: ``RaySamples/Frustums/RayBundle``,
    ``NeRFEncoding``, ``FieldHeadNames``, ``Field``, ``InstantiateConfig``; and
    ``tinycudann.Encoding`` mapped to nerfstudio's torch hash grid (A.3);
  * permissive stubs for every other ``nerfstudio.*``, ``nerfacc``, ``torchmetrics.*``,
    ``matplotlib.*`` ... attribute, so that ``from x import Y`` succeeds and ``class Z(Y)``
    is legal.  Stubs raise if they are ever *called* for arithmetic.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Optional, Type

import numpy as np
import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"
_STUB_ROOTS = ("nerfstudio", "nerfacc", "tinycudann", "torchmetrics", "matplotlib", "icosphere", "roma", "pyexr", "imageio", "lpips", "yacs", "wandb", "mediapy", "plotly", "viser", "tyro", "cv2", "PIL", "jaxtyping", "rich", "skimage", "OpenEXR", "Imath")


class _StubMeta(type):
    def __getattr__(cls, name):  # class attributes such as colors.WHITE
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(f"{cls.__name__}.{name}")

    def __getitem__(cls, item):  # typing-style subscripts: TensorType["bs", 3]
        return cls


def _make_stub(name: str):
    return _StubMeta(name.split(".")[-1], (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None, "_is_stub": True})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _make_stub(f"{self.__name__}.{name}")
        setattr(self, name, obj)
        return obj


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS and fullname not in sys.modules:
            try:  # prefer a real installed module when there is one (cv2, rich, tyro, ...)
                for finder in sys.meta_path:
                    if finder is self:
                        continue
                    spec = finder.find_spec(fullname, path, target) if hasattr(finder, "find_spec") else None
                    if spec is not None:
                        return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _populate(module)


# ------------------------------------------------------------------ real pieces [NS-mem]


class FieldHeadNames(Enum):
    RGB = "rgb"
    SH = "sh"
    DENSITY = "density"
    NORMALS = "normals"
    PRED_NORMALS = "pred_normals"
    UNCERTAINTY = "uncertainty"
    BACKGROUND_RGB = "background_rgb"
    TRANSIENT_RGB = "transient_rgb"
    TRANSIENT_DENSITY = "transient_density"
    SEMANTICS = "semantics"
    SDF = "sdf"
    ALPHA = "alpha"
    GRADIENT = "gradient"


@dataclass
class Frustums:
    origins: Any = None
    directions: Any = None
    starts: Any = None
    ends: Any = None
    pixel_area: Any = None
    offsets: Any = None

    def get_positions(self):
        return self.origins + self.directions * (self.starts + self.ends) / 2

    def get_start_positions(self):
        return self.origins + self.directions * self.starts

    @property
    def shape(self):
        return self.directions.shape[:-1]


@dataclass
class RaySamples:
    frustums: Frustums = None
    camera_indices: Any = None
    deltas: Any = None
    spacing_starts: Any = None
    spacing_ends: Any = None
    spacing_to_euclidean_fn: Any = None
    metadata: Any = None
    times: Any = None

    @property
    def shape(self):
        return self.frustums.shape

    def to(self, device):
        return self


@dataclass
class RayBundle:
    origins: Any = None
    directions: Any = None
    pixel_area: Any = None
    camera_indices: Any = None
    nears: Any = None
    fars: Any = None
    metadata: Any = None
    times: Any = None

    def __len__(self):
        return self.origins.shape[0]


class NeRFEncoding(nn.Module):
    """nerfstudio.field_components.encodings.NeRFEncoding, torch backend [NS-mem A.2]."""

    def __init__(self, in_dim, num_frequencies, min_freq_exp, max_freq_exp, include_input=False, implementation="torch"):
        super().__init__()
        self.in_dim, self.num_frequencies = in_dim, num_frequencies
        self.min_freq, self.max_freq, self.include_input = min_freq_exp, max_freq_exp, include_input

    def get_out_dim(self):
        return self.in_dim * self.num_frequencies * 2 + (self.in_dim if self.include_input else 0)

    def forward(self, in_tensor, covs=None):
        scaled = 2 * torch.pi * in_tensor
        freqs = 2 ** torch.linspace(self.min_freq, self.max_freq, self.num_frequencies, device=in_tensor.device)
        s = scaled[..., None] * freqs
        s = s.view(*s.shape[:-2], -1)
        enc = torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))
        if self.include_input:
            enc = torch.cat([enc, in_tensor], dim=-1)
        return enc


class TorchHashEncoding(nn.Module):
    """Stand-in for ``tinycudann.Encoding`` with nerfstudio's HashEncoding.pytorch_fwd
    semantics [NS-mem A.3]."""

    def __init__(self, n_input_dims=3, encoding_config=None, **_):
        super().__init__()
        c = encoding_config
        self.num_levels = c["n_levels"]
        self.features_per_level = c["n_features_per_level"]
        self.log2_hashmap_size = c["log2_hashmap_size"]
        self.hash_table_size = 2 ** c["log2_hashmap_size"]
        min_res = c["base_resolution"]
        growth = c["per_level_scale"]
        levels = torch.arange(self.num_levels)
        self.scalings = torch.floor(min_res * growth**levels)
        self.hash_offset = levels * self.hash_table_size
        self.hash_table = nn.Parameter((torch.rand(self.hash_table_size * self.num_levels, self.features_per_level) * 2 - 1) * 1e-3)
        self.n_output_dims = self.num_levels * self.features_per_level

    def hash_fn(self, t):
        t = t * torch.tensor([1, 2654435761, 805459861])
        x = torch.bitwise_xor(t[..., 0], t[..., 1])
        x = torch.bitwise_xor(x, t[..., 2])
        x %= self.hash_table_size
        x += self.hash_offset
        return x

    def forward(self, in_tensor):
        in_tensor = in_tensor[..., None, :]
        scaled = in_tensor * self.scalings.view(-1, 1)
        sc = torch.ceil(scaled).type(torch.int32)
        sf = torch.floor(scaled).type(torch.int32)
        offset = scaled - sf
        cat = torch.cat
        h0 = self.hash_fn(sc)
        h1 = self.hash_fn(cat([sc[..., 0:1], sf[..., 1:2], sc[..., 2:3]], dim=-1))
        h2 = self.hash_fn(cat([sf[..., 0:1], sf[..., 1:2], sc[..., 2:3]], dim=-1))
        h3 = self.hash_fn(cat([sf[..., 0:1], sc[..., 1:2], sc[..., 2:3]], dim=-1))
        h4 = self.hash_fn(cat([sc[..., 0:1], sc[..., 1:2], sf[..., 2:3]], dim=-1))
        h5 = self.hash_fn(cat([sc[..., 0:1], sf[..., 1:2], sf[..., 2:3]], dim=-1))
        h6 = self.hash_fn(sf)
        h7 = self.hash_fn(cat([sf[..., 0:1], sc[..., 1:2], sf[..., 2:3]], dim=-1))
        f0, f1, f2, f3 = self.hash_table[h0], self.hash_table[h1], self.hash_table[h2], self.hash_table[h3]
        f4, f5, f6, f7 = self.hash_table[h4], self.hash_table[h5], self.hash_table[h6], self.hash_table[h7]
        f03 = f0 * offset[..., 0:1] + f3 * (1 - offset[..., 0:1])
        f12 = f1 * offset[..., 0:1] + f2 * (1 - offset[..., 0:1])
        f56 = f5 * offset[..., 0:1] + f6 * (1 - offset[..., 0:1])
        f47 = f4 * offset[..., 0:1] + f7 * (1 - offset[..., 0:1])
        f0312 = f03 * offset[..., 1:2] + f12 * (1 - offset[..., 1:2])
        f4756 = f47 * offset[..., 1:2] + f56 * (1 - offset[..., 1:2])
        enc = f0312 * offset[..., 2:3] + f4756 * (1 - offset[..., 2:3])
        return torch.flatten(enc, start_dim=-2, end_dim=-1)


@dataclass
class InstantiateConfig:
    _target: Type = None

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


@dataclass
class FieldConfig(InstantiateConfig):
    pass


class Field(nn.Module):
    def forward(self, ray_samples, compute_normals=False):
        return self.get_outputs(ray_samples)


def _scale_dict(dictionary, coefficients):
    """nerfstudio.utils.misc.scale_dict [NS-mem]: multiply the entries that have a coefficient, in place."""
    for key in dictionary:
        if key in coefficients:
            dictionary[key] *= coefficients[key]
    return dictionary


def _populate(module: types.ModuleType) -> None:
    real = {
        "nerfstudio.utils.misc": dict(scale_dict=_scale_dict),
        "nerfstudio.utils": dict(misc=types.SimpleNamespace(scale_dict=_scale_dict)),   # `from nerfstudio.utils import misc`
        "nerfstudio.cameras.rays": dict(Frustums=Frustums, RaySamples=RaySamples, RayBundle=RayBundle),
        "nerfstudio.field_components.encodings": dict(NeRFEncoding=NeRFEncoding),
        "nerfstudio.field_components.field_heads": dict(FieldHeadNames=FieldHeadNames),
        "nerfstudio.fields.base_field": dict(Field=Field, FieldConfig=FieldConfig),
        "nerfstudio.configs.base_config": dict(InstantiateConfig=InstantiateConfig),
        "tinycudann": dict(Encoding=TorchHashEncoding),
    }
    for k, v in real.get(module.__name__, {}).items():
        setattr(module, k, v)


_installed = False


def install() -> None:
    """Idempotently install the stub finder and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    sys.meta_path.insert(0, _StubFinder())
    for p in (REFERENCE_ROOT, REFERENCE_ROOT + "/ns_reni"):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True
