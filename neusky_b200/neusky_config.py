"""nerfstudio method registration for the drop-in path (reference: pyproject.toml:17-20,
neusky/configs/neusky_config.py:33-242).

nerfstudio selects components by the ``_target`` of each dataclass config and builds them with ``config.setup(**kwargs)``.
``retarget(method)`` takes the REFERENCE's own ``MethodSpecification`` (so data parsers, data manager, pipeline, optimizers and
schedulers stay the reference's, unchanged) and points the four hot-path targets at this package:

    model._target                              -> neusky_b200.models.NeuSkyFactoModel
    model.sdf_field._target                    -> neusky_b200.fields.SDFAlbedoField
    model.illumination_field._target           -> neusky_b200.fields.RENIField
    visibility_field._target / .ddf_field._target -> neusky_b200.models.DDFModel / neusky_b200.fields.DirectionalDistanceField

``NeuSkyB200`` is that retargeted specification, registered as the ``neusky-b200`` method by this repo's pyproject.toml
(``ns-train neusky-b200 ...``).  It is only defined when nerfstudio AND the reference package are importable (neither is
in the build image: the guarded import keeps this module importable there, and tests/test_dropin_state_dict.py checks the
retargeting logic on stand-in config objects).
"""
from __future__ import annotations

import copy
from typing import Any

from . import fields as _fields
from . import models as _models


def retarget(method: Any) -> Any:
    """A deep copy of a reference MethodSpecification whose hot-path ``_target``s are this package's drop-in classes.  Every
    other field (and every other config object) is left as the reference set it."""
    m = copy.deepcopy(method)
    pipe = m.config.pipeline
    model = pipe.model
    model._target = _models.NeuSkyFactoModel
    model.sdf_field._target = _fields.SDFAlbedoField
    model.illumination_field._target = _fields.RENIField
    vf = getattr(pipe, "visibility_field", None)
    if vf is not None:
        vf._target = _models.DDFModel
        vf.ddf_field._target = _fields.DirectionalDistanceField
    m.config.method_name = str(getattr(m.config, "method_name", "neusky")) + "-b200"
    if hasattr(m, "description"):
        m.description = "NeuSky with the B200-native render-and-shade path (neusky_b200)."
    return m


try:  # pragma: no cover - neither package exists in the build image
    from neusky.configs.neusky_config import NeuSky as _ReferenceNeuSky  # noqa: F401  (imports nerfstudio)

    NeuSkyB200 = retarget(_ReferenceNeuSky)
except Exception:  # ImportError and whatever a half-installed nerfstudio raises
    NeuSkyB200 = None
