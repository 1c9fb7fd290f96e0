"""Drop-in boundary checks that need no GPU: the nn.Module drop-ins register the reference's ``state_dict``
names / shapes / requires_grad flags, take the reference's constructor arguments, round-trip the reference's own modules'
``state_dict`` (built live through oracle/ref_shim when /root/reference is present, else from the committed layout
tests/golden/state_dict_keys.json that tests/golden/make_golden.py wrote from those modules), expose the reference's
optimizer groups, and ``retarget`` repoints a method specification."""
import json
import os
from types import SimpleNamespace

import pytest
import torch

from neusky_b200 import fields as F
from neusky_b200 import models as M

HERE = os.path.dirname(os.path.abspath(__file__))
HAVE_REF = os.path.isdir("/root/reference/neusky")


def _layout(m):
    req = {n: bool(p.requires_grad) for n, p in m.named_parameters()}
    return {k: {"shape": list(v.shape), "requires_grad": req.get(k)} for k, v in m.state_dict().items()}


@pytest.fixture(scope="module")
def ref_layout():
    with open(os.path.join(HERE, "golden", "state_dict_keys.json")) as f:
        return json.load(f)


def test_ddf_field_layout_matches_reference(ref_layout):
    ours = F.DirectionalDistanceFieldConfig().setup(ddf_radius=1.0)
    assert _layout(ours) == ref_layout["DirectionalDistanceField"]
    assert ours.ddf_radius == 1.0


def test_reni_field_layout_matches_reference(ref_layout):
    assert _layout(F.RENIFieldConfig().setup(num_train_data=None, num_eval_data=None)) == ref_layout["RENIField"]
    got = _layout(F.RENIField(F.RENIFieldConfig(), num_train_data=7, num_eval_data=3, normalisations={"min_max": None, "log_domain": True}))
    assert got == ref_layout["RENIField_7_3"]


def test_sdf_albedo_field_layout():
    """SDFAlbedoField cannot be built from the reference here (nerfstudio's SDFField is absent); names follow
    neusky/fields/sdf_albedo_field.py:104-161 (aabb, embedding_appearance, encoding, glin*, deviation_network, clin* with
    nn.utils.weight_norm's weight_g / weight_v) and nerfstudio's LearnedVariance / Embedding."""
    f = F.SDFAlbedoFieldConfig().setup(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=5)
    sd = f.state_dict()
    want = {"aabb": (2, 3), "embedding_appearance.embedding.weight": (5, 32), "encoding.hash_table": (16 << 19, 2), "deviation_network.variance": (1,)}
    for l, (o, i) in enumerate([(256, 71), (256, 256), (257, 256)]):
        want.update({f"glin{l}.weight_g": (o, 1), f"glin{l}.weight_v": (o, i), f"glin{l}.bias": (o,)})
    for l, (o, i) in enumerate([(256, 295), (256, 256), (3, 256)]):
        want.update({f"clin{l}.weight_g": (o, 1), f"clin{l}.weight_v": (o, i), f"clin{l}.bias": (o,)})
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert not f.aabb.requires_grad
    # geometric init [NS-mem A.4]: sphere-like sdf bias, weight_g = row norm of weight_v
    assert torch.allclose(f.glin2.bias, torch.full((257,), -0.1))
    assert torch.allclose(f.glin1.weight_g, f.glin1.weight_v.norm(dim=1, keepdim=True))
    # a 0-dim variance (how neusky_b200.init writes it) loads into the [1] parameter
    sd2 = dict(sd)
    sd2["deviation_network.variance"] = torch.tensor(0.3)
    f.load_state_dict(sd2, strict=True)
    assert float(f.deviation_network.get_variance()) == pytest.approx(float(torch.exp(torch.tensor(3.0))), rel=1e-6)
    with pytest.raises(NotImplementedError):
        F.SDFAlbedoFieldConfig(num_layers=8).setup(aabb=torch.zeros(2, 3), num_images=1)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")
def test_state_dict_roundtrip_with_live_reference_modules():
    from oracle import ref_shim

    ref_shim.install()
    from neusky.fields.directional_distance_field import DirectionalDistanceField, DirectionalDistanceFieldConfig
    from reni.illumination_fields.reni_illumination_field import RENIField, RENIFieldConfig

    ref = DirectionalDistanceField(DirectionalDistanceFieldConfig(
        ddf_type="ddf", position_encoding_type="hash", direction_encoding_type="nerf", conditioning="FiLM", termination_output_activation="sigmoid",
        probability_of_hit_output_activation="sigmoid", hidden_layers=5, hidden_features=256, mapping_layers=5, mapping_features=256,
        num_attention_heads=8, num_attention_layers=6, predict_probability_of_hit=False), ddf_radius=1.0)
    ours = F.DirectionalDistanceField(F.DirectionalDistanceFieldConfig(), ddf_radius=1.0)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k

    cfg = dict(conditioning="Attention", invariant_function="VN", equivariance="SO2", axis_of_invariance="z", positional_encoding="NeRF",
               encoded_input="Directions", latent_dim=100, hidden_features=128, hidden_layers=9, mapping_layers=5, mapping_features=128,
               num_attention_heads=8, num_attention_layers=6, output_activation="None", last_layer_linear=True, fixed_decoder=True, trainable_scale=True)
    norm = {"min_max": None, "log_domain": True}
    ref_r = RENIField(RENIFieldConfig(**cfg), num_train_data=4, num_eval_data=2, normalisations=norm)
    ours_r = F.RENIField(F.RENIFieldConfig(**cfg), num_train_data=4, num_eval_data=2, normalisations=norm)
    ours_r.load_state_dict(ref_r.state_dict(), strict=True)
    ref_r.load_state_dict(ours_r.state_dict(), strict=True)
    assert bool(ours_r.log_domain) and ours_r._is_log_domain()
    # hold_decoder_fixed restores the flags it found (reni_illumination_field.py:157-196)
    ours_r.network.fc.weight.requires_grad_(True)
    with ours_r.hold_decoder_fixed():
        assert not ours_r.network.fc.weight.requires_grad and not ours_r.train_scale.requires_grad and ours_r.fixed_decoder
    assert ours_r.network.fc.weight.requires_grad and ours_r.train_scale.requires_grad


def _model(with_ddf=True):
    box = M.SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]]))
    ddf = M.DDFModelConfig().setup(ddf_radius=1.0) if with_ddf else None
    return M.NeuSkyFactoModelConfig().setup(scene_box=box, num_train_data=4, num_val_data=2, num_test_data=3, visibility_field=ddf, test_mode="val"), ddf


def test_model_parameter_names_and_param_groups():
    m, ddf = _model()
    g = m.get_param_groups()
    assert set(g) == {"fields", "proposal_networks", "illumination_field", "visibility_sigmoid"}          # neusky_model.py:379-398
    assert set(ddf.get_param_groups()) == {"ddf_field"}                                                     # ddf_model.py:151-156
    assert g["illumination_field"][0] is m.train_illumination_latents and g["illumination_field"][1] is m.train_scale
    assert g["visibility_sigmoid"] == [m.visibility_threshold] and float(m.visibility_threshold) == 2.0     # :234
    assert len(g["fields"]) == len(list(m.field.parameters())) and len(g["proposal_networks"]) == 10
    sd = m.state_dict()
    for k in ("train_illumination_latents", "train_scale", "eval_illumination_latents", "eval_scale", "eval_rotation", "visibility_threshold",
              "field.glin0.weight_v", "field.encoding.hash_table", "field.deviation_network.variance", "illumination_field.network.fc.weight",
              "illumination_field.log_domain", "visibility_field.field.ddf.final_layer.weight", "proposal_networks.1.mlp.1.bias"):
        assert k in sd, k
    assert not any(k.startswith("_train_step") for k in sd)
    assert m.eval_illumination_latents.shape == (2, 100, 3) and m.train_illumination_latents.shape == (4, 100, 3)   # test_mode "val" -> num_val_data
    assert all(not p.requires_grad for p in m.illumination_field.network.parameters())                     # fixed decoder (neusky_config.py:94)
    assert m.ddf_radius == 1.0
    # the eval / train latent switch of get_illumination_field (:400-410)
    m.train()
    assert m.get_illumination_field()[0] is m.train_illumination_latents
    m.eval()
    assert m.get_illumination_field()[0] is m.eval_illumination_latents
    # a second model loads the first one's state_dict strictly
    m2, _ = _model()
    m2.load_state_dict(sd, strict=True)


def test_reference_errors_are_kept():
    m, _ = _model(with_ddf=False)
    with pytest.raises(ValueError):
        m.load_illumination_decoder("/nonexistent/step-000050000.ckpt")                                     # neusky_model.py:283-284
    with pytest.raises(NotImplementedError):
        F.DirectionalDistanceFieldConfig(position_encoding_type="icosphere_hash").setup()                   # directional_distance_field.py:177-181
    f = F.SDFAlbedoFieldConfig().setup(aabb=torch.zeros(2, 3), num_images=1)
    rs = SimpleNamespace(camera_indices=None, frustums=None)
    with pytest.raises(AttributeError):
        f(rs)                                                                                                # sdf_albedo_field.py:218-219
    with pytest.raises(ValueError):                                                                          # no CPU path
        f.get_sdf_at_pos(torch.zeros(4, 3))


def test_retarget_points_the_hot_path_targets_here():
    from neusky_b200 import neusky_config as C

    class Ref:      # stand-ins for the reference's config objects (anything with a _target)
        pass

    mk = lambda **kw: SimpleNamespace(_target=Ref, **kw)
    model = mk(sdf_field=mk(), illumination_field=mk(), illumination_sampler=mk(), eval_num_rays_per_chunk=256)
    pipe = mk(model=model, visibility_field=mk(ddf_field=mk()), datamanager=mk())
    spec = SimpleNamespace(config=SimpleNamespace(method_name="neusky", pipeline=pipe), description="Base config for NeuSky.")
    out = C.retarget(spec)
    p = out.config.pipeline
    assert p.model._target is M.NeuSkyFactoModel and p.model.sdf_field._target is F.SDFAlbedoField and p.model.illumination_field._target is F.RENIField
    assert p.visibility_field._target is M.DDFModel and p.visibility_field.ddf_field._target is F.DirectionalDistanceField
    assert p.datamanager._target is Ref and p.model.illumination_sampler._target is Ref and p._target is Ref      # everything else untouched
    assert out.config.method_name == "neusky-b200" and spec.config.method_name == "neusky" and pipe.model._target is Ref   # deep copy
