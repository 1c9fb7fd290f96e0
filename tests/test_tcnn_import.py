"""tiny-cuda-nn hash-grid import (neusky_b200/tcnn_import.py).

tiny-cuda-nn is not in this image and the reference carries no fixtures for it, so tcnn's grid semantics are restated from memory and
these tests are SELF-CONSISTENCY tests (parity unpinned against tcnn itself):
  * CPU: level geometry, parameter re-layout, the plain-torch statement of the grid against an independent dense trilinear
    interpolation (torch grid_sample), node values, uint32 index arithmetic, the state-dict hook;
  * GPU: the stand-alone encode kernel, K2 (fp32 and tensor-core, incl. the smoothstep factor of the analytic normal) and K4 (fp32 and
    tensor-core) on an imported grid against the oracle with its hash encode swapped for the plain-torch statement."""
import numpy as np
import pytest
import torch

from neusky_b200 import init as nb_init
from neusky_b200 import tcnn_import as TI


# ------------------------------------------------------------------------------------------------------------------ CPU
def test_level_geometry_of_the_reference_configuration():
    """16 levels, base 16, max 2048, T = 2^19 (sdf_albedo_field.py:108-130): scale_0 = 15, resolution 16; levels are dense while
    res^3 fits 2^19, sizes are multiples of 8, offsets are cumulative, the fine levels own exactly T hashed entries."""
    lv = TI.tcnn_levels(16, 16, 2048, 19)
    assert len(lv) == 16
    assert lv[0].scale == 15.0 and lv[0].resolution == 16 and lv[0].dense and lv[0].size == 4096 and lv[0].offset == 0
    off = 0
    for l in lv:
        assert l.offset == off and l.size % 8 == 0 and l.size <= 1 << 19
        assert l.dense == (l.resolution ** 3 <= l.size)
        assert l.resolution == int(np.ceil(l.scale)) + 1
        off += l.size
    assert not lv[-1].dense and lv[-1].size == 1 << 19
    assert abs(lv[-1].scale - 2047.0) < 0.5
    assert TI.tcnn_num_params(lv) == off * 2
    # dense exactly up to resolution 80 (80^3 = 512000 <= 2^19 < 81^3)
    assert all(l.dense == (l.resolution <= 80) for l in lv)


def test_params_relayout_round_trip():
    lv = TI.tcnn_levels(6, 4, 64, 10)
    n = TI.tcnn_num_params(lv)
    flat = torch.arange(n, dtype=torch.float32).to(torch.float16)            # tcnn stores fp16
    tab = TI.tcnn_params_to_table(flat, lv, 10)
    T = 1 << 10
    assert tab.shape == (6 * T, 2) and tab.dtype == torch.float32
    for i, l in enumerate(lv):
        blk = tab[i * T:(i + 1) * T]
        assert torch.equal(blk[:l.size].reshape(-1), flat[l.offset * 2:(l.offset + l.size) * 2].float())
        if l.size < T:
            assert float(blk[l.size:].abs().max()) == 0.0
    with pytest.raises(ValueError):
        TI.tcnn_params_to_table(flat[:-2], lv, 10)


def _random_grid(levels, log2_T, seed):
    g = torch.Generator().manual_seed(seed)
    n = TI.tcnn_num_params(levels)
    return TI.tcnn_params_to_table(torch.randn(n, generator=g), levels, log2_T)


def test_dense_levels_equal_independent_trilinear_interpolation():
    """Linear weights on a dense level are plain trilinear interpolation of the res^3 volume at grid position x * scale + 0.5:
    checked against torch.nn.functional.grid_sample (align_corners=True), an implementation that shares no code with ours."""
    log2_T = 12
    lv = TI.tcnn_levels(4, 4, 16, log2_T)
    tab = _random_grid(lv, log2_T, 0)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(500, 3, generator=g) * 0.9 + 0.02
    got = TI.tcnn_grid_encode_torch(x, tab, lv, log2_T, smoothstep=False)
    T = 1 << log2_T
    checked = 0
    for i, l in enumerate(lv):
        if not l.dense:
            continue
        r = l.resolution
        vol = tab[i * T:i * T + r ** 3].reshape(r, r, r, 2).permute(3, 0, 1, 2)[None]       # [1, F, z, y, x]
        pos = x * np.float32(l.scale) + 0.5                                                  # continuous grid coordinates (x, y, z)
        if float(pos.max()) > r - 1:
            continue
        grid = (2.0 * pos / (r - 1) - 1.0).reshape(1, -1, 1, 1, 3)                           # grid_sample order = (x, y, z)
        ref = torch.nn.functional.grid_sample(vol, grid, mode="bilinear", align_corners=True).reshape(2, -1).T
        assert torch.allclose(got[:, 2 * i:2 * i + 2], ref, rtol=1e-4, atol=1e-5)
        checked += 1
    assert checked >= 2


def test_node_values_and_smoothstep_midpoints():
    """At grid nodes both interpolation modes return the stored value; halfway between two nodes smoothstep(0.5) = 0.5 = linear."""
    log2_T = 12
    lv = TI.tcnn_levels(3, 4, 8, log2_T)
    tab = _random_grid(lv, log2_T, 2)
    T = 1 << log2_T
    l0 = lv[0]
    node = torch.tensor([[1, 2, 1]], dtype=torch.float32)
    x = (node - 0.5) / np.float32(l0.scale)                           # pos = x * scale + 0.5 = node
    idx = int(node[0, 0] + node[0, 1] * l0.resolution + node[0, 2] * l0.resolution ** 2)
    for sm in (False, True):
        out = TI.tcnn_grid_encode_torch(x, tab, lv, log2_T, smoothstep=sm)
        assert torch.allclose(out[0, :2], tab[idx], atol=1e-5)
    xm = (node + torch.tensor([[0.5, 0.0, 0.0]]) - 0.5) / np.float32(l0.scale)
    a = TI.tcnn_grid_encode_torch(xm, tab, lv, log2_T, smoothstep=False)[0, :2]
    b = TI.tcnn_grid_encode_torch(xm, tab, lv, log2_T, smoothstep=True)[0, :2]
    assert torch.allclose(a, b, atol=1e-5) and torch.allclose(a, 0.5 * (tab[idx] + tab[idx + 1]), atol=1e-5)


def test_hashed_level_index_is_uint32_arithmetic():
    """A single-corner probe of a hashed level: the value read is the table entry at (x ^ y * P1 ^ z * P2) mod 2^32 mod size, in numpy
    uint32 arithmetic -- also for NEGATIVE inputs (the DDF feeds points of the sphere |q| = r, directional_distance_field.py:268)."""
    log2_T = 8
    lv = TI.tcnn_levels(2, 16, 64, log2_T)          # both levels hashed (16^3 > 256)
    assert not lv[0].dense
    T = 1 << log2_T
    tab = torch.zeros(2 * T, 2)
    tab[:T, 0] = torch.arange(T, dtype=torch.float32)
    for node in ([3, 5, 7], [-2, 4, -9]):
        x = (torch.tensor([node], dtype=torch.float32) - 0.5) / np.float32(lv[0].scale)
        out = TI.tcnn_grid_encode_torch(x, tab, lv, log2_T, smoothstep=False)
        u = np.array(node, dtype=np.int64).astype(np.uint32)
        with np.errstate(over="ignore"):
            h = int((u[0] ^ (u[1] * np.uint32(2654435761)) ^ (u[2] * np.uint32(805459861))) % np.uint32(lv[0].size))
        assert abs(float(out[0, 0]) - h) < 1e-3, (node, float(out[0, 0]), h)


def test_state_dict_hook_converts_params_entry():
    class _G:      # the attributes convert_tcnn_state_dict_entry reads from fields._HashGrid
        num_levels, base_res, max_res, log2_T, features = 4, 4, 32, 10, 2
        tcnn_levels = None

    lv = TI.tcnn_levels(4, 4, 32, 10)
    flat = torch.randn(TI.tcnn_num_params(lv)).half()
    sd = {"enc.params": flat, "other": torch.zeros(1)}
    grid = _G()
    TI.convert_tcnn_state_dict_entry(sd, "enc.", grid)
    assert "enc.params" not in sd and sd["enc.hash_table"].shape == (4 << 10, 2) and grid.tcnn_levels == lv
    sd2 = {"other": torch.zeros(1)}
    grid2 = _G()
    TI.convert_tcnn_state_dict_entry(sd2, "enc.", grid2)                  # nothing to convert: untouched
    assert list(sd2) == ["other"] and grid2.tcnn_levels is None


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neusky_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


class _tcnn_oracle:
    """Context manager: the oracle's hash encode replaced by the plain-torch statement of tcnn's grid (differentiable, so the oracle's
    autograd normal carries the smoothstep derivative)."""

    def __init__(self, levels, log2_T, smoothstep):
        self.levels, self.log2_T, self.smoothstep = levels, log2_T, smoothstep

    def __enter__(self):
        from oracle import neusky_oracle as O

        self._old = O.hash_encode
        O.hash_encode = lambda x, table, scalings, log2_T: TI.tcnn_grid_encode_torch(x, table, self.levels, self.log2_T, self.smoothstep)
        return O

    def __exit__(self, *a):
        from oracle import neusky_oracle as O

        O.hash_encode = self._old


@pytest.mark.gpu
@pytest.mark.parametrize("smoothstep", [True, False])
def test_encode_kernel_vs_torch_statement(dev, smoothstep):
    from neusky_b200 import ops

    log2_T = 15
    lv = TI.tcnn_levels(16, 16, 2048, log2_T)
    assert any(l.dense for l in lv) and any(not l.dense for l in lv)
    tab = _random_grid(lv, log2_T, 3)
    g = torch.Generator().manual_seed(4)
    x = torch.cat([torch.rand(3000, 3, generator=g), torch.rand(1000, 3, generator=g) * 2 - 1])      # [0,1] and negative coordinates
    ref = TI.tcnn_grid_encode_torch(x, tab, lv, log2_T, smoothstep)
    out = ops.hash_encode_tcnn(x.to(dev), tab.to(dev), TI.tcnn_level_meta(lv, dev), log2_T, smoothstep).cpu()
    assert float((out - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


def _trained_like(p, seed, tab):
    g = torch.Generator().manual_seed(seed)
    p = {k: v.clone() for k, v in p.items()}
    p["glin0.weight_v"][:, 3:] = 0.05 * torch.randn(256, 68, generator=g)
    p["glin0.weight_g"] = p["glin0.weight_v"].norm(dim=1, keepdim=True) * (0.8 + 0.4 * torch.rand(256, 1, generator=g))
    p["encoding.hash_table"] = tab * 0.2
    p["glin1.bias"] = 0.02 * torch.randn(256, generator=g)
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("smoothstep", [True, False])
def test_k2_on_imported_grid_vs_oracle(dev, smoothstep):
    """K2 (exact fp32 and tcgen05) on a tcnn grid: sdf, ANALYTIC normal (incl. d smoothstep / dt) and albedo against the oracle whose
    encode is the torch statement of tcnn's grid (normal through autograd).  Tolerances of tests/test_gpu_sdf.py."""
    from neusky_b200 import ops, packing

    log2_T = 15
    lv = TI.tcnn_levels(16, 16, 2048, log2_T)
    tab = _random_grid(lv, log2_T, 5)
    p = _trained_like(nb_init.init_sdf_params(3, log2_T=log2_T), 6, tab)
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(600, 3, generator=g) * 2 - 1) * 1.5            # inside and outside the unit cube (contraction branch)
    with _tcnn_oracle(lv, log2_T, smoothstep) as O:
        ref = O.sdf_field(x, p, O.hash_scalings(), log2_T)
        sc = O.hash_scalings().to(dev)
    pd = {k: v.to(dev) for k, v in p.items()}
    meta = TI.tcnn_level_meta(lv, dev)
    ex = ops.sdf_field(x.to(dev), packing.pack_sdf_simt(pd), pd["encoding.hash_table"], sc, log2_T, impl="simt", grid_meta=meta, smoothstep=smoothstep)
    assert torch.allclose(ex["sdf"].cpu(), ref["sdf"], rtol=1e-4, atol=1e-5), float((ex["sdf"].cpu() - ref["sdf"]).abs().max())
    gerr = (ex["gradient"].cpu() - ref["gradient"]).abs().max() / ref["gradient"].abs().max()
    assert float(gerr) <= 1e-3, float(gerr)
    assert torch.allclose(ex["albedo"].cpu(), ref["albedo"], rtol=1e-4, atol=1e-5)
    # the imported grid really is a different function of x than the nerfstudio grid on the same table
    ns = ops.sdf_field(x.to(dev), packing.pack_sdf_simt(pd), pd["encoding.hash_table"], sc, log2_T, impl="simt")
    assert float((ns["sdf"] - ex["sdf"]).abs().max()) > 1e-3
    tc = ops.sdf_field(x.to(dev), packing.pack_sdf_tc(pd), pd["encoding.hash_table"], sc, log2_T, impl="tc", grid_meta=meta, smoothstep=smoothstep)
    assert float((tc["sdf"] - ex["sdf"]).abs().max()) <= 2e-3
    assert float((tc["albedo"] - ex["albedo"]).abs().max()) <= 5e-3
    assert float(torch.nn.functional.cosine_similarity(tc["gradient"], ex["gradient"], dim=-1).min()) >= 0.999


@pytest.mark.gpu
def test_k4_on_imported_grid_vs_oracle(dev):
    """K4 (fp32 and CTA-pair tcgen05) with the DDF's position grid imported from tcnn: per-pair visibility / expected termination distance
    against oracle.compute_visibility with the swapped encode (the DDF feeds sphere points with negative coordinates)."""
    from neusky_b200.render import SkyShader

    log2_T = 14
    lv = TI.tcnn_levels(16, 16, 2048, log2_T)
    p = nb_init.init_ddf_params(11, final_gain=8.0, log2_T=log2_T, table_scale=0.1)
    p["position_encoding.hash_table"] = _random_grid(lv, log2_T, 8) * 0.1
    g = torch.Generator().manual_seed(9)
    R = 40
    pts = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * torch.rand(R, 1, generator=g) ** (1 / 3) * 0.9
    with _tcnn_oracle(lv, log2_T, True) as O:
        dirs = O.icosphere_directions(100)
        dirs = dirs[dirs[:, 2] > 0].contiguous()
        ref = O.compute_visibility(pts, dirs, p, O.hash_scalings(), log2_T, 1.0, 0.35, 25.0, only_upper=True)
    sh = SkyShader(p, None, device=dev, log2_T=log2_T)
    sh.grid_meta, sh.grid_smoothstep = TI.tcnn_level_meta(lv, dev), True
    sh.set_directions(dirs)
    D = dirs.shape[0]
    dummy = torch.zeros(R, 1, 3, device=dev)
    rad = torch.ones(1, D, 3, device=dev)
    for impl, tol_ddf, tol_vis in (("simt", 2e-4, 5e-4), ("tc2", 2.5e-3, 1e-2)):
        out = sh.shade(pts.to(dev), dummy, dummy, rad, want_vis=True, want_ddf=True, threshold=0.35, sigmoid_scale=25.0, impl=impl)
        e_ddf = float((out["expected_termination_dist"].cpu() - ref["expected_termination_dist"]).abs().max())
        e_vis = float((out["visibility"].cpu() - ref["visibility"]).abs().max())
        assert e_ddf <= tol_ddf and e_vis <= tol_vis, (impl, e_ddf, e_vis)
    sh.grid_meta = None
    ns = sh.shade(pts.to(dev), dummy, dummy, rad, want_vis=True, want_ddf=True, threshold=0.35, sigmoid_scale=25.0, impl="simt")
    assert float((ns["expected_termination_dist"].cpu() - ref["expected_termination_dist"]).abs().max()) > 1e-3      # a different function


@pytest.mark.gpu
def test_field_module_loads_a_tcnn_checkpoint_entry(dev):
    """SDFAlbedoField.load_state_dict with the reference's ``encoding.params`` (flat fp16 tcnn vector) in place of our ``encoding.hash_table``:
    the module switches to the imported grid, eval goes through the kernels with it, training raises."""
    from neusky_b200 import fields as F

    cfg = F.SDFAlbedoFieldConfig(log2_hashmap_size=12)
    f = F.SDFAlbedoField(cfg, aabb=torch.tensor([[-1.0] * 3, [1.0] * 3]), num_images=4).to(dev)
    lv = TI.tcnn_levels(f.encoding.num_levels, f.encoding.base_res, f.encoding.max_res, 12)
    sd = {k: v.clone() for k, v in f.state_dict().items()}
    flat = (torch.randn(TI.tcnn_num_params(lv)) * 0.05).half()
    del sd["encoding.hash_table"]
    sd["encoding.params"] = flat
    f.load_state_dict(sd, strict=True)
    assert f.encoding.tcnn_levels == lv
    assert torch.equal(f.encoding.hash_table.detach().cpu(), TI.tcnn_params_to_table(flat, lv, 12))
    f.eval()
    for p_ in f.parameters():
        p_.requires_grad_(False)
    x = (torch.rand(64, 3) * 2 - 1).to(dev)
    enc = f.encoding(x * 0.25 + 0.5).cpu()
    assert torch.allclose(enc, TI.tcnn_grid_encode_torch((x * 0.25 + 0.5).cpu(), f.encoding.hash_table.detach().cpu(), lv, 12, True), atol=2e-5)
    geo = f.forward_geonetwork(x)                                      # eval path: imported grid through K2
    assert geo.shape == (64, 1 + 256) and torch.isfinite(geo).all()
    for p_ in f.parameters():
        p_.requires_grad_(True)
    with pytest.raises(NotImplementedError):
        f.get_sdf_at_pos(x)                                            # differentiable path: not supported for an imported grid
