"""The captured training iteration (neusky_b200/graphed.py) against the same iterations run eagerly: identical inputs, identical
initial parameters, SGD updates (smooth in the gradients: Adam's first steps are lr * sign(g), which turns summation-order noise on
near-zero gradient elements into +-lr parameter differences and a few per cent of loss between ANY two runs) -> the losses and every
gradient bucket must agree iteration by iteration (to rounding level on the first iteration, where the parameters are identical; at
1e-4 afterwards), the graph must really have been replayed -- and must see the
optimizer's updates and each iteration's own inputs -- and a change of a baked-in host scalar must fall back to eager execution and
re-capture."""
import numpy as np
import pytest
import torch

from neusky_b200 import init as nb_init

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neusky_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


def _case(R, K, seed):
    g = torch.Generator().manual_seed(seed)
    o = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * (0.5 + 0.3 * torch.rand(R, 1, generator=g))
    d = torch.nn.functional.normalize(-o + 0.12 * torch.randn(R, 3, generator=g), dim=-1)
    return {"origins": o, "directions": d, "dnorm": torch.ones(R, 1), "cam": torch.randint(0, K, (R,), generator=g).to(torch.int32),
            "image": torch.rand(R, 3, generator=g), "fg": (torch.rand(R, generator=g) > 0.3).float(), "ground": (torch.rand(R, generator=g) > 0.7).float(),
            "sky": (torch.rand(R, generator=g) > 0.8).float()}


def _build(dev, proposal, with_fit, graph):
    from neusky_b200 import train as T
    from neusky_b200.ddf_fit import DDFFit, DDFSamplerConfig, VMFDDFSampler
    from neusky_b200.graphed import GraphedTrainIteration
    from neusky_b200.parallel import GradBucketReducer
    from oracle import sampler_oracle as SO

    log2_T, S, K = 14, 12, 3
    g = torch.Generator().manual_seed(41)
    sdf_p = nb_init.init_sdf_params(3, log2_T=log2_T)
    sdf_p["encoding.hash_table"] = (torch.rand(sdf_p["encoding.hash_table"].shape, generator=g) * 2 - 1) * 0.05
    sdf_p["deviation_network.variance"] = torch.tensor(0.25)
    ddf_p = nb_init.init_ddf_params(5, final_gain=8.0, log2_T=log2_T, table_scale=0.1)
    nets = [SO.init_proposal_net(1, table_scale=1.0, density_bias=1.0), SO.init_proposal_net(2, table_scale=1.0, density_bias=2.0)] if proposal else None
    step = T.NeuSkyTrainStep(sdf_p, ddf_p, nb_init.init_reni_params(8), num_cameras=K, device=dev, log2_T=log2_T, num_samples=S, split_geo=3, split=3,
                             threshold_init=0.4, proposal_params=nets, num_proposal_samples_per_ray=(32, 20))
    with torch.no_grad():
        step.latents.copy_(torch.randn(K, 100, 3, generator=g).to(dev))
    params = [p for p in step.parameters() if p.requires_grad]
    red = GradBucketReducer(params, big_bytes=64 << 10)
    opt = torch.optim.SGD(params, lr=1e-4)
    fit = None
    if with_fit:
        fit = DDFFit(step, sampler=VMFDDFSampler(DDFSamplerConfig(num_samples_on_sphere=2, num_rays_per_sample=8), ddf_sphere_radius=step.radius, device=dev))
    # the eager arm also keeps the DDF fitting pass AFTER the main pass on one stream (the reference's order); the graphed arm runs it as
    # a parallel branch on a second stream, so the comparison covers the fork / join as well as the capture
    return step, red, GraphedTrainIteration(step, red, opt, fit=fit, graph=graph, eager_warmup=1, overlap_fit=graph)


def _run(dev, proposal, with_fit, graph, n_iter, anneal_change_at=None, repeat_inputs=False):
    from scipy.spatial.transform import Rotation

    from oracle import neusky_oracle as O

    R, K = 16, 3
    step, red, it = _build(dev, proposal, with_fit, graph)
    base = O.icosphere_directions(100).double().numpy()
    rots = Rotation.random(n_iter, random_state=np.random.RandomState(5)).as_matrix()
    g = torch.Generator().manual_seed(77)
    torch.manual_seed(1234)                                   # the DDF-fit samplers draw from torch's CPU generator
    sky_o = (torch.tensor([0.0, -0.6, 0.1]).expand(8, 3) + 0.1 * torch.randn(8, 3, generator=g)).to(dev)
    sky_d = torch.nn.functional.normalize(torch.randn(8, 3, generator=g) + torch.tensor([0.0, 0.0, 1.0]), dim=-1).to(dev)
    losses, grads = [], []
    for i in range(n_iter):
        if anneal_change_at is not None and i == anneal_change_at:
            step.cos_anneal_ratio = 0.5
        if not (repeat_inputs and i > 1):                      # repeat_inputs: iterations 1, 2, ... all get iteration 1's inputs
            batch = _case(R, K, 100 + i)
            if proposal:
                batch["jitters"] = [torch.rand(R, generator=g) for _ in range(3)]
            dirs = torch.from_numpy((base @ rots[i]).astype(np.float32))
            gp, gd = torch.rand(27, 3, generator=g) * 2 - 1, torch.nn.functional.normalize(torch.randn(27, 3, generator=g), dim=-1)
        loss = it(batch, dirs, gp, gd, sky_o if with_fit else None, sky_d if with_fit else None)
        losses.append(float(loss))
        grads.append([b.detach().double().cpu() for b in red.buckets])
    return losses, grads, it


@pytest.mark.parametrize("proposal,with_fit", [(True, False), (False, True)])
def test_graphed_iteration_matches_eager(dev, proposal, with_fit):
    n = 5
    le, ge, it_e = _run(dev, proposal, with_fit, False, n)
    lg, gg, it_g = _run(dev, proposal, with_fit, True, n)
    assert it_e.replays == 0 and it_e.eager_steps == n
    assert it_g.captures == 1 and it_g.eager_steps == 1 and it_g.replays == n - 1 and it_g.kernels_in_graph > 50
    for i in range(n):
        tol_l, tol_g = (2e-6, 1e-5) if i == 0 else (1e-4, 5e-3)
        assert abs(le[i] - lg[i]) <= tol_l * max(1.0, abs(le[i])), (i, le, lg)
        for a, b in zip(ge[i], gg[i]):
            assert float((a - b).norm()) <= tol_g * float(a.norm()) + 1e-12, (i, float((a - b).norm()), float(a.norm()), le, lg)
    assert len(set(lg)) == n                                   # every replay saw its own inputs


def test_graphed_iteration_recaptures_when_a_baked_scalar_changes(dev):
    n = 6
    le, ge, _ = _run(dev, True, False, False, n, anneal_change_at=3)
    lg, gg, it_g = _run(dev, True, False, True, n, anneal_change_at=3)
    # iteration 0 eager, 1-2 replays of graph 1; the key changes at 3: eager again, then graph 2 for 4-5
    assert it_g.captures == 2 and it_g.eager_steps == 2 and it_g.replays == 4
    for i in range(n):
        assert abs(le[i] - lg[i]) <= 1e-4 * max(1.0, abs(le[i])), (i, le, lg)


def test_graph_replays_see_the_optimizer_updates(dev):
    """Identical inputs on consecutive replays: the loss must still move (the replayed kernels read the parameters the optimizer
    just updated, including the re-folded weight-norm weights and the re-packed proposal MLPs) and follow the eager run."""
    n = 4
    le, _, _ = _run(dev, True, False, False, n, repeat_inputs=True)
    lg, _, it_g = _run(dev, True, False, True, n, repeat_inputs=True)
    assert it_g.replays == n - 1
    assert lg[1] != lg[2] and lg[2] != lg[3], lg
    for i in range(n):
        assert abs(le[i] - lg[i]) <= 1e-4 * max(1.0, abs(le[i])), (le, lg)
