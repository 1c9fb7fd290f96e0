"""One NeuSky training iteration as ONE CUDA graph.

An `ns-train neusky` iteration at the README batch (1024 rays) is ~2900 kernel launches of a few microseconds each: run
eagerly it is bound by the HOST (Python + ctypes + torch dispatch: 42.9 ms of launch time for ~30 ms of kernels, the GPU
at 350 W; profiles/r02_bench_train_n1.json `host_launch_ms_per_step`).  The iteration has static shapes -- the icosphere is
centrally symmetric, so exactly half of its directions are in the upper hemisphere whatever the random rotation -- and,
after round 2's clean-up, no host synchronisation (no `.item()`, no pageable host->device copy, `inv_s` read on the device:
`nsk_neus_composite_*_dv`; scripts/train_sync_probe.py lists what is left), so zero-fill of the gradient buckets + forward
(neusky_model.py:553-1068) + DDF fitting pass (neusky_pipeline.py:272-289) + backward are captured once and replayed.

    it = GraphedTrainIteration(step, reducer, optimizer, fit=DDFFit(step))
    loss = it(batch_host, dirs_host, grid_positions, grid_dirs, sky_origins, sky_directions)      # per iteration

What stays outside the graph, per iteration: the host's random draws (icosphere rotation and its hemisphere compaction, the
DDF-fit samplers: torch's CPU generator, bit-exact with the reference) and their copies into the graph's static inputs, the
gradient all-reduce (`GradBucketReducer.finish()`: every bucket in order over NCCL, not overlapped with the backward -- 144 MiB
is ~0.3 ms of NVLink time) and the optimizer step.  Host scalars that are baked into kernel arguments (`cos_anneal_ratio`,
the proposal sampler's annealing slope) are part of the capture KEY together with every input shape: when the key changes the
iteration runs eagerly again and is re-captured once the key has been stable for `eager_warmup` iterations.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib

Tensor = torch.Tensor


class GraphedTrainIteration:
    def __init__(self, step, reducer, optimizer, fit=None, graph: bool = True, eager_warmup: int = 2, overlap_fit: bool = True):
        """`step`: train.NeuSkyTrainStep; `reducer`: parallel.GradBucketReducer over its parameters (the .grad views are the
        graph's static gradient buffers); `optimizer`: anything with `.step()`, stepped after the reduce (None: the caller steps);
        `fit`: ddf_fit.DDFFit or None.  Training only: every call runs a backward, so gradients must be enabled.
        `graph=False` runs every iteration eagerly through the same code (the comparison arm of bench.py and the tests).
        `overlap_fit`: run the DDF fitting pass and the RENI++ radiance decodes as parallel branches of the iteration (side streams,
        fork / join inside the captured graph) instead of in line with the main pass."""
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedTrainIteration needs a CUDA device (neusky_b200 has no CPU path)")
        self.step, self.red, self.opt, self.fit = step, reducer, optimizer, fit
        self.enabled, self.eager_warmup = bool(graph), max(1, int(eager_warmup))
        self._key, self._seen, self._graph = None, 0, None
        self._static: Dict[str, Tensor] = {}
        self.loss: Optional[Tensor] = None
        self.losses: Dict[str, Tensor] = {}
        self.kernels_in_graph = 0
        self.replays = self.eager_steps = self.captures = 0
        # Every iteration -- eager or captured -- runs on ONE dedicated stream: autograd's AccumulateGrad nodes remember the stream they
        # were created on, and a node created by an eager iteration on the legacy default stream that is still alive when the capture
        # starts makes the capture depend on the default stream (cudaErrorStreamCaptureImplicit).
        self._stream = torch.cuda.Stream(device=step.dev)
        self.overlap_fit = bool(overlap_fit)
        self._fit_stream = torch.cuda.Stream(device=step.dev)
        # parameters used by several branches get their gradient contributions from several streams: intended here (the engine orders
        # them); silence torch's per-backward warning about it
        _quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if _quiet is not None and self.overlap_fit:
            _quiet(False)
        self._aux_stream = torch.cuda.Stream(device=step.dev)

    # ------------------------------------------------------------------------------------------ host side of an iteration
    def _host_inputs(self, batch, dirs, grid_positions, grid_dirs, sky_origins, sky_directions) -> Dict[str, Tensor]:
        st = self.step
        d0, mask_u8, dirs_sel, sel = st.compact_directions(dirs)
        inp: Dict[str, Tensor] = {}
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                inp[f"batch.{k}"] = v.to(torch.int32) if k == "cam" else v
            elif isinstance(v, (list, tuple)):                       # e.g. "jitters": one [R] tensor per sampler level
                for i, t in enumerate(v):
                    inp[f"batch.{k}#{i}"] = t
        inp.update({"dirs": d0, "mask_u8": mask_u8, "dirs_sel": dirs_sel, "sel_index": sel})
        if grid_positions is not None:
            inp["grid_positions"], inp["grid_dirs"] = grid_positions, grid_dirs
        if self.fit is not None:
            o, d, mv = self.fit.draw_host()
            inp["fit.origins"], inp["fit.directions"] = o, d
            if mv is not None:
                inp["fit.multi_view_points"] = mv
            if sky_origins is not None:
                inp["fit.sky_origins"], inp["fit.sky_directions"] = sky_origins, sky_directions
        return inp

    def _key_of(self, inp: Dict[str, Tensor]):
        st = self.step
        shapes = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in inp.items()))
        return shapes + (float(st.cos_anneal_ratio), float(st.proposal_anneal), bool(torch.is_grad_enabled()))

    # ------------------------------------------------------------------------------------------ device side (captured)
    def _iteration(self):
        s, st = self._static, self.step
        st.dirs, st.mask_u8, st.dirs_sel, st.sel_index = s["dirs"], s["mask_u8"], s["dirs_sel"], s["sel_index"]
        batch: Dict[str, object] = {}
        for k in sorted(s):
            if not k.startswith("batch."):
                continue
            name, _, idx = k[6:].partition("#")
            if idx:
                batch.setdefault(name, []).append(s[k])              # keys sort as #0, #1, ... (fewer than ten levels)
            else:
                batch[name] = s[k]
        # Branches only where no collective is launched from inside the backward: an EAGER multi-GPU iteration all-reduces each bucket
        # from a gradient hook, and NCCL orders that collective after the hook's stream only -- gradients of the same bucket written
        # by another branch's stream would race with it.  Captured iterations (hooks silent, buckets reduced after the replay) and
        # single-GPU runs have no such collective.
        overlap = self.overlap_fit and (self.red.capturing or self.red.world == 1)
        st.aux_stream = self._aux_stream if overlap else None               # RENI++ radiance as a parallel branch of the main forward
        self.red.zero_grad()
        run_fit = lambda: self.fit(s.get("fit.sky_origins"), s.get("fit.sky_directions"), rays=(s["fit.origins"], s["fit.directions"]),
                                   multi_view_points=s.get("fit.multi_view_points"))
        fit_res = None
        if self.fit is not None and overlap:
            # The DDF fitting pass is independent of the main pass until the two losses are added, and its kernels are small (1024 rays,
            # 2304 DDF rows): it runs as a second BRANCH (its own stream; a fork / join inside the captured graph) next to the main
            # pass's 328k-row kernels instead of in front of them.  Autograd runs each branch's backward on the branch's stream.
            cur = torch.cuda.current_stream()
            shared = list(st.sdf_weights())                       # what both branches read is produced before the fork
            for f in st.proposal_fields or ():
                f.refresh()
            self._fit_stream.wait_stream(cur)
            with torch.cuda.stream(self._fit_stream):
                fit_res = run_fit()
            for t in shared:
                t.record_stream(self._fit_stream)
        loss, losses, _out = st(batch, grid_positions=s.get("grid_positions"), grid_dirs=s.get("grid_dirs"))
        if self.fit is not None:
            if fit_res is None:
                fit_res = run_fit()
            else:
                torch.cuda.current_stream().wait_stream(self._fit_stream)
                for t in (fit_res[0], *fit_res[1].values()):
                    t.record_stream(torch.cuda.current_stream())
            fl, fls = fit_res[0], fit_res[1]
            loss = loss + fl
            losses = {**losses, **{"ddf_fit." + k: v for k, v in fls.items()}}
        loss.backward()
        return loss.detach(), {k: v.detach() for k, v in losses.items()}

    def _capture(self) -> None:
        st = self.step
        # version-keyed host caches (weight-norm fold, packed proposal MLPs) must be rebuilt INSIDE the graph, not reused from an eager step
        st._sdf_w_key, st._sdf_w = None, None
        for f in st.proposal_fields or ():
            f._mlp_key = None
        torch.cuda.synchronize(st.dev)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launches
        self.red.capturing = True
        try:
            with torch.cuda.graph(g, stream=self._stream, capture_error_mode="thread_local"):
                self.loss, self.losses = self._iteration()
        finally:
            self.red.capturing = False
        self.kernels_in_graph = _lib.launches - n0        # ABI launches recorded into the graph (torch's own glue kernels not counted)
        _lib.launches = n0
        self._graph = g
        self.captures += 1

    # ------------------------------------------------------------------------------------------ one iteration
    def __call__(self, batch: Dict[str, Tensor], dirs: Tensor, grid_positions: Optional[Tensor] = None, grid_dirs: Optional[Tensor] = None,
                 sky_origins: Optional[Tensor] = None, sky_directions: Optional[Tensor] = None) -> Tensor:
        """batch: the ray batch of NeuSkyTrainStep.forward (host, ideally pinned, or device tensors); dirs [D,3]: this iteration's
        illumination directions (host).  Returns the iteration's total loss (a device scalar, overwritten by the next call); the
        parameters have been updated when it returns (asynchronously, on the current stream)."""
        st = self.step
        if not torch.is_grad_enabled():
            raise RuntimeError("GraphedTrainIteration: called under torch.no_grad(); an iteration includes the backward pass")
        inp = self._host_inputs(batch, dirs, grid_positions, grid_dirs, sky_origins, sky_directions)
        key = self._key_of(inp)
        if key != self._key:
            self._key, self._seen, self._graph = key, 0, None
            self._static = {k: torch.empty(tuple(v.shape), dtype=v.dtype, device=st.dev) for k, v in inp.items()}
        with torch.cuda.device(st.dev):
            caller = torch.cuda.current_stream()
            self._stream.wait_stream(caller)
            with torch.cuda.stream(self._stream):
                for k, v in inp.items():
                    self._static[k].copy_(v, non_blocking=True)
                if self._graph is None and self.enabled and self._seen >= self.eager_warmup:
                    self._capture()                          # records, does not run: the replay below is this iteration
                if self._graph is not None:
                    self.red.rearm()
                    self._graph.replay()
                    _lib.launches += self.kernels_in_graph
                    self.replays += 1
                else:
                    self.loss, self.losses = self._iteration()
                    self._seen += 1
                    self.eager_steps += 1
                self.red.finish()
                if self.opt is not None:
                    self.opt.step()
            caller.wait_stream(self._stream)                 # the caller's stream sees the updated parameters and the loss
        return self.loss
