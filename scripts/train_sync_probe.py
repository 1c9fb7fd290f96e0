"""Diagnostics: list every host<->device synchronisation inside one training iteration (bench.py --workload train).
torch.cuda.set_sync_debug_mode("warn") flags .item(), pageable host->device copies, nonzero, ...; each warning is printed
once per Python call site with the last frames of its stack.  What a CUDA-graph capture of the step must not contain."""
import os
import sys
import traceback
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

seen = {}


def show(message, category, filename, lineno, file=None, line=None):
    if "synchroniz" not in str(message):
        return
    st = [f for f in traceback.extract_stack()[:-1] if "/neusky_b200/" in f.filename or f.filename.endswith("bench.py")]
    key = tuple((f.filename, f.lineno) for f in st[-3:])
    if key in seen:
        seen[key] += 1
        return
    seen[key] = 1
    print("SYNC:", " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno} {f.name}" for f in reversed(st[-4:])), flush=True)


warnings.showwarning = show
warnings.simplefilter("always")

orig_forward = None
calls = {"n": 0}


def main():
    from neusky_b200.train import NeuSkyTrainStep

    fwd = NeuSkyTrainStep.forward

    def wrapped(self, *a, **k):
        calls["n"] += 1
        if calls["n"] == 3:          # after two eager warm-up iterations
            torch.cuda.set_sync_debug_mode("warn")
            print("---- sync debug on ----", flush=True)
        return fwd(self, *a, **k)

    NeuSkyTrainStep.forward = wrapped
    sys.argv = [sys.argv[0], "--workload", "train", "--steps", "2", "--warmup", "2"] + sys.argv[1:]
    bench.main()
    torch.cuda.set_sync_debug_mode("default")
    print("unique sync sites:", len(seen), "occurrences:", sum(seen.values()))


if __name__ == "__main__":
    main()
