"""Per-kernel roofline microbenchmarks at full size: one JSON object per kernel on stdout (see bench_kernels.py)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _peaks
from bench_kernels import kernel_rooflines

kernel_rooflines(torch.device("cuda:0"), _peaks(), 1.0, emit=lambda d: print(json.dumps(d), flush=True))
