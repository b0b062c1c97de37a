#!/usr/bin/env python
"""Run the C5 alpha workload a few times (for ncu captures of alpha_lines_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from radiobear_b200 import _lib
ctx = _lib.get_context(0)
print(bench.alpha_c5(ctx, torch.device('cuda', 0), reps=3))
