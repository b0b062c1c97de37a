"""FP64 pipe sharing on the GPU box: DFMA rate next to MUFU.RCP64H and with three register operands."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiobear_b200 import _lib
ctx = _lib.get_context(0)
iters = 20000
print('rb_probe_fp64_peak: {:.2f} TFLOP/s'.format(ctx.fp64_peak_tflops(iters)))
for rrr, nm in ((0, 0), (2, 0), (3, 0), (1, 0), (0, 2), (0, 4), (0, 8), (1, 2), (1, 4), (1, 8)):
    if True:
        ms = C.c_double(0.0)
        ctx.check(ctx.lib.rb_probe_fp64_mix(ctx.h, iters, nm, rrr, C.byref(ms)))
        n_thr = 148 * 8 * 256
        tf = 2.0 * 16 * iters * n_thr / (ms.value * 1e-3) / 1e12
        print('rrr={} mufu_per_16_dfma={}: {:.3f} ms  DFMA {:.2f} TFLOP/s  ({:.1f} clk per warp-trip per SMSP at 1.965 GHz)'.format(
            rrr, nm, ms.value, tf, ms.value * 1e-3 * 1.965e9 / iters / 16))
