"""Quick A/B of the C4 step on the GPU box: one process per library build / integration precision.

    [RB_LIB_PATH=radiobear_b200/lib/librb_<variant>.so] python tools/ab_quick.py <label> [f64|mixed] [reps]

Prints and appends to gpurun_out/ab_quick.jsonl one line: per-kernel device times (library event ring), step time
(CUDA events around geometry || alpha -> prepare -> integrate, L2 flushed between steps) and, when a reference
file /tmp/ab_ref_tb.npy exists (written by the first f64 run), the largest |Tb - reference| over the cube.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    label = sys.argv[1]
    precision = sys.argv[2] if len(sys.argv) > 2 else 'f64'
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    import torch
    import bench
    from radiobear_b200 import engine, _lib
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    ctx = _lib.get_context(0)
    ctx.enable_timing(True)
    ctx.set_rt_precision(precision)
    atm, freqs, grid = bench.workload()
    if os.environ.get("RB_AB_SAMEFREQ", ""):       # every frequency the same: no pace differences between the warps of a CTA
        freqs = np.full(len(freqs), float(os.environ["RB_AB_SAMEFREQ"]))
    cfg = atm.config
    n = len(grid)
    pts = np.stack([np.tile(grid, n), np.repeat(grid, n)], axis=1)
    if os.environ.get('RB_AB_SORT', ''):            # upper bound of what ordering the rays by work would give
        blk = int(os.environ['RB_AB_SORT'])
        key = np.hypot(pts[:, 0], pts[:, 1] * (71492.0 / 66854.0))
        if blk > 1:                                   # sort within blocks of `blk` consecutive rays only
            order = np.concatenate([s0 + np.argsort(key[s0:s0 + blk], kind='stable') for s0 in range(0, len(pts), blk)])
        else:
            order = np.argsort(key, kind='stable')
        pts = np.ascontiguousarray(pts[order])
    t64 = dict(dtype=torch.float64, device=dev)
    freqs_t = torch.tensor(freqs, **t64)
    T_t = torch.tensor(atm.gas[cfg.C['T']], **t64)
    P_t = torch.tensor(atm.gas[cfg.C['P']], **t64)
    gas_t = torch.tensor(atm.gas, **t64).contiguous()
    radius_t = torch.tensor(atm.property[cfg.LP['R']], **t64)
    nidx = atm.property[cfg.LP['N']]
    b_t = torch.tensor(pts, **t64).contiguous()
    L, F = atm.gas.shape[1], len(freqs)
    slab_t = torch.empty((L, F), **t64)
    tb_t = torch.empty((len(pts), F), dtype=torch.float64, device=dev)
    forms = [(c, f) for c, f in sorted(cfg.constituent_alpha.items()) if f is not None]
    other = {'h2': {'h2state': cfg.h2state}}
    orient = [float(cfg.orientation[0]), float(cfg.orientation[1])]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        engine.geometry_prefetch_dev(radius_t, nidx[0], nidx[1], b_t, cfg.Req, cfg.Rpol, orient, cfg.gtype, cfg.limb, ctx=ctx)
        engine.alpha_layers_dev(freqs_t, T_t, P_t, gas_t, cfg.C, formalisms=forms, other_dicts=other,
                                truncate_strength=cfg.truncate_strength, out=slab_t, freqs_host=freqs, ctx=ctx)
        engine.rt_batch_dev(radius_t, nidx[0], nidx[1], b_t, slab_t, T_t, cfg.Req, cfg.Rpol, orient, cfg.gtype, cfg.limb,
                            out_f32=False, tau_cut=engine.TAU_CUT, out=tb_t, ctx=ctx)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    k0 = {w: ctx.kernel_timed_count(w) for w in ('alpha', 'geometry', 'rt')}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i in range(reps):
        flush.zero_()
        ev[i][0].record()
        step()
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    out = {'label': label, 'precision': precision, 'lib': os.path.basename(_lib.LIB_PATH), 'reps': reps,
           'step_ms': float(np.mean(ms)), 'step_ms_min': float(np.min(ms))}
    for w in ('alpha', 'geometry', 'rt'):
        h = ctx.kernel_ms_history(w, ctx.kernel_timed_count(w) - k0[w])
        out[w + '_ms'] = float(np.mean(h)) if len(h) else None
    ctx.count_steps(True)
    step()
    torch.cuda.synchronize()
    out['steps_executed'] = ctx.count_steps(False)
    out['steps_small'] = ctx.count_small_steps()
    tb = tb_t.cpu().numpy()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    ref_fn = '/tmp/ab_ref_tb.npy'                      # 185 MB: stays on the GPU box
    if os.path.exists(ref_fn):
        ref = np.load(ref_fn)
        same_nan = bool(np.array_equal(np.isnan(tb), np.isnan(ref)))
        ok = ~np.isnan(ref)
        d = np.abs(tb - ref)[ok]
        out.update(max_abs_dTb_K=float(d.max()), mean_abs_dTb_K=float(d.mean()), nan_pattern_equal=same_nan,
                   p999_abs_dTb_K=float(np.quantile(d, 0.999)))
    elif precision == 'f64':
        np.save(ref_fn, tb)
    line = json.dumps(out)
    print(line, flush=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'ab_quick.jsonl'), 'a') as fh:
        fh.write(line + '\n')


if __name__ == '__main__':
    main()
