#!/usr/bin/env python
"""bench.py -- throughput of the RadioBEAR hot path on B200 (config C4 of BASELINE.json).

Workload: Jupiter full image, b = 0.005 (601 x 601 pixels, ~117 k on the disc), 64 frequencies
1..100 GHz, 1000-layer default atmosphere.  One "step" = one pass of the whole hot path:
alpha_lines (1000 layers x 64 freqs x ~1400 catalog lines) -> ray_geometry (every pixel) ->
rt_integrate (every on-disc pixel x 64 freqs x 999 segments).  Nothing is cached between steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1: launched by torchrun, one rank per GPU; image rows are sharded (balanced by on-disc pixels),
every rank recomputes the 0.5 MB alpha slab, one NCCL gather brings Tb to rank 0 ("strong" scaling:
the image is fixed).  Prints ONE JSON line on rank 0.

metric  Tb pixel*freq / s over the on-disc pixels of the cube (off-disc pixels are computed too and
        reported separately: they cost one findEdge scan).
value   device-resident: inputs in HBM, CUDA events around each step on the launching stream, L2 flushed
        between steps, max over ranks.
e2e     the public call `Planet.run(freqs, b=0.005)` with host buffers: host->device copies of the
        impact points / atmosphere and the device->host copy of the float32 cube are inside the timed
        region (wall clock between device synchronisations).
--impl reference  times the CPU restatement of the reference algorithm (oracle/, kind "port": the
        reference itself is pure Python living in /root/reference, which does not exist on the GPU box)
        on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BSTEP = 0.005
NFREQ = 64
METRIC = 'Tb pixel*freq/s (Jupiter image cube b=0.005, 64 freqs 1-100 GHz, on-disc pixels)'
UNIT = 'pixel*freq/s'
# rt_integrate_pairs_kernel (two frequencies per thread), per executed (ray, freq, segment) step (DESIGN.md 3.3;
# counted in the SASS of the two hot loops, profiles/r2_sass_rt_integrate_pairs.txt: 8 steps per trip):
#   table phase      6.5 DFMA + 2 DMUL + 2 DADD = 10.5 FP64 instructions, 17 flops   (ds_i + ds_i+1 is shared by the pair)
#   small-tau phase (tau < 2^-11)  4 DFMA + 1 DMUL + 1.5 DADD = 6.5 FP64 instructions, 10.5 flops
RT_FLOPS_PER_STEP = 17.0
RT_FP64_INSTR_PER_STEP = 10.5
RT_FLOPS_PER_SMALL_STEP = 10.5
RT_FP64_INSTR_PER_SMALL_STEP = 6.5
# dram__bytes_read.sum + dram__bytes_write.sum of one rt_integrate_pairs_kernel launch of this workload at N=1
RT_DRAM_BYTES_N1 = 1032517000 + 64695296
RT_DRAM_SOURCE = ('ncu --set full, profiles/r2_final_rt_integrate_pairs_ncu_full.txt (1.033 GB read + 0.065 GB written; the tile list '
                  'launched in one part instead of L2-sized parts reads 4.40 GB: profiles/r2_final_rt_kernels_one_part_ncu_full.txt)')
RT_DRAM_BYTES_N1_MIXED = 645195776
RT_DRAM_SOURCE_MIXED = 'ncu --set full, profiles/r1_mixed_rt_integrate_rays_mixed_ncu_full.txt (0.505 GB read + 0.140 GB written)'
# SASS instructions (cuobjdump, hot loops) and shared-memory wavefronts (128 B/clk/SM data pipe) per executed segment-step
# of the integration kernels, {precision: (phase A, phase B)}
RT_SASS_PER_STEP = {'f64': (10.0, 19.25), 'mixed': (10.0, 15.25)}
RT_SMEM_WAVEFRONTS_PER_STEP = {'f64': (2.5, 8.0), 'mixed': (2.0, 2.0)}
RT_KERNEL = {'f64': 'rt_integrate_pairs_kernel', 'mixed': 'rt_integrate_rays_mixed_kernel'}
WORKLOAD = 'C4: Jupiter full image b=0.005 (601x601 px) x 64 freqs 1-100 GHz, 1000 layers, alpha+geometry+RT per step'


def workload():
    from radiobear_b200.atmosphere import Atmosphere
    from radiobear_b200 import set_utils
    atm = Atmosphere.from_npz(os.path.join(ROOT, 'tests', 'golden', 'atm_jupiter.npz'), 'jupiter')
    freqs = np.linspace(1.0, 100.0, NFREQ)
    grid = set_utils.image_grid(BSTEP)
    return atm, freqs, grid


def on_disc_mask(grid, q):
    xx, yy = np.meshgrid(grid, grid)
    return xx**2 + (yy / q)**2 < 1.0


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# ------------------------------------------------------------------------------------------------
def _cpu_pixels(args):
    """Worker: geometry + integration of a few pixels with the oracle (one process = one core)."""
    blist, lay, seed = args
    from oracle import ray_oracle as ro, rt_oracle as rto
    atm, freqs, grid = workload()
    cfg = atm.config
    T = atm.gas[cfg.C['T']]
    out = []
    for b in blist:
        ray = ro.compute_ds(atm.property[cfg.LP['R']], atm.property[cfg.LP['N']], b, cfg.Req, cfg.Rpol, cfg.orientation,
                            cfg.gtype, cfg.limb)
        out.append(rto.integrate_ray(ray['ds'], ray['layer4ds'], lay, T))
    return np.array(out)


def _cpu_alpha(args):
    layers, = args
    from oracle import alpha_oracle as ao
    atm, freqs, grid = workload()
    cfg = atm.config
    return ao.get_layers(freqs, atm.gas, atm.cloud, cfg.C, cfg.Cl, cfg.constituent_alpha,
                         other_dicts={'h2': {'h2state': cfg.h2state}}, truncate_strength=cfg.truncate_strength,
                         layers=layers)


def cpu_sample(cores, n_pix_per_core=6, n_lay_per_core=24, pool=None):
    """Time the oracle on a bounded sample and extrapolate linearly to the full job.

    Returns (value pixel*freq/s for the whole job on `cores` cores, description, seconds spent)."""
    atm, freqs, grid = workload()
    q = atm.config.Rpol / atm.config.Req
    mask = on_disc_mask(grid, q)
    n_on = int(mask.sum())
    L = atm.gas.shape[1]
    rng = np.random.default_rng(0)
    iy, ix = np.nonzero(mask)
    pick = rng.choice(len(iy), n_pix_per_core * cores, replace=False)
    pix = [[grid[ix[k]], grid[iy[k]]] for k in pick]
    lay_all = sorted(rng.choice(L, min(L, n_lay_per_core * cores), replace=False).tolist())
    t00 = time.perf_counter()
    if pool is None:
        t0 = time.perf_counter()
        slab = _cpu_alpha((lay_all,))
        t_alpha = time.perf_counter() - t0
        # integration needs a full [F][L] slab: tile the sampled layers (timing only)
        lay_full = np.tile(slab, (1, L // slab.shape[1] + 1))[:, :L]
        t1 = time.perf_counter()
        _cpu_pixels((pix, lay_full, 0))
        t_pix = time.perf_counter() - t1
    else:
        chunks = [lay_all[i::cores] for i in range(cores)]
        t0 = time.perf_counter()
        slabs = pool.map(_cpu_alpha, [(c,) for c in chunks if c])
        t_alpha = time.perf_counter() - t0
        cat = np.concatenate(slabs, axis=1)
        lay_full = np.tile(cat, (1, L // cat.shape[1] + 1))[:, :L]
        t1 = time.perf_counter()
        pool.map(_cpu_pixels, [(pix[i::cores], lay_full, i) for i in range(cores)])
        t_pix = time.perf_counter() - t1
    full = t_alpha * (L / len(lay_all)) + t_pix * (n_on / len(pix))
    value = n_on * len(freqs) / full
    desc = ('oracle port, {} core(s): alpha on {} of {} layers x {} freqs in {:.2f} s, geometry+RT on {} of {} on-disc '
            'pixels x {} freqs in {:.2f} s, extrapolated linearly to the full cube ({:.0f} s)'
            .format(cores, len(lay_all), L, len(freqs), t_alpha, len(pix), n_on, len(freqs), t_pix, full))
    return value, desc, time.perf_counter() - t00


def stock_reference_record():
    """The UNMODIFIED reference timed next to the port on the same sample -- in the build container, where
    /root/reference exists (tools/time_stock_reference.py); the GPU box has no reference to run."""
    try:
        rec = json.load(open(os.path.join(ROOT, 'profiles', 'r2_stock_reference_timing.json')))
    except (OSError, ValueError):
        return None
    return {'kind': 'reference', 'value': rec['reference']['value'], 'unit': rec['reference']['unit'], 'cores': 1,
            'port_value_same_host': rec['port']['value'], 'port_over_reference': rec['port_over_reference'],
            'sample': rec['sample'], 'where': rec['host']['where'],
            'source': 'profiles/r2_stock_reference_timing.json (tools/time_stock_reference.py)'}


def run_reference(args):
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, desc = [], ''
    per = max(2, min(16, 256 // cores))
    lay = max(2, 96 // cores)
    with mp.get_context('spawn').Pool(cores) as pool:
        pool.map(_cpu_alpha, [([0],)] * cores)                       # warm the workers (imports, catalogs)
        for i in range(args.warmup + args.steps):
            v, desc, _ = cpu_sample(cores, n_pix_per_core=per, n_lay_per_core=lay, pool=pool)
            if i >= args.warmup:
                vals.append(v)
    value = float(np.mean(vals))
    atm, freqs, grid = workload()
    n_on = int(on_disc_mask(grid, atm.config.Rpol / atm.config.Req).sum())
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * n_on * NFREQ / value, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic image grid over the Jupiter default atmosphere fixture',
            'config': {'workload': WORKLOAD, 'pixels': 'on-disc'},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc,
                             'stock_reference': stock_reference_record()},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device, period_s=0.004):
        self.fn = tempfile.NamedTemporaryFile(prefix='rbclk', suffix='.csv', delete=False).name
        self.device = device
        self.p = None
        self.period_s = period_s
        self.thread = None
        self.samples = []          # (sm MHz, max sm MHz, reasons bitmask) from the NVML thread
        self._stop = False

    def _nvml_loop(self, nv, h):
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), float(mx), int(rs)))
            except Exception:
                pass
            time.sleep(self.period_s)

    def start(self):
        # the timed region is tens of milliseconds: nvidia-smi's fastest loop (100 ms) would sample it once at best,
        # so the same counters are polled through NVML from a thread; nvidia-smi stays as the fallback
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = None
            try:                               # CUDA_VISIBLE_DEVICES may renumber the devices: go by UUID when torch has it
                import torch
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(int(self.device)).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, 'encode') else uuid)
            except Exception:
                h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(int(self.device))
            self._stop = False
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=open(self.fn, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            bits = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40, 'sw_power_cap': 0x4}
            if self.samples:
                sm = [x[0] for x in self.samples]
                hi = sorted(sm)[len(sm) // 2:]      # upper half = samples under load
                seen = 0
                for x in self.samples:
                    seen |= x[2]
                out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(x[1] for x in self.samples),
                           reasons=sorted(k for k, v in bits.items() if seen & v), samples=len(sm), source='nvml')
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.fn):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        try:
            os.unlink(self.fn)
        except OSError:
            pass
        if sm:
            hi = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       source='nvidia-smi')
        return out


def alpha_c5(ctx, dev, reps=5, full=False, world=1, rank=0):
    """full=True: the untrimmed NH3 line lists (415 + 1301 + 4198 = 5914 lines, SURVEY 8d "full catalog").
    Secondary metric of BASELINE.json: alpha layer*freq*line / s on config C5 (synthetic 4096-layer
    atmosphere x 4096 freqs x NH3 catalog, formalism nh3_dbs_sjs; SURVEY 8d: T~U(80,1800) K,
    P log-U(1e-2,5e3) bar, X_NH3 log-U(1e-7,1e-3), X_H2 = 0.86, X_He = 0.135, seed 0).

    world > 1: the layers are split into contiguous blocks, one per rank (layers rather than
    frequencies because a CTA builds its layer's line tables once -- parallel.alpha_layers_sharded); every rank computes
    its [L/world][F] block and one NCCL all_gather writes the blocks straight into the full [L][F] slab on every GPU
    (what the ray integration of every rank needs).  Timed with CUDA events around compute + collective, max over
    ranks; `value` is the whole job."""
    import torch
    import torch.distributed as dist
    from radiobear_b200 import engine, catalogs, parallel
    from oracle import alpha_oracle as ao
    rng = np.random.default_rng(0)
    L = F = 4096
    n_gross = (1301 + 4198) if full else 399
    reps = max(2, reps // 2) if full else reps
    C = {'Z': 0, 'T': 1, 'P': 2, 'H2': 3, 'HE': 4, 'NH3': 5}
    gas = np.zeros((6, L))
    gas[C['T']] = rng.uniform(80.0, 1800.0, L)
    gas[C['P']] = 10**rng.uniform(-2.0, np.log10(5e3), L)
    gas[C['H2']], gas[C['HE']] = 0.86, 0.135
    gas[C['NH3']] = 10**rng.uniform(-7.0, -3.0, L)
    freqs = np.linspace(1.0, 100.0, F)
    P = gas[C['P']]
    # lines actually evaluated per (layer, freq): 814 below 400 bar, 1014 in the 400..2000 bar blend, 200 above
    # (full catalog: 5914 / 6114 / 200)
    nlines = np.where(P < 400.0, 415 + n_gross, np.where(P > 2000.0, 200, 615 + n_gross))
    n_br = np.where(P < 400.0, 415, np.where(P > 2000.0, 200, 615))          # Ben-Reuven line-evals
    n_gr = np.where(P > 2000.0, 0, n_gross)                                   # Gross line-evals
    evals = float(nlines.sum()) * F
    flops = (10.0 * float(n_br.sum()) + 8.0 * float(n_gr.sum())) * F          # + 1 reciprocal each (not counted)
    t64 = dict(dtype=torch.float64, device=dev)
    parts = parallel.partition_even(L, world)
    lo, hi = parts[rank]
    assert L % world == 0
    g_t = torch.tensor(gas[:, lo:hi], **t64).contiguous()
    f_t, T_t, P_t = torch.tensor(freqs, **t64), torch.tensor(gas[C['T']][lo:hi], **t64), torch.tensor(P[lo:hi], **t64)
    # the full slab; this rank's block of layers is out[lo:hi].  N > 1: a symmetric allocation, so that the absorption
    # kernel can store its values straight into every GPU's copy over NVLink (the all_gather inside the kernel);
    # RB_BENCH_ALPHA_GATHER=nccl (or no symmetric memory) keeps the kernel + NCCL all_gather pair
    sym = None
    if world > 1 and os.environ.get('RB_BENCH_ALPHA_GATHER', 'kernel') == 'kernel':
        sym = parallel.symmetric_slab(L, F, dev)
        ok = torch.tensor([1 if sym.usable else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            sym = None
    out = sym.tensor if sym is not None else torch.empty((L, F), **t64)
    local = out[lo:hi]
    forms = [('nh3', 'nh3_dbs_sjs')]
    ms, ms_k = [], []
    catalogs.use_full_nh3_catalog(full)
    try:
        for i in range(reps + 2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            if sym is not None:
                engine.alpha_layers_dev(f_t, T_t, P_t, g_t, C, formalisms=forms, freqs_host=freqs, ctx=ctx,
                                        scatter=(sym.ptrs, lo))
                e1.record()
                sym.barrier()                                # everybody's stores have landed everywhere
            else:
                engine.alpha_layers_dev(f_t, T_t, P_t, g_t, C, formalisms=forms, out=local, freqs_host=freqs, ctx=ctx)
                e1.record()
                if world > 1:
                    dist.all_gather_into_tensor(out, local)  # in place: block r of the output is rank r's input
            e2.record()
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(e0.elapsed_time(e2))
                ms_k.append(e0.elapsed_time(e1))
    finally:
        catalogs.use_full_nh3_catalog(False)
    tt = torch.tensor([float(np.mean(ms)), float(np.mean(ms_k))], **t64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t, t_kernel = float(tt[0].item()) * 1e-3, float(tt[1].item()) * 1e-3
    if rank != 0:
        return None
    # parity spot check against the oracle + its speed on one host core (layers spread over the three pressure
    # branches, frequencies over every rank's block)
    lay = [7, 1234, 4000]
    t0 = time.perf_counter()
    ref = ao.get_layers(freqs, gas, np.zeros((1, L)), C, {}, {'nh3': 'nh3_dbs_sjs'}, layers=lay,
                        cat=ao.LineCatalog(full_nh3=True) if full else None)
    t_cpu = time.perf_counter() - t0
    got = out[lay].cpu().numpy().T
    rel = float(np.nanmax(np.abs(got - ref) / np.abs(ref)))
    cpu_rate = float(nlines[lay].sum()) * F / t_cpu
    return {'workload': 'C5: 4096 layers x 4096 freqs x NH3 (nh3_dbs_sjs{}), synthetic'.format(
                ', full catalog: 415 + 1301 + 4198 lines' if full else ', shipped catalog: 415 + 201 + 198 lines'),
            'metric': 'alpha layer*freq*line/s',
            'value': evals / t, 'ms': t * 1e3, 'kernel_ms': t_kernel * 1e3, 'n_gpus': world,
            'sharding': 'none' if world == 1 else (
                'contiguous blocks of layers per rank; the kernel stores every value into the [L][F] slab of all GPUs over '
                'NVLink (symmetric memory) + one device-side barrier: every rank ends with the full slab' if sym is not None else
                'contiguous blocks of layers per rank + one all_gather (NCCL) in place into the [L][F] slab: every rank ends '
                'with the full slab'),
            'gather': 'none' if world == 1 else ('in-kernel peer stores' if sym is not None else 'nccl all_gather'),
            'line_evals': evals, 'max_rel_err_vs_oracle': rel,
            'fp64_tflops_algorithmic': flops / t / 1e12,
            'flops_per_line_eval': '10 (Ben-Reuven) / 8 (Gross) + 1 reciprocal (SURVEY 8d)',
            'cpu_port_value_1core': cpu_rate}


def retrieval_loop(atm, iters=20):
    """the retrieval inner loop of scripts/demo_batch.py -- `run('1:10:1', save_alpha='memory')` once, then
    `run('1:10:1', scale=<dict>, get_alpha='memory')` per iteration (disc-averaged, 10 frequencies): only the scale-sum
    over the cached per-constituent cube and the radiative transfer are redone.  The cube stays on the device
    (rb_alpha_rescale_resident); the same loop with the cube re-uploaded from the host cache every iteration
    (rb_alpha_scale_sum, round 1's path) and a full recomputation (get_alpha='calc') are timed beside it.
    Wall clock per Planet.run call, host buffers in and out."""
    from radiobear_b200.planet import Planet
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    L = atm.gas.shape[1]
    P = atm.gas[atm.config.C['P']]
    rng = np.random.default_rng(3)

    def scale_of(i):
        return {'nh3': list(0.5 + rng.random() + 0.3 * np.sin(np.log10(P) + i)), 'h2o': [float(1.0 + 0.05 * i)] * L}
    p.run('1:10:1', save_alpha='memory')
    scales = [scale_of(i) for i in range(iters + 3)]
    out = {}
    ref = None
    for mode in ('resident', 'host_cube', 'recompute') * 2:      # two passes: the second is reported (everything warm)
        tbs = []
        for i, sc in enumerate(scales):
            if i == 3:
                t0 = time.perf_counter()
            if mode == 'host_cube':
                p.alpha[0]._dev_cube = None               # forget the device copy: the cube is uploaded again
            if mode == 'recompute':
                tbs.append(np.array(p.run('1:10:1', scale=sc, reuse_override='false').Tb))
            else:
                tbs.append(np.array(p.run('1:10:1', scale=sc, get_alpha='memory', reuse_override='false').Tb))
        out[mode + '_ms_per_iteration'] = 1e3 * (time.perf_counter() - t0) / iters
        if mode == 'resident':
            ref = tbs
            resident_used = p.alpha[0]._res is not None
        else:
            out['max_abs_dTb_K_resident_vs_' + mode] = float(max(np.max(np.abs(a - b)) for a, b in zip(ref, tbs)))
        if mode == 'host_cube':
            p.run('1:10:1', save_alpha='memory')          # put the cube back on the device for whoever comes next
    out.update(workload="Planet.run('1:10:1', b='disc', scale=dict(nh3, h2o), get_alpha='memory') after one save_alpha='memory' "
                        "(scripts/demo_batch.py), {} layers x 10 freqs x {} constituents".format(L, len(p.alpha[0].ordered_constituents)),
               iterations=iters, resident_path_used=bool(resident_used),
               iterations_per_s=1e3 / out['resident_ms_per_iteration'])
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from radiobear_b200 import engine, parallel, _lib
    from radiobear_b200.planet import Planet

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- radiobear_b200 has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # keep stdout to the one JSON line: NCCL (and anything else below us) may print there, so fd 1 points at
    # stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        # NCCL_DEBUG is left as the launcher set it: its lines go to stderr (fd 1 points there until the JSON line)
        dist.init_process_group('nccl', device_id=dev)
    ctx = _lib.get_context(local)
    ctx.enable_timing(True)
    precision = ctx.rt_precision()          # 'f64' unless RB_RT_PRECISION=mixed (include/radiobear_b200.h)

    atm, freqs, grid = workload()
    cfg = atm.config
    q = cfg.Rpol / cfg.Req
    L, F, n = atm.gas.shape[1], len(freqs), len(grid)
    S = L - 1
    mask = on_disc_mask(grid, q)
    # RB_BENCH_EMULATE_WORLD=N: time rank 0's share of an N-rank run on one GPU (development aid; the line says so)
    emulate = int(os.environ.get('RB_BENCH_EMULATE_WORLD', '0')) if world == 1 else 0
    rparts = parallel.partition_rows(grid, q, emulate or world)
    r0, r1 = rparts[rank]
    pts_all = np.stack([np.tile(grid, n), np.repeat(grid, n)], axis=1)
    parts = [(a * n, b * n) for a, b in rparts]
    pts = pts_all[parts[rank][0]:parts[rank][1]]

    # ---- device-resident inputs --------------------------------------------------------------
    t64 = dict(dtype=torch.float64, device=dev)
    freqs_t = torch.tensor(freqs, **t64)
    T_t = torch.tensor(atm.gas[cfg.C['T']], **t64)
    P_t = torch.tensor(atm.gas[cfg.C['P']], **t64)
    gas_t = torch.tensor(atm.gas, **t64).contiguous()
    radius_t = torch.tensor(atm.property[cfg.LP['R']], **t64)
    nidx = atm.property[cfg.LP['N']]
    b_t = torch.tensor(pts, **t64).contiguous()
    slab_t = torch.empty((L, F), **t64)
    tb_t = torch.empty((len(pts), F), dtype=torch.float32, device=dev)
    forms = [(c, f) for c, f in sorted(cfg.constituent_alpha.items()) if f is not None]
    other = {'h2': {'h2state': cfg.h2state}}
    orient = [float(cfg.orientation[0]), float(cfg.orientation[1])]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # N > 1: the absorption slab is computed once per step by all ranks together -- every rank its block of layers,
    # stored by the kernel into the [L][F] slab of every GPU over NVLink (symmetric memory), one device-side barrier --
    # instead of N times redundantly.  Two slabs alternate so that a rank that runs ahead never overwrites the slab a
    # slower rank still integrates with.  RB_BENCH_STEP_ALPHA=replicated keeps every rank computing all layers.
    syms = None
    if world > 1 and os.environ.get('RB_BENCH_STEP_ALPHA', 'sharded') == 'sharded':
        syms = [parallel.SymmetricSlab(L, F, dev) for _ in range(2)]
        ok = torch.tensor([1 if all(s.usable for s in syms) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            syms = None
    if syms is not None:
        lo, hi = parallel.partition_even(L, world)[rank]
        T_blk, P_blk, gas_blk = T_t[lo:hi].contiguous(), P_t[lo:hi].contiguous(), gas_t[:, lo:hi].contiguous()
    step_no = [0]

    def step():
        # geometry first, on the library's side stream: it does not depend on the absorption and overlaps it
        engine.geometry_prefetch_dev(radius_t, nidx[0], nidx[1], b_t, cfg.Req, cfg.Rpol, orient, cfg.gtype, cfg.limb, ctx=ctx)
        if syms is not None:
            sym = syms[step_no[0] & 1]
            step_no[0] += 1
            engine.alpha_layers_dev(freqs_t, T_blk, P_blk, gas_blk, cfg.C, formalisms=forms, other_dicts=other,
                                    truncate_strength=cfg.truncate_strength, freqs_host=freqs, ctx=ctx,
                                    scatter=(sym.ptrs, lo))
            sym.barrier()                   # every rank's layers are in every GPU's slab
            slab = sym.tensor
        else:
            slab = slab_t
            engine.alpha_layers_dev(freqs_t, T_t, P_t, gas_t, cfg.C, formalisms=forms, other_dicts=other,
                                    truncate_strength=cfg.truncate_strength, out=slab_t, freqs_host=freqs, ctx=ctx)
        engine.rt_batch_dev(radius_t, nidx[0], nidx[1], b_t, slab, T_t, cfg.Req, cfg.Rpol, orient, cfg.gtype, cfg.limb,
                            out_f32=True, tau_cut=engine.TAU_CUT, out=tb_t, ctx=ctx)
        return tb_t                         # N > 1: the image stays row-sharded in HBM

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # pixel census from a real result (hit = anything but the cosmic background)
    tb_host = tb_t.cpu().numpy()
    n_on_local = int(np.sum(~(tb_host[:, 0] == np.float32(2.725))))
    n_nan_local = int(np.sum(np.isnan(tb_host[:, 0])))
    counts = torch.tensor([n_on_local, n_nan_local, len(pts)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    n_on, n_nan, n_all = [int(x) for x in counts.tolist()]

    sharding_note = ('image rows balanced by on-disc pixels; output stays row-sharded in HBM; ' +
                     ('absorption: layer blocks, all_gather inside the kernel over NVLink (symmetric memory) + one device-side '
                      'barrier per step; ' if syms is not None else 'absorption replicated on every rank (no collective); ') +
                     'e2e: every rank copies its rows into one shared pinned host image')
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    kcount0 = {w: ctx.kernel_timed_count(w) for w in ('alpha', 'geometry', 'rt')}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (untimed)
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    ms_per_step = float(total.item()) / args.steps
    value = n_on * F / (ms_per_step * 1e-3)

    # per-kernel device times of the timed steps (library-side CUDA events on the same stream)
    def per_step_ms(which):
        # a step may launch a kernel family several times (ray chunks): sum per step, average over the steps
        n = ctx.kernel_timed_count(which) - kcount0[which]
        h = ctx.kernel_ms_history(which, min(n, 256))
        return float(np.sum(h)) * (n / max(len(h), 1)) / args.steps, n // args.steps

    k_alpha, _ = per_step_ms('alpha')
    k_geo, _ = per_step_ms('geometry')
    k_rt, rt_chunks = per_step_ms('rt')
    fp64_peak = ctx.fp64_peak_tflops(20000) if rank == 0 else 0.0
    # executed (ray, freq, segment) steps of one step of this rank (the tau_cut exit skips the rest): untimed pass
    ctx.count_steps(True)
    step()
    torch.cuda.synchronize()
    steps_executed = ctx.count_steps(False)
    steps_small = ctx.count_small_steps()

    # ---- the same step with the integration in mixed precision (rb_set_rt_precision; not the headline) ----------
    rt_mixed = None
    if precision == 'f64' and os.environ.get('RB_BENCH_SKIP_MIXED') is None:
        tb_ref = tb_t.clone()
        ctx.set_rt_precision('mixed')
        for _ in range(3):
            step()
        barrier()
        km0 = ctx.kernel_timed_count('rt')
        evm = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.zero_()
            evm[i][0].record()
            step()
            evm[i][1].record()
        barrier()
        tm = torch.tensor([sum(a.elapsed_time(b) for a, b in evm)], dtype=torch.float64, device=dev)
        ok = ~torch.isnan(tb_ref)
        same_nan = bool(torch.equal(torch.isnan(tb_t), torch.isnan(tb_ref)))
        dmax = torch.tensor([float((tb_t[ok].double() - tb_ref[ok].double()).abs().max()) if bool(ok.any()) else 0.0],
                            dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        hm = ctx.kernel_ms_history('rt', min(ctx.kernel_timed_count('rt') - km0, 256))
        ctx.set_rt_precision('f64')
        ms_mixed = float(tm.item()) / args.steps
        rt_mixed = {'precision': 'mixed: FP64 optical depth, SFU ex2, FP32 (FFMA2) weights and chunk sums, FP64 totals',
                    'ms_per_step': ms_mixed, 'value': n_on * F / (ms_mixed * 1e-3), 'unit': UNIT,
                    'rt_integrate_ms': float(np.mean(hm)) if len(hm) else None,
                    'max_abs_dTb_K_vs_f64': float(dmax.item()), 'nan_pattern_equal': same_nan,
                    'note': 'same step, same inputs, same run; float32 Tb outputs compared (1 ulp of Tb = 3e-5 K at 300-500 K); '
                            'parity bar 0.01 K; opt-in (RB_RT_PRECISION=mixed / rb_set_rt_precision), not the headline'}

    # ---- end to end through the public API (host buffers) ----------------------------------------
    planet = Planet('jupiter', atmosphere=atm, verbose=False)
    # bytes this rank moves per Planet.run: impact points, atmosphere, alpha slab in; its Tb rows + alpha slab out
    h2d = pts.nbytes + atm.gas.nbytes + atm.cloud.nbytes + 3 * L * 8 + L * F * 8 + F * 8
    d2h = len(pts) * F * 4 + L * F * 8
    io = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(io)
    h2d, d2h = [int(x) for x in io.tolist()]
    for _ in range(2):
        planet.run(list(freqs), b=BSTEP, reuse_override='false')
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        rv = planet.run(list(freqs), b=BSTEP, reuse_override='false')
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n_on * F / float(e2e_t.item())

    # secondary metric: config C5 absorption, frequency-sharded over the ranks (every rank takes part)
    a5_short = alpha_c5(ctx, dev, world=world, rank=rank)
    a5_full = alpha_c5(ctx, dev, full=True, world=world, rank=rank) if os.environ.get('RB_BENCH_SKIP_C5_FULL') is None else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
        rt_ms = k_rt
        n_on_rank = n_on_local
        # algorithmic bytes of one rt_integrate launch: ds slab read for the on-disc rays +
        # alpha slab + T + float32 Tb out for every pixel of the rank
        # (the mixed kernel reads the float copy of ds: 4 bytes per segment, and 32-byte operand rows per (layer, freq))
        rt_bytes = n_on_rank * S * (8 if precision == 'f64' else 4) + F * L * (8 if precision == 'f64' else 32) + L * 8 + len(pts) * F * 4
        rt_steps_all = float(n_on_rank) * F * (S - 1)
        rt_flops = float(steps_executed - steps_small) * RT_FLOPS_PER_STEP + float(steps_small) * RT_FLOPS_PER_SMALL_STEP
        rt_instr = float(steps_executed - steps_small) * RT_FP64_INSTR_PER_STEP + float(steps_small) * RT_FP64_INSTR_PER_SMALL_STEP
        steps_b = float(steps_executed - steps_small)
        sm_hz = 1e6 * float((clocks or {}).get('sm_mhz') or 1965.0)
        n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        sass_a, sass_b = RT_SASS_PER_STEP[precision]
        wf_a, wf_b = RT_SMEM_WAVEFRONTS_PER_STEP[precision]
        warp_cycles = n_sms * sm_hz * (rt_ms * 1e-3)          # SM-cycles of the launch
        hbm = {'achieved': rt_bytes / (rt_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
               'frac': rt_bytes / (rt_ms * 1e-3) / 1e9 / hbm_peak, 'algorithmic_bytes': int(rt_bytes), 'peak_source': peak_src,
               'note': 'SURVEY 8d bytes: ds of the on-disc rays + alpha slab + T + Tb out; at F = 64 every ds byte feeds 64 '
                       'frequencies, so HBM is 4 % busy and does not bind'}
        common = {'kernel': RT_KERNEL[precision],
                  'traffic': (RT_DRAM_BYTES_N1 if precision == 'f64' else RT_DRAM_BYTES_N1_MIXED) if world == 1 else None,
                  'traffic_source': RT_DRAM_SOURCE if precision == 'f64' else RT_DRAM_SOURCE_MIXED,
                  'ms_per_launch': rt_ms, 'launches_per_step': rt_chunks, 'hbm': hbm,
                  # the other two resources the kernel runs close to: warp-instruction issue slots (4 schedulers per SM,
                  # 1 instruction per clock each) and the shared-memory data pipe (1 wavefront = 128 B per clock per SM)
                  'issue': {'sass_per_step_phase_a': sass_a, 'sass_per_step_phase_b': sass_b,
                            'frac': (float(steps_small) * sass_a + steps_b * sass_b) / 32.0 / (4.0 * warp_cycles),
                            'note': 'hot-loop instructions only (static SASS count x executed steps)'},
                  'smem': {'wavefronts_per_step_phase_a': wf_a, 'wavefronts_per_step_phase_b': wf_b,
                           'frac': (float(steps_small) * wf_a + steps_b * wf_b) / 32.0 / warp_cycles,
                           'peak': '128 B/clk/SM (B300_MICROARCH.md, LDS/STS)'},
                  'sm_mhz_used': sm_hz / 1e6,
                  'segment_steps_executed': int(steps_executed), 'segment_steps_small_tau': int(steps_small),
                  'segment_steps_all': rt_steps_all}
        if precision == 'f64':
            # the binding resource of the dominant kernel at F = 64: the FP64 pipe (compute-bound for F >~ 4)
            tf = rt_flops / (rt_ms * 1e-3) / 1e12
            roofline = {'bound': 'fp64', 'achieved': tf, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                        'frac': tf / fp64_peak if fp64_peak else None,
                        'peak_source': 'rb_probe_fp64_peak: independent DFMA chains on every SM, same box, same run '
                                       '(MEASURED_PEAKS.json has no FP64 entry)',
                        'flops_per_segment_step': RT_FLOPS_PER_STEP, 'fp64_instr_per_segment_step': RT_FP64_INSTR_PER_STEP,
                        'flops_per_small_tau_step': RT_FLOPS_PER_SMALL_STEP,
                        'fp64_instr_per_small_tau_step': RT_FP64_INSTR_PER_SMALL_STEP,
                        'pipe_frac': (rt_instr / (rt_ms * 1e-3) / 1e12) / (fp64_peak / 2.0) if fp64_peak else None,
                        'note': 'flops counted over the segment-steps actually executed (in-kernel counter, untimed pass); the '
                                'tau >= tau_cut exit skips the rest of each ray; pipe_frac counts FP64 instructions (a DADD / '
                                'DMUL occupies the pipe like a DFMA)'}
            roofline.update(common)
        else:
            roofline = {'bound': 'hbm', 'achieved': hbm['achieved'], 'peak': hbm_peak, 'unit': 'GB/s', 'frac': hbm['frac'],
                        'peak_source': peak_src,
                        'mixed': {'fp64_instr_per_step_phase_b': 2, 'mufu_per_step_phase_b': 1,
                                  'note': 'optical depth in FP64 (2 DFMA per step in phase B, none in phase A), exp on the SFU, '
                                          'weights and chunk sums in FP32 (FFMA2 / FMUL2), chunk sums accumulated in FP64; '
                                          'issue-bound, see issue / smem'}}
            roofline.update(common)
        cores = 1
        cpu_v, cpu_desc, _ = cpu_sample(cores, n_pix_per_core=int(os.environ.get('RB_BENCH_CPU_PIXELS', '256')), n_lay_per_core=128) if world == 1 else (None, None, None)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f64' if precision == 'f64' else 'f64 optical depth + f32 weights (mixed)', 'data': 'synthetic image grid over the Jupiter default atmosphere fixture',
            'config': {'workload': WORKLOAD, 'pixels': 'on-disc', 'on_disc_pixels': n_on, 'nan_limb_pixels': n_nan,
                       'all_pixels': n_all, 'layers': L, 'freqs': F, 'sharding': sharding_note,
                       'l2': 'flushed between timed steps (256 MiB write, untimed); ds slab (0.94 GB) exceeds L2',
                       'tb_dtype_out': 'f32', 'tau_cut': engine.TAU_CUT, 'rt_precision': precision,
                       **({'emulated_rank0_share_of_world': emulate} if emulate else {})},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': 1e3 * float(e2e_t.item()), 'api': 'Planet.run(freqs, b=0.005)'},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'kernels_ms': {'alpha_lines': k_alpha, 'ray_geometry': k_geo, 'rt_integrate': rt_ms,
                           'note': 'device time of each kernel family per step ({} integrate launch(es)); ray_geometry runs on a side '
                                   'stream concurrently with alpha_lines, so the sum exceeds ms_per_step'.format(rt_chunks)},
            'value_all_pixels': n_all * F / (ms_per_step * 1e-3),
        }
        if cpu_v is not None:
            line['cpu_baseline'] = {'value': cpu_v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': cpu_desc,
                                    'stock_reference': stock_reference_record()}
        if rt_mixed is not None:
            line['rt_mixed'] = rt_mixed
        line['retrieval_loop'] = retrieval_loop(atm)
        for key, a5 in (('alpha_c5', a5_short), ('alpha_c5_full', a5_full)):
            if a5 is not None:
                a5['fp64_peak_tflops_per_gpu'] = fp64_peak
                a5['fp64_frac'] = a5['fp64_tflops_algorithmic'] / (fp64_peak * world) if fp64_peak else None
                line[key] = a5
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()      # before stdout is handed back: NCCL_DEBUG=INFO reports the teardown on fd 1
    if rank == 0:
        # the one JSON line goes to the real stdout; fd 1 stays pointed at stderr until the process ends (NCCL and its
        # plugins still report their shutdown there with NCCL_DEBUG=INFO)
        sys.stdout.flush()
        os.write(saved_stdout, (json.dumps(line) + '\n').encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
