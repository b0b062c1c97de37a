"""Multi-GPU sharding: one process per GPU, torch.distributed (NCCL over NVLink; gloo in CPU tests).

The hot paths shard without any data-path collective:
* image / profile runs: contiguous blocks of image ROWS (or b-points) per rank, balanced by the number
  of on-disc pixels (off-disc pixels cost nothing); every rank holds the full alpha slab (0.5 MB at C4,
  recomputed locally in ~0.1 ms rather than broadcast); one final gather of Tb to rank 0.
* large absorption requests (alpha-dominated sweeps): contiguous blocks of LAYERS per rank (or of frequencies,
  `Alpha.shard_axis = 'freqs'`); one all_gather of the blocks when every rank needs the full slab.
This replaces the reference's manual `block=[i, N]` row chunking (set_utils.py:66-76,
scripts/image_block_pipeline.py).
"""
import numpy as np


def row_weights(grid, q, floor=1.0):
    """Work estimate per image row: on-disc pixel count (+ a small floor for the off-disc scan)."""
    grid = np.asarray(grid, dtype=np.float64)
    half = 1.0 - (grid / q)**2                      # x^2 < 1 - (y/q)^2
    w = np.zeros(len(grid))
    ok = half > 0.0
    step = np.abs(grid[1] - grid[0]) if len(grid) > 1 else 1.0
    w[ok] = 2.0 * np.sqrt(half[ok]) / step
    return w + floor


def partition_rows(grid, q, n, floor=1.0):
    """n contiguous [start, stop) row blocks with (nearly) equal summed weight."""
    w = row_weights(grid, q, floor)
    cum = np.concatenate(([0.0], np.cumsum(w)))
    total = cum[-1]
    cuts = [0]
    for i in range(1, n):
        target = total * i / n
        j = int(np.searchsorted(cum, target))
        j = min(max(j, cuts[-1]), len(w))
        cuts.append(j)
    cuts.append(len(w))
    return [(cuts[i], cuts[i + 1]) for i in range(n)]


def partition_even(count, n):
    """n contiguous [start, stop) blocks of `count` items (frequency blocks / b-lists)."""
    base, extra = divmod(count, n)
    out, s = [], 0
    for i in range(n):
        e = s + base + (1 if i < extra else 0)
        out.append((s, e))
        s = e
    return out


def gather_blocks(local, parts, dst=0, group=None):
    """Gather per-rank row blocks (torch tensors [rows_i, ...]) to rank `dst` -> [sum rows, ...] or None.

    Blocks are padded to the largest block so one dist.gather moves everything (NVLink: 92 MB for the
    full C4 cube, ~0.15 ms at 700 GB/s)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = [e - s for s, e in parts]
    if world == 1:
        return local
    mx = max(rows)
    pad = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[i][:rows[i]] for i in range(world)], dim=0)


def all_gather_freq_blocks(local_slab, parts, group=None):
    """All-gather frequency-sharded alpha slabs [L][F_i] -> full [L][F] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local_slab
    widths = [e - s for s, e in parts]
    mx = max(widths)
    L = local_slab.shape[0]
    pad = torch.zeros((L, mx), dtype=local_slab.dtype, device=local_slab.device)
    pad[:, :local_slab.shape[1]] = local_slab
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[i][:, :widths[i]] for i in range(world)], dim=1).contiguous()


# Below this many (layer, frequency) pairs the absorption of a request is recomputed on every rank: at C4
# (1000 x 64) the kernel takes 0.1 ms, less than one collective; at C5 (4096 x 4096) it takes 7.5 ms on one GPU.
ALPHA_SHARD_MIN_PAIRS = 1 << 20


def shard_alpha(L, F, world, policy='auto'):
    """Should `Alpha.get_layers` split this request over the ranks?  policy: 'auto' (by size), True / False
    (RB_ALPHA_SHARD=1 / 0 overrides 'auto')."""
    import os
    if world <= 1 or min(L, F) < world:
        return False
    env = os.environ.get('RB_ALPHA_SHARD')
    if policy == 'auto' and env is not None:
        policy = env not in ('0', 'false', 'False', '')
    if policy == 'auto':
        return L * F >= ALPHA_SHARD_MIN_PAIRS
    return bool(policy)


def all_gather_layer_blocks(local_slab, parts, group=None):
    """All-gather layer-sharded alpha slabs [L_i][F] -> full [L][F] on every rank.  The blocks are contiguous pieces of
    the [L][F] slab: with equal block sizes the collective writes straight into it (no padding, no transposing copy)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local_slab
    heights = [e - s for s, e in parts]
    F = local_slab.shape[1]
    if len(set(heights)) == 1:
        full = torch.empty((sum(heights), F), dtype=local_slab.dtype, device=local_slab.device)
        dist.all_gather_into_tensor(full, local_slab.contiguous(), group=group)
        return full
    mx = max(heights)
    pad = torch.zeros((mx, F), dtype=local_slab.dtype, device=local_slab.device)
    pad[:local_slab.shape[0]] = local_slab
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[i][:heights[i]] for i in range(world)], dim=0).contiguous()


def alpha_layers_sharded(compute_block, L, F, axis='layers', group=None):
    """Absorption slab [L][F] computed by all ranks: rank r computes the contiguous block `partition_even(n, world)[r]`
    of the layers (axis='layers', n = L) or of the frequencies (axis='freqs', n = F) with `compute_block(lo, hi)`
    (-> [hi - lo][F] / [L][hi - lo], numpy array or torch tensor) and one all_gather gives every rank the full slab --
    every rank traces rays afterwards (the reference's only counterpart is running blocks of the
    request by hand, set_utils.py:66-76).  Layers are the default: the kernel builds a layer's line tables once per
    CTA, so frequency blocks repeat that work on every rank (measured at C5, N = 2: 3.99 ms per rank against 3.65 for
    half of the one-GPU time) while layer blocks do not, and a block of layers is a contiguous piece of the slab.
    Returns a numpy [L][F] array on every rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    by_layers = axis == 'layers'
    parts = partition_even(L if by_layers else F, world)
    lo, hi = parts[rank]
    local = compute_block(lo, hi)
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    want = (hi - lo, F) if by_layers else (L, hi - lo)
    if tuple(t.shape) != want:
        raise ValueError('compute_block({}, {}) returned {}, expected {}'.format(lo, hi, tuple(t.shape), want))
    if dist.get_backend(group) == 'nccl' and not t.is_cuda:
        t = t.cuda()
    full = (all_gather_layer_blocks if by_layers else all_gather_freq_blocks)(t.contiguous(), parts, group=group)
    return full.cpu().numpy() if full.is_cuda else full.numpy()


class SymmetricSlab:
    """A [L][F] float64 slab that exists at the same place on every GPU of the process group and that every rank can
    store into directly over NVLink (torch.distributed._symmetric_memory): the target of the in-kernel all_gather of
    layer-sharded absorption runs (`engine.alpha_layers_dev(..., scatter=...)`).  `usable` is False when the symmetric
    allocation or the rendezvous is not available (gloo, no peer access); callers then fall back to all_gather."""

    def __init__(self, L, F, device, group=None):
        import torch
        import torch.distributed as dist
        self.usable, self.tensor, self.handle, self.ptrs = False, None, None, None
        self.shape = (int(L), int(F))
        try:
            import torch.distributed._symmetric_memory as symm
            if dist.get_backend(group) != 'nccl':
                return
            grp = group if group is not None else dist.group.WORLD
            self.tensor = symm.empty((int(L), int(F)), dtype=torch.float64, device=device)
            self.handle = symm.rendezvous(self.tensor, grp)
            self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
            self.usable = len(self.ptrs) == dist.get_world_size(group) and all(self.ptrs)
        except Exception as e:                                  # noqa: BLE001 -- any failure means "use the collective"
            self.error = repr(e)
            self.usable = False

    def barrier(self):
        """All ranks' stores into every copy of the slab are complete and visible (device-side barrier with system-scope
        release / acquire on the current stream)."""
        self.handle.barrier(channel=0)


_SLABS = {}


def symmetric_slab(L, F, device, group=None):
    key = (int(L), int(F), str(device))
    if key not in _SLABS:
        _SLABS[key] = SymmetricSlab(L, F, device, group)
    return _SLABS[key]


def alpha_layers_scatter(freqs, T, P, gas, gas_dict, cloud, cloud_dict, formalisms, other_dicts, units, scale,
                         truncate_strength, truncate_freq, group=None):
    """Layer-sharded absorption with the all_gather inside the kernel: every rank computes its block of layers from
    device-resident inputs and stores the values into the full [L][F] slab of every GPU over NVLink
    (`SymmetricSlab`, `rb_alpha_layers_dev_scatter`); one device-side barrier, then the slab is copied to the host.
    Returns the numpy [L][F] slab, or None when the ranks cannot set up the symmetric slab (the caller then uses
    alpha_layers_sharded)."""
    import torch
    import torch.distributed as dist
    from . import engine
    if dist.get_backend(group) != 'nccl' or not torch.cuda.is_available():
        return None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device('cuda', torch.cuda.current_device())
    L, F = len(T), len(freqs)
    sym = symmetric_slab(L, F, dev, group)
    ok = torch.tensor([1 if sym.usable else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        return None
    lo, hi = partition_even(L, world)[rank]
    t64 = dict(dtype=torch.float64, device=dev)
    if hi > lo:
        fr = np.ascontiguousarray(freqs, dtype=np.float64)
        sm = engine.scale_matrix(slice_scale(scale, lo, hi), [c for c, _ in formalisms], hi - lo)
        engine.alpha_layers_dev(
            torch.tensor(fr, **t64), torch.tensor(np.ascontiguousarray(T[lo:hi]), **t64),
            torch.tensor(np.ascontiguousarray(P[lo:hi]), **t64), torch.tensor(np.ascontiguousarray(gas[:, lo:hi]), **t64),
            gas_dict, None if cloud is None else torch.tensor(np.ascontiguousarray(cloud[:, lo:hi]), **t64), cloud_dict,
            formalisms=formalisms, other_dicts=other_dicts, units=units,
            scale_t=None if sm is None else torch.tensor(np.ascontiguousarray(sm), **t64),
            truncate_strength=truncate_strength, truncate_freq=truncate_freq, freqs_host=fr, scatter=(sym.ptrs, lo))
    sym.barrier()                       # every rank's rows are in every copy
    out = sym.tensor.cpu().numpy()
    sym.barrier()                       # nobody overwrites a copy that is still being read
    return out


def slice_scale(scale, lo, hi):
    """The part of a `scale` request (number / per-layer list / dict of per-layer lists, alpha.py:235-259) that belongs
    to layers [lo, hi)."""
    if isinstance(scale, dict):
        return {k: list(v)[lo:hi] for k, v in scale.items()}
    if isinstance(scale, (list, tuple, np.ndarray)):
        return list(scale)[lo:hi]
    return scale


def world_rank():
    """(world_size, rank) of the default process group; (1, 0) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except ImportError:
        pass
    return 1, 0


class _ShmFile:
    """A file under /dev/shm mapped read-write (plain POSIX shared memory; no multiprocessing resource
    tracker, whose bookkeeping assumes one creator *and* one owner process tree)."""
    DIR = '/dev/shm'

    def __init__(self, name, size, create):
        import mmap
        import os
        self.path = os.path.join(self.DIR, name)
        self.created = create
        flags = os.O_RDWR | ((os.O_CREAT | os.O_EXCL) if create else 0)
        fd = os.open(self.path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, size)
            self.mm = mmap.mmap(fd, size)
        finally:
            os.close(fd)
        self.buf = memoryview(self.mm)

    def close(self):
        import os
        try:
            self.buf.release()
            self.mm.close()
        except (BufferError, ValueError):
            pass                    # a caller still holds an array on it; the mapping goes with the process
        if self.created:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class SharedHostExchange:
    """Single-node delivery of row-sharded results to rank 0 without a device-side gather.

    Every rank copies its own block device -> host over its own PCIe link, straight into a POSIX
    shared-memory segment that all ranks have mapped (and page-locked with cudaHostRegister when the data
    is on a GPU); rank 0 hands a zero-copy numpy view of the segment to the caller.  Compared with
    gather-to-rank-0 + one D2H copy this removes the NVLink gather and spreads the 92 MB (C4) copy over
    N PCIe links.  Control words live in a small shared segment: rank 0 publishes (seq, segment id) for
    call number seq, every rank stores seq into its own `done` slot when its copy has landed, rank 0 waits
    for all slots.  Segments are recycled once nothing references the memory handed out earlier: every result is
    a view of a fresh "owner" array over the segment (numpy points the .base of every derived view -- reshape,
    slices, astype(copy=False) -- at that owner), and rank 0 keeps only a weak reference to the owner, so the
    segment is free exactly when the owner has been collected.
    """
    _CTRL_WORDS = 8 + 256

    def __init__(self, group=None):
        import os
        import uuid
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.seq = 0
        self.segments = {}          # seg_id -> [shared file, uint8 ndarray over it, pinned?]
        self.handed_out = {}        # rank 0: seg_id -> (weakref to the owner array of the result handed out, nbytes)
        self._cur = None
        self.base = None
        self.ctrl_shm = None
        msg = [None]
        if self.rank == 0:
            local_world = int(os.environ.get('LOCAL_WORLD_SIZE', self.world))
            if local_world == self.world and not os.environ.get('RB_NO_SHM') and self.world <= 256:
                try:
                    base = 'rb200_{}_{}'.format(os.getpid(), uuid.uuid4().hex[:8])
                    self.ctrl_shm = _ShmFile(base + '_ctrl', 8 * self._CTRL_WORDS, create=True)
                    msg[0] = base
                except OSError:
                    msg[0] = None
        dist.broadcast_object_list(msg, src=0, group=group)
        self.base = msg[0]
        if self.base is None:
            return
        if self.rank != 0:
            self.ctrl_shm = _ShmFile(self.base + '_ctrl', 8 * self._CTRL_WORDS, create=False)
        self.ctrl = np.ndarray((self._CTRL_WORDS,), dtype=np.int64, buffer=self.ctrl_shm.buf)
        if self.rank == 0:
            self.ctrl[:] = 0
        dist.barrier(group=group)
        import atexit
        atexit.register(self.close)

    @property
    def usable(self):
        return self.base is not None

    def _segment(self, seg_id, nbytes, create, pin):
        if seg_id not in self.segments:
            shm = _ShmFile('{}_{}_{}'.format(self.base, seg_id, nbytes), max(nbytes, 8), create)
            arr = np.ndarray((nbytes,), dtype=np.uint8, buffer=shm.buf)
            self.segments[seg_id] = [shm, arr, False]
        ent = self.segments[seg_id]
        if pin and not ent[2]:
            import torch
            err = torch.cuda.cudart().cudaHostRegister(ent[1].ctypes.data, max(nbytes, 8), 0)
            if int(err) != 0:
                raise RuntimeError('cudaHostRegister of the shared result segment failed: {}'.format(err))
            ent[2] = True
        return ent[1]

    @staticmethod
    def _spin(cond, what):
        import time
        t0 = time.perf_counter()
        n = 0
        while not cond():
            n += 1
            if n > 2000:
                time.sleep(20e-6)
                if time.perf_counter() - t0 > 120.0:
                    raise TimeoutError('SharedHostExchange: timed out waiting for ' + what)

    def begin(self, total_rows, tail_shape, dtype, pin):
        """Collective: agree on the result segment of this call; returns it as [total_rows, *tail_shape]
        (every rank sees the same memory; write only your own rows), page-locked for CUDA when `pin`."""
        self.seq += 1
        seq = self.seq
        dtype = np.dtype(dtype)
        nbytes = int(total_rows) * int(np.prod(tail_shape, dtype=np.int64)) * dtype.itemsize
        ctrl = self.ctrl
        if self.rank == 0:
            seg_id = None
            for sid, (ref, size) in self.handed_out.items():
                if size == nbytes and ref() is None:    # the owner array (hence every view of it) is gone
                    seg_id = sid
                    break
            if seg_id is None:
                seg_id = len(self.handed_out)
            seg = self._segment(seg_id, nbytes, create=seg_id not in self.segments, pin=pin)
            ctrl[1] = seg_id
            ctrl[2] = nbytes
            ctrl[0] = seq                               # publish (x86 stores are ordered)
        else:
            self._spin(lambda: ctrl[0] >= seq, 'rank 0 to publish the result segment')
            seg_id = int(ctrl[1])
            seg = self._segment(seg_id, nbytes, create=False, pin=pin)
        # a fresh owner per call: an ndarray straight over the shared file's buffer is the array numpy collapses
        # the .base of every derived view to
        owner = np.ndarray((max(nbytes, 0),), dtype=np.uint8, buffer=self.segments[seg_id][0].buf)
        self._cur = (seg_id, owner[:nbytes].view(dtype).reshape((int(total_rows),) + tuple(tail_shape)), owner)
        return self._cur[1]

    def finish(self):
        """This rank's rows are in place (all copies complete).  Rank 0 waits for everybody and returns the
        assembled array; the other ranks return None at once."""
        seq = self.seq
        ctrl = self.ctrl
        import weakref
        seg_id, root, owner = self._cur
        self._cur = None
        ctrl[8 + self.rank] = seq
        if self.rank != 0:
            return None
        self._spin(lambda: all(ctrl[8 + r] >= seq for r in range(self.world)), 'the other ranks to copy their blocks')
        self.handed_out[seg_id] = (weakref.ref(owner), owner.nbytes)
        return root

    def deliver(self, local, parts, tail_shape, dtype):
        """local: this rank's block (torch tensor [rows_i, *tail_shape], CUDA or CPU); parts: [start, stop)
        row blocks of all ranks.  Returns the assembled [sum rows, *tail_shape] array on rank 0, None elsewhere."""
        import torch
        on_gpu = bool(local.is_cuda)
        full = self.begin(parts[-1][1], tail_shape, dtype, pin=on_gpu)
        s, e = parts[self.rank]
        if e > s:
            torch.from_numpy(full[s:e]).copy_(local, non_blocking=on_gpu)
            if on_gpu:
                torch.cuda.current_stream(local.device).synchronize()
        return self.finish()

    def close(self):
        for shm, arr, pinned in list(self.segments.values()):
            if pinned:
                try:
                    import torch
                    torch.cuda.cudart().cudaHostUnregister(arr.ctypes.data)
                except Exception:
                    pass
        self.handed_out.clear()
        segs = list(self.segments.values())
        self.segments = {}
        for ent in segs:
            shm = ent[0]
            ent[1] = None
            shm.close()
        if self.ctrl_shm is not None:
            self.ctrl = None
            self.ctrl_shm.close()
            self.ctrl_shm = None


def nidx_of(atm):
    """(n at the top layer, n at the second layer): what the first-interface Snell step needs."""
    n = atm.property[atm.config.LP['N']]
    return [float(n[0]), float(n[1])]


_EXCHANGE = None


def host_exchange():
    """The process group's SharedHostExchange (created collectively on first use)."""
    global _EXCHANGE
    if _EXCHANGE is None:
        _EXCHANGE = SharedHostExchange()
    return _EXCHANGE


_ROW_BLOCKS = {}


# Host-output runs move every pixel of a row to the host, sky included, while only the on-disc pixels cost device time:
# a row weighs its on-disc pixels plus this fraction of all its pixels (measured on 8 B200 behind shared PCIe uplinks:
# the ranks holding the sky rows at the top and bottom of a disc image were copying 2.4 times the bytes of the others).
ROW_COPY_WEIGHT = 0.3


def point_blocks(npts, cfg, rows, world, host_output=True):
    """[start, stop) blocks of the flat point list for every rank: whole image rows balanced by on-disc
    pixels (plus, for results that go to the host, the pixels to copy) when `rows` = (grid of row y values, pixels
    per row) is given, else an even split.
    (A pure function of its arguments, asked for twice per Planet.run: the last few answers are kept.)"""
    if rows is not None:
        import os
        grid, ncol = rows
        q = cfg.Rpol / cfg.Req if cfg.gtype == 'ellipse' else 1.0
        floor = 1.0 + (float(os.environ.get('RB_ROW_COPY_WEIGHT', ROW_COPY_WEIGHT)) * ncol if host_output else 0.0)
        key = (np.asarray(grid, dtype=np.float64).tobytes(), int(ncol), float(q), int(world), floor)
        blocks = _ROW_BLOCKS.get(key)
        if blocks is None:
            blocks = [(a * ncol, b * ncol) for a, b in partition_rows(grid, q, world, floor)]
            if len(_ROW_BLOCKS) >= 8:
                _ROW_BLOCKS.pop(next(iter(_ROW_BLOCKS)))
            _ROW_BLOCKS[key] = blocks
        return list(blocks)
    return partition_even(npts, world)


def local_block(npts, cfg, rows, world, rank):
    return point_blocks(npts, cfg, rows, world)[rank]


def run_points_sharded(planet, pts, atm, alpha, out_f32=True, rows=None):
    """Brightness temperatures of impact points pts[R][2] computed by all ranks of the process group.

    Each rank takes a contiguous block (image rows balanced by on-disc pixels when `rows` = (grid, ncol)
    is given, else an even split), runs geometry + integration on its own GPU from device-resident
    inputs and copies its block into the shared host segment of SharedHostExchange (one node), or, across
    nodes, one gather over NCCL brings the blocks to rank 0.  Returns Tb[R][F] on rank 0 and None on the
    other ranks."""
    import torch
    import torch.distributed as dist
    from . import engine
    world, rank = world_rank()
    dev = torch.device('cuda', torch.cuda.current_device())
    cfg = atm.config
    parts = point_blocks(len(pts), cfg, rows, world)
    s, e = parts[rank]
    F = alpha.n_freqs
    ex = host_exchange()
    if ex.usable:
        # one node: every rank integrates its rows with the chunked host-output pipeline (D2H of ray chunk c
        # overlaps the integration of chunk c+1, over this rank's own PCIe link) straight into the shared,
        # page-locked result image
        full = ex.begin(parts[-1][1], (F,), np.float32 if out_f32 else np.float64, pin=True)
        if e > s:
            from . import raypath
            engine.rt_batch(b=np.asarray(pts[s:e], dtype=np.float64), alpha_slab=alpha.rt_slab(), T=atm.gas[cfg.C['T']],
                            out_f32=out_f32, out=full[s:e], **raypath._geometry_args(atm, cfg.orientation, None))
        return ex.finish()
    radius = torch.as_tensor(np.ascontiguousarray(atm.property[cfg.LP['R']]), device=dev)
    nidx = atm.property[cfg.LP['N']]
    T_t = torch.as_tensor(np.ascontiguousarray(atm.gas[cfg.C['T']]), device=dev)
    slab_t = torch.as_tensor(np.ascontiguousarray(alpha.slab), device=dev)
    if e > s:
        b_t = torch.as_tensor(np.array(pts[s:e], dtype=np.float64), device=dev)      # private writable copy
        from . import raypath
        local = engine.rt_batch_dev(radius, nidx[0], nidx[1], b_t, slab_t, T_t, cfg.Req, cfg.Rpol,
                                    [float(cfg.orientation[0]), float(cfg.orientation[1])], cfg.gtype,
                                    getattr(cfg, 'limb', 'shape'), out_f32=out_f32,
                                    gravity_model=raypath._geometry_args(atm, cfg.orientation, None).get('gravity_model'))
    else:
        local = torch.empty((0, F), dtype=torch.float32 if out_f32 else torch.float64, device=dev)
    # several nodes / no shared memory: gather over NCCL, rank 0 copies to pinned host memory
    full = gather_blocks(local, parts, dst=0)
    if rank != 0:
        return None
    out = engine.pinned_pool.get(tuple(full.shape), np.float32 if out_f32 else np.float64)
    host = torch.from_numpy(out)
    host.copy_(full, non_blocking=False)
    return out
