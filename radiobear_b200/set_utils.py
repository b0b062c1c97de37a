"""Frequency / impact-parameter request parsing.

Same request grammar as the reference (set_utils.py:12-88 set_b, :91-129 set_freq, :132-146
proc_string).  Two reference crashes are accepted here instead of raised:
a float `b` (full image grid; set_utils.py:78 calls len() on the float) and the log-sweep string
'start;stop;n' (set_utils.py:141 passes a float count to np.logspace).
"""
from argparse import Namespace

import numpy as np

from . import utils


_IMAGE_B_CACHE = {}


def image_grid(bstep):
    """Pixel coordinates of a full-disc image: -1.5..1.5 in steps of bstep, symmetric about 0
    (set_utils.py:65-66)."""
    pos = np.arange(bstep, 1.5 + bstep, bstep)
    return np.concatenate((-1.0 * pos[::-1], np.arange(0.0, 1.5 + bstep, bstep)))


def proc_string(srq):
    """'a,b,c' list | 'start:stop:step' arange | 'start;stop;n' logspace | number | file name."""
    if ',' in srq:
        return [float(x) for x in srq.split(',')]
    if ':' in srq:
        start, stop, step = [float(x) for x in srq.split(':')]
        return list(np.arange(start, stop + step / 2.0, step))
    if ';' in srq:
        start, stop, num = [float(x) for x in srq.split(';')]
        return list(np.logspace(np.log10(start), np.log10(stop), int(round(num))))
    try:
        return [float(srq)]
    except ValueError:
        return list(np.loadtxt(srq))


def set_freq(freqs, freqUnit='GHz'):
    """Return (list of frequencies in GHz, freqUnit)."""
    if isinstance(freqs, list):
        freqs = list(freqs)
    elif isinstance(freqs, np.ndarray):
        freqs = list(freqs)
    elif isinstance(freqs, str):
        freqs = proc_string(freqs)
    elif utils.isanynum(freqs):
        freqs = [float(freqs)]
    else:
        raise ValueError('Invalid format for frequency request')
    proc = utils.proc_unit(freqUnit)
    if freqUnit is not None and proc is not None and len(freqs) > 8:
        try:       # utils.convert_unit on the whole list: the same two IEEE operations per element
            arr = np.asarray(freqs, dtype=np.float64)
            if arr.ndim == 1:
                return (arr * utils.Units[freqUnit] / utils.Units[proc]).tolist(), freqUnit
        except (TypeError, ValueError):
            pass
    return [utils.convert_unit(f, freqUnit) for f in freqs], freqUnit


def set_b(b, block=(1, 1), **kwargs):
    """Process a b request -> Namespace(b, block, data_type, imSize).

    b: list of pairs | one pair | float (full image grid at that spacing) | 'disc' |
       'stamp:bres:xmin,xmax,ymin,ymax' | 'start:stop:step[<angle]' | 'n1,n2,..[<angle]'
    """
    block = list(block)
    rv = Namespace(b=None, data_type=None, block=block, imSize=None)
    if isinstance(b, str):
        req = b.lower()
        if req.startswith('dis'):
            rv.b, rv.data_type = [req], 'spectrum'
        elif req.startswith('stamp'):
            parts = req.split(':')
            bres = float(parts[1])
            xmin, xmax, ymin, ymax = [float(x) for x in parts[2].split(',')]
            outer = np.arange(xmin, xmax + bres / 2.0, bres)
            inner = np.arange(ymin, ymax + bres / 2.0, bres)
            rv.b = [[y, x] for x in outer for y in inner]          # set_utils.py:46-49
            rv.imSize = [len(inner), len(outer)]
            rv.data_type = 'image'
        else:
            mag, _, ang = req.partition('<')
            angle = utils.d2r(float(ang)) if ang else 0.0
            ab = kwargs['Rpol'] / kwargs['Req']
            rab = ab / np.sqrt(np.sin(angle)**2 + (ab * np.cos(angle))**2)
            rv.b = [[v * np.cos(angle), v * np.sin(angle)] for v in proc_string(mag) if v < 0.995 * rab]
            rv.data_type = 'profile'
        return rv
    if isinstance(b, float):
        hit = _IMAGE_B_CACHE.get((b, block[0], block[1]))
        if hit is not None:
            rv.b, rv.imSize, rv.data_type = hit[0], list(hit[1]), 'image'
            return rv
        grid = image_grid(b)
        nblk = abs(block[1])
        bsplit = len(grid) / nblk
        last = 0 if nblk == 1 else block[0] / nblk
        first = int((block[0] - 1) * bsplit)
        rows = [first + i for i in range(int(bsplit + last)) if first + i < len(grid)]
        # [R][2] array (x fastest, rows of constant y) -- indexable like the reference's list of pairs.
        # Pure function of (b, block): built once, kept in page-locked memory, handed out read-only.
        from . import hostmem
        pts = hostmem.pinned_copy(np.stack([np.tile(grid, len(rows)), np.repeat(grid[rows], len(grid))], axis=1))
        pts.flags.writeable = False
        if len(_IMAGE_B_CACHE) >= 4:
            _IMAGE_B_CACHE.pop(next(iter(_IMAGE_B_CACHE)))
        _IMAGE_B_CACHE[(b, block[0], block[1])] = (pts, (len(grid), len(rows)))
        rv.b = pts
        rv.imSize = [len(grid), len(rows)]
        rv.data_type = 'image'
        return rv
    shape_b = np.shape(b)
    if len(shape_b) == 1:
        rv.data_type, rv.b = 'spectrum', [list(b)]
    else:
        rv.data_type = 'spectrum' if shape_b[0] < 5 else 'profile'
        rv.b = b
    return rv
