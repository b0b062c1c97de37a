#!/usr/bin/env python
"""Repack the reference's line catalogs into one file: radiobear_b200/data/linecat.npz.

Input data (read-only, never modified): /root/reference/radiobear/constituents/
    nh3/ammonia.npz  (inversion / rotational / v2 lines used by nh3_hs, nh3_dbs, nh3_kd; nh3_hs.py:106-119)
    nh3/nh3.npz      (nh3_sjs.py:89-92)
    h2s/h2s.npz      (h2s_ddb.py:55-58)
    ph3/ph3jh.npz + ph3/PH3WGT.npz (ph3_jh.py:75-80)
    co/co.npz        (co_ddb.py:49-51)
The 15 H2O lines of h2o_bk.py:23-49 are constants of the formalism and live in the
oracle / kernel sources, not here.

Only numeric catalog DATA is repacked (float64 1-D arrays, unchanged values); no reference source
code is copied.  Run in the build container (the GPU box has no /root/reference):
    python tools/build_linecat.py
"""
import os
import sys
import numpy as np

REF = os.environ.get('RADIOBEAR_REFERENCE', '/root/reference')
SRC = os.path.join(REF, 'radiobear', 'constituents')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   'radiobear_b200', 'data', 'linecat.npz')

PACK = {
    'nh3_inv': ('nh3/ammonia.npz', ['fo', 'Io', 'Eo', 'gammaNH3o']),
    'nh3_rot': ('nh3/ammonia.npz', ['fo_rot', 'Io_rot', 'Eo_rot', 'gNH3_rot', 'gH2_rot', 'gHe_rot']),
    'nh3_v2': ('nh3/ammonia.npz', ['fo_v2', 'Io_v2', 'Eo_v2']),
    'nh3_sjs': ('nh3/nh3.npz', ['f0', 'I0', 'E', 'G0']),
    'h2s': ('h2s/h2s.npz', ['f0', 'I0', 'E', 'GH2S']),
    'ph3': ('ph3/ph3jh.npz', ['f0', 'I0', 'E']),
    'ph3_wgt': ('ph3/PH3WGT.npz', ['WgtI0', 'WgtFGB', 'WgtSB']),
    'co': ('co/co.npz', ['f0', 'I0', 'E']),
}


# The untrimmed NH3 line lists (ammonia_rotational.dat: 1301 lines, ammonia_rotovibrational.dat: 4198 lines; the
# shipped ammonia.npz keeps the first 201 / 198 of them, constituents/txt2npz.py:21-56 with rm = 1100, vm = 4000):
# BASELINE config C5 / SURVEY 8d "full catalog" variant, 415 + 1301 + 4198 = 5914 lines.
FULL = {
    'nh3_rot_full': ('nh3/ammonia_rotational.dat', 6),
    'nh3_v2_full': ('nh3/ammonia_rotovibrational.dat', 3),
}


def main():
    out = {}
    for name, (fn, ncols) in FULL.items():
        arr = np.loadtxt(os.path.join(SRC, fn), skiprows=1, unpack=True)
        assert arr.shape[0] == ncols, (fn, arr.shape)
        out[name] = np.ascontiguousarray(arr, dtype=np.float64)
        print('{:13s} {:34s} -> [{} x {}]'.format(name, fn, arr.shape[0], arr.shape[1]))
    for name, (fn, cols) in PACK.items():
        d = np.load(os.path.join(SRC, fn))
        arr = np.stack([np.asarray(d[c], dtype=np.float64) for c in cols])
        out[name] = arr
        out[name + '_cols'] = np.array(cols)
        print('{:8s} {:28s} -> [{} x {}]'.format(name, fn, arr.shape[0], arr.shape[1]))
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    sys.exit(main())
