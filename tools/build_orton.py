#!/usr/bin/env python
"""Repack the Orton H2 collision-induced-absorption tables the reference ships as text
(radiobear/constituents/h2/orton_H2.tables, read by h2_orton.py:17-123) into radiobear_b200/data/orton_h2.npz.
Values are copied unchanged (log of the absorption coefficients, wavenumbers in cm^-1); only the six tables
h2_orton.py uses (h2tableList, h2_orton.py:12) are kept.

    python tools/build_orton.py [/root/reference]
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NTAB = 6        # eh2h2, nh2h2, eh2he, nh2he, eh2ch4, nh2ch4


def main(ref='/root/reference'):
    fn = os.path.join(ref, 'radiobear', 'constituents', 'h2', 'orton_H2.tables')
    with open(fn) as fp:
        lines = fp.read().split('\n')
    head = lines[0].split()
    ntemp, tmax, tmin = int(head[0]), float(head[1]), float(head[2])
    nfreq = int(lines[1])
    nperl = len(lines[2].split())
    nfline = int(math.ceil(1.0 * nfreq / nperl))
    wn = np.array([float(v) for ln in lines[2:2 + nfline] for v in ln.split()])
    assert wn.shape == (nfreq,)
    first = 2 + nfline
    tabs = np.empty((NTAB, nfreq, ntemp))
    for ii in range(NTAB):
        for k in range(nfreq):
            tabs[ii, k] = [float(v) for v in lines[first + nfreq * ii + k].split()]
    out = os.path.join(ROOT, 'radiobear_b200', 'data', 'orton_h2.npz')
    np.savez_compressed(out, ntemp=ntemp, tmax=tmax, tmin=tmin, wavenumber=wn, logtab=tabs)
    print(out, os.path.getsize(out), 'bytes;', nfreq, 'wavenumbers', wn[0], '..', wn[-1], 'cm^-1;', ntemp, 'temperatures')


if __name__ == '__main__':
    main(*sys.argv[1:])
