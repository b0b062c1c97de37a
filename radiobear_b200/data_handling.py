"""DataReturn container (the object Planet.run returns).

Same attribute set and dtype conventions as the reference's Data (data_handling.py:11-49):
lists are stored as float32 arrays, a 'disc' request is stored as ['disc'].
"""
import copy

import numpy as np

from . import utils


class Data:
    allowed_parameters = ['f', 'freqUnit', 'b', 'Tb', 'header', 'start', 'stop', 'log', 'type', 'logfile']

    def __init__(self):
        for name in self.allowed_parameters:
            setattr(self, name, None)

    def set(self, par, val):
        if par not in self.allowed_parameters:
            print("{} not in valid data return list.".format(par))
            return
        if not isinstance(val, np.ndarray):      # arrays are adopted as-is (a 92 MB image cube is not copied)
            val = copy.copy(val)
        if par == 'b' and not isinstance(val, np.ndarray) and utils.b_type(val).startswith('dis'):
            self.b = ['disc']
        elif isinstance(val, list):
            setattr(self, par, np.asarray(val, dtype=np.float32))
        elif par == 'Tb' and isinstance(val, np.ndarray):
            setattr(self, par, val.astype(np.float32, copy=False))
        elif par == 'b' and isinstance(val, np.ndarray):
            setattr(self, par, val)          # image grids are shared read-only arrays (no per-run conversion)
        else:
            setattr(self, par, val)

    def __repr__(self):
        out = []
        for i, b in enumerate(self.b):
            out.append('b = {}:'.format(b))
            out.append(' f = ' + ' '.join('{:6.1f}'.format(x) for x in self.f) + '  GHz')
            out.append('Tb = ' + ' '.join('{:6.1f}'.format(x) for x in np.atleast_1d(self.Tb[i])) + '  K')
        return '\n'.join(out) + '\n'

    def show(self, include=('header', 'start', 'stop', 'f', 'b', 'Tb'), indent=1):
        tab = indent * '\t'
        for v in include:
            if v not in self.allowed_parameters:
                continue
            if v == 'header':
                print('{}<<<Header>>>'.format(tab))
                for k, h in (self.header or {}).items():
                    print('{}\t{:20s}     {}'.format(tab, k, h))
            elif v == 'f':
                print('{}<<<freq>>> {} {}'.format(tab, self.f, self.freqUnit))
            else:
                print('{}<<<{}>>>  {}'.format(tab, v, getattr(self, v)))
