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

    def show(self, include=('header', 'start', 'stop', 'f', 'b', 'Tb', 'log'), indent=1):
        tab = indent * '\t'
        for v in include:
            if v == 'log':
                self.show_log(indent=indent)
            elif v not in self.allowed_parameters:
                continue
            elif v == 'header':
                self.show_header(indent=indent)
            elif v == 'f':
                print('{}<<<freq>>> {} {}'.format(tab, self.f, self.freqUnit))
            else:
                print('{}<<<{}>>>  {}'.format(tab, v, getattr(self, v)))

    def show_header(self, indent=1):
        """data_handling.py:67-74."""
        print('{}<<<Header>>>'.format(indent * '\t'))
        for k, h in self.header.items():
            print('{}{:20s}     {}'.format((indent + 1) * '\t', k, h))
        print()

    def show_log(self, indent=1):
        """Close the run's log and print the file (data_handling.py:76-85)."""
        if self.log is not None:
            self.log.close()
        print('{}<<<Log>>>  {}'.format(indent * '\t', self.logfile))
        with open(self.logfile) as fp:
            for line in fp:
                print(line.strip())
        print("-----------------------------------\n")
