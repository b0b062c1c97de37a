"""Text output of a DataReturn in the reference's file formats.

Same layouts as the reference writer (fileIO.py:19-107) so that existing readers of
`Output/<Planet>_<type>_<timestamp>.dat` keep working on GPU-produced results:
  header    the header strings sorted by key, one per line;
  spectrum  `# GHz  K@b  <b columns>` then one row per frequency: freq (or wavelength in cm) + Tb per b;
  profile   `# b  K@GHz  <freq columns>` then one row per b: `bx by` + Tb per frequency;
  image     one row of tab-separated Tb per image row (single frequency; a multi-frequency cube writes
            one such block per frequency, separated by `# f = <freq>` lines -- the reference cannot produce
            those, planet_base.py:201-202).
"""
import os

import numpy as np

from . import utils

SPEED_OF_LIGHT = 2.9979E8     # m/s (utils.py:74)


class FileIO(object):
    def __init__(self, directory='Output', scratch_spec=None):
        self.directory = directory
        self.scratch_spec = scratch_spec

    def write(self, output_file, data, xaxis='frequency'):
        d = os.path.dirname(output_file)
        if d and not os.path.isdir(d):
            os.makedirs(d)
        kind = str(data.type).lower()
        with open(output_file, 'w') as fp:
            for key in sorted((data.header or {}).keys()):
                fp.write(data.header[key] + '\n')
            if kind == 'image':
                self._image(fp, data)
            elif kind == 'spectrum':
                self._spectrum(fp, data, xaxis.lower())
            elif kind == 'profile':
                self._profile(fp, data, xaxis.lower())
            else:
                print("Invalid output type: {}".format(data.type))
        return output_file

    @staticmethod
    def _wavelength_cm(f, unit):
        return 100.0 * SPEED_OF_LIGHT / (f * utils.Units[unit])

    def _spectrum(self, fp, data, xaxis):
        wave = xaxis.startswith('wave')
        head = '# cm   K@b  \t' if wave else '# {}  K@b  \t'.format(data.freqUnit)
        disc = utils.b_type(data.b) == 'disc'
        head += 'disc' if disc else ''.join('({:5.3f},{:5.3f})\t'.format(bv[0], bv[1]) for bv in data.b)
        fp.write(head.strip('\t') + '\n')
        Tb = np.atleast_2d(np.asarray(data.Tb))
        side = open(self.scratch_spec, 'w') if self.scratch_spec else None
        for i, f in enumerate(data.f):
            s = '{:.4f}\t  '.format(self._wavelength_cm(f, data.freqUnit)) if wave else '{:.2f}\t  '.format(f)
            s += ''.join('  {:9.4f}  \t'.format(Tb[j][i]) for j in range(Tb.shape[0]))
            s = s.strip() + '\n'
            fp.write(s)
            if side:
                side.write(s)
        if side:
            side.close()

    def _profile(self, fp, data, xaxis):
        wave = xaxis.startswith('wave')
        head = '# b  K@cm  \t' if wave else '# b  K@{} \t'.format(data.freqUnit)
        for f in data.f:
            head += '  {:.4f}   \t'.format(self._wavelength_cm(f, data.freqUnit)) if wave else '  {:9.4f}   \t'.format(f)
        fp.write(head.strip() + '\n')
        disc = utils.b_type(data.b) == 'disc'
        Tb = np.atleast_2d(np.asarray(data.Tb))
        for i, bv in enumerate(data.b):
            s = 'disc' if disc else '{:5.3f} {:5.3f}\t'.format(bv[0], bv[1])
            s += ''.join(' {:7.2f}\t '.format(Tb[i][j]) for j in range(Tb.shape[1]))
            fp.write(s.strip() + '\n')

    def _image(self, fp, data):
        Tb = np.asarray(data.Tb)
        planes = [Tb] if Tb.ndim == 2 else [Tb[:, :, k] for k in range(Tb.shape[2])]
        for k, plane in enumerate(planes):
            if len(planes) > 1:
                fp.write('# f = {}\n'.format(data.f[k]))
            for row in plane:
                fp.write('\t'.join('{:7.2f}'.format(v) for v in row).strip() + '\n')
