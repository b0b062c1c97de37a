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

    # ------------------------------------------------------------------ reading back (fileIO.py:109-330)
    def flist(self, files=None, tag=None):
        """File names to read: an index / list of indices into the sorted directory listing (filtered by `tag`),
        a name, a comma-separated string or a list of names (fileIO.py:109-155).  `None` takes every match
        (the reference prompts on stdin)."""
        names = sorted(os.path.join(self.directory, fn) for fn in os.listdir(self.directory)
                       if fn[0] != '.' and (tag is None or tag in fn)) if os.path.isdir(self.directory) else []
        if isinstance(files, str):
            return files.split(',')
        if isinstance(files, (list, tuple)) and files and isinstance(files[0], str):
            return list(files)
        if not names:
            print("No files found in {}".format(self.directory))
            return []
        if files is None:
            return names
        idx = [files] if isinstance(files, int) else list(files)
        return [names[i] for i in idx]

    def read_header(self, header_text):
        """'# key: value' lines -> dict; '#*' marks the keys copied onto the Data object (fileIO.py:305-330)."""
        header = {}
        self.starred_header_keys = []
        for line in header_text:
            body = line.strip('#').strip()
            if 'K@' in line:
                header['label-line'] = body
                continue
            key, value = (body.split(':', 1) + [''])[:2] if ':' in line else (body.split()[0] if body.split() else '', body)
            key, value = key.strip(), value.strip()
            if key in header:
                key = 'other_{}'.format(key)
            if key.startswith('*'):
                key = key.strip('*').strip()
                self.starred_header_keys.append(key)
            header[key] = value
        if 'res' in header:
            self.resolution = header['res'].split()[0]
        return header

    def read(self, fn=None, tag='dat', file_type='spectrum'):
        """Read spectrum / profile / image files written by `write` (or by the reference): fills self.files and
        self.data[filename] (Data with header, f, b, Tb -- Tb[b][f] like DataReturn; an image is Tb[row][col],
        or Tb[row][col][f] for the multi-frequency blocks this writer produces)."""
        from . import data_handling
        self.files, self.data = [], {}
        for filename in self.flist(fn, tag):
            if file_type.lower() != 'all' and file_type.lower() not in os.path.basename(filename).lower():
                continue
            try:
                lines = open(filename).read().splitlines()
            except IOError:
                print(filename + " not found - removing from list")
                continue
            d = data_handling.Data()
            d.header = self.read_header([ln for ln in lines if ln.startswith('#')])
            for key in self.starred_header_keys:
                if key in d.allowed_parameters:
                    setattr(d, key, d.header[key])
            kind = d.header.get('type', file_type).lower()
            rows = [[float(x) for x in ln.split()] for ln in lines if ln.strip() and not ln.startswith('#')
                    and not ln.startswith('disc')]
            if 'image' in kind:
                planes, cur = [], []
                for ln in lines:
                    if ln.startswith('# f ='):
                        if cur:
                            planes.append(cur)
                        cur = []
                    elif ln.strip() and not ln.startswith('#'):
                        cur.append([float(x) for x in ln.split()])
                planes.append(cur)
                d.Tb = np.array(planes[0]) if len(planes) == 1 else np.stack([np.array(p) for p in planes], axis=2)
                d.f = np.array([float(ln.split('=')[1]) for ln in lines if ln.startswith('# f =')])
                d.b = getattr(self, 'resolution', None)
            else:
                labels = d.header.get('label-line', '').split()[2:]
                tab = np.array(rows)
                if 'spectrum' in kind:
                    d.f = tab[:, 0]
                    d.Tb = tab[:, 1:].T
                    d.b = ['disc'] if labels == ['disc'] else \
                        np.array([[float(x) for x in lb.strip('()').split(',')] for lb in labels])
                else:                                   # profile: 'bx by  Tb(f0) Tb(f1) ...'
                    d.f = np.array([float(x) for x in labels])
                    d.b = tab[:, :2]
                    d.Tb = tab[:, 2:]
            d.type = kind
            self.files.append(filename)
            self.data[filename] = d
        return self.data

    def show(self, property='all'):
        for filename, d in getattr(self, 'data', {}).items():
            print(filename)
            keys = d.allowed_parameters if property == 'all' else ([property] if isinstance(property, str) else property)
            # the log of the run is printed when the file's header names one that is still there
            d.show(include=[k for k in keys if k != 'log' or (d.logfile and os.path.isfile(str(d.logfile)))])
