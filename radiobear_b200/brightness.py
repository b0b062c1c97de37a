"""Brightness: brightness temperature along rays, computed by the rt_integrate CUDA kernel.

API mirror of Brightness.single (brightness.py:30-126).  `single` is the R = 1 case of `batch`,
which runs geometry + integration for any number of impact points in two kernel launches.
"""
import os.path

import numpy as np

from . import engine
from . import logging as rblog
from . import raypath
from . import utils
from .engine import T_CMB


class Brightness:
    def __init__(self, config=None, log=None, verbose=True, **kwargs):
        self.verbose = verbose
        self.log = rblog.setup(log)
        if config is None or isinstance(config, str):
            from . import config as pcfg
            config = pcfg.planetConfig('x', configFile=config)
            config.update_config(**kwargs)
        self.config = config
        self.tau_cut = engine.TAU_CUT
        self.travel = None

    def _args(self, atm, orientation):
        a = raypath._geometry_args(atm, orientation, None)
        a['T'] = atm.gas[atm.config.C['T']]
        return a

    def prefetch(self, b, atm, orientation=None):
        """Start the ray geometry of a coming batch(b, ...) now (it does not need the absorption), so that it
        overlaps Alpha.get_layers.  `b` must be the same float64 array later handed to batch()."""
        b = np.asarray(b, dtype=np.float64)
        if b.ndim == 2 and b.shape[0] >= 4096 and b.flags.c_contiguous:
            engine.geometry_prefetch(b=b, **raypath._geometry_args(atm, orientation, None))

    def batch(self, b, freqs, atm, alpha, orientation=None, disc_average=False, out_f32=False, want_intW=False):
        """Tb[R][F] for impact points b[R][2]; off-planet rays give T_cmb, limb rays below the tangent
        shell give NaN exactly like the reference."""
        if not alpha.has_layers():
            alpha.get_layers(freqs, atm)
        if getattr(alpha.config, 'Doppler', False):
            pts = np.atleast_2d(np.asarray(b, dtype=np.float64))
            F = len(freqs)
            Tb = np.empty((len(pts), F), dtype=np.float32 if out_f32 else np.float64)
            iW = np.zeros((len(pts), F))
            for r, p in enumerate(pts):                       # one absorption slab pair per ray
                one = self._doppler_ray([float(p[0]), float(p[1])], freqs, atm, alpha, orientation, disc_average, False)
                Tb[r] = T_CMB if one is None else one['Tb'][0]
                iW[r] = 0.0 if one is None else one['integrated_W'][0]
            return {'Tb': Tb, 'integrated_W': iW} if want_intW else {'Tb': Tb}
        res = engine.rt_batch(b=np.atleast_2d(np.asarray(b, dtype=np.float64)), alpha_slab=alpha.rt_slab(),
                              disc_average=disc_average, out_f32=out_f32, tau_cut=self.tau_cut, want_intW=want_intW,
                              **self._args(atm, orientation))
        return res

    def _doppler_ray(self, b, freqs, atm, alpha, orientation, disc_average, profile):
        """One ray with Doppler-shifted absorption (brightness.py:80-96).  Step i of the ray joins layer i (absorption
        a0) and layer i + 1 (a1); the reference evaluates a1 at f / doppler[i] and a0 at f / doppler[i + 1], takes
        dtau = (a0 + a1) ds / 2 and builds the weights from a1.  (Its call goes to `alpha.get_alpha`, the former name of
        `get_alpha_from_calc`, and fails today: tests/golden/make_golden.py restores the name to generate the vectors.)
        Two absorption launches with per-layer frequencies, then the integration kernel with the pair of slabs.
        Leaves self.travel; returns None when the ray misses the planet."""
        self.travel = travel = raypath.compute_ds(atm, b, orientation)
        if travel.ds is None:
            return None
        n = len(travel.ds)
        L = atm.gas.shape[1]
        f = np.asarray(freqs, dtype=np.float64)
        dop = np.asarray(travel.doppler, dtype=np.float64)
        f1 = np.tile(f, (L, 1))
        f0 = np.tile(f, (L, 1))
        if n > 1:
            f1[1:n] = f[None, :] / dop[:n - 1, None]          # a1 of step i = layer i + 1 at f / doppler[i]
            f0[0:n - 1] = f[None, :] / dop[1:n, None]          # a0 of step i = layer i     at f / doppler[i + 1]
        slab1 = alpha.layers_at(f1, atm)
        slab0 = alpha.layers_at(f0, atm)
        ds = np.zeros((1, L - 1))
        ds[0, :n] = travel.ds
        res = engine.rt_integrate(ds, [n], slab1, atm.gas[atm.config.C['T']], disc_average=disc_average,
                                  tau_cut=0.0 if profile else self.tau_cut, want_intW=True, alpha0_slab=slab0,
                                  profile_ray=0 if profile else -1)
        if not profile:
            res = {'Tb': res[0], 'integrated_W': res[1]}
        return res

    def single(self, b, freqs, atm, alpha, orientation=None, taulimit=20.0):
        """Brightness temperature along one ray path -> list[F]; leaves .tau .W .Tb_lyr [F][S], .P .z [S],
        .integrated_W [F] and .travel for the ray like the reference."""
        disc_average = utils.b_type(b).startswith('dis')
        if disc_average:
            b = [0.0, 0.0]
        self.alpha = alpha
        self.freqs = freqs
        self.b = b
        if not alpha.has_layers():
            alpha.get_layers(freqs, atm)
        if getattr(alpha.config, 'Doppler', False):
            res = self._doppler_ray(b, freqs, atm, alpha, orientation, disc_average, True)
        else:
            res = engine.rt_batch(b=np.asarray([b], dtype=np.float64), alpha_slab=alpha.rt_slab(), disc_average=disc_average,
                                  tau_cut=0.0, want_intW=True, profile_ray=0, **self._args(atm, orientation))
            self.travel = raypath.compute_ds(atm, b, orientation)
        if self.travel.ds is None:
            print('Off planet')
            self.Tb = [utils.T_cmb for _ in freqs]
            return self.Tb
        n = len(self.travel.ds)
        C = atm.config.C
        P, z = atm.gas[C['P']], atm.gas[C['Z']]
        self.tau = res['tau'][:, :n]
        self.W = res['W'][:, :n]
        self.Tb_lyr = res['Tb_lyr'][:, :n]
        self.integrated_W = res['integrated_W'][0]
        self.P = np.concatenate(([P[0]], (P[:n - 1] + P[1:n]) / 2.0))
        self.z = np.concatenate(([z[0]], (z[:n - 1] + z[1:n]) / 2.0))
        self.Tb = [float(x) for x in res['Tb'][0]]
        if self.verbose:
            for f, w in zip(freqs, self.integrated_W):
                if w < 0.96:
                    print("Weight correction at {:.2f} is {:.4f} (showing below 0.96)".format(f, w))
        return self.Tb

    # ---- profile tables of the last ray (brightness.py:128-250): the text formats the reference's users read back ----
    def _table(self, filename, columns, tight=False):
        """One row per segment: repr(P) <tab> z (2 decimals) <tab> one repr(value) per frequency.  `columns[i][j]` is
        frequency i at segment j.  `tight` is the weight file's variant (no tab before 'GHz' or the line end)."""
        nF, n = len(self.freqs), len(self.P)
        lines = ['#P  \tz  \t' + ''.join('{:.2f}\t'.format(f) for f in self.freqs)]
        for j in range(n):
            lines.append('{}\t{:.2f}\t'.format(repr(self.P[j]), self.z[j])
                         + ''.join('{}\t'.format(repr(columns[i][j])) for i in range(nF)))
        if tight:
            lines = [ln.strip() for ln in lines]
        with open(filename, 'w') as fp:
            fp.write(lines[0] + 'GHz\n')
            for ln in lines[1:]:
                fp.write(ln + '\n')
        return '{} ({} x {})'.format(filename, nF, n)

    def saveAlpha(self, filename=None, path='.'):
        """Absorption of the layers the ray crossed, [segment][frequency] (brightness.py:166-183); the only one of the
        four tables that honours `path` in the reference, and it returns nothing."""
        self._table(os.path.join(path, 'alpha.out' if filename is None else filename), self.alpha.layers)

    def saveWeight(self, norm=False, filename=None, path='.'):
        """Weighting functions, optionally each divided by its maximum (brightness.py:185-208).  Written to the working
        directory like the reference does (`path` is accepted and unused there)."""
        W = self.W
        if norm:
            W = [np.asarray(w) / np.max(w) for w in W]
        else:
            W = [np.asarray(w) / 1.0 for w in W]
        return self._table('wgt.out' if filename is None else filename, W, tight=True)

    def saveTau(self, filename=None, path='.'):
        """Optical depth down to each segment (brightness.py:210-228); working directory, see saveWeight -- but `path`
        is joined (and the result dropped) there, so path=None raises TypeError here too."""
        filename = 'tau.out' if filename is None else filename
        os.path.join(path, filename)
        return self._table(filename, self.tau)

    def saveTblayer(self, filename=None, path='.'):
        """Brightness accumulated down to each segment (brightness.py:230-250); working directory, see saveTau."""
        filename = 'tblayer.out' if filename is None else filename
        os.path.join(path, filename)
        return self._table(filename, self.Tb_lyr)

    def savertm(self, tag=None, path=None):
        """All four tables (brightness.py:128-149).  The reference hands (filename, path) to saveWeight, whose first
        parameter is `norm`: with a tag the weights come out normalised, and the file is named `path` ('wgt.out' when
        path is None -- and then saveTau raises TypeError; a directory raises IsADirectoryError).  Kept as it is, so that
        the same call leaves the same files and the same exception (tests/golden/rtm_tables.npz)."""
        name = (lambda kind: None) if tag is None else (lambda kind: '{}_{}.out'.format(kind, tag))
        self.saveAlpha(name('alpha'), self.config.output_directory)
        self.saveWeight(norm=name('wgt') is not None, filename=path)
        self.saveTau(name('tau'), path)
        self.saveTblayer(name('tblayer'), path)

    def saveit(self):
        """One file per frequency: pressure, absorption, weight, optical depth, brightness per segment
        (brightness.py:151-164)."""
        for i, f in enumerate(self.freqs):
            filename = 'pawtt_{:.3f}.out'.format(f)
            print("{}:  Pressure, alpha, weight, tau, Tb".format(filename))
            with open(filename, 'w') as fp:
                for j in range(len(self.P)):
                    cols = (self.P[j], self.alpha.layers[i][j], self.W[i][j], self.tau[i][j], self.Tb_lyr[i][j])
                    fp.write('\t'.join(repr(c) for c in cols) + '\n')
