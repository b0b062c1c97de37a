"""Atmosphere provider: gas / cloud profiles -> atm.gas[16,L], atm.cloud[12,L], atm.property[11,L].

Host-side, one-off O(L) setup feeding the kernels (not a GPU path).  Two ways to get one:

* ``Atmosphere(planet, config=...).std()`` reads the same text profiles and `config.par` as the
  reference and follows its pipeline: readGas / readCloud (atm_base.py:39-118), regrid onto a log-P
  grid with linear interpolation + adiabatic inward extrapolation (regrid.py:11-176), the user's
  tweak module hook ``modify(gas, cloud, C, Cl)`` (atm_modify.py:12-21), computeProp
  (atm_base.py:120-194).
* ``Atmosphere.from_arrays(...)`` / ``from_npz(...)`` wraps arrays that already exist (fixtures,
  MCMC drivers, bench.py).
"""
import importlib
import os
import sys

import numpy as np

from . import logging as rblog
from . import utils

GRAV_CONST = 6.6738e-20   # km3/kg/s2   (chemistry.py:7)
R_GAS = 8.314462          # J/K/mol     (chemistry.py:8)
# per-constituent properties used by computeProp / regrid (chemistry.py:18-50)
AMU = {'H2': 2.016, 'HE': 4.003, 'H2S': 34.076, 'NH3': 17.030, 'H2O': 18.015, 'CH4': 16.04, 'PH3': 33.997,
       'NH4SH': 51.110}
CP_OVER_R = {'H2': 3.5, 'HE': 2.50, 'H2O': 4.00, 'NH3': 4.46, 'H2S': 4.01, 'CH4': 4.50}
REFRACTIVITY = {'H2': 124.43, 'HE': 35.832, 'H2S': 2247.0, 'NH3': 2700.0, 'CH4': 413.0}


def refractivity(key, T):
    """Refractivity coefficient of a constituent (chemistry.py:33-50; water depends on T)."""
    if key == 'H2O':
        return 245.0 + 1.28E6 / T
    return REFRACTIVITY.get(key, 0.0)


class Atmosphere:
    def __init__(self, planet, idnum=0, config='config.par', log=None, verbose=False, **kwargs):
        self.planet = planet.capitalize()
        self.verbose = verbose
        self.log = rblog.setup(log)
        if config is None or isinstance(config, str):
            from . import config as pcfg
            cfile = None if config is None else os.path.join(self.planet, config)
            config = pcfg.planetConfig(self.planet, configFile=cfile)
            config.update_config(**kwargs)
        self.config = config
        self.idnum = idnum
        self.gas = np.zeros((0, 0))
        self.cloud = np.zeros((0, 0))
        self.property = np.zeros((0, 0))
        self.nAtm = 0
        self.tweakComment = ''
        if getattr(config, 'path', None) and config.path not in sys.path:
            sys.path.insert(0, config.path)

    # ---------------------------------------------------------------- ready-made arrays
    @classmethod
    def from_arrays(cls, planet, config, gas, cloud, prop):
        self = cls(planet, config=config)
        self.gas = np.array(gas, dtype=np.float64)
        self.cloud = np.array(cloud, dtype=np.float64)
        self.property = np.array(prop, dtype=np.float64)
        self.nAtm = self.gas.shape[1]
        return self

    @classmethod
    def from_npz(cls, path, planet):
        """Atmosphere snapshot: npz with gas / cloud / property, the column maps (C_keys, Cl_keys, LP_keys)
        and the config scalars the hot path needs (written by tests/golden/make_golden.py:cfg_arrays or
        `Atmosphere.to_npz`)."""
        from . import config as pcfg
        d = np.load(path)
        cfg = pcfg.planetConfig(planet, configFile=None)
        cfg.C = {str(k): i for i, k in enumerate(d['C_keys'])}
        cfg.Cl = {str(k): i for i, k in enumerate(d['Cl_keys'])}
        cfg.LP = {str(k): i for i, k in enumerate(d['LP_keys'])}
        cfg.Req, cfg.Rpol = float(d['Req']), float(d['Rpol'])
        cfg.orientation = [float(x) for x in d['orientation']]
        cfg.gtype, cfg.limb = str(d['gtype']), str(d['limb'])
        cfg.h2state, cfg.coshape = str(d['h2state']), str(d['coshape'])
        for key in ('distance', 'GM_ref', 'p_ref'):
            if key in d.files:
                setattr(cfg, key, float(d[key]))
        for c, f in zip(d['alpha_constituents'], d['alpha_formalisms']):
            cfg.constituent_alpha[str(c)] = None if str(f) == 'none' else str(f)
        for gas in ('h2s', 'ph3'):
            v = float(d['truncate_strength_' + gas])
            cfg.truncate_strength[gas] = None if np.isnan(v) else v
            cfg.truncate_method[gas] = None if np.isnan(v) else 'truncate_strength'
        return cls.from_arrays(planet, cfg, d['gas'], d['cloud'], d['property'])

    def to_npz(self, path):
        c = self.config
        ca = {k: ('none' if v is None else v) for k, v in c.constituent_alpha.items()}
        nan = float('nan')
        np.savez_compressed(
            path, gas=self.gas, cloud=self.cloud, property=self.property,
            C_keys=np.array(sorted(c.C, key=lambda k: c.C[k])), Cl_keys=np.array(sorted(c.Cl, key=lambda k: c.Cl[k])),
            LP_keys=np.array(sorted(c.LP, key=lambda k: c.LP[k])), Req=c.Req, Rpol=c.Rpol,
            orientation=np.array(c.orientation[:2], dtype=float), gtype=c.gtype, limb=c.limb, h2state=c.h2state,
            coshape=c.coshape, alpha_constituents=np.array(sorted(ca)), alpha_formalisms=np.array([ca[k] for k in sorted(ca)]),
            truncate_strength_h2s=nan if c.truncate_strength.get('h2s') is None else c.truncate_strength['h2s'],
            truncate_strength_ph3=nan if c.truncate_strength.get('ph3') is None else c.truncate_strength['ph3'],
            distance=c.distance, GM_ref=c.GM_ref, p_ref=c.p_ref)

    # ---------------------------------------------------------------- file pipeline
    def _read_table(self, filename, cmap):
        """Columns of a whitespace table, keeping rows with the most common field count
        (utils.get_expected_number_of_entries, utils.py:174-195)."""
        rows = []
        with open(filename, 'r') as fp:
            for line in fp:
                vals = utils.data_line(line)
                if vals is not None:
                    rows.append(vals)
        counts = {}
        for r in rows:
            counts[len(r)] = counts.get(len(r), 0) + 1
        order = sorted(counts.values(), reverse=True)
        if len(order) > 1 and 2 * sum(order[1:]) > order[0]:
            raise ValueError("Not enough data lines in file: {}".format(order))
        width = max(counts, key=lambda k: counts[k])
        rows = np.array([r for r in rows if len(r) == width], dtype=np.float64)
        ncol = len(cmap)
        out = np.zeros((ncol, rows.shape[0]))
        use = min(ncol, width)
        out[:use] = rows[:, :use].T
        # re-index the column map by rank of its values (atm_base.py:63-66)
        for i, k in enumerate(sorted(cmap, key=lambda k: cmap[k])):
            cmap[k] = i
        P = out[cmap['P']]
        if not np.all(np.diff(P) > 0.0):
            out = np.fliplr(out)
            if not np.all(np.diff(out[cmap['P']]) > 0.0):
                raise ValueError("Pressure not monotonically increasing in {}.".format(filename))
        return np.ascontiguousarray(out)

    @staticmethod
    def _renorm_z(arr, cmap):
        """Deepest z becomes 0; DZ = |diff z| in cm (atm_base.py:205-219)."""
        z = arr[cmap['Z']]
        z -= z[-1]
        arr[cmap['DZ']] = np.append(np.array([0.0]), np.abs(np.diff(z)) * 1.0E5)

    def readGas(self):
        fn = os.path.join(self.config.path, self.config.gasFile[self.idnum])
        self.gas = self._read_table(fn, self.config.C)
        self._renorm_z(self.gas, self.config.C)
        self.nAtm = self.gas.shape[1]

    def readCloud(self):
        fn = os.path.join(self.config.path, self.config.cloudFile[self.idnum])
        self.cloud = self._read_table(fn, self.config.Cl)
        self._renorm_z(self.cloud, self.config.Cl)

    def _chem_keys(self):
        return [k for k in self.config.C if k not in ('P', 'T', 'Z', 'DZ')]

    def computeProp(self):
        """Derived layer properties (atm_base.py:120-194): Z R P GM AMU REFR N H LAPSE LAPSEP g."""
        C, LP = self.config.C, self.config.LP
        gas = self.gas
        P, T, Z = gas[C['P']], gas[C['T']], gas[C['Z']]
        L = gas.shape[1]
        prop = np.zeros((len(LP), L))
        iOffset = int(np.argmin(np.abs(P - self.config.p_ref)))    # first nearest (strict '<' scan)
        R = self.config.Req + Z - Z[iOffset]
        keys = self._chem_keys()
        amu = np.zeros(L)
        refr = np.zeros(L)
        for k in keys:
            amu = amu + AMU.get(k, 0.0) * gas[C[k]]
        for k in keys:
            refr = refr + refractivity(k, T) * gas[C[k]]
        refr = refr * P * (293.0 / T)
        GM = np.zeros(L)
        lapse = np.zeros(L)
        lapsep = np.zeros(L)
        if L > 1:
            rho = (amu[1:] * P[1:]) / (R_GAS * T[1:])
            dr = np.abs(np.diff(Z))
            dM = 1.0e11 * rho * (4.0 * np.pi * (R[1:]**2) * dr)
            GM[1:] = np.cumsum(GRAV_CONST * dM)
            dT, dP = np.abs(np.diff(T)), np.abs(np.diff(P))
            with np.errstate(divide='ignore', invalid='ignore'):
                lapse[1:] = dT / dr
                lapsep[1:] = dT / dP
        gm = self.config.GM_ref - (GM - GM[iOffset])
        g = gm / R**2
        with np.errstate(divide='ignore', invalid='ignore'):
            H = (R_GAS * T) / (g * amu) / 1000.0
        prop[LP['P']], prop[LP['Z']], prop[LP['R']] = P, Z, R
        prop[LP['AMU']], prop[LP['GM']] = amu, gm
        prop[LP['LAPSE']], prop[LP['LAPSEP']] = lapse, lapsep
        prop[LP['REFR']], prop[LP['N']] = refr, refr / 1.0E6 + 1.0
        prop[LP['H']], prop[LP['g']] = H, g
        self.property = prop

    # regrid.py:11-176
    def regrid(self, regridType=None, Pmin=None, Pmax=None):
        C, Cl = self.config.C, self.config.Cl
        if regridType is None:
            regridType = self.config.regridType
        if regridType is None or (isinstance(regridType, str) and regridType.lower() == 'none'):
            return 0
        P_in = self.gas[C['P']]
        if Pmin is None or Pmin == 'auto' or Pmin == 0:
            Pmin = P_in.min()
        if Pmax is None or Pmax == 0:
            Pmax = P_in.max()
        if isinstance(regridType, str):
            try:
                regridType = int(regridType)
            except ValueError:
                Pgrid = np.loadtxt(os.path.join(self.config.path, regridType))
                if not np.all(np.diff(Pgrid) > 0.0):
                    Pgrid = Pgrid[::-1]
                if not np.all(np.diff(Pgrid) > 0.0):
                    raise ValueError('Error in regrid')
        if isinstance(regridType, int):
            Pgrid = np.logspace(np.log10(Pmin), np.log10(Pmax), regridType)
        fill = -999.9
        gas = self._interp(self.gas, C, Pgrid, fill)
        if np.any(gas == fill):
            self.computeProp()
            gas = self._extrapolate(gas, fill)
        if np.any(gas == fill):
            raise ValueError("fillval still in gas!")
        cloud = self._interp(self.cloud, Cl, Pgrid, 0.0)
        self.gas, self.cloud = gas, cloud
        self._renorm_z(self.gas, C)
        self._renorm_z(self.cloud, Cl)
        self.nAtm = self.gas.shape[1]
        return 1

    @staticmethod
    def _interp(src, cmap, Pgrid, fill):
        """Linear interpolation in P of every column except P / DZ (regrid.py:97-118)."""
        out = np.zeros((src.shape[0], len(Pgrid)))
        Pin = src[cmap['P']]
        out[cmap['P']] = Pgrid
        inside = (Pgrid >= Pin[0]) & (Pgrid <= Pin[-1])
        # scipy interp1d (linear): y0 + (x - x0) * (y1 - y0) / (x1 - x0) on the bracketing pair
        hi = np.clip(np.searchsorted(Pin, Pgrid, side='left'), 1, len(Pin) - 1)
        lo = hi - 1
        slope_den = Pin[hi] - Pin[lo]
        for name, row in cmap.items():
            if name in ('P', 'DZ'):
                continue
            y = src[row]
            val = (y[hi] - y[lo]) / slope_den * (Pgrid - Pin[lo]) + y[lo]
            out[row] = np.where(inside, val, fill)
        return out

    def _extrapolate(self, gas, fill):
        """Inward: fixed mixing ratios + dry adiabat in hydrostatic equilibrium; outward: last slope
        (regrid.py:121-176)."""
        C, LP = self.config.C, self.config.LP
        old = self.gas
        keys = self._chem_keys()
        for k in keys:
            row = gas[C[k]]
            row[row == fill] = old[C[k]][-1]
        pDeep = old[C['P']][-1]
        g = self.property[LP['g']][-1]
        r = self.property[LP['R']][-1]
        P = gas[C['P']]
        for i in range(len(P)):
            p = P[i]
            if p < pDeep:
                continue
            prev = P[i - 1]
            dP = p - prev
            T = gas[C['T']][i - 1]
            z = gas[C['Z']][i - 1]
            amu = 0.0
            cp = 0.0
            for k in keys:
                cp += CP_OVER_R.get(k, 0.0) * gas[C[k]][i]
                amu += AMU.get(k, 0.0) * gas[C[k]][i]
            gas[C['T']][i] = T + T / (cp * p) * dP
            g = g + 2.0 * R_GAS * T * np.log(p / prev) / (r * amu) / 1000.0
            H = R_GAS * T / (amu * g) / 1000.0
            dz = H * dP / p
            r = r - dz
            gas[C['Z']][i] = z - dz
        if np.any(gas == fill):
            x = gas[C['P']]
            for name, row in C.items():
                if name in ('P', 'DZ'):
                    continue
                y = gas[row]
                i = int(np.argmax(y != fill))
                slope = (y[i + 1] - y[i]) / (x[i + 1] - x[i])
                for j in range(i - 1, -1, -1):
                    y[j] = y[j + 1] + slope * (x[j + 1] - x[j])
                    if y[j] <= 0.0:
                        y[j] = 1e-20
        return gas

    @staticmethod
    def is_present(c, tiny=1.0E-30):
        """(does any entry of the profile exceed `tiny`, the profile with everything at or below it raised to `tiny`)
        -- for log plots (atm_base.py:196-204)."""
        floor = np.maximum(np.asarray(c, dtype=np.float64), tiny)
        return bool((floor > tiny).any()), list(floor)

    def tweak(self):
        """Run the user's tweak module: modify(gas, cloud, C, Cl) -> (comment, gas, cloud)
        (atm_modify.py:12-21)."""
        name = self.config.tweakmodule
        if name is None or str(name).lower() == 'none':
            return
        mod = importlib.import_module(name)
        self.tweakComment, self.gas, self.cloud = mod.modify(self.gas, self.cloud, self.config.C, self.config.Cl)
        self.log.add(self.tweakComment, False)

    def simple(self, **kwargs):
        self.readGas()
        self.readCloud()
        self.computeProp()
        self.nAtm = self.gas.shape[1]

    def std(self, Pmin=None, Pmax=None, regridType=None, tweak=True, **kwargs):
        """Standard pipeline (atmosphere.py:66-116)."""
        self.readGas()
        self.readCloud()
        self.regrid(regridType=regridType, Pmin=self.config.pmin if Pmin is None else Pmin,
                    Pmax=self.config.pmax if Pmax is None else Pmax)
        if tweak:
            self.tweak()
        self.computeProp()
        return self.nAtm
