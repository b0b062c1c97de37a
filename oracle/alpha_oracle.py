"""Oracle (TEST INFRASTRUCTURE, not product code): absorption coefficients on the CPU.

Plain numpy restatement of the reference's per-layer absorption plugins and of the
``Alpha.get_layers`` driver.  Each function cites the reference file:line it follows
(paths relative to /root/reference/radiobear).  Arithmetic is float64 throughout and
keeps the reference's order of operations where it matters for rounding.

Plugin signature (same as the reference, constituents/<gas>/<formalism>.py):
    alpha(freq, T, P, X, P_dict, other_dict, **kwargs) -> ndarray[len(freq)]
kwargs: truncate_freq, truncate_strength, units ('invcm' | 'dBperkm'), cat (LineCatalog).

Pinned by tests/test_oracle_golden.py against vectors generated from the reference
(tests/golden/make_golden.py).
"""
import os
import numpy as np

OPTICALDEPTH_TO_DB = 434294.5   # cm^-1 -> dB/km (nh3_hs.py:46, h2s_ddb.py:83)

_DEFAULT_LINECAT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                'radiobear_b200', 'data', 'linecat.npz')


class LineCatalog:
    """Line catalogs (float64) with the reference's load-time truncation rules.

    h2s_ddb.py:23-38, ph3_jh.py:28-61: lines are kept when I0 > truncate_strength and
    f0 < truncate_freq.  The reference applies this once at first load (module cache).
    """

    def __init__(self, path=_DEFAULT_LINECAT, full_nh3=False):
        """full_nh3: the untrimmed rotational / roto-vibrational NH3 line lists (1301 / 4198 lines) instead of the
        201 / 198 the reference ships in ammonia.npz (constituents/txt2npz.py:21-56) -- SURVEY 8d "full catalog"."""
        d = np.load(path)
        self.raw = {k: np.array(d[k]) for k in d.files if not k.endswith('_cols')}
        if full_nh3:
            self.raw['nh3_rot'], self.raw['nh3_v2'] = self.raw['nh3_rot_full'], self.raw['nh3_v2_full']
        self._cache = {}

    def get(self, name, truncate_strength=None, truncate_freq=None):
        key = (name, truncate_strength, truncate_freq)
        if key in self._cache:
            return self._cache[key]
        if name in ('nh3_inv', 'nh3_rot', 'nh3_v2', 'nh3_sjs', 'co'):
            out = self.raw[name]          # these plugins never truncate
        elif name == 'h2s':
            a = self.raw['h2s']
            if truncate_strength is not None:
                a = a[:, a[1] > truncate_strength]
            if truncate_freq is not None:
                a = a[:, a[0] < truncate_freq]
            out = a
        elif name == 'ph3':
            a = np.vstack([self.raw['ph3'], self.raw['ph3_wgt']])   # f0 I0 E WgtI0 WgtFGB WgtSB
            # ph3_jh.py:28-33 uses "is not None" for the lines and :52 truthiness for the
            # weights; the two only differ for truncate_strength == 0 where the length
            # check (:56) raises in the reference.
            if truncate_strength is not None:
                a = a[:, a[1] > truncate_strength]
            if truncate_freq is not None:
                a = a[:, a[0] < truncate_freq]
            out = a
        else:
            raise KeyError(name)
        self._cache[key] = out
        return out


_default_cat = None


def default_catalog():
    global _default_cat
    if _default_cat is None:
        _default_cat = LineCatalog()
    return _default_cat


def _par(kwargs):
    """parameters.py:4-8 -- defaults units='dBperkm'."""
    units = kwargs.get('units', 'dBperkm')
    cat = kwargs.get('cat', None) or default_catalog()
    return units, cat, kwargs.get('truncate_strength', None), kwargs.get('truncate_freq', None)


# --------------------------------------------------------------------------------------
# NH3: Hanley/Steffes (nh3_hs) and Devaraj/Bellotti/Steffes (nh3_dbs)
# --------------------------------------------------------------------------------------
_NH3_INV_SETS = {
    # nh3_hs.py:139-170
    'hs': {
        'lo': dict(gnu_H2=1.640, gnu_He=0.75, gnu_NH3=0.852, GAMMA_H2=0.7756, GAMMA_He=0.666,
                   GAMMA_NH3=1.0, zeta_H2=1.262, zeta_He=0.3, zeta_NH3=0.5296, Z_H2=0.7964,
                   Z_He=0.667, Z_NH3=1.554, d=-0.0498, Con=0.9301),
        'hi': dict(gnu_H2=1.7465, gnu_He=0.9779, gnu_NH3=0.7298, GAMMA_H2=0.8202, GAMMA_He=1.0,
                   GAMMA_NH3=1.0, zeta_H2=1.2163, zeta_He=0.0291, zeta_NH3=0.5152, Z_H2=0.8873,
                   Z_He=0.8994, Z_NH3=2.0 / 3.0, d=-0.0627, Con=0.9862)},
    # nh3_dbs.py:135-166
    'dbs': {
        'lo': dict(gnu_H2=1.6937, gnu_He=0.6997, gnu_NH3=0.7523, GAMMA_H2=0.8085, GAMMA_He=1.0,
                   GAMMA_NH3=1.0, zeta_H2=1.3263, zeta_He=0.1607, zeta_NH3=0.6162, Z_H2=0.8199,
                   Z_He=0.0, Z_NH3=1.3832, d=-0.0139, Con=0.9619),
        'hi': dict(gnu_H2=1.7465, gnu_He=0.9779, gnu_NH3=0.7298, GAMMA_H2=0.8202, GAMMA_He=1.0,
                   GAMMA_NH3=1.0, zeta_H2=1.2163, zeta_He=0.0291, zeta_NH3=0.5152, Z_H2=0.8873,
                   Z_He=0.8994, Z_NH3=2.0 / 3.0, d=-0.0627, Con=0.9862)},
}
NH3_F_SPLIT = 30.0              # nh3_hs.py:58
_GHZ_TO_INVCM = 1 / 29.9792458  # nh3_hs.py:45
_HC = 19.858252418E-24          # nh3_hs.py:50
_KB = 1.38E-23                  # nh3_hs.py:51
_COEF_NH3 = 1.0E6 * 6.02297E23 / 8.31432E7   # nh3_hs.py:52-56
_TO = 300.0


def _nh3_consistent(freq, T, P, P_h2, P_he, P_nh3, cat, units, family, band, consts=None, clamp_le=False):
    """nh3_hs.py:91-312 / nh3_dbs.py:91-313 for one band ('lo': f<=30, 'hi': f>30); nh3_kd.py:115-351 passes
    its own pressure-dependent constants and clamps `<= 0` instead of `< 0` (nh3_kd.py:344-349)."""
    c = consts if consts is not None else _NH3_INV_SETS[family][band]
    fo, Io, Eo, gammaNH3o = cat.get('nh3_inv')
    fo_rot, Io_rot, Eo_rot, gNH3_rot, gH2_rot, gHe_rot = cat.get('nh3_rot')
    fo_v2, Io_v2, Eo_v2 = cat.get('nh3_v2')
    f = np.asarray(freq, dtype=np.float64)[None, :]        # [1, F]
    Tdiv = _TO / T
    eta = 3.0 / 2.0

    # inversion lines (nh3_hs.py:172-226)
    gH2 = c['gnu_H2'] * P_h2
    gHe = c['gnu_He'] * P_he
    gNH3 = c['gnu_NH3'] * P_nh3 * gammaNH3o
    gamma = gH2 * Tdiv**c['GAMMA_H2'] + gHe * Tdiv**c['GAMMA_He'] + gNH3 * (295.0 / T)**c['GAMMA_NH3']
    delt = c['d'] * gamma
    zH2 = c['zeta_H2'] * P_h2
    zHe = c['zeta_He'] * P_he
    zNH3 = c['zeta_NH3'] * P_nh3 * gammaNH3o
    zeta = zH2 * Tdiv**c['Z_H2'] + zHe * Tdiv**c['Z_He'] + zNH3 * (295.0 / T)**c['Z_NH3']
    expo = -(1.0 / T - 1.0 / _TO) * Eo * _HC / _KB
    ST = Io * np.exp(expo)
    alpha_noshape = c['Con'] * _COEF_NH3 * (P_nh3 / _TO) * (np.power(_TO / T, eta + 2.0)) * ST
    fo_m = fo[:, None]
    dnu = gamma[:, None]
    ce = zeta[:, None]
    pst = delt[:, None]
    Aa = (2.0 / np.pi) * np.square(f / fo_m)
    Bb = (dnu - ce) * np.square(f)
    Cc = dnu + ce
    Dd = np.square(fo_m + pst) + np.square(dnu) - np.square(ce)
    Ee = np.square(f)
    Jj = np.square(fo_m + pst)
    Gg = np.square(dnu)
    Hh = np.square(ce)
    Ii = 4.0 * np.square(f) * np.square(dnu)
    Ff = (Aa * (Bb + Cc * Dd)) / (np.square(Ee - Jj - Gg + Hh) + Ii)
    Fbr = (1.0 / _GHZ_TO_INVCM) * Ff
    alpha_inversion = alpha_noshape[:, None] * Fbr

    # rotational lines, Gross lineshape (nh3_hs.py:228-265)
    ST_rot = Io_rot * np.exp((1.0 / _TO - 1.0 / T) * Eo_rot * _HC / _KB)
    gamma_rot = (0.2984 * P_h2 * gH2_rot * Tdiv**0.8730 + 0.75 * P_he * gHe_rot * Tdiv**(2.0 / 3.0)
                 + 3.1789 * P_nh3 * gNH3_rot * Tdiv**1.0)
    dnu = gamma_rot[:, None]
    fo_m = fo_rot[:, None]
    Aa = (4.0 / np.pi) * np.square(f) * dnu
    Bb = np.square(np.square(fo_m) - np.square(f))
    Cc = 4.0 * np.square(f) * np.square(dnu)
    Fbr_rot = (1 / _GHZ_TO_INVCM) * (Aa / (Bb + Cc))
    alpha_rot = (2.4268 * _COEF_NH3 * (P_nh3 / _TO) * ((_TO / T)**(eta + 2.0)) * ST_rot)[:, None] * Fbr_rot

    # v2 roto-vibrational lines, Gross lineshape (nh3_hs.py:267-302)
    ST_v2 = Io_v2 * (np.exp((1. / _TO - 1. / T) * Eo_v2 * _HC / _KB))
    gamma_v2 = (P_h2 * 1.4) * Tdiv**0.73 + (P_he * 0.68) * (Tdiv**0.5716) + (P_nh3 * 9.5) * Tdiv**1.0
    dnu = np.full((len(fo_v2), 1), gamma_v2)
    fo_m = fo_v2[:, None]
    Aa = (4.0 / np.pi) * np.square(f) * dnu
    Bb = np.square(np.square(fo_m) - np.square(f))
    Cc = 4.0 * np.square(f) * np.square(dnu)
    Fbr_v2 = (1.0 / _GHZ_TO_INVCM) * (Aa / (Bb + Cc))
    alpha_v2 = (1.1206 * _COEF_NH3 * (P_nh3 / _TO) * ((_TO / T)**(eta + 2.0)) * ST_v2)[:, None] * Fbr_v2

    # total (nh3_hs.py:304-312): unit factor first, then the <0 clamp in output units
    a = np.sum(alpha_inversion, 0) + np.sum(alpha_rot, 0) + np.sum(alpha_v2, 0)
    if units == 'dBperkm':
        a = a * OPTICALDEPTH_TO_DB
    a = np.array(a)
    if clamp_le:
        a[a <= 0.0] = 1.0E-8
    else:
        a[a < 0.0] = 1.0E-8
    return a


def nh3_kd_constants(P):
    """Pressure switch of the inversion-line constants, linear between 12 and 20 bar (nh3_kd.py:160-195)."""
    hi = dict(gnu_H2=1.6361, gnu_He=0.4555, gnu_NH3=0.7298, GAMMA_H2=0.8, GAMMA_He=0.5, GAMMA_NH3=1.0,
              zeta_H2=1.1313, zeta_He=0.1, zeta_NH3=0.5152, Z_H2=0.6234, Z_He=0.5, Z_NH3=2.0 / 3.0, d=0.2, Con=1.3746)
    lo = dict(gnu_H2=1.7465, gnu_He=0.9779, gnu_NH3=0.7298, GAMMA_H2=0.8202, GAMMA_He=1.0, GAMMA_NH3=1.0,
              zeta_H2=1.2163, zeta_He=0.0291, zeta_NH3=0.5152, Z_H2=0.8873, Z_He=0.8994, Z_NH3=2.0 / 3.0,
              d=-0.0627, Con=0.9862)
    P_trans, dP_up, dP_down = 15.0, 5.0, 3.0
    if P > P_trans + dP_up:
        return hi
    if P <= P_trans - dP_down:
        return lo
    w = (15.0 - dP_down - P) / (dP_up + dP_down)
    out = {}
    for k in lo:
        out[k] = lo[k] if k in ('gnu_NH3', 'GAMMA_NH3', 'zeta_NH3', 'Z_NH3') else lo[k] + (lo[k] - hi[k]) * w
    return out


def nh3_kd(freq, T, P, X, P_dict, other_dict, **kwargs):
    """nh3_kd.py:115-351: the consistent model with pressure-switched constants and no 30 GHz split."""
    units, cat, _, _ = _par(kwargs)
    freq = np.array(freq, dtype=np.float64)
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_nh3 = P * X[P_dict['NH3']]
    return _nh3_consistent(freq, T, P, P_h2, P_he, P_nh3, cat, units, None, None, consts=nh3_kd_constants(P),
                           clamp_le=True)


def nh3_bg(freq, T, P, X, P_dict, other_dict, **kwargs):
    """nh3_bg.py:26-74: Berge-Gulkis Ben-Reuven sum over the nh3.npz catalog (scalar loops in the reference)."""
    units, cat, _, _ = _par(kwargs)
    T0 = 296.0
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_nh3 = P * X[P_dict['NH3']]
    f0, I0, E, G0 = cat.get('nh3_sjs')
    delta = -0.45 * P_nh3
    out = []
    for f in np.asarray(freq, dtype=np.float64):
        f2 = f**2
        acc = 0.0
        for i in range(len(f0)):
            gamma = pow((T0 / T), 2.0 / 3.0) * (2.318 * P_h2 + 0.790 * P_he + G0[i] * 0.750 * P_nh3)
            g2 = gamma**2
            zeta = pow((T0 / T), 2.0 / 3.0) * (1.920 * P_h2 + 0.300 * P_he + G0[i] * 0.490 * P_nh3)
            z2 = zeta**2
            ITG = I0[i] * np.exp(-((1.0 / T) - (1.0 / T0)) * E[i] * _HCK)
            num = (gamma - zeta) * f2 + (gamma + zeta) * (pow(f0[i] + delta, 2.0) + g2 - z2)
            den = pow((f2 - pow(f0[i] + delta, 2.0) - g2 + z2), 2.0) + 4.0 * f2 * g2
            acc += _GHZ * 2.0 * pow(f / f0[i], 2.0) * num / (np.pi * den) * ITG
        a = _COEF_GEISA * (P_nh3 / T0) * pow((T0 / T), 3.0 / 2.0 + 2) * acc
        if units == 'dBperkm':
            a *= OPTICALDEPTH_TO_DB
        out.append(a)
    return np.array(out)


def _nh3_split(family, freq, T, P, X, P_dict, other_dict, **kwargs):
    """nh3_hs.py:70-88 -- split at 30 GHz, concatenate lo then hi."""
    units, cat, _, _ = _par(kwargs)
    freq = np.array(freq, dtype=np.float64)
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_nh3 = P * X[P_dict['NH3']]
    out = None
    lo = freq[freq <= NH3_F_SPLIT]
    if len(lo):
        out = _nh3_consistent(lo, T, P, P_h2, P_he, P_nh3, cat, units, family, 'lo')
    hi = freq[freq > NH3_F_SPLIT]
    if len(hi):
        a_hi = _nh3_consistent(hi, T, P, P_h2, P_he, P_nh3, cat, units, family, 'hi')
        out = a_hi if out is None else np.concatenate((out, a_hi))
    return out


def nh3_hs(freq, T, P, X, P_dict, other_dict, **kwargs):
    return _nh3_split('hs', freq, T, P, X, P_dict, other_dict, **kwargs)


def nh3_dbs(freq, T, P, X, P_dict, other_dict, **kwargs):
    return _nh3_split('dbs', freq, T, P, X, P_dict, other_dict, **kwargs)


# --------------------------------------------------------------------------------------
# NH3: Spilker / Joiner-Steffes (nh3_sjs.py:26-128)
# --------------------------------------------------------------------------------------
_COEF_GEISA = 7.244E+21   # nh3_sjs.py:6, h2s_ddb.py:6, ph3_jh.py:7, co_ddb.py:6
_HCK = 1.438396
_GHZ = 29.9792458


def nh3_sjs(freq, T, P, X, P_dict, other_dict, **kwargs):
    units, cat, _, _ = _par(kwargs)
    T0 = 296.0
    fLower, fHigher, EPS = 26.0, 34.0, 1E-12
    Joiner, Spilker, Interp = 0, 1, 2
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_nh3 = P * X[P_dict['NH3']]
    Pscale = 1.0 + P / 1.0E5
    GH2, GHe, GNH3 = [1.690], [0.750], [0.6]
    ZH2, ZHe, ZNH3 = [1.350], [0.300], [0.200]
    C, D = [1.0], [-0.45]
    rexp = 8.79 * np.exp(-T / 83.0)
    GH2a = np.exp(9.024 - T / 20.3) - 0.9918 + P_h2
    with np.errstate(invalid='ignore'):
        GH2a = np.power(GH2a, rexp)     # NaN for a negative base (nh3_sjs.py:54-57 never raises)
    if GH2a < EPS:
        GH2.append(1.690), GHe.append(0.750), GNH3.append(0.60)
        ZH2.append(1.35), ZHe.append(0.30), ZNH3.append(0.20)
        C.append(1.00), D.append(-0.45)
    else:
        GH2a = 2.122 * np.exp(-T / 116.8) / GH2a
        GH2a = 2.34 * (1.0 - GH2a)
        GH2.append(GH2a), GHe.append(0.46 + T / 3000.0), GNH3.append(0.74)
        ZH2.append(5.7465 - 7.7644 * GH2a + 9.1931 * GH2a**2 - 5.6816 * GH2a**3 + 1.2307 * GH2a**4)
        ZHe.append(0.28 - T / 1750.0), ZNH3.append(0.50)
        C.append(-0.337 + T / 110.4 - T**2 / 70600.0), D.append(-0.45)
    for lst in (GH2, GHe, GNH3, ZH2, ZHe, ZNH3, C, D):
        lst.append(0.0)
    f0, I0, E, G0 = cat.get('nh3_sjs')
    n_dvl = 2.0 / 3.0
    n_int = 3.0 / 2.0
    ITG = I0 * np.exp(-((1.0 / T) - (1.0 / T0)) * E * _HCK)
    out = []
    for f in np.asarray(freq, dtype=np.float64):
        f2 = f**2
        if f <= fLower:
            use = Spilker
        elif f >= fHigher:
            use = Joiner
        else:
            use = Interp
            flfh = (fLower - fHigher) / (f - fLower)
            for lst in (GH2, GHe, GNH3, ZH2, ZHe, ZNH3, C, D):
                lst[Interp] = lst[Spilker] + (lst[Spilker] - lst[Joiner]) / flfh
        delta = D[use] * P_nh3
        gamma = np.power((T0 / T), n_dvl) * (GH2[use] * P_h2 + GHe[use] * P_he + G0 * GNH3[use] * P_nh3)
        g2 = gamma**2
        zeta = np.power((T0 / T), n_dvl) * (ZH2[use] * P_h2 + ZHe[use] * P_he + G0 * ZNH3[use] * P_nh3)
        z2 = zeta**2
        num = (gamma - zeta) * f2 + (gamma + zeta) * (np.power(f0 + delta, 2.0) + g2 - z2)
        den = np.power((f2 - np.power(f0 + delta, 2.0) - g2 + z2), 2.0) + 4.0 * f2 * g2
        shape = _GHZ * 2.0 * np.power(f / f0, 2.0) * num / (np.pi * den)
        out.append(np.sum(shape * ITG))
    a = _COEF_GEISA * (P_nh3 / T0) * pow((T0 / T), n_int + 2) * np.array(out) * Pscale
    if units == 'dBperkm':
        a = a * OPTICALDEPTH_TO_DB
    return a


def _nh3_pblend(lowp, freq, T, P, X, P_dict, other_dict, **kwargs):
    """nh3_hs_sjs.py:6-26 / nh3_dbs_sjs.py:6-26 -- pressure switch 400..2000 bar."""
    PLower, PHigher = 400.0, 2000.0
    if P < PLower:
        return lowp(freq, T, P, X, P_dict, other_dict, **kwargs)
    if P > PHigher:
        return nh3_sjs(freq, T, P, X, P_dict, other_dict, **kwargs)
    a2 = nh3_sjs(freq, T, P, X, P_dict, other_dict, **kwargs)
    a1 = lowp(freq, T, P, X, P_dict, other_dict, **kwargs)
    W = (P - PLower) / (PHigher - PLower)
    return W * a2 + (1.0 - W) * a1


def nh3_sjsd(freq, T, P, X, P_dict, other_dict, **kwargs):
    """nh3_sjsd.py:6-24: sjs outside 10..100 bar, triangular blend with nh3_kd (peak at 35 bar) inside."""
    PLower, PMid, PHigher = 10.0, 35.0, 100.0
    if P < PLower or P > PHigher:
        return nh3_sjs(freq, T, P, X, P_dict, other_dict, **kwargs)
    a1 = np.array(nh3_sjs(freq, T, P, X, P_dict, other_dict, **kwargs))
    a2 = np.array(nh3_kd(freq, T, P, X, P_dict, other_dict, **kwargs))
    W = (P - PLower) / (PMid - PLower) if P < PMid else 1 - (P - PMid) / (PHigher - PMid)
    return W * a2 + (1.0 - W) * a1


def nh3_hs_sjs(freq, T, P, X, P_dict, other_dict, **kwargs):
    return _nh3_pblend(nh3_hs, freq, T, P, X, P_dict, other_dict, **kwargs)


def nh3_dbs_sjs(freq, T, P, X, P_dict, other_dict, **kwargs):
    return _nh3_pblend(nh3_dbs, freq, T, P, X, P_dict, other_dict, **kwargs)


# --------------------------------------------------------------------------------------
# Ben-Reuven family with the GEISA prefactor: H2S, PH3 (and the VVW branch of CO)
# --------------------------------------------------------------------------------------
def _ben_reuven_sum(freq, f0, gamma, zeta, delta, ITG):
    """h2s_ddb.py:73-79 == ph3_jh.py:95-101 == nh3_sjs.py:119-123 (loop over freq)."""
    g2 = gamma**2
    z2 = zeta**2
    out = []
    for f in np.asarray(freq, dtype=np.float64):
        f2 = f**2
        num = (gamma - zeta) * f2 + (gamma + zeta) * (np.power(f0 + delta, 2.0) + g2 - z2)
        den = np.power((f2 - np.power(f0 + delta, 2.0) - g2 + z2), 2.0) + 4.0 * f2 * g2
        shape = _GHZ * 2.0 * np.power(f / f0, 2.0) * num / (np.pi * den)
        out.append(np.sum(shape * ITG))
    return np.array(out)


def h2s_ddb(freq, T, P, X, P_dict, other_dict, **kwargs):
    """h2s_ddb.py:42-87."""
    units, cat, tstr, tfrq = _par(kwargs)
    T0 = 296.0
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_h2s = P * X[P_dict['H2S']]
    f0, I0, E, GH2S = cat.get('h2s', tstr, tfrq)
    delta = 1.28 * P_h2s
    gamma = np.power((T0 / T), 0.7) * (1.960 * P_h2 + 1.200 * P_he + GH2S * P_h2s)
    zeta = gamma
    ITG = I0 * np.exp(-((1.0 / T) - (1.0 / T0)) * E * _HCK)
    s = _ben_reuven_sum(freq, f0, gamma, zeta, delta, ITG)
    a = _COEF_GEISA * (P_h2s / T0) * pow((T0 / T), 3.0 / 2.0 + 2) * s
    if units == 'dBperkm':
        a = a * OPTICALDEPTH_TO_DB
    return a


def ph3_jh(freq, T, P, X, P_dict, other_dict, **kwargs):
    """ph3_jh.py:64-108."""
    units, cat, tstr, tfrq = _par(kwargs)
    T0 = 300.0
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_ph3 = P * X[P_dict['PH3']]
    f0, I0, E, WgtI0, WgtFGB, WgtSB = cat.get('ph3', tstr, tfrq)
    gamma = (np.power((T0 / T), 2.0 / 3.0) * (3.2930 * P_h2 + 1.6803 * P_he) * WgtFGB
             + np.power((T0 / T), 1.0) * 4.2157 * P_ph3 * WgtSB)
    ITG = I0 * WgtI0 * np.exp(-((1.0 / T) - (1.0 / T0)) * E * _HCK)
    s = _ben_reuven_sum(freq, f0, gamma, 0.0, 0.0, ITG)
    a = _COEF_GEISA * (P_ph3 / T0) * np.power((T0 / T), 3.0 / 2.0 + 2) * s
    if units == 'dBperkm':
        a = a * OPTICALDEPTH_TO_DB
    return a


# --------------------------------------------------------------------------------------
# CO (co_ddb.py:22-99) -- Voigt (complex rational) + VVW.  With the shipped default
# coshape='voigt' the reference sums the bare Voigt shape WITHOUT the line intensity ITG
# (co_ddb.py:86-87); restated as-is.
# --------------------------------------------------------------------------------------
_AVOIGT = [122.60793178, 214.38238869, 181.92853309, 93.15558046, 30.18014220,
           5.91262621, 0.56418958, 0.0]
_BVOIGT = [122.60793178, 352.73062511, 457.33447878, 348.70391772, 170.35400182,
           53.99290691, 10.47985711, 1.0]


def co_ddb(freq, T, P, X, P_dict, other_dict, **kwargs):
    units, cat, _, _ = _par(kwargs)
    T0 = 296.0
    PLimits = [0.001, 0.1]
    coshape = other_dict.get('coshape', 'voigt')
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_co = P * X[P_dict['CO']]
    f0, I0, E = cat.get('co')
    gamma = pow((T0 / T), 0.7) * (1.960 * P_h2 + 1.200 * P_he + 6.000 * P_co)
    g2 = gamma**2
    ITG = I0 * np.exp(-((1.0 / T) - (1.0 / T0)) * E * _HCK)
    w = min(max((P - PLimits[0]) / (PLimits[1] - PLimits[0]), 0.0), 1.0)
    out = []
    for f in np.asarray(freq, dtype=np.float64):
        f2 = f**2
        shape_Voigt = np.zeros(len(f0))
        if P <= PLimits[1] or coshape == 'voigt' or coshape == 'diff':
            betaD = 4.3e-7 * np.sqrt(T / 28.0) * f
            num = np.zeros(len(f0), dtype='complex128')
            den = np.zeros(len(f0), dtype='complex128')
            xi = gamma / betaD + (1.0j) * (f - f0) / betaD
            for j in range(len(_AVOIGT)):
                num += _AVOIGT[j] * (xi**j)
                den += _BVOIGT[j] * (xi**j)
            val = num / den
            shape_Voigt = _GHZ * (1.0 / (np.sqrt(np.pi) * betaD)) * val.real
        shape_VVW = np.zeros(len(f0))
        if P >= PLimits[0] or coshape == 'vvw' or coshape == 'diff':
            num = gamma * f2 + gamma * (np.power(f0, 2.0) + g2)
            den = np.power((f2 - np.power(f0, 2.0) - g2), 2.0) + 4.0 * f2 * g2
            shape_VVW = _GHZ * 2.0 * np.power(f / f0, 2.0) * num / (np.pi * den)
        shape = w * shape_VVW + (1.0 - w) * shape_Voigt
        if coshape == 'voigt':
            out.append(np.sum(shape_Voigt))
        elif coshape == 'vvw':
            out.append(np.sum(shape_VVW))
        elif coshape == 'diff':
            out.append(np.sum(shape_Voigt - shape_VVW))
        else:
            out.append(np.sum(shape * ITG))
    a = _COEF_GEISA * (P_co / T0) * pow((T0 / T), 3.0 / 2.0 + 2) * np.array(out)
    if units == 'dBperkm':
        a = a * OPTICALDEPTH_TO_DB
    return a


# --------------------------------------------------------------------------------------
# H2O: Karpowicz/Steffes (h2o_bk.py:65-127, 130-187)
# --------------------------------------------------------------------------------------
H2O_LINES = {
    # h2o_bk.py:23-49
    'f_o': np.array([22.2351, 183.3101, 321.2256, 325.1529, 380.1974, 439.1508, 443.0183,
                     448.0011, 470.8890, 474.6891, 488.4911, 556.9360, 620.7008, 752.0332,
                     916.1712]),
    'I_o': np.array([0.1314E-13, 0.2279E-11, 0.8058E-13, 0.2701E-11, 0.2444E-10, 0.2185E-11,
                     0.4637E-12, 0.2568E-10, 0.8392E-12, 0.3272E-11, 0.6676E-12, 0.1535E-08,
                     0.1711E-10, 0.1014E-08, 0.4238E-10]),
    'E_o': np.array([2.144, 0.668, 6.179, 1.541, 1.048, 3.595, 5.048, 1.405, 3.597, 2.379,
                     2.852, 0.159, 2.391, 0.396, 1.441]),
    'w_s': np.array([0.01349, 0.01466, 0.01057, 0.01381, 0.01454, 0.009715, 0.00788,
                     0.01275, 0.00983, 0.01095, 0.01313, 0.01405, 0.011836, 0.01253,
                     0.01275]) / 0.001,
    'x_s': np.array([0.61, 0.85, 0.54, 0.74, 0.89, 0.62, 0.50, 0.67, 0.65, 0.64, 0.72, 1.0,
                     0.68, 0.84, 0.78]),
    'w_h2': np.array([2.395, 2.4000, 2.395, 2.395, 2.390, 2.395, 2.395, 2.395, 2.395, 2.395,
                      2.395, 2.395, 2.395, 2.395, 2.395]),
    'w_he': np.array([0.67, 0.71, 0.67, 0.67, 0.63, 0.67, 0.67, 0.67, 0.67, 0.67, 0.67, 0.67,
                      0.67, 0.67, 0.67]),
    'x_h2': np.array([0.900, 0.950, 0.900, 0.900, 0.850, 0.900, 0.900, 0.900, 0.900, 0.900,
                      0.900, 0.900, 0.900, 0.900, 0.900]),
    'x_he': np.array([0.515, 0.490, 0.515, 0.490, 0.540, 0.515, 0.515, 0.515, 0.515, 0.515,
                      0.515, 0.515, 0.515, 0.515, 0.515]),
}


def h2o_lines(truncate_strength=None, truncate_freq=None):
    """h2o_bk.py:51-62 (truthiness test: 0/None disable)."""
    d = {k: v.copy() for k, v in H2O_LINES.items()}
    if truncate_strength:
        use = d['I_o'] > truncate_strength
        d = {k: v[use] for k, v in d.items()}
    if truncate_freq:
        use = d['f_o'] < truncate_freq
        d = {k: v[use] for k, v in d.items()}
    return d


def h2o_bk(freq, T, P, X, P_dict, other_dict, **kwargs):
    units, _, tstr, tfrq = _par(kwargs)
    d = h2o_lines(tstr, tfrq)
    mbars_to_bars = 0.001
    inv_km_to_dB = 4.342945
    convert_to_km = 1e-4
    To = 300.0
    NA = 6.0221415e23
    M_amu = 8.314472 / 0.46151805
    isotope_partition = 0.997317
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_h2o = P * X[P_dict['H2O']]
    Theta = To / T
    density_h2o = (M_amu * P_h2o) / (8.314472e-5 * T)
    density_h2o = isotope_partition * (density_h2o / M_amu) * NA * (1.0 / 1e6)
    expo = d['E_o'] * (1.0 - Theta)
    S = d['I_o'] * (Theta**2.5) * np.exp(expo)
    df = (d['w_s'] * P_h2o * np.power(Theta, d['x_s'])
          + d['w_h2'] * P_h2 * np.power(Theta, d['x_h2'])
          + d['w_he'] * P_he * np.power(Theta, d['x_he']))
    # vvwlinecontribution_modified, h2o_bk.py:130-187 (shift SR == 0)
    f = np.asarray(freq, dtype=np.float64)[None, :]
    fo = d['f_o'][:, None]
    dfm = df[:, None]
    base = (df / (562500.0 + df**2))[:, None]
    A = np.square(f / fo) / np.pi
    B = dfm / (np.square(f - fo - 0.0) + np.square(dfm))
    Cc = dfm / (np.square(f + fo + 0.0) + np.square(dfm))
    F = A * (B - base + Cc - base)
    FSsum = np.sum(S[:, None] * F, 0)
    line_contribution = inv_km_to_dB * convert_to_km * density_h2o * FSsum
    Cf_he = ((1.0 / mbars_to_bars)**2) * 1.03562010226e-10
    Cf_h2 = ((1.0 / mbars_to_bars)**2) * 5.07722009423e-11
    Cs1 = 3.1e-07 * pow(Theta, 12.0)
    Cs2 = 0.0
    farr = np.asarray(freq, dtype=np.float64)
    Foreign_he = Cf_he * P_he * P_h2o * (farr**2) * pow(Theta, 3.0)
    Foreign_h2 = Cf_h2 * P_h2 * P_h2o * (farr**2) * pow(Theta, 3.0)
    Foreign = Foreign_he + Foreign_h2
    Self = Cs1 * ((P_h2o / mbars_to_bars)**2) * (farr**2.0) + Cs2 * (farr**2.0)
    a = line_contribution + inv_km_to_dB * Foreign + inv_km_to_dB * Self
    if units != 'dBperkm':
        a = a / 434294.5
    return np.asarray(a).flatten()


# --------------------------------------------------------------------------------------
# H2 collision-induced absorption (h2_jj_ddb.py:7-38, h2_jj.py:7-22)
# --------------------------------------------------------------------------------------
def _h2_core(freq, T, P, X, P_dict, pre, units):
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_ch4 = P * X[P_dict['CH4']]
    th = 273.0 / T
    out = []
    for f in np.asarray(freq, dtype=np.float64):
        cf = 3.9522E-14 * f**2 * P_h2 * pre
        a = cf * (P_h2 * pow(th, 3.12) + 1.382 * P_he * pow(th, 2.24) + 9.322 * P_ch4 * pow(th, 3.34))
        if units == 'dBperkm':
            a *= 434294.5
        out.append(a)
    return np.array(out)


def h2_jj_ddb(freq, T, P, X, P_dict, other_dict, **kwargs):
    units = kwargs.get('units', 'dBperkm')
    state = other_dict['h2state']
    if state == 'e':
        pre = min((T / 55.0)**2.7, 1.0)
        pre *= (T / 120.0)**0.55
        pre = min(pre, 1.0)
    elif state == 'n':
        pre = min((T / 40.0)**2.5, 1.0)
    else:
        return 0.0                      # 'INVALID H2STATE' (h2_jj_ddb.py:28-30)
    return _h2_core(freq, T, P, X, P_dict, pre, units)


def h2_jj(freq, T, P, X, P_dict, other_dict, **kwargs):
    return _h2_core(freq, T, P, X, P_dict, 1.0, kwargs.get('units', 'dBperkm'))


# --------------------------------------------------------------------------------------
# H2 CIA from Orton's tables (h2_orton.py:17-222)
# --------------------------------------------------------------------------------------
_ORTON_PATH = os.path.join(os.path.dirname(_DEFAULT_LINECAT), 'orton_h2.npz')
_ORTON = {}
_ORTON_TABLES = {'eh2h2': 0, 'nh2h2': 1, 'eh2he': 2, 'nh2he': 3, 'eh2ch4': 4, 'nh2ch4': 5}   # h2_orton.py:12
_ORTON_STATES = {'e': 0, 'n': 1}                                                               # h2_orton.py:13


def orton_tables(freqs):
    """readInputFiles (h2_orton.py:17-123): the tabulated temperatures and, per table and frequency, the
    absorption coefficients at those temperatures after the piece-wise quadratic interpolation in frequency
    through the three tabulated wavenumbers around f (extrapolated below the first one)."""
    if 'raw' not in _ORTON:
        d = np.load(_ORTON_PATH)
        _ORTON['raw'] = (int(d['ntemp']), float(d['tmax']), float(d['tmin']), np.array(d['wavenumber']), np.array(d['logtab']))
    nTemp, Tmax, Tmin, wn, logtab = _ORTON['raw']
    import math
    lTmx, lTmn = math.log(Tmax), math.log(Tmin)
    dlT = (lTmx - lTmn) / (nTemp - 1.0)
    ta = [lTmn]
    for i in range(nTemp - 1):                     # h2_orton.py:35-37: accumulated, not i * dlT
        ta.append(ta[i] + dlT)
    Ttab = np.array([math.exp(v) for v in ta])
    ftab = np.array([v * 29.9792458 for v in wn])  # h2_orton.py:49-54
    h2vab = np.zeros((len(_ORTON_TABLES), len(freqs), nTemp))
    for ii in range(len(_ORTON_TABLES)):
        for jj, f in enumerate(freqs):
            ifreq = int(np.where(ftab > f)[0][0])
            if ifreq == 0:
                ifreq = 1
            v1, v2, v3 = logtab[ii, ifreq - 1], logtab[ii, ifreq], logtab[ii, ifreq + 1]   # h2_orton.py:87-93
            X1 = ftab[ifreq - 1]
            X21 = ftab[ifreq] - ftab[ifreq - 1]
            X32 = ftab[ifreq + 1] - ftab[ifreq]
            X212 = ftab[ifreq]**2 - ftab[ifreq - 1]**2
            X322 = ftab[ifreq + 1]**2 - ftab[ifreq]**2
            for ll in range(nTemp):
                Y1 = math.exp(float(v1[ll]))
                Y21 = (math.exp(float(v2[ll])) - Y1)
                Y32 = (math.exp(float(v3[ll])) - math.exp(float(v2[ll])))
                DQ = X212 * X32 - X322 * X21
                AQ = (X32 * Y21 - X21 * Y32) / DQ
                BQ = (X212 * Y32 - X322 * Y21) / DQ
                CQ = Y1 - AQ * X1**2 - BQ * X1
                h2vab[ii, jj, ll] = AQ * f**2 + BQ * f + CQ
    return Ttab, h2vab


def h2_orton(freq, T, P, X, P_dict, other_dict, **kwargs):
    """h2_orton.py:126-222.  The text file of the reference prints 4-5 significant digits of log(alpha); the npz
    repack (tools/build_orton.py) holds the same numbers."""
    from scipy.interpolate import interp1d
    units = kwargs.get('units', 'dBperkm')
    T0, atm2bar = 273.0, 1.01325
    freq = [float(f) for f in freq]
    key = tuple(freq)
    if _ORTON.get('key') != key:
        _ORTON['key'] = key
        _ORTON['tab'] = orton_tables(freq)
    Ttab, h2vab = _ORTON['tab']
    st = _ORTON_STATES[other_dict['h2state'].lower()]
    xh2, xhe, xch4 = _ORTON_TABLES['eh2h2'] + st, _ORTON_TABLES['eh2he'] + st, _ORTON_TABLES['eh2ch4'] + st
    P_h2 = P * X[P_dict['H2']]
    P_he = P * X[P_dict['HE']]
    P_ch4 = P * X[P_dict['CH4']]
    out = []
    if T < Ttab[0]:                                 # quartic extrapolation below the table (h2_orton.py:150-178)
        nexp = 4.0
        X1 = Ttab[0]
        X21n = Ttab[1]**nexp - Ttab[0]**nexp
        for ii in range(len(freq)):
            v = []
            for jj in (0, 1):
                Tjj = Ttab[jj]
                a = ((P_h2 / atm2bar) * (h2vab[xh2, ii, jj] * P_h2 / atm2bar + h2vab[xhe, ii, jj] * P_he / atm2bar +
                                         h2vab[xch4, ii, jj] * P_ch4 / atm2bar) * (T0 / Tjj)**2)
                v.append(a)
            Y1 = v[0]
            Y21 = v[1] - v[0]
            AQ = -1.0 * Y21 / X21n
            CQ = Y1 + AQ * (X1**nexp)
            out.append(CQ - AQ * (T**nexp))
    elif T > Ttab[-1]:                              # h2_jj scaled to the table at its last temperature (:179-199)
        Tnear = Ttab[-1]
        jjnear = h2_jj(freq, Tnear, P, X, P_dict, other_dict)
        jj = h2_jj(freq, T, P, X, P_dict, other_dict)
        for ii in range(len(freq)):
            anear = ((P_h2 / atm2bar) * (h2vab[xh2, ii, -1] * P_h2 / atm2bar + h2vab[xhe, ii, -1] * P_he / atm2bar +
                                         h2vab[xch4, ii, -1] * P_ch4 / atm2bar) * (T0 / Tnear)**2)
            out.append(jj[ii] * (anear / jjnear[ii]))
    else:                                           # cubic spline in T through the 10 table points (:200-214)
        for ii in range(len(freq)):
            ah2 = interp1d(Ttab, h2vab[xh2, ii], kind='cubic')(T)
            ahe = interp1d(Ttab, h2vab[xhe, ii], kind='cubic')(T)
            ach4 = interp1d(Ttab, h2vab[xch4, ii], kind='cubic')(T)
            out.append((P_h2 / atm2bar) * (ah2 * P_h2 / atm2bar + ahe * P_he / atm2bar + ach4 * P_ch4 / atm2bar) * (T0 / T)**2)
    out = np.array(out, dtype=np.float64)
    if units == 'dBperkm':
        out = out * 434294.5
    return out


# --------------------------------------------------------------------------------------
# Clouds (clouds_idp.py:6-101)
# --------------------------------------------------------------------------------------
_FR = [1.0E8, 3.0E8, 1.0E9, 2.0E9, 3.0E9, 5.0E9, 1.0E10, 3.0E10, 1.0E11]
_EIMAG = [8.0E-3, 1.5E-3, 8.0E-4, 1.0E-3, 1.2E-3, 1.5E-3, 3.0E-3, 8.0E-3, 2.0E-2]


def _water_eps(freq, T):
    """clouds_idp.py:72-101."""
    Tc = T - 273.0
    fHz = freq * 1.0E9
    if Tc >= 0.0:
        RelT = 1.1109E-10 - Tc * 3.824E-12 + (Tc**2) * 6.938E-14 - (Tc**3) * 5.096E-16
        E0 = 88.045 - 0.4147 * Tc + (Tc**2) * 6.295E-4 + (Tc**3) * 1.075E-5
        if E0 < 0.0:
            E0 = 0.0
        EINF = 4.9
        E1 = EINF + (E0 - EINF) / (1.0 + (fHz * RelT)**2)
        E2 = fHz * RelT * (E0 - EINF) / (1.0 + (fHz * RelT)**2)
        if E2 < 0.0:
            E2 = 0.0
    else:
        E1 = 3.17
        LF = np.log10(fHz)
        for j in range(len(_FR) - 1):
            if _FR[j + 1] >= fHz:
                break
        LF0 = np.log10(_FR[j])
        LF1 = np.log10(_FR[j + 1])
        DLF = (LF - LF0) / (LF1 - LF0)
        X0 = np.log10(_EIMAG[j])
        X1 = np.log10(_EIMAG[j + 1])
        E2 = 10.0**(X0 + DLF * (X1 - X0))
    return E1 - E2 * 1.0j


def _acloud(k, fraction, e):
    K = (e - 1.0) / (e + 2.0)
    return 3.0 * k * fraction * (-K.imag)


def clouds_idp(freq, T, P, cloud, cloud_dict, other_dict, **kwargs):
    units = kwargs.get('units', 'dBperkm')
    out = []
    for f in np.asarray(freq, dtype=np.float64):
        a = 0.0
        k = (2.0 * np.pi * f / _GHZ)
        if other_dict.get('ice_p', 0.0) > 0.0:
            a += _acloud(k, cloud[cloud_dict['H2O']] / 0.9, _water_eps(f, T))
        if other_dict.get('water_p', 0.0) > 0.0:
            a += _acloud(k, cloud[cloud_dict['SOLN']] / 1.0, _water_eps(f, T))
        if other_dict.get('nh4sh_p', 0.0) > 0.0:
            a += _acloud(k, cloud[cloud_dict['NH4SH']] / 1.2, (1.7 - 0.005j)**2)
        if other_dict.get('nh3ice_p', 0.0) > 0.0:
            a += _acloud(k, cloud[cloud_dict['NH3']] / 1.6, (1.3 - 0.0001j)**2)
        if other_dict.get('h2sice_p', 0.0) > 0.0:
            a += _acloud(k, cloud[cloud_dict['H2S']] / 1.5, (1.15 - 0.0001j)**2)
        if other_dict.get('ch4', 0.0) > 0.0:      # key 'ch4' (not 'ch4_p'): clouds_idp.py:45
            a += _acloud(k, cloud[cloud_dict['CH4']] / 1.0, (1.3 - 0.00001j)**2)
        if a < 0.0:
            a = 0.0
        if units == 'dBperkm':
            a *= 434294.5
        out.append(a)
    return np.array(out)


FORMALISMS = {
    'nh3_hs': nh3_hs, 'nh3_dbs': nh3_dbs, 'nh3_sjs': nh3_sjs, 'nh3_kd': nh3_kd, 'nh3_bg': nh3_bg, 'nh3_sjsd': nh3_sjsd,
    'nh3_hs_sjs': nh3_hs_sjs, 'nh3_dbs_sjs': nh3_dbs_sjs,
    'h2s_ddb': h2s_ddb, 'ph3_jh': ph3_jh, 'co_ddb': co_ddb, 'h2o_bk': h2o_bk,
    'h2_jj_ddb': h2_jj_ddb, 'h2_jj': h2_jj, 'h2_orton': h2_orton, 'clouds_idp': clouds_idp,
}


# --------------------------------------------------------------------------------------
# Alpha.get_layers driver (alpha.py:151-305)
# --------------------------------------------------------------------------------------
def get_layers(freqs, gas, cloud, C, Cl, constituent_alpha, other_dicts=None, scale=False,
               truncate_strength=None, truncate_freq=None, units='invcm', cat=None,
               return_per_constituent=False, layers=None):
    """Restates Alpha.get_layers / get_single_layer / get_alpha_from_calc / total_layer_alpha.

    constituent_alpha: {constituent: formalism name or None}; called in sorted(constituent)
    order (alpha.py:83, 202).  Returns layers[F, L] (alpha.py:300) and optionally the
    per-constituent cube [L, F, C] (alpha.py:110-131, after scaling).
    """
    cat = cat or default_catalog()
    freqs = np.asarray(freqs, dtype=np.float64)
    used = {c: f for c, f in constituent_alpha.items() if f is not None}
    ordered = sorted(used.keys())
    other_dicts = other_dicts or {}
    truncate_strength = truncate_strength or {}
    truncate_freq = truncate_freq or {}
    L = gas.shape[1]
    idx = range(L) if layers is None else layers
    lscale = layer_scale(scale, L, ordered)
    total = np.zeros((len(freqs), len(idx)))
    cube = np.zeros((len(idx), len(freqs), len(ordered)))
    for n, layer in enumerate(idx):
        P = gas[C['P']][layer]
        T = gas[C['T']][layer]
        absorb = []
        for c in ordered:
            if c.lower().startswith('cloud'):
                X, D = cloud[:, layer], Cl
            else:
                X, D = gas[:, layer], C
            a = FORMALISMS[used[c]](freqs, T, P, X, D, other_dicts.get(c, {}),
                                    truncate_freq=truncate_freq.get(c), units=units, cat=cat,
                                    truncate_strength=truncate_strength.get(c))
            absorb.append(np.asarray(a, dtype=np.float64) * np.ones(len(freqs)))
        absorb = np.array(absorb).transpose()           # [F, C]
        ls = lscale[layer]
        if isinstance(ls, dict):
            for j, c in enumerate(ordered):
                if c in ls:
                    absorb[:, j] = absorb[:, j] * ls[c]
            tot = np.zeros(len(freqs))
            for j in range(len(ordered)):               # left-to-right sum (alpha.py:177-185)
                tot = tot + absorb[:, j]
        else:
            tot = absorb.sum(axis=1) * float(ls)
            absorb = absorb * float(ls)
        total[:, n] = tot
        cube[n] = absorb
    if return_per_constituent:
        return total, cube, ordered
    return total


def layer_scale(scale, N, ordered):
    """alpha.py:235-259."""
    if isinstance(scale, dict):
        for k, v in scale.items():
            if k not in ordered:
                raise ValueError("{} not found as constituent for alpha".format(k))
            if len(v) != N:
                raise ValueError("Incorrect scale for {}:  N {} vs {}".format(k, len(v), N))
        return [{k: v[i] for k, v in scale.items()} for i in range(N)]
    if isinstance(scale, (list, np.ndarray)):
        if len(scale) != N:
            raise ValueError("Incorrect number of scale layers.")
        return scale
    if isinstance(scale, bool) or isinstance(scale, (dict, list)):
        return [1.0] * N
    try:
        return [float(scale)] * N
    except (TypeError, ValueError):
        return [1.0] * N
