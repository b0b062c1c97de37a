"""Line catalogs for the absorption kernels.

Data: radiobear_b200/data/linecat.npz, repacked (values unchanged) from the reference's npz
catalogs by tools/build_linecat.py.  Truncation follows the reference's load-time rules:
h2s_ddb.py:23-38, ph3_jh.py:28-61 (`I0 > truncate_strength`, `f0 < truncate_freq`),
h2o_bk.py:51-62 (truthiness test).  nh3 / co catalogs are never truncated by the reference.
"""
import os

import numpy as np

LINECAT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'linecat.npz')

# The 15 H2O lines of the Karpowicz/Steffes model are constants of the formalism (h2o_bk.py:23-49).
# rows: f_o, I_o, E_o, w_s (already /mbars_to_bars), x_s, w_h2, w_he, x_h2, x_he
_H2O = np.array([
    [22.2351, 183.3101, 321.2256, 325.1529, 380.1974, 439.1508, 443.0183, 448.0011, 470.8890, 474.6891,
     488.4911, 556.9360, 620.7008, 752.0332, 916.1712],
    [0.1314E-13, 0.2279E-11, 0.8058E-13, 0.2701E-11, 0.2444E-10, 0.2185E-11, 0.4637E-12, 0.2568E-10,
     0.8392E-12, 0.3272E-11, 0.6676E-12, 0.1535E-08, 0.1711E-10, 0.1014E-08, 0.4238E-10],
    [2.144, 0.668, 6.179, 1.541, 1.048, 3.595, 5.048, 1.405, 3.597, 2.379, 2.852, 0.159, 2.391, 0.396, 1.441],
    [0.01349, 0.01466, 0.01057, 0.01381, 0.01454, 0.009715, 0.00788, 0.01275, 0.00983, 0.01095, 0.01313,
     0.01405, 0.011836, 0.01253, 0.01275],
    [0.61, 0.85, 0.54, 0.74, 0.89, 0.62, 0.50, 0.67, 0.65, 0.64, 0.72, 1.0, 0.68, 0.84, 0.78],
    [2.395, 2.4000, 2.395, 2.395, 2.390, 2.395, 2.395, 2.395, 2.395, 2.395, 2.395, 2.395, 2.395, 2.395, 2.395],
    [0.67, 0.71, 0.67, 0.67, 0.63, 0.67, 0.67, 0.67, 0.67, 0.67, 0.67, 0.67, 0.67, 0.67, 0.67],
    [0.900, 0.950, 0.900, 0.900, 0.850, 0.900, 0.900, 0.900, 0.900, 0.900, 0.900, 0.900, 0.900, 0.900, 0.900],
    [0.515, 0.490, 0.515, 0.490, 0.540, 0.515, 0.515, 0.515, 0.515, 0.515, 0.515, 0.515, 0.515, 0.515, 0.515]])
_H2O[3] = _H2O[3] / 0.001

_raw = None
# BASELINE config C5 / SURVEY 8d "full NH3 catalog": the untrimmed rotational (1301) and roto-vibrational (4198) line
# lists instead of the 201 / 198 lines the reference ships in ammonia.npz (constituents/txt2npz.py:21-56)
_full_nh3 = False


def use_full_nh3_catalog(flag=True):
    """Switch nh3_hs / nh3_dbs / nh3_kd (and the *_sjs blends) to the 5914-line NH3 catalog (off by default)."""
    global _full_nh3
    _full_nh3 = bool(flag)


def raw():
    global _raw
    if _raw is None:
        d = np.load(LINECAT_PATH)
        _raw = {k: np.array(d[k], dtype=np.float64) for k in d.files if not k.endswith('_cols')}
        _raw['h2o'] = _H2O.copy()
    return _raw


def table(name, truncate_strength=None, truncate_freq=None):
    """[ncols][nlines] float64 table for kernel catalog `name` after truncation."""
    r = raw()
    if name in ('nh3_rot', 'nh3_v2') and _full_nh3:
        return r[name + '_full']
    if name in ('nh3_inv', 'nh3_rot', 'nh3_v2', 'nh3_sjs', 'co'):
        return r[name]
    if name == 'h2s':
        a = r['h2s']
    elif name == 'ph3':
        a = np.vstack([r['ph3'], r['ph3_wgt']])
    elif name == 'h2o':
        a = r['h2o']
        # h2o_bk.py:51-62 tests truthiness: None and 0 both disable
        truncate_strength = truncate_strength or None
        truncate_freq = truncate_freq or None
    else:
        raise KeyError(name)
    if truncate_strength is not None:
        a = a[:, a[1] > truncate_strength]
    if truncate_freq is not None:
        a = a[:, a[0] < truncate_freq]
    return np.ascontiguousarray(a)


# which kernel catalogs a formalism reads, and which gas's truncation settings apply
FORMALISM_CATALOGS = {
    'nh3_hs': ['nh3_inv', 'nh3_rot', 'nh3_v2'], 'nh3_dbs': ['nh3_inv', 'nh3_rot', 'nh3_v2'],
    'nh3_sjs': ['nh3_sjs'], 'nh3_bg': ['nh3_sjs'], 'nh3_kd': ['nh3_inv', 'nh3_rot', 'nh3_v2'],
    'nh3_sjsd': ['nh3_inv', 'nh3_rot', 'nh3_v2', 'nh3_sjs'],
    'nh3_hs_sjs': ['nh3_inv', 'nh3_rot', 'nh3_v2', 'nh3_sjs'],
    'nh3_dbs_sjs': ['nh3_inv', 'nh3_rot', 'nh3_v2', 'nh3_sjs'],
    'h2s_ddb': ['h2s'], 'ph3_jh': ['ph3'], 'h2o_bk': ['h2o'], 'co_ddb': ['co'],
    'h2_jj_ddb': [], 'h2_jj': [], 'clouds_idp': [], 'h2_orton': ['h2_orton'],
}


# ---- Orton H2 CIA tables (h2_orton.py:17-123) ---------------------------------------------------
ORTON_PATH = os.path.join(os.path.dirname(LINECAT_PATH), 'orton_h2.npz')
ORTON_TABLES = {'eh2h2': 0, 'nh2h2': 1, 'eh2he': 2, 'nh2he': 3, 'eh2ch4': 4, 'nh2ch4': 5}     # h2_orton.py:12
_orton_raw = None


def _orton():
    global _orton_raw
    if _orton_raw is None:
        d = np.load(ORTON_PATH)
        nT, Tmax, Tmin = int(d['ntemp']), float(d['tmax']), float(d['tmin'])
        # h2_orton.py:28-42: log-spaced temperatures, accumulated step by step like the reference
        import math
        dlT = (math.log(Tmax) - math.log(Tmin)) / (nT - 1.0)
        ta = [math.log(Tmin)]
        for i in range(nT - 1):
            ta.append(ta[i] + dlT)
        _orton_raw = (np.array([math.exp(v) for v in ta]), np.array([v * 29.9792458 for v in d['wavenumber']]),
                      np.array(d['logtab']))
    return _orton_raw


def notaknot_spline(x, Y):
    """Not-a-knot cubic spline through (x[n], Y[n][m]) -- what scipy's interp1d(kind='cubic') builds.
    Returns c1, c2, c3 [n-1][m] of s(t) = Y[i] + c1 dt + c2 dt^2 + c3 dt^3 on [x[i], x[i+1]], dt = t - x[i]."""
    x = np.asarray(x, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    n = len(x)
    h = np.diff(x)
    d = np.diff(Y, axis=0) / h[:, None]
    A = np.zeros((n, n))
    rhs = np.zeros((n, Y.shape[1]))
    for i in range(1, n - 1):                       # continuity of s'' (unknowns: M = s'' at the knots)
        A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1], 2.0 * (h[i - 1] + h[i]), h[i]
        rhs[i] = 6.0 * (d[i] - d[i - 1])
    A[0, 0], A[0, 1], A[0, 2] = h[1], -(h[0] + h[1]), h[0]                      # s''' continuous at x[1]
    A[-1, -3], A[-1, -2], A[-1, -1] = h[-1], -(h[-2] + h[-1]), h[-2]            # ... and at x[n-2]
    M = np.linalg.solve(A, rhs)
    c1 = d - h[:, None] * (2.0 * M[:-1] + M[1:]) / 6.0
    c2 = M[:-1] / 2.0
    c3 = (M[1:] - M[:-1]) / (6.0 * h[:, None])
    return c1, c2, c3


def orton_table(freqs, h2state):
    """RB_CAT_H2_ORTON for the frequencies of a call: [121][F] (layout in include/radiobear_b200.h).

    Frequency: the quadratic through the three tabulated points around f, extrapolated below the first
    one (h2_orton.py:77-108, `ifreq = first ftab > f`, 0 -> 1).  Temperature: spline coefficients for the kernel."""
    Ttab, ftab, logtab = _orton()
    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    st = {'e': 0, 'n': 1}[str(h2state).lower()]
    if np.any(freqs >= ftab[-2]):
        raise ValueError('h2_orton: frequency beyond the tabulated range ({:.0f} GHz)'.format(ftab[-2]))
    ifreq = np.maximum(np.searchsorted(ftab, freqs, side='right'), 1)          # first ftab > f
    nT, F = len(Ttab), len(freqs)
    out = np.empty((nT + 3 * (nT + 3 * (nT - 1)), F))
    out[:nT] = Ttab[:, None]
    X1, X2, X3 = ftab[ifreq - 1], ftab[ifreq], ftab[ifreq + 1]
    X21, X32 = X2 - X1, X3 - X2
    X212, X322 = X2**2 - X1**2, X3**2 - X2**2
    DQ = X212 * X32 - X322 * X21
    for t, name in enumerate(('eh2h2', 'eh2he', 'eh2ch4')):
        tab = logtab[ORTON_TABLES[name] + st]
        E1, E2, E3 = np.exp(tab[ifreq - 1]), np.exp(tab[ifreq]), np.exp(tab[ifreq + 1])     # [F][nT]
        Y1, Y21, Y32 = E1, E2 - E1, E3 - E2
        AQ = (X32[:, None] * Y21 - X21[:, None] * Y32) / DQ[:, None]
        BQ = (X212[:, None] * Y32 - X322[:, None] * Y21) / DQ[:, None]
        CQ = Y1 - AQ * X1[:, None]**2 - BQ * X1[:, None]
        v = (AQ * freqs[:, None]**2 + BQ * freqs[:, None] + CQ).T               # [nT][F]
        c1, c2, c3 = notaknot_spline(Ttab, v)
        base = nT + t * (nT + 3 * (nT - 1))
        out[base:base + nT] = v
        coef = np.stack([c1, c2, c3], axis=1).reshape(3 * (nT - 1), F)          # interval-major: k*3 + m
        out[base + nT:base + nT + 3 * (nT - 1)] = coef
    return out


def upload(ctx, formalism, truncate_strength=None, truncate_freq=None, freqs=None, other=None):
    """Make sure the catalogs `formalism` needs are resident on the context's GPU."""
    if formalism == 'h2_orton':
        st = str((other or {}).get('h2state', 'e')).lower()
        if st not in ('e', 'n'):
            raise ValueError('INVALID H2STATE {!r}'.format(st))
        f = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
        key = (st, f.tobytes())
        if ctx.catalog_key.get('h2_orton', '__unset__') != key:
            ctx.set_catalog('h2_orton', orton_table(f, st), key=key)
        return
    for name in FORMALISM_CATALOGS[formalism]:
        key = (truncate_strength, truncate_freq) if name in ('h2s', 'ph3', 'h2o') else \
            (('full',) if (_full_nh3 and name in ('nh3_rot', 'nh3_v2')) else ())
        if ctx.catalog_key.get(name, '__unset__') == key:
            continue
        ctx.set_catalog(name, table(name, truncate_strength, truncate_freq), key=key)
