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
    'h2_jj_ddb': [], 'h2_jj': [], 'clouds_idp': [],
}


def upload(ctx, formalism, truncate_strength=None, truncate_freq=None):
    """Make sure the catalogs `formalism` needs are resident on the context's GPU."""
    for name in FORMALISM_CATALOGS[formalism]:
        key = (truncate_strength, truncate_freq) if name in ('h2s', 'ph3', 'h2o') else ()
        if ctx.catalog_key.get(name, '__unset__') == key:
            continue
        ctx.set_catalog(name, table(name, truncate_strength, truncate_freq), key=key)
