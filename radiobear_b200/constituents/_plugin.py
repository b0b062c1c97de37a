"""Shared body of the plugin shims."""
import numpy as np

from .. import engine
from . import parameters

NH3_F_SPLIT = 30.0
_REORDER = ('nh3_hs', 'nh3_dbs')                 # concatenate lo (<=30 GHz) then hi (nh3_hs.py:70-88)
_REORDER_BELOW_400 = ('nh3_hs_sjs', 'nh3_dbs_sjs')


def make_alpha(gas, formalism):
    """Build `alpha(freq, T, P, X, P_dict, other_dict, **kwargs)` for one formalism.

    kwargs (parameters.py:4-8 + alpha.py:210-213): units ('dBperkm' default | 'invcm'),
    truncate_strength, truncate_freq, path (ignored: catalogs ship with the package), verbose.
    Returns ndarray[len(freq)].
    """
    is_cloud = gas.startswith('cloud')

    def alpha(freq, T, P, X, P_dict, other_dict, **kwargs):
        par = parameters.setpar(kwargs)
        f = np.atleast_1d(np.asarray(freq, dtype=np.float64))
        X = np.asarray(X, dtype=np.float64)
        common = dict(formalisms=[(gas, formalism)], other_dicts={gas: dict(other_dict or {})}, units=par.units,
                      truncate_strength={gas: getattr(par, 'truncate_strength', None)},
                      truncate_freq={gas: getattr(par, 'truncate_freq', None)})
        if is_cloud:
            dummy = np.zeros((1, 1))
            out = engine.alpha_layers(f, [T], [P], dummy, {}, cloud=X, cloud_dict=P_dict, **common)[0]
        else:
            out = engine.alpha_layers(f, [T], [P], X, P_dict, **common)[0]
        if formalism in _REORDER or (formalism in _REORDER_BELOW_400 and P < 400.0):
            lo = f <= NH3_F_SPLIT
            out = np.concatenate((out[lo], out[~lo]))
        return out

    alpha.__name__ = 'alpha'
    alpha.__doc__ = 'Absorption of {} with formalism {} (see module docstring).'.format(gas, formalism)
    return alpha
