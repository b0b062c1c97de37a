"""Constituent absorption plugins: constituents/<gas>/<formalism>.py :: alpha(freq, T, P, X, P_dict, other_dict, **kwargs).

Same module names and call signature as the reference's plugins (alpha.py:59-68, 210-213), so
`config.par` lines such as `alpha nh3:nh3_hs_sjs h2s:h2s_ddb` keep their meaning.  Every module is a
thin shim over one C-ABI call (rb_alpha_layers with n_layers = 1); the arithmetic is in
csrc/alpha_kernels.cu.
"""
from . import parameters  # noqa
