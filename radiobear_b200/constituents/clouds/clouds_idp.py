"""clouds_idp -- Rayleigh cloud absorption, Debye water / tabulated ice permittivity (reference clouds/clouds_idp.py:6-101).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('clouds', 'clouds_idp')
