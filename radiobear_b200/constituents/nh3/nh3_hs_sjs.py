"""nh3_hs_sjs -- NH3 pressure switch: nh3_hs below 400 bar, nh3_sjs above 2000 bar, linear blend between (reference nh3/nh3_hs_sjs.py:6-26).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('nh3', 'nh3_hs_sjs')
