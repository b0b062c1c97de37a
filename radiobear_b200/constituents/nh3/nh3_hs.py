"""nh3_hs -- NH3, Hanley/Steffes consistent model: Ben-Reuven inversion lines + Gross rotational and v2 lines (reference nh3/nh3_hs.py:70-312).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('nh3', 'nh3_hs')
