"""nh3_kd -- NH3 consistent model with pressure-switched inversion constants, linear between 12 and 20 bar (reference nh3/nh3_kd.py:115-351).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('nh3', 'nh3_kd')
