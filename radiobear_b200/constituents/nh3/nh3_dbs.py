"""nh3_dbs -- NH3, Devaraj/Bellotti/Steffes constants of the consistent model (reference nh3/nh3_dbs.py:70-313).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('nh3', 'nh3_dbs')
