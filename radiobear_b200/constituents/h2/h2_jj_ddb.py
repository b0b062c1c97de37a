"""h2_jj_ddb -- H2-H2/He/CH4 collision-induced absorption with the equilibrium/normal pre-factor (reference h2/h2_jj_ddb.py:7-38).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('h2', 'h2_jj_ddb')
