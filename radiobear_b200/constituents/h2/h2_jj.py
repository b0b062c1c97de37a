"""h2_jj -- H2-H2/He/CH4 collision-induced absorption (reference h2/h2_jj.py:7-22).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('h2', 'h2_jj')
