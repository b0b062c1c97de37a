"""h2s_ddb -- H2S Ben-Reuven line sum with zeta = gamma (reference h2s/h2s_ddb.py:42-87).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('h2s', 'h2s_ddb')
