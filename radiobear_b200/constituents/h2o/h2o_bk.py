"""h2o_bk -- H2O Karpowicz/Steffes: 15 VVW lines + foreign and self continuum (reference h2o/h2o_bk.py:65-187).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('h2o', 'h2o_bk')
