"""co_ddb -- CO Voigt (complex rational) + VVW line sum (reference co/co_ddb.py:22-99).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('co', 'co_ddb')
