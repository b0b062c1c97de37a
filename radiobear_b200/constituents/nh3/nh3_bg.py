"""nh3_bg -- NH3 Berge-Gulkis Ben-Reuven line sum (reference nh3/nh3_bg.py:26-74).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('nh3', 'nh3_bg')
