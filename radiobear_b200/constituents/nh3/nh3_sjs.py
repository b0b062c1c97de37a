"""nh3_sjs -- NH3, Spilker (f<=26 GHz) / Joiner-Steffes (f>=34 GHz) Ben-Reuven model (reference nh3/nh3_sjs.py:26-128).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('nh3', 'nh3_sjs')
