"""nh3_sjsd -- NH3: nh3_sjs outside 10..100 bar, triangular blend with nh3_kd inside (reference nh3/nh3_sjsd.py:6-24).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('nh3', 'nh3_sjsd')
