"""ph3_jh -- PH3 weighted Van Vleck-Weisskopf-type line sum (reference ph3/ph3_jh.py:64-108).

Plugin shim: same signature as the reference module; the work is one launch of the
alpha_lines kernel (csrc/alpha_kernels.cu) through rb_alpha_layers.
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('ph3', 'ph3_jh')
