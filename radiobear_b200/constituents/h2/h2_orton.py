"""h2_orton -- H2-H2/He/CH4 collision-induced absorption from Orton's tables (reference h2/h2_orton.py:126-222).

Plugin shim with the signature every other plugin has: the reference's own module cannot be selected in
config.par because its `alpha()` rejects the truncate_* keywords Alpha passes (h2_orton.py:120 vs
alpha.py:210-213).  The piece-wise quadratic interpolation of the tables in frequency is done on the host per
frequency vector (radiobear_b200/catalogs.py:orton_table); the temperature branches (T^4 extrapolation below
40 K, cubic spline 40..400 K, scaled h2_jj above) run in the alpha_lines kernel (csrc/alpha_kernels.cu).
`other_dict['h2state']` selects equilibrium ('e') or normal ('n') hydrogen; `h2newset` is not needed (the
prepared table is keyed by the frequency vector).
"""
from radiobear_b200.constituents._plugin import make_alpha

alpha = make_alpha('h2', 'h2_orton')
