"""radiobear_b200 -- B200 (sm_100a) implementation of RadioBEAR's two hot paths.

Host code keeps the reference's Python API (Planet.run / Alpha.get_layers / Brightness.single /
raypath.compute_ds / constituents.<gas>.<formalism>.alpha); the arithmetic runs in hand-written CUDA
kernels behind the C ABI of include/radiobear_b200.h.  There is no CPU fallback.
"""
__version__ = '0.1.0'
