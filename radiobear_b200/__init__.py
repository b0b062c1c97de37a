"""radiobear_b200 -- B200 (sm_100a) implementation of RadioBEAR's two hot paths.

Host code keeps the reference's Python API (Planet.run / Alpha.get_layers / Brightness.single /
raypath.compute_ds / constituents.<gas>.<formalism>.alpha); the arithmetic runs in hand-written CUDA
kernels behind the C ABI of include/radiobear_b200.h.  There is no CPU fallback.
"""
__version__ = '0.1.0'

_SUBMODULES = ('planet', 'alpha', 'brightness', 'raypath', 'atmosphere', 'config', 'data_handling', 'fileIO', 'set_utils',
               'utils', 'logging', 'parallel', 'engine', 'constituents')


def __getattr__(name):
    """`import radiobear_b200 as rb; rb.planet.Planet('jupiter')` like the reference's package (radiobear/__init__.py:6),
    without loading anything before it is asked for."""
    if name in _SUBMODULES:
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError('module {!r} has no attribute {!r}'.format(__name__, name))
