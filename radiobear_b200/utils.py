"""Units, constants and small helpers shared by the host-side API mirror.

Mirrors the parts of the reference's utils.py the hot path touches: unit conversion of
frequency requests (utils.py:54-65), T_cmb (utils.py:77), alpha / layer units (utils.py:71-72),
b_type (utils.py:143-153), isanynum (utils.py:31-39).
"""
import math
import numbers

import numpy as np

T_cmb = 2.725
alphaUnit = 'invcm'
atmLayerUnit = 'km'
commentChars = ('!', '#', '$', '%', '&', '*')

# value of one unit expressed in the SI base unit of its dimension
Units = {'Hz': 1.0, 'kHz': 1.0e3, 'MHz': 1.0e6, 'GHz': 1.0e9,
         'm': 1.0, 'km': 1.0e3, 'cm': 1.0e-2, 'AU': 149597870691.0,
         'bars': 1.0, 'atm': 1.01325,
         'sec': 1.0, 'min': 60.0, 'hr': 3600.0, 'day': 86400.0, 'year': 31536000.0,
         'mpersec2': 1.0, 'cmpersec2': 0.01}
# processing unit -> the units that convert into it
processingUnits = {'GHz': ('GHz', 'Hz', 'kHz', 'MHz'), 'km': ('m', 'cm', 'AU', 'km'), 'bars': ('bars', 'atm'),
                   'sec': ('sec', 'min', 'hr', 'day', 'year'), 'mpersec2': ('mpersec2', 'cmpersec2')}


def proc_unit(supplied_unit):
    """Processing unit a supplied unit is converted to (None when unknown)."""
    for proc, members in processingUnits.items():
        if supplied_unit in members:
            return proc
    return None


def convert_unit(value, supplied_unit):
    """Convert `value` from supplied_unit to its processing unit; unknown units pass through."""
    proc = proc_unit(supplied_unit)
    if supplied_unit is None or proc is None:
        return value
    return value * Units[supplied_unit] / Units[proc]


processingFreqUnit = proc_unit('Hz')


def isanynum(x):
    """True for anything float() accepts that is not a bool / dict / list."""
    if isinstance(x, (bool, dict, list)):
        return False
    try:
        float(x)
    except (TypeError, ValueError):
        return False
    return True


def b_type(b):
    """'disc...' strings pass through lower-cased; otherwise classify by length."""
    if isinstance(b, str):
        return b.lower()
    if isinstance(b[0], str):
        return b[0].lower()
    n = len(b)
    if n > 20:
        return 'image'
    return 'profile' if n > 9 else 'spectrum'


def r2d(a):
    return a * 180.0 / math.pi


def d2r(a):
    return a * math.pi / 180.0


def r2asec(a):
    return 3600.0 * r2d(a)


def timer(dt):
    return dt.seconds + dt.microseconds / 1e6


def data_line(line):
    """Floats of a whitespace-separated data line; None for comments / short / non-numeric lines."""
    if not line or line[0] in commentChars or len(line) < 4:
        return None
    try:
        return [float(x) for x in line.split()]
    except ValueError:
        return None


def is_number(x):
    return isinstance(x, numbers.Number) and not isinstance(x, bool)


def as_float_list(x):
    return [float(v) for v in np.atleast_1d(x)]
