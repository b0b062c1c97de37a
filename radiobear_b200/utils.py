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


# radio band names by lower / upper edge in GHz (utils.py:79-82)
rfBands = {'HF': (0.003, 0.03), 'VHF': (0.03, 0.3), 'UHF': (0.3, 1.0), 'L': (1.0, 2.0), 'S': (2.0, 4.0), 'C': (4.0, 8.0),
           'X': (8.0, 12.0), 'Ku': (12.0, 18.0), 'K': (18.0, 26.5), 'Ka': (26.5, 40.0), 'Q': (40.0, 50.0), 'V': (50.0, 75.0),
           'W': (75.0, 110.0)}
get_data_from = data_line                                       # the reference's name (utils.py:152)


def getRFband(freq, unit='GHz'):
    """Band letter of a frequency, None outside 3 MHz - 110 GHz (utils.py:105-111)."""
    ghz = freq * Units[unit] / Units['GHz']
    return next((name for name, (lo, hi) in rfBands.items() if lo <= ghz < hi), None)


def invertDictionary(dic, reverse=False):
    """value -> key map and its sorted keys (utils.py:114-120)."""
    inverse = {v: k for k, v in dic.items()}
    return inverse, sorted(inverse, reverse=reverse)


def rb_path(add_path=None):
    """Directory of the package, optionally joined with a sub-path (utils.py:25-30)."""
    import os
    here = os.path.dirname(__file__)
    return here if add_path is None else os.path.join(here, add_path)


def ls(directory='Output', tag='dat', show=True, returnList=False):
    """Numbered list of the non-hidden files of a directory whose name contains `tag` (utils.py:123-136)."""
    import os
    files = [os.path.join(directory, name) for name in os.listdir(directory)
             if name[0] != '.' and (tag is None or (isinstance(tag, str) and tag in name))]
    if show:
        for i, name in enumerate(files):
            print('{}:  {}'.format(i, name))
    return files if returnList else None


def get_expected_number_of_entries(fp):
    """Most frequent number of floats per data line of an open table; ValueError when the other line lengths together
    occur more than half as often (utils.py:163-182).  Rewinds the file."""
    counts = {}
    for line in fp:
        vals = data_line(line)
        if vals is not None:
            counts[len(vals)] = counts.get(len(vals), 0) + 1
    fp.seek(0)
    by_count, ordered = invertDictionary(counts, reverse=True)
    if len(ordered) > 1 and 2 * sum(ordered[1:]) > ordered[0]:
        raise ValueError("Not enough data lines in file: {}".format(ordered))
    return by_count[ordered[0]]
