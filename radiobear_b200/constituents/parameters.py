"""Plugin keyword defaults (reference: constituents/parameters.py:4-8)."""
from argparse import Namespace


def setpar(kwargs):
    par = Namespace(units='dBperkm', path='./', verbose=False)
    for key, val in kwargs.items():
        setattr(par, key, val)
    return par
