"""Planet configuration: defaults + `<Planet>/config.par` + keyword overrides.

Reads the same `token value [unit]` files as the reference (config.py:66-139): scalar tokens with an
optional unit, list tokens, and dict tokens such as `alpha nh3:nh3_hs_sjs h2s:h2s_ddb` or the column
maps `constituents Z T P H2 ...` (position -> index).  Defaults come from
data/planet_defaults.json (tools/build_planet_defaults.py).
"""
import copy
import json
import os

from . import utils

_DEFAULTS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'planet_defaults.json')
_SPECIAL = {'true': True, 'false': False, 'none': None, 'null': None}
_defaults = None


def _load_defaults():
    global _defaults
    if _defaults is None:
        with open(_DEFAULTS_PATH) as fp:
            _defaults = json.load(fp)
    return _defaults


def set_single_val(val, unit=None):
    """String -> bool/None/list/float/int where it parses (config.py:11-32); floats get unit-converted."""
    if isinstance(val, str):
        low = val.lower()
        if low in _SPECIAL:
            val = _SPECIAL[low]
        elif ',' in val:
            val = val.split(',')
        else:
            kind = float if ('.' in val or 'e' in val or 'E' in val) else int
            try:
                val = kind(val)
            except ValueError:
                pass
    if isinstance(val, float):
        val = utils.convert_unit(val, unit)
    return val


class planetConfig:
    """Configuration namespace: one attribute per token (see `toks`)."""

    def __init__(self, planet, configFile=None, path=None):
        planet = planet.capitalize()
        self.planet = planet
        self.filename = configFile
        self.path = planet if path is None else path
        d = _load_defaults()
        self.toks = {tok: spec[0] for tok, spec in d['tokens'].items()}
        units = {spec[0]: spec[1] for spec in d['tokens'].values()}
        pl = d['planets'].get(planet, d['planets']['X'])
        for name, val in pl.items():
            val = copy.deepcopy(val)
            if isinstance(val, (str, float)):
                val = set_single_val(val, units.get(name))
            setattr(self, name, val)
        self.vwlat, self.vwdat = [0.0, 90.0], [0.0, 0.0]
        self.setConfig(configFile)

    def setConfig(self, configFile):
        if configFile is None:
            return 0
        try:
            fp = open(configFile, 'r')
        except IOError:
            print(configFile, ' not found.  Using defaults.')
            return 0
        with fp:
            for line in fp:
                if line[0] in utils.commentChars or len(line) < 4:
                    continue
                line = line.split('#', 1)[0]
                data = line.split()
                if not data:
                    continue
                tok = data.pop(0).lower()
                if tok not in self.toks:
                    print('token {} not found'.format(tok))
                    continue
                name = self.toks[tok]
                pre = getattr(self, name)
                if isinstance(pre, dict):
                    val = pre
                    for i, v in enumerate(data):
                        if ':' in v:
                            key, form = v.split(':')
                            val[key.lower().strip()] = set_single_val(form.lower().strip())
                        else:
                            val[v.strip()] = i
                elif isinstance(pre, list):
                    if len(data) == 1 and ',' in data[0]:
                        data = data[0].split(',')
                    val = [set_single_val(x) for x in data]
                else:
                    if not data:
                        print("{} didn't have an associated argument in config file.".format(tok))
                        continue
                    val = set_single_val(data[0], data[1] if len(data) == 2 else 'none')
                setattr(self, name, val)
        for cmap in (self.C, self.Cl):
            if 'DZ' not in cmap:
                cmap['DZ'] = len(cmap)
        try:
            with open(self.zonal, 'r') as zp:
                rows = [ln.split() for ln in zp if ln.split()]
            self.vwlat = [float(r[0]) for r in rows]
            self.vwdat = [float(r[1]) for r in rows]
        except (IOError, OSError, TypeError):
            self.vwlat, self.vwdat = [0.0, 90.0], [0.0, 0.0]
        return 1

    def update_config(self, key=None, value=None, **kwargs):
        """update_config(dict) | update_config([keys], [values]) | update_config(key, value) | kwargs."""
        pairs = []
        if isinstance(key, dict):
            pairs += list(key.items())
        elif isinstance(key, list):
            if len(key) != len(value):
                print("key/value pairs not matched.")
            else:
                pairs += list(zip(key, value))
        elif key is not None:
            pairs.append((key, value))
        pairs += list(kwargs.items())
        for k, v in pairs:
            setattr(self, self.toks.get(k, k), set_single_val(v))

    def show(self, print_it=True):
        s = 'Run parameters:\n'
        for key in sorted(self.toks):
            s += '\t{:20s}:  {}\n'.format(key, str(getattr(self, self.toks[key], None)))
        if print_it:
            print(s)
        return s

    def __str__(self):
        return self.show(print_it=False)
