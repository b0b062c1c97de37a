#!/usr/bin/env python
"""Restructure the reference's default configuration values into
radiobear_b200/data/planet_defaults.json  (per planet: {attribute: value}; plus the token table
{token: [attribute, unit]}).

Input (read-only): /root/reference/radiobear/default_config.json and default_state.json
(read by config.py:35-36, 50-59).  Only configuration DATA (planet radii, GM, token names, default
formalism names ...) is re-keyed; no source code is copied.  Run in the build container:
    python tools/build_planet_defaults.py
"""
import json
import os

REF = os.environ.get('RADIOBEAR_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'radiobear_b200', 'data',
                   'planet_defaults.json')


def main():
    toks = {}
    planets = {}
    for fn in ['default_config.json', 'default_state.json']:
        with open(os.path.join(REF, 'radiobear', fn)) as fp:
            d = json.load(fp)['toks']
        for tok, spec in d.items():
            toks[tok] = [spec['name'], spec['unit']]
            for planet, val in spec['default'].items():
                planets.setdefault(planet, {})[spec['name']] = val
    with open(OUT, 'w') as fp:
        json.dump({'tokens': toks, 'planets': planets}, fp, indent=1, sort_keys=True)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes;', len(toks), 'tokens;', sorted(planets))


if __name__ == '__main__':
    main()
