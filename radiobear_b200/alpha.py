"""Alpha: layer absorption for an atmosphere, computed by the sm_100a alpha_lines kernel.

API mirror of the reference's Alpha (alpha.py:13-305).  Where the reference loops over layers and
calls each constituent plugin per layer (alpha.py:298-300, 202-213), `get_layers` hands the whole
atmosphere to one kernel launch (engine.alpha_layers).  The formalism names of `config.par`
(`alpha nh3:nh3_hs_sjs ...`) keep their meaning; an unknown formalism is reported and dropped just
like a failed plugin import (alpha.py:69-72).
"""
import os
from argparse import Namespace

import numpy as np

from . import engine
from . import logging as rblog
from . import utils
from ._lib import FORMALISM_IDS


class Alpha:
    saved_fields = ['ordered_constituents', 'alpha_data', 'freqs', 'P']

    def __init__(self, idnum=0, config=None, log=None, load_formal=True, verbose=True, **kwargs):
        self.verbose = verbose
        self.log = rblog.setup(log)
        if config is None or isinstance(config, str):
            from . import config as pcfg
            config = pcfg.planetConfig('x', configFile=config)
            config.update_config(**kwargs)
        self.config = config
        self.constituentsAreAt = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'constituents')
        self.idnum = idnum
        self.reset_layers()
        self.alphafile = os.path.join(getattr(config, 'scratch_directory', 'Scratch'),
                                      'alpha{:04d}.npz'.format(idnum))
        self.memory = Namespace()
        self._slab = None           # [L][F] host copy of the slab (what .layers is the transpose view of)
        self._layers_view = None
        self._res = None            # engine.ResidentSlab: the slab (and, after save_alpha='memory', the cube) on the device
        self._dev_cube = None       # (ResidentSlab holding the cube, the host array it mirrors)
        # sharding of get_layers over the ranks of torch.distributed: 'auto' (large requests), True, False
        self.shard = kwargs.get('shard', kwargs.get('shard_freqs', 'auto'))
        self.shard_axis = kwargs.get('shard_axis', 'layers')    # 'layers' (default) or 'freqs'  (parallel.alpha_layers_sharded)
        if load_formal:
            self.setup_formalisms()

    # ------------------------------------------------------------------ formalisms
    def setup_formalisms(self):
        """Resolve config.constituent_alpha into kernel formalisms (alpha.py:49-102)."""
        self.constituent = {}
        self.truncate_strength = {}
        self.truncate_freq = {}
        cfg = self.config
        for c, absorber in cfg.constituent_alpha.items():
            if absorber is None:
                continue
            if absorber in FORMALISM_IDS:
                self.constituent[c] = absorber
            else:
                msg = "WARNING:  CAN'T LOAD " + str(absorber) + '.  '
                print(msg * 3)
                self.log.add("Can't load " + str(absorber), True)
            self.truncate_freq[c] = None
            self.truncate_strength[c] = None
            method = getattr(cfg, 'truncate_method', {}).get(c)
            if method is not None:
                for tm in method.split(','):
                    getattr(self, tm)[c] = getattr(cfg, tm)[c]
        self.ordered_constituents = sorted(self.constituent.keys())
        self.log.add('Using modules:', self.verbose)
        for c in self.constituent:
            self.log.add("\t{}: {} \t".format(c, self.constituent[c]), self.verbose)
        copy_back = {'h2': ['h2state', 'h2newset'],
                     'clouds': ['water_p', 'ice_p', 'nh4sh_p', 'nh3ice_p', 'h2sice_p', 'ch4_p'],
                     'co': ['coshape']}
        self.other_dict = {}
        for c in self.ordered_constituents:
            self.other_dict[c] = {k: getattr(cfg, k) for k in copy_back.get(c, []) if hasattr(cfg, k)}

    def formalisms(self):
        return [(c, self.constituent[c]) for c in self.ordered_constituents]

    def reset_layers(self):
        self.P = None
        self.freqs = None
        self._slab = None
        self._res = None

    # ------------------------------------------------------------------ the slab: on the device until somebody reads it
    # Planet.run -> Alpha.get_layers -> Brightness.batch never needs the absorption on the host: the kernel leaves it in
    # the context's resident buffer and the integration reads it there.  `.layers` / `.slab` copy it back on first use.
    def has_layers(self):
        return self._slab is not None or (self._res is not None and self._res.valid())

    def materialize(self):
        """Bring the device-resident slab to the host (called before another computation overwrites the buffer)."""
        if self._slab is None and self._res is not None and self._res.valid():
            self._slab = self._res.fetch()

    @property
    def slab(self):
        self.materialize()
        return self._slab

    @slab.setter
    def slab(self, value):
        self._slab, self._res = value, None

    @property
    def layers(self):
        """[F][L] view, indexable as layers[j][layer] like the reference (alpha.py:301)."""
        s = self.slab
        if s is None:
            return None
        if self._layers_view is None or self._layers_view.base is not s:
            self._layers_view = s.T
        return self._layers_view

    @layers.setter
    def layers(self, value):
        self.slab = None if value is None else np.ascontiguousarray(np.asarray(value, dtype=np.float64).T)

    def rt_slab(self):
        """What the integration should read: the resident handle while it is valid, else the host slab."""
        if self._res is not None and self._res.valid():
            return self._res
        return self.slab

    @property
    def n_freqs(self):
        return (self._slab if self._slab is not None else self._res).shape[1]

    # ------------------------------------------------------------------ cache (alpha.py:110-149)
    def save_alpha_data(self, save_type):
        if save_type == 'file':
            np.savez(self.alphafile, ordered_constituents=self.ordered_constituents, alpha_data=self.tosave,
                     freqs=self.freqs, P=self.P)
        elif save_type == 'memory':
            self.memory.ordered_constituents = self.ordered_constituents
            self.memory.alpha_data = self.tosave
            self.memory.freqs = self.freqs
            self.memory.P = self.P

    def read_alpha_data(self, save_type):
        if save_type == 'file':
            src = np.load(self.alphafile)
            for sf in self.saved_fields:
                setattr(self, sf, src[sf])
            # the npz holds a numpy string array; the reference looks constituents up by key (alpha.py:151-192), a
            # dict `scale` therefore works on a file cache too
            self.ordered_constituents = [str(c) for c in src['ordered_constituents']]
        elif save_type == 'memory':
            for sf in self.saved_fields:
                setattr(self, sf, getattr(self.memory, sf))

    def get_layer_scale(self, scale, N):
        """Validate a scale request; returns the [C][N] matrix the kernel applies (or None)."""
        return engine.scale_matrix(scale, self.ordered_constituents, N)

    # ------------------------------------------------------------------ the hot call
    def get_layers(self, freqs, atm, scale=False, get_alpha='calc', save_alpha='none'):
        """Compute (or re-scale cached) absorption for all layers: sets .layers[F, L], .P, .freqs."""
        self.reset_layers()
        self.freqs = freqs
        self._scale = scale
        C = atm.config.C
        self.P = atm.gas[C['P']]
        L = atm.gas.shape[1]
        from_cache = get_alpha in ('memory', 'file')
        to_cache = save_alpha in ('memory', 'file')
        self.log.add('{} layers'.format(L), self.verbose)
        slab = cube = res = None
        if from_cache:
            # cached per-constituent cube [L][F][C]: only the scale-sum is redone (alpha.py:224-225)
            self.read_alpha_data(get_alpha)
            dev = self._dev_cube
            if (get_alpha == 'memory' and not to_cache and dev is not None and dev[1] is self.alpha_data
                    and dev[0].cube_valid() and dev[0].cube_shape[0] == L):
                # the cube this cache holds is still on the device (left there by save_alpha='memory'): the retrieval
                # loop of scripts/demo_batch.py re-runs only the scale-sum and the radiative transfer, with no copy
                res = engine.alpha_rescale_resident(dev[0], self.get_layer_scale(scale, L), owner=self)
            else:
                out = engine.alpha_scale_sum(np.asarray(self.alpha_data, dtype=np.float64), self.get_layer_scale(scale, L),
                                             want_cube=to_cache)
                slab, cube = out if to_cache else (out, None)
        else:
            fr = np.asarray(freqs, dtype=np.float64)
            common = dict(cloud=atm.cloud if np.size(atm.cloud) else None, cloud_dict=atm.config.Cl,
                          formalisms=self.formalisms(), other_dicts=self.other_dict, units=utils.alphaUnit, scale=scale,
                          truncate_strength=self.truncate_strength, truncate_freq=self.truncate_freq)
            from . import parallel
            world, _ = parallel.world_rank()
            if not to_cache and parallel.shard_alpha(L, len(fr), world, self.shard):
                # one process per GPU: this rank's block of layers (or of frequencies), then one all_gather (every
                # rank traces rays with the full slab afterwards)
                cl = common['cloud']
                slab = None
                if self.shard_axis == 'layers':
                    if isinstance(scale, dict):                    # validated against all layers before it is cut
                        self.get_layer_scale(scale, L)
                    # NCCL ranks on one NVSwitch domain: the all_gather happens inside the absorption kernel
                    slab = parallel.alpha_layers_scatter(fr, atm.gas[C['T']], atm.gas[C['P']], atm.gas, C, cl,
                                                         atm.config.Cl, self.formalisms(), self.other_dict, utils.alphaUnit,
                                                         scale, self.truncate_strength, self.truncate_freq)
                if slab is not None:
                    pass
                elif self.shard_axis == 'freqs':
                    def block(lo, hi):
                        return engine.alpha_layers(fr[lo:hi], atm.gas[C['T']], atm.gas[C['P']], atm.gas, C, **common)
                else:
                    def block(lo, hi):
                        kw = dict(common, cloud=None if cl is None else np.ascontiguousarray(cl[:, lo:hi]),
                                  scale=parallel.slice_scale(scale, lo, hi))
                        if isinstance(scale, dict):                # validated against all layers before it is cut
                            self.get_layer_scale(scale, L)
                        return engine.alpha_layers(fr, atm.gas[C['T']][lo:hi], atm.gas[C['P']][lo:hi],
                                                   np.ascontiguousarray(atm.gas[:, lo:hi]), C, **kw)
                if slab is None:
                    slab = parallel.alpha_layers_sharded(block, L, len(fr), axis=self.shard_axis)
            elif os.environ.get('RB_ALPHA_RESIDENT', '1') == '0':     # A/B switch: results through host memory
                out = engine.alpha_layers(fr, atm.gas[C['T']], atm.gas[C['P']], atm.gas, C, want_cube=to_cache, **common)
                slab, cube = out if to_cache else (out, None)
            else:
                res = engine.alpha_layers_resident(fr, atm.gas[C['T']], atm.gas[C['P']], atm.gas, C, keep_cube=to_cache,
                                                   owner=self, **common)
                if to_cache:
                    cube = res.fetch_cube()
        self._slab, self._res = slab, res
        if to_cache:
            self.tosave = cube
            self.save_alpha_data(save_alpha)
            # the device copy of what the memory cache now holds (None when the cube went through the host path)
            self._dev_cube = (res, self.memory.alpha_data) if (save_alpha == 'memory' and res is not None) else None
            del self.tosave

    def layers_at(self, freq_matrix, atm, scale=None):
        """Total absorption slab[L][F] with every layer evaluated at its own frequencies freq_matrix[L][F] -- what the
        Doppler branch of Brightness.single asks for step by step (brightness.py:83-92: one plugin sweep per step and
        frequency at f / doppler); here one kernel launch (rb_alpha_desc::freqs_per_layer).  `scale` as in get_layers
        (default: the scale of the last get_layers call)."""
        C = atm.config.C
        if scale is None:
            scale = getattr(self, '_scale', False)
        return engine.alpha_layers(np.asarray(freq_matrix, dtype=np.float64), atm.gas[C['T']], atm.gas[C['P']], atm.gas, C,
                                   cloud=atm.cloud if np.size(atm.cloud) else None, cloud_dict=atm.config.Cl,
                                   formalisms=self.formalisms(), other_dicts=self.other_dict, units=utils.alphaUnit,
                                   scale=scale, truncate_strength=self.truncate_strength, truncate_freq=self.truncate_freq)

    def get_single_layer(self, freqs, layer, atm, lscale=1.0, units='invcm'):
        """Total absorption of one layer (alpha.py:218-233)."""
        C = atm.config.C
        sl = slice(layer, layer + 1)
        slab = engine.alpha_layers(np.asarray(freqs, dtype=np.float64), atm.gas[C['T']][sl], atm.gas[C['P']][sl],
                                   np.ascontiguousarray(atm.gas[:, sl]), C,
                                   cloud=np.ascontiguousarray(atm.cloud[:, sl]) if np.size(atm.cloud) else None,
                                   cloud_dict=atm.config.Cl, formalisms=self.formalisms(), other_dicts=self.other_dict,
                                   units=units, scale=self._one_layer_scale(lscale),
                                   truncate_strength=self.truncate_strength, truncate_freq=self.truncate_freq)
        return slab[0]

    def _one_layer_scale(self, lscale):
        """The scale of one layer as the reference hands it around (a number, or {constituent: number}, the value
        get_layer_scale builds per layer, alpha.py:235-259) -> the request form of a one-layer absorption call.  Names
        that are not constituents are ignored like in total_layer_alpha (alpha.py:182-184)."""
        if isinstance(lscale, dict):
            return {k: [float(v)] for k, v in lscale.items() if k in self.ordered_constituents}
        return lscale

    def total_layer_alpha(self, absorb, lscale):
        """Scale-sum of one layer's per-constituent absorption absorb[F][C] (the array get_alpha_from_calc returns,
        columns in ordered_constituents order) -> total[F] (alpha.py:151-192).  `lscale`: a number for all constituents
        or {constituent: number}; constituents that are not named keep their value, names that are not constituents are
        ignored like in the reference.  Device scale-sum (rb_alpha_scale_sum, L = 1)."""
        absorb = np.ascontiguousarray(absorb, dtype=np.float64)
        if absorb.ndim != 2:
            raise ValueError('absorb must be [F][C], got shape {}'.format(absorb.shape))
        ncon = absorb.shape[1]
        if utils.isanynum(lscale):
            m = np.full((ncon, 1), float(lscale))
        else:
            m = np.ones((ncon, 1))
            for j, name in enumerate(list(self.ordered_constituents)[:ncon]):
                if name in lscale:
                    m[j, 0] = float(lscale[name])
        total, scaled = engine.alpha_scale_sum(absorb[None], m, want_cube=True)
        if getattr(self, '_save_alpha_memfil', False):             # the reference's per-layer cache list
            self.tosave.append(scaled[0])
        return total[0]

    def get_alpha_from_calc(self, freqs, T, P, gas, gas_dict, cloud, cloud_dict, units='invcm'):
        """Per-constituent absorption of one (T, P, X) point -> [F][C] (alpha.py:194-216)."""
        _, cube = engine.alpha_layers(np.asarray(freqs, dtype=np.float64), [T], [P], np.asarray(gas, dtype=np.float64),
                                      gas_dict, cloud=None if cloud is None else np.asarray(cloud, dtype=np.float64),
                                      cloud_dict=cloud_dict, formalisms=self.formalisms(), other_dicts=self.other_dict,
                                      units=units, want_cube=True, truncate_strength=self.truncate_strength,
                                      truncate_freq=self.truncate_freq)
        return cube[0]
