"""Planet: the executive API -- `Planet(name).run(freqs, b)` -> DataReturn.

API mirror of planet.Planet / planet_base.PlanetBase (planet.py:31-160, planet_base.py:28-385).
The reference loops over b-points calling Brightness.single once per ray (planet.py:126-133); here
all rays of a request go through one geometry launch and one integration launch.  Requests the
reference itself cannot run are accepted: float `b` (full image, any number of frequencies ->
Tb[rows][cols][F]) and log-sweep frequency strings.
"""
import datetime
import os
import sys

import numpy as np

from . import alpha as rbalpha
from . import atmosphere as rbatm
from . import brightness as rbbright
from . import config as pcfg
from . import data_handling
from . import logging as rblog
from . import parallel
from . import set_utils
from . import utils

__all__ = ['Planet']
VERSION = '0.1.0 (radiobear_b200; API of radiobear 2.0.1)'


class Planet:
    planet_list = ['Jupiter', 'Saturn', 'Neptune', 'Uranus']

    def __init__(self, name, config_file='config.par', run_atm=True, load_formal=True, verbose=True,
                 atmosphere=None, **kwargs):
        """atmosphere: optional ready-made Atmosphere (or list) -- skips the file pipeline."""
        self.planet = name.capitalize()
        self.verbose = verbose
        self.load_formal = load_formal
        self.header = {}
        self.freqs = []
        self.freqUnit = None
        self.b = None
        self.data_type = None
        self.imSize = None
        self.bmap_loaded = False
        self.version = VERSION
        self.scale = None
        self.get_alpha = None
        self.save_alpha = None
        self.alpha_options = {'f': 'file', 'm': 'memory', 'n': 'none', 'c': 'none'}
        if atmosphere is not None:
            atmos = atmosphere if isinstance(atmosphere, list) else [atmosphere]
            self.config = atmos[0].config
            self.config.update_config(**kwargs)
            self.config_file = getattr(self.config, 'filename', None)
        else:
            self.config_file = os.path.join(self.planet, config_file)
            if verbose:
                print('Reading config file:  ', self.config_file)
            self.config = pcfg.planetConfig(self.planet, configFile=self.config_file)
            self.config.update_config(**kwargs)
            if self.config.path not in sys.path:
                sys.path.insert(0, self.config.path)
            atmos = None
        # log / data_return / atm / alpha / bright  (planet_base.py:61-113)
        if getattr(self.config, 'write_log_file', False):
            start = datetime.datetime.now()
            self.log = rblog.LogIt('{}/{}_{}.log'.format(self.config.log_directory, self.planet,
                                                         start.strftime("%Y%m%d_%H%M%S")))
            self.log.add(self.planet + ' start ' + str(start), self.verbose)
            for entry in (self.planet, config_file, self.config.show(print_it=False)):     # planet.py:47-52
                self.log.add(entry, False)
        else:
            self.log = None
        self.data_return = data_handling.Data()
        self.data_return.set('log', self.log)
        if atmos is None:
            for attr in ('gasFile', 'cloudFile'):
                if not isinstance(getattr(self.config, attr), list):
                    setattr(self.config, attr, [getattr(self.config, attr)])
            atmos = [rbatm.Atmosphere(self.planet, idnum=i, config=self.config, log=self.log, verbose=verbose)
                     for i in range(len(self.config.gasFile))]
            if run_atm:
                for atm in atmos:
                    getattr(atm, getattr(self.config, 'atm_run_type', 'std'))()
        self.atmos = atmos
        self.alpha = [rbalpha.Alpha(idnum=i, config=self.config, log=self.log, load_formal=load_formal,
                                    verbose=verbose) for i in range(len(self.atmos))]
        self.bright = rbbright.Brightness(config=self.config, log=self.log, verbose=verbose)
        from . import fileIO
        self.fIO = fileIO.FileIO(directory=getattr(self.config, 'output_directory', 'Output'))   # planet_base.py:115-118
        self.Tb = []
        self.rNorm = self.tip = self.rotate = None

    # ------------------------------------------------------------------ requests
    def set_freqs(self, freqs, freqUnit='GHz'):
        self.header['freqs'] = '# freqs request: {} {}'.format(str(freqs), freqUnit)
        freqs, freqUnit = set_utils.set_freq(freqs, freqUnit)
        self.data_return.set('f', freqs)
        self.data_return.set('freqUnit', freqUnit)
        return freqs, freqUnit

    def set_b(self, b=(0.0, 0.0), block=(1, 1)):
        self.header['b'] = '# b request:  {}  {}'.format(str(b) if not isinstance(b, np.ndarray) else 'array',
                                                         str(list(block)))
        rv = set_utils.set_b(b, block, Rpol=self.config.Rpol, Req=self.config.Req)
        self.b, self.block, self.data_type, self.imSize = rv.b, rv.block, rv.data_type, rv.imSize
        self.data_return.set('b', self.b)

    def map_b_to_atm(self, b):
        """Atmosphere index for an impact point (planet_base.py:183-193)."""
        if isinstance(b, str):
            return 0
        mod = getattr(self.config, 'bmapmodule', None)
        if mod is None or mod == 'nobmap':
            return 0
        if not self.bmap_loaded:
            __import__(mod)
            self.bmapModule = sys.modules[mod]
            self.bmap_loaded = True
        return self.bmapModule.bmap(b=b)

    def alpha_layers(self, freqs, atmos, scale=False, get_alpha='calc', save_alpha='none'):
        for i, atm in enumerate(atmos):
            self.alpha[i].reset_layers()
            self.alpha[i].get_layers(freqs=freqs, atm=atm, scale=scale, get_alpha=get_alpha, save_alpha=save_alpha)

    def check_reuse(self, freqs, scale, get_alpha, save_alpha, reuse_override='check'):
        """Skip the absorption step when nothing it depends on changed (planet_base.py:303-346)."""
        if reuse_override == 'true':
            return True
        if reuse_override == 'false':
            return False
        if get_alpha != self.get_alpha or save_alpha != self.save_alpha:
            return False
        if self.freqs is None or len(freqs) != len(self.freqs):
            return False
        if any((a - b) / a > 0.0001 for a, b in zip(sorted(self.freqs), sorted(freqs))):
            return False
        if utils.isanynum(scale) and utils.isanynum(self.scale) and (scale - self.scale) / scale > 0.0001:
            return False
        if not isinstance(scale, type(self.scale)):
            return False
        if isinstance(scale, (list, np.ndarray)):
            if len(scale) != len(self.scale) or any((a - b) / a > 0.0001 for a, b in zip(scale, self.scale)):
                return False
        if isinstance(scale, dict):
            if sorted(scale) != sorted(self.scale):
                return False
            for k in scale:
                if len(scale[k]) != len(self.scale[k]) or \
                        any((a - b) / a > 0.001 for a, b in zip(scale[k], self.scale[k])):
                    return False
        return True

    # ------------------------------------------------------------------ run
    def run(self, freqs, b='disc', scale=False, get_alpha='calc', save_alpha='none', freqUnit='GHz', block=(1, 1),
            reuse_override='check'):
        """Brightness temperature for the frequency and impact-parameter requests -> DataReturn."""
        get_alpha = self.alpha_options[get_alpha[0].lower()]
        save_alpha = self.alpha_options[save_alpha[0].lower()]
        freqs, freqUnit = self.set_freqs(freqs=freqs, freqUnit=freqUnit)
        reuse = self.check_reuse(freqs, scale, get_alpha, save_alpha, reuse_override=str(reuse_override).lower())
        t0 = datetime.datetime.now()
        self.set_b(b=b, block=block)
        self._pts = None
        local_pts = self._local_points()
        doppler = bool(getattr(self.config, 'Doppler', False))   # per-ray absorption: Brightness._doppler_ray, ray by ray
        if not reuse and local_pts is not None and not doppler:
            # the ray geometry does not depend on the absorption: start it first, on its own stream
            self.bright.prefetch(local_pts, self.atmos[0], self.config.orientation)
        if not reuse:
            self.freqs = freqs
            self.freqUnit = utils.proc_unit(freqUnit)
            self.scale, self.get_alpha, self.save_alpha = scale, get_alpha, save_alpha
            if self.log is not None:                                                      # planet.py:103-110
                unit = utils.proc_unit(freqUnit)
                self.log.add('{} at {} frequencies ({} - {} {})'.format(self.planet, len(freqs), freqs[0], freqs[-1], unit)
                             if self.verbose else '{} at {} {}'.format(self.planet, freqs[0], unit), self.verbose)
            self.alpha_layers(freqs=self.freqs, atmos=self.atmos, scale=scale, get_alpha=get_alpha,
                              save_alpha=save_alpha)
            if self.verbose:
                print("Absoprtion calc took {:.3f} s".format(utils.timer(datetime.datetime.now() - t0)))
        runStart = datetime.datetime.now()
        F = len(self.freqs)
        disc = isinstance(self.b[0], str)
        if disc:
            res = self.bright.batch([[0.0, 0.0]], self.freqs, self.atmos[0], self.alpha[0], self.config.orientation,
                                    disc_average=True)
            Tb = res['Tb']
        else:
            pts = self._pts if self._pts is not None else np.asarray(self.b, dtype=np.float64)
            # per-point atmosphere index (planet_base.py map_b_to_atm); None = everything uses atmos[0]
            which = np.array([self.map_b_to_atm(list(p)) for p in pts]) if \
                getattr(self.config, 'bmapmodule', 'nobmap') not in (None, 'nobmap') else None
            if which is not None and not which.any():
                which = None
            f32 = self.data_type == 'image'
            world, rank = parallel.world_rank()
            if world > 1 and len(pts) >= 64 * world and which is None and not doppler:
                # one process per GPU: shard rows / points, gather to rank 0 (None on the other ranks)
                rows = (pts[::self.imSize[0], 1], self.imSize[0]) if self.data_type == 'image' else None
                Tb = parallel.run_points_sharded(self, pts, self.atmos[0], self.alpha[0], out_f32=f32, rows=rows)
            elif which is None:
                Tb = self.bright.batch(pts, self.freqs, self.atmos[0], self.alpha[0], self.config.orientation,
                                       out_f32=f32)['Tb']
            else:
                Tb = np.empty((len(pts), F), dtype=np.float32 if f32 else np.float64)
                for j in np.unique(which):
                    sel = np.nonzero(which == j)[0]
                    res = self.bright.batch(pts[sel], self.freqs, self.atmos[j], self.alpha[j],
                                            self.config.orientation, out_f32=f32)
                    Tb[sel] = res['Tb']
        runStop = datetime.datetime.now()
        if self.log is not None:
            self.log.add('Run start ' + str(runStart), False)
            self.log.add('Run stop ' + str(runStop), False)
        if self.data_type == 'image' and Tb is not None:
            ncol, nrow = self.imSize[0], len(self.b) // self.imSize[0]
            Tb = Tb.reshape(nrow, ncol, F)
            if F == 1:
                Tb = Tb[:, :, 0]
        self.Tb = Tb
        radius = self.atmos[0].property[self.config.LP['R']]
        self.rNorm = float(radius[0])
        self.set_header(runStart, runStop)
        self.data_return.set('start', runStart)
        self.data_return.set('stop', runStop)
        self.data_return.set('Tb', None if self.Tb is None else np.asarray(self.Tb))
        self.data_return.set('type', self.data_type)
        self.data_return.set('header', self.header)
        if self.log is not None:
            self.data_return.set('logfile', self.log.logfile)
        if self.verbose:
            print("RT calc took {:.3f} s".format(utils.timer(runStop - runStart)))
        # output files are written by the process that holds the result (rank 0 of a sharded run); the name carries
        # the image block like the reference's (planet.py:142-146, planet_base.py:197-207)
        if getattr(self.config, 'write_output_files', False) and self.Tb is not None:
            block_postfix = '_'
            if self.data_type == 'image' and abs(self.block[1]) > 1:
                block_postfix = '_{:02d}of{:02d}_'.format(self.block[0], abs(self.block[1]))
            fn = '{}_{}{}{}.dat'.format(self.planet, self.data_type, block_postfix, runStart.strftime("%Y%m%d_%H%M%S"))
            self.fIO.write(os.path.join(self.config.output_directory, fn), self.data_return)
        return self.data_return

    def _local_points(self):
        """The impact points this process will integrate with atmos[0] in one batch (its row block when the
        process group shards the image), or None for disc-averaged / b-mapped / tiny requests."""
        if isinstance(self.b[0], str) or getattr(self.config, 'bmapmodule', 'nobmap') not in (None, 'nobmap'):
            return None
        pts = self._pts = np.asarray(self.b, dtype=np.float64)      # run() integrates this very array
        world, rank = parallel.world_rank()
        if world > 1 and len(pts) >= 64 * world:
            rows = (pts[::self.imSize[0], 1], self.imSize[0]) if self.data_type == 'image' else None
            s, e = parallel.local_block(len(pts), self.atmos[0].config, rows, world, rank)
            return pts[s:e] if e > s else None
        return pts

    def set_header(self, run_start, run_stop):
        from . import raypath
        f = 1.0 - self.config.Rpol / self.config.Req
        tip, rotate = raypath.computeAspect(self.config.orientation, f)
        self.tip, self.rotate = tip, rotate
        h = self.header
        h['orientation'] = '# orientation:   {}'.format(repr(self.config.orientation))
        h['aspect'] = '# aspect tip, rotate:  {:.4f}  {:.4f}'.format(utils.r2d(tip), utils.r2d(rotate))
        h['rNorm'] = '# rNorm: {}'.format(self.rNorm)
        if self.data_type == 'image':
            h['imgSize'] = '# imgSize: {}'.format(self.imSize)
            res = utils.r2asec(np.arctan(abs(self.b[1][0] - self.b[0][0]) * self.rNorm / self.config.distance))
            h['res'] = '# res:  {} arcsec'.format(res)
        h['data-type'] = '#* type:  {}'.format(self.data_type)
        h['gtype'] = '# gtype: {}'.format(self.config.gtype)
        h['radii'] = '# radii:  {:.1f}  {:.1f}  km'.format(self.config.Req, self.config.Rpol)
        h['distance'] = '# distance:  {} km'.format(self.config.distance)
        if getattr(self.config, 'write_log_file', False) and self.log is not None:
            h['log-file:'] = '#* logfile: {}'.format(self.log.logfile)
        h['start'] = "#* start: {:%Y-%m-%d %H:%M:%S}".format(run_start)
        h['stop'] = "#* stop: {:%Y-%m-%d %H:%M:%S}".format(run_stop)
