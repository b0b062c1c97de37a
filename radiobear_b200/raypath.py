"""Ray geometry: path length through every layer shell for an impact point b.

API mirror of raypath.compute_ds / Ray (raypath.py:20-36, 108-273).  The geometry itself runs in
the ray_geometry CUDA kernel (csrc/rt_kernels.cu), one thread per ray; `compute_ds_batch` exposes
the batched form the image / profile paths use.
"""
import numpy as np

from . import engine


class Ray:
    """Holds the ray parameters (same fields as the reference)."""

    allowed_parameters = ['ds', 'layer4ds', 'r4ds', 'P4ds', 'doppler', 'tip', 'rotate', 'rNorm']

    def __init__(self):
        for k in self.allowed_parameters:
            setattr(self, k, None)

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if k not in self.allowed_parameters:
                raise ValueError('{} not allowed Ray parameter'.format(k))
            setattr(self, k, v)


def computeAspect(Q, f=1.0):
    """[position angle, planetographic sub-earth latitude] (deg) -> (tip, rotate) in rad (raypath.py:39-44)."""
    tip = -Q[0] * np.pi / 180.0
    rotate = -np.arctan(np.tan(Q[1] * np.pi / 180.0) * (1.0 - f)**2)
    return tip, rotate


def _about_x_then(tip, rotate, v, x_last):
    ct, st, cr, sr = np.cos(tip), np.sin(tip), np.cos(rotate), np.sin(rotate)
    about_z = np.array([[ct, -st, 0.0], [st, ct, 0.0], [0.0, 0.0, 1.0]])
    about_x = np.array([[1.0, 0.0, 0.0], [0.0, cr, -sr], [0.0, sr, cr]])
    v = np.asarray(v, dtype=np.float64)
    return about_x @ (about_z @ v) if x_last else about_z @ (about_x @ v)


def rotate2planet(rotate, tip, b):
    """Observer frame -> planet frame: about z by `tip`, then about x by `rotate` (raypath.py:47-51; the order the ray
    trace kernel applies to the impact vector)."""
    return _about_x_then(tip, rotate, b, True)


def rotate2obs(rotate, tip, b):
    """The two rotations in the other order (raypath.py:54-57); rotate2obs(-rotate, -tip, .) undoes rotate2planet."""
    return _about_x_then(tip, rotate, b, False)


def _geometry_args(atm, orientation, gtype):
    cfg = atm.config
    LP = cfg.LP
    if gtype is None:
        gtype = cfg.gtype
    if orientation is None:
        orientation = cfg.orientation
    args = dict(radius=atm.property[LP['R']], refr_index=atm.property[LP['N']], Req=cfg.Req, Rpol=cfg.Rpol,
                orientation=[float(orientation[0]), float(orientation[1])], gtype=gtype,
                limb=getattr(cfg, 'limb', 'shape'))
    if gtype == 'gravity':
        # what Shape._calcGeoid / _gravity read from the planet (shape.py:141-221)
        args['gravity_model'] = dict(GM_profile=atm.property[LP['GM']], Jn=cfg.Jn, RJ=cfg.RJ, omega_m=cfg.omega_m,
                                     vwlat=cfg.vwlat, vwdat=cfg.vwdat)
    return args


def compute_ds_batch(atm, b, orientation=None, gtype=None):
    """ds[R][L-1] (km), nseg[R] (-1: ray misses the planet), (tip, rotate, rNorm)."""
    return engine.compute_ds(b=np.atleast_2d(np.asarray(b, dtype=np.float64)), **_geometry_args(atm, orientation, gtype))


def compute_ds(atm, b, orientation=None, gtype=None, verbose=False):
    """Path lengths for one impact point -> Ray (ds is None when the ray misses, raypath.py:126-140)."""
    path = Ray()
    ds, nseg, aspect = compute_ds_batch(atm, [b], orientation, gtype)
    n = int(nseg[0])
    if n < 0:
        return path
    C, LP = atm.config.C, atm.config.LP
    layers = list(range(n))
    radius = atm.property[LP['R']]
    # the descriptive fields of the ray (raypath.py:186-187, 224): shell radius at each step and the Doppler factor
    # 1 - (omega r cos(lat) + vw(lat)) sin(lng) / c of the point it starts at
    fields = engine.compute_ray_fields(b=np.atleast_2d(np.asarray(b, dtype=np.float64)),
                                       **_geometry_args(atm, orientation, gtype))[0]
    r4, lat, lng = fields[0, :n], fields[1, :n], fields[2, :n]
    cfg = atm.config
    vw = np.interp(lat, getattr(cfg, 'vwlat', [0.0, 90.0]), getattr(cfg, 'vwdat', [0.0, 0.0])) / 1000.0
    doppler = 1.0 - (getattr(cfg, 'omega_m', 0.0) * r4 * np.cos(np.radians(lat)) + vw) * np.sin(np.radians(lng)) / 3.0E5
    path.update(ds=list(ds[0, :n]), layer4ds=layers, r4ds=list(r4), P4ds=list(atm.gas[C['P']][:n]), doppler=list(doppler),
                tip=float(aspect[0]), rotate=float(aspect[1]), rNorm=float(radius[0]))
    return path
