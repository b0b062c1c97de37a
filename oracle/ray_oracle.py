"""Oracle (TEST INFRASTRUCTURE, not product code): ray geometry on the CPU.

Scalar numpy restatement of raypath.compute_ds (raypath.py:108-273), findEdge
(raypath.py:60-105), computeAspect / rotate2planet (raypath.py:39-51) and the
ellipse/sphere geoid Shape._calcEllipse (shape.py:223-274) with rotX/Y/Z (shape.py:277-298).
Keeps the reference's trigonometric evaluation order (the CUDA kernel uses an algebraic
form; agreement is a tested property, not an assumption).

Reference behaviours kept on purpose:
* only the first interface refracts (nratio = nr[0]/nr[1], then 1.0: raypath.py:158, 257);
* np.sqrt of a negative number gives NaN, so the 'tangent' branch never fires and limb rays
  carry NaN ds from the tangent depth onward (raypath.py:192-209);
* lat == 0 is replaced by 1e-6 rad in the ellipse (shape.py:231-233);
* geoid.rmag (radius of the shell at the *new* latitude) is used as rNow of the next step
  (raypath.py:181, 235).

Pinned by tests/test_oracle_golden.py against reference-generated ds vectors.
"""
import numpy as np

_XH = np.array([1.0, 0.0, 0.0])
_YH = np.array([0.0, 1.0, 0.0])
_ZH = np.array([0.0, 0.0, 1.0])


def _r2d(a):
    return a * 180.0 / np.pi


def _d2r(a):
    return a * np.pi / 180.0


def rotX(x, V):
    return np.dot(np.array([[1.0, 0.0, 0.0], [0.0, np.cos(x), -np.sin(x)], [0.0, np.sin(x), np.cos(x)]]), V)


def rotY(y, V):
    return np.dot(np.array([[np.cos(y), 0.0, np.sin(y)], [0.0, 1.0, 0.0], [-np.sin(y), 0.0, np.cos(y)]]), V)


def rotZ(z, V):
    return np.dot(np.array([[np.cos(z), -np.sin(z), 0.0], [np.sin(z), np.cos(z), 0.0], [0.0, 0.0, 1.0]]), V)


def compute_aspect(Q, f=1.0):
    """raypath.py:39-44."""
    tip = -Q[0] * np.pi / 180.0
    rotate = -np.arctan(np.tan(Q[1] * np.pi / 180.0) * (1.0 - f)**2)
    return tip, rotate


def rotate2planet(rotate, tip, b):
    """raypath.py:47-51."""
    return rotX(rotate, rotZ(tip, b))


class Ellipse:
    """State of Shape after _calcEllipse (shape.py:223-274): r, n, rmag."""

    def __init__(self, gtype, Req, Rpol):
        self.gtype = gtype
        self.q = Rpol / Req
        self.r = np.zeros(3)
        self.n = np.zeros(3)
        self.rmag = 0.0

    def calc(self, r, pclat, delta_lng):
        a = r
        b = self.q * r if self.gtype == 'ellipse' else r
        lat = _d2r(pclat)
        lng = _d2r(delta_lng)
        if lat == 0.0:
            lat = 1.0E-6
        norm = np.array([0.0, a * np.sin(lat), b * np.cos(lat)])
        norm = norm / np.linalg.norm(norm)
        self.n = rotY(lng, norm)
        r_vec = rotY(lng, np.array([0.0, b * np.sin(lat), a * np.cos(lat)]))
        self.r = r_vec
        self.rmag = np.linalg.norm(r_vec)
        return self.rmag


class Geoid:
    """State of Shape after _calcGeoid (shape.py:141-171) with _gravity (shape.py:173-221): the 'gravity' shape.

    Starts at the equatorial radius r and marches north (or south) in steps of 0.01 deg to the first grid latitude at
    or beyond pclat; every step evaluates the gravity vector (zonal harmonics Jn, rotation with the zonal winds) and
    moves along the local tangent.  r, n and rmag are those of the LAST grid latitude (the march is quantised: the
    latitude the shape is evaluated at is k * latstep, not pclat).  GM is np.interp(r, R profile, GM profile) exactly as
    the reference calls it (the R profile decreases with index; np.interp is used as is).
    """
    latstep0 = 0.01

    def __init__(self, R_profile, GM_profile, Jn, RJ, omega_m, vwlat, vwdat):
        from scipy.special import eval_legendre      # what scipy.special.legendre(i)(x) evaluates
        self._P = eval_legendre
        self.Rp, self.GMp = np.asarray(R_profile, dtype=float), np.asarray(GM_profile, dtype=float)
        self.Jn, self.RJ, self.omega_m = [float(x) for x in Jn], float(RJ), float(omega_m)
        self.vwlat, self.vwdat = np.asarray(vwlat, dtype=float), np.asarray(vwdat, dtype=float)
        self.r = np.zeros(3)
        self.n = np.zeros(3)
        self.t = np.zeros(3)
        self.rmag = 0.0
        self.gamma = 0.0

    def _gravity(self, pclat, delta_lng, r, GM, omega):
        P = self._P
        g_static = GM / r**2
        lat = _d2r(pclat)
        lng = _d2r(delta_lng)
        nsl = 1.0 if lat == 0.0 else np.sign(lat)
        dphi = nsl * 0.00001
        Sr = 0.0
        Sp = 0.0
        sp = np.sin(lat)
        sp1 = np.sin(lat + dphi)
        sp0 = np.sin(lat - dphi)
        for i in range(len(self.Jn)):
            Sr += (i + 1.0) * self.Jn[i] * pow(self.RJ / r, i) * P(i, sp)
            dP = (P(i, sp1) - P(i, sp)) / dphi
            dP += (P(i, sp) - P(i, sp0)) / dphi
            dP *= 0.5
            Sp += self.Jn[i] * pow(self.RJ / r, i) * dP
        gr = (g_static * (1.0 - Sr) - (2.0 / 3.0) * (omega**2.0) * r * (1.0 - P(2, sp)))
        dP = (3.0 * sp * np.sqrt(1.0 - sp**2))
        gp = (1.0 / 3.0) * (omega**2.0) * r * dP + g_static * Sp
        gamma = np.arctan2(gp, gr)
        self.n = rotY(lng, np.array([0.0, np.sin(lat + gamma), np.cos(lat + gamma)]))
        self.t = rotY(lng, np.array([0.0, np.cos(lat + gamma), -np.sin(lat + gamma)]))
        self.r = rotY(lng, np.array([0.0, r * np.sin(lat), r * np.cos(lat)]))
        self.rmag = np.linalg.norm(self.r)
        self.gamma = gamma

    def calc(self, r, pclat, delta_lng):
        nsp = 1.0 if pclat == 0.0 else np.sign(pclat)
        latstep = nsp * self.latstep0
        steps = np.arange(0.0, pclat + latstep, latstep)
        GM = np.interp(r, self.Rp, self.GMp)
        for latv in steps:
            vw = np.interp(latv, self.vwlat, self.vwdat) / 1000.0
            omega = self.omega_m + vw / (r * np.cos(_d2r(latv)))
            self._gravity(latv, delta_lng, r, GM, omega)
            r = np.linalg.norm(self.r + r * _d2r(latstep) * self.t)
        return self.rmag


class GeoidTable(Geoid):
    """The same 'gravity' shape for every layer at once.  A march of Shape._calcGeoid visits the grid latitudes
    k * 0.01 deg one after the other and returns the state of the last one, so every shape raypath.compute_ds asks for is
    an entry (layer, k, hemisphere) of one table; here the march of all layers advances together (numpy over layers,
    the operations of Geoid._gravity in the same order), which makes a gravity ray affordable for the tests.
    `calc(r, ...)` takes r = req[layer] (the only radii the ray loop passes)."""

    def __init__(self, R_profile, GM_profile, Jn, RJ, omega_m, vwlat, vwdat, max_abs_lat=90.0):
        Geoid.__init__(self, R_profile, GM_profile, Jn, RJ, omega_m, vwlat, vwdat)
        self.layer_of = {float(r): i for i, r in enumerate(self.Rp)}
        self.K = int(np.ceil((max_abs_lat + self.latstep0) / self.latstep0))
        self.tab = {}
        for nsp in (1.0, -1.0):
            latstep = nsp * self.latstep0
            r = self.Rp.copy()
            GM = np.array([np.interp(x, self.Rp, self.GMp) for x in self.Rp])
            rm = np.empty((self.K, len(r)))
            gm = np.empty((self.K, len(r)))
            for k in range(self.K):
                latv = 0.0 + k * latstep
                vw = np.interp(latv, self.vwlat, self.vwdat) / 1000.0
                omega = self.omega_m + vw / (r * np.cos(_d2r(latv)))
                rm[k] = r
                gm[k], tvec, rvec = self._gravity_vec(latv, r, GM, omega)
                r = np.sqrt((rvec[0] + r * _d2r(latstep) * tvec[0])**2 + (rvec[1] + r * _d2r(latstep) * tvec[1])**2)
            self.tab[nsp] = (rm, gm)

    def _gravity_vec(self, pclat, r, GM, omega):
        """Geoid._gravity for an array of radii at delta_lng = 0 -> gamma, (t_y, t_z), (r_y, r_z)."""
        P = self._P
        g_static = GM / r**2
        lat = _d2r(pclat)
        nsl = 1.0 if lat == 0.0 else np.sign(lat)
        dphi = nsl * 0.00001
        Sr = 0.0
        Sp = 0.0
        sp, sp1, sp0 = np.sin(lat), np.sin(lat + dphi), np.sin(lat - dphi)
        for i in range(len(self.Jn)):
            Sr = Sr + (i + 1.0) * self.Jn[i] * (self.RJ / r)**i * P(i, sp)
            dP = (P(i, sp1) - P(i, sp)) / dphi
            dP += (P(i, sp) - P(i, sp0)) / dphi
            dP *= 0.5
            Sp = Sp + self.Jn[i] * (self.RJ / r)**i * dP
        gr = (g_static * (1.0 - Sr) - (2.0 / 3.0) * (omega**2.0) * r * (1.0 - P(2, sp)))
        dP = (3.0 * sp * np.sqrt(1.0 - sp**2))
        gp = (1.0 / 3.0) * (omega**2.0) * r * dP + g_static * Sp
        gamma = np.arctan2(gp, gr)
        return gamma, (np.cos(lat + gamma), -np.sin(lat + gamma)), (r * np.sin(lat), r * np.cos(lat))

    def kindex(self, pclat):
        """Index of the last grid latitude of the march to pclat: len(np.arange(0, pclat + latstep, latstep)) - 1."""
        nsp = 1.0 if pclat == 0.0 else float(np.sign(pclat))
        latstep = nsp * self.latstep0
        return nsp, int(np.ceil((pclat + latstep) / latstep)) - 1

    def calc(self, r, pclat, delta_lng):
        if pclat != pclat:
            # a ray below its tangent shell carries NaN positions; the reference's march raises on np.arange(0, nan, nan)
            # here, the ellipse path (and the CUDA kernel for both shapes) carries the NaN on
            self.r = self.n = np.full(3, np.nan)
            self.rmag = np.nan
            return self.rmag
        l = self.layer_of[float(r)]
        nsp, k = self.kindex(pclat)
        rm, gm = self.tab[nsp]
        rk, gamma = rm[k, l], gm[k, l]
        lat = _d2r(0.0 + k * (nsp * self.latstep0))
        lng = _d2r(delta_lng)
        self.n = rotY(lng, np.array([0.0, np.sin(lat + gamma), np.cos(lat + gamma)]))
        self.r = rotY(lng, np.array([0.0, rk * np.sin(lat), rk * np.cos(lat)]))
        self.rmag = np.linalg.norm(self.r)
        self.gamma = gamma
        return self.rmag


def find_edge(b, rNorm, tip, rotate, geoid):
    """raypath.py:60-105."""
    tmp = (b[0]**2 + b[1]**2)
    zQ_Trial = np.arange(np.sqrt(1.0 - tmp) * 1.01, 0.0, -0.005)
    r_zQ, r_pclat = [], []
    hit = False
    for zQ in zQ_Trial:
        b_vec = rotate2planet(rotate, tip, np.array([b[0], b[1], zQ]))
        r1 = np.linalg.norm(b_vec) * rNorm
        r_zQ.append(r1)
        pclat = _r2d(np.arcsin(np.dot(b_vec, _YH) / np.linalg.norm(b_vec)))
        dlng = _r2d(np.arctan2(np.dot(b_vec, _XH), np.dot(b_vec, _ZH)))
        r2 = geoid.calc(rNorm, pclat, dlng)
        r_pclat.append(r2)
        if r1 < r2:
            hit = True
            break
    if not hit:
        return None, None
    xx = np.flipud(np.array(r_zQ) - np.array(r_pclat))
    yy = np.flipud(np.array(zQ_Trial[0:len(r_zQ)]))
    zQ = np.interp(0.0, xx, yy)
    bq = np.array([b[0], b[1], zQ])
    return rNorm * rotate2planet(rotate, tip, bq), bq


def compute_ds(req, nr, b, Req, Rpol, orientation=(0.0, 0.0), gtype='ellipse', limb='shape', gravity=None):
    """raypath.py:108-273.  req/nr: equatorial radius and refractive index per layer.
    gtype='gravity': `gravity` = dict(GM=GM profile per layer, Jn, RJ, omega_m, vwlat, vwdat).

    Returns dict(ds, layer4ds, r4ds, tip, rotate, rNorm) with ds=None when the ray misses.
    """
    out = dict(ds=None, layer4ds=None, r4ds=None, tip=None, rotate=None, rNorm=None)
    rNorm = req[0]
    if (b[0]**2 + b[1]**2) >= 1.0:
        return out
    mu = np.sqrt(1.0 - b[0]**2 - b[1]**2)
    f = 1.0 - Rpol / Req
    tip, rotate = compute_aspect(orientation, f)
    if gtype == 'gravity':
        geoid = gravity.get('table') or Geoid(req, gravity['GM'], gravity['Jn'], gravity['RJ'], gravity['omega_m'],
                                              gravity['vwlat'], gravity['vwdat'])
    else:
        geoid = Ellipse(gtype, Req, Rpol)
    edge, bq = find_edge(b, rNorm, tip, rotate, geoid)
    if edge is None:
        return out
    pclat = _r2d(np.arcsin(np.dot(edge, _YH) / np.linalg.norm(edge)))
    dlng = _r2d(np.arctan2(np.dot(edge, _XH), np.dot(edge, _ZH)))
    geoid.calc(rNorm, pclat, dlng)

    s = [rotate2planet(rotate, tip, np.array([0.0, 0.0, -1.0]))]
    n = [geoid.n]
    r = [geoid.r]
    with np.errstate(invalid='ignore'):
        t_inc = [np.arccos(-np.dot(s[-1], n[-1]))]
        nratio = nr[0] / nr[1]
        t_tran = [np.arcsin(nratio * np.sin(t_inc[-1]))]
    ds, layer4ds, r4ds = [], [], []
    i = 0
    layer = 0
    L = len(req)
    with np.errstate(invalid='ignore'):
        while True:
            s.append(nratio * s[i] + 1 * (nratio * np.cos(t_inc[i]) * n[i] - np.cos(t_tran[i]) * n[i]))
            rNow = geoid.rmag
            rNext = geoid.calc(req[layer + 1], pclat, dlng)
            rdots = np.dot(r[i], s[i + 1])
            dsm = -rdots - np.sqrt(rdots**2.0 + rNext**2.0 - rNow**2.0)
            ds_step = dsm                                   # direction is always 'ingress'
            if ds_step < 0.0:                               # raypath.py:212-216
                break
            if limb == 'sec':                               # raypath.py:218-219
                ds_step = abs(rNext - rNow) / mu
            ds.append(ds_step)
            layer4ds.append(layer)
            r4ds.append(rNow)
            rnext = r[i] + ds_step * s[i + 1]
            pclat = _r2d(np.arcsin(np.dot(rnext, _YH) / np.linalg.norm(rnext)))
            dlng = _r2d(np.arctan2(np.dot(rnext, _XH), np.dot(rnext, _ZH)))
            geoid.calc(req[layer + 1], pclat, dlng)
            r.append(rnext)
            n.append(geoid.n)
            layer += 1
            t_inc.append(np.arccos(-1 * np.dot(s[i + 1], n[i + 1])))
            i += 1
            if layer + 1 >= L:                              # IndexError exit, raypath.py:255-264
                break
            nratio = 1.0                                    # raypath.py:257
            t_tran.append(np.arcsin(nratio * np.sin(t_inc[-1])))
    out.update(ds=np.array(ds), layer4ds=np.array(layer4ds, dtype=np.int64), r4ds=np.array(r4ds),
               tip=tip, rotate=rotate, rNorm=rNorm)
    return out


def image_grid(bstep):
    """set_utils.py:65-77: pixel coordinates of a full image, rows of constant y, x fastest."""
    grid = -1.0 * np.flipud(np.arange(bstep, 1.5 + bstep, bstep))
    grid = np.concatenate((grid, np.arange(0.0, 1.5 + bstep, bstep)))
    return grid
