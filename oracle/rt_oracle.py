"""Oracle (TEST INFRASTRUCTURE, not product code): brightness-temperature integration on the CPU.

numpy restatement of Brightness.single (brightness.py:30-126): trapezoid optical depth,
weighting function W = a*exp(-tau) (point rays) or 2*a*E2(tau) (disc average, scipy.special.expn,
brightness.py:95-98), integral of W and of T*W, normalisation by the integrated weight and the
T_cmb floor (brightness.py:107-117).  Vectorised over frequency only; the segment recurrence is
kept sequential exactly as in the reference.

Pinned by tests/test_oracle_golden.py (scripts/benchmark.py:20-24 table + reference-generated
vectors).
"""
import numpy as np
from scipy.special import expn

T_CMB = 2.725          # utils.py:77
KM_TO_CM = 1.0E5       # utils.Units['km'] / utils.Units['cm'] (brightness.py:66)


def integrate_ray(ds, layer4ds, alpha_layers, T_layers, disc_average=False, return_profiles=False):
    """alpha_layers[F, L] in cm^-1; ds[S] in km.  Returns Tb[F] (and profiles when asked).

    ds is None (ray misses the planet) -> [T_cmb]*F (brightness.py:46-51).
    """
    F = alpha_layers.shape[0]
    if ds is None:
        return np.full(F, T_CMB)
    tau = np.zeros(F)
    W = np.zeros(F)
    Tb_lyr = np.zeros(F)
    integrated_W = np.zeros(F)
    if return_profiles:
        taus, Ws, Tbs = [tau.copy()], [W.copy()], [Tb_lyr.copy()]
    with np.errstate(invalid='ignore', over='ignore'):
        for i in range(len(ds) - 1):
            dscm = ds[i] * KM_TO_CM
            ii = layer4ds[i]
            ii1 = layer4ds[i + 1]
            a1 = alpha_layers[:, ii1]
            a0 = alpha_layers[:, ii]
            dtau = (a0 + a1) * dscm / 2.0
            tau = tau + dtau
            if disc_average:
                Wn = 2.0 * a1 * expn(2, tau)
            else:
                Wn = a1 * np.exp(-tau)
            integrated_W = integrated_W + (Wn + W) * dscm / 2.0
            Tb_lyr = Tb_lyr + (T_layers[ii1] * Wn + T_layers[ii] * W) * dscm / 2.0
            W = Wn
            if return_profiles:
                taus.append(tau.copy()), Ws.append(W.copy()), Tbs.append(Tb_lyr.copy())
    Tb = np.where(Tb_lyr < T_CMB, T_CMB, Tb_lyr / integrated_W)
    if return_profiles:
        return Tb, dict(tau=np.array(taus).T, W=np.array(Ws).T, Tb_lyr=np.array(Tbs).T,
                        integrated_W=integrated_W)
    return Tb


def integrate_ray_doppler(ds, layer4ds, doppler, freqs, alpha_at, T_layers, disc_average=False):
    """Brightness.single with config Doppler (brightness.py:80-96): step i evaluates the lower node's absorption a1 at
    f / doppler[i] and the upper node's a0 at f / doppler[i + 1] (as the reference writes it); everything below those
    two lines is the plain loop.  alpha_at(layer, freq_vector) -> total absorption [F] of that layer (the sum over the
    constituents of Alpha.get_alpha_from_calc, alpha.py:194-216: what the renamed `alpha.get_alpha` call asks for --
    tests/golden/make_golden.py section `doppler` restores that name to run the reference's branch).  Returns Tb[F]."""
    freqs = np.asarray(freqs, dtype=np.float64)
    F = len(freqs)
    if ds is None:
        return np.full(F, T_CMB)
    tau, W, Tb_lyr, integrated_W = np.zeros(F), np.zeros(F), np.zeros(F), np.zeros(F)
    with np.errstate(invalid='ignore', over='ignore'):
        for i in range(len(ds) - 1):
            dscm = ds[i] * KM_TO_CM
            ii, ii1 = layer4ds[i], layer4ds[i + 1]
            a1 = alpha_at(ii1, freqs / doppler[i])
            a0 = alpha_at(ii, freqs / doppler[i + 1])
            tau = tau + (a0 + a1) * dscm / 2.0
            Wn = 2.0 * a1 * expn(2, tau) if disc_average else a1 * np.exp(-tau)
            integrated_W = integrated_W + (Wn + W) * dscm / 2.0
            Tb_lyr = Tb_lyr + (T_layers[ii1] * Wn + T_layers[ii] * W) * dscm / 2.0
            W = Wn
    return np.where(Tb_lyr < T_CMB, T_CMB, Tb_lyr / integrated_W)
