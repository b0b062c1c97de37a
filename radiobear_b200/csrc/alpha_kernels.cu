// radiobear_b200 -- hot path A: per-layer x per-frequency absorption on sm_100a.
//
// One CTA owns one atmosphere layer (blockIdx.y) and a block of frequency groups (blockIdx.x).
//   phase 1  warp 0: the ~30 pow() calls a layer needs (one per lane)            -> smem
//   phase 2  all threads: per-(layer, line) hoisted terms (1 exp per line)        -> smem line tables
//   phase 3  each warp: lanes = 32*FPT consecutive frequencies, loop over the line tables
//            (broadcast LDS.128), FP64 FMA + MUFU.RCP64H/Newton reciprocal, FP64 accumulators;
//            when F is small the lines are split over K warps ("slices") and reduced through smem
//   phase 4  epilogue: continuum terms (H2 CIA, H2O continuum, clouds), NH3 pressure blend / clamp,
//            per-constituent scale-sum (alpha.py:151-192), coalesced stores to the [L][F] slab.
//
// Line shapes after hoisting (x = f^2; all prefactors folded into the table entries):
//   Ben-Reuven  (nh3 inversion, nh3_sjs, ph3, co-vvw):  x * (Bp*x + Cp) / ((x-A)^2 + G4*x)     4 doubles/line
//   Gross       (nh3 rot, nh3 v2, h2s [zeta==gamma]):   x *  N          / ((x-A)^2 + G4*x)     3 doubles/line
//   VVW         (h2o, 15 lines):                         x * cS*(df/((f-f0)^2+df^2) + df/((f+f0)^2+df^2) - 2*base)
// Reference formulas: nh3_hs.py:172-312, nh3_dbs.py:135-313, nh3_sjs.py:26-128, h2s_ddb.py:42-87,
// ph3_jh.py:64-108, h2o_bk.py:65-187, h2_jj_ddb.py:7-38, h2_jj.py:7-22, clouds_idp.py:6-101, co_ddb.py:22-99.
#include "rb_common.cuh"
#include <cmath>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- formalism constants --------------------------------------------------------------------
// nh3 inversion-line parameter sets [family: hs, dbs][band: lo (f<=30), hi (f>30)]
// order: gnu_H2 gnu_He gnu_NH3 GAMMA_H2 GAMMA_He GAMMA_NH3 zeta_H2 zeta_He zeta_NH3 Z_H2 Z_He Z_NH3 d Con
__constant__ double c_nh3inv[2][2][14] = {
    {{1.640, 0.75, 0.852, 0.7756, 0.666, 1.0, 1.262, 0.3, 0.5296, 0.7964, 0.667, 1.554, -0.0498, 0.9301},
     {1.7465, 0.9779, 0.7298, 0.8202, 1.0, 1.0, 1.2163, 0.0291, 0.5152, 0.8873, 0.8994, 2.0 / 3.0, -0.0627, 0.9862}},
    {{1.6937, 0.6997, 0.7523, 0.8085, 1.0, 1.0, 1.3263, 0.1607, 0.6162, 0.8199, 0.0, 1.3832, -0.0139, 0.9619},
     {1.7465, 0.9779, 0.7298, 0.8202, 1.0, 1.0, 1.2163, 0.0291, 0.5152, 0.8873, 0.8994, 2.0 / 3.0, -0.0627, 0.9862}}};

constexpr double kGHz = 29.9792458;          // cm^-1 <-> GHz (nh3_hs.py:45, nh3_sjs.py:9)
constexpr double kDb = 434294.5;             // cm^-1 -> dB/km
constexpr double kHcOverKb = 19.858252418E-24 / 1.38E-23;       // nh3_hs.py:50-51
constexpr double kCoefNH3 = 1.0E6 * 6.02297E23 / 8.31432E7;     // nh3_hs.py:52-56
constexpr double kCoefGeisa = 7.244E+21;     // nh3_sjs.py:6
constexpr double kHck = 1.438396;            // nh3_sjs.py:8
constexpr double kPi = 3.141592653589793;

// pow() jobs of phase 1 (index into s_pow[])
enum {
  PW_LO_GH2 = 0, PW_LO_GHE, PW_LO_ZH2, PW_LO_ZHE, PW_LO_GNH3, PW_LO_ZNH3,
  PW_HI_GH2, PW_HI_GHE, PW_HI_ZH2, PW_HI_ZHE, PW_HI_GNH3, PW_HI_ZNH3,
  PW_TDIV_35, PW_TDIV_0873, PW_TDIV_23, PW_TDIV_073, PW_TDIV_05716,
  PW_T296_23, PW_T296_35, PW_SPILKER, PW_T296_07,
  PW_TH_25, PW_TH_12, PW_TH_3, PW_H2_312, PW_H2_224, PW_H2_334, PW_H2_E27, PW_H2_E055, PW_H2_N25,
  PW_COUNT
};

struct AlphaK {
  int L, F, C;
  int form[RB_MAX_CONSTITUENTS];
  const double* freqs;
  int freq_stride;   // 0: freqs[F] for every layer; F: freqs[L][F] (rb_alpha_desc::freqs_per_layer)
  const double* T;
  const double* P;
  const double* gas[RB_NUM_GAS];     // row pointers (nullptr = absent -> mixing ratio 0)
  const double* cloud[RB_NUM_CLD];
  unsigned cloud_flags;
  int h2state, coshape, units;
  const double* scale;
  double* out_total;
  double* out_cube;
  // layer-sharded multi-GPU runs: this rank's layers are rows [peer_row0, peer_row0 + L) of a full slab that every GPU
  // holds; the epilogue stores each value into all of them over NVLink (peer pointers of a symmetric allocation)
  // instead of into out_total -- the all_gather happens inside the kernel, tile by tile behind the line sums
  double* peer_out[RB_MAX_PEERS];
  int n_peer;
  long long peer_row0;
  const int* order;   // launch order of the layers (alpha_order_kernel) or null: CTA y computes layer order[y]
  const double* cat[RB_NUM_CATALOGS];
  int ncat[RB_NUM_CATALOGS];
  // which families are present and the constituent slot each one fills (-1 = absent)
  int nh3_form, slot_nh3, slot_h2s, slot_ph3, slot_h2o, slot_h2, h2_form, slot_cld, slot_co;
  int nh3_family;  // 0 hs, 1 dbs (low-pressure formalism)
  // frequency classes present (host scan)
  int any_lo, any_hi, any_S, any_J, any_I;
  // tiling
  int fpt, K, ngroups;
  // smem layout (offsets in doubles)
  int off_inv_lo, off_inv_hi, off_rot_ag, off_rot_n, off_v2_ag, off_v2_n, off_sjs_S, off_sjs_J, off_sjs_I,
      off_h2s_ag, off_h2s_n, off_ph3, off_h2o, off_co, off_red;
};

// ---- H2 CIA from Orton's tables (h2_orton.py:126-222) --------------------------------------------
// tab: RB_CAT_H2_ORTON, [121][F]; fi: frequency index; xx = f^2; th312.. = (273/T)^3.12 / 2.24 / 3.34.
// Three temperature branches like the reference: below the table a T^4 extrapolation through its first two
// temperatures (:150-178), inside the not-a-knot cubic spline (:200-214), above h2_jj scaled to the table
// at its last temperature (:179-199; the dB/km factors of both h2_jj calls cancel).
__device__ __forceinline__ double h2_orton_alpha(const double* __restrict__ tab, int F, int fi, double T, double xx,
                                                 double P_h2, double P_he, double P_ch4, double th312, double th224,
                                                 double th334) {
  constexpr int NT = 10, TSTRIDE = 37;
  constexpr double T0 = 273.0, atm = 1.01325;
  auto col = [&](int c) { return tab[(size_t)c * F + fi]; };
  auto val = [&](int t, int j) { return col(NT + t * TSTRIDE + j); };
  auto pair = [&](double ah2, double ahe, double ach4, double Tt) {
    const double r = T0 / Tt;
    return ((P_h2 / atm) * (ah2 * P_h2 / atm + ahe * P_he / atm + ach4 * P_ch4 / atm) * (r * r));
  };
  const double Tlo = col(0), Thi = col(NT - 1);
  if (T < Tlo) {
    const double T1 = col(1);
    const double v0 = pair(val(0, 0), val(1, 0), val(2, 0), Tlo), v1 = pair(val(0, 1), val(1, 1), val(2, 1), T1);
    const double X1 = Tlo * Tlo * Tlo * Tlo;
    const double AQ = -1.0 * (v1 - v0) / (T1 * T1 * T1 * T1 - X1);
    const double CQ = v0 + AQ * X1;
    return CQ - AQ * (T * T * T * T);
  }
  if (T > Thi) {
    const double anear = pair(val(0, NT - 1), val(1, NT - 1), val(2, NT - 1), Thi);
    const double thn = T0 / Thi;
    const double cf = 3.9522E-14 * xx * P_h2;   // h2_jj.py:13-18 at T and at Thi
    const double jj = cf * (P_h2 * th312 + 1.382 * P_he * th224 + 9.322 * P_ch4 * th334);
    const double jjnear = cf * (P_h2 * pow(thn, 3.12) + 1.382 * P_he * pow(thn, 2.24) + 9.322 * P_ch4 * pow(thn, 3.34));
    return jj * (anear / jjnear);
  }
  int kk = 0;
#pragma unroll
  for (int i = 1; i < NT - 1; ++i) kk += (T >= col(i)) ? 1 : 0;   // interval [Ttab[kk], Ttab[kk+1]]
  const double dt = T - col(kk);
  double a[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int c = NT + t * TSTRIDE + NT + kk * 3;
    a[t] = fma(dt, fma(dt, fma(dt, col(c + 2), col(c + 1)), col(c)), val(t, kk));
  }
  return pair(a[0], a[1], a[2], T);
}

// ---- line loops --------------------------------------------------------------------------------
template <int FPT, int NEWTON>
__device__ __forceinline__ void loop_br4(const double4* __restrict__ tab, int n, int k, int K, const double (&x)[FPT],
                                         double (&acc)[FPT]) {
#pragma unroll 4
  for (int i = k; i < n; i += K) {
    const double4 e = tab[i];  // A, G4, Bp, Cp
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
      const double t = x[j] - e.x;
      const double den = fma(t, t, e.y * x[j]);
      const double num = fma(e.z, x[j], e.w);
      acc[j] = fma(num, rb_rcp<NEWTON>(den), acc[j]);
    }
  }
}

// per-lane table pointers (warps that straddle a frequency-class boundary)
template <int FPT, int NEWTON>
__device__ __forceinline__ void loop_br4_sel(const double4* const (&tab)[FPT], int n, int k, int K,
                                             const double (&x)[FPT], double (&acc)[FPT]) {
#pragma unroll 2
  for (int i = k; i < n; i += K) {
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
      if (tab[j] != nullptr) {
        const double4 e = tab[j][i];
        const double t = x[j] - e.x;
        const double den = fma(t, t, e.y * x[j]);
        const double num = fma(e.z, x[j], e.w);
        acc[j] = fma(num, rb_rcp<NEWTON>(den), acc[j]);
      }
    }
  }
}

template <int FPT, int NEWTON>
__device__ __forceinline__ void loop_gr3(const double2* __restrict__ ag, const double* __restrict__ nn, int n, int k,
                                         int K, const double (&x)[FPT], double (&acc)[FPT]) {
#pragma unroll 4
  for (int i = k; i < n; i += K) {
    const double2 e = ag[i];  // A, G4
    const double N = nn[i];
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
      const double t = x[j] - e.x;
      const double den = fma(t, t, e.y * x[j]);
      acc[j] = fma(N, rb_rcp<NEWTON>(den), acc[j]);
    }
  }
}

// nh3_sjs transition band 26 < f < 34 GHz: broadening parameters depend on f (nh3_sjs.py:100-111)
// table entry: gS, gJ, zS, zJ, F0 (= f0 + delta), Sp
template <int NEWTON>
__device__ __forceinline__ double loop_sjs_interp(const double* __restrict__ tab, int n, int k, int K, double f,
                                                  double x) {
  const double flfh = (26.0 - 34.0) / (f - 26.0);
  double acc = 0.0;
  for (int i = k; i < n; i += K) {
    const double* e = tab + 6 * i;
    const double g = e[0] + (e[0] - e[1]) / flfh;
    const double z = e[2] + (e[2] - e[3]) / flfh;
    const double A = e[4] * e[4] + g * g - z * z;
    const double num = (g - z) * x + (g + z) * A;
    const double t = x - A;
    const double den = fma(t, t, 4.0 * g * g * x);
    acc = fma(e[5] * num, rb_rcp<NEWTON>(den), acc);
  }
  return acc;
}

// ---- complex helpers for clouds / CO Voigt -----------------------------------------------------
struct cplx {
  double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  // Smith's algorithm (what numpy/C99 use) -- avoids overflow of |b|^2 for the degree-7 Voigt rational
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im / b.re, d = b.re + b.im * r;
    return {(a.re + a.im * r) / d, (a.im - a.re * r) / d};
  }
  const double r = b.re / b.im, d = b.re * r + b.im;
  return {(a.re * r + a.im) / d, (a.im * r - a.re) / d};
}

// clouds_idp.py:72-101 -- complex permittivity of water / ice
__device__ cplx water_eps(double f, double T) {
  const double Tc = T - 273.0;
  const double fHz = f * 1.0E9;
  double E1, E2;
  if (Tc >= 0.0) {
    const double RelT = 1.1109E-10 - Tc * 3.824E-12 + (Tc * Tc) * 6.938E-14 - (Tc * Tc * Tc) * 5.096E-16;
    double E0 = 88.045 - 0.4147 * Tc + (Tc * Tc) * 6.295E-4 + (Tc * Tc * Tc) * 1.075E-5;
    if (E0 < 0.0) E0 = 0.0;
    const double EINF = 4.9;
    const double w = fHz * RelT;
    E1 = EINF + (E0 - EINF) / (1.0 + w * w);
    E2 = w * (E0 - EINF) / (1.0 + w * w);
    if (E2 < 0.0) E2 = 0.0;
  } else {
    const double FR[9] = {1.0E8, 3.0E8, 1.0E9, 2.0E9, 3.0E9, 5.0E9, 1.0E10, 3.0E10, 1.0E11};
    const double EI[9] = {8.0E-3, 1.5E-3, 8.0E-4, 1.0E-3, 1.2E-3, 1.5E-3, 3.0E-3, 8.0E-3, 2.0E-2};
    E1 = 3.17;
    int j = 0;
    for (j = 0; j < 8; ++j)
      if (FR[j + 1] >= fHz) break;
    if (j > 7) j = 7;  // python: loop variable keeps its last value when nothing breaks
    const double LF = log10(fHz), LF0 = log10(FR[j]), LF1 = log10(FR[j + 1]);
    const double DLF = (LF - LF0) / (LF1 - LF0);
    const double X0 = log10(EI[j]), X1 = log10(EI[j + 1]);
    E2 = pow(10.0, X0 + DLF * (X1 - X0));
  }
  return {E1, -E2};
}
// clouds_idp.py:61-65: 3 k fraction * (-Im((e-1)/(e+2)))
__device__ __forceinline__ double acloud(double k, double fraction, cplx e) {
  const cplx K = cdiv({e.re - 1.0, e.im}, {e.re + 2.0, e.im});
  return 3.0 * k * fraction * (-K.im);
}

__device__ __forceinline__ double ld0(const double* p, int l) { return p ? p[l] : 0.0; }

// ================================================================================================
#ifndef RB_ALPHA_FPT4_CTAS
#define RB_ALPHA_FPT4_CTAS 1
#endif
template <int FPT, int NEWTON>
__global__ void __launch_bounds__(kThreads, FPT >= 4 ? RB_ALPHA_FPT4_CTAS : 2) alpha_lines_kernel(const __grid_constant__ AlphaK k) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_pow[PW_COUNT];

  const int l = k.order ? k.order[blockIdx.y] : (int)blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  const double T = k.T[l];
  const double P = k.P[l];
  const double P_h2 = P * ld0(k.gas[RB_GAS_H2], l);
  const double P_he = P * ld0(k.gas[RB_GAS_HE], l);
  const double P_ch4 = P * ld0(k.gas[RB_GAS_CH4], l);
  const double P_nh3 = P * ld0(k.gas[RB_GAS_NH3], l);
  const double P_h2o = P * ld0(k.gas[RB_GAS_H2O], l);
  const double P_h2s = P * ld0(k.gas[RB_GAS_H2S], l);
  const double P_ph3 = P * ld0(k.gas[RB_GAS_PH3], l);
  const double P_co = P * ld0(k.gas[RB_GAS_CO], l);
  const double Tdiv = 300.0 / T;

  // which NH3 branches this layer needs and how they blend: a = wb * a_b + (1 - wb) * a_a
  //   nh3_hs_sjs / nh3_dbs_sjs (nh3_hs_sjs.py:10-25): (a_a, a_b) = (low, sjs), wb = (P-400)/1600 in 400..2000 bar
  //   nh3_sjsd (nh3_sjsd.py:6-24):                    (a_a, a_b) = (sjs, kd),  triangular wb peaking at 35 bar
  const int nform = k.nh3_form;
  const bool nh3_blend = (nform == RB_F_NH3_HS_SJS || nform == RB_F_NH3_DBS_SJS);
  const bool sjsd_mix = (nform == RB_F_NH3_SJSD) && !(P < 10.0 || P > 100.0);
  const bool use_low = (nform == RB_F_NH3_HS || nform == RB_F_NH3_DBS || nform == RB_F_NH3_KD) ||
                       (nh3_blend && !(P > 2000.0)) || sjsd_mix;
  const bool use_sjs = (nform == RB_F_NH3_SJS || nform == RB_F_NH3_BG || nform == RB_F_NH3_SJSD) ||
                       (nh3_blend && !(P < 400.0));
  // inversion-line constants: tabulated per band (hs, dbs) or switched on pressure (kd, nh3_kd.py:160-195)
  double cl[14], ch[14];
  if (k.nh3_family < 2) {
#pragma unroll
    for (int i = 0; i < 14; ++i) { cl[i] = c_nh3inv[k.nh3_family][0][i]; ch[i] = c_nh3inv[k.nh3_family][1][i]; }
  } else {
    const double hiP[14] = {1.6361, 0.4555, 0.7298, 0.8, 0.5, 1.0, 1.1313, 0.1, 0.5152, 0.6234, 0.5, 2.0 / 3.0, 0.2, 1.3746};
    const double loP[14] = {1.7465, 0.9779, 0.7298, 0.8202, 1.0, 1.0, 1.2163, 0.0291, 0.5152, 0.8873, 0.8994, 2.0 / 3.0,
                            -0.0627, 0.9862};
    const double w = (15.0 - 3.0 - P) / (5.0 + 3.0);
#pragma unroll
    for (int i = 0; i < 14; ++i) {
      double v;
      if (P > 20.0) v = hiP[i];
      else if (P <= 12.0) v = loP[i];
      else v = (i == 2 || i == 5 || i == 8 || i == 11) ? loP[i] : loP[i] + (loP[i] - hiP[i]) * w;
      cl[i] = ch[i] = v;
    }
  }

  // ---- phase 1: pow() jobs, one per lane of warp 0 -------------------------------------------
  if (warp == 0 && lane < PW_COUNT) {
    const double t295 = 295.0 / T, t296 = 296.0 / T, th = 273.0 / T;
    double base = 1.0, ex = 0.0;
    switch (lane) {
      case PW_LO_GH2: base = Tdiv; ex = cl[3]; break;
      case PW_LO_GHE: base = Tdiv; ex = cl[4]; break;
      case PW_LO_ZH2: base = Tdiv; ex = cl[9]; break;
      case PW_LO_ZHE: base = Tdiv; ex = cl[10]; break;
      case PW_LO_GNH3: base = t295; ex = cl[5]; break;
      case PW_LO_ZNH3: base = t295; ex = cl[11]; break;
      case PW_HI_GH2: base = Tdiv; ex = ch[3]; break;
      case PW_HI_GHE: base = Tdiv; ex = ch[4]; break;
      case PW_HI_ZH2: base = Tdiv; ex = ch[9]; break;
      case PW_HI_ZHE: base = Tdiv; ex = ch[10]; break;
      case PW_HI_GNH3: base = t295; ex = ch[5]; break;
      case PW_HI_ZNH3: base = t295; ex = ch[11]; break;
      case PW_TDIV_35: base = Tdiv; ex = 3.5; break;
      case PW_TDIV_0873: base = Tdiv; ex = 0.8730; break;
      case PW_TDIV_23: base = Tdiv; ex = 2.0 / 3.0; break;
      case PW_TDIV_073: base = Tdiv; ex = 0.73; break;
      case PW_TDIV_05716: base = Tdiv; ex = 0.5716; break;
      case PW_T296_23: base = t296; ex = 2.0 / 3.0; break;
      case PW_T296_35: base = t296; ex = 3.5; break;
      case PW_SPILKER: base = exp(9.024 - T / 20.3) - 0.9918 + P_h2; ex = 8.79 * exp(-T / 83.0); break;
      case PW_T296_07: base = t296; ex = 0.7; break;
      case PW_TH_25: base = Tdiv; ex = 2.5; break;
      case PW_TH_12: base = Tdiv; ex = 12.0; break;
      case PW_TH_3: base = Tdiv; ex = 3.0; break;
      case PW_H2_312: base = th; ex = 3.12; break;
      case PW_H2_224: base = th; ex = 2.24; break;
      case PW_H2_334: base = th; ex = 3.34; break;
      case PW_H2_E27: base = T / 55.0; ex = 2.7; break;
      case PW_H2_E055: base = T / 120.0; ex = 0.55; break;
      case PW_H2_N25: base = T / 40.0; ex = 2.5; break;
    }
    s_pow[lane] = pow(base, ex);
  }
  __syncthreads();

  // ---- phase 2: line tables ---------------------------------------------------------------------
  if (k.slot_nh3 >= 0 && use_low) {
    const double expfac = -(1.0 / T - 1.0 / 300.0) * kHcOverKb;
    const double pcommon = kCoefNH3 * (P_nh3 / 300.0) * s_pow[PW_TDIV_35] * kGHz;
    {  // inversion lines, Ben-Reuven (nh3_hs.py:172-226)
      const double* fo = k.cat[RB_CAT_NH3_INV];
      const int n = k.ncat[RB_CAT_NH3_INV];
      const double *Io = fo + n, *Eo = fo + 2 * n, *g0 = fo + 3 * n;
      const double lo_g0 = cl[0] * P_h2 * s_pow[PW_LO_GH2] + cl[1] * P_he * s_pow[PW_LO_GHE];
      const double lo_gN = cl[2] * P_nh3 * s_pow[PW_LO_GNH3];
      const double lo_z0 = cl[6] * P_h2 * s_pow[PW_LO_ZH2] + cl[7] * P_he * s_pow[PW_LO_ZHE];
      const double lo_zN = cl[8] * P_nh3 * s_pow[PW_LO_ZNH3];
      const double hi_g0 = ch[0] * P_h2 * s_pow[PW_HI_GH2] + ch[1] * P_he * s_pow[PW_HI_GHE];
      const double hi_gN = ch[2] * P_nh3 * s_pow[PW_HI_GNH3];
      const double hi_z0 = ch[6] * P_h2 * s_pow[PW_HI_ZH2] + ch[7] * P_he * s_pow[PW_HI_ZHE];
      const double hi_zN = ch[8] * P_nh3 * s_pow[PW_HI_ZNH3];
      const double lo_pref = cl[13] * pcommon * (2.0 / kPi), hi_pref = ch[13] * pcommon * (2.0 / kPi);
      double4* tlo = reinterpret_cast<double4*>(smem + k.off_inv_lo);
      double4* thi = reinterpret_cast<double4*>(smem + k.off_inv_hi);
      for (int i = tid; i < n; i += kThreads) {
        const double f0 = fo[i];
        const double ST = Io[i] * exp(expfac * Eo[i]) / (f0 * f0);
        if (k.any_lo) {
          const double g = lo_g0 + lo_gN * g0[i], z = lo_z0 + lo_zN * g0[i];
          const double F0 = f0 + cl[12] * g;
          const double A = F0 * F0 + g * g - z * z;
          const double Sp = lo_pref * ST;
          tlo[i] = make_double4(A, 4.0 * g * g, Sp * (g - z), Sp * (g + z) * A);
        }
        if (k.any_hi) {
          const double g = hi_g0 + hi_gN * g0[i], z = hi_z0 + hi_zN * g0[i];
          const double F0 = f0 + ch[12] * g;
          const double A = F0 * F0 + g * g - z * z;
          const double Sp = hi_pref * ST;
          thi[i] = make_double4(A, 4.0 * g * g, Sp * (g - z), Sp * (g + z) * A);
        }
      }
    }
    {  // rotational lines, Gross (nh3_hs.py:228-265)
      const double* fo = k.cat[RB_CAT_NH3_ROT];
      const int n = k.ncat[RB_CAT_NH3_ROT];
      const double *Io = fo + n, *Eo = fo + 2 * n, *gN = fo + 3 * n, *gH = fo + 4 * n, *gHe = fo + 5 * n;
      const double c1 = 0.2984 * P_h2 * s_pow[PW_TDIV_0873], c2 = 0.75 * P_he * s_pow[PW_TDIV_23],
                   c3 = 3.1789 * P_nh3 * Tdiv;
      const double pref = 2.4268 * pcommon * (4.0 / kPi);
      double2* ag = reinterpret_cast<double2*>(smem + k.off_rot_ag);
      double* nn = smem + k.off_rot_n;
      for (int i = tid; i < n; i += kThreads) {
        const double g = c1 * gH[i] + c2 * gHe[i] + c3 * gN[i];
        const double ST = Io[i] * exp(expfac * Eo[i]);
        ag[i] = make_double2(fo[i] * fo[i], 4.0 * g * g);
        nn[i] = pref * ST * g;
      }
    }
    {  // v2 roto-vibrational lines, Gross (nh3_hs.py:267-302)
      const double* fo = k.cat[RB_CAT_NH3_V2];
      const int n = k.ncat[RB_CAT_NH3_V2];
      const double *Io = fo + n, *Eo = fo + 2 * n;
      const double g = (P_h2 * 1.4) * s_pow[PW_TDIV_073] + (P_he * 0.68) * s_pow[PW_TDIV_05716] + (P_nh3 * 9.5) * Tdiv;
      const double pref = 1.1206 * pcommon * (4.0 / kPi) * g;
      double2* ag = reinterpret_cast<double2*>(smem + k.off_v2_ag);
      double* nn = smem + k.off_v2_n;
      for (int i = tid; i < n; i += kThreads) {
        ag[i] = make_double2(fo[i] * fo[i], 4.0 * g * g);
        nn[i] = pref * Io[i] * exp(expfac * Eo[i]);
      }
    }
  }
  if (k.slot_nh3 >= 0 && use_sjs) {  // nh3_sjs.py:26-128
    const double th = s_pow[PW_T296_23];
    // Joiner / Spilker broadening sets (nh3_sjs.py:43-77)
    double SG_H2 = 1.690, SG_He = 0.750, SG_N = 0.60, SZ_H2 = 1.35, SZ_He = 0.30, SZ_N = 0.20;
    double GH2a = s_pow[PW_SPILKER];
    if (!(GH2a < 1E-12)) {
      GH2a = 2.122 * exp(-T / 116.8) / GH2a;
      GH2a = 2.34 * (1.0 - GH2a);
      SG_H2 = GH2a;
      SG_He = 0.46 + T / 3000.0;
      SG_N = 0.74;
      SZ_H2 = 5.7465 - 7.7644 * GH2a + 9.1931 * GH2a * GH2a - 5.6816 * GH2a * GH2a * GH2a +
              1.2307 * GH2a * GH2a * GH2a * GH2a;
      SZ_He = 0.28 - T / 1750.0;
      SZ_N = 0.50;
    }
    const double Sg0 = th * (SG_H2 * P_h2 + SG_He * P_he), SgN = th * SG_N * P_nh3;
    const double Sz0 = th * (SZ_H2 * P_h2 + SZ_He * P_he), SzN = th * SZ_N * P_nh3;
    // "Joiner" set; nh3_bg (nh3_bg.py:42-49) is the same Ben-Reuven sum with its own constants at every
    // frequency and without the deep-atmosphere pressure scale
    const bool bg = (nform == RB_F_NH3_BG);
    const double Jg0 = th * ((bg ? 2.318 : 1.690) * P_h2 + (bg ? 0.790 : 0.750) * P_he), JgN = th * (bg ? 0.750 : 0.6) * P_nh3;
    const double Jz0 = th * ((bg ? 1.920 : 1.350) * P_h2 + 0.300 * P_he), JzN = th * (bg ? 0.490 : 0.2) * P_nh3;
    const double delta = -0.45 * P_nh3;
    const double expfac = -((1.0 / T) - (1.0 / 296.0)) * kHck;
    const double pref = kCoefGeisa * (P_nh3 / 296.0) * s_pow[PW_T296_35] * (bg ? 1.0 : (1.0 + P / 1.0E5)) * kGHz * (2.0 / kPi);
    const double* f0p = k.cat[RB_CAT_NH3_SJS];
    const int n = k.ncat[RB_CAT_NH3_SJS];
    const double *I0 = f0p + n, *E = f0p + 2 * n, *G0 = f0p + 3 * n;
    double4* tS = reinterpret_cast<double4*>(smem + k.off_sjs_S);
    double4* tJ = reinterpret_cast<double4*>(smem + k.off_sjs_J);
    double* tI = smem + k.off_sjs_I;
    for (int i = tid; i < n; i += kThreads) {
      const double f0 = f0p[i];
      const double Sp = pref * I0[i] * exp(expfac * E[i]) / (f0 * f0);
      const double F0 = f0 + delta;
      const double gS = Sg0 + SgN * G0[i], zS = Sz0 + SzN * G0[i];
      const double gJ = Jg0 + JgN * G0[i], zJ = Jz0 + JzN * G0[i];
      if (k.any_S) {
        const double A = F0 * F0 + gS * gS - zS * zS;
        tS[i] = make_double4(A, 4.0 * gS * gS, Sp * (gS - zS), Sp * (gS + zS) * A);
      }
      if (k.any_J) {
        const double A = F0 * F0 + gJ * gJ - zJ * zJ;
        tJ[i] = make_double4(A, 4.0 * gJ * gJ, Sp * (gJ - zJ), Sp * (gJ + zJ) * A);
      }
      if (k.any_I) {
        double* e = tI + 6 * i;
        e[0] = gS; e[1] = gJ; e[2] = zS; e[3] = zJ; e[4] = F0; e[5] = Sp;
      }
    }
  }
  if (k.slot_h2s >= 0) {  // h2s_ddb.py:42-87: zeta == gamma -> num = 2 gamma (f0+delta)^2
    const double* f0p = k.cat[RB_CAT_H2S];
    const int n = k.ncat[RB_CAT_H2S];
    const double *I0 = f0p + n, *E = f0p + 2 * n, *GH2S = f0p + 3 * n;
    const double th = s_pow[PW_T296_07];
    const double delta = 1.28 * P_h2s;
    const double expfac = -((1.0 / T) - (1.0 / 296.0)) * kHck;
    const double pref = kCoefGeisa * (P_h2s / 296.0) * s_pow[PW_T296_35] * kGHz * (2.0 / kPi);
    double2* ag = reinterpret_cast<double2*>(smem + k.off_h2s_ag);
    double* nn = smem + k.off_h2s_n;
    for (int i = tid; i < n; i += kThreads) {
      const double f0 = f0p[i];
      const double g = th * (1.960 * P_h2 + 1.200 * P_he + GH2S[i] * P_h2s);
      const double F0 = f0 + delta;
      ag[i] = make_double2(F0 * F0, 4.0 * g * g);
      nn[i] = pref * I0[i] * exp(expfac * E[i]) / (f0 * f0) * (2.0 * g) * (F0 * F0);
    }
  }
  if (k.slot_ph3 >= 0) {  // ph3_jh.py:64-108: zeta = delta = 0
    const double* f0p = k.cat[RB_CAT_PH3];
    const int n = k.ncat[RB_CAT_PH3];
    const double *I0 = f0p + n, *E = f0p + 2 * n, *WI = f0p + 3 * n, *WF = f0p + 4 * n, *WS = f0p + 5 * n;
    const double gA = s_pow[PW_TDIV_23] * (3.2930 * P_h2 + 1.6803 * P_he), gB = Tdiv * 4.2157 * P_ph3;
    const double expfac = -((1.0 / T) - (1.0 / 300.0)) * kHck;
    const double pref = kCoefGeisa * (P_ph3 / 300.0) * s_pow[PW_TDIV_35] * kGHz * (2.0 / kPi);
    double4* t = reinterpret_cast<double4*>(smem + k.off_ph3);
    for (int i = tid; i < n; i += kThreads) {
      const double f0 = f0p[i];
      const double g = gA * WF[i] + gB * WS[i];
      const double Sp = pref * I0[i] * WI[i] * exp(expfac * E[i]) / (f0 * f0);
      const double A = f0 * f0 + g * g;
      t[i] = make_double4(A, 4.0 * g * g, Sp * g, Sp * g * A);
    }
  }
  if (k.slot_h2o >= 0) {  // h2o_bk.py:88-97: entry f0, df, cS, base
    const double* fo = k.cat[RB_CAT_H2O];
    const int n = k.ncat[RB_CAT_H2O];
    const double *Io = fo + n, *Eo = fo + 2 * n, *ws = fo + 3 * n, *xs = fo + 4 * n, *wh2 = fo + 5 * n,
                 *whe = fo + 6 * n, *xh2 = fo + 7 * n, *xhe = fo + 8 * n;
    double4* t = reinterpret_cast<double4*>(smem + k.off_h2o);
    for (int i = tid; i < n; i += kThreads) {
      const double S = Io[i] * s_pow[PW_TH_25] * exp(Eo[i] * (1.0 - Tdiv));
      const double df = ws[i] * P_h2o * pow(Tdiv, xs[i]) + wh2[i] * P_h2 * pow(Tdiv, xh2[i]) +
                        whe[i] * P_he * pow(Tdiv, xhe[i]);
      t[i] = make_double4(fo[i], df, S / (kPi * fo[i] * fo[i]), df / (562500.0 + df * df));
    }
  }
  if (k.slot_co >= 0) {  // co_ddb.py:44-58: entry f0, ITG
    const double* f0p = k.cat[RB_CAT_CO];
    const int n = k.ncat[RB_CAT_CO];
    const double *I0 = f0p + n, *E = f0p + 2 * n;
    const double expfac = -((1.0 / T) - (1.0 / 296.0)) * kHck;
    double2* t = reinterpret_cast<double2*>(smem + k.off_co);
    for (int i = tid; i < n; i += kThreads) t[i] = make_double2(f0p[i], I0[i] * exp(expfac * E[i]));
  }
  __syncthreads();

  // ---- phase 3: line sums ---------------------------------------------------------------------
  // A CTA keeps its layer's line tables for several frequency groups per warp (grid.x is sized for a few waves of
  // CTAs, not for one group per warp): at C5 the tables of a layer are built once instead of eight times, and the
  // barriers of phases 1-2, where the FP64 pipe idles, are paid once per layer.
  const int gpb = kWarps / k.K;  // frequency groups per block and trip
  const int slice = warp % k.K;
  const int K = k.K;
  const int trips = (k.ngroups + gpb * (int)gridDim.x - 1) / (gpb * (int)gridDim.x);   // the same for every warp
  for (int trip = 0; trip < trips; ++trip) {
  const int g = (trip * (int)gridDim.x + (int)blockIdx.x) * gpb + warp / k.K;
  double f[FPT], x[FPT];
  int fidx[FPT];
  bool valid[FPT];
#pragma unroll
  for (int j = 0; j < FPT; ++j) {
    fidx[j] = (g * FPT + j) * 32 + lane;
    valid[j] = (g < k.ngroups) && (fidx[j] < k.F);
    f[j] = valid[j] ? k.freqs[(size_t)l * k.freq_stride + fidx[j]] : 1.0;   // (freq_stride = F: per-layer frequencies)
    x[j] = f[j] * f[j];
  }
  double s_low[FPT], s_sjs[FPT], s_h2s[FPT], s_ph3[FPT], s_h2o[FPT], s_co[FPT];
#pragma unroll
  for (int j = 0; j < FPT; ++j) s_low[j] = s_sjs[j] = s_h2s[j] = s_ph3[j] = s_h2o[j] = s_co[j] = 0.0;

  if (g < k.ngroups) {
    if (k.slot_nh3 >= 0 && use_low) {
      const double4* tlo = reinterpret_cast<const double4*>(smem + k.off_inv_lo);
      const double4* thi = reinterpret_cast<const double4*>(smem + k.off_inv_hi);
      bool all_lo = true, all_hi = true;
#pragma unroll
      for (int j = 0; j < FPT; ++j) {
        all_lo = all_lo && (f[j] <= 30.0);
        all_hi = all_hi && (f[j] > 30.0);
      }
      all_lo = __all_sync(0xffffffffu, all_lo);
      all_hi = __all_sync(0xffffffffu, all_hi);
      const int n = k.ncat[RB_CAT_NH3_INV];
      if (all_lo) {
        loop_br4<FPT, NEWTON>(tlo, n, slice, K, x, s_low);
      } else if (all_hi) {
        loop_br4<FPT, NEWTON>(thi, n, slice, K, x, s_low);
      } else {
        const double4* sel[FPT];
#pragma unroll
        for (int j = 0; j < FPT; ++j) sel[j] = (f[j] <= 30.0) ? tlo : thi;
        loop_br4_sel<FPT, NEWTON>(sel, n, slice, K, x, s_low);
      }
      loop_gr3<FPT, NEWTON>(reinterpret_cast<const double2*>(smem + k.off_rot_ag), smem + k.off_rot_n,
                            k.ncat[RB_CAT_NH3_ROT], slice, K, x, s_low);
      loop_gr3<FPT, NEWTON>(reinterpret_cast<const double2*>(smem + k.off_v2_ag), smem + k.off_v2_n,
                            k.ncat[RB_CAT_NH3_V2], slice, K, x, s_low);
    }
    if (k.slot_nh3 >= 0 && use_sjs) {
      const double4* tS = reinterpret_cast<const double4*>(smem + k.off_sjs_S);
      const double4* tJ = reinterpret_cast<const double4*>(smem + k.off_sjs_J);
      bool all_S = true, all_J = true;
#pragma unroll
      for (int j = 0; j < FPT; ++j) {
        all_S = all_S && (f[j] <= 26.0);
        all_J = all_J && (f[j] >= 34.0);
      }
      all_S = __all_sync(0xffffffffu, all_S) && nform != RB_F_NH3_BG;
      all_J = __all_sync(0xffffffffu, all_J) || nform == RB_F_NH3_BG;
      const int n = k.ncat[RB_CAT_NH3_SJS];
      if (all_S) {
        loop_br4<FPT, NEWTON>(tS, n, slice, K, x, s_sjs);
      } else if (all_J) {
        loop_br4<FPT, NEWTON>(tJ, n, slice, K, x, s_sjs);
      } else {
        const double4* sel[FPT];
#pragma unroll
        for (int j = 0; j < FPT; ++j) sel[j] = (f[j] <= 26.0) ? tS : ((f[j] >= 34.0) ? tJ : nullptr);
        loop_br4_sel<FPT, NEWTON>(sel, n, slice, K, x, s_sjs);
#pragma unroll
        for (int j = 0; j < FPT; ++j)
          if (sel[j] == nullptr) s_sjs[j] = loop_sjs_interp<NEWTON>(smem + k.off_sjs_I, n, slice, K, f[j], x[j]);
      }
    }
    if (k.slot_h2s >= 0)
      loop_gr3<FPT, NEWTON>(reinterpret_cast<const double2*>(smem + k.off_h2s_ag), smem + k.off_h2s_n,
                            k.ncat[RB_CAT_H2S], slice, K, x, s_h2s);
    if (k.slot_ph3 >= 0)
      loop_br4<FPT, NEWTON>(reinterpret_cast<const double4*>(smem + k.off_ph3), k.ncat[RB_CAT_PH3], slice, K, x, s_ph3);
    if (k.slot_h2o >= 0) {  // vvwlinecontribution_modified, h2o_bk.py:130-187 (plain divides: 15 lines)
      const double4* t = reinterpret_cast<const double4*>(smem + k.off_h2o);
      for (int i = slice; i < k.ncat[RB_CAT_H2O]; i += K) {
        const double4 e = t[i];
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
          const double dm = f[j] - e.x, dp = f[j] + e.x;
          const double B = e.y / (dm * dm + e.y * e.y);
          const double Cc = e.y / (dp * dp + e.y * e.y);
          s_h2o[j] += e.z * (B - e.w + Cc - e.w);
        }
      }
    }
    if (k.slot_co >= 0) {  // co_ddb.py:60-90
      const double2* t = reinterpret_cast<const double2*>(smem + k.off_co);
      const double gamma = s_pow[PW_T296_07] * (1.960 * P_h2 + 1.200 * P_he + 6.000 * P_co);
      const double g2 = gamma * gamma;
      double w = (P - 0.001) / (0.1 - 0.001);
      w = w < 0.0 ? 0.0 : (w > 1.0 ? 1.0 : w);
      const bool do_voigt = (P <= 0.1) || k.coshape == 0 || k.coshape == 2;
      const bool do_vvw = (P >= 0.001) || k.coshape == 1 || k.coshape == 2;
      const double av[8] = {122.60793178, 214.38238869, 181.92853309, 93.15558046, 30.18014220, 5.91262621, 0.56418958, 0.0};
      const double bv[8] = {122.60793178, 352.73062511, 457.33447878, 348.70391772, 170.35400182, 53.99290691, 10.47985711, 1.0};
      for (int i = slice; i < k.ncat[RB_CAT_CO]; i += K) {
        const double2 e = t[i];  // f0, ITG
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
          double sV = 0.0, sW = 0.0;
          if (do_voigt) {
            const double betaD = 4.3e-7 * sqrt(T / 28.0) * f[j];
            const cplx xi = {gamma / betaD, (f[j] - e.x) / betaD};
            cplx num = {av[7], 0.0}, den = {bv[7], 0.0};
#pragma unroll
            for (int m = 6; m >= 0; --m) {
              num = cmul(num, xi); num.re += av[m];
              den = cmul(den, xi); den.re += bv[m];
            }
            const cplx val = cdiv(num, den);
            sV = kGHz * (1.0 / (sqrt(kPi) * betaD)) * val.re;
          }
          if (do_vvw) {
            const double num = gamma * x[j] + gamma * (e.x * e.x + g2);
            const double t0 = x[j] - e.x * e.x - g2;
            const double den = t0 * t0 + 4.0 * x[j] * g2;
            sW = kGHz * 2.0 * (x[j] / (e.x * e.x)) * num / (kPi * den);
          }
          double term;
          if (k.coshape == 0) term = sV;
          else if (k.coshape == 1) term = sW;
          else if (k.coshape == 2) term = sV - sW;
          else term = (w * sW + (1.0 - w) * sV) * e.y;
          s_co[j] += term;
        }
      }
    }
  }

  // ---- slice reduction through smem (K > 1: lines were split across K warps) ----------------------
  if (K > 1) {
    double* red = smem + k.off_red;  // [warp][6][FPT*32]
    double* mine = red + warp * (6 * FPT * 32);
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
      mine[(0 * FPT + j) * 32 + lane] = s_low[j];
      mine[(1 * FPT + j) * 32 + lane] = s_sjs[j];
      mine[(2 * FPT + j) * 32 + lane] = s_h2s[j];
      mine[(3 * FPT + j) * 32 + lane] = s_ph3[j];
      mine[(4 * FPT + j) * 32 + lane] = s_h2o[j];
      mine[(5 * FPT + j) * 32 + lane] = s_co[j];
    }
    __syncthreads();
    if (slice == 0) {
      for (int s = 1; s < K; ++s) {
        const double* o = red + (warp + s) * (6 * FPT * 32);
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
          s_low[j] += o[(0 * FPT + j) * 32 + lane];
          s_sjs[j] += o[(1 * FPT + j) * 32 + lane];
          s_h2s[j] += o[(2 * FPT + j) * 32 + lane];
          s_ph3[j] += o[(3 * FPT + j) * 32 + lane];
          s_h2o[j] += o[(4 * FPT + j) * 32 + lane];
          s_co[j] += o[(5 * FPT + j) * 32 + lane];
        }
      }
    }
  }
  if (K > 1 && trip + 1 < trips) __syncthreads();   // the reduction buffer is reused by the next trip
  if (slice != 0 || g >= k.ngroups) continue;

  // ---- phase 4: epilogue -------------------------------------------------------------------------
  const double unit = (k.units == RB_UNITS_DBPERKM) ? kDb : 1.0;
#pragma unroll
  for (int j = 0; j < FPT; ++j) {
    if (!valid[j]) continue;
    double a_fam[RB_MAX_CONSTITUENTS];
#pragma unroll
    for (int c = 0; c < RB_MAX_CONSTITUENTS; ++c) a_fam[c] = 0.0;
    const double xx = x[j], ff = f[j];
    if (k.slot_nh3 >= 0) {
      double a_low = 0.0, a_sjs = 0.0;
      if (use_low) {
        a_low = xx * s_low[j] * unit;
        // nh3_hs.py:311-312: `< 0 -> 1e-8` in output units; nh3_kd.py:344-349 uses `<= 0`
        if (a_low < 0.0 || (k.nh3_family == 2 && a_low <= 0.0)) a_low = 1.0E-8;
      }
      if (use_sjs) a_sjs = xx * s_sjs[j] * unit;
      double a;
      if (use_low && use_sjs) {
        if (nform == RB_F_NH3_SJSD) {
          const double W = (P < 35.0) ? (P - 10.0) / (35.0 - 10.0) : 1.0 - (P - 35.0) / (100.0 - 35.0);
          a = W * a_low + (1.0 - W) * a_sjs;
        } else {
          const double W = (P - 400.0) / (2000.0 - 400.0);
          a = W * a_sjs + (1.0 - W) * a_low;
        }
      } else {
        a = use_low ? a_low : a_sjs;
      }
      a_fam[k.slot_nh3] = a;
    }
    if (k.slot_h2s >= 0) a_fam[k.slot_h2s] = xx * s_h2s[j] * unit;
    if (k.slot_ph3 >= 0) a_fam[k.slot_ph3] = xx * s_ph3[j] * unit;
    if (k.slot_co >= 0) {
      const double pref = kCoefGeisa * (P_co / 296.0) * s_pow[PW_T296_35];
      a_fam[k.slot_co] = pref * s_co[j] * unit;
    }
    if (k.slot_h2o >= 0) {  // h2o_bk.py:83-125 (native unit dB/km)
      const double M_amu = 8.314472 / 0.46151805;
      double density = (M_amu * P_h2o) / (8.314472e-5 * T);
      density = 0.997317 * (density / M_amu) * 6.0221415e23 * (1.0 / 1e6);
      const double line = 4.342945 * 1e-4 * density * (xx * s_h2o[j]);
      const double Cf_he = 1.0e6 * 1.03562010226e-10, Cf_h2 = 1.0e6 * 5.07722009423e-11;
      const double th3 = s_pow[PW_TH_3];
      const double foreign = Cf_he * P_he * P_h2o * xx * th3 + Cf_h2 * P_h2 * P_h2o * xx * th3;
      const double p3 = P_h2o / 0.001;
      const double self = 3.1e-07 * s_pow[PW_TH_12] * (p3 * p3) * xx;
      double a = line + 4.342945 * foreign + 4.342945 * self;
      if (k.units != RB_UNITS_DBPERKM) a = a / kDb;
      a_fam[k.slot_h2o] = a;
    }
    if (k.slot_h2 >= 0 && k.h2_form == RB_F_H2_ORTON) {
      a_fam[k.slot_h2] = h2_orton_alpha(k.cat[RB_CAT_H2_ORTON], k.F, fidx[j], T, xx, P_h2, P_he, P_ch4,
                                        s_pow[PW_H2_312], s_pow[PW_H2_224], s_pow[PW_H2_334]) * unit;
    } else if (k.slot_h2 >= 0) {  // h2_jj_ddb.py:7-38 / h2_jj.py:7-22
      double pre = 1.0;
      if (k.h2_form == RB_F_H2_JJ_DDB) {
        if (k.h2state == 0) {
          pre = s_pow[PW_H2_E27];
          if (pre > 1.0) pre = 1.0;
          pre *= s_pow[PW_H2_E055];
          if (pre > 1.0) pre = 1.0;
        } else {
          pre = s_pow[PW_H2_N25];
          if (pre > 1.0) pre = 1.0;
        }
      }
      const double cf = 3.9522E-14 * xx * P_h2 * pre;
      a_fam[k.slot_h2] =
          cf * (P_h2 * s_pow[PW_H2_312] + 1.382 * P_he * s_pow[PW_H2_224] + 9.322 * P_ch4 * s_pow[PW_H2_334]) * unit;
    }
    if (k.slot_cld >= 0) {  // clouds_idp.py:6-58
      const double kk = 2.0 * kPi * ff / kGHz;
      double a = 0.0;
      if (k.cloud_flags & 1u) a += acloud(kk, ld0(k.cloud[RB_CLD_H2O], l) / 0.9, water_eps(ff, T));
      if (k.cloud_flags & 2u) a += acloud(kk, ld0(k.cloud[RB_CLD_SOLN], l) / 1.0, water_eps(ff, T));
      if (k.cloud_flags & 4u) a += acloud(kk, ld0(k.cloud[RB_CLD_NH4SH], l) / 1.2, cmul({1.7, -0.005}, {1.7, -0.005}));
      if (k.cloud_flags & 8u) a += acloud(kk, ld0(k.cloud[RB_CLD_NH3], l) / 1.6, cmul({1.3, -0.0001}, {1.3, -0.0001}));
      if (k.cloud_flags & 16u) a += acloud(kk, ld0(k.cloud[RB_CLD_H2S], l) / 1.5, cmul({1.15, -0.0001}, {1.15, -0.0001}));
      if (k.cloud_flags & 32u) a += acloud(kk, ld0(k.cloud[RB_CLD_CH4], l) / 1.0, cmul({1.3, -0.00001}, {1.3, -0.00001}));
      if (a < 0.0) a = 0.0;
      a_fam[k.slot_cld] = a * unit;
    }
    // scale-sum over constituents in call order (alpha.py:151-192)
    double total = 0.0;
    const size_t o = (size_t)l * k.F + fidx[j];
#pragma unroll
    for (int c = 0; c < RB_MAX_CONSTITUENTS; ++c) {
      if (c < k.C) {
        double v = a_fam[c];
        if (k.scale) v *= k.scale[(size_t)c * k.L + l];
        if (k.out_cube) k.out_cube[o * k.C + c] = v;
        total += v;
      }
    }
    if (k.n_peer > 0) {
      const size_t op = (size_t)(k.peer_row0 + l) * k.F + fidx[j];
#pragma unroll
      for (int p = 0; p < RB_MAX_PEERS; ++p)
        if (p < k.n_peer) k.peer_out[p][op] = total;
    } else {
      k.out_total[o] = total;
    }
  }
  }  // trips
}

// Launch order of the layers: the NH3 formalisms evaluate 1014 lines per frequency in the 400..2000 bar blend, 814
// below it and 200 above, so the CTAs of a launch differ in length by a factor of five; longest first (a counting sort
// of the layers by pressure class, layer order kept inside a class) leaves the short ones for the tail of the launch.
__global__ void __launch_bounds__(1024) alpha_order_kernel(const double* __restrict__ P, int L, int* __restrict__ order) {
  auto cls = [](double p) { return (p >= 400.0 && p <= 2000.0) ? 0 : (p < 400.0 ? 1 : 2); };
  // stable placement: thread t handles the contiguous layer range [t * per, (t + 1) * per); an exclusive scan of the
  // per-thread class counts (warp shuffles, then the 32 warp totals) gives every thread its write positions
  const int per = (L + blockDim.x - 1) / blockDim.x;
  const int l0 = min(L, (int)threadIdx.x * per), l1 = min(L, l0 + per);
  int mine[3] = {0, 0, 0};
  for (int l = l0; l < l1; ++l) ++mine[cls(P[l])];
  __shared__ int s_warp[3][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  int incl[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v = mine[c];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += u;
    }
    incl[c] = v;
    if (lane == 31) s_warp[c][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int v = lane < nwarps ? s_warp[c][lane] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
      }
      s_warp[c][lane] = v;                                   // inclusive over the warps
    }
  }
  __syncthreads();
  const int tot0 = s_warp[0][nwarps - 1], tot1 = s_warp[1][nwarps - 1];
  const int base[3] = {0, tot0, tot0 + tot1};
  int pos[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pos[c] = base[c] + (warp ? s_warp[c][warp - 1] : 0) + incl[c] - mine[c];
  for (int l = l0; l < l1; ++l) order[pos[cls(P[l])]++] = l;
}

int family_of(int form) {
  switch (form) {
    case RB_F_NH3_HS: case RB_F_NH3_DBS: case RB_F_NH3_SJS: case RB_F_NH3_HS_SJS: case RB_F_NH3_DBS_SJS:
    case RB_F_NH3_KD: case RB_F_NH3_SJSD: case RB_F_NH3_BG: return 0;
    case RB_F_H2S_DDB: return 1;
    case RB_F_PH3_JH: return 2;
    case RB_F_H2O_BK: return 3;
    case RB_F_H2_JJ_DDB: case RB_F_H2_JJ: case RB_F_H2_ORTON: return 4;
    case RB_F_CLOUDS_IDP: return 5;
    case RB_F_CO_DDB: return 6;
    default: return -1;
  }
}

}  // namespace

// d holds DEVICE pointers; freqs are also needed on the host for the class scan, so the caller
// passes a host copy through ctx scratch (see capi.cu).
int rb_launch_alpha(rb_context* ctx, const rb_alpha_desc* d, const double* h_freqs, double* out_total,
                    double* out_cube, int n_peer, double* const* peer_out, long long peer_row0) {
  AlphaK k{};
  k.L = d->n_layers; k.F = d->n_freqs; k.C = d->n_constituents;
  if (k.L <= 0 || k.F <= 0) return rb_fail(ctx, RB_ERR_INVALID, "alpha: n_layers and n_freqs must be positive");
  if (k.C <= 0 || k.C > RB_MAX_CONSTITUENTS)
    return rb_fail(ctx, RB_ERR_INVALID, "alpha: n_constituents must be 1..%d", RB_MAX_CONSTITUENTS);
  k.freqs = d->freqs; k.T = d->T; k.P = d->P;
  k.freq_stride = d->freqs_per_layer ? d->n_freqs : 0;
  for (int i = 0; i < RB_NUM_GAS; ++i) {
    const int c = d->gas_col[i];
    if (c >= d->gas_rows) return rb_fail(ctx, RB_ERR_INVALID, "alpha: gas_col[%d]=%d out of range", i, c);
    k.gas[i] = (c >= 0 && d->gas) ? d->gas + (size_t)c * k.L : nullptr;
  }
  for (int i = 0; i < RB_NUM_CLD; ++i) {
    const int c = d->cloud_col[i];
    if (c >= d->cloud_rows && d->cloud) return rb_fail(ctx, RB_ERR_INVALID, "alpha: cloud_col[%d]=%d out of range", i, c);
    k.cloud[i] = (c >= 0 && d->cloud) ? d->cloud + (size_t)c * k.L : nullptr;
  }
  k.cloud_flags = d->cloud_flags; k.h2state = d->h2state; k.coshape = d->coshape; k.units = d->units;
  k.scale = d->scale; k.out_total = out_total; k.out_cube = out_cube;
  k.n_peer = 0; k.peer_row0 = peer_row0;
  if (n_peer > 0) {
    if (n_peer > RB_MAX_PEERS || !peer_out) return rb_fail(ctx, RB_ERR_INVALID, "alpha: 1..%d peer slabs", RB_MAX_PEERS);
    k.n_peer = n_peer;
    for (int p = 0; p < n_peer; ++p) k.peer_out[p] = peer_out[p];
  }
  k.nh3_form = 0; k.slot_nh3 = k.slot_h2s = k.slot_ph3 = k.slot_h2o = k.slot_h2 = k.slot_cld = k.slot_co = -1;
  for (int c = 0; c < k.C; ++c) {
    const int fm = d->formalism[c];
    k.form[c] = fm;
    int* slot = nullptr;
    switch (family_of(fm)) {
      case 0: slot = &k.slot_nh3; k.nh3_form = fm; break;
      case 1: slot = &k.slot_h2s; break;
      case 2: slot = &k.slot_ph3; break;
      case 3: slot = &k.slot_h2o; break;
      case 4: slot = &k.slot_h2; k.h2_form = fm; break;
      case 5: slot = &k.slot_cld; break;
      case 6: slot = &k.slot_co; break;
      default: return rb_fail(ctx, RB_ERR_UNSUPPORTED, "alpha: formalism id %d is not built", fm);
    }
    if (*slot >= 0) return rb_fail(ctx, RB_ERR_INVALID, "alpha: two formalisms of the same gas family");
    *slot = c;
  }
  k.nh3_family = (k.nh3_form == RB_F_NH3_DBS || k.nh3_form == RB_F_NH3_DBS_SJS) ? 1
                 : (k.nh3_form == RB_F_NH3_KD || k.nh3_form == RB_F_NH3_SJSD) ? 2 : 0;
  for (int i = 0; i < RB_NUM_CATALOGS; ++i) { k.cat[i] = ctx->cat[i]; k.ncat[i] = ctx->cat_n[i]; }
  const bool need_low = k.slot_nh3 >= 0 && k.nh3_form != RB_F_NH3_SJS && k.nh3_form != RB_F_NH3_BG;
  const bool need_sjs = k.slot_nh3 >= 0 && k.nh3_form != RB_F_NH3_HS && k.nh3_form != RB_F_NH3_DBS &&
                        k.nh3_form != RB_F_NH3_KD;
  auto need_cat = [&](int id, const char* nm) -> int {
    if (!ctx->cat[id] || ctx->cat_n[id] <= 0) return rb_fail(ctx, RB_ERR_INVALID, "alpha: line catalog '%s' not set", nm);
    return RB_OK;
  };
  if (need_low) { RB_TRY(need_cat(RB_CAT_NH3_INV, "nh3_inv")); RB_TRY(need_cat(RB_CAT_NH3_ROT, "nh3_rot")); RB_TRY(need_cat(RB_CAT_NH3_V2, "nh3_v2")); }
  if (need_sjs) RB_TRY(need_cat(RB_CAT_NH3_SJS, "nh3_sjs"));
  if (k.slot_h2s >= 0) RB_TRY(need_cat(RB_CAT_H2S, "h2s"));
  if (k.slot_ph3 >= 0) RB_TRY(need_cat(RB_CAT_PH3, "ph3"));
  if (k.slot_h2o >= 0) RB_TRY(need_cat(RB_CAT_H2O, "h2o"));
  if (k.slot_co >= 0) RB_TRY(need_cat(RB_CAT_CO, "co"));
  if (k.slot_h2 >= 0 && k.h2_form == RB_F_H2_ORTON) {
    if (d->freqs_per_layer) return rb_fail(ctx, RB_ERR_UNSUPPORTED, "alpha: h2_orton with per-layer frequencies");
    RB_TRY(need_cat(RB_CAT_H2_ORTON, "h2_orton"));
    if (ctx->cat_n[RB_CAT_H2_ORTON] != k.F)
      return rb_fail(ctx, RB_ERR_INVALID, "alpha: the h2_orton table was prepared for %d frequencies, the call has %d",
                     ctx->cat_n[RB_CAT_H2_ORTON], k.F);
  }

  // frequency classes
  const long long n_hf = (long long)k.F * (d->freqs_per_layer ? k.L : 1);   // per-layer lists: the union of all rows
  for (long long i = 0; i < n_hf; ++i) {
    const double f = h_freqs[i];
    if (f <= 30.0) k.any_lo = 1; else k.any_hi = 1;
    if (f <= 26.0) k.any_S = 1; else if (f >= 34.0) k.any_J = 1; else k.any_I = 1;
  }
  if (k.nh3_form == RB_F_NH3_BG) { k.any_S = 0; k.any_I = 0; k.any_J = 1; }   // one constant set at every frequency
  // tiling: lanes = 32*FPT frequencies per warp; K line slices per group when F is small
  k.fpt = (k.F >= 512) ? 2 : 1;
  {
    // RB_ALPHA_FPT=4: four frequencies per lane (128 per warp) for wide sweeps -- half the table reads per line
    // evaluation of FPT = 2, at one CTA per SM (measured at C5: see DESIGN.md 3.1)
    const char* e = getenv("RB_ALPHA_FPT");
    if (e && atoi(e) == 4 && k.F >= 1024) k.fpt = 4;
    if (e && atoi(e) == 1) k.fpt = 1;
  }
  k.ngroups = (k.F + 32 * k.fpt - 1) / (32 * k.fpt);
  k.K = 1;
  while (k.K < kWarps && k.ngroups * k.K * 2 <= kWarps) k.K *= 2;
  const int gpb = kWarps / k.K;
  // CTAs per layer: one group per warp when that is what fills the GPU, else as few as leave ~4 CTAs per slot
  // (2 resident CTAs per SM), each looping over the remaining groups with the layer's tables in place
  int nblk_x = (k.ngroups + gpb - 1) / gpb;
  {
    const long long want = 8LL * ctx->num_sms;
    const int per_layer = (int)((want + k.L - 1) / k.L);
    const char* e = getenv("RB_ALPHA_BLOCKS_PER_LAYER");     // measurement aid: 0 = one group per warp (no trips)
    if (e && atoi(e) > 0) { if (atoi(e) < nblk_x) nblk_x = atoi(e); }
    else if (!e && per_layer < nblk_x) nblk_x = per_layer < 1 ? 1 : per_layer;
  }

  // smem layout
  int off = 0;
  auto take = [&](int ndoubles) { int o = off; off += (ndoubles + 1) & ~1; return o; };
  if (need_low) {
    k.off_inv_lo = take(4 * k.ncat[RB_CAT_NH3_INV]);
    k.off_inv_hi = take(4 * k.ncat[RB_CAT_NH3_INV]);
    k.off_rot_ag = take(2 * k.ncat[RB_CAT_NH3_ROT]); k.off_rot_n = take(k.ncat[RB_CAT_NH3_ROT]);
    k.off_v2_ag = take(2 * k.ncat[RB_CAT_NH3_V2]); k.off_v2_n = take(k.ncat[RB_CAT_NH3_V2]);
  }
  if (need_sjs) {
    k.off_sjs_S = take(4 * k.ncat[RB_CAT_NH3_SJS]);
    k.off_sjs_J = take(4 * k.ncat[RB_CAT_NH3_SJS]);
    k.off_sjs_I = take(k.any_I ? 6 * k.ncat[RB_CAT_NH3_SJS] : 0);
  }
  if (k.slot_h2s >= 0) { k.off_h2s_ag = take(2 * k.ncat[RB_CAT_H2S]); k.off_h2s_n = take(k.ncat[RB_CAT_H2S]); }
  if (k.slot_ph3 >= 0) k.off_ph3 = take(4 * k.ncat[RB_CAT_PH3]);
  if (k.slot_h2o >= 0) k.off_h2o = take(4 * k.ncat[RB_CAT_H2O]);
  if (k.slot_co >= 0) k.off_co = take(2 * k.ncat[RB_CAT_CO]);
  if (k.K > 1) k.off_red = take(kWarps * 6 * k.fpt * 32);
  const size_t smem_bytes = (size_t)off * sizeof(double);
  if (smem_bytes > ctx->smem_optin)
    return rb_fail(ctx, RB_ERR_INVALID, "alpha: line tables need %zu B of shared memory (limit %zu B)", smem_bytes,
                   ctx->smem_optin);

  if (ctx->alpha_newton < 0) {
    // measured on B200 (tools/probe_rcp.py): seed 9.8e-7, 1 step 9.6e-13, 2 steps 1.1e-16 max relative
    // error.  One step is 6 orders below the 1e-6 parity bar and 3 below the 1e-9 the tests hold.
    const char* e = getenv("RB_RCP_NEWTON");
    ctx->alpha_newton = (e && e[0] == '2') ? 2 : 1;
  }
  const int newton = ctx->alpha_newton;
  void (*kern)(const AlphaK) = nullptr;
  if (k.fpt == 4) kern = (newton == 1) ? alpha_lines_kernel<4, 1> : alpha_lines_kernel<4, 2>;
  else if (k.fpt == 2) kern = (newton == 1) ? alpha_lines_kernel<2, 1> : alpha_lines_kernel<2, 2>;
  else kern = (newton == 1) ? alpha_lines_kernel<1, 1> : alpha_lines_kernel<1, 2>;
  RB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  dim3 grid(nblk_x, k.L);
  if (k.L > 65535) return rb_fail(ctx, RB_ERR_INVALID, "alpha: n_layers > 65535 not supported");
  k.order = nullptr;
  if (k.slot_nh3 >= 0 && k.L >= 2 * ctx->num_sms && !getenv("RB_ALPHA_NO_ORDER")) {
    void* p_order;
    RB_TRY(rb_ensure(ctx, RB_BUF_ORDER, (size_t)k.L * sizeof(int), &p_order));
    alpha_order_kernel<<<1, 1024, 0, ctx->stream>>>(k.P, k.L, (int*)p_order);
    ctx->launches += 1;
    k.order = (const int*)p_order;
  }
  RB_CUDA(ctx, rb_time_begin(ctx, 0));
  kern<<<grid, kThreads, smem_bytes, ctx->stream>>>(k);
  RB_CUDA(ctx, cudaGetLastError());
  RB_CUDA(ctx, rb_time_end(ctx, 0));
  ctx->launches += 1;
  return RB_OK;
}


// ---- scale-sum of a cached per-constituent cube (alpha.py:151-192) ----------------------------------
namespace {
__global__ void alpha_scale_sum_kernel(const double* __restrict__ cube, const double* __restrict__ scale, int L, int F,
                                       int C, double* __restrict__ total, double* __restrict__ out_cube) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (layer, freq)
  if (idx >= (size_t)L * F) return;
  const int l = (int)(idx / F);
  double t = 0.0;
  for (int c = 0; c < C; ++c) {                                       // left-to-right like the reference
    double v = cube[idx * C + c];
    if (scale) v *= scale[(size_t)c * L + l];
    if (out_cube) out_cube[idx * C + c] = v;
    t += v;
  }
  total[idx] = t;
}
}  // namespace

int rb_launch_alpha_scale_sum(rb_context* ctx, const double* cube, const double* scale, int L, int F, int C,
                              double* total, double* out_cube) {
  const size_t n = (size_t)L * F;
  alpha_scale_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(cube, scale, L, F, C, total, out_cube);
  RB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return RB_OK;
}
