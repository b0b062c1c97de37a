// radiobear_b200 -- hot path B: ray geometry (raypath.compute_ds) and the fused
// optical-depth / weighting-function / brightness-temperature integration (Brightness.single).
//
// Data layout in HBM
//   ds slab   [R/32][S = L-1][32]  float64, tiled by groups of 32 rays: the geometry kernel (one thread per
//             ray) writes 256-byte rows coalesced, and the integration kernel (lanes = the same 32 rays)
//             walks one contiguous 256 KB tile per block -- sequential in DRAM pages and inside one or two
//             2 MB TLB pages (a plain [S][R] slab strides 2.9 MB per segment and thrashes the TLB).
//   alpha slab [L][F] float64, FREQUENCY fastest (written by alpha_lines_kernel): a warp reads one
//             256-byte row per segment, shared by all rays it carries; 0.5 MB at C4, lives in L2/L1.
//   Tb        [R][F] float32 or float64, coalesced stores.
//
// Geometry is evaluated in an algebraic (trig-free) form of the reference recurrence:
//   shell radius at parametric latitude lat: r*sqrt(q^2 sin^2 lat + cos^2 lat)    (shape.py:240-244)
//   sin lat = y/|r| of the current position                                        (raypath.py:231)
//   ds = -r.s - sqrt((r.s)^2 + rNext^2 - rNow^2)  (NaN below the tangent shell)    (raypath.py:193)
//   Snell refraction at the first interface only, nratio = 1 afterwards            (raypath.py:158,257)
// tests/test_ray_geometry.py checks it against the trigonometric oracle and the reference's vectors.
#include "rb_common.cuh"
#include <cmath>
#include <climits>

namespace {

constexpr double kTcmb = 2.725;   // utils.py:77
constexpr double kKmToCm = 1.0E5; // brightness.py:66

// offset of (ray r, segment 0) in the tiled ds slab; consecutive segments of a ray are 32 doubles apart
__host__ __device__ __forceinline__ size_t ds_tile_base(long long r, int S) {
  return ((size_t)(r >> 5) * (size_t)S) * 32 + (size_t)(r & 31);
}
constexpr int kDsStride = 32;

struct GeoK {
  int L;
  const double* radius;
  double n0, n1, q;
  double cz, sz, cx, sx;  // rotate2planet = rotX(rotate) . rotZ(tip)   (raypath.py:47-51)
  int limb;
  long long R, Rpad;
  const double* b;
  double* ds;
  float* dsf;    // optional (mixed-precision integration): [R/32][S][32] (float)ds_i, 0 below the last used one
  int* nseg;
  int* nanflag;  // 1 when a segment the integration uses (index <= nseg-2) is NaN -> Tb is NaN
  // compacted mode (see ray_edge_kernel): ds tiles, nseg and nanflag are indexed by the position in the list
  // of hitting rays instead of the ray index
  int* cidx;     // [<= R] list position -> ray index
  int* ncomp;    // length of the list (device scalar)
  double* zq;    // [R] edge depth found by findEdge, NaN for a ray that misses
  int* blkcnt;   // [ceil(R / kEdgeThreads)] hits per block of rays
  const double* dr2;  // [L-1] R_l+1^2 - R_l^2 (layer_dr2_kernel)
  // Streamed geometry: the integration of a tile may start while its rays are still being traced.  Every ray counts
  // itself into prog[tile][c] once its rows 0 .. 32 (c + 1) are in global memory (fence, then a relaxed atomic); the
  // integration CTA waits (one thread, acquire load) until all rays of its tile have counted before it copies chunk c.
  // A ray that ends counts into all remaining chunks.  In this mode the rows from the last used segment on are written
  // as 0 (the integration walks all S - 1 steps: it does not know the segment count before the trace has ended).
  int* prog;          // [tiles][npub] or null
  int npub;           // chunks per tile
};

__device__ __forceinline__ void rot2planet(const GeoK& g, double x, double y, double z, double& ox, double& oy,
                                           double& oz) {
  const double x1 = g.cz * x - g.sz * y, y1 = g.sz * x + g.cz * y;
  ox = x1;
  oy = g.cx * y1 - g.sx * z;
  oz = g.sx * y1 + g.cx * z;
}

// sin / cos of the (parametric) latitude used by Shape._calcEllipse for a position with y/|r| = v;
// lat == 0 is replaced by 1e-6 rad (shape.py:231-233)
__device__ __forceinline__ void lat_sc(double v, double& s, double& c) {
  if (v == 0.0) {
    s = 9.99999999999833333e-07;  // sin(1e-6)
    c = 0.9999999999995;          // cos(1e-6)
  } else {
    s = v;
    c = sqrt(fmax(0.0, 1.0 - v * v));
    if (v != v) c = v;  // NaN propagates
  }
}

// ---- findEdge (raypath.py:60-105): march zQ down by 0.005 until inside the outer shell ----------------
// rays with b^2 >= 1 (raypath.py:126-127; NaN impact parameters too) never hit
// (not inlined: the classification kernel and the plain geometry kernel must run the very same instructions, so that
//  compacted and plain launches give the same bits)
__device__ __noinline__ bool find_edge(const GeoK& g, double bx, double by, double& zq) {
  const double bb = bx * bx + by * by;
  const double q2 = g.q * g.q;
  const double rNorm = g.radius[0];
  zq = 0.0;
  if (!(bb < 1.0)) return false;
  const double z0 = sqrt(1.0 - bb) * 1.01;
  const int ntrial = (int)ceil(z0 / 0.005);  // len(np.arange(z0, 0, -0.005))
  double d_prev = 0.0, z_prev = 0.0;
  for (int t = 0; t < ntrial; ++t) {
    const double z = z0 + t * (-0.005);
    double px, py, pz;
    rot2planet(g, bx, by, z, px, py, pz);
    const double nb = sqrt(px * px + py * py + pz * pz);
    const double r1 = nb * rNorm;
    double s, c;
    lat_sc(py / nb, s, c);
    const double r2 = rNorm * sqrt(q2 * s * s + c * c);
    const double d = r1 - r2;
    if (r1 < r2) {
      // np.interp(0, [d, d_prev], [z, z_prev]); a first-trial hit returns z itself
      zq = (t == 0) ? z : ((z_prev - z) / (d_prev - d)) * (0.0 - d) + z;
      return true;
    }
    d_prev = d;
    z_prev = z;
  }
  return false;
}

// sqrt / scaled reciprocal of the layer march: SFU seed (MUFU.RSQ64H / MUFU.RCP64H read the high word of the
// operand only: relative error eps <= 2^-20) and ONE third-order correction, so that each sits on the dependent
// chain of a segment with four / three FP64 latencies instead of six (two coupled Newton steps):
//   sqrt(x) = g (1 - 2e)^-1/2 = g (1 + e + 3/2 e^2) + O(5/2 e^3),  g = x r0,  e = 1/2 - (r0/2) g   (|e| ~ eps)
//   c / x   = c r0 (1 + e + e^2) + O(e^3),                         e = 1 - x r0
// 2.5 eps^3 < 3e-18: below half an ulp.  x = 0 gives NaN (0 * inf) and x < 0 gives NaN like np.sqrt.
__device__ __forceinline__ double sqrt_seeded(double x) {
  double r0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  const double gq = x * r0, hq = 0.5 * r0;
  const double eq = fma(-hq, gq, 0.5);
  return fma(gq * eq, fma(1.5, eq, 1.0), gq);
}
__device__ __forceinline__ double rcp_seeded_scaled(double x, double c) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  const double eq = fma(-x, r0, 1.0);
  const double rc = r0 * c;
  return fma(rc, fma(eq, eq, eq), rc);
}

// R_l+1^2 - R_l^2 for every shell pair, once per launch instead of three FP64 instructions per (ray, segment)
__global__ void layer_dr2_kernel(const double* __restrict__ radius, int L, double* __restrict__ dr2) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < L - 1) dr2[l] = (radius[l + 1] - radius[l]) * (radius[l + 1] + radius[l]);
}

// ---- compaction of the rays that hit the planet (rays-major FP64 integration) --------------------------
// An image grid is 2/3 sky, and a tile of 32 consecutive pixels that straddles the limb carries idle lanes
// through every instruction of the layer march and of the integration (measured: 28.9 of 32 lanes active).
// ray_edge_kernel classifies every ray (findEdge) and counts the hits per block; ray_compact_kernel turns the
// counts into offsets and writes the list of hitting rays in their original order; the layer march and the
// integration then run over full tiles of that list and scatter their results to the original ray index.
constexpr int kEdgeThreads = 256;
__global__ void __launch_bounds__(kEdgeThreads) ray_edge_kernel(const __grid_constant__ GeoK g) {
  const long long r = (long long)blockIdx.x * kEdgeThreads + threadIdx.x;
  bool hit = false;
  double zq = 0.0;
  if (r < g.R) {
    hit = find_edge(g, g.b[2 * r], g.b[2 * r + 1], zq);
    g.zq[r] = hit ? zq : nan("");
  }
  const int n = __syncthreads_count(hit);
  if (threadIdx.x == 0) g.blkcnt[blockIdx.x] = n;
}

// One CTA per super-block of kSortBlock rays: counting sort of its hits by projected radius (256 bins), written at the
// super-block's offset in the list (the hits of all earlier blocks).  The order of the rays inside one bin is whatever
// the shared-memory atomics make it; nothing depends on it (every ray is computed on its own, only its lane changes).
constexpr int kSortThreads = 1024;
constexpr int kSortBins = 256;
__global__ void __launch_bounds__(kSortThreads) ray_compact_kernel(const __grid_constant__ GeoK g) {
  __shared__ int s_part[kSortThreads / 32];
  __shared__ int s_hist[kSortBins], s_cur[kSortBins];
  __shared__ int s_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // hits in the edge blocks (kEdgeThreads rays each) before this super-block
  const int first_blk = (int)blockIdx.x * (kSortBlock / kEdgeThreads);
  int acc = 0;
  for (int j = tid; j < first_blk; j += kSortThreads) acc += g.blkcnt[j];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) s_part[warp] = acc;
  if (tid < kSortBins) { s_hist[tid] = 0; s_cur[tid] = 0; }
  __syncthreads();
  if (tid == 0) {
    int b = 0;
    for (int w = 0; w < kSortThreads / 32; ++w) b += s_part[w];
    s_base = b;
  }
  const double iq2 = 1.0 / (g.q * g.q);
  constexpr int kPer = kSortBlock / kSortThreads;
  int bin[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const long long r = (long long)blockIdx.x * kSortBlock + i * kSortThreads + tid;
    bin[i] = -1;
    if (r < g.R) {
      const double z = g.zq[r];
      if (z == z) {
        const double bx = g.b[2 * r], by = g.b[2 * r + 1];
        const double rho2 = fmin(bx * bx + by * by * iq2, 1.0);    // 0 at the disc centre, ~1 at the limb
        // bins of equal width in mu = sqrt(1 - rho^2): the path geometry changes fastest near the limb
        bin[i] = min(kSortBins - 1, (int)((1.0 - sqrt(1.0 - rho2)) * kSortBins));
        atomicAdd(&s_hist[bin[i]], 1);
      }
    }
  }
  __syncthreads();
  if (warp == 0) {                                           // exclusive scan of the 256 bin counts by one warp
    int carry = 0;
    for (int c = 0; c < kSortBins; c += 32) {
      const int v = s_hist[c + lane];
      int x = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      s_hist[c + lane] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0 && blockIdx.x == gridDim.x - 1) g.ncomp[0] = s_base + carry;
    // where the copy-out pipeline enters the list: the tile that holds the first hit of the super-block its plain
    // ray order starts with (rb_progress_shift)
    if (lane == 0 && (long long)blockIdx.x * kSortBlock == 32LL * rb_progress_shift(g.R)) g.ncomp[1] = s_base >> 5;
  }
  __syncthreads();
  const int base = s_base;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    if (bin[i] >= 0) {
      const long long r = (long long)blockIdx.x * kSortBlock + i * kSortThreads + tid;
      g.cidx[base + s_hist[bin[i]] + atomicAdd(&s_cur[bin[i]], 1)] = (int)r;
    }
  }
}

// one thread per ray (COMPACT: per entry of the list of hitting rays)
#ifndef RB_GEO_CTAS
#define RB_GEO_CTAS 8
#endif
// 64 registers: 8 CTAs of 128 rays per SM = 151 k rays in flight, so that the ~117 k rays of the C4 image are ONE
// wave (every ray is the same ~1000-step chain: a second, nearly empty wave doubles the kernel time -- measured with
// 80 registers, 6 CTAs per SM: 918 CTAs on 888 slots, 0.50 ms)
// (one kernel for the plain and the compacted launch -- g.cidx selects -- so that both run the very same instructions
//  per ray: two template instantiations were contracted into FMAs differently by the compiler in the start-up and
//  restart arithmetic, and 743 rays of the C4 image differed in the last bits of ds)
// RB_GEO_FUSED = 2 / 1: the layer march with one / two latencies taken out of the dependent chain of every segment (see the
// speculative block).  Built and measured (profiles/r2_ab_geometry_chain.txt): 0.52 / 0.54 ms against 0.48 ms on C4 and
// slower on a rank's share of eight GPUs too -- the extra FP64 instructions and spills cost more than the latencies save.  Off.
#ifndef RB_GEO_FUSED
#define RB_GEO_FUSED 0
#endif
__global__ void __launch_bounds__(128, RB_GEO_CTAS) ray_geometry_kernel(const __grid_constant__ GeoK g) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // streamed trace: count the CTAs that have started.  The integration that waits on g.prog is released (stream
  // memory operation on this counter) only when every CTA of the trace is resident or done, so that its waiting CTAs
  // can never keep a CTA of the trace off the machine.
  if (g.prog && threadIdx.x == 0) atomicAdd(g.prog + (size_t)(g.Rpad / 32) * g.npub, 1);
  const int S = g.L - 1;
  const bool COMPACT = g.cidx != nullptr;
  long long r = t;
  bool inrange, hit;
  double bx = 2.0, by = 2.0, zq = 0.0;
  if (COMPACT) {
    inrange = hit = t < (long long)*g.ncomp;
    if (!inrange) return;
    r = g.cidx[t];
    bx = g.b[2 * r]; by = g.b[2 * r + 1];
    zq = g.zq[r];
  } else {
    inrange = r < g.R;
    if (inrange) { bx = g.b[2 * r]; by = g.b[2 * r + 1]; }
    hit = find_edge(g, bx, by, zq);
    // Lanes leave the trial loop at different iterations; without this the early ones run ahead into the
    // layer loop and the warp stays split for all ~1000 segments (measured: 18.7 active lanes per issue).
    __syncwarp();
    if (inrange) g.nanflag[r] = 0;
    if (!hit) {
      if (inrange) g.nseg[r] = -1;
      return;
    }
  }
  const double q2 = g.q * g.q;
  const double rNorm = g.radius[0];
  const double mu = sqrt(fmax(0.0, 1.0 - bx * bx - by * by));
  // edge position, then the shell point / normal the reference starts from (raypath.py:141-156)
  double ex, ey, ez;
  rot2planet(g, bx, by, zq, ex, ey, ez);
  ex *= rNorm; ey *= rNorm; ez *= rNorm;
  double px, py, pz;  // current position r_i
  double v, hs, sl, clng;
  {
    const double ne = sqrt(ex * ex + ey * ey + ez * ez);
    v = ey / ne;
    hs = sqrt(ex * ex + ez * ez);
    sl = (hs > 0.0) ? ex / hs : 0.0;
    clng = (hs > 0.0) ? ez / hs : 1.0;
    double s, c;
    lat_sc(v, s, c);
    // geoid.r = rotY(lng, [0, b sin lat, a cos lat])
    px = rNorm * c * sl;
    py = g.q * rNorm * s;
    pz = rNorm * c * clng;
  }
  double s_lat, c_lat;
  lat_sc(v, s_lat, c_lat);
  double shape = sqrt(q2 * s_lat * s_lat + c_lat * c_lat);  // rmag / req at the current latitude
  // outward normal n = rotY(lng, [0, a sin, b cos]/norm)
  double nx, ny, nz;
  {
    const double inv = 1.0 / sqrt(s_lat * s_lat + q2 * c_lat * c_lat);
    nx = g.q * c_lat * sl * inv;
    ny = s_lat * inv;
    nz = g.q * c_lat * clng * inv;
  }
  // start direction (0,0,-1) rotated into the planet frame
  double sx, sy, sz;
  rot2planet(g, 0.0, 0.0, -1.0, sx, sy, sz);
  // first interface: Snell with nratio = n0/n1 (raypath.py:157-159, 176-177)
  {
    const double nratio = g.n0 / g.n1;
    const double ci = -(sx * nx + sy * ny + sz * nz);      // cos(t_inc)
    const double st = nratio * sqrt(fmax(0.0, 1.0 - ci * ci));
    double ct = sqrt(1.0 - st * st);                        // cos(asin(st)); NaN if st > 1
    if (fabs(ci) > 1.0) ct = nan("");                       // arccos out of range
    const double w = nratio * ci - ct;
    sx = nratio * sx + w * nx;
    sy = nratio * sy + w * ny;
    sz = nratio * sz + w * nz;
  }
  int layer = 0;
  int count = 0;
  const long long o = COMPACT ? t : r;                       // where this ray's results live
  double* out = g.ds + ds_tile_base(o, S);
  // mixed-precision integration (rt_integrate_rays_mixed_kernel) also wants every segment as a float; the
  // one below the last used segment (i = n-1, brightness.py:65 stops at n-2) is stored as 0 so that the
  // trapezoid weight ds_i + ds_i+1 of the last node needs no special case
  float* const outf = g.dsf ? g.dsf + ds_tile_base(o, S) : nullptr;
  const double e2 = 1.0 - q2;
  const double sin6 = 9.99999999999833333e-07, cos6 = 0.9999999999995;   // sin / cos of 1e-6 rad
  const double shape0sq = q2 * sin6 * sin6 + cos6 * cos6;               // lat == 0 -> 1e-6 rad (shape.py:231-233)
  // Between two (rare) direction changes the ray is a straight line r(t) = r0 + t s entering the planet
  // (r.s < 0), and the recurrence of raypath.py:176-237 collapses to ONE scalar, X = (r.s)^2:
  //   ds_i = -r.s - sqrt((r.s)^2 + rNext^2 - rNow^2)          (raypath.py:193)
  //   (r.s)_i+1 = (r.s)_i + ds_i = -sqrt(X_i+1),   X_i+1 = X_i + shape2_i (R_i+1^2 - R_i^2)
  //   ds_i = sqrt(X_i) - sqrt(X_i+1)
  //   shape2 = (shell radius / equatorial radius)^2 = 1 - (1 - q^2) y^2 / |r|^2           (shape.py:240-244)
  //   |r|^2 = X + perp2,   y = y0 + t sy = c0 - sy sqrt(X)      (t = -sqrt(X) - r0.s)
  // The kernel is bound by the latency of this chain (one thread per ray, ~1000 sequential segments):
  //   X' = fma -> sqrt (seed + 4) -> y' (1) -> y'^2 (1) -> shape2' (1)      = SFU seed + 8 FP64 latencies,
  // the reciprocal of |r|^2 (seed + 4) running beside the square root.
  double rd0 = px * sx + py * sy + pz * sz;                  // r0.s  (< 0: ingress)
  double perp2 = (px * px + py * py + pz * pz) - rd0 * rd0;  // |r0|^2 - (r0.s)^2: constant along the line
  double c0 = py - rd0 * sy;
  double X = rd0 * rd0;
  double sq = fabs(rd0);                                     // sqrt(X)
  double shape2 = shape * shape;
  bool outward = !(rd0 < 0.0);                               // only possible after a reflection (below)
  int cool = 0;                                              // careful steps to take before speculating again
  int pubc = 0;                                              // streamed geometry: chunks published so far
  while (layer < S) {
    if (g.prog && layer >= (pubc + 1) * kGeoPub + 1) {       // rows 0 .. 32 (pubc + 1) are written
      __threadfence();
      atomicAdd(g.prog + (size_t)(t >> 5) * g.npub + pubc, 1);
      ++pubc;
    }
    // Speculative block of kSpec segments: the plain recurrence with every exceptional condition (ray leaves,
    // NaN below the tangent shell, y == 0, grazing incidence) folded into one flag that nothing waits for until
    // the end of the block -- no branch sits between two segments of the chain.  A flagged block is discarded
    // and its segments are redone one at a time by the careful step below (same arithmetic, same values).
    constexpr int kSpec = 4;
    if (cool == 0 && !outward && g.limb != RB_LIMB_SEC && layer + kSpec <= S) {
      const double X0 = X, sq0 = sq, sh0 = shape2;
      double dsv[kSpec];
      bool bad = false;
#if RB_GEO_FUSED
      // The chain with two latencies taken out of every segment (X' -> sqrt -> y^2 -> X'' instead of
      // X' -> sqrt -> y -> y^2 -> shape2 -> X''):
      //   y^2 = (c0 - sy sqrt(X'))^2 = (c0^2 + sy^2 X') - 2 c0 sy sqrt(X')          one FMA behind the square root
      //   X'' = X' + shape2 dR2' = (X' + dR2') + y^2 (inve dR2')                     one FMA behind y^2
      // (the bracketed terms depend on X' only and run beside the square root).  Rounding differs from the plain form
      // by a few 1e-16 of |r|^2 resp. X; shape2 itself is only needed as the state the block ends with.
#if RB_GEO_FUSED == 1
      const double c02 = c0 * c0, sy2 = sy * sy, m2 = -2.0 * c0 * sy;
#endif
      double Xn = fma(shape2, g.dr2[layer], X);
#pragma unroll
      for (int j = 0; j < kSpec; ++j) {
        const double sqn = sqrt_seeded(Xn);
        const double inve = rcp_seeded_scaled(Xn + perp2, -e2);
        const double ds = sq - sqn;
        const double yn = fma(-sy, sqn, c0);
#if RB_GEO_FUSED == 1
        const double yy = fma(m2, sqn, fma(sy2, Xn, c02));
#else
        const double yy = yn * yn;                             // RB_GEO_FUSED == 2: only the second of the two
#endif
        const double syy = sy * yn;
        const double d = fma(g.q, -sqn - syy, syy);
        bad |= !(ds >= 0.0) | (yn == 0.0) | !(d <= 0.0);
        dsv[j] = ds;
        X = Xn; sq = sqn;
        if (j + 1 < kSpec) {
          const double dn = g.dr2[layer + j + 1];
          Xn = fma(yy, inve * dn, Xn + dn);
        } else {
          shape2 = fma(yy, inve, 1.0);
        }
      }
#else
#pragma unroll
      for (int j = 0; j < kSpec; ++j) {
        const double Xn = fma(shape2, g.dr2[layer + j], X);
        const double sqn = sqrt_seeded(Xn);
        const double inve = rcp_seeded_scaled(Xn + perp2, -e2);
        const double ds = sq - sqn;
        const double yn = fma(-sy, sqn, c0);
        const double syy = sy * yn;
        const double d = fma(g.q, -sqn - syy, syy);
        bad |= !(ds >= 0.0) | (yn == 0.0) | !(d <= 0.0);
        dsv[j] = ds;
        shape2 = fma(yn * yn, inve, 1.0);
        X = Xn; sq = sqn;
      }
#endif
      if (!bad) {
#pragma unroll
        for (int j = 0; j < kSpec; ++j) {
          out[(size_t)(layer + j) * kDsStride] = dsv[j];
          if (outf) outf[(size_t)(layer + j) * kDsStride] = (float)dsv[j];
        }
        layer += kSpec; count += kSpec;
        continue;
      }
      X = X0; sq = sq0; shape2 = sh0;
      cool = kSpec;
    }
    if (cool > 0) --cool;
    // ---- careful step: one segment with every test of the reference in place ----
    const double Xn = fma(shape2, g.dr2[layer], X);
    double sqn = sqrt_seeded(Xn);                            // NaN for Xn < 0 like np.sqrt
    if (Xn == 0.0) sqn = 0.0;
    double ds = sq - sqn;
    if (outward) ds = -sq - sqn;                             // r.s > 0: the reference's dsm is negative -> stop
    if (ds < 0.0) break;                                     // raypath.py:212-216
    if (g.limb == RB_LIMB_SEC) {
      const double sh = sqrt(shape2);
      ds = fabs(g.radius[layer + 1] * sh - g.radius[layer] * sh) / mu;   // raypath.py:218-219 (also replaces a NaN)
    }
    out[(size_t)layer * kDsStride] = ds;
    ++count;
    if (outf) outf[(size_t)layer * kDsStride] = (float)ds;
    if (ds != ds) {
      // below the tangent shell: every later segment is NaN too (np.sqrt of a negative number,
      // raypath.py:192-209 never reaches its `except`), so just fill the rest of the column
      for (++layer; layer < S; ++layer) out[(size_t)layer * kDsStride] = ds;
      count = S;
      break;
    }
    // latitude of the new point -> shape factor of the next shell pair (raypath.py:228-237)
    const double yn = fma(-sy, sqn, c0);
    const double inve = rcp_seeded_scaled(Xn + perp2, -e2);
    double shape2n = fma(yn * yn, inve, 1.0);
    if (yn == 0.0) shape2n = shape0sq;
    // incidence on the next shell with nratio = 1 (raypath.py:176-177, 246, 257): the reference's
    // arccos / arcsin pair gives s += (cos t_inc - |cos t_inc|) n, i.e. nothing unless cos t_inc < 0.
    // n is parallel to (q x, y, q z), so cos t_inc < 0  <=>  d = q (r.s - sy y) + sy y > 0.
    const double syy = sy * yn;
    const double d = fma(g.q, -sqn - syy, syy);
    X = Xn; sq = sqn; shape2 = shape2n;
    ++layer;
    if (g.limb == RB_LIMB_SEC || !(d <= 0.0) || outward) {
      // grazing ray (rare) or the secant mode (position advances by the secant ds, not along the chord):
      // go through the vector form and restart the line at the current point.  (For y == 0 the reference's
      // normal carries a 1e-6 y-component; it can only change the sign test within 1e-6 rad of tangency,
      // where the next segment is NaN anyway, so the plain test is used.)
      const double tt = (g.limb == RB_LIMB_SEC) ? ds : ((outward ? sqn : -sqn) - rd0);
      px = fma(tt, sx, px); py = fma(tt, sy, py); pz = fma(tt, sz, pz);
      if (g.limb != RB_LIMB_SEC && !outward) py = yn;
      const double nr2 = px * px + py * py + pz * pz;
      if (g.limb == RB_LIMB_SEC) shape2 = (py == 0.0) ? shape0sq : fma(-e2 * (py * py), 1.0 / nr2, 1.0);
      double ux, uy, uz;
      if (py == 0.0) {
        const double hxz = sqrt(px * px + pz * pz);
        ux = g.q * cos6 * px / hxz; uy = sin6; uz = g.q * cos6 * pz / hxz;
      } else {
        ux = g.q * px; uy = py; uz = g.q * pz;
      }
      const double invn = 1.0 / sqrt(ux * ux + uy * uy + uz * uz);
      ux *= invn; uy *= invn; uz *= invn;
      const double ci = -(sx * ux + sy * uy + sz * uz);
      if (!(ci >= 0.0)) {
        const double w = (fabs(ci) > 1.0) ? nan("") : 2.0 * ci;
        sx += w * ux; sy += w * uy; sz += w * uz;
      }
      rd0 = px * sx + py * sy + pz * sz;
      perp2 = nr2 - rd0 * rd0;
      c0 = py - rd0 * sy;
      X = rd0 * rd0;
      sq = fabs(rd0);
      outward = rd0 > 0.0;
    }
  }
  // Brightness.single uses ds[0 .. n-2] (brightness.py:65); NaN persists once it appears, so the last
  // used segment tells whether the ray is a NaN ray
  int first_nan = -1;
  if (count >= 2) {
    const double last_used = out[(size_t)(count - 2) * kDsStride];
    if (last_used != last_used) first_nan = 0;
  }
  g.nseg[o] = count;
  if (outf && count >= 1) outf[(size_t)(count - 1) * kDsStride] = 0.0f;
  // Brightness.single uses ds[0 .. n-2] (brightness.py:65); a NaN there makes every frequency NaN
  g.nanflag[o] = (first_nan >= 0 && first_nan <= count - 2) ? 1 : 0;
  if (g.prog) {
    // nothing lies below the last used segment: zeros from row count - 1 on (a NaN ray keeps its NaN rows), then the
    // ray counts into every chunk it has not published yet
    if (first_nan < 0)
      for (int l = max(count - 1, 0); l < S; ++l) out[(size_t)l * kDsStride] = 0.0;
    __threadfence();
    for (; pubc < g.npub; ++pubc) atomicAdd(g.prog + (size_t)(t >> 5) * g.npub + pubc, 1);
  }
}

// ---- descriptive ray fields of the compute_ds API (Ray.r4ds, and the latitude / longitude Ray.doppler is made of) ---
// raypath.py:186-187, 224: per step the reference records rNowMag (the shell radius at the latitude of the current
// point) and a Doppler factor of that point's latitude / longitude.  They describe the ray and feed nothing on the
// hot path, so they are not carried through ray_geometry_kernel: this kernel walks the positions r_i+1 = r_i + ds_i s
// (raypath.py:227) along the segments the geometry kernel produced, in the reference's own vector form.
__global__ void ray_fields_kernel(const __grid_constant__ GeoK g, double* __restrict__ out) {   // out [R][3][S]
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= g.R) return;
  const int S = g.L - 1;
  double* o_r = out + (size_t)r * 3 * S;
  double* o_lat = o_r + S;
  double* o_lng = o_lat + S;
  const int n = g.nseg[r];
  const double q2 = g.q * g.q, rNorm = g.radius[0];
  double zq;
  if (n <= 0 || !find_edge(g, g.b[2 * r], g.b[2 * r + 1], zq)) return;
  double ex, ey, ez;
  rot2planet(g, g.b[2 * r], g.b[2 * r + 1], zq, ex, ey, ez);
  const double kDeg = 57.295779513082320877;
  // the shell point / normal the reference starts from (raypath.py:141-156), then Snell at the first interface
  double v = ey / sqrt(ex * ex + ey * ey + ez * ez), sl, cl;
  lat_sc(v, sl, cl);
  const double hs = sqrt(ex * ex + ez * ez);
  const double sln = (hs > 0.0) ? ex / hs : 0.0, cln = (hs > 0.0) ? ez / hs : 1.0;
  double px = rNorm * cl * sln, py = g.q * rNorm * sl, pz = rNorm * cl * cln;
  const double inv = 1.0 / sqrt(sl * sl + q2 * cl * cl);
  const double nx = g.q * cl * sln * inv, ny = sl * inv, nz = g.q * cl * cln * inv;
  double sx, sy, sz;
  rot2planet(g, 0.0, 0.0, -1.0, sx, sy, sz);
  {
    const double nratio = g.n0 / g.n1;
    const double ci = -(sx * nx + sy * ny + sz * nz);
    const double st = nratio * sqrt(fmax(0.0, 1.0 - ci * ci));
    const double w = nratio * ci - sqrt(1.0 - st * st);
    sx = nratio * sx + w * nx; sy = nratio * sy + w * ny; sz = nratio * sz + w * nz;
  }
  const double* ds = g.ds + ds_tile_base(r, S);
  // step 0 is described by the edge point (raypath.py:141-146), step i + 1 by r_i+1 = r_i + ds_i s (raypath.py:227-231)
  o_r[0] = rNorm * sqrt(q2 * sl * sl + cl * cl);
  o_lat[0] = asin(v) * kDeg;
  o_lng[0] = atan2(ex, ez) * kDeg;
  for (int i = 0; i + 1 < n; ++i) {
    const double d = ds[(size_t)i * kDsStride];
    px = fma(d, sx, px); py = fma(d, sy, py); pz = fma(d, sz, pz);
    const double nr = sqrt(px * px + py * py + pz * pz);
    lat_sc(py / nr, sl, cl);
    o_r[i + 1] = g.radius[i + 1] * sqrt(q2 * sl * sl + cl * cl);   // Shape.rmag of the new point (shape.py:240-244)
    o_lat[i + 1] = asin(py / nr) * kDeg;
    o_lng[i + 1] = atan2(px, pz) * kDeg;
  }
}

// ==== 'gravity' shape (Shape._calcGeoid / _gravity, shape.py:141-221) =======================================
// The reference finds the shell radius at latitude pclat by marching from the equator in steps of 0.01 deg: at
// every grid latitude it evaluates the gravity vector (zonal harmonics, rotation + zonal wind), takes the local
// tangent and steps along it; what it returns is the state at the LAST grid latitude (k = len(np.arange(0, pclat +
// step, step)) - 1), not at pclat.  So every shape the ray loop can ask for is an entry of a table indexed by
// (hemisphere, k, layer): geoid_table_kernel fills it with one march per (layer, hemisphere) -- the operations of
// _gravity in their order -- and ray_geometry_gravity_kernel is raypath.compute_ds in the reference's own vector /
// trigonometric form with table look-ups for Shape.calcShape (the scalar recurrence of ray_geometry_kernel is
// specific to the ellipse).  Descriptive kernel, not tuned: the reference spends minutes per ray here.
struct GravK {
  int L, K, nJ, nvw;
  const double* radius;   // [L]
  const double* GM;       // [L]
  const double* Jn;       // [nJ]
  const double* vwlat;    // [nvw]
  const double* vwdat;
  double RJ, omega_m, latstep;
  double* rmag;           // [2][K][L]
  double* gamma;          // [2][K][L]
};

// np.interp(x, xp, fp) for ascending xp
__device__ double np_interp(const double* __restrict__ xp, const double* __restrict__ fp, int n, double x) {
  if (x <= xp[0]) return fp[0];
  if (x >= xp[n - 1]) return fp[n - 1];
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  const double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
  return slope * (x - xp[lo]) + fp[lo];
}

// Legendre polynomial P_n(x) (scipy.special.legendre(n)(x)) by the three-term recurrence
__device__ double legendre_p(int n, double x) {
  if (n == 0) return 1.0;
  double pm = 1.0, p = x;
  for (int m = 1; m < n; ++m) {
    const double pn = ((2.0 * m + 1.0) * x * p - m * pm) / (m + 1.0);
    pm = p; p = pn;
  }
  return p;
}

__global__ void geoid_table_kernel(const __grid_constant__ GravK g) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int hemi = blockIdx.y;                                // 0: north (and the equator), 1: south
  if (l >= g.L) return;
  const double kD2R = 3.14159265358979323846 / 180.0;
  const double latstep = hemi ? -g.latstep : g.latstep;
  const double dlat = latstep * kD2R;
  const double GM = g.GM[l];
  double r = g.radius[l];
  for (int k = 0; k < g.K; ++k) {
    const double latv = 0.0 + k * latstep;
    const double vw = np_interp(g.vwlat, g.vwdat, g.nvw, latv) / 1000.0;
    const double omega = g.omega_m + vw / (r * cos(latv * kD2R));
    // ---- _gravity(latv, lng, r, GM, omega) (shape.py:173-221)
    const double g_static = GM / (r * r);
    const double lat = latv * kD2R;
    const double nsl = (lat == 0.0) ? 1.0 : (lat > 0.0 ? 1.0 : -1.0);
    const double dphi = nsl * 0.00001;
    const double sp = sin(lat), sp1 = sin(lat + dphi), sp0 = sin(lat - dphi);
    double Sr = 0.0, Sp = 0.0;
    for (int i = 0; i < g.nJ; ++i) {
      const double pw = pow(g.RJ / r, (double)i);
      const double P = legendre_p(i, sp);
      Sr += (i + 1.0) * g.Jn[i] * pw * P;
      double dP = (legendre_p(i, sp1) - P) / dphi;
      dP += (P - legendre_p(i, sp0)) / dphi;
      dP *= 0.5;
      Sp += g.Jn[i] * pw * dP;
    }
    const double gr = g_static * (1.0 - Sr) - (2.0 / 3.0) * (omega * omega) * r * (1.0 - legendre_p(2, sp));
    const double dP2 = 3.0 * sp * sqrt(1.0 - sp * sp);
    const double gp = (1.0 / 3.0) * (omega * omega) * r * dP2 + g_static * Sp;
    const double gamma = atan2(gp, gr);
    const double ry = r * sin(lat), rz = r * cos(lat);
    const size_t o = ((size_t)hemi * g.K + k) * g.L + l;
    g.rmag[o] = sqrt(ry * ry + rz * rz);                      // Shape.rmag = |r_vec|
    g.gamma[o] = gamma;
    // ---- step along the tangent (shape.py:163-166)
    const double ty = cos(lat + gamma), tz = -sin(lat + gamma);
    const double ny = ry + r * dlat * ty, nz = rz + r * dlat * tz;
    r = sqrt(ny * ny + nz * nz);
  }
}

// Shape.calcShape(atm, radius[layer], pclat, dlng) for gtype 'gravity': shell radius, position and normal
struct GeoidState { double rmag, rx, ry, rz, nx, ny, nz; };
__device__ GeoidState geoid_lookup(const GravK& g, int layer, double pclat, double dlng) {
  const double kD2R = 3.14159265358979323846 / 180.0;
  const double nsp = (pclat == 0.0) ? 1.0 : (pclat > 0.0 ? 1.0 : -1.0);
  const double latstep = nsp * g.latstep;
  int k = (int)ceil((pclat + latstep) / latstep) - 1;         // len(np.arange(0, pclat + latstep, latstep)) - 1
  GeoidState st;
  if (!(pclat == pclat) || k >= g.K) {                        // NaN latitude (the ray is gone) / beyond the table
    st.rmag = st.rx = st.ry = st.rz = st.nx = st.ny = st.nz = nan("");
    return st;
  }
  if (k < 0) k = 0;
  const size_t o = ((size_t)(nsp < 0.0 ? 1 : 0) * g.K + k) * g.L + layer;
  const double rk = g.rmag[o], gamma = g.gamma[o];
  const double lat = (0.0 + k * latstep) * kD2R, lng = dlng * kD2R;
  const double vy = rk * sin(lat), vz = rk * cos(lat);
  // rotY(lng, [0, vy, vz]) = [sin(lng) vz, vy, cos(lng) vz]   (shape.py:285-290)
  st.rx = sin(lng) * vz; st.ry = vy; st.rz = cos(lng) * vz;
  st.rmag = sqrt(st.rx * st.rx + st.ry * st.ry + st.rz * st.rz);
  const double my = sin(lat + gamma), mz = cos(lat + gamma);
  st.nx = sin(lng) * mz; st.ny = my; st.nz = cos(lng) * mz;
  return st;
}

// raypath.compute_ds (raypath.py:108-273) for the 'gravity' shape, one thread per ray, plain ray order
__global__ void __launch_bounds__(128) ray_geometry_gravity_kernel(const __grid_constant__ GeoK g, const __grid_constant__ GravK gv,
                                                                   double* __restrict__ fields) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= g.R) return;
  const int S = g.L - 1;
  const double kDeg = 180.0 / 3.14159265358979323846;
  const double bx = g.b[2 * r], by = g.b[2 * r + 1];
  const double rNorm = g.radius[0];
  double* out = g.ds + ds_tile_base(r, S);
  g.nanflag[r] = 0;
  g.nseg[r] = -1;
  const double bb = bx * bx + by * by;
  if (!(bb < 1.0)) return;
  const double mu = sqrt(1.0 - bx * bx - by * by);
  // ---- findEdge (raypath.py:60-105) with the geoid of the outer shell
  double zq = 0.0;
  {
    const double z0 = sqrt(1.0 - bb) * 1.01;
    const int ntrial = (int)ceil(z0 / 0.005);
    double d_prev = 0.0, z_prev = 0.0;
    bool hit = false;
    for (int t = 0; t < ntrial; ++t) {
      const double z = z0 + t * (-0.005);
      double px, py, pz;
      rot2planet(g, bx, by, z, px, py, pz);
      const double nb = sqrt(px * px + py * py + pz * pz);
      const double r1 = nb * rNorm;
      const GeoidState st = geoid_lookup(gv, 0, asin(py / nb) * kDeg, atan2(px, pz) * kDeg);
      const double d = r1 - st.rmag;
      if (r1 < st.rmag) {
        zq = (t == 0) ? z : ((z_prev - z) / (d_prev - d)) * (0.0 - d) + z;
        hit = true;
        break;
      }
      d_prev = d; z_prev = z;
    }
    if (!hit) return;
  }
  double ex, ey, ez;
  rot2planet(g, bx, by, zq, ex, ey, ez);
  ex *= rNorm; ey *= rNorm; ez *= rNorm;
  double pclat = asin(ey / sqrt(ex * ex + ey * ey + ez * ez)) * kDeg;
  double dlng = atan2(ex, ez) * kDeg;
  GeoidState st = geoid_lookup(gv, 0, pclat, dlng);
  double px = st.rx, py = st.ry, pz = st.rz;                  // r[0] = geoid.r
  double sx, sy, sz;
  rot2planet(g, 0.0, 0.0, -1.0, sx, sy, sz);
  double t_inc = acos(-(sx * st.nx + sy * st.ny + sz * st.nz));
  double nratio = g.n0 / g.n1;
  double t_tran = asin(nratio * sin(t_inc));
  int layer = 0, count = 0;
  while (true) {
    const double w = nratio * cos(t_inc) - cos(t_tran);       // s += ... n  (raypath.py:179-180)
    sx = nratio * sx + w * st.nx; sy = nratio * sy + w * st.ny; sz = nratio * sz + w * st.nz;
    const double rNow = st.rmag;
    const GeoidState nx_ = geoid_lookup(gv, layer + 1, pclat, dlng);
    const double rNext = nx_.rmag;
    const double rdots = px * sx + py * sy + pz * sz;
    double ds = -rdots - sqrt(rdots * rdots + rNext * rNext - rNow * rNow);
    if (ds < 0.0) break;                                      // raypath.py:212-216
    if (g.limb == RB_LIMB_SEC) ds = fabs(rNext - rNow) / mu;  // raypath.py:218-219
    out[(size_t)layer * kDsStride] = ds;
    if (fields) {
      double* f = fields + (size_t)r * 3 * S;
      f[layer] = rNow; f[S + layer] = pclat; f[2 * S + layer] = dlng;
    }
    ++count;
    if (ds != ds) {
      // below the tangent shell every later segment is NaN (the reference's march raises on the NaN latitude here;
      // the ellipse path carries the NaN on, and so does this one)
      for (++layer; layer < S; ++layer) out[(size_t)layer * kDsStride] = ds;
      count = S;
      break;
    }
    px = fma(ds, sx, px); py = fma(ds, sy, py); pz = fma(ds, sz, pz);
    const double nr = sqrt(px * px + py * py + pz * pz);
    pclat = asin(py / nr) * kDeg;
    dlng = atan2(px, pz) * kDeg;
    st = geoid_lookup(gv, layer + 1, pclat, dlng);
    ++layer;
    t_inc = acos(-(sx * st.nx + sy * st.ny + sz * st.nz));
    if (layer + 1 >= g.L) break;                              // raypath.py:255-264 (IndexError exit)
    nratio = 1.0;                                             // raypath.py:257
    t_tran = asin(nratio * sin(t_inc));
  }
  g.nseg[r] = count;
  if (count >= 2) {
    const double last_used = out[(size_t)(count - 2) * kDsStride];
    if (last_used != last_used) g.nanflag[r] = 1;
  }
}

// [S][Rpad] slab -> [R][S] ray-major (only for the compute_ds API that returns Ray.ds)
__global__ void ds_transpose_kernel(const double* __restrict__ slab, long long R, long long Rpad, int S,
                                    const int* __restrict__ nseg, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int s0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int s = s0 + j;
    const long long r = r0 + threadIdx.x;
    tile[j][threadIdx.x] = (s < S && r < R) ? slab[ds_tile_base(r, S) + (size_t)s * kDsStride] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + j;
    const int s = s0 + threadIdx.x;
    if (r < R && s < S) {
      const int n = nseg[r];
      out[r * S + s] = (s < n) ? tile[threadIdx.x][j] : 0.0;
    }
  }
}

__global__ void ds_to_slab_kernel(const double* __restrict__ in, long long R, long long Rpad, int S,
                                  const int* __restrict__ nseg, int* __restrict__ nanflag,
                                  double* __restrict__ slab) {
  __shared__ double tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int s0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + j;
    const int s = s0 + threadIdx.x;
    const double v = (r < R && s < S) ? in[r * S + s] : 0.0;
    tile[j][threadIdx.x] = v;
    if (v != v && s <= nseg[r] - 2) nanflag[r] = 1;  // benign race: every writer stores 1
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int s = s0 + j;
    const long long r = r0 + threadIdx.x;
    if (s < S && r < Rpad) slab[ds_tile_base(r, S) + (size_t)s * kDsStride] = tile[threadIdx.x][j];
  }
}

// ---- E2(x): exponential integral of order 2 (scipy.special.expn(2, x), brightness.py:96) ----------
// Same algorithm as the Cephes expn routine scipy wraps: power series for x <= 1, continued fraction
// for x > 1.
__device__ double expn2(double x) {
  const double EUL = 0.57721566490153286060, MACHEP = 1.11022302462515654042E-16, BIG = 1.44115188075855872E+17;
  if (x != x) return x;
  if (x < 0.0) return nan("");
  if (x > 7.09782712893383996843E2) return 0.0;
  if (x == 0.0) return 1.0;  // 1/(n-1)
  if (x > 1.0) {
    int kk = 1;
    double pkm2 = 1.0, qkm2 = x, pkm1 = 1.0, qkm1 = x + 2.0, ans = pkm1 / qkm1, t;
    do {
      kk += 1;
      double yk, xk;
      if (kk & 1) { yk = 1.0; xk = 2.0 + (kk - 1) / 2; }
      else { yk = x; xk = kk / 2; }
      const double pk = pkm1 * yk + pkm2 * xk;
      const double qk = qkm1 * yk + qkm2 * xk;
      if (qk != 0) { const double r = pk / qk; t = fabs((ans - r) / r); ans = r; }
      else t = 1.0;
      pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
      if (fabs(pk) > BIG) { pkm2 /= BIG; pkm1 /= BIG; qkm2 /= BIG; qkm1 /= BIG; }
    } while (t > MACHEP);
    return ans * exp(-x);
  }
  // power series
  double psi = -EUL - log(x) + 1.0;  // + sum_{i=1}^{n-1} 1/i with n = 2
  const double z = -x;
  double xk = 0.0, yk = 1.0, pk = 1.0 - 2.0, ans = 1.0 / pk, t;
  do {
    xk += 1.0;
    yk *= z / xk;
    pk += 1.0;
    if (pk != 0.0) ans += yk / pk;
    t = (ans != 0.0) ? fabs(yk / ans) : 1.0;
  } while (t > MACHEP);
  return z * psi - ans;  // pow(z, n-1) * psi / Gamma(n) - ans
}

struct RtK {
  int L, F;
  long long R, Rpad;
  const double* alpha;  // [L][F]
  const double* T;      // [L]
  const double4* prep;  // [F/8][L-1][8] interleaved loop operands (rays-major kernel only, see rt_prepare_kernel)
  unsigned long long* step_counter;  // optional: [0] (ray, freq, segment) steps actually integrated, [1] of which in phase A
  const double* exp_tab;  // 2^(j/1024) (rays-major kernel only)
  const float* dsf;               // mixed-precision rays-major kernel: float segments (see GeoK::dsf)
  const void* prepm;              // mixed-precision rays-major kernel: operand rows (see rt_prepare_mixed_kernel)
  RtProgress progress;    // optional per-chunk completion counters (rays-major kernel only)
  const double2* prep2;   // [F/16][L-1][8][3] pair operands (rt_integrate_pairs_kernel, see rt_prepare_pairs_kernel)
  unsigned fgroups, ntiles;  // rays-major launches are 1-D: block = tile * fgroups + frequency group
  const double* alpha0;      // Doppler form (rb_rt_desc::alpha0): dtau_i = (alpha0[i] + alpha[i+1]) ds_i / 2, or null
  const int* geo_prog;       // streamed geometry: [tiles][geo_npub] rays past each chunk (see GeoK::prog), or null
  int geo_npub;
  int fg_reverse;            // launch the frequency groups last to first (behind a running trace: rb_launch_integrate)
  unsigned nparts;           // launch order: the tile list in nparts parts, inside a part frequency group by frequency
                             // group (see rb_launch_integrate); 0: block = tile * fgroups + frequency group
  unsigned part_tiles;       // != 0: the device cuts the list into parts of about part_tiles tile blocks (<= nparts parts)
  unsigned tile_blocks;      // upper bound of the CTAs per frequency group (the grid holds nparts more per group)
  const int* cidx;        // compacted launch: list position -> ray index (null: ds / nseg / nanflag are per ray index)
  const int* ncomp;       // compacted launch: length of the list
  const double* ds;     // [S][Rpad]
  const int* nseg;      // [R]
  const int* nanflag;   // [R]
  void* out_Tb;         // [R][F]
  double* out_intW;     // [R][F] or null
  int out_f32;
  int disc;
  double tau_cut;
  // profile outputs for one ray (RPT = 1 launches only)
  long long profile_ray;
  double *out_tau, *out_W, *out_Tblyr;  // [F][S]
};

// thread = (frequency lane, RPT consecutive rays); blockDim = (32, WY); grid = (ray groups, freq groups)
template <int RPT, bool DISC, bool PROFILE>
__global__ void __launch_bounds__(128) rt_integrate_kernel(const __grid_constant__ RtK k) {
  const int fi = blockIdx.y * 32 + threadIdx.x;
  const long long rg = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const long long r0 = rg * RPT;
  if (r0 >= k.R) return;
  const bool fvalid = fi < k.F;
  const int f = fvalid ? fi : k.F - 1;
  const int S = k.L - 1;

  int n[RPT];
  int nmax = 0;
  bool isnan_ray[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    n[j] = (r0 + j < k.R) ? k.nseg[r0 + j] : -1;
    isnan_ray[j] = (r0 + j < k.R) && !PROFILE && k.nanflag[r0 + j] != 0;
    if (!isnan_ray[j]) nmax = max(nmax, n[j]);  // rays that end NaN need no integration at all
  }
  double tau[RPT], Wp[RPT], iW[RPT], Tb[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) tau[j] = Wp[j] = iW[j] = Tb[j] = 0.0;

  if (PROFILE && fvalid && nmax > 0) {
    k.out_tau[(size_t)f * S] = 0.0; k.out_W[(size_t)f * S] = 0.0; k.out_Tblyr[(size_t)f * S] = 0.0;
  }
  double a0 = k.alpha0 ? k.alpha0[f] : k.alpha[f];
  double T0 = k.T[0];
  // a ray stops once tau >= tau_cut, by the high words (the rule of the rays-major kernels; INFINITY disables it)
  const int cut_hi = __double2hiint(k.tau_cut);
  // brightness.py:65: for i in range(len(ds) - 1)
  for (int i = 0; i + 1 < nmax; ++i) {
    const double a1 = k.alpha[(size_t)(i + 1) * k.F + f];
    const double T1 = k.T[i + 1];
    double dsv[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) dsv[j] = k.ds[ds_tile_base(r0 + j, S) + (size_t)i * kDsStride];
    const double asum = a0 + a1;
    bool live = false;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      if (i + 1 < n[j] && !isnan_ray[j] && __double2hiint(tau[j]) < cut_hi) {
        const double h = dsv[j] * kKmToCm * 0.5;        // ds/2 in cm
        tau[j] = tau[j] + asum * h;                      // dtau = (a0 + a1) * ds / 2
        const double W = DISC ? 2.0 * a1 * expn2(tau[j]) : a1 * exp(-tau[j]);
        iW[j] += (W + Wp[j]) * h;
        Tb[j] += (T1 * W + T0 * Wp[j]) * h;
        Wp[j] = W;
        live = true;
      }
      if (PROFILE && fvalid && i + 1 < n[j]) {
        k.out_tau[(size_t)f * S + i + 1] = tau[j];
        k.out_W[(size_t)f * S + i + 1] = Wp[j];
        k.out_Tblyr[(size_t)f * S + i + 1] = Tb[j];
      }
    }
    if (!PROFILE && !live) break;
    a0 = k.alpha0 ? k.alpha0[(size_t)(i + 1) * k.F + f] : a1;   // (Doppler: the upper node of the next step has its own value)
    T0 = T1;
  }
  if (!fvalid) return;
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const long long r = r0 + j;
    if (r >= k.R) continue;
    double v;
    double w = iW[j];
    if (n[j] < 0) v = kTcmb;                               // off planet (brightness.py:46-51)
    else if (isnan_ray[j]) v = w = nan("");                // NaN segment below the tangent shell
    else v = (Tb[j] < kTcmb) ? kTcmb : Tb[j] / iW[j];      // brightness.py:109-113
    const size_t o = (size_t)r * k.F + f;
    if (k.out_f32) reinterpret_cast<float*>(k.out_Tb)[o] = (float)v;
    else reinterpret_cast<double*>(k.out_Tb)[o] = v;
    if (k.out_intW) k.out_intW[o] = (n[j] < 0) ? 0.0 : w;
  }
}

// ---- disc-averaged integration (brightness.py:95-96: W = 2 a E2(tau)) -----------------------------------------
// One CTA per (ray, frequency).  The exponential integral costs hundreds of instructions per layer (Cephes series /
// continued fraction), and a thread that walks the ~1000 layers of a ray one after the other spends 16 ms on it
// (rt_integrate_kernel<1, true, *>, round 1).  Only the optical depth is a recurrence: thread 0 accumulates it (one
// FMA-class step per layer, the operations and the order of the sequential kernel), all threads evaluate E2 of the
// layers in parallel, and thread 0 adds the weighting-function and brightness sums in layer order again -- the same
// operations in the same order as the sequential kernel, so the same bits.  Optional profile outputs
// (Brightness.tau / .W / .Tb_lyr, brightness.py:118-120) as in rt_integrate_kernel<1, true, true>.
constexpr int kDiscThreads = 256;
constexpr size_t disc_smem_bytes(int L) { return 6 * (size_t)L * sizeof(double); }
__global__ void __launch_bounds__(kDiscThreads) rt_disc_kernel(const __grid_constant__ RtK k, int profile) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  double* const s_tau = reinterpret_cast<double*>(s_raw);     // [L] tau after step i at index i + 1
  double* const s_W = s_tau + k.L;                            // [L] W after step i at index i + 1
  double* const s_a = s_W + k.L;                              // [L] this frequency's column of the alpha slab
  double* const s_h = s_a + k.L;                              // [L] ds_i / 2 in cm
  double* const s_T = s_h + k.L;                              // [L]
  double* const s_a0 = s_T + k.L;                             // [L] upper-node absorption of step i (alpha0, else = s_a)
  __shared__ int s_last;                                      // number of steps taken
  const int f = blockIdx.y;
  const long long r = blockIdx.x;
  const int S = k.L - 1;
  const int n = k.nseg[r];
  const bool nanray = !profile && k.nanflag[r] != 0;
  const int nsteps = (n > 0 && !nanray) ? n - 1 : 0;          // brightness.py:65: i = 0 .. len(ds) - 2
  const int cut_hi = __double2hiint(k.tau_cut);
  const double* ds = k.ds + ds_tile_base(r, S);
  // the operands of the two sequential passes, fetched by all threads (the passes themselves then only touch
  // shared memory: a thread that walks global memory alone waits an L2 round trip per layer)
  for (int i = threadIdx.x; i <= nsteps && i < k.L; i += kDiscThreads) {
    s_a[i] = k.alpha[(size_t)i * k.F + f];
    s_a0[i] = k.alpha0 ? k.alpha0[(size_t)i * k.F + f] : s_a[i];
    s_T[i] = k.T[i];
    if (i < nsteps) s_h[i] = ds[(size_t)i * kDsStride] * kKmToCm * 0.5;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tau = 0.0;
    int i = 0;
    s_tau[0] = 0.0;
    for (; i < nsteps; ++i) {
      if (!(__double2hiint(tau) < cut_hi)) break;             // the rule of every integration kernel of this file
      tau = tau + (s_a0[i] + s_a[i + 1]) * s_h[i];
      s_tau[i + 1] = tau;
    }
    s_last = i;
  }
  __syncthreads();
  const int last = s_last;
  for (int i = threadIdx.x; i < last; i += kDiscThreads) s_W[i + 1] = 2.0 * s_a[i + 1] * expn2(s_tau[i + 1]);
  if (threadIdx.x == 0) s_W[0] = 0.0;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double iW = 0.0, Tb = 0.0, Wp = 0.0, T0 = s_T[0];
  if (profile && n > 0) { k.out_tau[(size_t)f * S] = 0.0; k.out_W[(size_t)f * S] = 0.0; k.out_Tblyr[(size_t)f * S] = 0.0; }
  for (int i = 0; i < (profile ? nsteps : last); ++i) {
    if (i < last) {
      const double h = s_h[i];
      const double W = s_W[i + 1], T1 = s_T[i + 1];
      iW += (W + Wp) * h;
      Tb += (T1 * W + T0 * Wp) * h;
      Wp = W;
      T0 = T1;
    }
    if (profile) {                                            // beyond tau_cut the profiles repeat their last values
      k.out_tau[(size_t)f * S + i + 1] = s_tau[min(i, last - 1) + 1];
      k.out_W[(size_t)f * S + i + 1] = Wp;
      k.out_Tblyr[(size_t)f * S + i + 1] = Tb;
    }
  }
  double v, w = iW;
  if (n < 0) v = kTcmb;
  else if (nanray) v = w = nan("");
  else v = (Tb < kTcmb) ? kTcmb : Tb / iW;
  const size_t o = (size_t)r * k.F + f;
  if (k.out_f32) reinterpret_cast<float*>(k.out_Tb)[o] = (float)v;
  else reinterpret_cast<double*>(k.out_Tb)[o] = v;
  if (k.out_intW) k.out_intW[o] = (n < 0) ? 0.0 : w;
}

// ---- exp(-tau) for the weighting function ------------------------------------------------------------
// tau >= 0, N = kExpTab.  exp(-tau) = 2^k * 2^(j/N) * exp(r) with -tau N log2(e) = N k + j + N r/ln2,
// |r| <= h = ln2/(2N), the 2^(j/N) table in shared memory, exp(r) by an economised polynomial (the first
// dropped Taylor term is folded onto the lower coefficients, Chebyshev-style):
//   RB_EXP_DEG 2:  1 + (1 + h^2/8) r + r^2/2                         |rel. error| <= h^3/24
//   RB_EXP_DEG 3:  (1 - h^4/192) + r + (1/2 + h^2/24) r^2 + r^3/6    |rel. error| <= h^4/192
// Branch-free; libdevice exp() costs ~25 FP64 instructions with its special-case paths.  The error bound of
// the shipped combination is printed in DESIGN.md 3.3; it is far inside the 0.01 K bar (tests hold 1e-4 K
// against the reference and 1e-7 K between kernels).  tau beyond the underflow point gives exactly 0; NaN
// is handled by the caller.  The constants are fetched once per thread into registers (see pin()).
// The table size trades shared-memory wavefronts for polynomial degree: neighbouring lanes are neighbouring
// rays, their tau differ by a few per cent, so with a small table their entries fall into a short window
// of consecutive entries (16 consecutive doubles never conflict), with a large one they are random.
#ifndef RB_EXP_TABLOG
#define RB_EXP_TABLOG 10
#endif
#ifndef RB_EXP_DEG
#define RB_EXP_DEG 2
#endif
constexpr int kExpTabLog = RB_EXP_TABLOG;
constexpr int kExpTab = 1 << kExpTabLog;
// RB_EXP_REP copies of every entry, interleaved ([entry][copy]): lane l reads copy l % RB_EXP_REP.  With 16 copies
// each lane of a half-warp owns its own pair of banks, so a lookup costs exactly two shared-memory wavefronts
// whatever the 32 indices are (a single copy: ~5.5 wavefronts for 32 random entries of a 1024-entry table).
#ifndef RB_EXP_REP
#define RB_EXP_REP 1
#endif
constexpr int kExpRep = RB_EXP_REP;
static_assert(kExpRep == 1 || kExpRep == 2 || kExpRep == 4 || kExpRep == 8 || kExpRep == 16, "copies per table entry");
constexpr int kExpTabDoubles = kExpTab * kExpRep;
constexpr double kExpH = 0.69314718055994530942 / (2.0 * kExpTab);
__device__ double c_expc[8] = {
    -1.4426950408889634074 * kExpTab,   // -N log2(e)
    6755399441055744.0,                 // 2^52 + 2^51: adding it rounds to the nearest integer (kept in the low word)
    -0.69314718055994530942 / kExpTab,  // -ln2/N
    1.0 / 6.0,
#if RB_EXP_DEG == 2
    0.5, 1.0, 1.0 + kExpH * kExpH / 8.0,
#else
    0.5 + kExpH * kExpH / 24.0, 1.0, 1.0 - kExpH * kExpH * kExpH * kExpH / 192.0,
#endif
    0.0};
// Small optical depths (about half of all executed segment-steps: tau grows exponentially with depth) need no
// range reduction: on 0 <= tau <= H = 2^-11 the economised quadratic
//   exp(-tau) ~ E [1 - (1 + m^2/8)(tau - m) + (tau - m)^2 / 2],  m = H/2, E = exp(-m)
// is within m^3/24 = 6e-13 (relative; the table path holds 1.6e-12) -- two FMAs instead of the table path's
// six FP64 instructions, one shared-memory read and four integer instructions.  Measured on C4: H = 2^-15
// 4.22 ms, 2^-13 4.11 ms, 2^-11 3.95 ms against 4.50 ms without the path (RB_EXP_SMALL_LOG = 0).
#ifndef RB_EXP_SMALL_LOG
#define RB_EXP_SMALL_LOG 11
#endif
constexpr double kSmallH = (RB_EXP_SMALL_LOG > 0) ? 1.0 / (double)(1 << (RB_EXP_SMALL_LOG > 0 ? RB_EXP_SMALL_LOG : 1)) : 0.0;
constexpr double kSmallM = 0.5 * kSmallH;
constexpr double exp_neg_small(double m) {   // exp(-m) for m << 1 (compile time)
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 12; ++i) { term *= -m / i; sum += term; }
  return sum;
}
constexpr double kSmallE = exp_neg_small(kSmallM);
constexpr double kSmallK1 = 1.0 + kSmallM * kSmallM / 8.0;
__device__ double c_small[4] = {
    kSmallE * (1.0 + kSmallK1 * kSmallM + 0.5 * kSmallM * kSmallM),   // a0
    -kSmallE * (kSmallK1 + kSmallM),                                  // a1
    0.5 * kSmallE,                                                    // a2
    0.0};
// The same polynomial with the square completed, E/2 [(tau - m - k1)^2 + (2 - k1^2)], as the pair kernel evaluates it:
// one DADD and one DFMA that read two registers each (the Horner form is two DFMAs with three register operands, which
// the register file feeds at 2/3 of the pipe rate -- tools/probe_pipes.py); the factor E/2 is applied to the
// accumulators once, when the ray leaves the small-tau phase.
__device__ double c_small2[4] = {
    -(kSmallM + kSmallK1),                                            // cb: u = tau + cb
    2.0 - kSmallK1 * kSmallK1,                                        // cd: p = u^2 + cd
    0.5 * kSmallE,                                                    // a2
    0.0};
// mixed-precision kernel: -2^23 log2 e and the rounding constant 2^52 + 2^51 with 0x3F000000 folded into its low word
__device__ double c_mxc[2] = {-1.4426950408889634074 * 8388608.0, 6755399441055744.0 + 1056964608.0};
#if RB_EXP_DEG == 2
#define RB_EXP_POLY(x) fma(fma((x), c2, ce), (x), c1)
#else
#define RB_EXP_POLY(x) fma(fma(fma((x), c3, c2), (x), c1), (x), ce)
#endif

__global__ void exp_tab_init_kernel(double* tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kExpTabDoubles) tab[t] = exp2((double)(t / kExpRep) / (double)kExpTab);
}

// per-(layer, freq) operands of the integration loop, hoisted out of the per-ray work and interleaved so
// that one 32-byte shared-memory read fetches them (kHalfCm = 0.5 * 1e5 folds ds [km] -> ds/2 [cm],
// brightness.py:66).  Grouped by blocks of 8 frequencies so that the operands one CTA needs for a chunk
// of segments are contiguous in HBM:
//   prep[fg][i][fl] = { (a_i + a_i+1) kHalfCm,  a_i+1 kHalfCm,  T_i+1 a_i+1 kHalfCm,  0 }
//   f = 8 fg + fl,  i = 0 .. L-2,  zero for f >= F
__global__ void rt_prepare_kernel(const double* __restrict__ alpha, const double* __restrict__ T, int L, int F,
                                  int ngroups, double4* __restrict__ prep) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int Lm1 = L - 1;
  if (idx >= ngroups * Lm1 * 8) return;
  const int fl = idx & 7;
  const int i = (idx >> 3) % Lm1;
  const int fg = (idx >> 3) / Lm1;
  const int f = fg * 8 + fl;
  double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
  if (f < F) {
    const double kHalfCm = 0.5 * kKmToCm;
    const double a0 = alpha[(size_t)i * F + f], a1 = alpha[(size_t)(i + 1) * F + f];
    v = make_double4((a0 + a1) * kHalfCm, a1 * kHalfCm, (T[i + 1] * a1) * kHalfCm, 0.0);
  }
  prep[idx] = v;
}

// fetch a loop-invariant FP64 constant into a register pair with a plain global load: ptxas cannot
// re-materialise that inside the loop (it does so for immediates and constant-bank values with
// UMOV + MOV / LDC, which costs issue slots in this issue-bound kernel)
__device__ __forceinline__ double pin(const double* p) {
  double x;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(x) : "l"(p));
  return x;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
template <int OFF>
__device__ __forceinline__ void cp_async16_at(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%2], 16;" ::"r"(dst), "l"(src), "n"(OFF) : "memory");
}
template <int N, int ROUND = 4096>
__device__ __forceinline__ void cp_rounds(unsigned dst, const void* src) {
  if constexpr (N > 0) {
    cp_rounds<N - 1, ROUND>(dst, src);
    cp_async16_at<(N - 1) * ROUND>(dst, src);
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- ray tile of a CTA of the rays-major kernels ----------------------------------------------------------
// Plain launch: CTA y owns rays 32 y .. 32 y + 31.  Compacted launch (k.cidx): CTA y owns 32 consecutive entries
// of the list of rays that hit the planet; ds / nseg / nanflag are indexed by list position, results are scattered
// to the ray index.  When the host pipelines the copy-out (k.progress.done) the launch order is rotated to the
// middle of the ray list: an image starts and ends with rows of sky, so the centre rows -- the long ones -- start
// first and the chunks complete evenly in time.
struct RayTile {
  unsigned fg, by;     // frequency group / position in launch order (frequency groups of a tile are adjacent: they
                       // share the ds tile through L2)
  unsigned tile;       // tile of the ds slab
  long long t;         // this lane: index into ds tiles / nseg / nanflag
  long long r;         // this lane: ray index (row of out_Tb)
  bool in;             // this lane carries a ray
  bool dead;           // whole CTA beyond the end of the compacted list
  int first_r, last_r; // compacted launch: ray index of the first / last ray of the tile
};
// (tiles_per_cta > 1: the CTA's warps take consecutive tiles, `sub` = this warp's)
__device__ __forceinline__ RayTile map_ray_tile(const RtK& k, unsigned tiles_per_cta = 1, unsigned sub = 0) {
  RayTile m;
  m.dead = false;
  m.first_r = m.last_r = 0;
  unsigned bcta;
  bool beyond = false;
  if (k.nparts) {
    // CTAs (tile blocks) that exist: known on the device only for a compacted launch
    const unsigned nt = k.cidx ? ((unsigned)((*k.ncomp + 31) >> 5) + tiles_per_cta - 1) / tiles_per_cta : k.tile_blocks;
    // (part_tiles: parts of about that many tile blocks, at most k.nparts of them -- the host sized the grid for
    //  k.nparts without knowing how many rays hit)
    const unsigned np = k.part_tiles ? min(k.nparts, max(1u, (nt + k.part_tiles / 2) / k.part_tiles)) : k.nparts;
    const unsigned cs = (nt + np - 1) / np;                    // tile blocks per part
    const unsigned per = cs * k.fgroups;
    const unsigned part = per ? blockIdx.x / per : np;
    const unsigned local = blockIdx.x - part * per;
    m.fg = cs ? local / cs : 0;
    bcta = part * cs + (local - m.fg * cs);
    if (k.fg_reverse && m.fg < k.fgroups) m.fg = k.fgroups - 1 - m.fg;
    beyond = part >= np || bcta >= nt;
  } else {
    bcta = blockIdx.x / k.fgroups;
    m.fg = blockIdx.x - bcta * k.fgroups;
  }
  m.by = bcta * tiles_per_cta + sub;
  if (k.cidx) {
    const int nc = *k.ncomp;
    const unsigned ntile = (unsigned)((nc + 31) >> 5);
    m.dead = beyond || m.by >= ntile;
    unsigned tile = m.by + (k.progress.done ? (unsigned)k.ncomp[1] : 0u);
    if (tile >= ntile) tile -= ntile;
    m.tile = tile;
    m.t = (long long)tile * 32 + threadIdx.x;
    m.in = !m.dead && m.t < nc;
    m.r = m.in ? k.cidx[m.t] : 0;
    if (!m.dead && k.progress.done) {
      // the tile's rays come from anywhere in its super-block(s) of the sorted list: smallest / largest ray index
      // (all 32 lanes of every warp get here together: blockDim.x == 32)
      m.first_r = (int)__reduce_min_sync(0xffffffffu, m.in ? (unsigned)m.r : 0x7fffffffu);
      m.last_r = (int)__reduce_max_sync(0xffffffffu, m.in ? (unsigned)m.r : 0u);
    }
  } else {
    m.dead = beyond || m.by >= k.ntiles;
    unsigned tile = m.by + (unsigned)k.progress.shift;
    if (tile >= k.ntiles) tile -= k.ntiles;
    m.tile = tile;
    m.t = m.r = (long long)tile * 32 + threadIdx.x;
    m.in = !m.dead && m.r < k.R;
  }
  return m;
}

// chunks (bit c) of the copy-out pipeline whose rows intersect the ray-index tiles [m0, m1] (compacted launches:
// a tile of the list spans rays of several image rows).  Chunk c is the tiles (p + shift) mod ntiles,
// p in [cut[c], cut[c+1]), of the plain ray order.
__device__ __forceinline__ unsigned progress_touched(const RtProgress& pg, int ntiles, int m0, int m1) {
  unsigned mask = 0;
  for (int c = 0; c < pg.nchunks; ++c) {
    const int len = pg.cut[c + 1] - pg.cut[c];
    if (len <= 0) continue;
    int a = pg.cut[c] + pg.shift;
    if (a >= ntiles) a -= ntiles;
    const int e = a + len;                               // [a, e) possibly wrapping past ntiles
    const bool hit = (e <= ntiles) ? (m0 < e && m1 >= a) : ((m1 >= a) || (m0 < e - ntiles));
    if (hit) mask |= 1u << c;
  }
  return mask;
}

// this CTA's results are in global memory: count it in its chunk(s) (the host's copy stream waits on the counters
// with a stream memory operation and then copies the chunk device -> host)
// (per_warp: the warp is the unit that finished a tile -- rt_integrate_pairs_kernel<true>)
__device__ __forceinline__ void progress_report(const RtK& k, const RayTile& m, int tid, bool per_warp = false) {
  if (!k.progress.done) return;
  __threadfence();
  if (per_warp) { __syncwarp(); if (threadIdx.x != 0) return; }
  else { __syncthreads(); if (tid != 0) return; }
  if (k.cidx) {
    unsigned mask = progress_touched(k.progress, (int)k.ntiles, m.first_r >> 5, m.last_r >> 5);
    for (; mask; mask &= mask - 1) atomicAdd(k.progress.done + (__ffs(mask) - 1), 1u);
  } else {
    int c = 0;
    while (c + 1 < k.progress.nchunks && (int)m.by >= k.progress.cut[c + 1]) ++c;
    atomicAdd(k.progress.done + c, 1u);
  }
}

// Compacted launch, before the integration: done[c] = kProgressTarget - (CTAs that will report into chunk c), so that
// every counter ends at kProgressTarget -- the value the host waits for, known without reading the list back.
__global__ void __launch_bounds__(256) rt_progress_init_kernel(RtProgress pg, const int* __restrict__ cidx,
                                                               const int* __restrict__ ncomp, int ntiles,
                                                               unsigned fgroups) {
  __shared__ unsigned s_cnt[kMaxProgressChunks];
  if (threadIdx.x < kMaxProgressChunks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int nc = *ncomp;
  const int nct = (nc + 31) >> 5;
  for (int tl = threadIdx.x; tl < nct; tl += blockDim.x) {
    int first = 0x7fffffff, lastr = 0;
    for (int e = tl * 32; e < min(nc, tl * 32 + 32); ++e) {
      const int rr = cidx[e];
      first = min(first, rr);
      lastr = max(lastr, rr);
    }
    unsigned mask = progress_touched(pg, ntiles, first >> 5, lastr >> 5);
    for (; mask; mask &= mask - 1) atomicAdd(&s_cnt[__ffs(mask) - 1], 1u);
  }
  __syncthreads();
  if ((int)threadIdx.x < pg.nchunks) pg.done[threadIdx.x] = kProgressTarget - s_cnt[threadIdx.x] * fgroups;
}

// rays that miss the planet see the sky (brightness.py:46-51); compacted launches never visit them
__global__ void rt_fill_miss_kernel(const double* __restrict__ zq, long long R, int F, void* out_Tb, double* out_intW,
                                    int out_f32) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * F) return;
  const long long r = idx / F;
  const double z = zq[r];
  if (z == z) return;
  if (out_f32) reinterpret_cast<float*>(out_Tb)[idx] = (float)kTcmb;
  else reinterpret_cast<double*>(out_Tb)[idx] = kTcmb;
  if (out_intW) out_intW[idx] = 0.0;
}
// the same with four frequencies per thread and 16-byte stores (F a multiple of 4, outputs 16-byte aligned)
__global__ void rt_fill_miss4_kernel(const double* __restrict__ zq, long long R, int F4, void* out_Tb, double* out_intW,
                                     int out_f32) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = R * F4;
  if (idx >= n) return;
  const long long r = n < 0x7fffffffLL ? (long long)((unsigned)idx / (unsigned)F4) : idx / F4;
  const double z = zq[r];
  if (z == z) return;
  if (out_f32) {
    const float t = (float)kTcmb;
    reinterpret_cast<float4*>(out_Tb)[idx] = make_float4(t, t, t, t);
  } else {
    double2* o = reinterpret_cast<double2*>(out_Tb) + 2 * idx;
    o[0] = make_double2(kTcmb, kTcmb);
    o[1] = make_double2(kTcmb, kTcmb);
  }
  if (out_intW) {
    double2* w = reinterpret_cast<double2*>(out_intW) + 2 * idx;
    w[0] = make_double2(0.0, 0.0);
    w[1] = make_double2(0.0, 0.0);
  }
}

// thread = (ray, frequency); lanes = 32 consecutive rays; the CTA's 8 warps are the 8 frequencies of one
// frequency group, all working on the same 32 rays.  Both operand streams are staged through shared memory
// in chunks of kChunk segments with cp.async (LDGSTS), three buffers deep:
//   ds tile    (kChunk+1) x 32 doubles   -- one contiguous piece of the tiled ds slab
//   prep tile   kChunk x 8 double4       -- one contiguous piece of the grouped operand slab
// so the segment loop only touches shared memory (LDS latency, no scoreboard-limited global loads in flight)
// and HBM / L2 latency is covered two chunks ahead.  One __syncthreads_or per chunk both publishes the
// landed chunk and tells the CTA when every (ray, freq) has finished (tau > tau_cut or out of segments).
// Frequency groups of one ray group are adjacent in launch order (blockIdx.x) so they share the ds tile
// through L2.  The trapezoid sums are regrouped by node:
//   sum_i (W_i+1 + W_i) h_i = sum_j W_j (h_j-1 + h_j),  W_0 = 0.
// FP64 work per (ray, freq, segment): 1 (tau) + 9 (exp) + 1 (ds_i + ds_i+1) + 3 (weights) = 14 instructions.
#ifndef RB_RT_CHUNK
#define RB_RT_CHUNK 32
#endif
constexpr int kChunk = RB_RT_CHUNK;          // segments per staged tile (multiple of 16)
static_assert(kChunk % 16 == 0 && kChunk >= 16, "a chunk is a whole number of 4 KB copy rounds");
constexpr int kStages = 2;
constexpr int kTileDs = (kChunk + 1) * 32;   // doubles per ds tile
constexpr int kTilePp = kChunk * 8;          // double4 per operand tile
constexpr size_t kRaysSmemBytes = kStages * (kTileDs * sizeof(double) + kTilePp * sizeof(double4));

#ifndef RB_RT_GROUP
#define RB_RT_GROUP 4
#endif
#ifndef RB_RT_CTAS
#define RB_RT_CTAS 4
#endif
constexpr int kGroup = RB_RT_GROUP;          // segments per group = independent exp chains per thread
__global__ void __launch_bounds__(256, RB_RT_CTAS) rt_integrate_rays_kernel(const __grid_constant__ RtK k) {
  // dynamic shared memory: [ table | ds tiles x kStages | operand tiles x kStages ]
  extern __shared__ __align__(16) unsigned char s_raw[];
  double* const s_tab = reinterpret_cast<double*>(s_raw);                    // 2^(j/N), copied from k.exp_tab with chunk 0
  double* const s_ds = s_tab + kExpTabDoubles;
  double4* const s_pp = reinterpret_cast<double4*>(s_ds + kStages * kTileDs);
  const int tid = threadIdx.y * 32 + threadIdx.x;

  const int S = k.L - 1;
  // ray tile of this CTA (see RayTile)
  const RayTile rt_ = map_ray_tile(k);
  if (rt_.dead) return;
  const unsigned tile = rt_.tile;
  const long long tpos = rt_.t;       // where this ray's ds / nseg / nanflag live
  const long long r = rt_.r;          // ray index (output row)
  const int f = rt_.fg * 8 + threadIdx.y;
  const bool valid = rt_.in && (f < k.F);
  const int n = valid ? k.nseg[tpos] : -1;
  const bool nanray = valid && k.nanflag[tpos] != 0;
  const int steps = (valid && !nanray) ? n - 1 : 0;        // brightness.py:65: i = 0 .. len(ds)-2

  // Per-thread copy plan for one chunk, computed once: the chunk's 33 ds rows and 32 operand rows are each
  // one contiguous piece of global memory (8448 B / 8192 B), cut into 16-byte cp.async pieces: two per
  // thread for each stream plus a third ds piece for the first 16 threads.  No bounds checks: rows past the
  // end of a tile are never consumed and both slabs are allocated with kRtSlackBytes of slack.
  const char* src_ds = reinterpret_cast<const char*>(k.ds + (size_t)tile * S * 32) + tid * 16;
  const char* src_pp = reinterpret_cast<const char*>(k.prep + (size_t)rt_.fg * S * 8) + tid * 16;
  const unsigned dst_ds = (unsigned)__cvta_generic_to_shared(s_ds) + tid * 16;
  const unsigned dst_pp = (unsigned)__cvta_generic_to_shared(s_pp) + tid * 16;
  auto issue = [&](int c) {
    const unsigned bd = (c & 1) ? (unsigned)(kTileDs * sizeof(double)) : 0u;
    const unsigned bp = (c & 1) ? (unsigned)(kTilePp * sizeof(double4)) : 0u;
    // kChunk / 16 rounds of 4 KB (256 threads x 16 B) per stream, then the extra ds row
    cp_rounds<kChunk / 16>(dst_ds + bd, src_ds);
    cp_rounds<kChunk / 16>(dst_pp + bp, src_pp);
    if (tid < 16) cp_async16_at<(kChunk / 16) * 4096>(dst_ds + bd, src_ds);
    cp_async_commit();
    src_ds += kChunk * 32 * sizeof(double);
    src_pp += kChunk * 8 * sizeof(double4);
  };
  // a CTA with nothing to integrate (off the planet, NaN rays) skips the pipeline altogether
  bool live = steps > 0;
  const bool any_live = __syncthreads_or(live);
  if (any_live) {
    // the exponential table rides in the first copy group
    for (int q = tid; q < kExpTabDoubles / 2; q += 256) cp_async16(s_tab + 2 * q, k.exp_tab + 2 * q);
    issue(0);
  }

  // cA and c2 each meet another constant in one FMA (a DFMA takes a single uniform-register operand), so
  // they are loaded through a thread-dependent (always zero) offset to keep them in vector registers
  const int vz = threadIdx.x >> 5;   // blockDim.x == 32
  const double cA = pin(c_expc + 0 + vz), cM = pin(c_expc + 1), cL = pin(c_expc + 2), c3 = pin(c_expc + 3),
               c2 = pin(c_expc + 4 + vz), c1 = pin(c_expc + 5), ce = pin(c_expc + 6);
  // 32-bit shared-window address of the table, computed once (ptxas otherwise re-derives the CTA's shared
  // window base with S2UR / UMOV / ULEA in every iteration)
  unsigned tab_base;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(tab_base) : "l"(s_tab));
  tab_base += (threadIdx.x & (kExpRep - 1)) * 8u;            // this lane's copy of the table (see kExpRep)
  // One integer compare on the high word of tau (a non-negative double: it orders like its high word; a NaN
  // compares as "beyond") ends the ray once tau >= tau_cut -- the same rule in every integration kernel of this
  // file (tau_cut to the 20 mantissa bits of its high word: exact for 5, 50, ...).  tau_cut is capped at 707
  // (beyond it 2^k leaves the normal range and exp(-tau) is 0 for every purpose), so the hot loop needs neither
  // an underflow select nor an FP64 compare.  The step that crosses the threshold is finished on a cold path.
  const int cut_hi = __double2hiint(fmin(k.tau_cut, 707.0));
  double tau = 0.0, iW = 0.0, Tb = 0.0;
  int i = 0;
  double last_dd = 0.0, last_qy = 0.0, last_qz = 0.0;    // operands of the crossing step

  // one segment: consumes ds_i = dcur, ds_i+1 = dnxt and the operands q of segment i
#define RB_RT_STEP(dcur, dnxt, q)                                                                              \
  {                                                                                                            \
    tau = fma((q).x, (dcur), tau);                                                                             \
    double nd = fma(tau, cA, cM);                                                                              \
    const int ni = __double2loint(nd);                                                                         \
    nd -= cM;                                                                                                  \
    if (__double2hiint(tau) >= cut_hi) {                                                                       \
      last_dd = (dcur) + (dnxt); last_qy = (q).y; last_qz = (q).z;                                             \
      stop = true;                                                                                             \
      break;                                                                                                   \
    }                                                                                                          \
    const double rr = fma(nd, cL, -tau);                                                                       \
    const double p = RB_EXP_POLY(rr);                                                                          \
    double tj;                                                                                                 \
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(tab_base + (ni & (kExpTab - 1)) * (8 * kExpRep)));                  \
    const double v = p * tj;                                                                                   \
    const double e = __hiloint2double(__double2hiint(v) + ((ni << (20 - kExpTabLog)) & 0xFFF00000), __double2loint(v)); \
    const double w = e * ((dcur) + (dnxt));                                                                    \
    iW = fma((q).y, w, iW);                                                                                    \
    Tb = fma((q).z, w, Tb);                                                                                    \
  }

  // e^-tau * dd for a step known to be below the threshold (no checks): independent of the other steps of
  // a group, so ptxas interleaves the four chains (ILP 4 instead of one ~100-cycle dependent chain per step)
#define RB_RT_WEIGHT(tauv, dd, w)                                                                              \
  {                                                                                                            \
    double nd_ = fma((tauv), cA, cM);                                                                          \
    const int ni_ = __double2loint(nd_);                                                                       \
    nd_ -= cM;                                                                                                 \
    const double rr_ = fma(nd_, cL, -(tauv));                                                                  \
    double p_ = RB_EXP_POLY(rr_);                                                                              \
    double tj_;                                                                                                \
    unsigned ta_, ex_;                                                                                         \
    asm("{ .reg .b32 t; and.b32 t, %2, %4; mad.lo.u32 %0, t, %6, %3; and.b32 %1, %2, %5; }"                     \
        : "=r"(ta_), "=r"(ex_) : "r"(ni_), "r"(tab_base), "n"(kExpTab - 1), "n"(~(kExpTab - 1)), "n"(8 * kExpRep)); \
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj_) : "r"(ta_));                                                     \
    const double v_ = p_ * tj_;                                                                                \
    int hi_;                                                                                                   \
    asm("mad.lo.s32 %0, %1, %3, %2;" : "=r"(hi_) : "r"(ex_), "r"(__double2hiint(v_)), "n"(1 << (20 - kExpTabLog))); \
    const double e_ = __hiloint2double(hi_, __double2loint(v_));                                               \
    w = e_ * (dd);                                                                                             \
  }

  bool stop = false;
  bool small = true;                                       // this ray is still in the small-tau phase A
  int ia = 0;                                              // segments integrated in phase A
  for (int c = 0; any_live; ++c) {
    cp_async_wait<0>();                                    // this thread's pieces of chunk c have landed
    if (!__syncthreads_or(live)) break;                    // ... everybody's have; chunk c-1 is fully consumed
    issue(c + 1);                                          // refill the buffer chunk c-1 used
    if (live) {
      double* dsb = s_ds + (c % kStages) * kTileDs + threadIdx.x;
      const double4* ppb = s_pp + (c % kStages) * kTilePp + threadIdx.y;
      // nothing lies below the last node: make ds_steps read as 0 in this tile (the 8 warps of the ray all
      // store the same 0 before their own read of it)
      const int zrow = steps - c * kChunk;
      if (zrow <= kChunk) dsb[zrow * 32] = 0.0;
      const int m = min(kChunk, steps - i);
      int u = 0;
      const double* dp = dsb;
      const double4* qp = ppb;
      // Phase A (per ray, while tau < 2^-13; tau only grows): groups of kGroup segments with the short
      // polynomial -- no range reduction, no table, no tau_cut test.  The group that crosses the bound is
      // left to phase B.  The three coefficients are re-read per chunk so that they are not live in phase B.
      if (RB_EXP_SMALL_LOG > 0 && small) {
        const double sa0 = pin(c_small + 0), sa1 = pin(c_small + 1 + vz), sa2 = pin(c_small + 2);
        constexpr int small_hi = (int)(((0x3FFull - RB_EXP_SMALL_LOG) << 20));   // high word of 2^-RB_EXP_SMALL_LOG
#pragma unroll 1
        for (; u + kGroup <= m; u += kGroup, dp += kGroup * 32, qp += kGroup * 8) {
          double d[kGroup + 1], t[kGroup];
          double4 q[kGroup];
#pragma unroll
          for (int j = 0; j <= kGroup; ++j) d[j] = dp[j * 32];
#pragma unroll
          for (int j = 0; j < kGroup; ++j) q[j] = qp[j * 8];
          t[0] = fma(q[0].x, d[0], tau);
#pragma unroll
          for (int j = 1; j < kGroup; ++j) t[j] = fma(q[j].x, d[j], t[j - 1]);
          if (__double2hiint(t[kGroup - 1]) >= small_hi) { small = false; break; }
#pragma unroll
          for (int j = 0; j < kGroup; ++j) {
            const double w = fma(fma(t[j], sa2, sa1), t[j], sa0) * (d[j] + d[j + 1]);
            iW = fma(q[j].y, w, iW);
            Tb = fma(q[j].z, w, Tb);
          }
          tau = t[kGroup - 1];
        }
        ia += u;                                             // phase-A steps of this chunk (feeds rb_count_steps)
      }
      // Phase B: groups of kGroup segments: the optical depths first (one dependent FMA each), one threshold
      // test on the deepest, then kGroup independent exp / accumulate chains through the table
      if (!(RB_EXP_SMALL_LOG > 0) || !small) {
#pragma unroll 1
        for (; u + kGroup <= m; u += kGroup, dp += kGroup * 32, qp += kGroup * 8) {
          double d[kGroup + 1], t[kGroup], w[kGroup];
          double4 q[kGroup];
#pragma unroll
          for (int j = 0; j <= kGroup; ++j) d[j] = dp[j * 32];
#pragma unroll
          for (int j = 0; j < kGroup; ++j) q[j] = qp[j * 8];
          t[0] = fma(q[0].x, d[0], tau);
#pragma unroll
          for (int j = 1; j < kGroup; ++j) t[j] = fma(q[j].x, d[j], t[j - 1]);
          if (__double2hiint(t[kGroup - 1]) >= cut_hi) break;   // tau_cut is crossed inside this group: go step by step
#pragma unroll
          for (int j = 0; j < kGroup; ++j) RB_RT_WEIGHT(t[j], d[j] + d[j + 1], w[j]);
#pragma unroll
          for (int j = 0; j < kGroup; ++j) { iW = fma(q[j].y, w[j], iW); Tb = fma(q[j].z, w[j], Tb); }
          tau = t[kGroup - 1];
        }
      }
      // remainder of the chunk / the group that crosses tau_cut: one segment at a time with the test
      for (; u < m; ++u) {
        const double da = dsb[u * 32], db = dsb[(u + 1) * 32];
        const double4 q0 = ppb[u * 8];
        RB_RT_STEP(da, db, q0);
      }
      i += stop ? (u + 1) : m;                               // segments integrated so far (exact: feeds rb_count_steps)
      live = !stop && i < steps;
    }
  }
#undef RB_RT_WEIGHT
#undef RB_RT_STEP
  if (stop) {
    // finish the step that crossed tau_cut: full exp(-tau) with the underflow guard
    const double tc = fmin(tau, 800.0);
    double nd = fma(tc, cA, cM);
    const int ni = __double2loint(nd);
    nd -= cM;
    const double rr = fma(nd, cL, -tc);
    const double p = RB_EXP_POLY(rr);
    const double v = p * s_tab[(ni & (kExpTab - 1)) * kExpRep];
    double e = __hiloint2double(__double2hiint(v) + ((ni << (20 - kExpTabLog)) & 0xFFF00000), __double2loint(v));
    if ((unsigned)__double2hiint(nd) > (unsigned)__double2hiint(-1022.0 * kExpTab)) e = 0.0;
    const double w = e * last_dd;
    iW = fma(last_qy, w, iW);
    Tb = fma(last_qz, w, Tb);
  }
  cp_async_wait<0>();
  if (k.step_counter) {   // measurement aid (bench.py): executed segment-steps, one atomic per warp
    unsigned long long done = (unsigned long long)i;   // rays that crossed tau_cut count the whole last group
    unsigned long long done_a = (unsigned long long)ia;
    for (int o = 16; o > 0; o >>= 1) {
      done += __shfl_down_sync(0xffffffffu, done, o);
      done_a += __shfl_down_sync(0xffffffffu, done_a, o);
    }
    if (threadIdx.x == 0) { atomicAdd(k.step_counter, done); atomicAdd(k.step_counter + 1, done_a); }
  }
  if (valid) {
    double vout, wout = iW;
    if (n < 0) vout = kTcmb;                                 // off planet (brightness.py:46-51)
    else if (nanray || tau != tau) vout = wout = nan("");    // NaN segment below the tangent shell / NaN alpha
    else vout = (Tb < kTcmb) ? kTcmb : Tb / iW;              // brightness.py:109-113
    const size_t o = (size_t)r * k.F + f;
    if (k.out_f32) reinterpret_cast<float*>(k.out_Tb)[o] = (float)vout;
    else reinterpret_cast<double*>(k.out_Tb)[o] = vout;
    if (k.out_intW) k.out_intW[o] = (n < 0) ? 0.0 : wout;
  }
  progress_report(k, rt_, tid);
}

// ====================================================================================================
// Two frequencies per thread (F >= 16): rt_integrate_pairs_kernel.
//
// Same arithmetic, operation for operation, as rt_integrate_rays_kernel (results are bit-identical), other
// decomposition: a thread owns one ray and TWO adjacent frequencies, a CTA 32 rays x 16 frequencies.  What that
// buys (the one-frequency kernel ends at 59 % of its issue slots with the shared-memory data pipe as its
// busiest unit, 0.81 wavefronts per clock -- profiles/r1_final_rt_integrate_rays_ncu_full.txt):
//   * every ds value read from shared memory (LDS.64, two wavefronts per warp) feeds two chains, the carried
//     ds_i+1 of a group is kept in a register, and ds_i + ds_i+1 is added once for both frequencies;
//   * the operands of a pair are one 48-byte row {asum_a, asum_b | a'_a, T a'_a | a'_b, T a'_b}: three LDS.128
//     per segment for two steps instead of four reads;
//   * half as many CTAs per ray tile: half the per-chunk work (copy plan, wait, barrier vote) per step and half
//     the L2 -> shared-memory traffic of the ds tiles.
// A thread runs its two frequencies together while both are live: phase A while both optical depths are below
// 2^-RB_EXP_SMALL_LOG, phase B until one of them crosses tau_cut inside a group; the crossing is resolved one
// segment at a time, and the frequency that is left alone finishes on the single-segment path (adjacent
// frequencies stop a few layers apart).  A frequency beyond F (odd F) is a ghost: zero operands, ends with its
// partner.
// CTAs per SM the pair kernel is compiled for: 3 x 256 threads leave 80 registers per thread (4 x 256: 64 registers
// and spills in the loops -- measured 4.07 against 3.74 ms on C4)
#ifndef RB_RTP_CTAS
#define RB_RTP_CTAS (3 * 8 / RB_RT_WARPS)
#endif
// RB_RT_RING = 1: the operand streams of the pair kernel are staged by bulk copies (cp.async.bulk, the TMA engine)
// into a ring of kPStages buffers guarded by mbarriers instead of cp.async + one CTA barrier per chunk: whichever warp
// finds the next buffer free issues the copy of the next chunk (one thread, two instructions), and a warp that is ahead
// of the others -- its frequencies are still in the cheap small-tau phase, or already finished -- keeps going for up to
// kPStages - 2 chunks instead of waiting at the barrier (see the ring protocol in the kernel).
// Measured on B200 (C4, profiles/r2_ab_ring.jsonl), same bits as the barrier version: 3 x 16 segments 3.96 ms,
// 4 x 16 (5 CTAs/SM) 3.97, 4 x 8 4.78, 3 x 32 (5 CTAs/SM) 3.80 against 3.49 ms with cp.async + one barrier per 32
// segments -- a warp parked at a CTA barrier costs nothing, a warp polling an mbarrier takes issue slots from the warps
// that have work, and the kernel is short of exactly those.  Kept as a build option, off by default.
#ifndef RB_RT_RING
#define RB_RT_RING 0
#endif
#ifndef RB_RTP_CHUNK
#define RB_RTP_CHUNK (RB_RT_RING ? 16 : RB_RT_CHUNK)
#endif
#ifndef RB_RTP_STAGES
#define RB_RTP_STAGES (RB_RT_RING ? 4 : 2)
#endif
constexpr int kPChunk = RB_RTP_CHUNK;                // segments per staged tile of the pair kernel
constexpr int kPStages = RB_RTP_STAGES;
constexpr int kPTileDs = (kPChunk + 1) * 32;         // doubles per ds tile (one extra row: ds_i+1 of the last segment)
static_assert(kPChunk % 4 == 0 && kPChunk >= 8, "pair kernel: trips of four segments");
static_assert(RB_RT_RING ? kPStages >= 3 : kPStages == 2, "ring: at least three buffers; barrier version: two");
constexpr int kPairRow = 3;                          // double2 per (segment, pair)
constexpr int kPairThreads = 32 * kPairWarps;
// Decomposition "tiles" (rt_integrate_pairs_kernel<true>): the warps of a CTA share the FREQUENCY PAIR and each takes
// its own ray tile (four tiles that are neighbours in the ordered list), staging its own ds chunks; nothing is shared
// but the exponential table, so there is no CTA barrier in the loop and the warps of a CTA -- same frequencies, nearly
// the same rays -- finish together.  The ds tiles are then fetched once per frequency pair instead of once per eight
// frequencies (16 GB instead of 8 GB from L2 per C4 launch: 14 % of the crossbar), in chunks of kTChunk segments.
// Measured on B200 (C4, same bits): 3.36 ms against 3.10 ms for the shared-rays decomposition -- the barrier waits it
// removes cost no issue slots, the extra copy instructions and chunk turns (8 segments instead of 32) do.  Kept as a
// run-time option (RB_RT_TILES=1), off by default.
#ifndef RB_RTT_CHUNK
#define RB_RTT_CHUNK 8
#endif
constexpr int kTChunk = RB_RTT_CHUNK;
static_assert(kTChunk % 4 == 0 && kTChunk >= 8, "tiles decomposition: trips of four segments");
constexpr size_t kTilesWarpBytes = 2 * ((kTChunk + 1) * 32 * sizeof(double) + kTChunk * 3 * sizeof(double2));
constexpr size_t kTilesSmemBytes = kPairWarps * kTilesWarpBytes;
static_assert(kRtSlackBytes >= (size_t)(kTChunk + 1) * 32 * sizeof(double) + 256, "ds slab slack covers one over-read chunk");
constexpr int kPairTileQ = kPChunk * kPairWarps * kPairRow;    // double2 per operand tile
constexpr int kPairRound = kPairThreads * 16;        // bytes one round of 16-byte copies of the whole CTA moves
constexpr size_t kPairStageBytes = kPTileDs * sizeof(double) + kPairTileQ * sizeof(double2);
constexpr size_t kPairsSmemBytes = kPStages * kPairStageBytes + (RB_RT_RING ? 2 * kPStages * 8 + 16 : 0);
static_assert(RB_RT_RING || (kPChunk * 32 * sizeof(double)) % kPairRound == 0, "ds tile: whole copy rounds");
static_assert(RB_RT_RING || (kPairTileQ * sizeof(double2)) % kPairRound == 0, "pair operand tile: whole copy rounds");
static_assert((kPTileDs * sizeof(double)) % 16 == 0 && (kPairTileQ * sizeof(double2)) % 16 == 0, "bulk copies: 16-byte sizes");
// (the ring issues no chunk beyond the one that holds row S - 1: at most one chunk of over-read in both versions)
static_assert(kRtSlackBytes >= (size_t)(kPChunk + 1) * 32 * sizeof(double) + 256, "ds slab slack covers one over-read chunk");
static_assert(kRtSlackBytes >= (size_t)kPChunk * 8 * kPairRow * sizeof(double2), "pair operand slack covers one over-read chunk");
static_assert(kRtSlackBytes >= kChunk * 8 * sizeof(double4), "operand slack covers one over-read chunk");

// pair operands:  prep2[fg][i][p] = { asum_a, asum_b }, { a'_a, T_i+1 a'_a }, { a'_b, T_i+1 a'_b },
//   a = kPairFreqs fg + 2 p, b = a + 1; asum = (a_i + a_i+1) kHalfCm, a' = a_i+1 kHalfCm; zero for f >= F
__global__ void rt_prepare_pairs_kernel(const double* __restrict__ alpha, const double* __restrict__ T, int L, int F,
                                        int ngroups, int pw /* pairs per row */, double2* __restrict__ prep2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int Lm1 = L - 1;
  if (idx >= ngroups * Lm1 * pw) return;
  const int p = idx % pw;
  const int i = (idx / pw) % Lm1;
  const int fg = (idx / pw) / Lm1;
  const double kHalfCm = 0.5 * kKmToCm;
  double v[2][3];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = fg * 2 * pw + 2 * p + h;
    v[h][0] = v[h][1] = v[h][2] = 0.0;
    if (f < F) {
      const double a0 = alpha[(size_t)i * F + f], a1 = alpha[(size_t)(i + 1) * F + f];
      v[h][0] = (a0 + a1) * kHalfCm; v[h][1] = a1 * kHalfCm; v[h][2] = (T[i + 1] * a1) * kHalfCm;
    }
  }
  double2* o = prep2 + (size_t)idx * kPairRow;
  o[0] = make_double2(v[0][0], v[1][0]);
  o[1] = make_double2(v[0][1], v[0][2]);
  o[2] = make_double2(v[1][1], v[1][2]);
}

// shared-memory reads at a 32-bit shared-window address plus a compile-time offset (one LDS each, no address
// arithmetic in the loops)
template <unsigned OFF>
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
template <unsigned OFF>
__device__ __forceinline__ double2 lds_v2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF));
  return v;
}

#if RB_RT_RING
// ---- mbarrier / bulk-copy primitives of the ring (PTX; shared-window addresses) ---------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {     // may block for a bounded time
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {         // never blocks
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#endif

template <bool TILES, bool STREAM = false>
__global__ void __launch_bounds__(kPairThreads, RB_RTP_CTAS) rt_integrate_pairs_kernel(const __grid_constant__ RtK k) {
  static_assert(!(TILES && RB_RT_RING), "the tiles decomposition stages with cp.async");
  // TILES = false: the CTA's warps are kPairWarps frequency pairs on one ray tile (shared staging buffers);
  // TILES = true:  the CTA's warps are kPairWarps ray tiles on one frequency pair (private staging buffers per warp)
  constexpr int kCh = TILES ? kTChunk : kPChunk;                 // segments per staged chunk
  constexpr int kSt = TILES ? 2 : kPStages;                      // staging buffers
  constexpr int kRowPairs = TILES ? 1 : kPairWarps;              // frequency pairs per operand row
  constexpr int kTileD = (kCh + 1) * 32;                         // doubles per ds buffer
  constexpr int kTileQ = kCh * kRowPairs * kPairRow;             // double2 per operand buffer
  // dynamic shared memory: [ table | ds buffers x kSt | operand buffers x kSt ]  (TILES: the last two per warp)
  extern __shared__ __align__(16) unsigned char s_raw[];
  double* const s_tab = reinterpret_cast<double*>(s_raw);
  double* const s_ds = s_tab + kExpTabDoubles + (TILES ? threadIdx.y * (kSt * (kTileD + 2 * kTileQ)) : 0);
  double2* const s_q = reinterpret_cast<double2*>(s_ds + kSt * kTileD);
  const int tid = threadIdx.y * 32 + threadIdx.x;

  const int S = k.L - 1;
  const RayTile rt_ = TILES ? map_ray_tile(k, kPairWarps, threadIdx.y) : map_ray_tile(k);
  if (!TILES && rt_.dead) return;
  const unsigned tile = rt_.tile;
  const long long tpos = rt_.t;
  const long long r = rt_.r;
  const int fA = rt_.fg * (2 * kRowPairs) + 2 * (TILES ? 0 : (int)threadIdx.y);
  const bool validA = rt_.in && (fA < k.F), validB = rt_.in && (fA + 1 < k.F);
  // Streamed geometry (GeoK::prog): the trace of this tile may still be running.  Segment count and NaN flag are read
  // at the end; until then every ray walks all S - 1 steps (the trace writes zeros from the last used segment on)
  // and chunk c of the tile is copied only after all its rays have counted themselves past it.
  constexpr bool streamed = !TILES && STREAM;
  int n = validA ? (streamed ? S : k.nseg[tpos]) : -1;
  bool nanray = validA && !streamed && k.nanflag[tpos] != 0;
  const int steps = (validA && !nanray) ? n - 1 : 0;        // brightness.py:65: i = 0 .. len(ds)-2
  bool geo_done = false;                                     // (thread 0) the trace of this tile has ended
  auto geo_wait = [&](int c) {                               // one thread: chunk c of this tile is complete
    if (geo_done || c >= k.geo_npub) return;
    const int geo_lanes = min(32, *k.ncomp - (int)tile * 32);
    const int* base = k.geo_prog + (size_t)tile * k.geo_npub;
    int v;
    // every ray counts into the last chunk when its trace ends: once that counter is full nothing is left to wait
    // for (the usual case for all but the first CTAs of a launch)
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(base + k.geo_npub - 1) : "memory");
    if (v >= geo_lanes) { geo_done = true; return; }
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(base + c) : "memory");
      if (v >= geo_lanes) break;
      if (++spins > (1u << 26)) __trap();                    // a trace that never comes: fail, do not hang
    }
  };

  // bit 0: frequency a is live, bit 1: frequency b (a ghost b beyond F rides along with zero operands)
  int mode = (steps > 0) ? 3 : 0;
#if RB_RT_RING
  // ---- ring protocol ------------------------------------------------------------------------------------------
  // Chunk n of the tile (kPChunk segments: one contiguous piece of the ds tile and one of the operand slab) lives in
  // buffer n % kPStages.  full[b] completes when the two bulk copies of the chunk have landed (transaction bytes);
  // empty[b] when all warps of the CTA have released the chunk (one arrival per warp).  After releasing a chunk, lane 0
  // of a warp issues every following chunk whose buffer is free (s_issued counts them; an atomic claims a chunk), so
  // the copies run up to kPStages - 1 chunks ahead of the slowest warp and the fastest warp at most kPStages - 2
  // chunks ahead of it.  A warp whose rays are all finished keeps releasing chunks without work until s_live -- the
  // number of warps with unfinished rays -- reaches 0.
  unsigned long long* const s_bar = reinterpret_cast<unsigned long long*>(s_q + kPStages * kPairTileQ);
  int* const s_ctl = reinterpret_cast<int*>(s_bar + 2 * kPStages);     // [0] live warps, [1] chunks issued so far
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(s_bar);
  auto bar_full = [&](int b) { return bar0 + 8u * (unsigned)b; };
  auto bar_empty = [&](int b) { return bar0 + 8u * (unsigned)(kPStages + b); };
  const char* const g_ds = reinterpret_cast<const char*>(k.ds + (size_t)tile * S * 32);
  const char* const g_q = reinterpret_cast<const char*>(k.prep2 + (size_t)rt_.fg * S * kPairWarps * kPairRow);
  const unsigned sh_ds = (unsigned)__cvta_generic_to_shared(s_ds), sh_q = (unsigned)__cvta_generic_to_shared(s_q);
  constexpr unsigned kDsBytes = kPTileDs * sizeof(double), kQBytes = kPairTileQ * sizeof(double2);
  const int last_chunk = (S - 1) / kPChunk;                  // no ray uses rows beyond S - 1
  auto issue = [&](int n) {                                  // one thread
    const int b = n % kPStages;
    mbar_expect_tx(bar_full(b), kDsBytes + kQBytes);
    bulk_g2s(sh_ds + (unsigned)b * kDsBytes, g_ds + (size_t)n * (kPChunk * 32 * sizeof(double)), kDsBytes, bar_full(b));
    bulk_g2s(sh_q + (unsigned)b * kQBytes, g_q + (size_t)n * kQBytes, kQBytes, bar_full(b));
  };
  if (tid == 0) {
    for (int b = 0; b < kPStages; ++b) { mbar_init(bar_full(b), 1); mbar_init(bar_empty(b), kPairWarps); }
    s_ctl[0] = 0; s_ctl[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  bool warp_live = __any_sync(0xffffffffu, mode != 0);
  if (threadIdx.x == 0 && warp_live) atomicAdd(&s_ctl[0], 1);
  __syncthreads();
  const bool any_live = s_ctl[0] > 0;
  if (any_live) {
    for (int q = tid; q < kExpTabDoubles / 2; q += kPairThreads) cp_async16(s_tab + 2 * q, k.exp_tab + 2 * q);
    cp_async_commit();
    if (tid == 0) {
      const int n0 = min(kPStages - 1, last_chunk + 1);      // the first chunks: every buffer but one
      for (int n = 0; n < n0; ++n) issue(n);
      s_ctl[1] = n0;
    }
    cp_async_wait<0>();
  }
  __syncthreads();                                           // table landed, s_ctl[1] published
#else
  // copy plan (see rt_integrate_rays_kernel): a ds chunk is (kCh + 1) x 256 B, an operand chunk kCh x kRowPairs x 48 B,
  // both contiguous in global memory, moved in 16-byte pieces by the threads that share the buffers (TILES: the warp)
  constexpr int kCopyThreads = TILES ? 32 : kPairThreads;
  const int ctid = TILES ? (int)threadIdx.x : tid;
  const char* src_ds = reinterpret_cast<const char*>(k.ds + (size_t)tile * S * 32) + ctid * 16;
  const char* src_q = reinterpret_cast<const char*>(k.prep2 + (size_t)rt_.fg * S * kRowPairs * kPairRow) + ctid * 16;
  const unsigned dst_ds = (unsigned)__cvta_generic_to_shared(s_ds) + ctid * 16;
  const unsigned dst_q = (unsigned)__cvta_generic_to_shared(s_q) + ctid * 16;
  auto issue = [&](int c) {
    const unsigned bd = (c & 1) ? (unsigned)(kTileD * sizeof(double)) : 0u;
    const unsigned bq = (c & 1) ? (unsigned)(kTileQ * sizeof(double2)) : 0u;
    constexpr int kRound = kCopyThreads * 16;
    constexpr int kDsBytes = kTileD * (int)sizeof(double), kQBytes = kTileQ * (int)sizeof(double2);
    cp_rounds<kDsBytes / kRound, kRound>(dst_ds + bd, src_ds);
    if (ctid * 16 < kDsBytes % kRound) cp_async16_at<(kDsBytes / kRound) * kRound>(dst_ds + bd, src_ds);
    cp_rounds<kQBytes / kRound, kRound>(dst_q + bq, src_q);
    if (ctid * 16 < kQBytes % kRound) cp_async16_at<(kQBytes / kRound) * kRound>(dst_q + bq, src_q);
    cp_async_commit();
    src_ds += kCh * 32 * sizeof(double);
    src_q += kTileQ * sizeof(double2);
  };
  bool any_live;
  if (TILES) {
    // the table is the one thing the warps share: all threads fetch it, then every warp is on its own
    for (int q = tid; q < kExpTabDoubles / 2; q += kPairThreads) cp_async16(s_tab + 2 * q, k.exp_tab + 2 * q);
    any_live = __any_sync(0xffffffffu, mode != 0);
    if (any_live) issue(0); else cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (!any_live) {                                           // nothing to integrate in this warp's tile
      mode = 0;
    }
  } else {
    if (streamed && tid == 0) geo_wait(0);
    any_live = __syncthreads_or(mode != 0);
    if (any_live) {
      for (int q = tid; q < kExpTabDoubles / 2; q += kPairThreads) cp_async16(s_tab + 2 * q, k.exp_tab + 2 * q);
      issue(0);
    }
  }
#endif

  const int vz = threadIdx.x >> 5;   // blockDim.x == 32: always zero, unknown to ptxas (see pin())
  // cM = 2^52 + 2^51, 1 and 1/2 have an all-zero low word: they are encoded in the instruction (a DFMA / DADD takes
  // a 32-bit immediate for the high word) and cost no register; the full-mantissa constants are pinned (see pin())
  const double cA = pin(c_expc + 0 + vz), cL = pin(c_expc + 2), ce = pin(c_expc + 6);
  constexpr double cM = 6755399441055744.0;
#if RB_EXP_DEG == 2
  constexpr double c2 = 0.5, c1 = 1.0;
#else
  const double c3 = pin(c_expc + 3), c2 = pin(c_expc + 4 + vz);
  constexpr double c1 = 1.0;
#endif
  unsigned tab_base;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(tab_base) : "l"(s_tab));
  tab_base += (threadIdx.x & (kExpRep - 1)) * 8u;            // this lane's copy of the table (see kExpRep)
  const int cut_hi = __double2hiint(fmin(k.tau_cut, 707.0));   // tau >= tau_cut <=> hi(tau) >= cut_hi (see above)
  double tauA = 0.0, iWA = 0.0, TbA = 0.0, tauB = 0.0, iWB = 0.0, TbB = 0.0;
  int i = 0;                 // segments consumed so far (both frequencies walk together)
  int nf = 0, ns = 0;        // rb_count_steps: segments taken by the pair loops (two steps each) / single steps
  int ia = 0;                // steps in phase A (both frequencies counted)
  bool small = true;
  // the small-tau phase accumulates its weights without the factor E/2 of the polynomial (see c_small2): applied once,
  // when the ray leaves the phase (or ends inside it)
  auto end_small = [&]() {
    if (RB_EXP_SMALL_LOG > 0 && small) {
      const double a2 = pin(c_small2 + 2);
      iWA *= a2; TbA *= a2; iWB *= a2; TbB *= a2;
    }
    small = false;
  };

  // e^-tau * dd for a step known to be below the threshold (the table path of rt_integrate_rays_kernel);
  // ndraw = fma(tau, cA, cM) (the caller may have it from the tau_cut test)
  auto weight_nd = [&](double tauv, double ndraw, double dd) -> double {
    const int ni_ = __double2loint(ndraw);
    const double nd_ = ndraw - cM;
    const double rr_ = fma(nd_, cL, -tauv);
    const double p_ = RB_EXP_POLY(rr_);
    double tj_;
    unsigned ta_, ex_;
    asm("{ .reg .b32 t; and.b32 t, %2, %4; mad.lo.u32 %0, t, %6, %3; and.b32 %1, %2, %5; }"
        : "=r"(ta_), "=r"(ex_) : "r"(ni_), "r"(tab_base), "n"(kExpTab - 1), "n"(~(kExpTab - 1)), "n"(8 * kExpRep));
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj_) : "r"(ta_));
    const double v_ = p_ * tj_;
    int hi_;
    asm("mad.lo.s32 %0, %1, %3, %2;" : "=r"(hi_) : "r"(ex_), "r"(__double2hiint(v_)), "n"(1 << (20 - kExpTabLog)));
    return __hiloint2double(hi_, __double2loint(v_)) * dd;
  };
  auto weight = [&](double tauv, double dd) -> double { return weight_nd(tauv, fma(tauv, cA, cM), dd); };
  // one step of frequency h (0: a, 1: b) whose optical depth tk is known: the table exponential with the tau_cut
  // test of rt_integrate_rays_kernel's single-step path.  The step that crosses tau_cut is still accumulated (full
  // exp(-tau) with the underflow guard), then the frequency stops.
  auto step_known = [&](int h, double tk, double dd, double2 op) {
    double nd = fma(tk, cA, cM);
    int ni = __double2loint(nd);
    nd -= cM;
    const bool crossed = __double2hiint(tk) >= cut_hi;
    double tc = tk;
    if (crossed) {
      tc = fmin(tk, 800.0);
      nd = fma(tc, cA, cM);
      ni = __double2loint(nd);
      nd -= cM;
    }
    const double rr = fma(nd, cL, -tc);
    const double p = RB_EXP_POLY(rr);
    const double v = p * s_tab[(ni & (kExpTab - 1)) * kExpRep];
    double e = __hiloint2double(__double2hiint(v) + ((ni << (20 - kExpTabLog)) & 0xFFF00000), __double2loint(v));
    if (crossed && (unsigned)__double2hiint(nd) > (unsigned)__double2hiint(-1022.0 * kExpTab)) e = 0.0;
    const double w = e * dd;
    if (h) { iWB = fma(op.x, w, iWB); TbB = fma(op.y, w, TbB); }
    else { iWA = fma(op.x, w, iWA); TbA = fma(op.y, w, TbA); }
    ++ns;
    if (crossed) mode &= ~(1 << h);
  };

  // shared-window addresses of this thread's ds column / operand rows in stage 0
  const unsigned ds_a0 = (unsigned)__cvta_generic_to_shared(s_ds + threadIdx.x);
  const unsigned q_a0 = (unsigned)__cvta_generic_to_shared(s_q + (TILES ? 0 : threadIdx.y) * kPairRow);
  constexpr unsigned kRowB = kRowPairs * kPairRow * sizeof(double2);   // bytes between the operand rows of two segments
  constexpr int small_hi = (int)(((0x3FFull - (RB_EXP_SMALL_LOG > 0 ? RB_EXP_SMALL_LOG : 1)) << 20));

  for (int c = 0; any_live; ++c) {
#if RB_RT_RING
    {
      // wait for chunk c; give up when no warp of the CTA has work left (nobody will issue it any more)
      const unsigned fb = bar_full(c % kPStages), par = (unsigned)(c / kPStages) & 1u;
      bool got = __all_sync(0xffffffffu, mbar_try_wait(fb, par));
      while (!got) {
        if (__any_sync(0xffffffffu, *reinterpret_cast<volatile int*>(&s_ctl[0]) <= 0)) break;
        got = __all_sync(0xffffffffu, mbar_try_wait(fb, par));
      }
      if (!got) break;
    }
    bool wrote_zero = false;
#else
    cp_async_wait<0>();
    if (TILES) {
      // this warp's pieces of chunk c have landed (every lane waited for its own); the vote also tells whether any of
      // its rays is still going, and that all lanes are done with the buffer chunk c + 1 goes into
      __syncwarp();
      if (!__any_sync(0xffffffffu, mode != 0)) break;
    } else {
      if (streamed && tid == 0) geo_wait(c + 1);             // the chunk about to be copied (the others wait at the barrier)
      if (!__syncthreads_or(mode != 0)) break;
    }
    issue(c + 1);
#endif
    if (mode != 0) {
      double* dsb = s_ds + (c % kSt) * kTileD + threadIdx.x;
      const double2* qb = s_q + (c % kSt) * kTileQ + (TILES ? 0 : threadIdx.y) * kPairRow;
      const int zrow = steps - c * kCh;
#if RB_RT_RING
      wrote_zero = zrow <= kCh;
#endif
      if (zrow <= kCh) dsb[zrow * 32] = 0.0;                 // nothing lies below the last node
      const int m = min(kCh, steps - i);
      int u = 0;
      if (mode == 3 && m >= 4) {
        // Trips of 2 x (two segments x two frequencies).  The optical depths are updated in place; a group the
        // fast loops cannot finish (it leaves the small-tau range / a frequency crosses tau_cut inside it) is
        // handed, with its four optical depths, to finish_group, which takes it one step at a time.
        const unsigned dbase = ds_a0 + (unsigned)((c % kSt) * kTileD * sizeof(double));
        unsigned dpa = dbase;
        unsigned qpa = q_a0 + (unsigned)((c % kSt) * kTileQ * sizeof(double2));
        const unsigned dlast = dbase + 256u * (unsigned)(m - 4);   // last trip start with four segments left
        double d0 = lds_f64<0>(dpa), d1, d2, d3, tA0, tB0;
        double2 s0, s1;
        // the group of segments (u, u+1) = rows QO, QO + kRowB of the operand tile, segment lengths D0, D1, D2
#define RB_PAIR_TAUS(D0, D1, QO)                                                        \
        s0 = lds_v2<(QO)>(qpa); s1 = lds_v2<(QO) + kRowB>(qpa);                         \
        tA0 = fma(s0.x, (D0), tauA); tB0 = fma(s0.y, (D0), tauB);                       \
        tauA = fma(s1.x, (D1), tA0); tauB = fma(s1.y, (D1), tB0);
#define RB_PAIR_ACC(QO, WA0, WB0, WA1, WB1)                                             \
        {                                                                               \
          const double2 a0 = lds_v2<(QO) + 16>(qpa), b0 = lds_v2<(QO) + 32>(qpa);       \
          const double2 a1 = lds_v2<(QO) + kRowB + 16>(qpa), b1 = lds_v2<(QO) + kRowB + 32>(qpa); \
          iWA = fma(a0.x, (WA0), iWA); TbA = fma(a0.y, (WA0), TbA);                     \
          iWB = fma(b0.x, (WB0), iWB); TbB = fma(b0.y, (WB0), TbB);                     \
          iWA = fma(a1.x, (WA1), iWA); TbA = fma(a1.y, (WA1), TbA);                     \
          iWB = fma(b1.x, (WB1), iWB); TbB = fma(b1.y, (WB1), TbB);                     \
        }
        auto finish_group = [&](double e0, double e1, double e2, unsigned qa) {
          const double dd0 = e0 + e1, dd1 = e1 + e2;
          const double tA1 = tauA, tB1 = tauB;
          if (mode & 1) { tauA = tA0; step_known(0, tA0, dd0, lds_v2<16>(qa)); }
          if (mode & 2) { tauB = tB0; step_known(1, tB0, dd0, lds_v2<32>(qa)); }
          if (mode & 1) { tauA = tA1; step_known(0, tA1, dd1, lds_v2<kRowB + 16>(qa)); }
          if (mode & 2) { tauB = tB1; step_known(1, tB1, dd1, lds_v2<kRowB + 32>(qa)); }
          nf -= 2;                                           // counted step by step in ns
        };
        if (RB_EXP_SMALL_LOG > 0 && small) {
          const double scb = pin(c_small2 + 0 + vz), scd = pin(c_small2 + 1);
          auto wa = [&](double t, double dd) { const double u_ = t + scb; return fma(u_, u_, scd) * dd; };
#define RB_PAIR_WA(T, DD) wa((T), (DD))
#pragma unroll 1
          while (dpa <= dlast) {
            d1 = lds_f64<256>(dpa); d2 = lds_f64<512>(dpa);
            RB_PAIR_TAUS(d0, d1, 0)
            if (max(__double2hiint(tauA), __double2hiint(tauB)) >= small_hi) goto small_exit_0;
            {
              const double dd0 = d0 + d1, dd1 = d1 + d2;
              RB_PAIR_ACC(0, RB_PAIR_WA(tA0, dd0), RB_PAIR_WA(tB0, dd0), RB_PAIR_WA(tauA, dd1), RB_PAIR_WA(tauB, dd1))
            }
            d3 = lds_f64<768>(dpa); d0 = lds_f64<1024>(dpa);
            RB_PAIR_TAUS(d2, d3, 2 * kRowB)
            if (max(__double2hiint(tauA), __double2hiint(tauB)) >= small_hi) goto small_exit_1;
            {
              const double dd0 = d2 + d3, dd1 = d3 + d0;
              RB_PAIR_ACC(2 * kRowB, RB_PAIR_WA(tA0, dd0), RB_PAIR_WA(tB0, dd0), RB_PAIR_WA(tauA, dd1), RB_PAIR_WA(tauB, dd1))
            }
            dpa += 1024; qpa += 4 * kRowB;
          }
          ia += 2 * (int)((dpa - dbase) >> 8);
          goto small_done;
        small_exit_1:
          { const double t0 = d0; d0 = d2; d1 = d3; d2 = t0; }
          dpa += 512; qpa += 2 * kRowB;
        small_exit_0:
          ia += 2 * (int)((dpa - dbase) >> 8);
          end_small();
          finish_group(d0, d1, d2, qpa);
          d0 = d2; dpa += 512; qpa += 2 * kRowB;
        small_done:;
#undef RB_PAIR_WA
        }
        if ((!(RB_EXP_SMALL_LOG > 0) || !small) && mode == 3) {
#pragma unroll 1
          while (dpa <= dlast) {
            d1 = lds_f64<256>(dpa); d2 = lds_f64<512>(dpa);
            RB_PAIR_TAUS(d0, d1, 0)
            // a frequency crosses tau_cut inside this group?
            if (max(__double2hiint(tauA), __double2hiint(tauB)) >= cut_hi) goto cut_exit_0;
            {
              const double dd0 = d0 + d1, dd1 = d1 + d2;
              const double wA0 = weight(tA0, dd0), wB0 = weight(tB0, dd0);
              const double wA1 = weight(tauA, dd1), wB1 = weight(tauB, dd1);
              RB_PAIR_ACC(0, wA0, wB0, wA1, wB1)
            }
            d3 = lds_f64<768>(dpa); d0 = lds_f64<1024>(dpa);
            RB_PAIR_TAUS(d2, d3, 2 * kRowB)
            if (max(__double2hiint(tauA), __double2hiint(tauB)) >= cut_hi) goto cut_exit_1;
            {
              const double dd0 = d2 + d3, dd1 = d3 + d0;
              const double wA0 = weight(tA0, dd0), wB0 = weight(tB0, dd0);
              const double wA1 = weight(tauA, dd1), wB1 = weight(tauB, dd1);
              RB_PAIR_ACC(2 * kRowB, wA0, wB0, wA1, wB1)
            }
            dpa += 1024; qpa += 4 * kRowB;
          }
          goto cut_done;
        cut_exit_1:
          { const double t0 = d0; d0 = d2; d1 = d3; d2 = t0; }
          dpa += 512; qpa += 2 * kRowB;
        cut_exit_0:
          finish_group(d0, d1, d2, qpa);
          d0 = d2; dpa += 512; qpa += 2 * kRowB;
        cut_done:;
        }
#undef RB_PAIR_TAUS
#undef RB_PAIR_ACC
        u = (int)((dpa - dbase) >> 8);
        nf += u;
      }
      // one segment at a time: the chunk remainder, and the frequency that is left alone after its partner stopped
      // (these steps take the table exponential: a ray still in the small-tau phase leaves it here)
      if (u < m && mode != 0) end_small();
#pragma unroll 1
      for (; u < m && mode != 0; ++u) {
        const double dcur = dsb[u * 32];
        const double dd = dcur + dsb[(u + 1) * 32];
        const double2* qr = qb + (size_t)u * kRowPairs * kPairRow;
        const double2 s0 = qr[0];
        if (mode & 1) { tauA = fma(s0.x, dcur, tauA); step_known(0, tauA, dd, qr[1]); }
        if (mode & 2) { tauB = fma(s0.y, dcur, tauB); step_known(1, tauB, dd, qr[2]); }
        // a ghost frequency (beyond F: zero operands) never crosses: it ends with its partner
        if (mode == 2 && !validB) mode = 0;
      }
      i += u;
      if (i >= steps || (mode == 2 && !validB)) mode = 0;
    }
#if RB_RT_RING
    {
      // release chunk c; the zero written over ds_steps (generic proxy) must be ordered before the bulk copy (async
      // proxy) that refills the buffer
      if (__any_sync(0xffffffffu, wrote_zero)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const bool still = __any_sync(0xffffffffu, mode != 0);
      if (threadIdx.x == 0) {
        if (warp_live && !still) atomicSub(&s_ctl[0], 1);
        mbar_arrive(bar_empty(c % kPStages));
        // issue every following chunk whose buffer has been released by all warps (chunk n reuses the buffer of chunk
        // n - kPStages); whoever releases a buffer last finds it free here, so no chunk is left unissued
        while (*reinterpret_cast<volatile int*>(&s_ctl[0]) > 0) {
          const int nx = *reinterpret_cast<volatile int*>(&s_ctl[1]);
          if (nx > last_chunk) break;
          if (nx >= kPStages && !mbar_test(bar_empty(nx % kPStages), (unsigned)(nx / kPStages - 1) & 1u)) break;
          if (atomicCAS(&s_ctl[1], nx, nx + 1) == nx) issue(nx);
        }
      }
      warp_live = still;
    }
#endif
  }
#if RB_RT_RING
  __syncthreads();
  if (any_live && tid == 0) {
    // bulk copies still in flight must land before the CTA gives up its shared memory: the last chunk of every buffer
    const int issued = s_ctl[1];
    for (int n = max(0, issued - kPStages); n < issued; ++n)
      while (!mbar_try_wait(bar_full(n % kPStages), (unsigned)(n / kPStages) & 1u)) {}
  }
#else
  cp_async_wait<0>();
#endif
  end_small();              // a ray that ended inside the small-tau phase
  if (streamed) {
    // the trace of the whole tile has ended (every ray counts into the last chunk when it ends): segment counts and
    // NaN flags are final
    if (tid == 0) geo_wait(k.geo_npub - 1);
    __syncthreads();
    if (validA) { n = k.nseg[tpos]; nanray = k.nanflag[tpos] != 0; }
  }
  if (k.step_counter) {   // measurement aid (bench.py): executed (ray, freq, segment) steps, one atomic per warp
    unsigned long long done = 2ull * (unsigned long long)nf + (unsigned long long)ns;
    unsigned long long done_a = (unsigned long long)ia;
    for (int o = 16; o > 0; o >>= 1) {
      done += __shfl_down_sync(0xffffffffu, done, o);
      done_a += __shfl_down_sync(0xffffffffu, done_a, o);
    }
    if (threadIdx.x == 0) { atomicAdd(k.step_counter, done); atomicAdd(k.step_counter + 1, done_a); }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (!(h ? validB : validA)) continue;
    const double tau = h ? tauB : tauA, iW = h ? iWB : iWA, Tb = h ? TbB : TbA;
    double vout, wout = iW;
    if (n < 0) vout = kTcmb;                                 // off planet (brightness.py:46-51)
    else if (nanray || tau != tau) vout = wout = nan("");    // NaN segment below the tangent shell / NaN alpha
    else vout = (Tb < kTcmb) ? kTcmb : Tb / iW;              // brightness.py:109-113
    const size_t o = (size_t)r * k.F + fA + h;
    if (k.out_f32) reinterpret_cast<float*>(k.out_Tb)[o] = (float)vout;
    else reinterpret_cast<double*>(k.out_Tb)[o] = vout;
    if (k.out_intW) k.out_intW[o] = (n < 0) ? 0.0 : wout;
  }
  if (TILES && rt_.dead) return;                             // a warp beyond the end of the list has no tile to report
  progress_report(k, rt_, tid, TILES);
}

// ====================================================================================================
// Mixed-precision variant of the rays-major integration (rb_set_rt_precision(ctx, RB_RT_MIXED)).
//
// Data.Tb is float32 in the reference (data_handling.py:46-47: 3e-5 K per ulp at 300 K) and the parity bar is
// 0.01 K, but the optical depth must be carried in FP64: exp(-tau) turns an absolute error of tau into a relative
// error of every later term, and a float32 running sum over ~500 segments loses 1e-5.  So tau stays in FP64 and
// the rest moves to the FP32 and SFU pipes, which idle in the FP64 kernel -- and the 8 KB exponential table with
// its bank-conflicting lookups (about half of the FP64 kernel's shared-memory wavefronts) disappears:
//   phase B (tau >= 2^-RB_RTM_SMALL_LOG), per (ray, freq, segment)
//     t   = fma(asum, ds, t)                              FP64; ds is the float copy widened with two integer
//                                                         instructions (its 2^-24 rounding moves tau by < 1e-8)
//     nd  = fma(t, -2^23 log2 e, 2^52 + 2^51 + 0x3F000000)  FP64: low word ni = k 2^23 + x + 0x3F000000, x in [0, 2^23)
//     y   = (ni & 0x7FFFFF) | 0x3F800000                  LOP3: the float 1 + x 2^-23, built from the bits of x (no
//                                                         FP64 -> FP32 conversion: those run at a quarter of the DFMA rate)
//     e   = ex2.approx(y) as bits + ni - y                MUFU.EX2 + IADD3: 2^(1 + x 2^-23) scaled by 2^(k-1)
//     w   = e (ds_i + ds_i+1);  iW += a' w;  Tb += T a' w        FP32 (FADD, FMUL, 2 FFMA) on float copies of ds
//                                                         (geometry kernel) and of the operands (prepare kernel)
//   phase A (tau < 2^-RB_RTM_SMALL_LOG, half of the executed steps): everything in FP32, tau included (its absolute
//     error stays below 1e-8 there), exp(-t) = 1 - t + t^2/2 (truncation t^3/6 < 2^-26 for t < 2^-8).
// The FP32 partial sums of one chunk (<= 32 segments) are added to FP64 accumulators at the end of the chunk.
// Shared-memory wavefronts per (warp, segment): 2 in phase A, 2.25 in phase B (FP64 kernel: 4.5 / ~8.5) -- the
// 128 B/clk shared-memory pipe is what both kernels run out of first (ncu: l1tex data pipe 74-81 % busy).
//
// Operand rows (rt_prepare_mixed_kernel): 32 bytes per frequency, 256 bytes per (frequency group, segment):
//   { double asum | float a' | float T a' || float a' | float T a' | float asum | 0 }     (all x 0.5e5)
// phase B reads the first 16 bytes, phase A the second.
// Shared-memory stage (x 2): kMxChunk operand rows | float ds tile (kMxChunk + 1) x 32 floats = 12.1 KB at 32
// segments per chunk; the FP64 ds slab is not read at all.
#ifndef RB_RTM_CTAS
#define RB_RTM_CTAS 4
#endif
#ifndef RB_RTM_SMALL_LOG
#define RB_RTM_SMALL_LOG 8
#endif
#ifndef RB_RTM_CHUNK
#define RB_RTM_CHUNK 32
#endif
#ifndef RB_RTM_GROUP
#define RB_RTM_GROUP 4
#endif
constexpr int kMxG = RB_RTM_GROUP;                        // segments per group (2 or 4): independent weight chains per thread
static_assert(kMxG == 2 || kMxG == 4 || kMxG == 8, "groups are processed in pairs");
constexpr int kMxChunk = RB_RTM_CHUNK;                    // segments per staged tile: 32 or 64
static_assert(kMxChunk == 32 || kMxChunk == 64, "copy plan below");
constexpr int kMxRow = 256;                               // bytes per operand row
constexpr int kMxTileP = kMxChunk * kMxRow;               // operand rows
constexpr int kMxTileF = (kMxChunk + 1) * 32 * 4;         // float segment lengths, one extra row (ds_i+1 of the last)
constexpr int kMxStage = kMxTileP + kMxTileF;
constexpr size_t kMixedSmemBytes = 2 * (size_t)kMxStage;
static_assert(kMxTileP % 4096 == 0 && (kMxTileF - 128) % 4096 == 0 && kMxStage % 16 == 0, "the copy plan below is written for these sizes");
constexpr double kMxTauMax = 85.0;   // 2^(k-1) stays a normal float while tau log2 e <= 126
// Every term beyond tau = 25 is below e^-25 (1.4e-11) x T (<= 2000 K) x dtau, i.e. < 3e-8 K in Tb: three orders below
// the resolution of this mode's FP32 partial sums (1e-5 K), so the mixed kernel never integrates deeper than that
// whatever tau_cut asks for (the FP64 kernel's default, 50, is the same argument at FP64 resolution).  7 % of the
// segment-steps of C4 lie between tau = 25 and tau = 50.
constexpr double kMxTauCut = 25.0;

struct alignas(16) MxOperand {
  double asum;                  // (a_i + a_i+1) 0.5e5
  float ay, az;                 // a_i+1 0.5e5,  T_i+1 a_i+1 0.5e5   (an aligned pair: one FFMA2 operand)
  float ay_f, az_f;             // the same pair again, for phase A's 16-byte read
  float asum_f, pad;
};
static_assert(sizeof(MxOperand) == 32, "operand layout");

__global__ void rt_prepare_mixed_kernel(const double* __restrict__ alpha, const double* __restrict__ T, int L, int F,
                                        int ngroups, MxOperand* __restrict__ prepm) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int Lm1 = L - 1;
  if (idx >= ngroups * Lm1 * 8) return;
  const int fl = idx & 7;
  const int i = (idx >> 3) % Lm1;
  const int fg = (idx >> 3) / Lm1;
  const int f = fg * 8 + fl;
  double vx = 0.0, vy = 0.0, vz = 0.0;
  if (f < F) {
    const double kHalfCm = 0.5 * kKmToCm;
    const double a0 = alpha[(size_t)i * F + f], a1 = alpha[(size_t)(i + 1) * F + f];
    vx = (a0 + a1) * kHalfCm; vy = a1 * kHalfCm; vz = (T[i + 1] * a1) * kHalfCm;
  }
  MxOperand o;
  o.asum = vx; o.ay = (float)vy; o.az = (float)vz;
  o.ay_f = (float)vy; o.az_f = (float)vz; o.asum_f = (float)vx; o.pad = 0.0f;
  prepm[idx] = o;                                            // idx = (fg (L-1) + i) 8 + fl
}

// packed FP32 pairs (sm_100 FFMA2 / FMUL2): one issue slot for two operations
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
#ifndef RB_RTM_PACKED
#define RB_RTM_PACKED 1      // 0: scalar FFMA / FMUL instead of FFMA2 / FMUL2 (measured: 2.33 against 2.29 ms)
#endif
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
#if RB_RTM_PACKED
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
#else
  float a0, a1, b0, b1, c0, c1;
  unpack2(a, a0, a1); unpack2(b, b0, b1); unpack2(c, c0, c1);
  return pack2(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
#endif
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
#if RB_RTM_PACKED
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
#else
  float a0, a1, b0, b1;
  unpack2(a, a0, a1); unpack2(b, b0, b1);
  return pack2(a0 * b0, a1 * b1);
#endif
}

// one 16-byte shared-memory read that ptxas cannot split into narrower ones (a split costs wavefronts: the
// kernel's scarcest resource after issue slots is the 128 B/clk shared-memory crossbar)
__device__ __forceinline__ ulonglong2 lds128(const void* p) {
  ulonglong2 v;
  asm("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

// float -> double of a non-negative normal float with two integer instructions (F2F.F64.F32 runs at a quarter of the
// DFMA rate): exponent rebias 127 -> 1023 and mantissa shift.  0 maps to 2^-127 (used only as a segment length: harmless).
__device__ __forceinline__ double f2d_bits(float f) {
  const unsigned b = __float_as_uint(f);
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}

// exp(-tau) as a float for 0 <= tau <= kMxTauMax (see the header comment);
// cA = -2^23 log2 e, cM = 2^52 + 2^51 + 0x3F000000, fmask = 0x7FFFFF in a register (one LOP3 does the and-or)
__device__ __forceinline__ float exp_neg_mixed(double tau, double cA, double cM, unsigned fmask) {
  const unsigned ni = (unsigned)__double2loint(fma(tau, cA, cM));
  unsigned y;
  asm("lop3.b32 %0, %1, %2, 0x3F800000, 0xEA;" : "=r"(y) : "r"(ni), "r"(fmask));   // (ni & fmask) | 0x3F800000
  float ev;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ev) : "f"(__uint_as_float(y)));
  return __uint_as_float(__float_as_uint(ev) + ni - y);
}

__global__ void __launch_bounds__(256, RB_RTM_CTAS) rt_integrate_rays_mixed_kernel(const __grid_constant__ RtK k) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int S = k.L - 1;
  const RayTile rt_ = map_ray_tile(k);                      // never a compacted launch (k.cidx == null)
  const unsigned tile = rt_.tile;
  const long long r = rt_.r;
  const int f = rt_.fg * 8 + threadIdx.y;
  const bool valid = rt_.in && (f < k.F);
  const int n = valid ? k.nseg[r] : -1;
  const bool nanray = valid && k.nanflag[r] != 0;
  const int steps = (valid && !nanray) ? n - 1 : 0;        // brightness.py:65: i = 0 .. len(ds)-2

  // copy plan of one chunk: kMxChunk / 16 rounds of 256 x 16 bytes for the operand rows, kMxChunk / 32 rounds and 8
  // extra pieces for the float segments, each stream one contiguous piece of global memory; unchecked like the
  // FP64 kernel's (rows past the end of a tile are never consumed, the slabs carry kRtSlackBytes of slack)
  const char* src_p = reinterpret_cast<const char*>(k.prepm) + (size_t)rt_.fg * S * kMxRow + tid * 16;
  const char* src_f = reinterpret_cast<const char*>(k.dsf + (size_t)tile * S * 32) + tid * 16;
  const unsigned dst = (unsigned)__cvta_generic_to_shared(s_raw) + tid * 16;
  auto issue = [&](int c) {
    const unsigned o = dst + ((c & 1) ? (unsigned)kMxStage : 0u);
    cp_rounds<kMxChunk / 16>(o, src_p);
    cp_rounds<kMxChunk / 32>(o + kMxTileP, src_f);
    if (tid < 8) cp_async16_at<kMxChunk * 128>(o + kMxTileP, src_f);
    cp_async_commit();
    src_p += kMxTileP;
    src_f += kMxChunk * 32 * 4;
  };
  bool live = steps > 0;
  const bool any_live = __syncthreads_or(live);
  if (any_live) issue(0);

  // loop-invariant constants fetched into registers (see pin()): c_mxc = {cA, cM}
  const double cA = pin(c_mxc + 0), cM = pin(c_mxc + 1);
  unsigned fmask;
  asm volatile("mov.b32 %0, 0x007FFFFF;" : "=r"(fmask));
  constexpr float kSmallF = 1.0f / (float)(1 << RB_RTM_SMALL_LOG);
  const double cutd = fmin(k.tau_cut, kMxTauCut);
  const int cut_hi = __double2hiint(cutd);                 // positive doubles order like their high words
  double tau = 0.0, iW = 0.0, Tb = 0.0;
  float tauf = 0.0f;
  unsigned long long acc = 0ull;                            // packed FP32 {sum a' w, sum T a' w} of the current chunk
  int i = 0, ia = 0;
  bool stop = false;
  bool small = cutd > (double)kSmallF;                     // phase A ignores tau_cut: only run it below the cut

  for (int c = 0; any_live; ++c) {
    cp_async_wait<0>();                                    // this thread's pieces of chunk c have landed
    if (!__syncthreads_or(live)) break;                    // ... everybody's have; chunk c-1 is fully consumed
    issue(c + 1);                                          // refill the buffer chunk c-1 used
    // a staged tile is consumed in pieces of 32 segments: the FP32 partial sums are flushed after each
#pragma unroll 1
    for (int h = 0; h < kMxChunk / 32 && live; ++h) {
      const unsigned char* st = s_raw + (c & 1) * kMxStage;
      const MxOperand* qp = reinterpret_cast<const MxOperand*>(st) + h * 32 * 8 + threadIdx.y;
      const float* fp = reinterpret_cast<const float*>(st + kMxTileP) + h * 32 * 32 + threadIdx.x;
      const int m = min(32, steps - i);
      int u = 0;
      if (small) {
        // phase A: all FP32, groups of kMxG segments; the group that reaches 2^-RB_RTM_SMALL_LOG is left to phase B
#pragma unroll 1
        for (; u + kMxG <= m; u += kMxG, fp += kMxG * 32, qp += kMxG * 8) {
          float d[kMxG + 1], t[kMxG];
          ulonglong2 q[kMxG];                                   // .x = {a', T a'} (FFMA2 operand), .y = {asum, 0}
#pragma unroll
          for (int j = 0; j <= kMxG; ++j) d[j] = fp[j * 32];
#pragma unroll
          for (int j = 0; j < kMxG; ++j) q[j] = lds128(&qp[j * 8].ay_f);
          t[0] = fmaf(__uint_as_float((unsigned)q[0].y), d[0], tauf);
#pragma unroll
          for (int j = 1; j < kMxG; ++j) t[j] = fmaf(__uint_as_float((unsigned)q[j].y), d[j], t[j - 1]);
          if (!(t[kMxG - 1] < kSmallF)) { small = false; break; }   // also leaves on NaN
#pragma unroll
          for (int j = 0; j < kMxG; j += 2) {
            // two segments per packed instruction: w = (1 - t + t^2/2) (ds_j + ds_j+1)
            const unsigned long long tt = pack2(t[j], t[j + 1]);
            const unsigned long long pp = ffma2(ffma2(tt, pack2(0.5f, 0.5f), pack2(-1.0f, -1.0f)), tt, pack2(1.0f, 1.0f));
            float w0, w1;
            unpack2(fmul2(pp, pack2(d[j] + d[j + 1], d[j + 1] + d[j + 2])), w0, w1);
            acc = ffma2(q[j].x, pack2(w0, w0), acc);
            acc = ffma2(q[j + 1].x, pack2(w1, w1), acc);
          }
          tauf = t[kMxG - 1];
        }
        ia += u;
        if (u < m) small = false;                            // chunk remainder (end of the ray): one by one below
        if (!small) tau = (double)tauf;
      }
      if (!small) {
        // phase B: FP64 optical depth, SFU exponential, FP32 weights
#pragma unroll 1
        for (; u + kMxG <= m; u += kMxG, fp += kMxG * 32, qp += kMxG * 8) {
          double t[kMxG];
          float d[kMxG + 1];
          unsigned long long qa[kMxG];                           // {a', T a'}
#pragma unroll
          for (int j = 0; j <= kMxG; ++j) d[j] = fp[j * 32];
#pragma unroll
          for (int j = 0; j < kMxG; ++j) {
            // one 16-byte read: {double asum, float a', float T a'}
            const ulonglong2 q = lds128(&qp[j * 8].asum);
            qa[j] = q.y;
            t[j] = fma(__longlong_as_double((long long)q.x), f2d_bits(d[j]), j ? t[j > 0 ? j - 1 : 0] : tau);
          }
          if (__double2hiint(t[kMxG - 1]) >= cut_hi) break;          // tau_cut (or a NaN) inside this group: one by one below
#pragma unroll
          for (int j = 0; j < kMxG; j += 2) {
            float w0, w1;
            unpack2(fmul2(pack2(exp_neg_mixed(t[j], cA, cM, fmask), exp_neg_mixed(t[j + 1], cA, cM, fmask)),
                          pack2(d[j] + d[j + 1], d[j + 1] + d[j + 2])), w0, w1);
            acc = ffma2(qa[j], pack2(w0, w0), acc);
            acc = ffma2(qa[j + 1], pack2(w1, w1), acc);
          }
          tau = t[kMxG - 1];
        }
        // chunk remainder / the group that crosses tau_cut: one segment at a time; the crossing step is included
        for (; u < m; fp += 32, qp += 8) {
          const float d0 = fp[0];
          tau = fma(qp->asum, f2d_bits(d0), tau);
          const bool cross = __double2hiint(tau) >= cut_hi;   // true for NaN as well (tau stays NaN -> NaN output)
          const double tc = cross ? fmin(tau, kMxTauMax) : tau;
          const float w = exp_neg_mixed(tc, cA, cM, fmask) * (d0 + fp[32]);
          acc = ffma2(pack2(qp->ay, qp->az), pack2(w, w), acc);
          ++u;
          if (cross) { stop = true; break; }
        }
      }
      i += u;                                                // segments integrated so far (feeds rb_count_steps)
      live = !stop && i < steps;
      float iWf, Tbf;                                        // FP32 partial sums of these 32 segments -> FP64 accumulators
      unpack2(acc, iWf, Tbf);
      iW += (double)iWf;
      Tb += (double)Tbf;
      acc = 0ull;
    }
  }
  cp_async_wait<0>();
  if (small) tau = (double)tauf;                             // a NaN optical depth that never left phase A
  if (k.step_counter) {
    unsigned long long done = (unsigned long long)i;
    unsigned long long done_a = (unsigned long long)ia;
    for (int o = 16; o > 0; o >>= 1) {
      done += __shfl_down_sync(0xffffffffu, done, o);
      done_a += __shfl_down_sync(0xffffffffu, done_a, o);
    }
    if (threadIdx.x == 0) { atomicAdd(k.step_counter, done); atomicAdd(k.step_counter + 1, done_a); }
  }
  if (valid) {
    double vout, wout = iW;
    if (n < 0) vout = kTcmb;                                 // off planet (brightness.py:46-51)
    else if (nanray || tau != tau) vout = wout = nan("");    // NaN segment below the tangent shell / NaN alpha
    else vout = (Tb < kTcmb) ? kTcmb : Tb / iW;              // brightness.py:109-113
    const size_t o = (size_t)r * k.F + f;
    if (k.out_f32) reinterpret_cast<float*>(k.out_Tb)[o] = (float)vout;
    else reinterpret_cast<double*>(k.out_Tb)[o] = vout;
    if (k.out_intW) k.out_intW[o] = (n < 0) ? 0.0 : wout;
  }
  progress_report(k, rt_, tid);
}
}  // namespace

int rb_launch_geometry(rb_context* ctx, const RtLaunch& g) {
  if (g.gtype == RB_GTYPE_GRAVITY) return rb_launch_gravity_geometry(ctx, g, nullptr);
  GeoK k{};
  k.L = g.L; k.radius = g.radius; k.n0 = g.n0; k.n1 = g.n1; k.q = g.q;
  k.cz = g.rot[0]; k.sz = g.rot[1]; k.cx = g.rot[2]; k.sx = g.rot[3];
  k.limb = g.limb; k.R = g.R; k.Rpad = g.Rpad; k.b = g.b; k.ds = g.ds; k.dsf = g.dsf; k.nseg = g.nseg; k.nanflag = g.nanflag;
  const int threads = 128;
  const long long blocks = (g.R + threads - 1) / threads;
  void* p_dr2;
  RB_TRY(rb_ensure(ctx, RB_BUF_DR2, (size_t)g.L * 8, &p_dr2));
  RB_CUDA(ctx, rb_time_begin(ctx, 1));
  layer_dr2_kernel<<<(g.L + 255) / 256, 256, 0, ctx->stream>>>(g.radius, g.L, (double*)p_dr2);
  ctx->launches += 1;
  k.dr2 = (const double*)p_dr2;
  if (g.compact) {
    // classify (findEdge), list the rays that hit, march only those (full warps, full tiles for the integration)
    if (!g.cidx || !g.ncomp || !g.zq || !g.blkcnt) return rb_fail(ctx, RB_ERR_INVALID, "geometry: compaction buffers missing");
    k.cidx = g.cidx; k.ncomp = g.ncomp; k.zq = g.zq; k.blkcnt = g.blkcnt;
    const unsigned eblocks = (unsigned)((g.R + kEdgeThreads - 1) / kEdgeThreads);
    ray_edge_kernel<<<eblocks, kEdgeThreads, 0, ctx->stream>>>(k);
    ray_compact_kernel<<<(unsigned)((g.R + kSortBlock - 1) / kSortBlock), kSortThreads, 0, ctx->stream>>>(k);
    ctx->launches += 2;
    if (g.prog) {
      k.prog = g.prog;
      k.npub = (g.L - 1 + kGeoPub - 1) / kGeoPub;
      RB_CUDA(ctx, cudaMemsetAsync(g.prog, 0, ((size_t)(g.Rpad / 32) * k.npub + 1) * sizeof(int), ctx->stream));
    }
    if (g.mid_event) RB_CUDA(ctx, cudaEventRecord(g.mid_event, ctx->stream));
  }
  ray_geometry_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(k);
  RB_CUDA(ctx, cudaGetLastError());
  RB_CUDA(ctx, rb_time_end(ctx, 1));
  ctx->launches += 1;
  return RB_OK;
}

int rb_launch_ray_fields(rb_context* ctx, const RtLaunch& g, double* out) {
  GeoK k{};
  k.L = g.L; k.radius = g.radius; k.n0 = g.n0; k.n1 = g.n1; k.q = g.q;
  k.cz = g.rot[0]; k.sz = g.rot[1]; k.cx = g.rot[2]; k.sx = g.rot[3];
  k.limb = g.limb; k.R = g.R; k.Rpad = g.Rpad; k.b = g.b; k.ds = g.ds; k.nseg = g.nseg; k.nanflag = g.nanflag;
  ray_fields_kernel<<<(unsigned)((g.R + 127) / 128), 128, 0, ctx->stream>>>(k, out);
  RB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return RB_OK;
}

int rb_launch_ds_transpose(rb_context* ctx, const RtLaunch& g, double* out) {
  const int S = g.L - 1;
  dim3 grid((unsigned)((g.R + 31) / 32), (S + 31) / 32), block(32, 8);
  ds_transpose_kernel<<<grid, block, 0, ctx->stream>>>(g.ds, g.R, g.Rpad, S, g.nseg, out);
  RB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return RB_OK;
}

int rb_launch_ds_to_slab(rb_context* ctx, const double* in, int64_t R, int64_t Rpad, int S, const int* nseg,
                         int* nanflag, double* slab) {
  dim3 grid((unsigned)((Rpad + 31) / 32), (S + 31) / 32), block(32, 8);
  RB_CUDA(ctx, cudaMemsetAsync(nanflag, 0, sizeof(int) * Rpad, ctx->stream));
  ds_to_slab_kernel<<<grid, block, 0, ctx->stream>>>(in, R, Rpad, S, nseg, nanflag, slab);
  RB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return RB_OK;
}

// Decide the integration path for a request of R_total rays and, for the rays-major path, build the
// per-(layer,freq) operand slab once (shared by all ray chunks of the request).
// rb_set_rt_tuning / RB_RT_PAIRS force the one- / two-frequency kernel (default: whichever wastes fewer frequency slots).
static bool choose_pairs(const rb_context* ctx, int F) {
  if (ctx->rt_pairs == 0 || F < 2) return false;
  if (ctx->rt_pairs == 1) return true;
  // slots executed per useful frequency; the pair kernel does ~1.2x the work per issue slot
  const double fill1 = (double)F / (8.0 * ((F + 7) / 8)), fill2 = (double)F / ((double)kPairFreqs * ((F + kPairFreqs - 1) / kPairFreqs));
  return fill2 * 1.2 >= fill1;
}

int rb_rt_prepare(rb_context* ctx, int L, const rb_rt_desc* rt, int64_t R_total, bool profile, bool have_pairs,
                  RtPrep* out) {
  out->use_rays = !profile && !rt->disc_average && R_total >= 512;
  out->mixed = out->use_rays && have_pairs && ctx->rt_precision == RB_RT_MIXED;
  out->pairs = false;
  out->tiles = false;
  out->prep = nullptr;
  const int F = rt->n_freqs;
  const int ngroups = (F + 7) / 8;
  out->fgroups = ngroups;
  if (!out->use_rays) return RB_OK;
  const int nel = ngroups * (L - 1) * 8;
  void* scratch;
  if (out->mixed) {
    RB_TRY(rb_ensure(ctx, RB_BUF_PREP, (size_t)ngroups * (L - 1) * kMxRow + kRtSlackBytes, &scratch));
    rt_prepare_mixed_kernel<<<(nel + 255) / 256, 256, 0, ctx->stream>>>(rt->alpha, rt->T, L, F, ngroups,
                                                                         (MxOperand*)scratch);
    RB_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    out->prep = scratch;
    return RB_OK;
  }
  if (!ctx->exp_tab) {
    RB_CUDA(ctx, cudaMalloc(&ctx->exp_tab, kExpTabDoubles * sizeof(double)));
    exp_tab_init_kernel<<<(kExpTabDoubles + 255) / 256, 256, 0, ctx->stream>>>(ctx->exp_tab);
    ctx->launches += 1;
  }
  out->pairs = choose_pairs(ctx, F);
  out->tiles = out->pairs && ctx->rt_tiles == 1 && !RB_RT_RING;
  out->fgroups = out->pairs ? (out->tiles ? (F + 1) / 2 : (F + kPairFreqs - 1) / kPairFreqs) : ngroups;
  if (out->pairs) {
    const int pw = out->tiles ? 1 : kPairWarps;               // frequency pairs per operand row
    const int ng2 = out->fgroups;
    const int nel2 = ng2 * (L - 1) * pw;
    RB_TRY(rb_ensure(ctx, RB_BUF_PREP, (size_t)nel2 * kPairRow * sizeof(double2) + kRtSlackBytes, &scratch));
    rt_prepare_pairs_kernel<<<(nel2 + 255) / 256, 256, 0, ctx->stream>>>(rt->alpha, rt->T, L, F, ng2, pw, (double2*)scratch);
  } else {
    RB_TRY(rb_ensure(ctx, RB_BUF_PREP, (size_t)ngroups * (L - 1) * 8 * sizeof(double4) + kRtSlackBytes, &scratch));
    rt_prepare_kernel<<<(nel + 255) / 256, 256, 0, ctx->stream>>>(rt->alpha, rt->T, L, F, ngroups, (double4*)scratch);
  }
  RB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  out->prep = scratch;
  return RB_OK;
}

static GravK grav_of(const rb_context* ctx) {
  GravK v{};
  v.L = ctx->grav.L; v.K = ctx->grav.K; v.latstep = ctx->grav.latstep;
  v.rmag = ctx->grav.rmag; v.gamma = ctx->grav.gamma;
  return v;
}

int rb_build_geoid_table(rb_context* ctx, int L, int K, int nJ, int nvw, const double* d_radius, const double* d_GM,
                         const double* d_Jn, const double* d_vwlat, const double* d_vwdat, double RJ, double omega_m,
                         double latstep) {
  GravK v = grav_of(ctx);
  v.L = L; v.K = K; v.nJ = nJ; v.nvw = nvw; v.radius = d_radius; v.GM = d_GM; v.Jn = d_Jn; v.vwlat = d_vwlat; v.vwdat = d_vwdat;
  v.RJ = RJ; v.omega_m = omega_m; v.latstep = latstep;
  geoid_table_kernel<<<dim3((L + 63) / 64, 2), 64, 0, ctx->stream>>>(v);
  RB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return RB_OK;
}

int rb_launch_gravity_geometry(rb_context* ctx, const RtLaunch& g, double* out_fields) {
  if (!ctx->grav.rmag || ctx->grav.L != g.L)
    return rb_fail(ctx, RB_ERR_INVALID, "geometry: gtype 'gravity' needs rb_set_gravity_model for this %d-layer profile", g.L);
  GeoK k{};
  k.L = g.L; k.radius = g.radius; k.n0 = g.n0; k.n1 = g.n1; k.q = g.q;
  k.cz = g.rot[0]; k.sz = g.rot[1]; k.cx = g.rot[2]; k.sx = g.rot[3];
  k.limb = g.limb; k.R = g.R; k.Rpad = g.Rpad; k.b = g.b; k.ds = g.ds; k.nseg = g.nseg; k.nanflag = g.nanflag;
  RB_CUDA(ctx, rb_time_begin(ctx, 1));
  ray_geometry_gravity_kernel<<<(unsigned)((g.R + 127) / 128), 128, 0, ctx->stream>>>(k, grav_of(ctx), out_fields);
  RB_CUDA(ctx, cudaGetLastError());
  RB_CUDA(ctx, rb_time_end(ctx, 1));
  ctx->launches += 1;
  return RB_OK;
}

// The sky fill only needs the classification of the rays (zq) and touches no output element the integration writes:
// it runs on a side stream beside the integration (memory-bound next to an FP64-bound kernel).  Whoever reads the
// outputs waits with rb_join_fill_miss.
int rb_launch_fill_miss(rb_context* ctx, const RtLaunch& g, int F, void* out_Tb, double* out_intW, int out_f32) {
  // RB_FILL_STREAM: 0 = on the context stream, ahead of the integration; 1 = side stream; 2 = side stream of the
  // highest priority (its CTAs get SM slots ahead of the thousands of integration CTAs launched right after it, so
  // the copy-out stream, which waits for the fill, is released at the start of the integration and not in its middle)
  {
    const char* e = getenv("RB_FILL_STREAM");                 // (read per call: tests switch it)
    ctx->fill_mode = e ? atoi(e) : 2;
    if (ctx->fill_mode < 0 || ctx->fill_mode > 2) ctx->fill_mode = 2;
  }
  for (auto& e : ctx->fill_ev)
    if (!e) RB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaStream_t user = ctx->stream, side = user;
  if (ctx->fill_mode == 1) {
    if (!ctx->aux[1]) RB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux[1], cudaStreamNonBlocking));
    side = ctx->aux[1];
  } else if (ctx->fill_mode == 2) {
    if (!ctx->fill_stream) {
      int lo = 0, hi = 0;
      RB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      RB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->fill_stream, cudaStreamNonBlocking, hi));
    }
    side = ctx->fill_stream;
  }
  if (side != user) {
    RB_CUDA(ctx, cudaEventRecord(ctx->fill_ev[0], user));
    RB_CUDA(ctx, cudaStreamWaitEvent(side, ctx->fill_ev[0], 0));
  }
  struct Restore { rb_context* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{ctx, user};
  ctx->stream = side;
  const long long nout = (long long)g.R * F;
  const bool aligned = (((uintptr_t)out_Tb | (uintptr_t)out_intW) & 15) == 0;
  if (F % 4 == 0 && aligned)
    rt_fill_miss4_kernel<<<(unsigned)((nout / 4 + 255) / 256), 256, 0, ctx->stream>>>(g.zq, g.R, F / 4, out_Tb, out_intW, out_f32);
  else
    rt_fill_miss_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, ctx->stream>>>(g.zq, g.R, F, out_Tb, out_intW, out_f32);
  RB_CUDA(ctx, cudaGetLastError());
  RB_CUDA(ctx, cudaEventRecord(ctx->fill_ev[1], side));
  ctx->launches += 1;
  return RB_OK;
}
int rb_join_fill_miss(rb_context* ctx, cudaStream_t stream) {
  RB_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->fill_ev[1], 0));
  return RB_OK;
}

int rb_launch_progress_init(rb_context* ctx, const RtLaunch& g, const RtProgress& pg, unsigned fgroups) {
  if (!g.compact || !pg.done) return rb_fail(ctx, RB_ERR_INVALID, "rt: progress init needs a compacted launch");
  rt_progress_init_kernel<<<1, 256, 0, ctx->stream>>>(pg, g.cidx, g.ncomp, (int)((g.R + 31) / 32), fgroups);
  RB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return RB_OK;
}

// opt a kernel into more than 48 KB of dynamic shared memory, once per context (= per device of the process)
template <typename K>
static int opt_in_smem(rb_context* ctx, K kernel, size_t bytes, int slot) {
  if (bytes > 48 * 1024 && !ctx->smem_opted[slot]) {
    RB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->smem_opted[slot] = true;
  }
  return RB_OK;
}

int rb_launch_integrate(rb_context* ctx, const RtLaunch& g, const rb_rt_desc* rt, const RtPrep& prep,
                        const RtProgress* progress, void* out_Tb,
                        double* out_intW, int64_t profile_ray, double* out_tau, double* out_W, double* out_Tblyr) {
  RtK k{};
  bool fill_pending = false;
  k.L = g.L; k.F = rt->n_freqs; k.R = g.R; k.Rpad = g.Rpad;
  k.alpha = rt->alpha; k.T = rt->T; k.ds = g.ds; k.nseg = g.nseg; k.nanflag = g.nanflag;
  k.alpha0 = rt->alpha0;
  if (rt->alpha0 && prep.use_rays)
    return rb_fail(ctx, RB_ERR_UNSUPPORTED, "rt: alpha0 (Doppler form) is built for requests below 512 rays");
  k.out_Tb = out_Tb; k.out_intW = out_intW; k.out_f32 = rt->out_f32; k.disc = rt->disc_average;
  k.tau_cut = (rt->tau_cut > 0.0) ? rt->tau_cut : INFINITY;
  k.profile_ray = profile_ray; k.out_tau = out_tau; k.out_W = out_W; k.out_Tblyr = out_Tblyr;
  const int fgroups = (k.F + 31) / 32;
  if (g.compact && (profile_ray >= 0 || !prep.use_rays || prep.mixed))
    return rb_fail(ctx, RB_ERR_INVALID, "rt: compacted geometry needs the FP64 rays-major integration");
  RB_CUDA(ctx, rb_time_begin(ctx, 2));
  if (profile_ray >= 0) {
    if (g.R != 1) return rb_fail(ctx, RB_ERR_INVALID, "rt: profile outputs need a single-ray launch");
    dim3 grid(1, fgroups), block(32, 1);
    if (k.disc && disc_smem_bytes(k.L) <= 200 * 1024) {
      RB_TRY(opt_in_smem(ctx, rt_disc_kernel, 200 * 1024, 3));
      rt_disc_kernel<<<dim3(1, k.F), kDiscThreads, disc_smem_bytes(k.L), ctx->stream>>>(k, 1);
    } else if (k.disc) rt_integrate_kernel<1, true, true><<<grid, block, 0, ctx->stream>>>(k);
    else rt_integrate_kernel<1, false, true><<<grid, block, 0, ctx->stream>>>(k);
  } else if (!prep.use_rays) {
    const int wy = 4;
    const long long gx = (g.R + wy - 1) / wy;
    if (gx > 2147483647LL) return rb_fail(ctx, RB_ERR_INVALID, "rt: too many rays for one launch");
    dim3 grid((unsigned)gx, fgroups), block(32, wy);
    if (k.disc && g.R <= 65535 && disc_smem_bytes(k.L) <= 200 * 1024) {
      RB_TRY(opt_in_smem(ctx, rt_disc_kernel, 200 * 1024, 3));
      rt_disc_kernel<<<dim3((unsigned)g.R, k.F), kDiscThreads, disc_smem_bytes(k.L), ctx->stream>>>(k, 0);
    } else if (k.disc) rt_integrate_kernel<1, true, false><<<grid, block, 0, ctx->stream>>>(k);
    else rt_integrate_kernel<1, false, false><<<grid, block, 0, ctx->stream>>>(k);
  } else {
    // rays-major mapping: CTAs of 32 rays x 8 (16: pair kernel) frequencies in a 1-D grid, the frequency groups
    // of a ray tile adjacent in launch order; operands prepared by rb_rt_prepare
    k.step_counter = ctx->step_counter;
    if (progress) k.progress = *progress;
    k.fgroups = (unsigned)prep.fgroups;
    k.ntiles = (unsigned)((g.R + 31) / 32);
    k.tile_blocks = prep.tiles ? (k.ntiles + kPairWarps - 1) / kPairWarps : k.ntiles;
    // Launch order: frequency group by frequency group (the low-frequency groups integrate twice as many layers as
    // the others: they go first, and the CTAs that are resident together do the same kind of work) instead of the
    // frequency groups of a ray tile next to each other -- measured on C4: 3.10 -> 2.93 ms on one GPU, 0.48 -> 0.40 ms
    // for a rank's share on eight, although every ds tile then comes from DRAM once per frequency group.  With the
    // copy-out pipeline the tile list is cut into parts that are launched one after the other, so that the copy chunks
    // still complete in order.  Device-resident outputs: parts of ~80 MB of ds tiles, which stay in L2 (126 MB) for
    // the eight frequency groups of the part -- C4 in one part reads 4.4 GB from DRAM for 0.94 GB of ds (2.895 ms),
    // in 12 parts 2.915 ms (profiles/r2_ab_parts.txt).  RB_RT_PARTS: 0 = the old order, n = n parts.
    {
      const char* e = getenv("RB_RT_PARTS");
      const unsigned long long tile_bytes = (unsigned long long)(k.L - 1) * 256ull * (prep.tiles ? kPairWarps : 1);
      k.part_tiles = 0;
      if (e || progress) {
        k.nparts = e ? (unsigned)atoi(e) : 12u;
      } else if (g.prog) {
        k.nparts = 1;                                          // behind a running trace: measured in one part (N = 2 / 4 / 8)
      } else {
        // (a compacted launch knows its number of tiles on the device only: k.nparts is the bound for a list of all rays)
        k.part_tiles = (unsigned)(80000000ull / tile_bytes);
        if (k.part_tiles < 1) k.part_tiles = 1;
        k.nparts = (k.tile_blocks + k.part_tiles / 2) / k.part_tiles;
        if (k.nparts < 1) k.nparts = 1;
      }
      if (k.nparts > k.tile_blocks) k.nparts = k.tile_blocks ? k.tile_blocks : 1u;
    }
    const unsigned long long nblocks = (unsigned long long)k.fgroups * (k.tile_blocks + k.nparts);
    if (nblocks > 2147483647ULL) return rb_fail(ctx, RB_ERR_INVALID, "rt: too many (ray tile, frequency group) blocks for one launch");
    dim3 block(32, prep.pairs ? kPairWarps : 8), grid((unsigned)nblocks);
    if (prep.mixed) {
      if (!g.dsf) return rb_fail(ctx, RB_ERR_INVALID, "rt: mixed-precision integration without the float ds slab");
      k.dsf = g.dsf;
      k.prepm = prep.prep;
      RB_TRY(opt_in_smem(ctx, rt_integrate_rays_mixed_kernel, kMixedSmemBytes, 0));
      rt_integrate_rays_mixed_kernel<<<grid, block, kMixedSmemBytes, ctx->stream>>>(k);
      RB_CUDA(ctx, cudaGetLastError());
      RB_CUDA(ctx, rb_time_end(ctx, 2));
      ctx->launches += 1;
      return RB_OK;
    }
    if (g.compact) {
      // sky pixels first: the compacted launch never visits them (the copy-out pipeline has filled them already,
      // before its copy stream was released: rb_launch_fill_miss)
      k.cidx = g.cidx; k.ncomp = g.ncomp;
      if (g.prog && prep.pairs && !prep.tiles && !RB_RT_RING && kPChunk == kGeoPub) {
        k.geo_prog = g.prog;
        k.geo_npub = (k.L - 1 + kGeoPub - 1) / kGeoPub;
        // RB_RT_STREAM_ORDER=1 launches the frequency groups last to first (the high frequencies reach tau_cut in the
        // upper half of the atmosphere: their CTAs need only the chunks the trace publishes first).  Measured on rank
        // shares of C4 (1/8, 1/4, 1/2 of the image): 0.63 -> 0.70, 1.01 -> 1.09, 1.80 -> 1.89 ms -- the long
        // low-frequency CTAs then form the tail of the launch, which costs more than the earlier start gains.
        const char* e = getenv("RB_RT_STREAM_ORDER");
        k.fg_reverse = e ? atoi(e) : 0;
      } else if (g.prog) {
        return rb_fail(ctx, RB_ERR_INVALID, "rt: a streamed trace needs the FP64 pair kernel");
      }
      if (!progress) { RB_TRY(rb_launch_fill_miss(ctx, g, k.F, out_Tb, out_intW, k.out_f32)); fill_pending = true; }
    }
    k.exp_tab = ctx->exp_tab;
    if (prep.pairs && prep.tiles) {
      k.prep2 = (const double2*)prep.prep;
      constexpr size_t smem = kExpTabDoubles * sizeof(double) + kTilesSmemBytes;
      static_assert(smem <= 227 * 1024, "shared memory limit");
#if !RB_RT_RING
      RB_TRY(opt_in_smem(ctx, rt_integrate_pairs_kernel<true>, smem, 0));
      rt_integrate_pairs_kernel<true><<<grid, block, smem, ctx->stream>>>(k);
#endif
    } else if (prep.pairs) {
      k.prep2 = (const double2*)prep.prep;
      constexpr size_t smem = kExpTabDoubles * sizeof(double) + kPairsSmemBytes;
      static_assert(smem <= 227 * 1024, "shared memory limit");
      if (k.geo_prog) {
        RB_TRY(opt_in_smem(ctx, rt_integrate_pairs_kernel<false, true>, smem, 4));
        rt_integrate_pairs_kernel<false, true><<<grid, block, smem, ctx->stream>>>(k);
      } else {
        RB_TRY(opt_in_smem(ctx, rt_integrate_pairs_kernel<false>, smem, 1));
        rt_integrate_pairs_kernel<false><<<grid, block, smem, ctx->stream>>>(k);
      }
    } else {
      k.prep = (const double4*)prep.prep;
      constexpr size_t smem = kExpTabDoubles * sizeof(double) + kRaysSmemBytes;
      static_assert(smem <= 227 * 1024, "shared memory limit");
      RB_TRY(opt_in_smem(ctx, rt_integrate_rays_kernel, smem, 2));
      rt_integrate_rays_kernel<<<grid, block, smem, ctx->stream>>>(k);
    }
  }
  RB_CUDA(ctx, cudaGetLastError());
  if (fill_pending) RB_TRY(rb_join_fill_miss(ctx, ctx->stream));
  RB_CUDA(ctx, rb_time_end(ctx, 2));
  ctx->launches += 1;
  return RB_OK;
}
