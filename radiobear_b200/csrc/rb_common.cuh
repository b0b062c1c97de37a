// radiobear_b200 -- shared device helpers and the context object behind the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/radiobear_b200.h"

#define RB_NUM_SMS_B200 148

// ---- error plumbing -------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct rb_context {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int num_sms = RB_NUM_SMS_B200;
  size_t smem_optin = 0;
  // catalogs on device: SoA [ncols][nlines]
  double* cat[RB_NUM_CATALOGS] = {nullptr};
  int cat_n[RB_NUM_CATALOGS] = {0};
  int cat_cols[RB_NUM_CATALOGS] = {0};
  // grow-only scratch buffers
  DevBuf buf[32];
  // timing
  bool timing = false;
  static constexpr int kEvRing = 256;
  cudaEvent_t ev[3][kEvRing][2] = {};
  int64_t ev_count[3] = {0, 0, 0};
  int ev_slot = 0;  // slot of the launch being timed
  // ray-chunk pipeline (geometry of chunk c+1 overlaps integration of chunk c and the D2H of chunk c-1)
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fill_ev[2] = {nullptr, nullptr};   // fork / join of the sky fill (rb_launch_fill_miss)
  cudaStream_t fill_stream = nullptr;            // highest-priority stream of the sky fill (RB_FILL_STREAM=2)
  int fill_mode = 2;                             // RB_FILL_STREAM of the last sky fill
  std::vector<cudaEvent_t> pipe_ev;
  int rt_chunks = 0;  // 0 = automatic
  int rt_precision = 0;  // RB_RT_F64 / RB_RT_MIXED: arithmetic of the rays-major integration (rb_set_rt_precision)
  int rt_pairs = -1;     // rb_set_rt_tuning: two frequencies per thread (-1 automatic)
  int rt_compact = 1;    // rb_set_rt_tuning: integrate the compacted list of rays that hit the planet
  int rt_stream = -1;    // RB_RT_STREAM_GEOMETRY: integrate behind a trace that is still running (-1: small requests)
  int rt_tiles = 0;      // RB_RT_TILES: pair kernel with one frequency pair x four ray tiles per CTA (0: four pairs x one tile)
  bool smem_opted[5] = {false, false, false, false, false};  // kernels opted into > 48 KB of dynamic shared memory
  int alpha_newton = -1; // Newton steps of the line reciprocal (RB_RCP_NEWTON), read once per context
  unsigned long long* step_counter = nullptr;  // device counters of integrated segment-steps (measurement aid)
  unsigned long long* step_counter_buf = nullptr;  // their allocation (rb_count_steps)
  int64_t last_small_steps = 0;                // of the count last read: steps taken in the small-tau phase
  // geometry computed ahead of the rt call that will use it (rb_geometry_prefetch[_dev]); single use
  struct GeoTicket {
    bool valid = false;
    int64_t R = 0;
    const void* b = nullptr;
    rb_geometry_desc g{};
    cudaEvent_t done = nullptr;
    cudaEvent_t mid = nullptr;     // list of hitting rays and progress counters are ready, the trace starts
    bool streamed = false;         // the trace publishes its progress (RtLaunch::prog)
    int32_t* prog = nullptr;       // its progress counters, the counter of started CTAs and that counter's final value
    int32_t* started = nullptr;
    unsigned started_target = 0;
  } ticket;
  // device-resident absorption (rb_alpha_layers_resident / rb_alpha_rescale_resident): the slab and the
  // per-constituent cube stay in RB_BUF_RES_TOTAL / RB_BUF_RES_CUBE between calls; generations count overwrites
  struct Resident {
    int L = 0, F = 0;            // slab [L][F]
    int cL = 0, cF = 0, cC = 0;  // cube [cL][cF][cC]
    uint64_t slab_gen = 0, cube_gen = 0;   // 0 = nothing resident
    uint64_t counter = 0;
  } res;
  // 'gravity' geoid (rb_set_gravity_model): the shape table [2 hemispheres][K grid latitudes][L layers] of
  // (shell radius, gamma) that Shape._calcGeoid's march visits (shape.py:141-221), built once per model
  struct Gravity {
    int L = 0, K = 0;
    double latstep = 0.0;
    double* rmag = nullptr;   // device [2][K][L]
    double* gamma = nullptr;  // device [2][K][L]
    double r0 = 0.0, rlast = 0.0;   // first / last radius the table was built for (checked at launch)
  } grav;
  double* exp_tab = nullptr;  // 2^(j/1024), j = 0..1023: copied into shared memory by every integration CTA
};

// record the start / stop event of one launch of kernel family `which` (no-ops unless timing is on)
inline cudaError_t rb_time_begin(rb_context* ctx, int which) {
  if (!ctx->timing) return cudaSuccess;
  ctx->ev_slot = (int)(ctx->ev_count[which] % rb_context::kEvRing);
  return cudaEventRecord(ctx->ev[which][ctx->ev_slot][0], ctx->stream);
}
inline cudaError_t rb_time_end(rb_context* ctx, int which) {
  if (!ctx->timing) return cudaSuccess;
  cudaError_t e = cudaEventRecord(ctx->ev[which][ctx->ev_slot][1], ctx->stream);
  ctx->ev_count[which] += 1;
  return e;
}

enum {
  RB_BUF_FREQS = 0, RB_BUF_T, RB_BUF_P, RB_BUF_GAS, RB_BUF_CLOUD, RB_BUF_SCALE, RB_BUF_TOTAL, RB_BUF_CUBE,
  RB_BUF_RADIUS, RB_BUF_B, RB_BUF_DS, RB_BUF_NSEG, RB_BUF_TB, RB_BUF_INTW, RB_BUF_PROFILE, RB_BUF_MISC, RB_BUF_PREP,
  RB_BUF_FLAGS, RB_BUF_CIDX, RB_BUF_ZQ, RB_BUF_BLKCNT, RB_BUF_DR2, RB_BUF_RES_TOTAL, RB_BUF_RES_CUBE, RB_BUF_ORDER, RB_BUF_PROG,
  RB_BUF_ALPHA0
};

// The rays-major integration kernel prefetches whole 32-segment chunks of the ds slab and of the operand
// slab without bounds checks (the rows it may over-read are never consumed); both buffers carry this slack.
constexpr size_t kRtSlackBytes = 128 * 256;
#define RB_MAX_PEERS 8   // GPUs of one NVSwitch domain a layer-sharded absorption run stores into
// Compaction of the rays that hit the planet (ray_compact_kernel): within super-blocks of kSortBlock consecutive rays
// the list is ordered by projected radius, so that the 32 rays of a tile (= the lanes of a warp of the layer march and
// of the integration) have nearly the same path: they change phase and finish together.  A tile of the list may then
// hold rays from anywhere in its super-block(s), and how many tiles of the list touch a copy-out chunk is only known on
// the device: the chunk counters of a compacted launch start at kProgressTarget minus that number
// (rt_progress_init_kernel) and the copy stream waits for kProgressTarget.
constexpr int kSortBlock = 8192;
constexpr unsigned kProgressTarget = 0x40000000u;
// The copy-out pipeline starts the integration in the middle of the ray list (see run_rt_pipeline) -- at the first ray
// tile of a super-block, so that the list of hitting rays can be entered at the same place: the first hit of that
// super-block (ray_compact_kernel leaves its list position in ncomp[1]).  Plain ray tiles (32 rays) to skip:
__host__ __device__ inline int rb_progress_shift(long long R) {
  const long long ntiles = (R + 31) / 32, per = kSortBlock / 32;
  return (int)(((ntiles / 2) / per) * per);
}
// warps (= frequency pairs) per CTA of rt_integrate_pairs_kernel: a CTA is 32 rays x 2 kPairWarps frequencies.
// The warps of a CTA advance chunk by chunk together (one barrier per 32 segments) although their frequencies reach
// the table phase and tau_cut at different layers; 4 warps (8 adjacent frequencies, 6 CTAs per SM) wait less for each
// other than 8 (16 frequencies, 3 CTAs per SM): 3.48 against 3.74 ms on C4, same bits (profiles/r2_ab_variants.jsonl).
#ifndef RB_RT_WARPS
#define RB_RT_WARPS 4
#endif
constexpr int kPairWarps = RB_RT_WARPS;
constexpr int kPairFreqs = 2 * kPairWarps;

int rb_fail(rb_context* ctx, int code, const char* fmt, ...);
int rb_ensure(rb_context* ctx, int which, size_t bytes, void** out);

#define RB_CUDA(ctx, call)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return rb_fail((ctx), RB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                      \
  } while (0)

#define RB_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != RB_OK) return _s; \
  } while (0)

// ---- kernel launchers implemented in the .cu files ---------------------------------------------
// d holds DEVICE pointers; h_freqs is a host copy of d->freqs (frequency-class scan)
int rb_launch_alpha(rb_context* ctx, const rb_alpha_desc* d, const double* h_freqs, double* out_total,
                    double* out_cube, int n_peer = 0, double* const* peer_out = nullptr, long long peer_row0 = 0);

int rb_launch_alpha_scale_sum(rb_context* ctx, const double* cube, const double* scale, int L, int F, int C,
                              double* total, double* out_cube);

struct RtLaunch {
  // geometry
  int L;
  const double* radius;  // device [L]
  double n0, n1, q;      // q = Rpol/Req (1 for sphere)
  double rot[4];         // cos(tip), sin(tip), cos(rotate), sin(rotate)
  int limb;
  int gtype;             // RB_GTYPE_*
  // rays
  int64_t R;
  int64_t Rpad;
  const double* b;  // device [R][2]
  double* ds;       // device slab [Rpad/32][L-1][32] (tiled by 32 rays, see rt_kernels.cu)
  float* dsf;       // device slab of float segments, same tiling (mixed-precision integration only, else null)
  int32_t* nseg;    // device [R]
  int32_t* nanflag; // device [R]: ray carries a NaN segment the integration would use
  // compacted geometry (FP64 rays-major integration of >= 512 point rays): the layer march and the integration
  // run over the list of rays that hit the planet; ds tiles / nseg / nanflag are indexed by list position
  bool compact;
  int32_t* cidx;    // device [R]: list position -> ray index
  int32_t* ncomp;   // device scalar: length of the list
  double* zq;       // device [R]: findEdge depth per ray, NaN = misses the planet
  int32_t* blkcnt;  // device [ceil(R / 256)]
  // streamed geometry (see rt_kernels.cu, "streamed geometry"): per (tile of the list, chunk of 32 segments) the number
  // of rays whose trace has passed the chunk; the integration starts while the trace is still running
  int32_t* prog;    // device [Rpad / 32][kGeoPubChunksMax] or null
  cudaEvent_t mid_event;   // recorded on the launching stream right before ray_geometry_kernel (or null)
};
constexpr int kGeoPub = 32;             // segments per published chunk (= the staging chunk of the pair kernel)
int rb_launch_geometry(rb_context* ctx, const RtLaunch& g);
int rb_launch_gravity_geometry(rb_context* ctx, const RtLaunch& g, double* out_fields);
int rb_build_geoid_table(rb_context* ctx, int L, int K, int nJ, int nvw, const double* d_radius, const double* d_GM,
                         const double* d_Jn, const double* d_vwlat, const double* d_vwdat, double RJ, double omega_m,
                         double latstep);
int rb_launch_fill_miss(rb_context* ctx, const RtLaunch& g, int F, void* out_Tb, double* out_intW, int out_f32);
int rb_join_fill_miss(rb_context* ctx, cudaStream_t stream);
int rb_launch_ray_fields(rb_context* ctx, const RtLaunch& g, double* out);
int rb_launch_ds_transpose(rb_context* ctx, const RtLaunch& g, double* out_ds_raymajor /*[R][L-1] device*/);
int rb_launch_ds_to_slab(rb_context* ctx, const double* ds_raymajor, int64_t R, int64_t Rpad, int S, const int* nseg,
                         int* nanflag, double* slab);
// Progress reporting of one integration launch: ray tiles [cut[c], cut[c+1]) form chunk c; every CTA adds 1 to
// done[c] of its chunk when its results are in global memory, so that a copy stream can wait (stream memory
// operation) until a chunk is complete and move it to the host while the same launch keeps integrating.
constexpr int kMaxProgressChunks = 16;
struct RtProgress {
  int nchunks = 0;
  int shift = 0;              // CTA y processes ray tile (y + shift) mod #tiles (see run_rt_pipeline)
  int cut[kMaxProgressChunks + 1] = {0};   // in processing order
  unsigned* done = nullptr;   // device, nchunks counters, zeroed before the launch
};

// compacted launches: start the chunk counters at (len + 4) * fgroups minus the CTAs that will report into them
int rb_launch_progress_init(rb_context* ctx, const RtLaunch& g, const RtProgress& pg, unsigned fgroups);

struct RtPrep {
  bool use_rays = false;       // rays-major kernel (R >= 512, point rays) or the lanes = frequency kernel
  bool mixed = false;          // rays-major kernel in mixed precision (operand rows of rt_prepare_mixed_kernel)
  bool pairs = false;          // FP64 rays-major kernel with two frequencies per thread (pair operand rows)
  bool tiles = false;          // ... in the decomposition where the warps of a CTA share the frequency pair, not the rays
  int fgroups = 0;             // frequency groups = CTAs (pair kernel, tiles: warps) that integrate one ray tile
  const void* prep = nullptr;  // operand slab of the rays-major kernel
};
int rb_rt_prepare(rb_context* ctx, int L, const rb_rt_desc* rt /*device pointers*/, int64_t R_total, bool profile,
                  bool have_pairs /*RtLaunch::dsf was written*/, RtPrep* out);
int rb_launch_integrate(rb_context* ctx, const RtLaunch& g, const rb_rt_desc* rt /*device pointers*/, const RtPrep& prep,
                        const RtProgress* progress,
                        void* out_Tb, double* out_intW, int64_t profile_ray, double* out_tau, double* out_W,
                        double* out_Tblyr);

// ---- device math helpers -----------------------------------------------------------------------
#ifdef __CUDACC__
// Reciprocal for the line-shape denominators.  MUFU.RCP64H seed (rcp.approx.ftz.f64, >= 20 good
// bits) refined by Newton steps in FP64; NEWTON = 2 is below 1 ulp-ish (2^-80 before rounding),
// NEWTON = 1 gives ~2^-40 which is 6 orders below the 1e-6 parity bar.
template <int NEWTON>
__device__ __forceinline__ double rb_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#pragma unroll
  for (int i = 0; i < NEWTON; ++i) {
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
}
#endif
