/*
 * radiobear_b200 -- C ABI of the B200 (sm_100a) hot paths of RadioBEAR.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.  Each entry
 * point names the reference interface it replaces (paths relative to david-deboer/radiobear
 * v2.0.1, radiobear/...).  The reference is pure Python, so the "FFI" a maintainer adds is a
 * ctypes stub; see INTEGRATION.md.  The Python host in radiobear_b200/ binds exactly these.
 *
 * Conventions
 *   - all floating point arrays are float64 unless stated; frequencies GHz, T in K, P in bar,
 *     lengths km, absorption in cm^-1 ('invcm') or dB/km.
 *   - every call returns RB_OK (0) or an error code; rb_last_error() gives the message.
 *     Nothing is printed from kernels.  No pointer is retained after a call returns.
 *   - functions without suffix take HOST pointers and copy in/out on the context stream
 *     (synchronous on return).  "_dev" functions take DEVICE pointers and only enqueue work on
 *     the context stream (asynchronous; call rb_synchronize or sync the stream yourself).
 *   - one context per (process, GPU); a context is not thread-safe.
 */
#ifndef RADIOBEAR_B200_H
#define RADIOBEAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_ABI_VERSION 1

/* status codes */
#define RB_OK 0
#define RB_ERR_INVALID 1      /* bad argument (Python raises ValueError) */
#define RB_ERR_CUDA 2         /* CUDA runtime failure (RuntimeError) */
#define RB_ERR_NOMEM 3        /* device allocation failed (MemoryError) */
#define RB_ERR_UNSUPPORTED 4  /* formalism / option not built (NotImplementedError) */

typedef struct rb_context rb_context;

/* ---- context --------------------------------------------------------------------------- */
int rb_abi_version(void);
/* Create a context on CUDA device `device`.  Fails (RB_ERR_CUDA) when no sm_100 GPU is there:
 * there is no CPU fallback. */
int rb_create(int device, rb_context** ctx);
void rb_destroy(rb_context* ctx);
const char* rb_last_error(const rb_context* ctx);
/* Launch on an existing CUDA stream (cudaStream_t as void*).  NULL is the legacy default stream (what
 * torch.cuda.current_stream().cuda_stream returns for torch's default stream), NOT "no stream". */
int rb_set_stream(rb_context* ctx, void* cuda_stream);
/* Go back to the context's own non-blocking stream (the state after rb_create). */
int rb_use_own_stream(rb_context* ctx);
int rb_synchronize(rb_context* ctx);
/* Number of kernel launches issued through this context since creation (bench bookkeeping). */
int64_t rb_launch_count(const rb_context* ctx);
/* Device time (ms, CUDA events on the context stream) of the last call of each kernel family;
 * which: 0 alpha_lines, 1 ray_geometry, 2 rt_integrate.  Valid after rb_synchronize when timing
 * was enabled with rb_enable_timing(ctx, 1). */
int rb_enable_timing(rb_context* ctx, int on);
double rb_last_kernel_ms(rb_context* ctx, int which);
/* Durations (ms) of the most recent launches of a kernel family, oldest first; returns how many were
 * written (<= max_out, <= 256).  Synchronises on the last event. */
int rb_kernel_ms_history(rb_context* ctx, int which, double* out_ms, int max_out);
/* Number of timed launches of a kernel family so far (a request may be cut into several ray chunks). */
int64_t rb_kernel_timed_count(const rb_context* ctx, int which);
/* Ray-chunk pipeline depth for large requests: n chunks whose geometry / integration / device->host copy
 * overlap on three streams (0 = automatic, 1 = no pipelining). */
int rb_set_rt_chunks(rb_context* ctx, int n);
/* Arithmetic of the batched ray integration (Brightness.single over >= 512 point rays, brightness.py:60-113).
 *   RB_RT_F64   : every operation in FP64 (|Tb - reference| ~ 1e-9 K).
 *   RB_RT_MIXED : the optical depth tau stays FP64; exp(-tau) on the SFU (ex2.approx), weights and per-chunk
 *                 partial sums in FP32, chunk sums accumulated in FP64.  Error against RB_RT_F64 is at the level of
 *                 one float32 ulp of Tb (Data.Tb is float32 in the reference, data_handling.py:46-47), two orders
 *                 inside the 0.01 K parity bar; see DESIGN.md 3.3 for the measured figure.
 * Disc-averaged runs, profile outputs, small batches and rb_rt_integrate always run in FP64.  The initial value is
 * RB_RT_DEFAULT_PRECISION unless the environment variable RB_RT_PRECISION is "f64" or "mixed". */
#define RB_RT_F64 0
#define RB_RT_MIXED 1
#define RB_RT_DEFAULT_PRECISION RB_RT_F64
int rb_set_rt_precision(rb_context* ctx, int precision);
int rb_get_rt_precision(const rb_context* ctx);
/* Work decomposition of the batched FP64 ray integration (results do not depend on it: bit-identical Tb).
 *   pairs   : 1 = two frequencies per thread (CTAs of 32 rays x 16 frequencies), 0 = one (32 x 8), -1 = automatic
 *             (whichever wastes fewer frequency slots for the request's F).  Environment: RB_RT_PAIRS.
 *   compact : 1 = trace and integrate only the rays that hit the planet, as full tiles of a compacted list, and
 *             scatter the results (requests of >= 512 point rays), 0 = walk the rays in the order given.
 *             Environment: RB_RT_COMPACT.
 * Kept as a call for A/B measurements and for the tests that compare the decompositions. */
int rb_set_rt_tuning(rb_context* ctx, int pairs, int compact);
/* Streamed ray trace (results do not depend on it).  A geometry started ahead of its consumer with
 * rb_geometry_prefetch[_dev] can publish its progress per (tile of 32 rays, chunk of 32 segments); the FP64 pair
 * integration of the following rb_rt_batch* call then starts as soon as the list of hitting rays exists and follows the
 * running trace chunk by chunk instead of waiting for its end.  The trace is a chain of ~1000 dependent steps per ray
 * however few rays there are, so this pays for small requests (a rank's rows of an image shared by several GPUs).
 *   mode : 1 = always, 0 = never, -1 = automatic (requests of at most 200 000 rays).
 * Environment: RB_RT_STREAM_GEOMETRY = 0 / 1.  Replaces nothing in the reference (raypath.py:108-273 and
 * brightness.py:30-126 run one after the other there). */
int rb_set_rt_stream_geometry(rb_context* ctx, int mode);
/* Measurement aid: count the (ray, freq, segment) steps the integration kernel actually executes (the
 * tau_cut early exit skips the rest).  enable = 1 resets and starts counting, 0 stops; the count is returned
 * (after synchronising the context stream). */
int64_t rb_count_steps(rb_context* ctx, int enable);
/* Of the count returned by the last rb_count_steps call: the steps taken in the kernel's small-optical-depth
 * phase (tau < 2^-11: short polynomial instead of the table exponential). */
int64_t rb_count_small_steps(const rb_context* ctx);

/* ---- line catalogs --------------------------------------------------------------------- *
 * Replaces the per-plugin npz readers: nh3_hs.py:62-67, nh3_sjs.py:17-23, h2s_ddb.py:14-39,
 * ph3_jh.py:18-61, co_ddb.py:14-19.  Truncation (truncate_strength / truncate_freq) is applied
 * by the caller before upload.  cols is row-major [ncols][nlines] (host memory).            */
#define RB_CAT_NH3_INV 0 /* fo, Io, Eo, gammaNH3o                                  (4 cols) */
#define RB_CAT_NH3_ROT 1 /* fo_rot, Io_rot, Eo_rot, gNH3_rot, gH2_rot, gHe_rot      (6 cols) */
#define RB_CAT_NH3_V2 2  /* fo_v2, Io_v2, Eo_v2                                    (3 cols) */
#define RB_CAT_NH3_SJS 3 /* f0, I0, E, G0                                          (4 cols) */
#define RB_CAT_H2S 4     /* f0, I0, E, GH2S                                        (4 cols) */
#define RB_CAT_PH3 5     /* f0, I0, E, WgtI0, WgtFGB, WgtSB                        (6 cols) */
#define RB_CAT_CO 6      /* f0, I0, E                                              (3 cols) */
#define RB_CAT_H2O 7     /* f_o, I_o, E_o, w_s, x_s, w_h2, w_he, x_h2, x_he  (9 cols, h2o_bk.py:23-49) */
/* Orton H2 CIA table prepared for the frequencies of the call (one "line" per frequency, in call order):
 * cols 0..9 the tabulated temperatures; then for each of the 3 pair tables of the chosen h2 state (h2-h2,
 * h2-he, h2-ch4; h2_orton.py:12-13) 10 values at those temperatures (after the quadratic interpolation in
 * frequency, h2_orton.py:77-108) followed by 9 x 3 coefficients (c1, c2, c3 of every interval) of the
 * not-a-knot cubic spline through them (scipy interp1d kind='cubic', h2_orton.py:203-211): 10 + 3 x 37 cols. */
#define RB_CAT_H2_ORTON 8
#define RB_NUM_CATALOGS 9
int rb_set_catalog(rb_context* ctx, int catalog, int nlines, int ncols, const double* cols);

/* ---- absorption (hot path A) ----------------------------------------------------------- *
 * Formalism ids = module names under constituents/<gas>/ (alpha.py:59-68).                  */
#define RB_F_NONE 0
#define RB_F_NH3_HS 1      /* nh3/nh3_hs.py:70-312      */
#define RB_F_NH3_DBS 2     /* nh3/nh3_dbs.py:70-313     */
#define RB_F_NH3_SJS 3     /* nh3/nh3_sjs.py:26-128     */
#define RB_F_NH3_HS_SJS 4  /* nh3/nh3_hs_sjs.py:6-26    */
#define RB_F_NH3_DBS_SJS 5 /* nh3/nh3_dbs_sjs.py:6-26   */
#define RB_F_H2S_DDB 6     /* h2s/h2s_ddb.py:42-87      */
#define RB_F_PH3_JH 7      /* ph3/ph3_jh.py:64-108      */
#define RB_F_H2O_BK 8      /* h2o/h2o_bk.py:65-187      */
#define RB_F_H2_JJ_DDB 9   /* h2/h2_jj_ddb.py:7-38      */
#define RB_F_H2_JJ 10      /* h2/h2_jj.py:7-22          */
#define RB_F_CLOUDS_IDP 11 /* clouds/clouds_idp.py:6-101 */
#define RB_F_CO_DDB 12     /* co/co_ddb.py:22-99        */
#define RB_F_NH3_KD 13     /* nh3/nh3_kd.py:115-351     */
#define RB_F_NH3_SJSD 14   /* nh3/nh3_sjsd.py:6-24      */
#define RB_F_NH3_BG 15     /* nh3/nh3_bg.py:26-74       */
#define RB_F_H2_ORTON 16   /* h2/h2_orton.py:126-222 (table prepared by the host: RB_CAT_H2_ORTON) */
#define RB_NUM_FORMALISMS 17
#define RB_MAX_CONSTITUENTS 8

#define RB_UNITS_INVCM 0
#define RB_UNITS_DBPERKM 1

/* gas rows used by the plugins (P_dict lookups, e.g. nh3_hs.py:101-103) */
enum { RB_GAS_H2 = 0, RB_GAS_HE, RB_GAS_CH4, RB_GAS_NH3, RB_GAS_H2O, RB_GAS_H2S, RB_GAS_PH3, RB_GAS_CO,
       RB_NUM_GAS };
/* cloud rows used by clouds_idp.py:17-47 */
enum { RB_CLD_H2O = 0, RB_CLD_SOLN, RB_CLD_NH4SH, RB_CLD_NH3, RB_CLD_H2S, RB_CLD_CH4, RB_NUM_CLD };

typedef struct rb_alpha_desc {
  int32_t n_layers;                        /* L */
  int32_t n_freqs;                         /* F */
  int32_t n_constituents;                  /* C <= RB_MAX_CONSTITUENTS, in CALL order = sorted(constituent), alpha.py:83 */
  int32_t formalism[RB_MAX_CONSTITUENTS];  /* RB_F_* per constituent */
  const double* freqs;                     /* [F] GHz */
  const double* T;                         /* [L] K    (atm.gas[C['T']]) */
  const double* P;                         /* [L] bar  (atm.gas[C['P']]) */
  const double* gas;                       /* [gas_rows][L] mixing ratios (atm.gas) */
  int32_t gas_rows;
  int32_t gas_col[RB_NUM_GAS];             /* row of H2,HE,CH4,NH3,H2O,H2S,PH3,CO in `gas` (P_dict); -1 = absent */
  const double* cloud;                     /* [cloud_rows][L] cloud densities (atm.cloud) or NULL */
  int32_t cloud_rows;
  int32_t cloud_col[RB_NUM_CLD];           /* rows in `cloud` (cloud_dict); -1 = absent */
  uint32_t cloud_flags;                    /* bit i set <=> other_dict enables species i (clouds_idp.py:17-45),
                                              bit order: ice_p(H2O) water_p(SOLN) nh4sh_p nh3ice_p h2sice_p ch4 */
  int32_t h2state;                         /* 0 'e', 1 'n' (h2_jj_ddb.py:17-30) */
  int32_t coshape;                         /* 0 voigt, 1 vvw, 2 diff, 3 other (co_ddb.py:39-42) */
  int32_t units;                           /* RB_UNITS_* (parameters.py:4-8) */
  const double* scale;                     /* [C][L] per-constituent per-layer scale or NULL (alpha.py:151-192, 235-259) */
  const double* freqs_host;                /* _dev calls only: optional HOST copy of freqs (saves one small D2H + sync) */
  int32_t freqs_per_layer;                 /* 1: freqs (and freqs_host) is [L][F], row l = the frequencies layer l is
                                              evaluated at -- Doppler-shifted absorption, where every step of a ray sees
                                              f / doppler (brightness.py:80-92).  0: one list [F] for all layers.  Not
                                              with h2_orton (its table is prepared per frequency list). */
} rb_alpha_desc;

/* Alpha.get_layers (alpha.py:261-305): total absorption for every (layer, freq).
 *   out_total : [L][F]  (layer-major "alpha slab"; Alpha.layers[F][L] is its transpose view)
 *   out_cube  : [L][F][C] per-constituent absorption after scaling (alpha.py:110-131) or NULL
 * A single plugin call (constituents/<gas>/<formalism>.alpha, alpha.py:210) is n_layers = 1.   */
int rb_alpha_layers(rb_context* ctx, const rb_alpha_desc* desc, double* out_total, double* out_cube);
int rb_alpha_layers_dev(rb_context* ctx, const rb_alpha_desc* desc, double* out_total, double* out_cube);
/* Layer-sharded absorption over the GPUs of one NVSwitch domain (one process per GPU): desc describes THIS rank's block
 * of layers (device pointers, like rb_alpha_layers_dev); peer_slabs[n_peers] are the device addresses of the full
 * [L_total][F] slab on every GPU, this one included, as mapped into this process (a symmetric / peer-accessible
 * allocation, e.g. torch.distributed._symmetric_memory); the kernel stores every value it computes into row
 * first_row + l of all of them, so the all_gather of the blocks happens inside the kernel, over
 * NVLink, behind the line sums.  The caller synchronises the ranks afterwards (a barrier with system-scope release /
 * acquire, e.g. the symmetric-memory barrier) before anybody reads the slab. */
int rb_alpha_layers_dev_scatter(rb_context* ctx, const rb_alpha_desc* desc, int32_t n_peers, const uint64_t* peer_slabs,
                                int64_t first_row);
/* Alpha.total_layer_alpha on a cached per-constituent cube (get_alpha='memory'/'file', alpha.py:151-192,
 * 224-225): out_total[l][f] = sum_c scale[c][l] * cube[l][f][c]; when out_cube != NULL it receives the scaled
 * cube (what a following save_alpha stores).  scale may be NULL (all ones).  Host pointers. */
int rb_alpha_scale_sum(rb_context* ctx, int32_t n_layers, int32_t n_freqs, int32_t n_constituents,
                       const double* cube, const double* scale, double* out_total, double* out_cube);

/* ---- device-resident absorption ------------------------------------------------------------
 * The reference's absorption cache keeps the per-constituent cube in host memory between runs (save_alpha /
 * get_alpha = 'memory', alpha.py:110-149) so that a retrieval loop only re-does the scale-sum and the radiative
 * transfer (scripts/demo_batch.py).  Here the slab [L][F] and the cube [L][F][C] stay in buffers owned by the context:
 * no copy back, no synchronisation; rb_rt_batch_resident integrates straight from the resident slab.  Every overwrite
 * of the slab / cube gets a new generation number (> 0), so a caller can tell whether "its" data is still there. */
/* rb_alpha_layers (host pointers in desc) into the resident slab and, when keep_cube, the resident cube. */
int rb_alpha_layers_resident(rb_context* ctx, const rb_alpha_desc* desc, int32_t keep_cube, uint64_t* out_slab_generation,
                             uint64_t* out_cube_generation);
/* Alpha.total_layer_alpha on the resident cube (alpha.py:151-192): resident slab[l][f] = sum_c scale[c][l] cube[l][f][c].
 * scale: host [C][L] or NULL (all ones).  The cube is not modified. */
int rb_alpha_rescale_resident(rb_context* ctx, const double* scale, uint64_t* out_slab_generation);
/* slab_shape[2] = {L, F}, cube_shape[3] = {L, F, C}; a generation of 0 means nothing is resident.  Any may be NULL. */
int rb_alpha_resident_info(rb_context* ctx, int32_t* slab_shape, uint64_t* slab_generation, int32_t* cube_shape,
                           uint64_t* cube_generation);
/* Copy the resident slab [L][F] / cube [L][F][C] to host memory (either may be NULL); synchronises. */
int rb_alpha_fetch(rb_context* ctx, double* out_total, double* out_cube);

/* ---- ray geometry + radiative transfer (hot path B) ------------------------------------- */
#define RB_GTYPE_ELLIPSE 0 /* shape.py:223-274 */
#define RB_GTYPE_SPHERE 1  /* 'circle' / 'sphere' */
#define RB_GTYPE_GRAVITY 2 /* shape.py:141-221; needs rb_set_gravity_model */
#define RB_LIMB_SHAPE 0
#define RB_LIMB_SEC 1      /* raypath.py:218-219 */

typedef struct rb_geometry_desc {
  int32_t n_layers;         /* L */
  const double* radius;     /* [L] equatorial radius of each layer, km  (atm.property[LP['R']], raypath.py:124) */
  double n0, n1;            /* refractive index of the two outermost layers (raypath.py:158) */
  double Req, Rpol;         /* config.Req / config.Rpol, km */
  double orientation[2];    /* [position angle, sub-earth latitude] degrees (raypath.py:39-44) */
  int32_t gtype;            /* RB_GTYPE_* */
  int32_t limb;             /* RB_LIMB_* */
} rb_geometry_desc;

/* The 'gravity' shape (Shape._calcGeoid / _gravity, shape.py:141-221): the geoid is marched from the equator in steps of
 * latstep degrees with the gravity vector of the zonal harmonics Jn, the rotation omega_m and the zonal winds vw(lat);
 * the shape the ray loop gets is that of the last grid latitude of the march.  The call builds, on the device, the table
 * of every (layer, grid latitude, hemisphere) the march can return; geometry requests with gtype = RB_GTYPE_GRAVITY
 * then look it up.  Host pointers, copied.
 *   radius   [L] equatorial radius of each layer (the profile later passed in rb_geometry_desc.radius)
 *   GM_layer [L] GM used for the march that starts at radius[l]: np.interp(radius[l], R profile, GM profile) as
 *                shape.py:156-157 calls it
 *   vwlat, vwdat [n_vw] zonal wind table (config.vwlat / vwdat, m/s), latitudes ascending */
typedef struct rb_gravity_model {
  int32_t n_layers;
  const double* radius;
  const double* GM_layer;
  int32_t n_J;
  const double* Jn;
  double RJ;
  double omega_m;
  int32_t n_vw;
  const double* vwlat;
  const double* vwdat;
  double latstep;           /* degrees; the reference's default_gravcalc_latstep is 0.01 */
  double max_lat;           /* degrees covered by the table (<= 90) */
} rb_gravity_model;
int rb_set_gravity_model(rb_context* ctx, const rb_gravity_model* model);

/* raypath.compute_ds (raypath.py:108-273) for n_rays impact points b[R][2].
 *   out_ds   : [R][L-1] path length per layer, km (NaN from the tangent depth on, as in the reference)
 *   out_nseg : [R] number of valid segments; -1 = ray misses the planet (Ray.ds is None)
 *   out_aspect: [3] tip, rotate (rad), rNorm (km)                                              */
int rb_compute_ds(rb_context* ctx, const rb_geometry_desc* geom, int64_t n_rays, const double* b,
                  double* out_ds, int32_t* out_nseg, double* out_aspect);

/* The descriptive per-step fields of raypath.compute_ds beside ds (raypath.py:186-187, 224): for every ray
 *   out_fields[r][0][i] = Ray.r4ds[i], the shell radius at the latitude of the point where step i starts, km
 *   out_fields[r][1][i], out_fields[r][2][i] = planetocentric latitude / longitude of that point, degrees (what
 *   Ray.doppler is computed from, raypath.py:186); [R][3][L-1], zero beyond the ray's last step.  Host pointers. */
int rb_compute_ray_fields(rb_context* ctx, const rb_geometry_desc* geom, int64_t n_rays, const double* b,
                          double* out_fields);

typedef struct rb_rt_desc {
  int32_t n_freqs;          /* F */
  const double* alpha;      /* [L][F] alpha slab in cm^-1 (rb_alpha_layers out_total) */
  const double* T;          /* [L] K */
  int32_t disc_average;     /* 1: W = 2 a E2(tau) (brightness.py:95-96) -- use with b = (0,0) */
  int32_t out_f32;          /* 1: out_Tb is float32 (Data.Tb dtype, data_handling.py:46-47) else float64 */
  double tau_cut;           /* stop a ray once tau > tau_cut (exp(-tau) no longer representable in the sums);
                               <= 0 disables.  The reference integrates every layer. */
  const double* alpha0;     /* NULL, or a second [L][F] slab: dtau of step i becomes (alpha0[i] + alpha[i+1]) ds_i / 2 while
                               the weights keep alpha[i+1] -- the Doppler form of brightness.py:80-96, where the upper node
                               of a step is evaluated at its own shifted frequency (a0 there) instead of re-using the lower
                               node of the step before.  rb_rt_integrate[_profile] only (requests below 512 rays). */
} rb_rt_desc;

/* Brightness.single over a batch of rays (brightness.py:30-126): geometry + tau/W/Tb integration.
 *   out_Tb          : [R][F] brightness temperature (off-planet rays: 2.725, brightness.py:46-51)
 *   out_integrated_W: [R][F] or NULL
 *   profile_ray >= 0 additionally returns for that ray out_tau / out_W / out_Tb_lyr, each [F][L-1]
 *                    (Brightness.tau / .W / .Tb_lyr, brightness.py:118-120); pass -1 / NULL otherwise. */
int rb_rt_batch(rb_context* ctx, const rb_geometry_desc* geom, const rb_rt_desc* rt, int64_t n_rays,
                const double* b, void* out_Tb, double* out_integrated_W, int64_t profile_ray, double* out_tau,
                double* out_W, double* out_Tb_lyr);
/* rb_rt_batch reading the resident absorption slab (rb_alpha_layers_resident / rb_alpha_rescale_resident) instead of
 * rt->alpha, which is ignored; geom->n_layers and rt->n_freqs must match the resident shape. */
int rb_rt_batch_resident(rb_context* ctx, const rb_geometry_desc* geom, const rb_rt_desc* rt, int64_t n_rays,
                         const double* b, void* out_Tb, double* out_integrated_W, int64_t profile_ray, double* out_tau,
                         double* out_W, double* out_Tb_lyr);
/* Device-pointer variant: geom->radius, rt->alpha, rt->T, b, out_* are device pointers. */
int rb_rt_batch_dev(rb_context* ctx, const rb_geometry_desc* geom, const rb_rt_desc* rt, int64_t n_rays,
                    const double* b, void* out_Tb, double* out_integrated_W);

/* Start the ray geometry (raypath.compute_ds of every ray, raypath.py:108-273) of the NEXT rb_rt_batch[_dev]
 * call now, on an internal stream forked from the context stream, so that it overlaps whatever the caller
 * enqueues in between -- typically rb_alpha_layers[_dev], which does not depend on it.  The next
 * rb_rt_batch (after rb_geometry_prefetch: host geom->radius / b) or rb_rt_batch_dev (after
 * rb_geometry_prefetch_dev: device pointers) with the same n_rays, b pointer and geometry descriptor waits
 * for that stream and skips its own geometry launch; any other ray call drops the prefetched geometry.  The
 * caller must not change b[] in between.  Purely an ordering optimisation: results are identical. */
int rb_geometry_prefetch(rb_context* ctx, const rb_geometry_desc* geom, int64_t n_rays, const double* b);
int rb_geometry_prefetch_dev(rb_context* ctx, const rb_geometry_desc* geom, int64_t n_rays, const double* b);

/* Integration only, for caller-supplied segments (Brightness.single with a given Ray):
 *   ds : [R][S] km host, nseg[R]; layer4ds is implicit 0..nseg-1 (raypath.py:222-225).          */
int rb_rt_integrate(rb_context* ctx, const rb_rt_desc* rt, int32_t n_layers, int64_t n_rays, int32_t n_seg,
                    const double* ds, const int32_t* nseg, void* out_Tb, double* out_integrated_W);
/* The same with the profile outputs of rb_rt_batch for one ray (Brightness.tau / .W / .Tb_lyr, brightness.py:118-120):
 * profile_ray >= 0 and out_tau / out_W / out_Tb_lyr each [F][S], or -1 / NULL. */
int rb_rt_integrate_profile(rb_context* ctx, const rb_rt_desc* rt, int32_t n_layers, int64_t n_rays, int32_t n_seg,
                            const double* ds, const int32_t* nseg, void* out_Tb, double* out_integrated_W,
                            int64_t profile_ray, double* out_tau, double* out_W, double* out_Tb_lyr);

/* ---- measurement probes (bench.py / tests) ------------------------------------------------- */
/* Sustained FP64 FMA rate of the device in TFLOP/s (2 flops per DFMA, 16 independent chains per thread,
 * CUDA events): the roofline denominator of the FP64-bound kernels. */
int rb_probe_fp64_peak(rb_context* ctx, int iters, double* out_tflops);
/* Measurement aid: time in ms of `iters` trips of 16 DFMA + n_mufu (0, 2, 4, 8) MUFU.RCP64H per thread, 8 x 256 threads
 * per SM; rrr: 0 one register operand, 1 three distinct register pairs per DFMA, 2 two + an immediate, 3 three with two
 * shared by all chains (n_mufu = 0 only for 2 and 3).  Shows what the reciprocal seed and the register
 * file cost the FP64 pipe. */
int rb_probe_fp64_mix(rb_context* ctx, int iters, int n_mufu, int rrr, double* out_ms);
/* y[i] = reciprocal of x[i] as the line loops compute it: MUFU.RCP64H seed + `newton` (0..2) Newton steps. */
int rb_probe_rcp(rb_context* ctx, int newton, int n, const double* x, double* y);

#ifdef __cplusplus
}
#endif
#endif /* RADIOBEAR_B200_H */
