// radiobear_b200 -- the extern "C" boundary (include/radiobear_b200.h): context, catalogs,
// host<->device staging around the kernels.  No torch, no C++ types in the signatures.
#include "rb_common.cuh"
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cuda.h>

int rb_fail(rb_context* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

int rb_ensure(rb_context* ctx, int which, size_t bytes, void** out) {
  DevBuf& b = ctx->buf[which];
  if (bytes == 0) bytes = 16;
  if (b.cap < bytes) {
    if (b.p) {
      cudaStreamSynchronize(ctx->stream);
      cudaFree(b.p);
      b.p = nullptr;
      b.cap = 0;
    }
    size_t want = bytes + (bytes >> 3);  // 12.5 % headroom
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&b.p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return rb_fail(ctx, RB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = want;
  }
  *out = b.p;
  return RB_OK;
}

extern "C" {

int rb_abi_version(void) { return RB_ABI_VERSION; }

int rb_create(int device, rb_context** out) {
  if (!out) return RB_ERR_INVALID;
  *out = nullptr;
  rb_context* ctx = new rb_context();
  *out = ctx;  // returned even on failure so that rb_last_error works; caller destroys it
  ctx->device = device;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return rb_fail(ctx, RB_ERR_CUDA, "no CUDA device available (%s): radiobear_b200 has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= count) return rb_fail(ctx, RB_ERR_INVALID, "device %d out of range (0..%d)", device, count - 1);
  RB_CUDA(ctx, cudaSetDevice(device));
  cudaDeviceProp prop;
  RB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return rb_fail(ctx, RB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                   prop.major, prop.minor);
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  RB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  ctx->rt_precision = RB_RT_DEFAULT_PRECISION;
  if (const char* pe = getenv("RB_RT_PRECISION")) {
    if (!strcmp(pe, "mixed")) ctx->rt_precision = RB_RT_MIXED;
    else if (!strcmp(pe, "f64")) ctx->rt_precision = RB_RT_F64;
    else return rb_fail(ctx, RB_ERR_INVALID, "RB_RT_PRECISION must be f64 or mixed, got '%s'", pe);
  }
  if (const char* e = getenv("RB_RT_PAIRS")) ctx->rt_pairs = atoi(e) < 0 ? -1 : (atoi(e) ? 1 : 0);
  if (const char* e = getenv("RB_RT_COMPACT")) ctx->rt_compact = atoi(e) ? 1 : 0;
  if (const char* e = getenv("RB_RT_TILES")) ctx->rt_tiles = atoi(e) ? 1 : 0;
  for (int i = 0; i < 3; ++i)
    for (int r = 0; r < rb_context::kEvRing; ++r)
      for (int j = 0; j < 2; ++j) RB_CUDA(ctx, cudaEventCreate(&ctx->ev[i][r][j]));
  return RB_OK;
}

void rb_destroy(rb_context* ctx) {
  if (!ctx) return;
  if (ctx->own_stream) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->buf)
      if (b.p) cudaFree(b.p);
    for (auto& c : ctx->cat)
      if (c) cudaFree(c);
    if (ctx->exp_tab) cudaFree(ctx->exp_tab);
    if (ctx->grav.rmag) cudaFree(ctx->grav.rmag);
    if (ctx->grav.gamma) cudaFree(ctx->grav.gamma);
    if (ctx->step_counter_buf) cudaFree(ctx->step_counter_buf);
    if (ctx->ticket.done) cudaEventDestroy(ctx->ticket.done);
    if (ctx->ticket.mid) cudaEventDestroy(ctx->ticket.mid);
    for (auto e : ctx->fill_ev) if (e) cudaEventDestroy(e);
    if (ctx->fill_stream) cudaStreamDestroy(ctx->fill_stream);
    for (int i = 0; i < 3; ++i)
      for (int r = 0; r < rb_context::kEvRing; ++r)
        for (int j = 0; j < 2; ++j)
          if (ctx->ev[i][r][j]) cudaEventDestroy(ctx->ev[i][r][j]);
    for (auto& a : ctx->aux)
      if (a) cudaStreamDestroy(a);
    for (auto& e : ctx->pipe_ev) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->own_stream);
  }
  delete ctx;
}

const char* rb_last_error(const rb_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rb_set_stream(rb_context* ctx, void* s) {
  if (!ctx) return RB_ERR_INVALID;
  ctx->stream = reinterpret_cast<cudaStream_t>(s);  // NULL = legacy default stream
  return RB_OK;
}

int rb_use_own_stream(rb_context* ctx) {
  if (!ctx) return RB_ERR_INVALID;
  ctx->stream = ctx->own_stream;
  return RB_OK;
}

int rb_synchronize(rb_context* ctx) {
  if (!ctx) return RB_ERR_INVALID;
  RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RB_OK;
}

int64_t rb_launch_count(const rb_context* ctx) { return ctx ? ctx->launches : 0; }

int rb_enable_timing(rb_context* ctx, int on) {
  if (!ctx) return RB_ERR_INVALID;
  ctx->timing = on != 0;
  return RB_OK;
}

static double ring_ms(rb_context* ctx, int which, int64_t idx) {
  const int r = (int)(idx % rb_context::kEvRing);
  float ms = 0.f;
  if (cudaEventSynchronize(ctx->ev[which][r][1]) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, ctx->ev[which][r][0], ctx->ev[which][r][1]) != cudaSuccess) return -1.0;
  return ms;
}

double rb_last_kernel_ms(rb_context* ctx, int which) {
  if (!ctx || which < 0 || which > 2 || ctx->ev_count[which] == 0) return -1.0;
  return ring_ms(ctx, which, ctx->ev_count[which] - 1);
}

int rb_kernel_ms_history(rb_context* ctx, int which, double* out_ms, int max_out) {
  if (!ctx || which < 0 || which > 2 || !out_ms || max_out <= 0) return 0;
  int64_t n = ctx->ev_count[which];
  if (n > rb_context::kEvRing) n = rb_context::kEvRing;
  if (n > max_out) n = max_out;
  const int64_t first = ctx->ev_count[which] - n;
  for (int64_t i = 0; i < n; ++i) out_ms[i] = ring_ms(ctx, which, first + i);
  return (int)n;
}

int64_t rb_kernel_timed_count(const rb_context* ctx, int which) {
  return (ctx && which >= 0 && which <= 2) ? ctx->ev_count[which] : 0;
}

int64_t rb_count_steps(rb_context* ctx, int enable) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  unsigned long long h[2] = {0, 0};
  if (!ctx->step_counter_buf && cudaMalloc(&ctx->step_counter_buf, sizeof(h)) != cudaSuccess) return -1;
  unsigned long long* const dev = ctx->step_counter_buf;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->step_counter) cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost);
  if (enable) {
    cudaMemset(dev, 0, sizeof(h));
    ctx->step_counter = dev;
  } else {
    ctx->step_counter = nullptr;
  }
  ctx->last_small_steps = (int64_t)h[1];
  return (int64_t)h[0];
}

int64_t rb_count_small_steps(const rb_context* ctx) { return ctx ? ctx->last_small_steps : -1; }

int rb_set_rt_chunks(rb_context* ctx, int n) {
  if (!ctx || n < 0) return RB_ERR_INVALID;
  ctx->rt_chunks = n;
  return RB_OK;
}

static void drop_ticket(rb_context* ctx);

int rb_set_rt_precision(rb_context* ctx, int precision) {
  if (!ctx) return RB_ERR_INVALID;
  if (precision != RB_RT_F64 && precision != RB_RT_MIXED)
    return rb_fail(ctx, RB_ERR_INVALID, "rt precision %d unknown (RB_RT_F64 / RB_RT_MIXED)", precision);
  if (precision != ctx->rt_precision) drop_ticket(ctx);   // a prefetched geometry has the other slab layout
  ctx->rt_precision = precision;
  return RB_OK;
}

int rb_get_rt_precision(const rb_context* ctx) { return ctx ? ctx->rt_precision : -1; }

int rb_set_rt_tuning(rb_context* ctx, int pairs, int compact) {
  if (!ctx) return RB_ERR_INVALID;
  if (pairs < -1 || pairs > 1 || compact < 0 || compact > 1)
    return rb_fail(ctx, RB_ERR_INVALID, "rt tuning: pairs must be -1 / 0 / 1 and compact 0 / 1");
  if (compact != ctx->rt_compact) drop_ticket(ctx);        // a prefetched geometry has the other indexing
  ctx->rt_pairs = pairs;
  ctx->rt_compact = compact;
  return RB_OK;
}

int rb_set_rt_stream_geometry(rb_context* ctx, int mode) {
  if (!ctx) return RB_ERR_INVALID;
  if (mode < -1 || mode > 1) return rb_fail(ctx, RB_ERR_INVALID, "rt stream geometry: mode must be -1 / 0 / 1");
  ctx->rt_stream = mode < 0 ? 2 : mode;
  return RB_OK;
}

int rb_set_catalog(rb_context* ctx, int catalog, int nlines, int ncols, const double* cols) {
  if (!ctx) return RB_ERR_INVALID;
  static const int want_cols[RB_NUM_CATALOGS] = {4, 6, 3, 4, 4, 6, 3, 9, 121};
  if (catalog < 0 || catalog >= RB_NUM_CATALOGS) return rb_fail(ctx, RB_ERR_INVALID, "catalog id %d unknown", catalog);
  if (ncols != want_cols[catalog])
    return rb_fail(ctx, RB_ERR_INVALID, "catalog %d needs %d columns, got %d", catalog, want_cols[catalog], ncols);
  if (nlines < 0 || (nlines > 0 && !cols)) return rb_fail(ctx, RB_ERR_INVALID, "catalog %d: bad line count / pointer", catalog);
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->cat[catalog]) {
    cudaFree(ctx->cat[catalog]);
    ctx->cat[catalog] = nullptr;
  }
  ctx->cat_n[catalog] = nlines;
  ctx->cat_cols[catalog] = ncols;
  if (nlines == 0) return RB_OK;
  const size_t bytes = (size_t)nlines * ncols * sizeof(double);
  RB_CUDA(ctx, cudaMalloc(&ctx->cat[catalog], bytes));
  RB_CUDA(ctx, cudaMemcpy(ctx->cat[catalog], cols, bytes, cudaMemcpyHostToDevice));
  return RB_OK;
}

// ---------------------------------------------------------------------------------------------------
static int check_alpha_desc(rb_context* ctx, const rb_alpha_desc* d, const double* out_total) {
  if (!ctx) return RB_ERR_INVALID;
  if (!d || !out_total) return rb_fail(ctx, RB_ERR_INVALID, "alpha: null descriptor / output");
  if (d->n_layers <= 0 || d->n_freqs <= 0) return rb_fail(ctx, RB_ERR_INVALID, "alpha: n_layers and n_freqs must be positive");
  if (!d->freqs || !d->T || !d->P) return rb_fail(ctx, RB_ERR_INVALID, "alpha: freqs / T / P must not be null");
  if (d->n_constituents <= 0 || d->n_constituents > RB_MAX_CONSTITUENTS)
    return rb_fail(ctx, RB_ERR_INVALID, "alpha: n_constituents must be 1..%d", RB_MAX_CONSTITUENTS);
  if (d->units != RB_UNITS_INVCM && d->units != RB_UNITS_DBPERKM) return rb_fail(ctx, RB_ERR_INVALID, "alpha: bad units");
  return RB_OK;
}

int rb_alpha_layers_dev(rb_context* ctx, const rb_alpha_desc* d, double* out_total, double* out_cube) {
  RB_TRY(check_alpha_desc(ctx, d, out_total));
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  // the frequency-class scan needs the (tiny) frequency list on the host
  if (d->freqs_host) return rb_launch_alpha(ctx, d, d->freqs_host, out_total, out_cube);
  const size_t n_hf = (size_t)d->n_freqs * (d->freqs_per_layer ? (size_t)d->n_layers : 1);
  std::vector<double> hf(n_hf);
  RB_CUDA(ctx, cudaMemcpyAsync(hf.data(), d->freqs, sizeof(double) * n_hf, cudaMemcpyDeviceToHost, ctx->stream));
  RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return rb_launch_alpha(ctx, d, hf.data(), out_total, out_cube);
}

// Host-pointer absorption call.  resident: the results stay in the context's RB_BUF_RES_* buffers, nothing is
// copied back and the stream is not synchronised (rb_alpha_layers_resident); else they land in out_total / out_cube.
static int alpha_layers_host(rb_context* ctx, const rb_alpha_desc* d, double* out_total, double* out_cube, bool resident,
                             bool keep_cube) {
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t L = d->n_layers, F = d->n_freqs, C = d->n_constituents;
  rb_alpha_desc dd = *d;
  void *p_f, *p_T, *p_P, *p_gas = nullptr, *p_cloud = nullptr, *p_scale = nullptr, *p_tot, *p_cube = nullptr;
  const bool want_cube = resident ? keep_cube : (out_cube != nullptr);
  const size_t n_f = F * (d->freqs_per_layer ? L : 1);       // per-layer frequency lists: [L][F]
  RB_TRY(rb_ensure(ctx, RB_BUF_FREQS, n_f * 8, &p_f));
  RB_TRY(rb_ensure(ctx, RB_BUF_T, L * 8, &p_T));
  RB_TRY(rb_ensure(ctx, RB_BUF_P, L * 8, &p_P));
  RB_TRY(rb_ensure(ctx, resident ? RB_BUF_RES_TOTAL : RB_BUF_TOTAL, L * F * 8, &p_tot));
  cudaStream_t s = ctx->stream;
  RB_CUDA(ctx, cudaMemcpyAsync(p_f, d->freqs, n_f * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_T, d->T, L * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_P, d->P, L * 8, cudaMemcpyHostToDevice, s));
  if (d->gas && d->gas_rows > 0) {
    RB_TRY(rb_ensure(ctx, RB_BUF_GAS, (size_t)d->gas_rows * L * 8, &p_gas));
    RB_CUDA(ctx, cudaMemcpyAsync(p_gas, d->gas, (size_t)d->gas_rows * L * 8, cudaMemcpyHostToDevice, s));
  }
  if (d->cloud && d->cloud_rows > 0) {
    RB_TRY(rb_ensure(ctx, RB_BUF_CLOUD, (size_t)d->cloud_rows * L * 8, &p_cloud));
    RB_CUDA(ctx, cudaMemcpyAsync(p_cloud, d->cloud, (size_t)d->cloud_rows * L * 8, cudaMemcpyHostToDevice, s));
  }
  if (d->scale) {
    RB_TRY(rb_ensure(ctx, RB_BUF_SCALE, C * L * 8, &p_scale));
    RB_CUDA(ctx, cudaMemcpyAsync(p_scale, d->scale, C * L * 8, cudaMemcpyHostToDevice, s));
  }
  if (want_cube) RB_TRY(rb_ensure(ctx, resident ? RB_BUF_RES_CUBE : RB_BUF_CUBE, L * F * C * 8, &p_cube));
  dd.freqs = (const double*)p_f; dd.T = (const double*)p_T; dd.P = (const double*)p_P;
  dd.gas = (const double*)p_gas; dd.cloud = (const double*)p_cloud; dd.scale = (const double*)p_scale;
  if (resident) {                       // whatever happens below, the old contents are gone
    ctx->res.slab_gen = 0;
    if (want_cube) ctx->res.cube_gen = 0;
  }
  RB_TRY(rb_launch_alpha(ctx, &dd, d->freqs, (double*)p_tot, (double*)p_cube));
  if (resident) {
    ctx->res.L = (int)L; ctx->res.F = (int)F;
    ctx->res.slab_gen = ++ctx->res.counter;
    if (want_cube) {
      ctx->res.cL = (int)L; ctx->res.cF = (int)F; ctx->res.cC = (int)C;
      ctx->res.cube_gen = ++ctx->res.counter;
    }
    return RB_OK;
  }
  RB_CUDA(ctx, cudaMemcpyAsync(out_total, p_tot, L * F * 8, cudaMemcpyDeviceToHost, s));
  if (out_cube) RB_CUDA(ctx, cudaMemcpyAsync(out_cube, p_cube, L * F * C * 8, cudaMemcpyDeviceToHost, s));
  RB_CUDA(ctx, cudaStreamSynchronize(s));
  return RB_OK;
}

int rb_alpha_layers(rb_context* ctx, const rb_alpha_desc* d, double* out_total, double* out_cube) {
  RB_TRY(check_alpha_desc(ctx, d, out_total));
  return alpha_layers_host(ctx, d, out_total, out_cube, false, false);
}

int rb_alpha_layers_dev_scatter(rb_context* ctx, const rb_alpha_desc* d, int32_t n_peers, const uint64_t* peer_slabs,
                                int64_t first_row) {
  static const double dummy = 0.0;
  RB_TRY(check_alpha_desc(ctx, d, &dummy));
  if (n_peers < 1 || n_peers > RB_MAX_PEERS || !peer_slabs || first_row < 0)
    return rb_fail(ctx, RB_ERR_INVALID, "alpha scatter: 1..%d peer slabs, first_row >= 0", RB_MAX_PEERS);
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  double* peers[RB_MAX_PEERS];
  for (int p = 0; p < n_peers; ++p) {
    if (!peer_slabs[p]) return rb_fail(ctx, RB_ERR_INVALID, "alpha scatter: null peer slab %d", p);
    peers[p] = (double*)(uintptr_t)peer_slabs[p];
  }
  if (d->freqs_host) return rb_launch_alpha(ctx, d, d->freqs_host, nullptr, nullptr, n_peers, peers, first_row);
  const size_t n_hf = (size_t)d->n_freqs * (d->freqs_per_layer ? (size_t)d->n_layers : 1);
  std::vector<double> hf(n_hf);
  RB_CUDA(ctx, cudaMemcpyAsync(hf.data(), d->freqs, sizeof(double) * n_hf, cudaMemcpyDeviceToHost, ctx->stream));
  RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return rb_launch_alpha(ctx, d, hf.data(), nullptr, nullptr, n_peers, peers, first_row);
}

int rb_alpha_layers_resident(rb_context* ctx, const rb_alpha_desc* d, int32_t keep_cube, uint64_t* out_slab_generation,
                             uint64_t* out_cube_generation) {
  static const double dummy = 0.0;
  RB_TRY(check_alpha_desc(ctx, d, &dummy));
  RB_TRY(alpha_layers_host(ctx, d, nullptr, nullptr, true, keep_cube != 0));
  if (out_slab_generation) *out_slab_generation = ctx->res.slab_gen;
  if (out_cube_generation) *out_cube_generation = keep_cube ? ctx->res.cube_gen : 0;
  return RB_OK;
}

int rb_alpha_rescale_resident(rb_context* ctx, const double* scale, uint64_t* out_slab_generation) {
  if (!ctx) return RB_ERR_INVALID;
  if (!ctx->res.cube_gen) return rb_fail(ctx, RB_ERR_INVALID, "rescale_resident: no per-constituent cube is resident");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = ctx->res.cL, F = ctx->res.cF, C = ctx->res.cC;
  void *p_tot, *p_scale = nullptr;
  RB_TRY(rb_ensure(ctx, RB_BUF_RES_TOTAL, (size_t)L * F * 8, &p_tot));
  ctx->res.slab_gen = 0;
  if (scale) {
    RB_TRY(rb_ensure(ctx, RB_BUF_SCALE, (size_t)C * L * 8, &p_scale));
    RB_CUDA(ctx, cudaMemcpyAsync(p_scale, scale, (size_t)C * L * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  RB_TRY(rb_launch_alpha_scale_sum(ctx, (const double*)ctx->buf[RB_BUF_RES_CUBE].p, (const double*)p_scale, L, F, C,
                                   (double*)p_tot, nullptr));
  ctx->res.L = L; ctx->res.F = F;
  ctx->res.slab_gen = ++ctx->res.counter;
  if (out_slab_generation) *out_slab_generation = ctx->res.slab_gen;
  return RB_OK;
}

int rb_alpha_resident_info(rb_context* ctx, int32_t* slab_shape, uint64_t* slab_generation, int32_t* cube_shape,
                           uint64_t* cube_generation) {
  if (!ctx) return RB_ERR_INVALID;
  if (slab_shape) { slab_shape[0] = ctx->res.L; slab_shape[1] = ctx->res.F; }
  if (slab_generation) *slab_generation = ctx->res.slab_gen;
  if (cube_shape) { cube_shape[0] = ctx->res.cL; cube_shape[1] = ctx->res.cF; cube_shape[2] = ctx->res.cC; }
  if (cube_generation) *cube_generation = ctx->res.cube_gen;
  return RB_OK;
}

int rb_alpha_fetch(rb_context* ctx, double* out_total, double* out_cube) {
  if (!ctx) return RB_ERR_INVALID;
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (out_total) {
    if (!ctx->res.slab_gen) return rb_fail(ctx, RB_ERR_INVALID, "alpha_fetch: no slab is resident");
    RB_CUDA(ctx, cudaMemcpyAsync(out_total, ctx->buf[RB_BUF_RES_TOTAL].p, (size_t)ctx->res.L * ctx->res.F * 8,
                                 cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (out_cube) {
    if (!ctx->res.cube_gen) return rb_fail(ctx, RB_ERR_INVALID, "alpha_fetch: no cube is resident");
    RB_CUDA(ctx, cudaMemcpyAsync(out_cube, ctx->buf[RB_BUF_RES_CUBE].p,
                                 (size_t)ctx->res.cL * ctx->res.cF * ctx->res.cC * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RB_OK;
}

int rb_alpha_scale_sum(rb_context* ctx, int32_t L, int32_t F, int32_t C, const double* cube, const double* scale,
                       double* out_total, double* out_cube) {
  if (!ctx) return RB_ERR_INVALID;
  if (!cube || !out_total || L <= 0 || F <= 0 || C <= 0) return rb_fail(ctx, RB_ERR_INVALID, "scale_sum: bad arguments");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)L * F;
  void *p_cube, *p_scale = nullptr, *p_tot;
  RB_TRY(rb_ensure(ctx, RB_BUF_CUBE, n * C * 8, &p_cube));
  RB_TRY(rb_ensure(ctx, RB_BUF_TOTAL, n * 8, &p_tot));
  cudaStream_t s = ctx->stream;
  RB_CUDA(ctx, cudaMemcpyAsync(p_cube, cube, n * C * 8, cudaMemcpyHostToDevice, s));
  if (scale) {
    RB_TRY(rb_ensure(ctx, RB_BUF_SCALE, (size_t)C * L * 8, &p_scale));
    RB_CUDA(ctx, cudaMemcpyAsync(p_scale, scale, (size_t)C * L * 8, cudaMemcpyHostToDevice, s));
  }
  RB_TRY(rb_launch_alpha_scale_sum(ctx, (const double*)p_cube, (const double*)p_scale, L, F, C, (double*)p_tot,
                                   out_cube ? (double*)p_cube : nullptr));
  RB_CUDA(ctx, cudaMemcpyAsync(out_total, p_tot, n * 8, cudaMemcpyDeviceToHost, s));
  if (out_cube) RB_CUDA(ctx, cudaMemcpyAsync(out_cube, p_cube, n * C * 8, cudaMemcpyDeviceToHost, s));
  RB_CUDA(ctx, cudaStreamSynchronize(s));
  return RB_OK;
}

// ---------------------------------------------------------------------------------------------------
static int make_geometry(rb_context* ctx, const rb_geometry_desc* g, int64_t R, RtLaunch* out) {
  if (!g) return rb_fail(ctx, RB_ERR_INVALID, "geometry: null descriptor");
  if (g->n_layers < 2) return rb_fail(ctx, RB_ERR_INVALID, "geometry: need at least 2 layers");
  if (R <= 0) return rb_fail(ctx, RB_ERR_INVALID, "geometry: n_rays must be positive");
  if (!(g->Req > 0.0) || !(g->Rpol > 0.0)) return rb_fail(ctx, RB_ERR_INVALID, "geometry: Req / Rpol must be positive");
  if (g->gtype != RB_GTYPE_ELLIPSE && g->gtype != RB_GTYPE_SPHERE && g->gtype != RB_GTYPE_GRAVITY)
    return rb_fail(ctx, RB_ERR_UNSUPPORTED, "geometry: gtype %d not built (ellipse / sphere / gravity only)", g->gtype);
  if (g->gtype == RB_GTYPE_GRAVITY && (!ctx->grav.rmag || ctx->grav.L != g->n_layers))
    return rb_fail(ctx, RB_ERR_INVALID, "geometry: gtype 'gravity' needs rb_set_gravity_model for this %d-layer profile",
                   g->n_layers);
  out->gtype = g->gtype;
  out->L = g->n_layers;
  out->n0 = g->n0; out->n1 = g->n1;
  out->q = (g->gtype == RB_GTYPE_ELLIPSE) ? g->Rpol / g->Req : 1.0;
  // computeAspect (raypath.py:39-44): f = 1 - Rpol/Req is used for every gtype
  const double f = 1.0 - g->Rpol / g->Req;
  const double tip = -g->orientation[0] * M_PI / 180.0;
  const double rotate = -atan(tan(g->orientation[1] * M_PI / 180.0) * (1.0 - f) * (1.0 - f));
  out->rot[0] = cos(tip); out->rot[1] = sin(tip); out->rot[2] = cos(rotate); out->rot[3] = sin(rotate);
  out->limb = g->limb;
  out->R = R;
  out->Rpad = (R + 31) & ~(int64_t)31;
  return RB_OK;
}

static void aspect_out(const rb_geometry_desc* g, double rNorm, double* out) {
  const double f = 1.0 - g->Rpol / g->Req;
  out[0] = -g->orientation[0] * M_PI / 180.0;
  out[1] = -atan(tan(g->orientation[1] * M_PI / 180.0) * (1.0 - f) * (1.0 - f));
  out[2] = rNorm;
}

// ---- prefetched geometry (rb_geometry_prefetch[_dev]) ---------------------------------------------------
static bool same_geometry(const rb_geometry_desc& a, const rb_geometry_desc& b) {
  return a.n_layers == b.n_layers && a.radius == b.radius && a.n0 == b.n0 && a.n1 == b.n1 && a.Req == b.Req &&
         a.Rpol == b.Rpol && a.orientation[0] == b.orientation[0] && a.orientation[1] == b.orientation[1] &&
         a.gtype == b.gtype && a.limb == b.limb;
}

// true (and the ticket is consumed, ctx->stream made to wait for it) when the prefetched geometry is the one
// this call needs; any other state drops the ticket: the caller is about to overwrite the ds buffers
// (streamed: a hit on a trace that publishes its progress is NOT waited for here -- the caller decides with
//  join_streamed_ticket whether its integration can follow the running trace)
static bool take_ticket(rb_context* ctx, const rb_geometry_desc* g, int64_t R, const void* b, bool* streamed = nullptr) {
  auto& t = ctx->ticket;
  if (streamed) *streamed = false;
  if (!t.valid) return false;
  const bool hit = g && t.R == R && t.b == b && same_geometry(t.g, *g);
  t.valid = false;
  if (hit && t.streamed && streamed) { *streamed = true; return true; }
  cudaStreamWaitEvent(ctx->stream, t.done, 0);   // hit: the geometry is ready; miss: it no longer writes the buffers
  return hit;
}
static void drop_ticket(rb_context* ctx) { take_ticket(ctx, nullptr, -1, nullptr); }

typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWaitValue32Fn stream_wait_value();
// follow == true: ctx->stream continues once the list of hitting rays exists and every CTA of the trace has started
// (the integration then waits chunk by chunk on the device: RtLaunch::prog); false: once the trace has ended
static int join_streamed_ticket(rb_context* ctx, bool follow) {
  auto& t = ctx->ticket;
  if (!follow) { RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, t.done, 0)); return RB_OK; }
  RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, t.mid, 0));
  if (stream_wait_value()((CUstream)ctx->stream, (CUdeviceptr)(uintptr_t)t.started, t.started_target,
                          CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
    return rb_fail(ctx, RB_ERR_CUDA, "rt: cuStreamWaitValue32 failed");
  return RB_OK;
}

// The ds buffer of a ray request: the tiled FP64 slab, followed -- when the context integrates in mixed
// precision -- by the float slab of the same tiling (each with its over-read slack).
static size_t ds_bytes(const rb_context* ctx, size_t S, int64_t Rpad) {
  const size_t one = S * (size_t)Rpad * 8 + kRtSlackBytes;
  return ctx->rt_precision == RB_RT_MIXED ? one + S * (size_t)Rpad * 4 + kRtSlackBytes : one;
}
static void bind_ds(const rb_context* ctx, RtLaunch& L, void* p_ds, void* p_n) {
  const size_t S = L.L - 1;
  L.ds = (double*)p_ds;
  L.dsf = ctx->rt_precision == RB_RT_MIXED ? (float*)((char*)p_ds + S * (size_t)L.Rpad * 8 + kRtSlackBytes) : nullptr;
  L.nseg = (int32_t*)p_n;
  L.nanflag = (int32_t*)p_n + L.Rpad;
}

// Requests that go to the FP64 rays-major integration (>= 512 point rays; rb_rt_prepare) trace and integrate
// only the rays that hit the planet, as full tiles of a compacted list (rt_kernels.cu: ray_edge_kernel).
// rb_set_rt_tuning / RB_RT_COMPACT=0 switch it off (A/B measurements, tests of the plain path).
static bool want_compact(const rb_context* ctx, int64_t R) {
  return ctx->rt_compact == 1 && R >= 512 && R < 2000000000LL && ctx->rt_precision == RB_RT_F64;
}
// RB_RT_STREAM_GEOMETRY=0/1 (default: requests of at most kStreamMaxRays rays)
constexpr int64_t kStreamMaxRays = 200000;
static bool want_stream(rb_context* ctx, int64_t R) {
  if (ctx->rt_stream < 0) {
    const char* e = getenv("RB_RT_STREAM_GEOMETRY");
    ctx->rt_stream = e ? (atoi(e) ? 1 : 0) : 2;
  }
  if (ctx->rt_precision != RB_RT_F64 || ctx->rt_tiles) return false;
  return ctx->rt_stream == 1 || (ctx->rt_stream == 2 && R <= kStreamMaxRays);
}
static int bind_compact(rb_context* ctx, RtLaunch& L) {
  L.compact = want_compact(ctx, L.R) && L.gtype != RB_GTYPE_GRAVITY;   // the gravity march walks the plain ray order
  if (L.gtype == RB_GTYPE_GRAVITY) L.dsf = nullptr;                    // ... and feeds the FP64 integration only
  L.cidx = nullptr; L.ncomp = nullptr; L.zq = nullptr; L.blkcnt = nullptr;
  if (!L.compact) return RB_OK;
  void *p_c, *p_z, *p_k;
  RB_TRY(rb_ensure(ctx, RB_BUF_CIDX, ((size_t)L.Rpad + 32) * 4, &p_c));
  RB_TRY(rb_ensure(ctx, RB_BUF_ZQ, (size_t)L.Rpad * 8, &p_z));
  RB_TRY(rb_ensure(ctx, RB_BUF_BLKCNT, ((size_t)L.Rpad / 256 + 2) * 4, &p_k));
  L.cidx = (int32_t*)p_c;
  L.ncomp = (int32_t*)p_c + L.Rpad;
  L.zq = (double*)p_z;
  L.blkcnt = (int32_t*)p_k;
  return RB_OK;
}

static int geometry_prefetch(rb_context* ctx, const rb_geometry_desc* g, int64_t R, const double* b, bool host) {
  if (!ctx) return RB_ERR_INVALID;
  if (!b || !g || !g->radius) return rb_fail(ctx, RB_ERR_INVALID, "geometry_prefetch: null pointer");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  RtLaunch L{};
  RB_TRY(make_geometry(ctx, g, R, &L));
  const size_t S = L.L - 1;
  auto& t = ctx->ticket;
  drop_ticket(ctx);
  void *p_rad = nullptr, *p_b = nullptr, *p_ds, *p_n;
  if (host) {
    RB_TRY(rb_ensure(ctx, RB_BUF_RADIUS, L.L * 8, &p_rad));
    RB_TRY(rb_ensure(ctx, RB_BUF_B, (size_t)R * 16, &p_b));
  }
  RB_TRY(rb_ensure(ctx, RB_BUF_DS, ds_bytes(ctx, S, L.Rpad), &p_ds));
  RB_TRY(rb_ensure(ctx, RB_BUF_NSEG, (size_t)L.Rpad * 8, &p_n));
  if (!ctx->aux[0]) {
    // highest priority: the few, long-running geometry CTAs get their SM slots ahead of the absorption
    // kernel they overlap with (which has thousands of short CTAs to fill in around them)
    int lo = 0, hi = 0;
    RB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    RB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->aux[0], cudaStreamNonBlocking, hi));
  }
  if (!t.done) RB_CUDA(ctx, cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
  cudaStream_t user = ctx->stream, sG = ctx->aux[0];
  // fork: everything already enqueued on the context stream (the producers of b, earlier users of the
  // buffers) comes first
  RB_CUDA(ctx, cudaEventRecord(t.done, user));
  RB_CUDA(ctx, cudaStreamWaitEvent(sG, t.done, 0));
  if (host) {
    RB_CUDA(ctx, cudaMemcpyAsync(p_rad, g->radius, L.L * 8, cudaMemcpyHostToDevice, sG));
    RB_CUDA(ctx, cudaMemcpyAsync(p_b, b, (size_t)R * 16, cudaMemcpyHostToDevice, sG));
    L.radius = (const double*)p_rad; L.b = (const double*)p_b;
  } else {
    L.radius = g->radius; L.b = b;
  }
  bind_ds(ctx, L, p_ds, p_n);
  RB_TRY(bind_compact(ctx, L));
  // Small requests (a rank's rows of an image shared by 4 or 8 GPUs): the trace is a chain of ~1000 dependent steps
  // per ray whatever the number of rays, so let the integration start behind it (RtLaunch::prog)
  t.streamed = false;
  if (L.compact && want_stream(ctx, R) && stream_wait_value()) {
    const size_t npub = (S + kGeoPub - 1) / kGeoPub, nprog = (size_t)(L.Rpad / 32) * npub;
    void* p_g;
    RB_TRY(rb_ensure(ctx, RB_BUF_PROG, (nprog + 1) * sizeof(int), &p_g));
    if (!t.mid) RB_CUDA(ctx, cudaEventCreateWithFlags(&t.mid, cudaEventDisableTiming));
    L.prog = (int32_t*)p_g;
    L.mid_event = t.mid;
    t.streamed = true;
    t.prog = L.prog;
    t.started = L.prog + nprog;
    t.started_target = (unsigned)((R + 127) / 128);          // CTAs of ray_geometry_kernel
  }
  ctx->stream = sG;
  const int status = rb_launch_geometry(ctx, L);
  ctx->stream = user;
  RB_TRY(status);
  RB_CUDA(ctx, cudaEventRecord(t.done, sG));
  t.valid = true; t.R = R; t.b = b; t.g = *g;
  return RB_OK;
}

int rb_geometry_prefetch(rb_context* ctx, const rb_geometry_desc* g, int64_t R, const double* b) {
  return geometry_prefetch(ctx, g, R, b, true);
}
int rb_geometry_prefetch_dev(rb_context* ctx, const rb_geometry_desc* g, int64_t R, const double* b) {
  return geometry_prefetch(ctx, g, R, b, false);
}

int rb_compute_ds(rb_context* ctx, const rb_geometry_desc* g, int64_t R, const double* b, double* out_ds,
                  int32_t* out_nseg, double* out_aspect) {
  if (!ctx) return RB_ERR_INVALID;
  if (!b || !out_ds || !out_nseg || !g || !g->radius) return rb_fail(ctx, RB_ERR_INVALID, "compute_ds: null pointer");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  RtLaunch L{};
  RB_TRY(make_geometry(ctx, g, R, &L));
  drop_ticket(ctx);                                          // a prefetched geometry shares these buffers
  const size_t S = L.L - 1;
  void *p_rad, *p_b, *p_ds, *p_n, *p_o;
  RB_TRY(rb_ensure(ctx, RB_BUF_RADIUS, L.L * 8, &p_rad));
  RB_TRY(rb_ensure(ctx, RB_BUF_B, (size_t)R * 16, &p_b));
  RB_TRY(rb_ensure(ctx, RB_BUF_DS, S * L.Rpad * 8 + kRtSlackBytes, &p_ds));
  RB_TRY(rb_ensure(ctx, RB_BUF_NSEG, (size_t)L.Rpad * 8, &p_n));
  RB_TRY(rb_ensure(ctx, RB_BUF_MISC, (size_t)R * S * 8, &p_o));
  cudaStream_t s = ctx->stream;
  RB_CUDA(ctx, cudaMemcpyAsync(p_rad, g->radius, L.L * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_b, b, (size_t)R * 16, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemsetAsync(p_ds, 0, S * L.Rpad * 8, s));
  L.radius = (const double*)p_rad; L.b = (const double*)p_b; L.ds = (double*)p_ds; L.nseg = (int32_t*)p_n; L.nanflag = (int32_t*)p_n + L.Rpad;
  RB_TRY(rb_launch_geometry(ctx, L));
  RB_TRY(rb_launch_ds_transpose(ctx, L, (double*)p_o));
  RB_CUDA(ctx, cudaMemcpyAsync(out_ds, p_o, (size_t)R * S * 8, cudaMemcpyDeviceToHost, s));
  RB_CUDA(ctx, cudaMemcpyAsync(out_nseg, p_n, (size_t)R * 4, cudaMemcpyDeviceToHost, s));
  RB_CUDA(ctx, cudaStreamSynchronize(s));
  if (out_aspect) aspect_out(g, g->radius[0], out_aspect);
  return RB_OK;
}

int rb_set_gravity_model(rb_context* ctx, const rb_gravity_model* m) {
  if (!ctx) return RB_ERR_INVALID;
  if (!m || m->n_layers < 2 || !m->radius || !m->GM_layer || m->n_J < 3 || !m->Jn || m->n_vw < 2 || !m->vwlat || !m->vwdat)
    return rb_fail(ctx, RB_ERR_INVALID, "gravity model: null pointer / too few entries");
  if (!(m->latstep > 0.0) || !(m->max_lat > 0.0) || m->max_lat > 90.0 || !(m->RJ > 0.0))
    return rb_fail(ctx, RB_ERR_INVALID, "gravity model: latstep, max_lat (<= 90) and RJ must be positive");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = m->n_layers;
  const int K = (int)ceil((m->max_lat + m->latstep) / m->latstep);
  RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->grav.rmag) { cudaFree(ctx->grav.rmag); ctx->grav.rmag = nullptr; }
  if (ctx->grav.gamma) { cudaFree(ctx->grav.gamma); ctx->grav.gamma = nullptr; }
  ctx->grav.L = 0;
  const size_t nent = 2 * (size_t)K * L;
  if (cudaMalloc(&ctx->grav.rmag, nent * 8) != cudaSuccess || cudaMalloc(&ctx->grav.gamma, nent * 8) != cudaSuccess) {
    cudaGetLastError();
    return rb_fail(ctx, RB_ERR_NOMEM, "gravity model: %zu MB for the shape table", 2 * nent * 8 >> 20);
  }
  // inputs: radius | GM | Jn | vwlat | vwdat in one scratch buffer
  const size_t nin = 2 * (size_t)L + m->n_J + 2 * (size_t)m->n_vw;
  void* p;
  RB_TRY(rb_ensure(ctx, RB_BUF_MISC, nin * 8, &p));
  double* d = (double*)p;
  double *d_rad = d, *d_GM = d + L, *d_J = d + 2 * L, *d_vl = d_J + m->n_J, *d_vd = d_vl + m->n_vw;
  cudaStream_t s = ctx->stream;
  RB_CUDA(ctx, cudaMemcpyAsync(d_rad, m->radius, (size_t)L * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(d_GM, m->GM_layer, (size_t)L * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(d_J, m->Jn, (size_t)m->n_J * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(d_vl, m->vwlat, (size_t)m->n_vw * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(d_vd, m->vwdat, (size_t)m->n_vw * 8, cudaMemcpyHostToDevice, s));
  ctx->grav.L = L; ctx->grav.K = K; ctx->grav.latstep = m->latstep;
  ctx->grav.r0 = m->radius[0]; ctx->grav.rlast = m->radius[L - 1];
  RB_TRY(rb_build_geoid_table(ctx, L, K, m->n_J, m->n_vw, d_rad, d_GM, d_J, d_vl, d_vd, m->RJ, m->omega_m, m->latstep));
  RB_CUDA(ctx, cudaStreamSynchronize(s));      // the scratch inputs may be reused by the next call
  return RB_OK;
}

int rb_compute_ray_fields(rb_context* ctx, const rb_geometry_desc* g, int64_t R, const double* b, double* out_fields) {
  if (!ctx) return RB_ERR_INVALID;
  if (!b || !out_fields || !g || !g->radius) return rb_fail(ctx, RB_ERR_INVALID, "ray_fields: null pointer");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  RtLaunch L{};
  RB_TRY(make_geometry(ctx, g, R, &L));
  drop_ticket(ctx);
  const size_t S = L.L - 1;
  void *p_rad, *p_b, *p_ds, *p_n, *p_o;
  RB_TRY(rb_ensure(ctx, RB_BUF_RADIUS, L.L * 8, &p_rad));
  RB_TRY(rb_ensure(ctx, RB_BUF_B, (size_t)R * 16, &p_b));
  RB_TRY(rb_ensure(ctx, RB_BUF_DS, S * L.Rpad * 8 + kRtSlackBytes, &p_ds));
  RB_TRY(rb_ensure(ctx, RB_BUF_NSEG, (size_t)L.Rpad * 8, &p_n));
  RB_TRY(rb_ensure(ctx, RB_BUF_MISC, (size_t)R * 3 * S * 8, &p_o));
  cudaStream_t s = ctx->stream;
  RB_CUDA(ctx, cudaMemcpyAsync(p_rad, g->radius, L.L * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_b, b, (size_t)R * 16, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemsetAsync(p_o, 0, (size_t)R * 3 * S * 8, s));
  L.radius = (const double*)p_rad; L.b = (const double*)p_b; L.ds = (double*)p_ds; L.nseg = (int32_t*)p_n; L.nanflag = (int32_t*)p_n + L.Rpad;
  if (L.gtype == RB_GTYPE_GRAVITY) {
    RB_TRY(rb_launch_gravity_geometry(ctx, L, (double*)p_o));   // the gravity march knows these fields itself
  } else {
    RB_TRY(rb_launch_geometry(ctx, L));
    RB_TRY(rb_launch_ray_fields(ctx, L, (double*)p_o));
  }
  RB_CUDA(ctx, cudaMemcpyAsync(out_fields, p_o, (size_t)R * 3 * S * 8, cudaMemcpyDeviceToHost, s));
  RB_CUDA(ctx, cudaStreamSynchronize(s));
  return RB_OK;
}

static int check_rt(rb_context* ctx, const rb_rt_desc* rt, const void* out_Tb, bool allow_alpha0 = false) {
  if (!rt || !out_Tb) return rb_fail(ctx, RB_ERR_INVALID, "rt: null descriptor / output");
  if (rt->n_freqs <= 0) return rb_fail(ctx, RB_ERR_INVALID, "rt: n_freqs must be positive");
  if (!rt->alpha || !rt->T) return rb_fail(ctx, RB_ERR_INVALID, "rt: alpha / T must not be null");
  if (rt->alpha0 && !allow_alpha0)
    return rb_fail(ctx, RB_ERR_UNSUPPORTED, "rt: alpha0 (Doppler form: one absorption slab pair per ray) goes through rb_rt_integrate");
  return RB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Result pipeline.  Geometry runs once for all rays; for HOST outputs the integration is cut into ray
// chunks (multiples of 32 rays = whole ds tiles) on a second stream so that the device->host copy of chunk
// c (third stream) overlaps the integration of chunk c+1.  Both streams are forked from / joined to
// ctx->stream with events, so callers still see one in-order stream.  (Overlapping geometry(c+1) with
// integrate(c) was measured and is slower: the latency-bound geometry kernel starves when it shares SMs.)
// All pointers are DEVICE pointers except h_out / h_intW.
// RB_TRACE=1: host-clock trace of the stages of rb_rt_batch (stderr), for tools/e2e_breakdown.py
static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
static bool tracing() { static int t = -1; if (t < 0) t = getenv("RB_TRACE") ? 1 : 0; return t == 1; }
#define RB_TRACE_AT(label) do { if (tracing()) fprintf(stderr, "[rb_trace] %-28s %.3f ms\n", label, now_ms() - trace_t0); } while (0)

// cuStreamWaitValue32 through the runtime's driver entry point lookup (the library links no libcuda)
static StreamWaitValue32Fn stream_wait_value() {
  static StreamWaitValue32Fn fn = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (StreamWaitValue32Fn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

static int pipe_setup(rb_context* ctx, int nch) {
  for (int i = 1; i < 3; ++i)   // aux[0] is the (high-priority) geometry prefetch stream
    if (!ctx->aux[i]) RB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
  const size_t need = 4 + 2 * (size_t)nch;
  while (ctx->pipe_ev.size() < need) {
    cudaEvent_t e;
    RB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->pipe_ev.push_back(e);
  }
  return RB_OK;
}

static int choose_chunks(const rb_context* ctx, int64_t R, bool rays_path, bool host_out) {
  if (!rays_path || !host_out || R < 16384) return 1;
  int n = ctx->rt_chunks;
  if (n <= 0) {
    // Planet.run of C4 on one GPU (361 201 rays, 93 MB out): 4 / 6 / 8 / 12 / 16 chunks 4.08 / 3.97 / 3.93 / 3.86 / 3.88 ms
    // (tools/sessions/gpu_session_r2aj.sh); a rank's share on 2-8 GPUs was tuned with 6
    const char* e = getenv("RB_RT_CHUNKS");
    n = e ? atoi(e) : (R >= 200000 ? 12 : 6);
  }
  if (n < 1) n = 1;
  if (n > 64) n = 64;
  return n;
}

static int run_rt_pipeline(rb_context* ctx, const RtLaunch& full_in, const rb_rt_desc* rd, void* d_out, double* d_intW,
                           void* h_out, double* h_intW, bool have_geometry, bool streamed) {
  RtLaunch full = full_in;
  const int64_t R = full.R;
  const size_t S = full.L - 1, F = rd->n_freqs, esz = rd->out_f32 ? 4 : 8;
  RtPrep prep;
  RB_TRY(rb_rt_prepare(ctx, full.L, rd, R, false, full.dsf != nullptr, &prep));
  // a prefetched trace that publishes its progress: the FP64 pair kernel can follow it while it runs; every other
  // consumer waits for its end
  const bool follow = streamed && full.compact && prep.use_rays && prep.pairs && !prep.tiles;
  if (streamed) RB_TRY(join_streamed_ticket(ctx, follow));
  if (follow) full.prog = ctx->ticket.prog;
  struct Rejoin {   // whatever follows on the context stream comes after the end of the trace
    rb_context* c; bool on;
    ~Rejoin() { if (on) cudaStreamWaitEvent(c->stream, c->ticket.done, 0); }
  } rejoin{ctx, follow};
  if (full.compact && !prep.use_rays) {
    // a disc-averaged request over >= 512 rays: the lanes = frequency kernel walks the plain ray order
    full.compact = false;
    have_geometry = false;
  }
  if (!have_geometry) RB_TRY(rb_launch_geometry(ctx, full));
  int nch = choose_chunks(ctx, R, prep.use_rays, h_out != nullptr);
  const bool one_launch = stream_wait_value() && nch <= kMaxProgressChunks && !getenv("RB_RT_SPLIT_LAUNCHES");
  if (full.compact && !one_launch) nch = 1;   // the per-chunk launches below cut the plain ray order
  if (nch == 1) {
    RB_TRY(rb_launch_integrate(ctx, full, rd, prep, nullptr, d_out, d_intW, -1, nullptr, nullptr, nullptr));
    if (h_out) RB_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, (size_t)R * F * esz, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_intW) RB_CUDA(ctx, cudaMemcpyAsync(h_intW, d_intW, (size_t)R * F * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return RB_OK;
  }
  RB_TRY(pipe_setup(ctx, nch));
  cudaStream_t user = ctx->stream, sI = ctx->aux[1], sC = ctx->aux[2];
  cudaEvent_t* ev = ctx->pipe_ev.data();
  const int64_t ntiles = (R + 31) / 32;
  auto cut_at = [&](int j) {
    // shrinking chunks: the copy of the last chunk is the only one nothing overlaps, so keep it small
    return (j >= nch) ? ntiles : (int64_t)llround((double)ntiles * (1.0 - pow(1.0 - (double)j / nch, 1.6)));
  };
  if (one_launch) {
    // One integration launch; its CTAs count themselves into per-chunk counters when their results are in
    // global memory, and the copy stream waits on each counter (stream memory operation) before moving
    // that chunk to the host: the copies overlap the same launch, no launch tails between chunks.
    void* p_flags;
    RB_TRY(rb_ensure(ctx, RB_BUF_FLAGS, kMaxProgressChunks * sizeof(unsigned), &p_flags));
    RtProgress pg;
    pg.nchunks = nch;
    pg.done = (unsigned*)p_flags;
    // CTAs run in launch order; a disc image ends with rows that miss the planet (no work), which would all
    // complete -- and queue for copy-out -- at the very end of the launch.  Start in the middle of the ray
    // list instead (the centre row of an image) and wrap around: the cheap tiles are done early and the
    // chunks complete evenly in time.
    pg.shift = rb_progress_shift(R);
    for (int c = 0; c <= nch; ++c) pg.cut[c] = (int)cut_at(c);
    const unsigned fgroups = (unsigned)prep.fgroups;
    // value the counter of a chunk of `len` ray tiles ends at: one count per CTA (len * fgroups); compacted launches
    // start the counters at their deficit so that they end at kProgressTarget (rt_progress_init_kernel)
    RB_CUDA(ctx, cudaMemsetAsync(p_flags, 0, kMaxProgressChunks * sizeof(unsigned), user));
    // compacted launch: the counters start at their deficits (how many CTAs will report into each chunk is only
    // known on the device)
    // (and the sky pixels are written before the copy stream may move anything: a chunk without a single hit is
    // complete from the start)
    if (full.compact) {
      RB_TRY(rb_launch_progress_init(ctx, full, pg, fgroups));
      RB_TRY(rb_launch_fill_miss(ctx, full, (int)F, d_out, d_intW, rd->out_f32));
    }
    RB_CUDA(ctx, cudaEventRecord(ev[0], user));             // counters are set; the copy stream may start waiting
    RB_CUDA(ctx, cudaStreamWaitEvent(sC, ev[0], 0));
    if (full.compact) RB_TRY(rb_join_fill_miss(ctx, sC));   // (the sky fill runs on its own stream)
    RB_TRY(rb_launch_integrate(ctx, full, rd, prep, &pg, d_out, d_intW, -1, nullptr, nullptr, nullptr));
    auto copy_tiles = [&](int64_t t0, int64_t t1) -> int {   // memory-order tiles [t0, t1)
      const int64_t r0 = t0 * 32, r1 = (t1 * 32 < R) ? t1 * 32 : R;
      if (r1 <= r0) return RB_OK;
      RB_CUDA(ctx, cudaMemcpyAsync((char*)h_out + (size_t)r0 * F * esz, (char*)d_out + (size_t)r0 * F * esz,
                                   (size_t)(r1 - r0) * F * esz, cudaMemcpyDeviceToHost, sC));
      if (h_intW)
        RB_CUDA(ctx, cudaMemcpyAsync(h_intW + (size_t)r0 * F, d_intW + (size_t)r0 * F, (size_t)(r1 - r0) * F * 8,
                                     cudaMemcpyDeviceToHost, sC));
      return RB_OK;
    };
    for (int c = 0; c < nch; ++c) {
      const int64_t p0 = pg.cut[c], p1 = pg.cut[c + 1];      // processing-order tiles of the chunk
      if (p1 <= p0) continue;
      const unsigned target = full.compact ? kProgressTarget : (unsigned)(p1 - p0) * fgroups;
      if (stream_wait_value()((CUstream)sC, (CUdeviceptr)(uintptr_t)(pg.done + c), target,
                              CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
        return rb_fail(ctx, RB_ERR_CUDA, "rt: cuStreamWaitValue32 failed");
      const int64_t m0 = (p0 + pg.shift) % ntiles, len = p1 - p0;
      if (m0 + len <= ntiles) {
        RB_TRY(copy_tiles(m0, m0 + len));
      } else {
        RB_TRY(copy_tiles(m0, ntiles));
        RB_TRY(copy_tiles(0, m0 + len - ntiles));
      }
    }
    RB_CUDA(ctx, cudaEventRecord(ev[3], sC));
    RB_CUDA(ctx, cudaStreamWaitEvent(user, ev[3], 0));
    return RB_OK;
  }
  RB_CUDA(ctx, cudaEventRecord(ev[0], user));               // inputs, operands and geometry are ready
  RB_CUDA(ctx, cudaStreamWaitEvent(sI, ev[0], 0));
  RB_CUDA(ctx, cudaStreamWaitEvent(sC, ev[0], 0));
  // fallback: one integration launch per chunk (chunk boundaries are cut by tiles of 32 rays)
  int status = RB_OK;
  ctx->stream = sI;
  for (int c = 0; c < nch && status == RB_OK; ++c) {
    const int64_t t0 = cut_at(c), t1 = cut_at(c + 1);
    const int64_t r0 = t0 * 32, r1 = (t1 * 32 < R) ? t1 * 32 : R;
    if (r1 <= r0) continue;
    RtLaunch Lc = full;
    Lc.R = r1 - r0;
    Lc.Rpad = (Lc.R + 31) & ~(int64_t)31;
    Lc.b = full.b + 2 * r0;
    Lc.ds = full.ds + (size_t)t0 * S * 32;
    if (full.dsf) Lc.dsf = full.dsf + (size_t)t0 * S * 32;
    Lc.nseg = full.nseg + r0;
    Lc.nanflag = full.nanflag + r0;
    char* oc = (char*)d_out + (size_t)r0 * F * esz;
    double* wc = d_intW ? d_intW + (size_t)r0 * F : nullptr;
    status = rb_launch_integrate(ctx, Lc, rd, prep, nullptr, oc, wc, -1, nullptr, nullptr, nullptr);
    if (status != RB_OK) break;
    cudaEventRecord(ev[4 + c], sI);
    cudaStreamWaitEvent(sC, ev[4 + c], 0);
    cudaMemcpyAsync((char*)h_out + (size_t)r0 * F * esz, oc, (size_t)(r1 - r0) * F * esz, cudaMemcpyDeviceToHost, sC);
    if (h_intW) cudaMemcpyAsync(h_intW + (size_t)r0 * F, wc, (size_t)(r1 - r0) * F * 8, cudaMemcpyDeviceToHost, sC);
  }
  ctx->stream = user;
  cudaEventRecord(ev[2], sI); cudaStreamWaitEvent(user, ev[2], 0);
  cudaEventRecord(ev[3], sC); cudaStreamWaitEvent(user, ev[3], 0);
  RB_TRY(status);
  RB_CUDA(ctx, cudaGetLastError());
  return RB_OK;
}

int rb_rt_batch_dev(rb_context* ctx, const rb_geometry_desc* g, const rb_rt_desc* rt, int64_t R, const double* b,
                    void* out_Tb, double* out_intW) {
  if (!ctx) return RB_ERR_INVALID;
  RB_TRY(check_rt(ctx, rt, out_Tb));
  if (!b || !g || !g->radius) return rb_fail(ctx, RB_ERR_INVALID, "rt: null pointer");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  RtLaunch L{};
  RB_TRY(make_geometry(ctx, g, R, &L));
  const size_t S = L.L - 1;
  void *p_ds, *p_n;
  RB_TRY(rb_ensure(ctx, RB_BUF_DS, ds_bytes(ctx, S, L.Rpad), &p_ds));
  RB_TRY(rb_ensure(ctx, RB_BUF_NSEG, (size_t)L.Rpad * 8, &p_n));
  L.radius = g->radius; L.b = b;
  bind_ds(ctx, L, p_ds, p_n);
  RB_TRY(bind_compact(ctx, L));
  bool streamed;
  const bool have_geometry = take_ticket(ctx, g, R, b, &streamed);
  return run_rt_pipeline(ctx, L, rt, out_Tb, out_intW, nullptr, nullptr, have_geometry, streamed);
}

static int rt_batch_host(rb_context* ctx, const rb_geometry_desc* g, const rb_rt_desc* rt, int64_t R, const double* b,
                         void* out_Tb, double* out_intW, int64_t profile_ray, double* out_tau, double* out_W,
                         double* out_Tblyr, bool resident_alpha) {
  if (!ctx) return RB_ERR_INVALID;
  if (resident_alpha) {
    if (!rt || !out_Tb || rt->n_freqs <= 0 || !rt->T) return rb_fail(ctx, RB_ERR_INVALID, "rt: null descriptor / output");
    if (!ctx->res.slab_gen) return rb_fail(ctx, RB_ERR_INVALID, "rt: no absorption slab is resident");
    if (ctx->res.F != rt->n_freqs || !g || ctx->res.L != g->n_layers)
      return rb_fail(ctx, RB_ERR_INVALID, "rt: the resident slab is [%d][%d], the request needs [%d][%d]", ctx->res.L,
                     ctx->res.F, g ? g->n_layers : -1, rt->n_freqs);
  } else {
    RB_TRY(check_rt(ctx, rt, out_Tb));
  }
  if (!b || !g || !g->radius) return rb_fail(ctx, RB_ERR_INVALID, "rt: null pointer");
  if (profile_ray >= 0 && (!out_tau || !out_W || !out_Tblyr)) return rb_fail(ctx, RB_ERR_INVALID, "rt: profile outputs null");
  if (profile_ray >= R) return rb_fail(ctx, RB_ERR_INVALID, "rt: profile_ray out of range");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  RtLaunch L{};
  RB_TRY(make_geometry(ctx, g, R, &L));
  const size_t nL = L.L, S = nL - 1, F = rt->n_freqs;
  const size_t esz = rt->out_f32 ? 4 : 8;
  void *p_rad, *p_b, *p_ds, *p_n, *p_alpha, *p_T, *p_tb, *p_iw = nullptr, *p_prof = nullptr;
  RB_TRY(rb_ensure(ctx, RB_BUF_RADIUS, nL * 8, &p_rad));
  RB_TRY(rb_ensure(ctx, RB_BUF_B, (size_t)R * 16, &p_b));
  RB_TRY(rb_ensure(ctx, RB_BUF_DS, ds_bytes(ctx, S, L.Rpad), &p_ds));
  RB_TRY(rb_ensure(ctx, RB_BUF_NSEG, (size_t)L.Rpad * 8, &p_n));
  if (resident_alpha) p_alpha = ctx->buf[RB_BUF_RES_TOTAL].p;
  else RB_TRY(rb_ensure(ctx, RB_BUF_TOTAL, nL * F * 8, &p_alpha));
  RB_TRY(rb_ensure(ctx, RB_BUF_T, nL * 8, &p_T));
  RB_TRY(rb_ensure(ctx, RB_BUF_TB, (size_t)R * F * esz, &p_tb));
  if (out_intW) RB_TRY(rb_ensure(ctx, RB_BUF_INTW, (size_t)R * F * 8, &p_iw));
  cudaStream_t s = ctx->stream;
  const double trace_t0 = now_ms();
  bool streamed;
  const bool have_geometry = take_ticket(ctx, g, R, b, &streamed);   // radius and b were staged by the prefetch
  if (!have_geometry) {
    RB_CUDA(ctx, cudaMemcpyAsync(p_rad, g->radius, nL * 8, cudaMemcpyHostToDevice, s));
    RB_CUDA(ctx, cudaMemcpyAsync(p_b, b, (size_t)R * 16, cudaMemcpyHostToDevice, s));
  }
  if (!resident_alpha) RB_CUDA(ctx, cudaMemcpyAsync(p_alpha, rt->alpha, nL * F * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_T, rt->T, nL * 8, cudaMemcpyHostToDevice, s));
  RB_TRACE_AT("inputs enqueued");
  L.radius = (const double*)p_rad; L.b = (const double*)p_b;
  bind_ds(ctx, L, p_ds, p_n);
  RB_TRY(bind_compact(ctx, L));
  rb_rt_desc rd = *rt;
  rd.alpha = (const double*)p_alpha; rd.T = (const double*)p_T;
  RB_TRY(run_rt_pipeline(ctx, L, &rd, p_tb, (double*)p_iw, out_Tb, out_intW, have_geometry, streamed));
  RB_TRACE_AT("pipeline enqueued");
  if (profile_ray >= 0) {
    // re-run the selected ray alone with the profile-writing variant (Brightness.tau/.W/.Tb_lyr)
    RB_TRY(rb_ensure(ctx, RB_BUF_PROFILE, 3 * F * S * 8 + F * 8, &p_prof));
    RB_CUDA(ctx, cudaMemsetAsync(p_prof, 0, 3 * F * S * 8 + F * 8, s));
    // (its geometry is traced again on its own: the batch above may have run over the compacted ray list)
    RtLaunch L1 = L;
    L1.R = 1; L1.Rpad = 32;
    L1.compact = false;
    L1.dsf = nullptr;
    L1.nanflag = L.nseg + 32;
    L1.b = L.b + 2 * profile_ray;
    RB_TRY(rb_launch_geometry(ctx, L1));
    double* pp = (double*)p_prof;
    rd.out_f32 = 0;
    RtPrep prof_prep;
    RB_TRY(rb_rt_prepare(ctx, L.L, &rd, 1, true, false, &prof_prep));
    RB_TRY(rb_launch_integrate(ctx, L1, &rd, prof_prep, nullptr, pp + 3 * F * S, nullptr, 0, pp, pp + F * S, pp + 2 * F * S));
    RB_CUDA(ctx, cudaMemcpyAsync(out_tau, pp, F * S * 8, cudaMemcpyDeviceToHost, s));
    RB_CUDA(ctx, cudaMemcpyAsync(out_W, pp + F * S, F * S * 8, cudaMemcpyDeviceToHost, s));
    RB_CUDA(ctx, cudaMemcpyAsync(out_Tblyr, pp + 2 * F * S, F * S * 8, cudaMemcpyDeviceToHost, s));
  }
  RB_CUDA(ctx, cudaStreamSynchronize(s));
  RB_TRACE_AT("synchronised");
  return RB_OK;
}

int rb_rt_batch(rb_context* ctx, const rb_geometry_desc* g, const rb_rt_desc* rt, int64_t R, const double* b,
                void* out_Tb, double* out_intW, int64_t profile_ray, double* out_tau, double* out_W,
                double* out_Tblyr) {
  return rt_batch_host(ctx, g, rt, R, b, out_Tb, out_intW, profile_ray, out_tau, out_W, out_Tblyr, false);
}

int rb_rt_batch_resident(rb_context* ctx, const rb_geometry_desc* g, const rb_rt_desc* rt, int64_t R, const double* b,
                         void* out_Tb, double* out_intW, int64_t profile_ray, double* out_tau, double* out_W,
                         double* out_Tblyr) {
  return rt_batch_host(ctx, g, rt, R, b, out_Tb, out_intW, profile_ray, out_tau, out_W, out_Tblyr, true);
}

int rb_rt_integrate_profile(rb_context* ctx, const rb_rt_desc* rt, int32_t n_layers, int64_t R, int32_t n_seg,
                            const double* ds, const int32_t* nseg, void* out_Tb, double* out_intW, int64_t profile_ray,
                            double* out_tau, double* out_W, double* out_Tblyr) {
  if (!ctx) return RB_ERR_INVALID;
  RB_TRY(check_rt(ctx, rt, out_Tb, true));
  if (!ds || !nseg || R <= 0) return rb_fail(ctx, RB_ERR_INVALID, "rt_integrate: null pointer / no rays");
  if (n_layers < 2 || n_seg != n_layers - 1) return rb_fail(ctx, RB_ERR_INVALID, "rt_integrate: n_seg must be n_layers - 1");
  if (profile_ray >= R || (profile_ray >= 0 && (!out_tau || !out_W || !out_Tblyr)))
    return rb_fail(ctx, RB_ERR_INVALID, "rt_integrate: profile_ray out of range / profile outputs missing");
  RB_CUDA(ctx, cudaSetDevice(ctx->device));
  drop_ticket(ctx);
  RtLaunch L{};
  L.L = n_layers; L.R = R; L.Rpad = (R + 31) & ~(int64_t)31;
  const size_t nL = n_layers, S = n_seg, F = rt->n_freqs;
  const size_t esz = rt->out_f32 ? 4 : 8;
  void *p_in, *p_ds, *p_n, *p_alpha, *p_alpha0 = nullptr, *p_T, *p_tb, *p_iw = nullptr, *p_prof = nullptr;
  RB_TRY(rb_ensure(ctx, RB_BUF_MISC, (size_t)R * S * 8, &p_in));
  RB_TRY(rb_ensure(ctx, RB_BUF_DS, S * L.Rpad * 8 + kRtSlackBytes, &p_ds));
  RB_TRY(rb_ensure(ctx, RB_BUF_NSEG, (size_t)L.Rpad * 8, &p_n));
  RB_TRY(rb_ensure(ctx, RB_BUF_TOTAL, nL * F * 8, &p_alpha));
  if (rt->alpha0) RB_TRY(rb_ensure(ctx, RB_BUF_ALPHA0, nL * F * 8, &p_alpha0));
  RB_TRY(rb_ensure(ctx, RB_BUF_T, nL * 8, &p_T));
  RB_TRY(rb_ensure(ctx, RB_BUF_TB, (size_t)R * F * esz, &p_tb));
  if (out_intW) RB_TRY(rb_ensure(ctx, RB_BUF_INTW, (size_t)R * F * 8, &p_iw));
  cudaStream_t s = ctx->stream;
  RB_CUDA(ctx, cudaMemcpyAsync(p_in, ds, (size_t)R * S * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_n, nseg, (size_t)R * 4, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_alpha, rt->alpha, nL * F * 8, cudaMemcpyHostToDevice, s));
  if (rt->alpha0) RB_CUDA(ctx, cudaMemcpyAsync(p_alpha0, rt->alpha0, nL * F * 8, cudaMemcpyHostToDevice, s));
  RB_CUDA(ctx, cudaMemcpyAsync(p_T, rt->T, nL * 8, cudaMemcpyHostToDevice, s));
  L.ds = (double*)p_ds; L.nseg = (int32_t*)p_n; L.nanflag = (int32_t*)p_n + L.Rpad;
  RB_TRY(rb_launch_ds_to_slab(ctx, (const double*)p_in, R, L.Rpad, (int)S, L.nseg, L.nanflag, (double*)p_ds));
  rb_rt_desc rd = *rt;
  rd.alpha = (const double*)p_alpha; rd.alpha0 = (const double*)p_alpha0; rd.T = (const double*)p_T;
  RtPrep prep;
  RB_TRY(rb_rt_prepare(ctx, L.L, &rd, R, false, false, &prep));
  RB_TRY(rb_launch_integrate(ctx, L, &rd, prep, nullptr, p_tb, (double*)p_iw, -1, nullptr, nullptr, nullptr));
  RB_CUDA(ctx, cudaMemcpyAsync(out_Tb, p_tb, (size_t)R * F * esz, cudaMemcpyDeviceToHost, s));
  if (out_intW) RB_CUDA(ctx, cudaMemcpyAsync(out_intW, p_iw, (size_t)R * F * 8, cudaMemcpyDeviceToHost, s));
  if (profile_ray >= 0) {
    // the selected ray once more, alone, with the profile-writing variant (Brightness.tau / .W / .Tb_lyr)
    RB_TRY(rb_ensure(ctx, RB_BUF_PROFILE, 3 * F * S * 8 + F * 8, &p_prof));
    RB_CUDA(ctx, cudaMemsetAsync(p_prof, 0, 3 * F * S * 8 + F * 8, s));
    RtLaunch L1 = L;
    L1.R = 1; L1.Rpad = 32;
    L1.nanflag = L.nseg + 32;
    RB_CUDA(ctx, cudaMemcpyAsync(p_n, nseg + profile_ray, 4, cudaMemcpyHostToDevice, s));
    RB_TRY(rb_launch_ds_to_slab(ctx, (const double*)p_in + (size_t)profile_ray * S, 1, 32, (int)S, L1.nseg, L1.nanflag,
                                (double*)p_ds));
    double* pp = (double*)p_prof;
    rd.out_f32 = 0;
    RtPrep prof_prep;
    RB_TRY(rb_rt_prepare(ctx, L.L, &rd, 1, true, false, &prof_prep));
    RB_TRY(rb_launch_integrate(ctx, L1, &rd, prof_prep, nullptr, pp + 3 * F * S, nullptr, 0, pp, pp + F * S, pp + 2 * F * S));
    RB_CUDA(ctx, cudaMemcpyAsync(out_tau, pp, F * S * 8, cudaMemcpyDeviceToHost, s));
    RB_CUDA(ctx, cudaMemcpyAsync(out_W, pp + F * S, F * S * 8, cudaMemcpyDeviceToHost, s));
    RB_CUDA(ctx, cudaMemcpyAsync(out_Tblyr, pp + 2 * F * S, F * S * 8, cudaMemcpyDeviceToHost, s));
  }
  RB_CUDA(ctx, cudaStreamSynchronize(s));
  return RB_OK;
}

int rb_rt_integrate(rb_context* ctx, const rb_rt_desc* rt, int32_t n_layers, int64_t R, int32_t n_seg, const double* ds,
                    const int32_t* nseg, void* out_Tb, double* out_intW) {
  return rb_rt_integrate_profile(ctx, rt, n_layers, R, n_seg, ds, nseg, out_Tb, out_intW, -1, nullptr, nullptr, nullptr);
}

}  // extern "C"
