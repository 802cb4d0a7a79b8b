// psp_runtime.cu -- context, error reporting, device vectors, options.
#include "psp_common.cuh"

namespace psp {

static thread_local char g_err[512] = "";
static Ctx g_ctx;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

static thread_local Ctx *t_ctx = nullptr;         // set in the worker threads of the multi-GPU layer
Ctx &ctx() { return t_ctx ? *t_ctx : g_ctx; }
void set_thread_ctx(Ctx *c) { t_ctx = c; }

static int do_init(int device) {
  Ctx &c = g_ctx;
  if (c.ready && c.device == device) return PSP_OK;
  if (c.ready) { set_error("psp_init: already initialised on device %d", c.device); return PSP_ERR_ARG; }
  PSP_TRY(init_ctx(c, device));
  // PSP_OPTS="name=value,name=value": tuning knobs from the environment (see psp_set_option)
  if (const char *opts = getenv("PSP_OPTS")) {
    char buf[256];
    strncpy(buf, opts, sizeof buf - 1);
    buf[sizeof buf - 1] = 0;
    for (char *tok = strtok(buf, ","); tok; tok = strtok(nullptr, ",")) {
      char *eq = strchr(tok, '=');
      if (!eq) continue;
      *eq = 0;
      psp_set_option(tok, atoll(eq + 1));
    }
  }
  if (const char *g = getenv("PSP_GPUS")) psp_set_option("gpus", atoll(g));
  return PSP_OK;
}

int init_ctx(Ctx &c, int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s); pysparse_b200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return PSP_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) { set_error("psp_init: device %d out of range (%d devices)", device, count); return PSP_ERR_ARG; }
  PSP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PSP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library ships sm_100a code only",
              device, prop.name, prop.major, prop.minor);
    return PSP_ERR_NO_DEVICE;
  }
  c.device = device;
  c.sms = prop.multiProcessorCount;
  PSP_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  PSP_CUDA(cudaMalloc((void **)&c.partials, sizeof(double) * (kRedSlots + 1) * kMaxGrid));
  PSP_CUDA(cudaMalloc((void **)&c.tickets, sizeof(unsigned) * 16));
  PSP_CUDA(cudaMemset(c.tickets, 0, sizeof(unsigned) * 16));
  PSP_CUDA(cudaMalloc((void **)&c.state, sizeof(State) + 256));
  PSP_CUDA(cudaMemset(c.state, 0, sizeof(State) + 256));
  PSP_CUDA(cudaHostAlloc((void **)&c.h_state, 2 * sizeof(State), cudaHostAllocDefault));
  PSP_CUDA(cudaEventCreate(&c.ev0));
  PSP_CUDA(cudaEventCreate(&c.ev1));
  PSP_CUDA(cudaEventCreateWithFlags(&c.evpoll[0], cudaEventDisableTiming));
  PSP_CUDA(cudaEventCreateWithFlags(&c.evpoll[1], cudaEventDisableTiming));
  c.ready = true;
  return PSP_OK;
}

int ensure_init() {
  if (t_ctx) return PSP_OK;            // a rank thread: its context exists and its device is current
  if (g_ctx.ready) {
    cudaSetDevice(g_ctx.device);
    return PSP_OK;
  }
  int dev = 0;
  if (const char *s = getenv("PSP_DEVICE")) dev = atoi(s);
  else if (const char *s2 = getenv("LOCAL_RANK")) dev = atoi(s2);
  return do_init(dev);
}

int ensure_stage(size_t elems) {
  Ctx &c = ctx();
  if (c.stage_elems >= elems) return PSP_OK;
  if (c.stage) cudaFreeHost(c.stage);
  c.stage = nullptr; c.stage_elems = 0;
  size_t want = elems < (1u << 16) ? (1u << 16) : elems;
  cudaError_t e = cudaHostAlloc((void **)&c.stage, sizeof(double) * want, cudaHostAllocDefault);
  if (e != cudaSuccess) { set_error("cudaHostAlloc(%zu doubles) failed: %s", want, cudaGetErrorString(e)); return PSP_ERR_NOMEM; }
  c.stage_elems = want;
  return PSP_OK;
}

int pool_get(int slot, size_t elems, double **out) {
  Ctx &c = ctx();
  if (elems == 0) elems = 1;
  if (c.pool_cap[slot] < elems || c.pool_cap[slot] > 4 * elems + (1u << 20)) {   // grow, or give back a much larger block
    if (c.pool[slot]) cudaFree(c.pool[slot]);
    c.pool[slot] = nullptr; c.pool_cap[slot] = 0;
    cudaError_t e = cudaMalloc((void **)&c.pool[slot], sizeof(double) * elems);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc(%zu doubles) failed: %s", elems, cudaGetErrorString(e));
      return PSP_ERR_NOMEM;
    }
    c.pool_cap[slot] = elems;
  }
  *out = c.pool[slot];
  return PSP_OK;
}

__global__ void k_fill(double *p, size_t n, double v) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace psp

using namespace psp;

extern "C" int psp_init(int device) { return do_init(device); }

extern "C" int psp_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

extern "C" int psp_current_device(void) { return g_ctx.ready ? g_ctx.device : -1; }
extern "C" const char *psp_last_error(void) { return g_err; }
extern "C" const char *psp_version(void) { return "pysparse_b200 0.1.0 (sm_100a)"; }

extern "C" long long psp_launch_count(int reset) {
  long long v = g_ctx.launches;
  if (reset) g_ctx.launches = 0;
  return v;
}

extern "C" int psp_synchronize(void) {
  PSP_TRY(ensure_init());
  PSP_CUDA(cudaStreamSynchronize(ctx().stream));
  return PSP_OK;
}

extern "C" double psp_last_device_ms(void) { return g_ctx.last_ms; }

extern "C" int psp_host_alloc(size_t bytes, void **out) {
  if (!out) { set_error("psp_host_alloc: NULL out"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return PSP_ERR_NOMEM; }
  return PSP_OK;
}

extern "C" int psp_host_free(void *p) {
  if (p) PSP_CUDA(cudaFreeHost(p));
  return PSP_OK;
}

extern "C" int psp_vec_alloc(size_t n, double **dptr) {
  if (!dptr) { set_error("psp_vec_alloc: NULL out"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  cudaError_t e = cudaMalloc((void **)dptr, sizeof(double) * (n ? n : 1));
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu doubles) failed: %s", n, cudaGetErrorString(e)); return PSP_ERR_NOMEM; }
  return PSP_OK;
}

extern "C" int psp_vec_free(double *dptr) {
  if (dptr) PSP_CUDA(cudaFree(dptr));
  return PSP_OK;
}

// host -> device, element stride `inc` on the host side (NumPy strided views)
extern "C" int psp_vec_upload(double *dptr, const double *host, size_t n, long inc) {
  if (n == 0) return PSP_OK;
  if (!dptr || !host) { set_error("psp_vec_upload: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  Ctx &c = ctx();
  if (inc == 1) {
    PSP_CUDA(cudaMemcpyAsync(dptr, host, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
    PSP_CUDA(cudaStreamSynchronize(c.stream));
    return PSP_OK;
  }
  PSP_TRY(ensure_stage(n));
  for (size_t i = 0; i < n; i++) c.stage[i] = host[(long)i * inc];
  PSP_CUDA(cudaMemcpyAsync(dptr, c.stage, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  return PSP_OK;
}

extern "C" int psp_vec_download(const double *dptr, double *host, size_t n, long inc) {
  if (n == 0) return PSP_OK;
  if (!dptr || !host) { set_error("psp_vec_download: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  Ctx &c = ctx();
  if (inc == 1) {
    PSP_CUDA(cudaMemcpyAsync(host, dptr, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    PSP_CUDA(cudaStreamSynchronize(c.stream));
    return PSP_OK;
  }
  PSP_TRY(ensure_stage(n));
  PSP_CUDA(cudaMemcpyAsync(c.stage, dptr, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  for (size_t i = 0; i < n; i++) host[(long)i * inc] = c.stage[i];
  return PSP_OK;
}

extern "C" int psp_vec_fill(double *dptr, size_t n, double value) {
  if (n == 0) return PSP_OK;
  PSP_TRY(ensure_init());
  k_fill<<<grid_for(n), kThreads, 0, ctx().stream>>>(dptr, n, value);
  PSP_LAUNCH_CHECK();
  return PSP_OK;
}

extern "C" int psp_event_record(int idx) {
  if (idx < 0 || idx >= 8) { set_error("psp_event_record: index out of range"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  if (!g_ctx.user_ev[idx]) PSP_CUDA(cudaEventCreate(&g_ctx.user_ev[idx]));
  PSP_CUDA(cudaEventRecord(g_ctx.user_ev[idx], g_ctx.stream));
  return PSP_OK;
}

extern "C" int psp_event_elapsed_ms(int start, int stop, double *ms) {
  if (start < 0 || start >= 8 || stop < 0 || stop >= 8 || !ms || !g_ctx.user_ev[start] || !g_ctx.user_ev[stop]) {
    set_error("psp_event_elapsed_ms: bad arguments");
    return PSP_ERR_ARG;
  }
  PSP_CUDA(cudaEventSynchronize(g_ctx.user_ev[stop]));
  float f = 0.f;
  PSP_CUDA(cudaEventElapsedTime(&f, g_ctx.user_ev[start], g_ctx.user_ev[stop]));
  *ms = f;
  return PSP_OK;
}

extern "C" int psp_spmv_timing(int sample_every, int *count, double *total_ms) {
  PSP_TRY(ensure_init());
  Ctx &c = g_ctx;
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  double tot = 0.0;
  for (int i = 0; i < c.spmv_pairs; i++) {
    float f = 0.f;
    PSP_CUDA(cudaEventElapsedTime(&f, c.spmv_ev[2 * i], c.spmv_ev[2 * i + 1]));
    tot += f;
  }
  if (count) *count = c.spmv_pairs;
  if (total_ms) *total_ms = tot;
  c.spmv_pairs = 0;
  c.spmv_seen = 0;
  c.spmv_sample = sample_every;
  return PSP_OK;
}

extern "C" int psp_set_option(const char *name, long long value) {
  if (!name) return PSP_ERR_ARG;
  Ctx &c = g_ctx;
  if (!strcmp(name, "spmv_variant")) c.spmv_variant = value;
  else if (!strcmp(name, "spmv_ctas")) c.spmv_ctas = value;
  else if (!strcmp(name, "spmv_format")) c.spmv_format = value;
  else if (!strcmp(name, "gpus")) {
    if (value < 1 || value > kMaxRanks) { set_error("gpus must be 1..%d", kMaxRanks); return PSP_ERR_ARG; }
    c.gpus = value;
  }
  else if (!strcmp(name, "poll_chunk")) c.poll_chunk = value;
  else if (!strcmp(name, "l2_stream")) c.l2_stream = value;
  else if (!strcmp(name, "pcg_persist")) c.pcg_persist = value;
  else if (!strcmp(name, "ssor_engine")) c.ssor_engine = value;
  else if (!strcmp(name, "dist_debug")) c.dist_debug = value;
  else if (!strcmp(name, "dist_push")) c.dist_push = value;
  else if (!strcmp(name, "bicg_persist")) c.bicg_persist = value;
  else if (!strcmp(name, "gmres_persist")) c.gmres_persist = value;
  else if (!strcmp(name, "ssor_cluster")) c.ssor_cluster = (value >= 1 && value <= 16) ? value : 8;
  else { set_error("unknown option '%s'", name); return PSP_ERR_ARG; }
  return PSP_OK;
}

extern "C" long long psp_get_option(const char *name) {
  Ctx &c = g_ctx;
  if (!name) return -1;
  if (!strcmp(name, "spmv_variant")) return c.spmv_variant;
  if (!strcmp(name, "spmv_ctas")) return c.spmv_ctas;
  if (!strcmp(name, "spmv_format")) return c.spmv_format;
  if (!strcmp(name, "gpus")) return c.gpus;
  if (!strcmp(name, "poll_chunk")) return c.poll_chunk;
  if (!strcmp(name, "sms")) return c.sms;
  if (!strcmp(name, "l2_stream")) return c.l2_stream;
  if (!strcmp(name, "pcg_persist")) return c.pcg_persist;
  if (!strcmp(name, "ssor_engine")) return c.ssor_engine;
  if (!strcmp(name, "ssor_cluster")) return c.ssor_cluster;
  if (!strcmp(name, "dist_debug")) return c.dist_debug;
  if (!strcmp(name, "dist_push")) return c.dist_push;
  if (!strcmp(name, "bicg_persist")) return c.bicg_persist;
  if (!strcmp(name, "gmres_persist")) return c.gmres_persist;
  return -1;
}
