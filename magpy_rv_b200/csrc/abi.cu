// C ABI of magpy_rv_b200 (see include/magpy_rv_b200.h): problem handle, path selection, host loop of
// the blocked factorisation, dense helpers, FP64 tensor-pipe microbenchmark.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost ~ns unless a profiler is attached

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/magpy_rv_b200.h"
#include "blocked.cuh"
#include "dag.cuh"
#include "common.cuh"
#include "dense_ops.cuh"
#include "mean_model.cuh"
#include "small_logl.cuh"
#include "mid_logl.cuh"
#include "host_sampler.h"
#include "device_sampler.cuh"

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(-2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define SMALL_N_MAX 166   // largest N whose (N + 1) x N matrix + 4 N point / residual doubles fit 227 KB of shared memory

static unsigned long long g_buf_gen = 0;   // bumped by every (re)allocation: captured CUDA graphs hold raw pointers

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    ++g_buf_gen;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) return fail(-3, "cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() { return reinterpret_cast<T*>(p); }
};

// Pinned host staging (the per-batch parameters in, log L / status out): a pageable cudaMemcpyAsync stages and
// synchronises inside the driver (~10 us each, four per batch); one pinned buffer each way makes the small
// batches of the tutorial-size runs one true async copy in and one out.
struct PinBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    ++g_buf_gen;
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaHostAlloc(&p, need, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(-3, "cudaHostAlloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
};

// Optional per-category kernel timing with CUDA events on the launching stream (bench.py roofline).
enum { CAT_GEN = 0, CAT_POTRF, CAT_TRSM, CAT_UPDATE_PANEL, CAT_UPDATE_MID, CAT_UPDATE_TRAIL, CAT_MEAN, CAT_FINAL, CAT_FUSED, CAT_DAG, CAT_COUNT };
static const char* const kCatNames[CAT_COUNT] = {"gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final", "fused", "dag"};
struct ProfRec {
  int cat;
  cudaEvent_t e0, e1;
  double alg_flops, exec_flops;
};
struct Prof {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  double ns[CAT_COUNT] = {0}, alg[CAT_COUNT] = {0}, exec[CAT_COUNT] = {0};
  int64_t n[CAT_COUNT] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  // every kernel category is also an NVTX range on the host (gen, potrf, trsm, update_*, mean, final, fused, dag):
  // nsys / ncu --nvtx show the phases of a batch; the reference has only a wall-clock print (mcmc.py:778,868)
  void begin(int cat, cudaStream_t st, double alg_f, double exec_f) {
    nvtxRangePushA(kCatNames[cat]);
    if (!on) return;
    ProfRec r{cat, get(), get(), alg_f, exec_f};
    cudaEventRecord(r.e0, st);
    recs.push_back(r);
  }
  void end(cudaStream_t st) {
    nvtxRangePop();
    if (!on) return;
    cudaEventRecord(recs.back().e1, st);
  }
  void collect() {  // call after the stream has been synchronised
    for (ProfRec& r : recs) {
      float ms = 0.f;
      if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        ns[r.cat] += (double)ms * 1e6;
        alg[r.cat] += r.alg_flops;
        exec[r.cat] += r.exec_flops;
        n[r.cat]++;
      }
      pool.push_back(r.e0);
      pool.push_back(r.e1);
    }
    recs.clear();
  }
  void reset() {
    collect();
    for (int c = 0; c < CAT_COUNT; ++c) ns[c] = alg[c] = exec[c] = 0.0, n[c] = 0;
  }
  void destroy() {
    collect();
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    pool.clear();
  }
};

struct mrv_problem {
  int device = 0;
  Prof prof;
  ProblemDesc desc;
  bool has_kepler = false;
  int n_pad = 0, nt = 0;
  double *d_t = nullptr, *d_y = nullptr, *d_yerr = nullptr, *d_flags = nullptr;
  cudaStream_t stream = nullptr;  // used by the host-pointer entry points
  // options
  int path_opt = MRV_PATH_AUTO;
  int64_t wave_opt = 0;
  int panel_opt = 0;
  int num_sms = 148;
  int potrf_impl = 0;  // 0 = blocked rank-16 sweep with DMMA trailing updates, 1 = rank-1 sweep
  int tiled_impl = 0;  // tiled path: 0 = automatic, 1 = one launch per tile column and kernel (run_wave), 2 = persistent dataflow kernel
  int band_opt = 0;    // UPDATE rasterisation band (tile rows), 0 = default
  int panel_inner_opt = 0;  // inner panel width of the two-level scheme, 0 = default (8)
  int cs_store = 1;    // streaming stores for updated C tiles
  int64_t ensemble_opt = 0; // option "ensemble_walkers": size of the WHOLE ensemble when this handle evaluates one shard of it --
                            // the size-dependent kernel choices below then do not depend on the number of shards, so a walker's
                            // log L is bitwise the same on 1, 2, 4 or 8 GPUs (0 = the batch is the ensemble)
  int small_minb_opt = 0;   // 2 / 3 / 4 forces the register-cap variant of the fused kernel (0 = by shared memory)
  int small_blk_min = 32;   // below this the rank-1 shared-memory kernel is used (option "small_blk_min"); measured:
                            // the packed blocked sweep wins from N = 32 up (tools/small_probe.py), equal at 16
  int lookahead_opt = 1;    // diagonal-first look-ahead inside a panel (auxiliary stream per sub-wave)
  cudaStream_t la_stream[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t la_evT[4] = {nullptr, nullptr, nullptr, nullptr}, la_evO[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t budget_bytes = 0;  // wave workspace budget, sampled at the first blocked-path call
  int substreams_opt = 0;  // 0 = auto (two half-waves on two streams up to 32 tile columns), 1 = one stream
  cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  // Keplerian mean models (tiled path): the Newton kernel runs on its own stream next to the generation of tile
  // column 0, which does not need the residuals; the first diagonal-tile kernel waits for it
  int mean_overlap_opt = 1;
  bool mean_pending = false;
  cudaStream_t mean_stream = nullptr;
  cudaEvent_t ev_mean_fork = nullptr, ev_mean = nullptr;
  // workspace
  DevBuf hp, modpar, model, logl, status;                      // host-API staging (hp holds hp | modpar, logl holds logL | status)
  PinBuf pin_in, pin_out;
  DevBuf resid, escratch, phys, nonfinite, logdet, quad, fstatus, zbuf, factor, counter, midws, zcols, rowdone, phase, cpar;
  // host-buffer entry point, one-CTA-per-walker paths: copy in -> kernel(s) -> copy out replayed as ONE CUDA graph
  // launch (three stream operations per batch otherwise; the batch of a tutorial-size MCMC iteration is ~60 us of GPU work)
  int graph_opt = 1;
  cudaGraphExec_t graph_exec = nullptr;
  int64_t graph_W = -1, graph_ens = -1;
  int graph_to_ecc = -1, graph_state = 0;   // state: 0 = nothing seen, 1 = this key ran eagerly once (buffers exist), 2 = captured
  unsigned long long graph_buf_gen = 0, opt_gen = 0, graph_opt_gen = 0;
  int64_t graph_launches = 0, graph_replays = 0;
  int64_t wave_cur = 0;
  int64_t launches = 0;
  int last_path = 0;
  bool attrs_set = false;
};

// ------------------------------------------------------------------------------------------------
extern "C" const char* mrv_last_error(void) { return g_err.c_str(); }
extern "C" int mrv_abi_version(void) { return MRV_ABI_VERSION; }
extern "C" int mrv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int require_device(int device) {
  int n = mrv_device_count();
  if (n <= 0) return fail(-4, "no CUDA device available (magpy_rv_b200 has no CPU fallback)");
  if (device < 0 || device >= n) return fail(-4, "device %d out of range (have %d)", device, n);
  CU(cudaSetDevice(device));
  return 0;
}

static int kernel_nh(int id) {
  static const int nh[7] = {2, 2, 3, 4, 5, 2, 3};
  return (id >= 0 && id < 7) ? nh[id] : -1;
}

static int fill_desc(ProblemDesc& D, int64_t N, int kernel_id, int variant, int nh, const mrv_model_op* ops, int n_ops,
                     int nm, const mrv_prior* priors, int n_priors) {
  if (N <= 0 || N > (1 << 20)) return fail(-1, "N=%lld out of range", (long long)N);
  if (kernel_nh(kernel_id) < 0) return fail(-1, "unknown kernel id %d", kernel_id);
  if (nh != kernel_nh(kernel_id)) return fail(-1, "kernel %d needs %d hyper-parameters, got %d", kernel_id, kernel_nh(kernel_id), nh);
  if (n_ops < 0 || n_ops > MRV_MAX_OPS) return fail(-1, "n_ops=%d out of range (max %d)", n_ops, MRV_MAX_OPS);
  if (nm < 0 || nm > MRV_MAX_MODPAR) return fail(-1, "nm=%d out of range (max %d)", nm, MRV_MAX_MODPAR);
  if (n_priors < 0 || n_priors > MRV_MAX_PRIORS) return fail(-1, "n_priors=%d out of range (max %d)", n_priors, MRV_MAX_PRIORS);
  memset(&D, 0, sizeof D);
  D.N = (int)N;
  D.kernel_id = kernel_id;
  D.variant = variant;
  D.nh = nh;
  D.nm = nm;
  D.n_ops = n_ops;
  D.n_priors = n_priors;
  static const int npar[4] = {1, 4, 1, 5};
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].type < 0 || ops[i].type > 3) return fail(-1, "unknown model op type %d", ops[i].type);
    if (ops[i].param_offset < 0 || ops[i].param_offset + npar[ops[i].type] > nm)
      return fail(-1, "model op %d reads parameters [%d,%d) outside nm=%d", i, ops[i].param_offset,
                  ops[i].param_offset + npar[ops[i].type], nm);
    D.ops[i].type = ops[i].type;
    D.ops[i].param_offset = ops[i].param_offset;
    D.ops[i].flag_val = ops[i].flag_val;
  }
  for (int i = 0; i < n_priors; ++i) {
    if (priors[i].type < 0 || priors[i].type > 3) return fail(-1, "unknown prior type %d", priors[i].type);
    const int lim = priors[i].source == 0 ? nh : nm;
    if (priors[i].source < 0 || priors[i].source > 1 || priors[i].index < 0 || priors[i].index >= lim)
      return fail(-1, "prior %d refers to parameter %d of source %d (limit %d)", i, priors[i].index, priors[i].source, lim);
    D.priors[i].type = priors[i].type;
    D.priors[i].source = priors[i].source;
    D.priors[i].index = priors[i].index;
    D.priors[i].p0 = priors[i].p0;
    D.priors[i].p1 = priors[i].p1;
    D.priors[i].p2 = priors[i].p2;
  }
  return 0;
}

static int upload(double** dst, const double* src, size_t n, size_t n_alloc = 0) {
  n_alloc = std::max<size_t>(std::max(n, n_alloc), 1);   // n_alloc > n: zero padded (whole 128-point tiles can be copied)
  CU(cudaMalloc((void**)dst, n_alloc * sizeof(double)));
  if (n_alloc > n) CU(cudaMemset(*dst, 0, n_alloc * sizeof(double)));
  CU(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int mrv_problem_create(mrv_problem** out, int device, const double* t, const double* y, const double* yerr,
                                  const double* flags, int64_t N, int kernel_id, int kernel_variant, int nh,
                                  const mrv_model_op* ops, int n_ops, int nm, const mrv_prior* priors, int n_priors) {
  if (!out || !t || !y || !yerr) return fail(-1, "null argument");
  *out = nullptr;
  if (int rc = require_device(device)) return rc;
  mrv_problem* p = new (std::nothrow) mrv_problem();
  if (!p) return fail(-3, "out of host memory");
  p->device = device;
  if (int rc = fill_desc(p->desc, N, kernel_id, kernel_variant, nh, ops, n_ops, nm, priors, n_priors)) {
    delete p;
    return rc;
  }
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].type == MRV_MODEL_KEPLER) p->has_kepler = true;
    if (ops[i].type == MRV_MODEL_OFFSET && !flags) {
      delete p;
      return fail(-1, "an OFFSET model op needs a flags array");
    }
  }
  p->nt = (int)((N + TILE - 1) / TILE);
  p->n_pad = p->nt * TILE;
  p->desc.t_ref = t[N / 2];   // phase origin of the periodic covariance terms: phases stay small for any time offset (BJD ...)
  int rc = 0;
  if ((rc = upload(&p->d_t, t, N, p->n_pad)) || (rc = upload(&p->d_y, y, N)) || (rc = upload(&p->d_yerr, yerr, N)) ||
      (flags && (rc = upload(&p->d_flags, flags, N)))) {
    mrv_problem_destroy(p);
    return rc;
  }
  cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    mrv_problem_destroy(p);
    return fail(-2, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return 0;
}

extern "C" void mrv_problem_destroy(mrv_problem* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaFree(p->d_t);
  cudaFree(p->d_y);
  cudaFree(p->d_yerr);
  cudaFree(p->d_flags);
  DevBuf* bufs[] = {&p->hp, &p->modpar, &p->model, &p->logl, &p->status, &p->resid, &p->escratch, &p->phys,
                    &p->nonfinite, &p->logdet, &p->quad, &p->fstatus, &p->zbuf, &p->factor, &p->counter, &p->midws,
                    &p->zcols, &p->rowdone, &p->phase, &p->cpar};
  for (DevBuf* b : bufs) b->release();
  p->pin_in.release();
  p->pin_out.release();
  p->prof.destroy();
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  if (p->stream) cudaStreamDestroy(p->stream);
  for (int i = 0; i < 3; ++i) {
    if (p->aux_stream[i]) cudaStreamDestroy(p->aux_stream[i]);
    if (p->ev_join[i]) cudaEventDestroy(p->ev_join[i]);
  }
  for (int i = 0; i < 4; ++i) {
    if (p->la_stream[i]) cudaStreamDestroy(p->la_stream[i]);
    if (p->la_evT[i]) cudaEventDestroy(p->la_evT[i]);
    if (p->la_evO[i]) cudaEventDestroy(p->la_evO[i]);
  }
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->mean_stream) cudaStreamDestroy(p->mean_stream);
  if (p->ev_mean_fork) cudaEventDestroy(p->ev_mean_fork);
  if (p->ev_mean) cudaEventDestroy(p->ev_mean);
  delete p;
}

// ------------------------------------------------------------------------------------------------
// kernel attribute setup (opt-in shared memory sizes), once per process per device is enough but it
// is cheap, so once per handle.
// ------------------------------------------------------------------------------------------------
// opt in to the largest dynamic shared memory a kernel can get: 227 KiB minus whatever it declares statically
template <class K>
static int allow_max_dynamic_smem(K kernel) {
  cudaFuncAttributes fa;
  CU(cudaFuncGetAttributes(&fa, kernel));
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
  return 0;
}

static int set_attrs(mrv_problem* p) {
  if (p->attrs_set) return 0;
  if (int rc = allow_max_dynamic_smem(fused_small_logl_kernel)) return rc;
  if (int rc = allow_max_dynamic_smem(fused_small_blk_kernel<2>)) return rc;
  if (int rc = allow_max_dynamic_smem(fused_small_blk_kernel<3>)) return rc;
  if (int rc = allow_max_dynamic_smem(fused_small_blk_kernel<4>)) return rc;
  if (int rc = allow_max_dynamic_smem(mid_logl_kernel<7>)) return rc;
  if (int rc = allow_max_dynamic_smem(mid_logl_kernel<15>)) return rc;
  CU(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM_BYTES));
  CU(cudaFuncSetAttribute(potrf_tile_blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_BLK_SMEM_BYTES));
  CU(cudaFuncSetAttribute(gemm_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_TMA_SMEM_BYTES));
  CU(cudaFuncSetAttribute(tile_dag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DAG_SMEM_BYTES));
  CU(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));
  p->attrs_set = true;
  return 0;
}

// Panel width (tile columns per trailing update).  Measured optimum (tools/panel_probe.py): 8 up to
// N = 4096, 12 at 8192, 24 at 16384 -- wider panels amortise the C-tile traffic and the per-tile
// fill/drain of the trailing GEMM, until the in-panel updates (shorter K) start to dominate.
static int effective_panel(const mrv_problem* p) {
  if (p->panel_opt > 0) return p->panel_opt;
  return std::min(24, std::max(8, (p->nt * 3) / 16));
}

// one point per thread up to 512 points: the Keplerian Newton loop is a chain of up to 200 block-wide
// reductions per walker, so fewer points per thread shortens every link
static int g_mean_threads_opt = 0;   // option "mean_threads" (probe only): 0 = automatic
static unsigned mean_threads(int N) { return g_mean_threads_opt ? (unsigned)g_mean_threads_opt : (N > 256 ? 512u : 256u); }

static int choose_path(const mrv_problem* p, int64_t W = 0) {
  if (p->path_opt == MRV_PATH_FUSED_SMEM) return MRV_PATH_FUSED_SMEM;
  if (p->path_opt == MRV_PATH_BLOCKED) return MRV_PATH_BLOCKED;
  if (p->path_opt == MRV_PATH_MID) return MRV_PATH_MID;
  // measured crossovers (profiles/README.md): fused shared-memory kernels up to 144 points, the
  // one-CTA-per-walker streamed-panel kernel up to 512, the tiled wave factorisation above
  if (p->desc.N <= MRV_SMALL_BLK_N_MAX) {
    // above 112 points only one fused CTA fits an SM; the streamed-panel kernel keeps two walkers per SM resident and,
    // with its one-sweep diagonal blocks, is 1.1-1.35x ahead as soon as the ensemble needs more than one round of fused
    // CTAs (N = 128 x 1024 walkers: 3.09 M against 2.50 M evals/s; up to 148 walkers the fused kernel's latency wins)
    if (p->desc.N > 112 && W > p->num_sms) return MRV_PATH_MID;
    return MRV_PATH_FUSED_SMEM;
  }
  if (p->desc.N > MID_N_MAX) return MRV_PATH_BLOCKED;
  // 320 < N <= 512 (round 2, tools/crossover_probe.py, profiles/r2s_crossover_probe.txt): the tiled path -- the dataflow
  // kernel from 3 tile columns on -- pads N to a multiple of 128, the one-CTA-per-walker kernel runs in rounds of num_sms
  // walkers.  Tiled wins when N fills its tiles: always at N = 384 / 512 (1.07-1.37x), at N / n_pad >= 0.87 when the rounds
  // are badly filled (256 or 512 walkers: 86 %), >= 0.90 for well filled rounds (1024 walkers), >= 0.94 for exactly one
  // full round (148 walkers, the streamed-panel kernel's best case).
  if (p->desc.N > 320) {
    const int64_t Wn = W > 0 ? W : p->num_sms;
    const int64_t rounds = (Wn + p->num_sms - 1) / p->num_sms;
    const double fill = (double)Wn / (double)(rounds * p->num_sms);
    const double thr = fill >= 0.995 ? 0.94 : (fill >= 0.93 ? 0.90 : 0.87);
    if ((double)p->desc.N >= thr * (double)p->n_pad) return MRV_PATH_BLOCKED;
  }
  return MRV_PATH_MID;
}

// Blocked path for one wave of `B` walkers starting at walker w0.
// `slot0`: first walker slot of this (sub-)wave inside the factor / z workspaces of the current wave.
static int run_wave(mrv_problem* p, const double* d_hp, int64_t w0, int64_t B, cudaStream_t st, const SolveOut& so,
                    int64_t slot0 = 0, int sub = 0) {
  const ProblemDesc& D = p->desc;
  const int nt = p->nt, N = D.N;
  const size_t walker_stride = tile_index(nt - 1, nt - 1) * TILE_ELEMS + TILE_ELEMS;
  double* A = p->factor.as<double>() + (size_t)slot0 * walker_stride;
  double* resid = p->resid.as<double>() + (size_t)w0 * p->n_pad;
  double* zbuf = p->zbuf.as<double>() + (size_t)slot0 * TILE;
  double* logdet = p->logdet.as<double>() + w0;
  double* quad = p->quad.as<double>() + w0;
  int* fstatus = p->fstatus.as<int>() + w0;
  const double* hp_wave = d_hp + (size_t)w0 * D.nh;
  double* zout = so.z ? so.z + (size_t)w0 * so.ldz : nullptr;
  const int PW = effective_panel(p);

  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.N = N;
  g.nt = nt;
  g.kernel_id = D.kernel_id;
  g.variant = D.variant;
  g.nh = D.nh;
  g.Abase = A;
  g.walker_stride = walker_stride;
  g.t = p->d_t;
  g.yerr = p->d_yerr;
  g.hp_wave = hp_wave;
  g.ph_wave = p->phase.as<double>() + (size_t)w0 * 2 * p->n_pad;
  g.resid = resid;
  g.ldr = p->n_pad;
  g.zbuf = zbuf;
  // band height x panel width x 128 KiB = A rows that should stay L2-resident while the B rows stream
  // past once per band: ~48 MiB measured best for DRAM traffic (ncu, profiles/README.md) at <= 0.5 % speed cost
  g.band = p->band_opt > 0 ? p->band_opt : std::min(64, std::max(14, 384 / PW));
  g.cs_store = p->cs_store;

  auto tiles_in_cols = [&](int c0, int c1) {
    long long n = 0;
    for (int j = c0; j < c1; ++j) n += nt - j;
    return (unsigned)n;
  };

  // Diagonal-first look-ahead inside a panel: the diagonal tile of column k is updated first and factored on the
  // main stream while the off-diagonal tiles of the column are updated on an auxiliary stream, so the
  // latency-bound diagonal-tile kernel (one CTA per walker, ~60 us) leaves the critical path.
  // measured on top of the two half-waves (tools/substream_probe.py): +5 % at N = 1024 x 128 walkers, +0..1 % for
  // other small waves, -2 % at 640 x 512 -- so only (sub-)waves of at most 64 walkers use it (option 2 forces it)
  const bool la = nt <= 64 && (p->lookahead_opt == 2 || (p->lookahead_opt == 1 && B <= 64));
  cudaStream_t aux = nullptr;
  if (la) {
    if (!p->la_stream[sub]) CU(cudaStreamCreateWithFlags(&p->la_stream[sub], cudaStreamNonBlocking));
    if (!p->la_evT[sub]) CU(cudaEventCreateWithFlags(&p->la_evT[sub], cudaEventDisableTiming));
    if (!p->la_evO[sub]) CU(cudaEventCreateWithFlags(&p->la_evO[sub], cudaEventDisableTiming));
    aux = p->la_stream[sub];
  }
  auto launch_gemm = [&](dim3 grid, cudaStream_t st) {
    gemm_tile_kernel<<<grid, GEMM_TMA_THREADS, GEMM_TMA_SMEM_BYTES, st>>>(g);
  };
  const double T3 = (double)TILE * TILE * TILE;
  // algorithmic / executed FLOPs of an UPDATE launch over target columns [c0, c1) with np operand columns
  auto update_flops = [&](int c0, int c1, int np, double* exec_f) {
    double alg_f = 0.0, ex = 0.0;
    for (int j = c0; j < c1; ++j) {
      const double offdiag = (double)(nt - j - 1);
      alg_f += (offdiag * 2.0 * T3 + (double)TILE * (TILE + 1) * TILE) * np;
      ex += (offdiag + 1.0) * 2.0 * T3 * np;
    }
    *exec_f = ex * (double)B;
    return alg_f * (double)B;
  };
  p->prof.begin(CAT_GEN, st, 0.0, 0.0);
  gen_column_kernel<<<dim3(nt, (unsigned)B), 256, 0, st>>>(N, nt, 0, D.kernel_id, D.variant, D.nh, p->d_t, p->d_yerr,
                                                            hp_wave, g.ph_wave, A, walker_stride);
  p->prof.end(st);
  p->launches++;
  if (p->mean_pending) CU(cudaStreamWaitEvent(st, p->ev_mean, 0));   // residuals are first read by the diagonal-tile kernel
  // part: 0 = every target tile of the columns, 1 = only the diagonal tile of column c0, 2 = column c0 without it
  auto launch_update = [&](int cat, int c0, int c1, int pb, int pe, cudaStream_t s, int part = 0) {
    g.mode = GEMM_UPDATE;
    g.col_begin = c0;
    g.col_end = c1;
    g.p_begin = pb;
    g.p_end = pe;
    g.gen = (pb == 0);   // the operand range starts at column 0 <=> this is the first touch of the targets
    g.x_off = part == 2 ? 1 : 0;
    double ex = 0.0;
    double al = update_flops(c0, c1, pe - pb, &ex);
    unsigned tiles = tiles_in_cols(c0, c1);
    if (part != 0) {
      const double np_ = (double)(pe - pb), diag_al = (double)TILE * (TILE + 1) * TILE * np_ * (double)B,
                   diag_ex = 2.0 * T3 * np_ * (double)B;
      al = part == 1 ? diag_al : al - diag_al;
      ex = part == 1 ? diag_ex : ex - diag_ex;
      tiles = part == 1 ? 1u : tiles - 1u;
    }
    p->prof.begin(cat, s, al, ex);
    launch_gemm(dim3(tiles, (unsigned)B), s);
    p->prof.end(s);
    p->launches++;
    g.x_off = 0;
  };
  // Two-level panels: outer panels of PW tile columns get one big trailing update (K = 128 PW);
  // inside an outer panel, inner panels of PWI columns are factored column by column (short-K
  // in-panel updates) and then applied to the rest of the outer panel in one mid-level update
  // (K = 128 PWI), so only ~PWI/2 of every PW columns' worth of flops runs at short K.
  // (measured, tools/panel_probe.py: the mid-level updates run at the same rate as the in-panel ones, so the
  // second level buys nothing -- 24/8: 33.19, 24/12: 33.24, 24/24: 33.40 TFLOP/s at N = 16384; default = one level)
  const int PWI = std::min(PW, p->panel_inner_opt > 0 ? p->panel_inner_opt : PW);
  for (int P0 = 0; P0 < nt; P0 += PW) {
    const int P1 = std::min(P0 + PW, nt);
    for (int p0 = P0; p0 < P1; p0 += PWI) {
      const int p1 = std::min(p0 + PWI, P1);
      for (int k = p0; k < p1; ++k) {
        const bool split = la && k > p0 && k + 1 < nt;
        if (split) {
          launch_update(CAT_UPDATE_PANEL, k, k + 1, p0, k, st, 1);
          CU(cudaStreamWaitEvent(aux, p->la_evT[sub], 0));          // recorded after TRSM(k - 1)
          launch_update(CAT_UPDATE_PANEL, k, k + 1, p0, k, aux, 2);
          CU(cudaEventRecord(p->la_evO[sub], aux));
        } else if (k > p0) {
          launch_update(CAT_UPDATE_PANEL, k, k + 1, p0, k, st);
        }
        p->prof.begin(CAT_POTRF, st, (T3 / 3.0 + 2.0 * TILE * TILE) * (double)B, 1.5 * T3 * (double)B);
        if (p->potrf_impl == 0)
          potrf_tile_blk_kernel<<<(unsigned)B, POTRF_BLK_THREADS, POTRF_BLK_SMEM_BYTES, st>>>(
              k, N, A, walker_stride, resid, p->n_pad, zbuf, logdet, quad, fstatus, zout, so.ldz);
        else
          potrf_tile_kernel<<<(unsigned)B, 256, POTRF_SMEM_BYTES, st>>>(k, N, A, walker_stride, resid, p->n_pad, zbuf,
                                                                        logdet, quad, fstatus, zout, so.ldz);
        p->prof.end(st);
        p->launches++;
        if (k + 1 < nt) {
          if (split) CU(cudaStreamWaitEvent(st, p->la_evO[sub], 0));
          g.mode = GEMM_TRSM;
          g.k = k;
          p->prof.begin(CAT_TRSM, st, (T3 + 2.0 * TILE * TILE) * (nt - k - 1) * (double)B,
                        2.0 * T3 * (nt - k - 1) * (double)B);
          launch_gemm(dim3(nt - k - 1, (unsigned)B), st);
          p->prof.end(st);
          p->launches++;
          if (la) CU(cudaEventRecord(p->la_evT[sub], st));
        }
      }
      if (p1 < P1) launch_update(CAT_UPDATE_MID, p1, P1, p0, p1, st);
    }
    if (P1 < nt) launch_update(CAT_UPDATE_TRAIL, P1, nt, P0, P1, st);
  }
  CU(cudaGetLastError());
  return 0;
}

// Tiled path as ONE persistent dataflow kernel per wave (dag.cuh).  Left-looking by tiles: every tile streams its
// operands once per tile -- 16 flop per operand byte at EVERY N, 2.3 TB/s at the DMMA peak, from L2 while a wave's factors
// fit (N <= 4096) and from HBM above.  Round 1 assumed the right-looking panels of run_wave were needed beyond N = 4096
// (operand re-use); measured at the end of round 2 the dataflow kernel is ahead at every size it can index (up to 128
// tile columns): N = 8192 x 128 walkers 93.6 % of the DMMA peak against 89.5 %, N = 16384 93.8 % (waves of 32) / 95.1 %
// (waves of 64) against 92.1 %, bitwise identical results (tools/big_dag_probe.py).
static bool use_dag(const mrv_problem* p, int64_t wave = 0) {
  (void)wave;
  if (p->tiled_impl == 1) return false;
  if (p->tiled_impl == 2) return p->nt <= DAG_MAX_TILES;
  return p->nt >= DAG_AUTO_MIN_TILES && p->nt <= DAG_MAX_TILES;
}

static int run_wave_dag(mrv_problem* p, const double* d_hp, int64_t w0, int64_t B, cudaStream_t st, const SolveOut& so, int wave_index) {
  const ProblemDesc& D = p->desc;
  const int nt = p->nt;
  DagArgs g;
  memset(&g, 0, sizeof g);
  g.N = D.N;
  g.nt = nt;
  g.B = (int)B;
  {
    const double hp0[MRV_MAX_HP] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
    g.kind = make_cov_params(D.kernel_id, D.variant, hp0).kind;
  }
  const long long total = (long long)B * nt * (nt + 1) / 2;
  if (total > 0x7fffffffLL) return fail(-1, "wave of %lld walkers x %d tile columns has too many tile tasks", (long long)B, nt);
  g.total_tasks = (int)total;
  g.Abase = p->factor.as<double>();
  g.walker_stride = tile_index(nt - 1, nt - 1) * TILE_ELEMS + TILE_ELEMS;
  g.t = p->d_t;
  g.yerr = p->d_yerr;
  g.ph_wave = p->phase.as<double>() + (size_t)w0 * 2 * p->n_pad;
  g.cpar_wave = p->cpar.as<double>() + (size_t)w0 * 8;
  g.resid = p->resid.as<double>() + (size_t)w0 * p->n_pad;
  g.ldr = p->n_pad;
  g.zcols = p->zcols.as<double>();
  g.rowdone = p->rowdone.as<int>() + (size_t)w0 * nt;          // cleared for all walkers by phase_table_kernel
  g.abort_flag = p->counter.as<int>();                         // [0] abort flag of the batch, [1 + wave] task counters
  g.counter = p->counter.as<int>() + 1 + wave_index;
  g.acc_logdet = p->logdet.as<double>() + w0;
  g.acc_quad = p->quad.as<double>() + w0;
  g.status = p->fstatus.as<int>() + w0;
  g.zout = so.z ? so.z + (size_t)w0 * so.ldz : nullptr;
  g.ldz = so.ldz;
  const unsigned grid = (unsigned)std::min<long long>(total, p->num_sms);
  const double fl = ((double)D.N * D.N * D.N / 3.0 + 2.0 * D.N * D.N) * (double)B;
  const double T3 = (double)TILE * TILE * TILE;
  // executed: every tile (i,k) runs k full tile products, off-diagonal tiles one more for the solve (5/8 of it issued)
  double ex = 0.0;
  for (int k = 0; k < nt; ++k) ex += (double)(nt - k) * 2.0 * T3 * k + (double)(nt - k - 1) * 2.0 * T3 * 0.625 + 1.5 * T3;
  p->prof.begin(CAT_DAG, st, fl, ex * (double)B);
  tile_dag_kernel<<<grid, DAG_THREADS, DAG_SMEM_BYTES, st>>>(g);
  p->prof.end(st);
  p->launches++;
  CU(cudaGetLastError());
  return 0;
}

// Core: everything on device, asynchronous on `st`.
static int run_logl(mrv_problem* p, const double* d_hp, const double* d_modpar, const double* d_model_given, int64_t W,
                    int to_ecc, double* d_logl, int* d_status, cudaStream_t st, SolveOut so = SolveOut{nullptr, 0, nullptr, 0}) {
  if (W <= 0) return 0;
  if (W > 65535 * 64) return fail(-1, "W=%lld too large", (long long)W);
  CU(cudaSetDevice(p->device));
  if (int rc = set_attrs(p)) return rc;
  const ProblemDesc& D = p->desc;
  const int64_t Wsel = p->ensemble_opt > 0 ? std::max<int64_t>(p->ensemble_opt, W) : W;
  const int path = choose_path(p, Wsel);
  p->last_path = path;
  if (path == MRV_PATH_FUSED_SMEM) {
    if (D.N > SMALL_N_MAX) return fail(-1, "fused shared-memory path needs N <= %d (N=%d)", SMALL_N_MAX, D.N);
    const bool blk = p->potrf_impl == 0 && D.N >= p->small_blk_min && D.N <= MRV_SMALL_BLK_N_MAX;
    const size_t smem = blk ? small_blk_smem_bytes(D.N) : small_smem_bytes(D.N);
    const double fl = ((double)D.N * D.N * D.N / 3.0 + 2.0 * D.N * D.N) * (double)W;
    p->prof.begin(CAT_FUSED, st, fl, fl);
    if (blk) {
      // as many resident CTAs as the shared memory allows (1 KiB reserved per CTA), register cap to match
      const int fit = (int)((227 * 1024) / (smem + 1024));
      // ... but no tighter than the ensemble needs: a single round of CTAs runs fastest with the loosest cap
      const int rounds = (int)((Wsel + p->num_sms - 1) / p->num_sms);
      int minb = p->small_minb_opt > 0 ? p->small_minb_opt : std::max(2, std::min(4, std::min(fit, rounds)));
      // the 4 / 3-CTA variants size their sweep for at most 64 / 80 columns (small_logl.cuh)
      if (minb >= 4 && sblk_cols(D.N) > 64) minb = 3;
      if (minb == 3 && sblk_cols(D.N) > 80) minb = 2;
      auto kern = minb >= 4 ? fused_small_blk_kernel<4> : (minb == 3 ? fused_small_blk_kernel<3> : fused_small_blk_kernel<2>);
      kern<<<(unsigned)W, MRV_SMALL_THREADS, smem, st>>>(D, p->d_t, p->d_y, p->d_yerr, p->d_flags, d_hp, d_modpar,
                                                         d_model_given, to_ecc, d_logl, d_status, so);
    }
    else
      fused_small_logl_kernel<<<(unsigned)W, MRV_SMALL_THREADS, smem, st>>>(D, p->d_t, p->d_y, p->d_yerr, p->d_flags,
                                                                            d_hp, d_modpar, d_model_given, to_ecc,
                                                                            d_logl, d_status, so);
    p->prof.end(st);
    p->launches++;
    CU(cudaGetLastError());
    return 0;
  }

  if (path == MRV_PATH_MID) {
    if (D.N > MID_N_MAX) return fail(-1, "streamed-panel path needs N <= %d (N=%d)", MID_N_MAX, D.N);
    const int nw = mid_warps(D.N);
    const size_t smem = mid_smem_bytes(D.N);
    const int per_sm = (nw == 7 && 2 * (smem + 1024) <= (size_t)227 * 1024) ? 2 : 1;
    const unsigned grid = (unsigned)std::min<int64_t>(W, (int64_t)p->num_sms * per_sm);
    const size_t ws_stride = mid_ws_doubles(D.N);
    if (int rc = p->midws.ensure(std::max<size_t>(1, (size_t)grid * ws_stride) * sizeof(double))) return rc;
    // Keplerian mean models: residuals of ALL walkers first (one CTA each, many per SM), so the up to 200
    // block-wide Newton reductions per walker do not serialise in front of every factorisation
    const double* resid_ready = nullptr;
    const double* phys_ready = nullptr;
    const int* nonfinite_ready = nullptr;
    if (p->has_kepler && d_model_given == nullptr) {
      int rc = 0;
      if ((rc = p->resid.ensure((size_t)W * p->n_pad * sizeof(double)))) return rc;
      if ((rc = p->escratch.ensure((size_t)W * p->n_pad * sizeof(double)))) return rc;
      if ((rc = p->phys.ensure(std::max<size_t>(1, (size_t)W * D.nm) * sizeof(double)))) return rc;
      if ((rc = p->nonfinite.ensure((size_t)W * sizeof(int)))) return rc;
      p->prof.begin(CAT_MEAN, st, 0.0, 0.0);
      mean_residual_kernel<<<(unsigned)W, mean_threads(D.N), 0, st>>>(D, p->d_t, p->d_y, p->d_flags, d_modpar, nullptr, to_ecc,
                                                        p->resid.as<double>(), p->n_pad, p->escratch.as<double>(),
                                                        p->phys.as<double>(), p->nonfinite.as<int>());
      p->prof.end(st);
      p->launches++;
      resid_ready = p->resid.as<double>();
      phys_ready = p->phys.as<double>();
      nonfinite_ready = p->nonfinite.as<int>();
    }
    const double fl = ((double)D.N * D.N * D.N / 3.0 + 2.0 * D.N * D.N) * (double)W;
    p->prof.begin(CAT_FUSED, st, fl, fl);
    if (nw == 7)
      mid_logl_kernel<7><<<grid, 256, smem, st>>>(D, p->d_t, p->d_y, p->d_yerr, p->d_flags, d_hp, d_modpar, d_model_given,
                                                  to_ecc, (int)W, d_logl, d_status, so, p->midws.as<double>(), ws_stride,
                                                  resid_ready, p->n_pad, phys_ready, nonfinite_ready);
    else
      mid_logl_kernel<15><<<grid, 512, smem, st>>>(D, p->d_t, p->d_y, p->d_yerr, p->d_flags, d_hp, d_modpar, d_model_given,
                                                   to_ecc, (int)W, d_logl, d_status, so, p->midws.as<double>(), ws_stride,
                                                   resid_ready, p->n_pad, phys_ready, nonfinite_ready);
    p->prof.end(st);
    p->launches++;
    CU(cudaGetLastError());
    return 0;
  }

  // ---- blocked path ----
  const int nt = p->nt;
  const size_t walker_bytes = (tile_index(nt - 1, nt - 1) + 1) * TILE_ELEMS * sizeof(double);
  int64_t B = p->wave_opt;
  if (B <= 0) {
    // the memory budget is sampled once per handle (cudaMemGetInfo costs milliseconds right after another
    // handle released its workspace, and the answer only matters for the first allocation)
    if (p->budget_bytes == 0) {
      size_t free_b = 0, total_b = 0;
      CU(cudaMemGetInfo(&free_b, &total_b));
      free_b += p->factor.bytes;  // what we already hold can be reused
      // at most half of what is free, capped at 72 GiB: 64 walkers per wave at N = 16384 (1.03 GiB of packed tiles each;
      // one wave of 128 = 132 GiB measured 95.6 % of the DMMA peak against 95.2 % -- not worth two thirds of the HBM)
      p->budget_bytes = std::max<size_t>(1, std::min<size_t>((size_t)(0.5 * (double)free_b), (size_t)72 << 30));
    }
    B = (int64_t)std::max<size_t>(1, p->budget_bytes / walker_bytes);
  }
  B = std::min<int64_t>(B, W);
  B = std::min<int64_t>(B, 65535);
  {  // equal-sized waves (e.g. 128 walkers with room for 39 -> 4 x 32, not 39+39+39+11)
    const int64_t nwaves = (W + B - 1) / B;
    B = (W + nwaves - 1) / nwaves;
  }
  p->wave_cur = B;
  int rc = 0;
  if ((rc = p->resid.ensure((size_t)W * p->n_pad * sizeof(double)))) return rc;
  if (p->has_kepler && (rc = p->escratch.ensure((size_t)W * p->n_pad * sizeof(double)))) return rc;
  if ((rc = p->phys.ensure(std::max<size_t>(1, (size_t)W * D.nm) * sizeof(double)))) return rc;
  if ((rc = p->nonfinite.ensure((size_t)W * sizeof(int)))) return rc;
  if ((rc = p->logdet.ensure((size_t)W * sizeof(double)))) return rc;
  if ((rc = p->quad.ensure((size_t)W * sizeof(double)))) return rc;
  if ((rc = p->fstatus.ensure((size_t)W * sizeof(int)))) return rc;
  if ((rc = p->zbuf.ensure((size_t)B * TILE * sizeof(double)))) return rc;
  const int64_t n_waves = (W + B - 1) / B;
  if ((rc = p->counter.ensure((size_t)(n_waves + 2) * sizeof(int)))) return rc;
  if ((rc = p->factor.ensure((size_t)B * walker_bytes))) return rc;
  const bool dag = use_dag(p, B);
  if (dag) {
    if ((rc = p->zcols.ensure((size_t)B * nt * TILE * sizeof(double)))) return rc;
    if ((rc = p->rowdone.ensure((size_t)W * nt * sizeof(int)))) return rc;
  }

  if ((rc = p->phase.ensure((size_t)W * 2 * p->n_pad * sizeof(double)))) return rc;
  if ((rc = p->cpar.ensure((size_t)W * 8 * sizeof(double)))) return rc;
  p->prof.begin(CAT_GEN, st, 0.0, 0.0);
  phase_table_kernel<<<dim3((unsigned)std::min(8, (p->n_pad + 255) / 256), (unsigned)std::min<int64_t>(W, 65535)), 256, 0, st>>>(
      D.N, p->n_pad, (int)W, D.t_ref, D.kernel_id, D.variant, D.nh, p->d_t, d_hp, p->phase.as<double>(), p->cpar.as<double>(),
      p->logdet.as<double>(), p->quad.as<double>(), p->fstatus.as<int>(), dag ? p->rowdone.as<int>() : nullptr, nt,
      p->counter.as<int>(), (int)(n_waves + 2));
  p->prof.end(st);
  p->launches++;
  cudaStream_t mst = st;
  p->mean_pending = p->has_kepler && d_model_given == nullptr && p->mean_overlap_opt != 0 && !dag;
  if (p->mean_pending) {
    if (!p->mean_stream) CU(cudaStreamCreateWithFlags(&p->mean_stream, cudaStreamNonBlocking));
    if (!p->ev_mean_fork) CU(cudaEventCreateWithFlags(&p->ev_mean_fork, cudaEventDisableTiming));
    if (!p->ev_mean) CU(cudaEventCreateWithFlags(&p->ev_mean, cudaEventDisableTiming));
    mst = p->mean_stream;
    CU(cudaEventRecord(p->ev_mean_fork, st));      // after the parameter upload and the previous call's consumers
    CU(cudaStreamWaitEvent(mst, p->ev_mean_fork, 0));
  }
  p->prof.begin(CAT_MEAN, mst, 0.0, 0.0);
  mean_residual_kernel<<<(unsigned)W, mean_threads(D.N), 0, mst>>>(D, p->d_t, so.given_is_resid ? nullptr : p->d_y, p->d_flags, d_modpar,
                                                    d_model_given, to_ecc,
                                                    p->resid.as<double>(), p->n_pad, p->escratch.as<double>(),
                                                    p->phys.as<double>(), p->nonfinite.as<int>());
  p->prof.end(mst);
  p->launches++;
  CU(cudaGetLastError());
  if (p->mean_pending) CU(cudaEventRecord(p->ev_mean, mst));
  // Mid sizes: a wave is cut into two halves on two streams, so the latency-bound
  // diagonal-tile kernels and the small launches near the end of every tile column of one half overlap
  // with the GEMMs of the other (per-walker results do not depend on the split: nothing is shared).
  // measured (tools/substream_probe.py, medians of interleaved runs): two half-waves give +5..12 % up to
  // 16 tile columns and +3..9 % for waves of <= 64 walkers up to N = 8192; with 128 walkers at N = 4096 the
  // difference is inside the +-2 % run-to-run noise, and three or four streams never add anything (and
  // occasionally lose 10..40 %), so the automatic choice is two streams exactly where they were a robust win
  int nsub = p->substreams_opt > 0 ? p->substreams_opt : (((nt <= 16 && B >= 16) || (nt <= 64 && B >= 16 && B <= 64)) ? 2 : 1);
  nsub = std::max(1, std::min(nsub, 4));
  if (nsub > 1) {
    if (!p->ev_fork) CU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < nsub - 1; ++i) {
      if (!p->aux_stream[i]) CU(cudaStreamCreateWithFlags(&p->aux_stream[i], cudaStreamNonBlocking));
      if (!p->ev_join[i]) CU(cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming));
    }
  }
  for (int64_t w0 = 0; w0 < W; w0 += B) {
    const int64_t Bw = std::min<int64_t>(B, W - w0);
    const int parts = (int)std::min<int64_t>(nsub, Bw);
    if (dag) {
      if ((rc = run_wave_dag(p, d_hp, w0, Bw, st, so, (int)(w0 / B)))) return rc;
    } else if (parts > 1) {
      CU(cudaEventRecord(p->ev_fork, st));
      int64_t off = 0;
      for (int i = 0; i < parts; ++i) {
        const int64_t n = Bw / parts + (i < Bw % parts ? 1 : 0);
        cudaStream_t si = i == 0 ? st : p->aux_stream[i - 1];
        if (i > 0) CU(cudaStreamWaitEvent(si, p->ev_fork, 0));
        if ((rc = run_wave(p, d_hp, w0 + off, n, si, so, off, i))) return rc;
        if (i > 0) {
          CU(cudaEventRecord(p->ev_join[i - 1], si));
          CU(cudaStreamWaitEvent(st, p->ev_join[i - 1], 0));
        }
        off += n;
      }
    } else if ((rc = run_wave(p, d_hp, w0, Bw, st, so))) {
      return rc;
    }
  }
  p->prof.begin(CAT_FINAL, st, 0.0, 0.0);
  finalize_logl_kernel<<<(unsigned)((W + 127) / 128), 128, 0, st>>>(D, (int)W, d_hp, p->phys.as<double>(),
                                                                    p->logdet.as<double>(), p->quad.as<double>(),
                                                                    p->nonfinite.as<int>(), p->fstatus.as<int>(), d_logl,
                                                                    d_status, so.logdet, dag ? p->counter.as<int>() : nullptr);
  p->prof.end(st);
  p->launches++;
  CU(cudaGetLastError());
  return 0;
}

extern "C" int mrv_logl_batch_device(mrv_problem* p, const double* d_hp, const double* d_modpar, int64_t W, int to_ecc,
                                     double* d_logL_out, int32_t* d_status_out, void* stream) {
  if (!p || !d_hp || !d_modpar || !d_logL_out || !d_status_out) return fail(-1, "null argument");
  return run_logl(p, d_hp, d_modpar, nullptr, W, to_ecc, d_logL_out, d_status_out, (cudaStream_t)stream);
}

// Host-buffer evaluation in two halves, so that one host thread can keep several devices busy (mrv_shard_logl_batch):
// `enqueue` stages the parameters and enqueues copy in -> kernels -> copy out on the handle's stream (or replays the
// captured graph), `finish` waits for the stream and hands the results over.
static int host_logl_finish(mrv_problem* p, int64_t W, double* logL_out, int32_t* status_out) {
  if (W <= 0) return 0;
  CU(cudaSetDevice(p->device));
  CU(cudaStreamSynchronize(p->stream));
  memcpy(logL_out, p->pin_out.p, (size_t)W * sizeof(double));
  memcpy(status_out, static_cast<char*>(p->pin_out.p) + (size_t)W * sizeof(double), (size_t)W * sizeof(int));
  return 0;
}

static int host_logl_enqueue(mrv_problem* p, const double* hp, const double* modpar, const double* model_y, int64_t W, int to_ecc) {
  if (!p || !hp || !modpar) return fail(-1, "null argument");
  if (W <= 0) return 0;
  struct Range {
    Range() { nvtxRangePushA("mrv_logl_batch"); }
    ~Range() { nvtxRangePop(); }
  } nvtx_range;
  CU(cudaSetDevice(p->device));
  const ProblemDesc& D = p->desc;
  int rc = 0;
  const size_t n_hp = (size_t)W * D.nh, n_mp = (size_t)W * D.nm;
  const size_t in_bytes = (n_hp + std::max<size_t>(1, n_mp)) * sizeof(double);
  const size_t out_bytes = (size_t)W * (sizeof(double) + sizeof(int));
  if ((rc = p->hp.ensure(in_bytes)) || (rc = p->logl.ensure(out_bytes))) return rc;
  if ((rc = p->pin_in.ensure(in_bytes)) || (rc = p->pin_out.ensure(out_bytes))) return rc;
  cudaStream_t st = p->stream;
  double* h_in = static_cast<double*>(p->pin_in.p);
  memcpy(h_in, hp, n_hp * sizeof(double));
  if (n_mp > 0) memcpy(h_in + n_hp, modpar, n_mp * sizeof(double));
  double* d_hp = p->hp.as<double>();
  double* d_mp = d_hp + n_hp;
  double* d_logl = p->logl.as<double>();
  int* d_status = reinterpret_cast<int*>(d_logl + W);

  // CUDA-graph replay for the single-stream paths: first call of a (W, to_ecc) key runs eagerly (allocations, function
  // attributes), the second is captured, later ones are one cudaGraphLaunch.  Any reallocation or option change drops it.
  const int64_t Wsel = p->ensemble_opt > 0 ? std::max<int64_t>(p->ensemble_opt, W) : W;
  const int path_now = choose_path(p, Wsel);
  // (the tiled path qualifies when it runs as the dataflow kernel: phase table -> mean -> one kernel per wave -> finalize on
  // one stream, workspaces sized by the first, eager call; at C3 the five launches + two copies cost 0.1 ms of a 1.08 ms call)
  const bool graphable = p->graph_opt && !model_y && !p->prof.on &&
                         (path_now == MRV_PATH_FUSED_SMEM || path_now == MRV_PATH_MID ||
                          (path_now == MRV_PATH_BLOCKED && use_dag(p) && p->nt <= 32));
  const bool same_key = graphable && p->graph_W == W && p->graph_to_ecc == to_ecc && p->graph_buf_gen == g_buf_gen &&
                        p->graph_opt_gen == p->opt_gen && p->graph_ens == p->ensemble_opt;
  if (same_key && p->graph_state == 2) {
    CU(cudaGraphLaunch(p->graph_exec, st));
    p->launches += p->graph_launches;
    p->graph_replays++;
    p->last_path = path_now;
    return 0;
  }
  const bool capture = same_key && p->graph_state == 1;
  if (capture) CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int64_t launches0 = p->launches;
  cudaError_t ce = cudaMemcpyAsync(d_hp, h_in, (n_hp + n_mp) * sizeof(double), cudaMemcpyHostToDevice, st);
  const double* d_model = nullptr;
  if (ce == cudaSuccess && model_y) {
    if ((rc = p->model.ensure((size_t)W * D.N * sizeof(double)))) return rc;
    ce = cudaMemcpyAsync(p->model.p, model_y, (size_t)W * D.N * sizeof(double), cudaMemcpyHostToDevice, st);
    d_model = p->model.as<double>();
  }
  if (ce == cudaSuccess) rc = run_logl(p, d_hp, d_mp, d_model, W, to_ecc, d_logl, d_status, st);
  if (ce == cudaSuccess && rc == 0) ce = cudaMemcpyAsync(p->pin_out.p, d_logl, out_bytes, cudaMemcpyDeviceToHost, st);
  if (capture) {
    cudaGraph_t graph = nullptr;
    cudaError_t ee = cudaStreamEndCapture(st, &graph);     // always end the capture, also on errors
    if (p->graph_exec) {
      cudaGraphExecDestroy(p->graph_exec);
      p->graph_exec = nullptr;
    }
    p->graph_state = 0;
    if (ce == cudaSuccess && rc == 0 && ee == cudaSuccess && graph != nullptr &&
        cudaGraphInstantiate(&p->graph_exec, graph, 0) == cudaSuccess) {
      p->graph_state = 2;
      p->graph_launches = p->launches - launches0;
      p->launches = launches0;                             // nothing ran yet
    }
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc) return rc;
    if (p->graph_state != 2) {
      p->graph_opt = 0;                                    // capture is not possible here: stay eager for this handle
      return host_logl_enqueue(p, hp, modpar, model_y, W, to_ecc);
    }
    CU(cudaGraphLaunch(p->graph_exec, st));
    p->launches += p->graph_launches;
  } else {
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(-2, "mrv_logl_batch copy / launch failed: %s", cudaGetErrorString(ce));
    if (graphable) {                                       // remember the key: the next identical call is captured
      p->graph_W = W;
      p->graph_ens = p->ensemble_opt;
      p->graph_to_ecc = to_ecc;
      p->graph_buf_gen = g_buf_gen;
      p->graph_opt_gen = p->opt_gen;
      p->graph_state = 1;
    } else {
      p->graph_state = 0;
    }
  }
  return 0;
}

static int host_logl(mrv_problem* p, const double* hp, const double* modpar, const double* model_y, int64_t W, int to_ecc,
                     double* logL_out, int32_t* status_out) {
  if (!logL_out || !status_out) return fail(-1, "null argument");
  if (int rc = host_logl_enqueue(p, hp, modpar, model_y, W, to_ecc)) return rc;
  return host_logl_finish(p, W, logL_out, status_out);
}

extern "C" int mrv_logl_batch(mrv_problem* p, const double* hp, const double* modpar, int64_t W, int to_ecc,
                              double* logL_out, int32_t* status_out) {
  return host_logl(p, hp, modpar, nullptr, W, to_ecc, logL_out, status_out);
}

extern "C" int mrv_logl_given_model(mrv_problem* p, const double* hp, const double* modpar, const double* model_y,
                                    int64_t W, double* logL_out, int32_t* status_out) {
  if (!model_y) return fail(-1, "null model_y");
  return host_logl(p, hp, modpar, model_y, W, 0, logL_out, status_out);
}

// ------------------------------------------------------------------------------------------------
// Walker sharding inside the library (SURVEY.md 8(b) mrv_shard_init): one host thread, one problem handle per
// device.  The ensemble is cut into contiguous blocks (sizes differ by at most one, as sharding.partition), every
// device gets its block enqueued asynchronously, then the results are collected -- the "all-gather" of this layout is
// the host buffer the caller already owns, so no NCCL and no torch are needed (a maintainer who patches the reference
// as in INTEGRATION.md B shards by passing more than one device).  Per-walker results are bitwise those of one device:
// every handle selects its kernels by the size of the WHOLE ensemble.
// ------------------------------------------------------------------------------------------------
struct mrv_shard {
  std::vector<mrv_problem*> parts;
};

extern "C" int mrv_shard_create(mrv_shard** out, int n_dev, const int* devices, const double* t, const double* y,
                                const double* yerr, const double* flags, int64_t N, int kernel_id, int kernel_variant, int nh,
                                const mrv_model_op* ops, int n_ops, int nm, const mrv_prior* priors, int n_priors) {
  if (!out || !devices || n_dev <= 0) return fail(-1, "mrv_shard_create: need at least one device");
  *out = nullptr;
  mrv_shard* sh = new (std::nothrow) mrv_shard();
  if (!sh) return fail(-3, "out of host memory");
  for (int i = 0; i < n_dev; ++i) {
    mrv_problem* p = nullptr;
    const int rc = mrv_problem_create(&p, devices[i], t, y, yerr, flags, N, kernel_id, kernel_variant, nh, ops, n_ops, nm, priors, n_priors);
    if (rc) {
      for (mrv_problem* q : sh->parts) mrv_problem_destroy(q);
      delete sh;
      return rc;
    }
    sh->parts.push_back(p);
  }
  *out = sh;
  return 0;
}

extern "C" void mrv_shard_destroy(mrv_shard* sh) {
  if (!sh) return;
  for (mrv_problem* p : sh->parts) mrv_problem_destroy(p);
  delete sh;
}

extern "C" int mrv_shard_count(mrv_shard* sh) { return sh ? (int)sh->parts.size() : 0; }

extern "C" mrv_problem* mrv_shard_part(mrv_shard* sh, int i) {
  return (sh && i >= 0 && i < (int)sh->parts.size()) ? sh->parts[(size_t)i] : nullptr;
}

extern "C" int mrv_shard_logl_batch(mrv_shard* sh, const double* hp, const double* modpar, int64_t W, int to_ecc,
                                    double* logL_out, int32_t* status_out) {
  if (!sh || !hp || !modpar || !logL_out || !status_out) return fail(-1, "null argument");
  if (W <= 0) return 0;
  const int64_t G = (int64_t)sh->parts.size();
  const int nh = sh->parts[0]->desc.nh, nm = sh->parts[0]->desc.nm;
  const int64_t base = W / G, rem = W % G;
  int rc = 0;
  int64_t start = 0;
  for (int64_t r = 0; r < G; ++r) {          // enqueue everywhere first ...
    const int64_t n = base + (r < rem ? 1 : 0);
    mrv_problem* p = sh->parts[(size_t)r];
    p->ensemble_opt = W;
    if (n > 0 && (rc = host_logl_enqueue(p, hp + start * nh, modpar + start * nm, nullptr, n, to_ecc))) return rc;
    start += n;
  }
  start = 0;
  for (int64_t r = 0; r < G; ++r) {          // ... then collect in walker order
    const int64_t n = base + (r < rem ? 1 : 0);
    if (n > 0 && (rc = host_logl_finish(sh->parts[(size_t)r], n, logL_out + start, status_out + start))) return rc;
    start += n;
  }
  return 0;
}

// z_b = L_b^-1 rhs_b and ln det K_b for W (hyper-parameter, right-hand side) pairs: the forward half
// of cho_solve, exposed for GPLikelihood.predict (gp_likelihood.py:434-440).
extern "C" int mrv_solve_batch(mrv_problem* p, const double* hp, const double* rhs, int64_t W, double* z_out,
                               double* logdet_out, int32_t* status_out) {
  if (!p || !hp || !rhs || !z_out || !status_out) return fail(-1, "null argument");
  if (W <= 0) return 0;
  CU(cudaSetDevice(p->device));
  const ProblemDesc& D = p->desc;
  int rc = 0;
  DevBuf zbuf, ldbuf, mpbuf;
  auto done = [&](int code) {
    zbuf.release();
    ldbuf.release();
    mpbuf.release();
    return code;
  };
  if ((rc = p->hp.ensure((size_t)W * D.nh * sizeof(double)))) return rc;
  if ((rc = p->model.ensure((size_t)W * D.N * sizeof(double)))) return rc;
  if ((rc = p->logl.ensure((size_t)W * sizeof(double)))) return rc;
  if ((rc = p->status.ensure((size_t)W * sizeof(int)))) return rc;
  if ((rc = zbuf.ensure((size_t)W * D.N * sizeof(double)))) return done(rc);
  if ((rc = ldbuf.ensure((size_t)W * sizeof(double)))) return done(rc);
  if ((rc = mpbuf.ensure(std::max<size_t>(1, (size_t)W * D.nm) * sizeof(double)))) return done(rc);
  cudaStream_t st = p->stream;
  cudaError_t e = cudaMemcpyAsync(p->hp.p, hp, (size_t)W * D.nh * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->model.p, rhs, (size_t)W * D.N * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(mpbuf.p, 0, mpbuf.bytes, st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_solve_batch copy-in: %s", cudaGetErrorString(e)));
  // priors are part of LogL, not of the solve: evaluate with an empty prior list
  const int n_priors_saved = p->desc.n_priors;
  p->desc.n_priors = 0;
  SolveOut so{zbuf.as<double>(), (long long)D.N, ldbuf.as<double>(), 1};
  rc = run_logl(p, p->hp.as<double>(), mpbuf.as<double>(), p->model.as<double>(), W, 0, p->logl.as<double>(),
                p->status.as<int>(), st, so);
  p->desc.n_priors = n_priors_saved;
  if (rc) return done(rc);
  e = cudaMemcpyAsync(z_out, zbuf.p, (size_t)W * D.N * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && logdet_out)
    e = cudaMemcpyAsync(logdet_out, ldbuf.p, (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(status_out, p->status.p, (size_t)W * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_solve_batch copy-out: %s", cudaGetErrorString(e)));
  return done(0);
}

// ------------------------------------------------------------------------------------------------
// R right-hand sides against ONE factorisation (GPLikelihood.predict: K is factored once, then L^-1 [y | Ks^T]).
// After a tiled factorisation of one walker the workspace holds L(i,k) in tiles (i,k), i > k, and inv(L_kk) in the
// diagonal tiles (operand order).  Forward substitution by tile columns: z_k = inv(L_kk) r_k, r_i -= L(i,k) z_k.
// One CTA owns MRHS right-hand sides for the whole sweep (no inter-CTA dependency); thread = one tile row x half of
// the CTA's right-hand sides; every tile is read once per CTA as coalesced 32-byte pieces (from L2 after the first CTA).
// ------------------------------------------------------------------------------------------------
#define MRHS 8
__global__ void __launch_bounds__(256)
multi_rhs_forward_kernel(int nt, const double* __restrict__ A, double* __restrict__ Z, long long ldz, int R) {
  __shared__ double zk[TILE][MRHS];   // [column][right-hand side]
  const int tid = threadIdx.x, row = tid & (TILE - 1), half = tid >> 7;
  const int r0 = blockIdx.x * MRHS;
  const int v0 = 4 * half;            // this thread's four right-hand sides: r0 + v0 .. r0 + v0 + 3
  auto tile_times_zk = [&](const double* __restrict__ tile, double (&acc)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 0.0;
#pragma unroll 4
    for (int cg = 0; cg < TILE / 4; ++cg) {
      const double2* src = reinterpret_cast<const double2*>(tile + (((size_t)cg * 16 + (row >> 3)) << 5) + ((row & 7) << 2));
      const double2 a01 = src[0], a23 = src[1];
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double2 z01 = *reinterpret_cast<const double2*>(&zk[4 * cg + u][v0]);
        const double2 z23 = *reinterpret_cast<const double2*>(&zk[4 * cg + u][v0 + 2]);
        acc[0] = fma(a[u], z01.x, acc[0]);
        acc[1] = fma(a[u], z01.y, acc[1]);
        acc[2] = fma(a[u], z23.x, acc[2]);
        acc[3] = fma(a[u], z23.y, acc[3]);
      }
    }
  };
  for (int k = 0; k < nt; ++k) {
    __syncthreads();
    // r_k of the CTA's right-hand sides -> zk (as the operand of the diagonal product)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = v0 + j;
      zk[row][v] = (r0 + v < R) ? Z[(size_t)(r0 + v) * ldz + (size_t)k * TILE + row] : 0.0;
    }
    __syncthreads();
    double acc[4];
    tile_times_zk(A + tile_index(k, k) * TILE_ELEMS, acc);   // inv(L_kk) is stored with zeros above the diagonal
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = v0 + j;
      zk[row][v] = acc[j];
      if (r0 + v < R) Z[(size_t)(r0 + v) * ldz + (size_t)k * TILE + row] = acc[j];
    }
    __syncthreads();
    for (int i = k + 1; i < nt; ++i) {
      tile_times_zk(A + tile_index(i, k) * TILE_ELEMS, acc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = v0 + j;
        if (r0 + v < R) Z[(size_t)(r0 + v) * ldz + (size_t)i * TILE + row] -= acc[j];
      }
    }
  }
}

// z_r = L^-1 rhs_r for R right-hand sides and ONE hyper-parameter set: K is factorised once (tiled path, one
// walker), then every right-hand side is swept through the stored factor.  The building block of
// GPLikelihood.predict (gp_likelihood.py:429-441; the reference factorises K twice there, mrv_solve_batch
// with a broadcast hp did it R times).
extern "C" int mrv_solve_multi(mrv_problem* p, const double* hp, const double* rhs, int64_t R, double* z_out,
                               double* logdet_out, int32_t* status_out) {
  if (!p || !hp || !rhs || !z_out || !status_out) return fail(-1, "null argument");
  if (R <= 0) return 0;
  CU(cudaSetDevice(p->device));
  const ProblemDesc& D = p->desc;
  const int N = D.N;
  int rc = 0;
  DevBuf zwork, ldbuf, mpbuf, r0buf;
  auto done = [&](int code) {
    zwork.release();
    ldbuf.release();
    mpbuf.release();
    r0buf.release();
    return code;
  };
  const size_t ldz = (size_t)p->n_pad;
  if ((rc = p->hp.ensure((size_t)D.nh * sizeof(double)))) return rc;
  if ((rc = p->logl.ensure(sizeof(double)))) return rc;
  if ((rc = p->status.ensure(sizeof(int)))) return rc;
  if ((rc = zwork.ensure((size_t)R * ldz * sizeof(double)))) return done(rc);
  if ((rc = ldbuf.ensure(sizeof(double)))) return done(rc);
  if ((rc = mpbuf.ensure(std::max<size_t>(1, (size_t)D.nm) * sizeof(double)))) return done(rc);
  if ((rc = r0buf.ensure((size_t)N * sizeof(double)))) return done(rc);
  cudaStream_t st = p->stream;
  cudaError_t e = cudaMemcpyAsync(p->hp.p, hp, (size_t)D.nh * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(zwork.p, 0, zwork.bytes, st);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(zwork.p, ldz * sizeof(double), rhs, (size_t)N * sizeof(double), (size_t)N * sizeof(double), (size_t)R,
                          cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r0buf.p, rhs, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(mpbuf.p, 0, mpbuf.bytes, st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_solve_multi copy-in: %s", cudaGetErrorString(e)));
  // factorise once: tiled path, one walker, no priors; the factor stays in the workspace (slot 0)
  const int n_priors_saved = p->desc.n_priors, path_saved = p->path_opt;
  const int64_t ens_saved = p->ensemble_opt;
  p->desc.n_priors = 0;
  p->path_opt = MRV_PATH_BLOCKED;
  p->ensemble_opt = 0;
  SolveOut so{nullptr, 0, ldbuf.as<double>(), 1};
  rc = run_logl(p, p->hp.as<double>(), mpbuf.as<double>(), r0buf.as<double>(), 1, 0, p->logl.as<double>(), p->status.as<int>(), st, so);
  p->desc.n_priors = n_priors_saved;
  p->path_opt = path_saved;
  p->ensemble_opt = ens_saved;
  if (rc) return done(rc);
  multi_rhs_forward_kernel<<<(unsigned)((R + MRHS - 1) / MRHS), 256, 0, st>>>(p->nt, p->factor.as<double>(), zwork.as<double>(),
                                                                              (long long)ldz, (int)R);
  p->launches++;
  e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(z_out, (size_t)N * sizeof(double), zwork.p, ldz * sizeof(double), (size_t)N * sizeof(double), (size_t)R,
                          cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && logdet_out) e = cudaMemcpyAsync(logdet_out, ldbuf.p, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(status_out, p->status.p, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_solve_multi: %s", cudaGetErrorString(e)));
  return done(0);
}

// mean_p = V_p . z ;  cov_pq = Kss_pq - V_p . V_q  (full) or  var_p = Kss_pp - |V_p|^2.
// One warp per output entry, fixed-order reduction.
__global__ void predict_reduce_kernel(const double* __restrict__ V, const double* __restrict__ z,
                                      const double* __restrict__ Kss, long long Np, long long N, int full,
                                      double* __restrict__ mean_out, double* __restrict__ cov_out) {
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long n_cov = full ? Np * Np : Np;
  if (warp_global < Np) {  // means
    const double* vp = V + warp_global * N;
    double s = 0.0;
    for (long long i = lane; i < N; i += 32) s = fma(vp[i], z[i], s);
    s = warp_sum(s);
    if (lane == 0) mean_out[warp_global] = s;
  } else if (warp_global < Np + n_cov) {
    const long long e = warp_global - Np;
    const long long pi = full ? e / Np : e, qi = full ? e % Np : e;
    const double *vp = V + pi * N, *vq = V + qi * N;
    double s = 0.0;
    for (long long i = lane; i < N; i += 32) s = fma(vp[i], vq[i], s);
    s = warp_sum(s);
    if (lane == 0) cov_out[e] = Kss[pi * Np + qi] - s;
  }
}

extern "C" int mrv_predict_reduce(int device, const double* V, const double* z, const double* Kss, int64_t Np, int64_t N,
                                  int full_cov, double* mean_out, double* cov_out) {
  if (!V || !z || !Kss || !mean_out || !cov_out) return fail(-1, "null argument");
  if (int rc = require_device(device)) return rc;
  if (Np <= 0 || N <= 0) return 0;
  double *d_V = nullptr, *d_z = nullptr, *d_K = nullptr, *d_m = nullptr, *d_c = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_V);
    cudaFree(d_z);
    cudaFree(d_K);
    cudaFree(d_m);
    cudaFree(d_c);
  };
  int rc = 0;
  const size_t n_cov = full_cov ? (size_t)Np * Np : (size_t)Np;
  if ((rc = upload(&d_V, V, (size_t)Np * N)) || (rc = upload(&d_z, z, N)) || (rc = upload(&d_K, Kss, (size_t)Np * Np))) {
    cleanup();
    return rc;
  }
  cudaError_t e = cudaMalloc((void**)&d_m, Np * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_c, n_cov * sizeof(double));
  if (e != cudaSuccess) {
    cleanup();
    return fail(-3, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  const long long warps = Np + (long long)n_cov;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  predict_reduce_kernel<<<blocks, 256>>>(d_V, d_z, d_K, Np, N, full_cov, d_m, d_c);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(mean_out, d_m, Np * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(cov_out, d_c, n_cov * sizeof(double), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(-2, "mrv_predict_reduce: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" int mrv_model_eval(int device, const double* t, const double* flags, int64_t N, const mrv_model_op* ops,
                              int n_ops, int nm, const double* modpar, int64_t W, int to_ecc, double* model_out,
                              double* modpar_phys_out) {
  if (!t || !modpar || !model_out) return fail(-1, "null argument");
  if (int rc = require_device(device)) return rc;
  ProblemDesc D;
  if (int rc = fill_desc(D, N, 0, 0, 2, ops, n_ops, nm, nullptr, 0)) return rc;
  bool kep = false;
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].type == MRV_MODEL_OFFSET && !flags) return fail(-1, "an OFFSET model op needs a flags array");
    kep |= ops[i].type == MRV_MODEL_KEPLER;
  }
  if (W <= 0) return 0;
  double *d_t = nullptr, *d_f = nullptr, *d_mp = nullptr, *d_out = nullptr, *d_E = nullptr, *d_phys = nullptr;
  int rc = 0;
  auto cleanup = [&]() {
    cudaFree(d_t);
    cudaFree(d_f);
    cudaFree(d_mp);
    cudaFree(d_out);
    cudaFree(d_E);
    cudaFree(d_phys);
  };
  if ((rc = upload(&d_t, t, N)) || (flags && (rc = upload(&d_f, flags, N))) ||
      (rc = upload(&d_mp, modpar, (size_t)W * nm))) {
    cleanup();
    return rc;
  }
  cudaError_t e = cudaMalloc((void**)&d_out, (size_t)W * N * sizeof(double));
  if (e == cudaSuccess && kep) e = cudaMalloc((void**)&d_E, (size_t)W * N * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_phys, std::max<size_t>(1, (size_t)W * nm) * sizeof(double));
  if (e != cudaSuccess) {
    cleanup();
    return fail(-3, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  mean_residual_kernel<<<(unsigned)W, 256>>>(D, d_t, nullptr, d_f, d_mp, nullptr, to_ecc, d_out, (int)N, d_E, d_phys,
                                             nullptr);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(model_out, d_out, (size_t)W * N * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && modpar_phys_out && nm > 0)
    e = cudaMemcpy(modpar_phys_out, d_phys, (size_t)W * nm * sizeof(double), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(-2, "mrv_model_eval: %s", cudaGetErrorString(e));
  return 0;
}

// Keplerian.ecc_anomaly / true_anomaly on caller-supplied vectors (reference models.py:618-669).
static int kepler_vector_op(int device, const double* in, int64_t N, double ecc, int max_itr, double* out, bool anomaly) {
  if (!in || !out) return fail(-1, "null argument");
  if (N < 0 || N > (int64_t)1 << 30) return fail(-1, "bad length %lld", (long long)N);
  if (int rc = require_device(device)) return rc;
  if (N == 0) return 0;
  double *d_in = nullptr, *d_out = nullptr;
  int rc = upload(&d_in, in, (size_t)N);
  if (rc) return rc;
  cudaError_t e = cudaMalloc((void**)&d_out, (size_t)N * sizeof(double));
  if (e == cudaSuccess) {
    if (anomaly) kepler_anomaly_kernel<<<1, 512>>>(d_in, (int)N, ecc, max_itr, d_out);
    else true_anomaly_kernel<<<(unsigned)std::min<int64_t>((N + 255) / 256, 1184), 256>>>(d_in, (int)N, ecc, d_out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(-2, "kepler vector op: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int mrv_kepler_anomaly(int device, const double* M, int64_t N, double ecc, int max_itr, double* E_out) {
  return kepler_vector_op(device, M, N, ecc, max_itr, E_out, true);
}

extern "C" int mrv_true_anomaly(int device, const double* E, int64_t N, double ecc, double* nu_out) {
  return kepler_vector_op(device, E, N, ecc, 0, nu_out, false);
}

static int dense_cov(int device, int kernel_id, int variant, const double* hp, int nh, const double* x1, const double* x2,
                     const double* dist_e, const double* dist_se, int64_t n1, int64_t n2, int diag_mode,
                     const double* diag_add, double* K_out) {
  if (!hp || !K_out) return fail(-1, "null argument");
  if (kernel_nh(kernel_id) < 0 || nh != kernel_nh(kernel_id)) return fail(-1, "kernel %d needs %d hyper-parameters, got %d", kernel_id, kernel_nh(kernel_id), nh);
  if ((diag_mode & 2) && !diag_add) return fail(-1, "diag_mode bit 1 needs diag_add");
  if (int rc = require_device(device)) return rc;
  if (n1 <= 0 || n2 <= 0) return 0;
  const bool from_dist = dist_e != nullptr;
  const size_t total = (size_t)n1 * n2;
  double *d_a = nullptr, *d_b = nullptr, *d_d = nullptr, *d_K = nullptr;
  int rc = 0;
  auto cleanup = [&]() {
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_d);
    cudaFree(d_K);
  };
  if (from_dist) {
    if ((rc = upload(&d_a, dist_e, total)) || (rc = upload(&d_b, dist_se, total))) {
      cleanup();
      return rc;
    }
  } else {
    if ((rc = upload(&d_a, x1, n1)) || (rc = upload(&d_b, x2, n2))) {
      cleanup();
      return rc;
    }
  }
  if ((diag_mode & 2) && (rc = upload(&d_d, diag_add, n2))) {
    cleanup();
    return rc;
  }
  cudaError_t e = cudaMalloc((void**)&d_K, total * sizeof(double));
  if (e != cudaSuccess) {
    cleanup();
    return fail(-3, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  const CovParams cp = make_cov_params(kernel_id, variant, hp);
  const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  covmatrix_kernel<<<blocks, 256>>>(cp, from_dist ? 1 : 0, d_a, n1, d_b, n2, d_a, d_b, diag_mode, d_d, d_K);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(K_out, d_K, total * sizeof(double), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(-2, "covmatrix: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int mrv_covmatrix(int device, int kernel_id, int kernel_variant, const double* hp, int nh, const double* x1,
                             int64_t n1, const double* x2, int64_t n2, int diag_mode, const double* diag_add,
                             double* K_out) {
  if (!x1 || !x2) return fail(-1, "null argument");
  return dense_cov(device, kernel_id, kernel_variant, hp, nh, x1, x2, nullptr, nullptr, n1, n2, diag_mode, diag_add, K_out);
}

extern "C" int mrv_covmatrix_from_dist(int device, int kernel_id, int kernel_variant, const double* hp, int nh,
                                       const double* dist_e, const double* dist_se, int64_t n1, int64_t n2, int diag_mode,
                                       const double* diag_add, double* K_out) {
  if (!dist_e || !dist_se) return fail(-1, "null argument");
  return dense_cov(device, kernel_id, kernel_variant, hp, nh, nullptr, nullptr, dist_e, dist_se, n1, n2, diag_mode,
                   diag_add, K_out);
}

extern "C" int mrv_distances(int device, const double* t1, int64_t n1, const double* t2, int64_t n2, double* dist_e_out,
                             double* dist_se_out) {
  if (!t1 || !t2 || !dist_e_out || !dist_se_out) return fail(-1, "null argument");
  if (int rc = require_device(device)) return rc;
  if (n1 <= 0 || n2 <= 0) return 0;
  const size_t total = (size_t)n1 * n2;
  double *d_a = nullptr, *d_b = nullptr, *d_e = nullptr, *d_s = nullptr;
  int rc = 0;
  auto cleanup = [&]() {
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_e);
    cudaFree(d_s);
  };
  if ((rc = upload(&d_a, t1, n1)) || (rc = upload(&d_b, t2, n2))) {
    cleanup();
    return rc;
  }
  cudaError_t e = cudaMalloc((void**)&d_e, total * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_s, total * sizeof(double));
  if (e != cudaSuccess) {
    cleanup();
    return fail(-3, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  distances_kernel<<<blocks, 256>>>(d_a, n1, d_b, n2, d_e, d_s);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(dist_e_out, d_e, total * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(dist_se_out, d_s, total * sizeof(double), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(-2, "distances: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Device-resident sampler loop (device_sampler.cuh): proposals, batched likelihood and accept / reject
// are enqueued back to back on the handle's stream for all iterations; the host only uploads the
// starting ensemble and the per-iteration split assignment and downloads the histories at the end.
// ------------------------------------------------------------------------------------------------
extern "C" int mrv_mcmc_device(mrv_problem* p, const double* hp0, const double* mp0, int64_t W, int64_t iterations,
                               double a, const uint8_t* hp_vary, const uint8_t* mp_vary, const uint8_t* split_all,
                               uint64_t seed, double* hp_hist, double* mp_hist, double* logl_hist, double* acc_hist,
                               int32_t* fail_info) {
  if (!p || !hp0 || !mp0 || !hp_vary || !mp_vary || !split_all || !hp_hist || !mp_hist || !logl_hist || !acc_hist || !fail_info)
    return fail(-1, "null argument");
  if (W <= 1 || iterations < 0) return fail(-1, "need W > 1 walkers and iterations >= 0");
  CU(cudaSetDevice(p->device));
  const ProblemDesc& D = p->desc;
  const int nh = D.nh, nm = D.nm;
  SamplerArgs A;
  memset(&A, 0, sizeof A);
  A.W = (int)W;
  A.nh = nh;
  A.nm = nm;
  A.a = a;
  A.rng.seed = seed;
  for (int i = 0; i < D.n_ops; ++i)
    if (D.ops[i].type == MRV_MODEL_KEPLER) A.kep[A.n_kep++] = D.ops[i].param_offset;
  for (int i = 0; i < nh; ++i) A.hp_vary[i] = hp_vary[i];
  for (int i = 0; i < nm; ++i) A.mp_vary[i] = mp_vary[i];

  const size_t rows = (size_t)iterations + 1;
  DevBuf b_hp0, b_mp0, b_l0, b_hp, b_mp, b_l, b_st, b_split, b_hh, b_mh, b_lh, b_ah, b_fail;
  DevBuf* all[] = {&b_hp0, &b_mp0, &b_l0, &b_hp, &b_mp, &b_l, &b_st, &b_split, &b_hh, &b_mh, &b_lh, &b_ah, &b_fail};
  auto done = [&](int code) {
    for (DevBuf* b : all) b->release();
    return code;
  };
  int rc = 0;
  if ((rc = b_hp0.ensure((size_t)W * nh * 8)) || (rc = b_mp0.ensure(std::max<size_t>(1, (size_t)W * nm) * 8)) ||
      (rc = b_l0.ensure((size_t)W * 8)) || (rc = b_hp.ensure((size_t)W * nh * 8)) ||
      (rc = b_mp.ensure(std::max<size_t>(1, (size_t)W * nm) * 8)) || (rc = b_l.ensure((size_t)W * 8)) ||
      (rc = b_st.ensure((size_t)W * 4)) || (rc = b_split.ensure(std::max<size_t>(1, (size_t)iterations * W))) ||
      (rc = b_hh.ensure(rows * W * nh * 8)) || (rc = b_mh.ensure(std::max<size_t>(1, rows * W * nm) * 8)) ||
      (rc = b_lh.ensure(rows * W * 8)) || (rc = b_ah.ensure(rows * W * 8)) || (rc = b_fail.ensure(8 * sizeof(int))))
    return done(rc);
  cudaStream_t st = p->stream;
  cudaError_t e = cudaMemcpyAsync(b_hp0.p, hp0, (size_t)W * nh * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && nm > 0) e = cudaMemcpyAsync(b_mp0.p, mp0, (size_t)W * nm * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && iterations > 0)
    e = cudaMemcpyAsync(b_split.p, split_all, (size_t)iterations * W, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(b_fail.p, 0, 8 * sizeof(int), st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_mcmc_device upload: %s", cudaGetErrorString(e)));
  // initial state = history row 0 (accepted = 1, as the reference records)
  if ((rc = run_logl(p, b_hp0.as<double>(), b_mp0.as<double>(), nullptr, W, 1, b_l0.as<double>(), b_st.as<int>(), st))) return done(rc);
  std::vector<int> st0((size_t)W);
  e = cudaMemcpyAsync(st0.data(), b_st.p, (size_t)W * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_mcmc_device initial evaluation: %s", cudaGetErrorString(e)));
  fail_info[0] = fail_info[1] = fail_info[2] = 0;
  for (int64_t w = 0; w < W; ++w)
    if (st0[(size_t)w] != 0) {
      fail_info[0] = 1;
      fail_info[1] = -1;
      fail_info[2] = st0[(size_t)w];
      return done(0);
    }
  e = cudaMemcpyAsync(b_hh.p, b_hp0.p, (size_t)W * nh * 8, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && nm > 0) e = cudaMemcpyAsync(b_mh.p, b_mp0.p, (size_t)W * nm * 8, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b_lh.p, b_l0.p, (size_t)W * 8, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_mcmc_device history init: %s", cudaGetErrorString(e)));
  {
    std::vector<double> ones((size_t)W, 1.0);
    e = cudaMemcpyAsync(b_ah.p, ones.data(), (size_t)W * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(fail(-2, "mrv_mcmc_device history init: %s", cudaGetErrorString(e)));
  }
  const unsigned blocks = (unsigned)((W + 127) / 128);
  for (int64_t it = 0; it < iterations; ++it) {
    stretch_propose_kernel<<<blocks, 128, 0, st>>>(A, (unsigned)it, b_split.as<unsigned char>() + (size_t)it * W,
                                                   b_hp0.as<double>(), b_mp0.as<double>(), b_hp.as<double>(),
                                                   b_mp.as<double>(), b_fail.as<int>() + 4);
    if ((rc = run_logl(p, b_hp.as<double>(), b_mp.as<double>(), nullptr, W, 1, b_l.as<double>(), b_st.as<int>(), st)))
      return done(rc);
    stretch_accept_kernel<<<blocks, 128, 0, st>>>(A, (unsigned)it, (long long)it + 1, b_hp.as<double>(), b_mp.as<double>(),
                                                  b_l.as<double>(), b_st.as<int>(), b_hp0.as<double>(),
                                                  b_mp0.as<double>(), b_l0.as<double>(), b_hh.as<double>(),
                                                  b_mh.as<double>(), b_lh.as<double>(), b_ah.as<double>(),
                                                  b_fail.as<int>());
    p->launches += 2;
  }
  e = cudaGetLastError();
  int fl[8] = {0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(hp_hist, b_hh.p, rows * W * nh * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && nm > 0) e = cudaMemcpyAsync(mp_hist, b_mh.p, rows * W * nm * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(logl_hist, b_lh.p, rows * W * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(acc_hist, b_ah.p, rows * W * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(fl, b_fail.p, sizeof fl, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(fail(-2, "mrv_mcmc_device: %s", cudaGetErrorString(e)));
  fail_info[0] = fl[0];
  fail_info[1] = fl[1];
  fail_info[2] = fl[2];
  if (fl[4] != 0) return done(fail(-5, "mrv_mcmc_device: proposal kernel could not find a %s", fl[4] == 1 ? "partner walker in another split" : "valid stretch for some walker"));
  return done(0);
}

extern "C" int mrv_gelman_rubin(int device, const double* chains, int64_t n_steps, int64_t n_chains, int64_t n_par,
                                int64_t burn_in, double* R_out) {
  if (!chains || !R_out) return fail(-1, "null argument");
  if (n_par <= 0) return 0;
  if (burn_in < 0 || n_steps - burn_in < 2) return fail(-1, "need at least 2 steps after the burn-in (steps=%lld, burn_in=%lld)", (long long)n_steps, (long long)burn_in);
  if (n_chains < 2) return fail(-1, "need at least 2 chains");
  if (int rc = require_device(device)) return rc;
  double *d_c = nullptr, *d_R = nullptr;
  int rc = 0;
  if ((rc = upload(&d_c, chains, (size_t)n_steps * n_chains * n_par))) return rc;
  cudaError_t e = cudaMalloc((void**)&d_R, n_par * sizeof(double));
  if (e == cudaSuccess) {
    gelman_rubin_kernel<<<(unsigned)n_par, 256>>>(d_c, n_steps, n_chains, n_par, burn_in, d_R);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(R_out, d_R, n_par * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_c);
  cudaFree(d_R);
  if (e != cudaSuccess) return fail(-2, "mrv_gelman_rubin: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" int mrv_set_option(mrv_problem* p, const char* key, int64_t value) {
  if (!p || !key) return fail(-1, "null argument");
  p->opt_gen++;   // a captured graph of the host-buffer call froze the old choices
  if (!strcmp(key, "path")) {
    if (value < 0 || value > 3) return fail(-1, "path must be 0 (auto), 1 (fused smem), 2 (blocked) or 3 (streamed panel)");
    p->path_opt = (int)value;
  } else if (!strcmp(key, "wave_walkers")) {
    p->wave_opt = value;
  } else if (!strcmp(key, "panel_tiles")) {
    p->panel_opt = (int)value;
  } else if (!strcmp(key, "panel_inner")) {
    p->panel_inner_opt = (int)value;
  } else if (!strcmp(key, "update_band")) {
    p->band_opt = (int)value;
  } else if (!strcmp(key, "cs_store")) {
    p->cs_store = (int)value;   // 1: streaming (evict-first) stores of updated C tiles
  } else if (!strcmp(key, "ensemble_walkers")) {
    if (value < 0) return fail(-1, "ensemble_walkers must be >= 0");
    p->ensemble_opt = value;
  } else if (!strcmp(key, "mean_threads")) {
    if (value != 0 && (value < 64 || value > 1024 || value % 32)) return fail(-1, "mean_threads must be 0 or a multiple of 32 in 64..1024");
    g_mean_threads_opt = (int)value;
  } else if (!strcmp(key, "mean_overlap")) {
    p->mean_overlap_opt = (int)value;   // 0: Keplerian mean kernel on the main stream, 1: next to the column-0 generation
  } else if (!strcmp(key, "small_minb")) {
    p->small_minb_opt = (int)value;
  } else if (!strcmp(key, "small_blk_min")) {
    p->small_blk_min = (int)value;
  } else if (!strcmp(key, "lookahead")) {
    p->lookahead_opt = (int)value;
  } else if (!strcmp(key, "substreams")) {
    p->substreams_opt = (int)value;
  } else if (!strcmp(key, "cuda_graph")) {
    p->graph_opt = value != 0;   // 1 (default): replay the host-buffer call of the one-CTA-per-walker paths as a CUDA graph
  } else if (!strcmp(key, "tiled_impl")) {
    if (value < 0 || value > 2) return fail(-1, "tiled_impl must be 0 (automatic), 1 (launch per tile column) or 2 (persistent dataflow kernel)");
    p->tiled_impl = (int)value;
  } else if (!strcmp(key, "potrf_impl")) {
    if (value < 0 || value > 1) return fail(-1, "potrf_impl must be 0 (blocked DMMA sweep) or 1 (rank-1 sweep)");
    p->potrf_impl = (int)value;
  } else if (!strcmp(key, "profile")) {
    cudaSetDevice(p->device);
    p->prof.reset();
    p->prof.on = value != 0;
  } else {
    return fail(-1, "unknown option '%s'", key);
  }
  return 0;
}

extern "C" int64_t mrv_get_info(mrv_problem* p, const char* key) {
  if (!p || !key) return -1;
  if (!strcmp(key, "path")) return p->last_path ? p->last_path : choose_path(p);
  if (!strcmp(key, "launches")) return p->launches;
  if (!strcmp(key, "graph_replays")) return p->graph_replays;
  if (!strcmp(key, "wave_walkers")) return p->wave_cur;
  if (!strcmp(key, "tiled_impl")) return use_dag(p, p->wave_cur) ? 2 : 1;
  if (!strcmp(key, "n_pad")) return p->n_pad;
  if (!strcmp(key, "panel_tiles")) return effective_panel(p);
  if (!strcmp(key, "small_n_max")) return SMALL_N_MAX;
  if (!strcmp(key, "mid_n_max")) return MID_N_MAX;
  if (!strncmp(key, "dag_t", 5)) {   // phase timers of the dataflow kernel (only with -DDAG_TIMING); index 16 resets
    unsigned long long v[16];
    cudaSetDevice(p->device);
    if (cudaMemcpyFromSymbol(v, g_dag_timing, sizeof v) != cudaSuccess) return -1;
    const int i = atoi(key + 5);
    if (i < 0 || i > 16) return -1;
    if (i == 16) {
      memset(v, 0, sizeof v);
      cudaMemcpyToSymbol(g_dag_timing, v, sizeof v);
      return 0;
    }
    return (int64_t)v[i];
  }
  if (!strncmp(key, "small_t", 7)) {   // phase timers of the fused small-N kernel (only with -DSMALL_TIMING); index 8 resets
    long long v[8];
    cudaSetDevice(p->device);
    if (cudaMemcpyFromSymbol(v, g_small_timing, sizeof v) != cudaSuccess) return -1;
    const int i = atoi(key + 7);
    if (i < 0 || i > 8) return -1;
    if (i == 8) {
      memset(v, 0, sizeof v);
      cudaMemcpyToSymbol(g_small_timing, v, sizeof v);
      return 0;
    }
    return v[i];
  }
  if (!strncmp(key, "mid_t", 5)) {   // phase timers of the streamed-panel kernel (only with -DMID_TIMING); reading resets
    long long v[32];
    cudaSetDevice(p->device);
    if (cudaMemcpyFromSymbol(v, g_mid_timing, sizeof v) != cudaSuccess) return -1;
    const int i = atoi(key + 5);
    if (i < 0 || i > 32) return -1;
    if (i == 32) {
      memset(v, 0, sizeof v);
      cudaMemcpyToSymbol(g_mid_timing, v, sizeof v);
      return 0;
    }
    return v[i];
  }
  if (!strncmp(key, "prof_", 5)) {
    // prof_<ns|alg|exec|n>_<category>; the caller must have synchronised the stream it launched on
    cudaSetDevice(p->device);
    p->prof.collect();
    const char* rest = key + 5;
    const char* us = strchr(rest, '_');
    if (!us) return -1;
    const std::string what(rest, us - rest);
    for (int c = 0; c < CAT_COUNT; ++c) {
      if (!strcmp(us + 1, kCatNames[c])) {
        if (what == "ns") return (int64_t)p->prof.ns[c];
        if (what == "alg") return (int64_t)p->prof.alg[c];
        if (what == "exec") return (int64_t)p->prof.exec[c];
        if (what == "n") return p->prof.n[c];
      }
    }
    return -1;
  }
  if (!strcmp(key, "workspace_bytes")) {
    int64_t s = 0;
    DevBuf* bufs[] = {&p->hp, &p->modpar, &p->model, &p->logl, &p->status, &p->resid, &p->escratch, &p->phys,
                      &p->nonfinite, &p->logdet, &p->quad, &p->fstatus, &p->zbuf, &p->factor, &p->counter, &p->midws,
                      &p->zcols, &p->rowdone, &p->phase, &p->cpar};
    for (DevBuf* b : bufs) s += (int64_t)b->bytes;
    return s;
  }
  return -1;
}

// ------------------------------------------------------------------------------------------------
// FP64 MMA issue-rate microbenchmark: every warp keeps 8 independent m16n8k16 accumulator chains.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
  double acc[8][4];
  double a[8], b[4];
#pragma unroll
  for (int v = 0; v < 8; ++v) a[v] = 1.0 + 1e-9 * (threadIdx.x + v);
#pragma unroll
  for (int v = 0; v < 4; ++v) b[v] = 1.0 - 1e-9 * (threadIdx.x + v);
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[c][v] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) dmma_16816(acc[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int v = 0; v < 4; ++v) s += acc[c][v];
  if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
  double acc[16];
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = (double)c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < 16; ++c) s += acc[c];
  if (s == 123.456) sink[0] = s;
}

static double time_peak(int variant, int iters) {
  double* sink = nullptr;
  if (cudaMalloc((void**)&sink, 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = 148 * 2;
  double best = -1.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (variant == 0) dmma_peak_kernel<<<blocks, 256>>>(iters, sink);
    else dfma_peak_kernel<<<blocks, 256>>>(iters, sink);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      best = -1.0;
      break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops;
    if (variant == 0) flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * (256 / 32) * blocks;
    else flops = 2.0 * 16.0 * iters * 256 * blocks;
    const double rate = flops / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return best;
}

// Covariance-evaluation rate: every thread keeps 8 independent evaluations of k(d) in flight (the ILP the
// factorisation kernels use), 16 warps per SM.  bench.py turns the rate into the cost of one matrix entry in
// FP64 flop-equivalents (vector-pipe peak / rate) for the small-N roofline.
__global__ void __launch_bounds__(512) cov_rate_kernel(CovParams cp, int iters, double* sink) {
  double d[8], d2[8], o[8], acc = 0.0;
#pragma unroll
  for (int e = 0; e < 8; ++e) d[e] = 0.37 * (threadIdx.x + 1) + 1.3 * e + 1e-3 * blockIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d[e] += 0.0131;
      d2[e] = d[e] * d[e];
    }
    cov_from_dist_n<8>(cp, d, d2, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += o[e];
  }
  if (acc == 123.456) sink[0] = acc;
}

// Same through the fast per-point evaluation the factorisation kernels use (cov_pair_n).
template <bool LOCKSTEP>
__global__ void __launch_bounds__(512) cov_pair_rate_kernel(CovParams cp, int iters, double* sink) {
  double ti[8], ci[8], si[8], tj[8], cj[8], sj[8], o[8], acc = 0.0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    ti[e] = 0.37 * (threadIdx.x + 1) + 1.3 * e + 1e-3 * blockIdx.x;
    tj[e] = 0.11 * e;
    cov_point_phase(cp, ti[e], ci[e], si[e]);
    cov_point_phase(cp, tj[e], cj[e], sj[e]);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ti[e] += 0.0131;
      ci[e] = fma(ci[e], 0.999999, 1e-7 * si[e]);   // keeps the inputs changing without leaving [-1, 1]
    }
    cov_pair_n<8, LOCKSTEP>(cp, ti, ci, si, tj, cj, sj, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += o[e];
  }
  if (acc == 123.456) sink[0] = acc;
}

extern "C" double mrv_cov_eval_rate(int device, int kernel_id, int kernel_variant, const double* hp, int nh) {
  if (require_device(device)) return -1.0;
  if (!hp || kernel_nh(kernel_id) < 0 || nh != kernel_nh(kernel_id)) return -1.0;
  const CovParams cp = make_cov_params(kernel_id, kernel_variant % 100, hp);   // variant + 100: the library-function form
  double* sink = nullptr;
  if (cudaMalloc((void**)&sink, 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = 148, iters = 8000;   // one 16-warp block per SM for both forms (equal occupancy)
  double best = -1.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (kernel_variant >= 200) cov_pair_rate_kernel<true><<<blocks, 512>>>(cp, iters, sink);    // per-point form, lock-step exp (tile kernels)
    else if (kernel_variant >= 100) cov_rate_kernel<<<blocks, 512>>>(cp, iters, sink);          // library sinpi / exp on distances
    else cov_pair_rate_kernel<false><<<blocks, 512>>>(cp, iters, sink);                         // per-point form, library exp
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      best = -1.0;
      break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = 8.0 * iters * 512.0 * blocks / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return best;
}

extern "C" double mrv_fp64_mma_peak(int device, int iters) {
  if (require_device(device)) return -1.0;
  if (iters <= 0) iters = 20000;
  return time_peak(0, iters);
}

extern "C" double mrv_fp64_fma_peak(int device, int iters) {
  if (require_device(device)) return -1.0;
  if (iters <= 0) iters = 20000;
  return time_peak(1, iters);
}
