/*
 * magpy_rv_b200 -- C ABI of the B200-native GP log-likelihood hot path of MAGPy-RV.
 *
 * The reference (frescigno/magpy_rv 1.0.4) is pure Python and has no FFI; this header IS the
 * drop-in boundary.  Each entry point below names the reference interface it replaces
 * (file:line relative to /root/reference/).  Plain pointers and sizes only: no C++ or torch types
 * cross this boundary.  All floating point is IEEE binary64.
 *
 * Conventions
 *   - return value 0 = success; < 0 = API / CUDA error, text via mrv_last_error() (thread local).
 *   - per-walker `status` follows LAPACK potrf `info`: 0 ok, k > 0 = the leading minor of order k
 *     is not positive definite (the reference raises numpy.linalg.LinAlgError from cho_factor,
 *     gp_likelihood.py:271), -1 = non-finite covariance / residual input (cho_factor
 *     check_finite=True raises ValueError), -3 = internal error (the dataflow factorisation kernel gave up
 *     waiting for a tile dependency; never produced by valid input -- the Python shim raises RuntimeError).
 *   - host-pointer entry points copy in/out synchronously; *_device entry points take device
 *     pointers and enqueue on `stream` (a cudaStream_t passed as void*, NULL = default stream).
 *   - a handle is bound to one CUDA device and is not thread safe.
 *   - there is no CPU fallback: without a CUDA device every call fails with an error.
 */
#ifndef MAGPY_RV_B200_H
#define MAGPY_RV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRV_ABI_VERSION 1

/* Covariance kernels, in KERNELS order (src/magpy_rv/kernels.py:29-37). */
enum {
  MRV_KERNEL_COSINE = 0,         /* kernels.py:229   hp = [gp_amp, gp_per]                         */
  MRV_KERNEL_EXPSQUARED = 1,     /* kernels.py:356   hp = [gp_amp, gp_timescale]                   */
  MRV_KERNEL_EXPSINSQUARED = 2,  /* kernels.py:489   hp = [gp_amp, gp_timescale, gp_per]           */
  MRV_KERNEL_QUASIPER = 3,       /* kernels.py:629   hp = [gp_per, gp_perlength, gp_explength, gp_amp] */
  MRV_KERNEL_JITTERQUASIPER = 4, /* kernels.py:774   hp = [... QuasiPer ..., gp_jit]               */
  MRV_KERNEL_MATERN5 = 5,        /* kernels.py:909   hp = [gp_amp, gp_timescale]                   */
  MRV_KERNEL_MATERN3 = 6         /* kernels.py:1042  hp = [gp_amp, gp_timescale, gp_jit]           */
};
/* kernel_variant: 0 = arithmetic exactly as coded in the reference.  1 (Matern5 only) = the
 * opt-in textbook Matern 5/2, amp^2 (1 + sqrt5 d/ts + 5 d^2/(3 ts^2)) exp(-sqrt5 d/ts); the
 * reference's as-coded expression (kernels.py:909) is not a valid covariance. */

/* Mean-model ops, evaluated and summed in model_list order (src/magpy_rv/mcmc_aux.py:71-94). */
enum {
  MRV_MODEL_NONE = 0,   /* No_Model    models.py:266      1 param ("no")                      */
  MRV_MODEL_POLY = 1,   /* Polynomial  models.py:370      4 params a0..a3                     */
  MRV_MODEL_OFFSET = 2, /* Offset      models.py:467-479  1 param; applies where flags==flag_val */
  MRV_MODEL_KEPLER = 3  /* Keplerian   models.py:618-697  5 params P,K,ecc|Sk,omega|Ck,t0     */
};
typedef struct mrv_model_op {
  int32_t type;         /* MRV_MODEL_*                                                  */
  int32_t param_offset; /* first slot of this model in the flat model-parameter vector  */
  double flag_val;      /* OFFSET only: 1 for "offset"/"offset_0", j+1 for "offset_j"   */
} mrv_model_op;

/* Priors (src/magpy_rv/parameters.py:316-329, 383-402, 460-479, 527-544), summed in prior_list
 * order (gp_likelihood.py:393-400).  `source` 0 reads hp[index]; 1 reads the model-parameter
 * vector AFTER the Sk,Ck -> ecc,omega conversion of get_model(to_ecc=True) (mcmc_aux.py:84-88),
 * which is what the reference's priors see. */
enum { MRV_PRIOR_GAUSSIAN = 0, MRV_PRIOR_JEFFREY = 1, MRV_PRIOR_MODJEFFREY = 2, MRV_PRIOR_UNIFORM = 3 };
typedef struct mrv_prior {
  int32_t type;   /* MRV_PRIOR_*                                         */
  int32_t source; /* 0 = kernel hyper-parameter, 1 = model parameter     */
  int32_t index;  /* slot in hp / model-parameter vector                 */
  int32_t pad;
  double p0, p1, p2; /* GAUSSIAN mu,sigma | JEFFREY min,max | MODJEFFREY min,max,knee | UNIFORM min,max */
} mrv_prior;

typedef struct mrv_problem mrv_problem;

/* Build a likelihood problem: uploads t, y, yerr (and flags, may be NULL) once and fixes the
 * kernel, the mean-model op list and the prior list.  Replaces the per-walker object churn of
 * MCMC.compute (src/magpy_rv/mcmc.py:421-458: par_create/mod_create/parameter/GPLikelihood) and
 * GPLikelihood.__init__ (gp_likelihood.py:64-146).  nh / nm = number of kernel hyper-parameters /
 * model parameters per walker. */
int mrv_problem_create(mrv_problem** out, int device, const double* t, const double* y,
                       const double* yerr, const double* flags, int64_t N, int kernel_id,
                       int kernel_variant, int nh, const mrv_model_op* ops, int n_ops, int nm,
                       const mrv_prior* priors, int n_priors);
void mrv_problem_destroy(mrv_problem* p);

/* log-likelihood + log-priors of W walkers.  Replaces the body of MCMC.compute
 * (mcmc.py:410-461) and the initial loop of MCMC.__init__ (mcmc.py:221-261), i.e. per walker
 * get_model(..., to_ecc) -> GPLikelihood.LogL(prior_list) (gp_likelihood.py:250-278, 357-402).
 * hp [W, nh], modpar [W, nm] row-major; with to_ecc != 0 the ecc/omega slots of every Keplerian
 * hold Sk, Ck.  logL_out [W], status_out [W]. */
int mrv_logl_batch(mrv_problem* p, const double* hp, const double* modpar, int64_t W, int to_ecc,
                   double* logL_out, int32_t* status_out);
int mrv_logl_batch_device(mrv_problem* p, const double* d_hp, const double* d_modpar, int64_t W,
                          int to_ecc, double* d_logL_out, int32_t* d_status_out, void* stream);

/* Walker sharding inside the library (one host thread, one handle per device; SURVEY.md 8(b) `mrv_shard_init`).
 * mrv_shard_create replicates the problem on `n_dev` devices; mrv_shard_logl_batch cuts the ensemble into contiguous
 * blocks, enqueues every block on its device and collects log L / status in walker order -- the sharded form of
 * MCMC.compute (mcmc.py:410-461) for callers without a process group (the torch.distributed / NCCL form is
 * magpy_rv_b200/sharding.py).  Results are bitwise those of a single device. */
typedef struct mrv_shard mrv_shard;
int mrv_shard_create(mrv_shard** out, int n_dev, const int* devices, const double* t, const double* y,
                     const double* yerr, const double* flags, int64_t N, int kernel_id, int kernel_variant, int nh,
                     const mrv_model_op* ops, int n_ops, int nm, const mrv_prior* priors, int n_priors);
void mrv_shard_destroy(mrv_shard* sh);
int mrv_shard_count(mrv_shard* sh);
mrv_problem* mrv_shard_part(mrv_shard* sh, int i);   /* borrowed: options / info of one device's handle */
int mrv_shard_logl_batch(mrv_shard* sh, const double* hp, const double* modpar, int64_t W, int to_ecc,
                         double* logL_out, int32_t* status_out);

/* Same, but the caller supplies the mean-model curve(s) model_y [W, N] instead of the op list
 * (GPLikelihood accepts an arbitrary model_y, gp_likelihood.py:64,140-141); priors read modpar
 * as given (no conversion). */
int mrv_logl_given_model(mrv_problem* p, const double* hp, const double* modpar,
                         const double* model_y, int64_t W, double* logL_out, int32_t* status_out);

/* Forward half of cho_solve for W (hyper-parameter set, right-hand side) pairs: z_b = L_b^-1 rhs_b
 * with K_b = L_b L_b^T built from hp_b and the problem's t / yerr, plus ln det K_b.  The building
 * block of GPLikelihood.predict (gp_likelihood.py:406-481: mean = Ks K^-1 y = (L^-1 Ks^T)^T (L^-1 y),
 * cov = Kss - (L^-1 Ks^T)^T (L^-1 Ks^T)); the reference factorises K twice there (:434, :439).
 * hp [W, nh], rhs [W, N]; z_out [W, N], logdet_out [W] (may be NULL), status_out [W]. */
int mrv_solve_batch(mrv_problem* p, const double* hp, const double* rhs, int64_t W, double* z_out,
                    double* logdet_out, int32_t* status_out);

/* The same for R right-hand sides and ONE hyper-parameter set hp [nh]: K is factorised once and every right-hand side
 * is swept through the stored factor (what GPLikelihood.predict needs: L^-1 [y | Ks^T], gp_likelihood.py:429-441).
 * rhs [R, N]; z_out [R, N], logdet_out [1] (may be NULL), status_out [1]. */
int mrv_solve_multi(mrv_problem* p, const double* hp, const double* rhs, int64_t R, double* z_out,
                    double* logdet_out, int32_t* status_out);

/* Prediction reductions on the device: V [Np, N] (rows L^-1 Ks_p^T), z [N] (= L^-1 y), Kss [Np, Np]:
 * mean_out[p] = V_p . z;  cov_out = Kss - V V^T ([Np, Np]) when full_cov, else the Np variances
 * Kss_pp - |V_p|^2 (gp_likelihood.py:435-441, 476-481). */
int mrv_predict_reduce(int device, const double* V, const double* z, const double* Kss, int64_t Np,
                       int64_t N, int full_cov, double* mean_out, double* cov_out);

/* Mean-model curves for W parameter vectors over an arbitrary time vector (no handle): get_model
 * (mcmc_aux.py:52-97) and the Model.model() methods (models.py:266, 370, 467-479, 674-697).
 * flags may be NULL when no OFFSET op is present.  model_out [W, N]; modpar_phys_out (may be NULL)
 * [W, nm] receives the parameter vector after the optional Sk,Ck -> ecc,omega conversion. */
int mrv_model_eval(int device, const double* t, const double* flags, int64_t N,
                   const mrv_model_op* ops, int n_ops, int nm, const double* modpar, int64_t W,
                   int to_ecc, double* model_out, double* modpar_phys_out);

/* Keplerian.ecc_anomaly(M, ecc, max_itr) (models.py:618-650): Newton iteration of Kepler's equation on the
 * whole vector M [N] with the reference's stopping rule ||E - E0||_1 <= 1e-10; E_out [N]. */
int mrv_kepler_anomaly(int device, const double* M, int64_t N, double ecc, int max_itr, double* E_out);

/* Keplerian.true_anomaly(E, ecc) (models.py:653-669): nu = 2 atan(sqrt((1 + ecc) / (1 - ecc)) tan(E / 2)). */
int mrv_true_anomaly(int device, const double* E, int64_t N, double ecc, double* nu_out);

/* Dense covariance K [n1, n2] row-major between two time vectors: GPLikelihood.compute_kernel
 * (gp_likelihood.py:173-230) = compute_distances (kernels.py:51-78) + compute_covmatrix.
 * diag_mode bit 0: add gp_jit^2 (JitterQuasiPer / Matern3, kernels.py:778-783, 1044-1050) on
 * i == j; bit 1: then add diag_add[j] (the caller's errors^2, length n2) on i == j
 * (kernels.py:234-237 etc.).  Only meaningful for n1 == n2; the reference's broadcasting corner
 * cases for non-square K are resolved by the Python host. */
int mrv_covmatrix(int device, int kernel_id, int kernel_variant, const double* hp, int nh,
                  const double* x1, int64_t n1, const double* x2, int64_t n2, int diag_mode,
                  const double* diag_add, double* K_out);

/* Same map applied to caller-supplied dense distance matrices: Kernel.compute_covmatrix(dist_e,
 * dist_se, errors) (kernels.py:205, 332, 464, 603, 747, 885, 1017). */
int mrv_covmatrix_from_dist(int device, int kernel_id, int kernel_variant, const double* hp, int nh,
                            const double* dist_e, const double* dist_se, int64_t n1, int64_t n2,
                            int diag_mode, const double* diag_add, double* K_out);

/* |t1_i - t2_j| and (t1_i - t2_j)^2, both [n1, n2] row-major: compute_distances (kernels.py:51-78). */
int mrv_distances(int device, const double* t1, int64_t n1, const double* t2, int64_t n2,
                  double* dist_e_out, double* dist_se_out);

/* Device-resident sampler (opt-in, NOT a parity entry): the whole loop of run_MCMC (mcmc.py:800-830:
 * split_step -> compute -> compare -> reset, same stretch move, validity redraws and accept rule) runs
 * on the GPU with a counter-based Philox stream instead of the host's `random` / `np.random`, so
 * nothing crosses PCIe between iterations.  hp0 [W, nh] / mp0 [W, nm]: starting ensemble (Sk, Ck in
 * the ecc / omega slots); hp_vary [nh] / mp_vary [nm]: 0 keeps a parameter fixed; split_all
 * [iterations, W]: split index of every walker per iteration (the reference's shuffled
 * arange(W) % n_splits); histories [(iterations + 1), W, nh | nm | 1 | 1] as MCMC.hparameter_list,
 * model_parameter_list, logL_list, accepted.  fail_info[0] != 0: a likelihood evaluation reported
 * status fail_info[2] at iteration fail_info[1] (-1 = the starting ensemble) -- where the reference
 * raises LinAlgError / ValueError. */
int mrv_mcmc_device(mrv_problem* p, const double* hp0, const double* mp0, int64_t W, int64_t iterations,
                    double a, const uint8_t* hp_vary, const uint8_t* mp_vary, const uint8_t* split_all,
                    uint64_t seed, double* hp_hist, double* mp_hist, double* logl_hist, double* acc_hist,
                    int32_t* fail_info);

/* Gelman-Rubin statistic per parameter over the walker chains, chains [n_steps, n_chains, n_par]
 * row-major (the layout of MCMC.hparameter_list / model_parameter_list), the first burn_in steps
 * dropped.  NOT a parity entry: the reference's gelman_rubin_calc (mcmc.py:571-662) mislabels the axes
 * and drops into pdb; this computes the statistic it intends,
 * R = ((1 - 1/L) W + B/L) / W, and 1 for a parameter that never moves (mcmc.py:632-638). */
int mrv_gelman_rubin(int device, const double* chains, int64_t n_steps, int64_t n_chains, int64_t n_par,
                     int64_t burn_in, double* R_out);

/* Host-only helper (no GPU work): the sequential scan of one stretch-move split, MCMC.split_step
 * (mcmc.py:322-376).  Chains chain_begin..n_chains-1 consume the uniforms u[0..n_u) in order,
 * z = (u (a-1) + 1)^2 / a, proposal = Xj + dX z, redrawing while a hyper-parameter is negative or
 * mcmc_aux.parameter_check (mcmc_aux.py:187-231) fails for a Keplerian starting at one of
 * kepler_offsets.  Xj_* / dX_* / new_* are [n_chains, nh|nm] row-major, z_out [n_chains].  Returns
 * the number of uniforms consumed by COMPLETED chains; *chains_done = first chain not completed
 * (== n_chains when all are done; otherwise call again with more uniforms). */
int64_t mrv_stretch_scan(const double* u, int64_t n_u, int64_t chain_begin, int64_t n_chains, int32_t nh,
                         int32_t nm, const double* Xj_hp, const double* dX_hp, const double* Xj_mod,
                         const double* dX_mod, double a, const int32_t* kepler_offsets, int32_t n_kepler,
                         int64_t* chains_done, double* z_out, double* new_hp, double* new_mod);

/* One whole split of MCMC.split_step on the ensemble arrays (reference mcmc.py:343-394): chain c is walker
 * own[c], stretches towards walker partner[c]; the proposal (own value where vary == 0) is stored at row
 * dest[c] of hp_out [W, nh] / mod_out [W, nm], in chain order; z_out is [n_chains].  Same uniform-buffer
 * protocol and return value as mrv_stretch_scan; returns -1 when nh or nm exceeds 64. */
int64_t mrv_stretch_split(const double* u, int64_t n_u, int64_t chain_begin, int64_t n_chains, int32_t nh,
                          int32_t nm, const double* hp0, const double* mod0, const int64_t* own,
                          const int64_t* partner, const int64_t* dest, const uint8_t* hp_vary,
                          const uint8_t* mod_vary, double a, const int32_t* kepler_offsets, int32_t n_kepler,
                          int64_t* chains_done, double* z_out, double* hp_out, double* mod_out);

/* Introspection / tuning (no reference counterpart). */
enum { MRV_PATH_AUTO = 0, MRV_PATH_FUSED_SMEM = 1, MRV_PATH_BLOCKED = 2, MRV_PATH_MID = 3 };
/* Option keys: "path" (MRV_PATH_*; automatic by N and ensemble size), "tiled_impl" (0 automatic = persistent dataflow kernel for
 * 3 ... 128 tile columns, 1 launch-per-column loop, 2 dataflow kernel), "wave_walkers", "ensemble_walkers" (size of the whole
 * ensemble when this handle evaluates one shard of it: kernel selection then does not depend on the number of shards),
 * "cuda_graph" (replay of the host-buffer call), "profile" (CUDA-event timers per kernel category, read with "prof_*" infos),
 * and tuning knobs of the launch-per-column path / the fused kernel ("panel_tiles", "panel_inner", "update_band", "substreams",
 * "potrf_impl", "mean_overlap", "mean_threads", "small_minb", "small_blk_min").  Results never depend on them beyond the
 * documented 1e-9 (bitwise for everything but "path" and "potrf_impl"). */
int mrv_set_option(mrv_problem* p, const char* key, int64_t value);
int64_t mrv_get_info(mrv_problem* p, const char* key); /* "path", "launches", "wave_walkers", "n_pad", "workspace_bytes" */
int mrv_device_count(void);
const char* mrv_last_error(void);
int mrv_abi_version(void);

/* FP64 tensor-pipe (DMMA) issue-rate microbenchmark: returns achieved FLOP/s on `device`, or a
 * negative value on error.  Used by bench.py as the measured FP64 roofline denominator
 * (MEASURED_PEAKS.json has no FP64 entry). */
double mrv_fp64_mma_peak(int device, int iters);
/* Same for the vector FP64 pipe (DFMA), for comparison. */
double mrv_fp64_fma_peak(int device, int iters);
/* Covariance evaluations per second of kernel `kernel_id` (kernels.py:229...1042 bodies, 8 evaluations in flight per
 * thread, 16 warps per SM): bench.py states the generation cost of one matrix entry as vector-pipe peak / this rate
 * (FP64 flop-equivalents per entry) in the small-N roofline.  < 0 on error. */
double mrv_cov_eval_rate(int device, int kernel_id, int kernel_variant, const double* hp, int nh);

#ifdef __cplusplus
}
#endif
#endif /* MAGPY_RV_B200_H */
