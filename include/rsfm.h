/* rsfm.h -- C ABI of the B200 RSF-MCMC hot path (librsfm.so).
 *
 * The reference (SaumikDana/Bayesian-Markov-chain-Monte-Carlo) has no FFI layer:
 * its boundary is two duck-typed Python protocols.  This
 * header is what a Python facade binds with ctypes underneath those protocols;
 * every entry point cites the reference code it replaces.  Plain pointers and
 * sizes only: no torch types.  All `*_dev` pointers are CUDA device pointers
 * owned by the caller; `stream` is a cudaStream_t passed as void*.  Functions
 * return 0 on success and a negative rsfm_status on error, never throw, and do
 * not synchronise the stream unless stated.
 *
 * Devices and threads: a sampler lives on the device that was current in
 * rsfm_create; that device must be current in every later call on it
 * (rsfm_destroy switches by itself).  A sampler is used from one thread at a
 * time; different samplers and rsfm_forward_batch may be used concurrently.
 * rsfm_last_error is per thread.
 *
 * Layouts are chain-minor ("SoA") so that one warp = 32 consecutive chains reads
 * and writes 256 contiguous bytes:
 *     params   [P][C]          q0, proposals  [d][C] or [n_iters][d][C]
 *     series   [n_out][C]      (time-major acc output)
 *     samples  [n_iters][d][C]
 */
#ifndef RSFM_H
#define RSFM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSFM_ABI_VERSION 4
#define RSFM_MAX_PARAMS 3
#define RSFM_MAX_GROUPS 4

typedef enum {
    RSFM_OK = 0,
    RSFM_ERR_INVALID = -1,     /* bad argument */
    RSFM_ERR_CUDA = -2,        /* CUDA runtime error, see rsfm_last_error() */
    RSFM_ERR_NO_DEVICE = -3,   /* no sm_100 device: there is NO CPU fallback */
    RSFM_ERR_STATE = -4        /* sampler used before rsfm_init */
} rsfm_status;

/* load-point velocity */
enum { RSFM_LOAD_SINE_DECAY = 0,   /* RateStateModel.py:327-329 (reference) */
       RSFM_LOAD_VSTEP = 1,        /* piecewise-constant extension */
       RSFM_LOAD_TABLE = 2 };      /* tabulated extension: V_l/V_ref - 1 at t_start + i load_dt (load_table_dev,
                                      n_load_table entries), piecewise linear in between, constant beyond the ends */

/* state-evolution law */
enum { RSFM_LAW_AGING = 0,         /* Dieterich: theta' = 1 - v theta/Dc, RateStateModel.py:340 (reference) */
       RSFM_LAW_SLIP = 1 };        /* Ruina: theta' = -(v theta/Dc) ln(v theta/Dc) (extension; every step is scored
                                      with the general-range stages, about 3x the cost of the aging law) */

/* time-integration mode */
enum { RSFM_INTEG_PARITY = 0,      /* fresh dop853 call + hinit per output interval:
                                      scipy ode('dop853') as driven by RateStateModel.py:380-389 */
       RSFM_INTEG_CARRY = 1 };     /* carry h and the FSAL stage across output points */

/* proposal-covariance adaptation */
enum { RSFM_ADAPT_NONE = 0,        /* list-typed priors: update silently dead (quirk q2) */
       RSFM_ADAPT_COMPAT = 1,      /* dict-typed priors: MCMC.py:200-204 as written (quirk q3) */
       RSFM_ADAPT_POOLED = 2 };    /* covariance pooled over all chains (and ranks): rsfm_pooled_partials +
                                      rsfm_pooled_update, every step on the device */

/* observable scored by the sum of squares */
enum { RSFM_OBS_ACC = 0,           /* backward-difference acceleration, RateStateModel.py:388 (reference) */
       RSFM_OBS_MU = 1 };          /* friction coefficient mu_k = y[0], RateStateModel.py:385 (extension:
                                      the series the reference integrates and stores but never compares) */

/* which model constant the scalar (d = 1) chain samples ("named parameters (a, b, Dc, k1 ...)") */
enum { RSFM_PARAM_DC = 0,          /* Dc: what MCMC.SSqcalc sets on the model, MCMC.py:381 (reference) */
       RSFM_PARAM_K1 = 1 };        /* the radiation-damping coefficient k1 (RateStateModel.py:171, 349-353), with
                                      Dc = cfg->dc_fixed for every chain (extension) */

/* which instantiation of the solver kernels runs */
enum { RSFM_VARIANT_AUTO = 0,      /* stiff variant for RSFM_LOAD_VSTEP, default one otherwise */
       RSFM_VARIANT_DEFAULT = 1,
       RSFM_VARIANT_STIFF = 2 };

/* per-chain solver status word ("flag, not hang") */
enum { RSFM_CHAIN_OK = 0,
       RSFM_CHAIN_NMAX = 2,        /* dop853 idid = -2, "larger nsteps is needed" (quirk q9) */
       RSFM_CHAIN_HSMALL = 3,      /* dop853 idid = -3 */
       RSFM_CHAIN_NONFINITE = 4,
       RSFM_CHAIN_EARLY = 8 };     /* solver-internal: proposal rejected before the end of the series */

/* Model + solver constants.  Defaults (rsfm_cfg_defaults) are the reference's
 * RateStateModel.py:5-11,167-184 and the tolerances of :374. */
typedef struct rsfm_cfg {
    double  a, b, mu_ref, V_ref, k1;     /* :167-171; a, b are per-chain when sampled */
    double  t_start, t_final;            /* :174-175 */
    double  delta_t;                     /* :177, computed by the host exactly as the reference */
    double  mu_t_zero;                   /* :180 */
    double  vstep_period, vstep_factor;  /* RSFM_LOAD_VSTEP only; both must be positive (RSFM_ERR_INVALID otherwise) */
    double  rtol, atol;                  /* :374 */
    double  n0;                          /* MCMC.py:97 */
    double  lo[RSFM_MAX_PARAMS];         /* strict box prior, MCMC.py:98,318-320 */
    double  hi[RSFM_MAX_PARAMS];
    int32_t n_out;                       /* int(floor((t_final-t_start)/delta_t)), :358 (quirk q8) */
    int32_t nmax;                        /* scipy nsteps (500) */
    int32_t radiation_damping;           /* :183 */
    int32_t loading;                     /* RSFM_LOAD_* */
    int32_t integ_mode;                  /* RSFM_INTEG_* */
    int32_t n_params;                    /* d: 1 = (Dc), 3 = (a, b, Dc) */
    int32_t n_prior_len;                 /* len(qpriors): 3 list form, 2 dict form, MCMC.py:261 (q5) */
    int32_t adapt_interval;              /* MCMC.py:58 */
    int32_t adapt_mode;                  /* RSFM_ADAPT_* */
    int32_t spec_depth;                  /* speculative kernel of rsfm_run, 2^spec_depth lanes per chain: 0 auto (by
                                            chain count), 1 off, 2..5 forced; results never depend on it */
    int32_t observable;                  /* RSFM_OBS_* */
    int32_t solver_variant;              /* RSFM_VARIANT_* (tests / tuning; results agree within the parity gates) */
    int32_t stiff_exact;                 /* stiff variant: score every step that left the fast ranges with the
                                            general-range stages instead of taking exploding trial steps as rejected */
    int32_t block_threads;               /* threads per block of the one-thread-per-chain kernels: 0 auto, or
                                            32 / 64 / 96 / 128 (tuning; results never depend on it) */
    int32_t chain_groups;                /* RSFM_ADAPT_POOLED only: serve the chains by this many launches on the
                                            sampler's own streams (0 auto: up to 4 when >= 8,192 chains each, group
                                            boundaries on multiples of RSFM_POOL_GROUP; 1 off); see rsfm_join.
                                            A sampler the speculative kernel serves (spec_depth >= 2, or auto with
                                            n_params = 1 and <= 18,944 chains) has one group.
                                            Results never depend on it */
    int32_t round_packing;               /* d = 3 one-thread-per-chain kernel: pack the in-bounds proposals of a block into
                                            its lowest threads at the start of every solve round (0 auto = on for
                                            128-thread blocks and a series of <= 1,024 points, 1 off); results never
                                            depend on it */
    int32_t state_law;                   /* RSFM_LAW_* */
    int32_t n_load_table;                /* RSFM_LOAD_TABLE: entries (>= 2), spacing and DEVICE pointer of the table; */
    double  load_dt;                     /*   read by every call that gets this cfg and by rsfm_init, which tabulates */
    const double *load_table_dev;        /*   it for the sampler: keep it valid and unchanged while they run */
    int32_t sampled_param;               /* RSFM_PARAM_*.  RSFM_PARAM_K1 (n_params = 1, default solver variant only):
                                            every per-chain scalar this header calls "dc" -- dc_dev of
                                            rsfm_forward_batch, q / q0 / proposals / samples of the sampler, lo[0], hi[0]
                                            -- is k1, and Dc is dc_fixed */
    double  dc_fixed;                    /* RSFM_PARAM_K1 only: the common Dc (> 0) */
} rsfm_cfg;

typedef struct rsfm_sampler rsfm_sampler;   /* opaque; owns per-chain device state */

int          rsfm_abi_version(void);
const char  *rsfm_last_error(void);
void         rsfm_cfg_defaults(rsfm_cfg *cfg);

/* Number of CUDA devices usable by this library (compute capability 10.x). */
int          rsfm_device_count(void);

/* Release the per-process device caches (nominal loading tables of rsfm_forward_batch, the
 * rsfm_init workspace kept between samplers).  No reference counterpart: the reference holds no
 * device memory.  Safe at any time no rsfm call is in flight on another thread. */
int          rsfm_trim(void);

/* Batched RateStateModel.evaluate() (RateStateModel.py:188-395) + optional SSE
 * (MCMC.py:387).  dc_dev [C] is required; a_dev, b_dev [C] may be NULL (cfg->a,
 * cfg->b are used).  Outputs, each optional (NULL to skip):
 *   acc_out_dev [n_out][C]  the observable: backward-difference acceleration (:388), or mu_k (:385, entry 0 =
 *                           mu_ref, :367) with cfg->observable = RSFM_OBS_MU
 *   t_out_dev   [n_out][C]  output times as accumulated by the solver (:384)
 *   sse_out_dev [C]         sum_k (acc_k - data_k)^2, needs data_dev [n_out]
 *   status_dev  [C]         RSFM_CHAIN_*
 *   filled_dev  [C]         entries written before a failure (n_out when ok)
 *   nrhs_dev, nstep_dev [C] executed RHS evaluations / attempted steps */
int rsfm_forward_batch(const rsfm_cfg *cfg, int32_t C,
                       const double *dc_dev, const double *a_dev, const double *b_dev,
                       const double *data_dev,
                       double *acc_out_dev, double *t_out_dev, double *sse_out_dev,
                       int32_t *status_dev, int32_t *filled_dev,
                       uint64_t *nrhs_dev, uint64_t *nstep_dev, void *stream);

/* Sampler object = the state of C independent MCMC.sample() runs (MCMC.py:391-544).
 * chain_id0 is the global id of local chain 0 (Philox key = seed, counter =
 * global chain id, iteration, slot), so results do not depend on the sharding.
 * All device state lives in one buffer taken from (and, on rsfm_destroy, returned to)
 * the library's cache; rsfm_destroy synchronises the device first.  rsfm_init
 * synchronises `stream` before it returns. */
rsfm_sampler *rsfm_create(const rsfm_cfg *cfg, int32_t C, uint64_t seed, uint64_t chain_id0);
void          rsfm_destroy(rsfm_sampler *s);

/* compute_initial_covariance + first SSqcalc (MCMC.py:206-266, 464-468):
 * sigma2_0 = SSE(q0)/(N - n_prior_len), Vstart = sigma2_0 (X'X)^-1 from forward
 * differences with relative step 1e-6; proposal factor = chol(Vstart).
 * q0_dev [d][C]; data_dev [n_out] is copied into the sampler. */
int rsfm_init(rsfm_sampler *s, const double *q0_dev, const double *data_dev, void *stream);

/* n_iters iterations of the loop MCMC.py:494-527 for every chain, one launch.
 * Outputs (each optional): samples_out_dev [n_iters][d][C], sigma2_out_dev
 * [n_iters][C], accept_out_dev [n_iters][C].  draws_out_dev, if not NULL, receives
 * the random draws actually used, [n_iters][d+2][C] = (proposal[d], U, unit gamma),
 * so that the CPU oracle can replay the chain (U is NaN when not drawn, q10). */
int rsfm_run(rsfm_sampler *s, int32_t n_iters,
             double *samples_out_dev, double *sigma2_out_dev, uint8_t *accept_out_dev,
             double *draws_out_dev, void *stream);

/* Chain groups (RSFM_ADAPT_POOLED, large batches).  A launch of rsfm_run there is one adaptation interval long,
 * and the last, partly empty wave of its blocks would leave most SMs idle (1.15 waves at 65,536 chains: 18 % of the
 * launch).  The sampler therefore serves its chains in rsfm_chain_groups(s) groups, each by its own launch on a
 * stream the sampler owns; groups have no barrier in common, so one group's tail overlaps the others' next
 * launches.  ORDERING: with more than one group, the effects of rsfm_run (state and output buffers) are ordered on
 * the caller's `stream` only behind the next call on the same sampler that reads them back -- rsfm_pooled_partials
 * (the pooled-adaptation pipeline calls it after every rsfm_run), every rsfm_get_* / rsfm_set_*, rsfm_run_deterministic
 * -- or an explicit rsfm_join(s, stream).  rsfm_join also makes the groups' next launches wait for the work queued
 * on `stream` so far (a full meeting point, e.g. around a timed region).  No reference counterpart. */
int rsfm_chain_groups(const rsfm_sampler *s);
int rsfm_join(rsfm_sampler *s, void *stream);

/* log2 of the lanes per chain rsfm_run will use for this sampler (0 = plain one-thread-per-chain
 * kernel; g >= 1: 2^g lanes evaluate 2^g nodes of the tree of the chain's next iterations
 * concurrently, chosen best-first by a predictor of the accept / reject decisions).  Never
 * affects results. */
int rsfm_spec_depth(const rsfm_sampler *s);

/* Same loop with host-supplied randomness:
 * proposals_dev [n_iters][d][C] (absolute proposals, or standard normals z when
 * proposals_are_z != 0: q' = q + L z), uniforms_dev [n_iters][C] (consumed only
 * for in-bounds proposals), gammas_dev [n_iters][C] unit-scale Gamma((n0+N)/2). */
int rsfm_run_deterministic(rsfm_sampler *s, int32_t n_iters,
                           const double *proposals_dev, int32_t proposals_are_z,
                           const double *uniforms_dev, const double *gammas_dev,
                           double *samples_out_dev, double *sigma2_out_dev,
                           uint8_t *accept_out_dev, void *stream);

/* Per-chain state, for output, checkpoint/resume and tests.  Any pointer may be
 * NULL.  q [d][C], sse [C], sigma2 [C], chol [d(d+1)/2][C] (row-major lower
 * triangle; d = 1: the proposal VARIANCE as the reference stores it), accepted
 * [C], status [C], nrhs/nstep [C] cumulative executed work. */
int rsfm_get_state(rsfm_sampler *s, double *q_dev, double *sse_dev, double *sigma2_dev,
                   double *chol_dev, uint32_t *accepted_dev, int32_t *status_dev,
                   uint64_t *nrhs_dev, uint64_t *nstep_dev, void *stream);
int rsfm_set_state(rsfm_sampler *s, const double *q_dev, const double *sse_dev,
                   const double *sigma2_dev, const double *chol_dev, int64_t iteration,
                   void *stream);
int64_t rsfm_iteration(const rsfm_sampler *s);
/* RSFM_ADAPT_COMPAT only: the per-chain ring of the last adapt_interval samples, [adapt_interval][C], slot
 * (iteration % adapt_interval) = the sample of that iteration (MCMC.py:200: qparams[:, -adapt_interval:]).
 * Part of the state a checkpoint needs to continue a chain in the middle of an adaptation window. */
int rsfm_get_ring(rsfm_sampler *s, double *ring_dev, void *stream);
int rsfm_set_ring(rsfm_sampler *s, const double *ring_dev, void *stream);

/* Work totals since rsfm_init, summed over chains on the device and copied to
 * out_host[9] = (forward solves of the chains = in-bounds proposals decided, RHS
 * evaluations executed, steps attempted, accepted moves, chains with a non-zero
 * status, solves stopped early because rejection was already certain, solves
 * executed including speculative ones, RHS evaluations and steps of the deciding
 * solves only).  Synchronises the stream. */
int rsfm_get_totals(rsfm_sampler *s, uint64_t *out_host, void *stream);

/* Pooled adaptation (extension; generalises MCMC.py:162-204, 523-527).
 * Local sufficient statistics out_dev[1 + d + d(d+1)/2] = (n, sum q, sum qq^T lower) accumulated
 * over all chains and iterations since the last reset.  Stream-ordered, no synchronisation. */
int rsfm_get_suffstats(rsfm_sampler *s, double *out_dev, int32_t reset, void *stream);
/* Install one common proposal factor for every chain from HOST memory (synchronises the stream). */
int rsfm_set_proposal_chol(rsfm_sampler *s, const double *chol_host /* [d(d+1)/2] */, void *stream);

/* The same statistics as partial sums over fixed groups of RSFM_POOL_GROUP chains, aligned on the GLOBAL
 * chain id and reduced in a fixed order: out_dev [rsfm_pooled_groups(s)][RSFM_POOL_ROWS] with row layout
 * (n, sum q (d), sum qq^T (lower, row-major), zero padding).  Ranks all-gather these rows (NCCL) in rank order
 * = global chain order; every rank then holds the same array whatever the number of ranks, so the pooled
 * moments -- and with them the chains -- do not depend on the sharding (when shard boundaries are multiples of
 * RSFM_POOL_GROUP; otherwise they agree up to FP64 summation order).  reset != 0 clears the per-chain sums.
 * Stream-ordered, no synchronisation. */
#define RSFM_POOL_GROUP 1024
#define RSFM_POOL_ROWS 16
int rsfm_pooled_groups(const rsfm_sampler *s);
int rsfm_pooled_partials(rsfm_sampler *s, double *out_dev, int32_t reset, void *stream);

/* Device-side proposal update from gathered partials (no host round trip):
 *   moments_dev [1 + d + d(d+1)/2] += sum over parts_dev[n_parts][RSFM_POOL_ROWS] (sequential, fixed order)
 *       when accumulate != 0 (parts_dev may be NULL: nothing is added);
 *   when install != 0: V = (2.38^2/d) cov(moments) (ddof 1) (1 + 1e-10 on the diagonal), closed-form Cholesky
 *       (d <= 3); if V is finite and positive definite the common factor replaces every chain's proposal
 *       factor (d = 1: the proposal VARIANCE, as the reference stores it), otherwise the proposal stays (q4).
 *       install == 2 only FORMS the factor; a later rsfm_pooled_install(s, stream) puts it in place.  The pipeline
 *       of adaptation.PooledAdaptation forms the factor of interval j - 1 before it launches interval j and installs
 *       it before interval j + 1: with chain groups (rsfm_chain_groups) the install then waits for the forming of
 *       the factor only, never for the other groups' running launches.
 *   factor_out_dev (optional) [1 + d(d+1)/2] receives (valid ? 1 : 0, factor).
 * Stream-ordered, no synchronisation; every rank computes the same factor from the same moments. */
int rsfm_pooled_update(rsfm_sampler *s, const double *parts_dev, int32_t n_parts, double *moments_dev,
                       int32_t accumulate, int32_t install, double *factor_out_dev, void *stream);
int rsfm_pooled_install(rsfm_sampler *s, void *stream);

/* Diagnostics over a samples buffer [n][d][C] (device): per-chain mean, variance
 * (ddof 1) and autocovariance-based effective sample size (Geyer initial positive
 * sequence, lags < max_lag), for parameter p.  Outputs [C] each, optional. */
int rsfm_chain_diagnostics(const double *samples_dev, int32_t n, int32_t d, int32_t C, int32_t p,
                           int32_t max_lag, double *mean_dev, double *var_dev, double *ess_dev,
                           void *stream);

/* Gaussian kernel density estimate of samples_dev [n] on grid_dev [G] with the given
 * bandwidth (s.d. of the kernel); pdf_out_dev [G].  Replaces scipy.stats.gaussian_kde(...).pdf
 * as used by RSF.plot_dist (RSF.py:734-737); the Scott bandwidth is chosen by the caller. */
int rsfm_kde_grid(const double *samples_dev, int64_t n, const double *grid_dev, int32_t G, double bandwidth,
                  double *pdf_out_dev, void *stream);

/* Test hooks: what the device computes, element by element, for the parity tests against the CPU oracle.
 * rsfm_philox_raw: n Philox4x32-10 blocks, in_dev [n][6] = (counter[4], key[2]) -> out_dev [n][4] (the
 *   Random123 known-answer vectors go through the device function the kernels use).
 * rsfm_philox_draws: the draws iteration `iter0 + i` of chain `chain_id0 + c` consumes (csrc/philox.cuh stream
 *   layout): out_dev [n_iters][6][C] = (z0, z1, z2, U, unit Gamma(gamma_shape), number of gamma attempts).
 * rsfm_rhs_eval: friction(t, y) (RateStateModel.py:277-355) for n states, general == 0: the fast (series) form
 *   with its fallback, exactly as the solver evaluates a single RHS; general != 0: the reference's formulas.
 *   t/mu/theta/dc/a/b [n] -> out_dev [3][n] = (mu', theta', V'). */
int rsfm_philox_raw(const uint32_t *in_dev, uint32_t *out_dev, int32_t n, void *stream);
int rsfm_philox_draws(uint64_t seed, uint64_t chain_id0, int32_t C, uint32_t iter0, int32_t n_iters,
                      double gamma_shape, double *out_dev, void *stream);
int rsfm_rhs_eval(const rsfm_cfg *cfg, int32_t n, const double *t_dev, const double *mu_dev,
                  const double *theta_dev, const double *dc_dev, const double *a_dev, const double *b_dev,
                  int32_t general, double *out_dev, void *stream);

/* FP64 roofline denominator: runs a dependent-free DFMA kernel for about
 * `millis` ms on the current device and returns the sustained FP64 FMA rate in
 * FLOP/s (2 flops per DFMA) through *flops_out.  Synchronises. */
int rsfm_measure_fp64_peak(double millis, double *flops_out);

#ifdef __cplusplus
}
#endif
#endif /* RSFM_H */
