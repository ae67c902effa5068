// rsfm_kernels.cu -- CUDA kernels and the C ABI (include/rsfm.h) of the RSF-MCMC
// hot path for B200 (sm_100a).  There is no CPU fallback: every entry point that
// computes needs a compute-capability-10 device and fails loudly otherwise.
//
// Kernels
//   rsf_forward_kernel   batched RateStateModel.evaluate() (+ SSE)      a9, a10, a11, a4
//   rsf_init_kernel      compute_initial_covariance + first SSqcalc     a8
//   rsf_mcmc_kernel      K fused iterations of MCMC.sample's loop       a1-a7
//   suffstats_kernel     pooled (n, sum q, sum qq^T) over chains        8e
//   chain_diag_kernel    per-chain mean / var / ESS                     8d (ESS/s)
//   dfma_peak_kernel     FP64 roofline denominator
// (a-numbers: SURVEY.md section 8a.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <new>

#include <cuda_runtime.h>

#include "rsfm.h"
#include "rsfm_device.cuh"
#include "philox.cuh"

using namespace rsfm;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, const char *detail)
{
    snprintf(g_err, sizeof(g_err), fmt, detail ? detail : "");
    return code;
}
#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return set_err(RSFM_ERR_CUDA, #expr ": %s", cudaGetErrorString(e__)); \
    } while (0)

extern "C" int rsfm_abi_version(void) { return RSFM_ABI_VERSION; }
extern "C" const char *rsfm_last_error(void) { return g_err; }

extern "C" void rsfm_cfg_defaults(rsfm_cfg *c)
{
    memset(c, 0, sizeof(*c));
    c->a = 0.011; c->b = 0.014; c->mu_ref = 0.6; c->V_ref = 1.0; c->k1 = 1.0e-7;     // RateStateModel.py:5-9
    c->t_start = 0.0; c->t_final = 50.0;                                             // :10-11
    c->delta_t = (c->t_final - c->t_start) / 500;                                    // :177
    c->n_out = (int32_t)floor((c->t_final - c->t_start) / c->delta_t);               // :358
    c->mu_t_zero = 0.6;
    c->vstep_period = 1000.0; c->vstep_factor = 10.0;
    c->rtol = 1e-6; c->atol = 1e-10; c->nmax = 500;                                  // :374, scipy nsteps
    c->n0 = 0.01;                                                                    // MCMC.py:97
    for (int i = 0; i < RSFM_MAX_PARAMS; i++) { c->lo[i] = 0.0; c->hi[i] = 10000.0; }
    c->radiation_damping = 1;
    c->loading = RSFM_LOAD_SINE_DECAY;
    c->integ_mode = RSFM_INTEG_PARITY;
    c->n_params = 1;
    c->n_prior_len = 3;
    c->adapt_interval = 10;
    c->adapt_mode = RSFM_ADAPT_NONE;
    c->spec_depth = 0;
}

extern "C" int rsfm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

static int require_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return set_err(RSFM_ERR_NO_DEVICE, "no CUDA device: librsfm has no CPU fallback%s", "");
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
        cudaGetLastError();
        return set_err(RSFM_ERR_NO_DEVICE, "current device is not sm_100 (B200): librsfm has no other code path%s", "");
    }
    return RSFM_OK;
}

static int check_cfg(const rsfm_cfg *c)
{
    if (!c) return set_err(RSFM_ERR_INVALID, "cfg is NULL%s", "");
    if (c->n_out < 1 || c->nmax < 1 || !(c->delta_t > 0.0)) return set_err(RSFM_ERR_INVALID, "bad grid in cfg%s", "");
    if (c->n_params != 1 && c->n_params != 3) return set_err(RSFM_ERR_INVALID, "n_params must be 1 or 3%s", "");
    if (c->loading != RSFM_LOAD_SINE_DECAY && c->loading != RSFM_LOAD_VSTEP && c->loading != RSFM_LOAD_TABLE)
        return set_err(RSFM_ERR_INVALID, "bad loading selector%s", "");
    if (c->loading == RSFM_LOAD_TABLE && (!c->load_table_dev || c->n_load_table < 2 || !(c->load_dt > 0.0)))
        return set_err(RSFM_ERR_INVALID, "RSFM_LOAD_TABLE needs load_table_dev, n_load_table >= 2 and load_dt > 0%s", "");
    if (c->state_law != RSFM_LAW_AGING && c->state_law != RSFM_LAW_SLIP)
        return set_err(RSFM_ERR_INVALID, "bad state_law selector%s", "");
    if (c->integ_mode != RSFM_INTEG_PARITY && c->integ_mode != RSFM_INTEG_CARRY)
        return set_err(RSFM_ERR_INVALID, "bad integ_mode%s", "");
    if (c->loading == RSFM_LOAD_VSTEP && !(c->vstep_period > 0.0))
        return set_err(RSFM_ERR_INVALID, "vstep_period must be positive%s", "");
    if (c->loading == RSFM_LOAD_VSTEP && !(c->vstep_factor > 0.0))
        return set_err(RSFM_ERR_INVALID, "vstep_factor must be positive%s", "");
    if (!(c->rtol > 0.0) || !(c->atol >= 0.0)) return set_err(RSFM_ERR_INVALID, "rtol must be positive and atol non-negative%s", "");
    if (!(c->a > 0.0) || !(c->V_ref > 0.0)) return set_err(RSFM_ERR_INVALID, "a and V_ref must be positive%s", "");
    // the per-chain sample ring of the reference's windowed update is the only user of adapt_interval in here
    // (the reference accepts any value and, with list priors, never uses it: MCMC.py:58, 523-527)
    if (c->adapt_mode == RSFM_ADAPT_COMPAT && (c->adapt_interval < 2 || c->adapt_interval > 64))
        return set_err(RSFM_ERR_INVALID, "adapt_interval must be in [2, 64] with RSFM_ADAPT_COMPAT%s", "");
    if (c->adapt_mode != RSFM_ADAPT_NONE && c->adapt_mode != RSFM_ADAPT_COMPAT && c->adapt_mode != RSFM_ADAPT_POOLED)
        return set_err(RSFM_ERR_INVALID, "bad adapt_mode%s", "");
    if (c->observable != RSFM_OBS_ACC && c->observable != RSFM_OBS_MU)
        return set_err(RSFM_ERR_INVALID, "bad observable selector%s", "");
    if (c->solver_variant < RSFM_VARIANT_AUTO || c->solver_variant > RSFM_VARIANT_STIFF)
        return set_err(RSFM_ERR_INVALID, "bad solver_variant%s", "");
    if (c->block_threads != 0 && c->block_threads != 32 && c->block_threads != 64 && c->block_threads != 96 &&
        c->block_threads != 128)
        return set_err(RSFM_ERR_INVALID, "block_threads must be 0, 32, 64, 96 or 128%s", "");
    if (c->spec_depth < 0 || c->spec_depth > 5) return set_err(RSFM_ERR_INVALID, "spec_depth must be in [0, 5]%s", "");
    if (c->round_packing < 0 || c->round_packing > 1) return set_err(RSFM_ERR_INVALID, "round_packing must be 0 (auto) or 1 (off)%s", "");
    if (c->chain_groups < 0 || c->chain_groups > RSFM_MAX_GROUPS)
        return set_err(RSFM_ERR_INVALID, "chain_groups must be in [0, 4]%s", "");
    if (c->sampled_param != RSFM_PARAM_DC && c->sampled_param != RSFM_PARAM_K1)
        return set_err(RSFM_ERR_INVALID, "bad sampled_param selector%s", "");
    if (c->sampled_param == RSFM_PARAM_K1) {
        if (c->n_params != 1) return set_err(RSFM_ERR_INVALID, "RSFM_PARAM_K1 samples one scalar: n_params must be 1%s", "");
        if (!(c->dc_fixed > 0.0)) return set_err(RSFM_ERR_INVALID, "RSFM_PARAM_K1 needs dc_fixed > 0%s", "");
        if (c->solver_variant == RSFM_VARIANT_STIFF)
            return set_err(RSFM_ERR_INVALID, "RSFM_PARAM_K1 runs the default solver variant only%s", "");
    }
    return RSFM_OK;
}

static ModelK make_model(const rsfm_cfg *c)
{
    ModelK M;
    memset(&M, 0, sizeof(M));          // padding too: ModelK is a memcmp key of the nominal-table cache
    M.mu_ref = c->mu_ref; M.V_ref = c->V_ref; M.k1 = c->k1; M.t_start = c->t_start;
    M.delta_t = c->delta_t; M.mu_t_zero = c->mu_t_zero;
    M.rtol = c->rtol; M.atol = c->atol; M.vstep_period = c->vstep_period; M.vstep_factor = c->vstep_factor;
    M.n_out = c->n_out; M.nmax = c->nmax; M.damping = c->radiation_damping; M.loading = c->loading;
    M.integ_mode = c->integ_mode;
    M.vstep_lnf = (c->loading == RSFM_LOAD_VSTEP) ? log(c->vstep_factor) : 0.0;
    M.vstep_rfac = (c->loading == RSFM_LOAD_VSTEP) ? 1.0 / c->vstep_factor : 1.0;
    M.vstep_rper = (c->loading == RSFM_LOAD_VSTEP) ? 1.0 / c->vstep_period : 1.0;
    M.stiff_exact = c->stiff_exact != 0;
    M.observable = c->observable;
    M.state_law = c->state_law;
    if (c->loading == RSFM_LOAD_TABLE) { M.load_n = c->n_load_table; M.load_dt = c->load_dt; M.load_tab = c->load_table_dev; }
    if (c->sampled_param == RSFM_PARAM_K1) M.dc_fixed = c->dc_fixed;
    return M;
}

// Which instantiation of the solver runs: the kernels exist twice.  VS = false keeps the general
// (state-by-state) interval path exactly as the non-stiff configurations were tuned with it (rsf_interval_plain);
// VS = true uses rsf_interval_general, which re-bases the friction law on the current load level, does not score
// the exploding trial steps of the stability-limited regime and has a cheaper step-size controller -- the variant
// for velocity-step loading (cfg 4), where nearly every interval is a general one.
// cfg->solver_variant overrides the choice (tests, tuning).
static bool stiff_variant(const rsfm_cfg *c)
{
    if (c->solver_variant == RSFM_VARIANT_DEFAULT) return false;
    if (c->sampled_param == RSFM_PARAM_K1) return false;    // k1 chains exist as default-variant kernels only
    if (c->solver_variant == RSFM_VARIANT_STIFF) return true;
    return c->loading == RSFM_LOAD_VSTEP;
}

// The stiff variant runs one-warp blocks: with a streamed series (cfg 4) the block barriers at tile boundaries
// would otherwise make every warp of a block wait for the slowest lane of all of them.  (Measured and dropped,
// profiles/microbench/forward_stiff_r1b.txt: spreading the chains over more, partly empty warps -- the kernel is
// instruction-fetch bound, a second warp per sub-partition doubles the time -- and a rolled stage loop.)
static const int STIFF_BLOCK = 32;

// chains per block: small batches are spread over more SMs (the kernel is latency
// bound there), large batches use 128-thread blocks.
static int pick_block(int C, const rsfm_cfg *c)
{
    if (c->block_threads) return c->block_threads;       // tuning / experiments only
    if (C <= 148 * 32) return 32;
    if (C <= 148 * 64 * 2) return 64;
    return 128;
}

// ---------------------------------------------------------------------------
// nominal loading table: stage values of L(t) on the time grid every non-stiff chain walks
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) loading_table_kernel(ModelK M, double *__restrict__ nom)
{
    // keys first: the same recurrence the solver runs (t_{k+1} = t_k + ((t_k + delta_t) - t_k))
    if (threadIdx.x == 0) {
        double t = M.t_start;
        for (int k = 1; k < M.n_out; k++) {
            const double xend = t + M.delta_t;
            const double h = xend - t;
            nom[(size_t)k * NOM_STRIDE + 11] = t;
            nom[(size_t)k * NOM_STRIDE + 12] = h;
            t = t + h;
        }
    }
    __syncthreads();
    const long long total = (long long)(M.n_out - 1) * 11;
    for (long long j = threadIdx.x; j < total; j += blockDim.x) {
        const int k = 1 + (int)(j / 11), i = (int)(j % 11);
        const double t = nom[(size_t)k * NOM_STRIDE + 11], h = nom[(size_t)k * NOM_STRIDE + 12];
        nom[(size_t)k * NOM_STRIDE + i] = loading_of(M, __dadd_rn(t, __dmul_rn(TB.c[i], h)));
    }
}

// Small per-process cache of nominal tables for the stateless rsfm_forward_batch (a sampler owns its
// own).  Entries are keyed by the model constants and the device; they are immutable once built, so
// concurrent readers on any stream only need to wait for the build event.
struct NomEntry { ModelK M; int device; double *ptr; cudaEvent_t ready; unsigned long long stamp; };
static NomEntry g_nom[8];
static int g_nom_n = 0;
static unsigned long long g_nom_clock = 0;
static std::mutex g_nom_mu;

static int get_nominal_table(const ModelK &M, cudaStream_t stream, const double **out)
{
    std::lock_guard<std::mutex> lock(g_nom_mu);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    for (int i = 0; i < g_nom_n; i++) {
        if (g_nom[i].device == dev && memcmp(&g_nom[i].M, &M, sizeof(ModelK)) == 0) {
            g_nom[i].stamp = ++g_nom_clock;
            CUDA_TRY(cudaStreamWaitEvent(stream, g_nom[i].ready, 0));
            *out = g_nom[i].ptr;
            return RSFM_OK;
        }
    }
    int slot = g_nom_n;
    if (g_nom_n == 8) {                       // evict the least recently used entry (rare: needs a full sync)
        slot = 0;
        for (int i = 1; i < 8; i++) if (g_nom[i].stamp < g_nom[slot].stamp) slot = i;
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(g_nom[slot].ptr);
        cudaEventDestroy(g_nom[slot].ready);
    } else {
        g_nom_n++;
    }
    NomEntry &e = g_nom[slot];
    memset(&e, 0, sizeof(e));
    memcpy(&e.M, &M, sizeof(ModelK));
    e.device = dev; e.stamp = ++g_nom_clock;
    if (cudaMalloc((void **)&e.ptr, sizeof(double) * NOM_STRIDE * (size_t)M.n_out) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming) != cudaSuccess) {
        g_nom_n = slot == g_nom_n - 1 ? g_nom_n - 1 : g_nom_n;
        e.device = -1;
        return set_err(RSFM_ERR_CUDA, "nominal table allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    loading_table_kernel<<<1, 1024, 0, stream>>>(M, e.ptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e.ready, stream));
    *out = e.ptr;
    return RSFM_OK;
}

// Device-buffer cache for the sampler state slab (rsfm_create) and rsfm_init's base trajectories.  A
// fresh cudaMalloc/cudaFree pair of more than a few MB goes through the driver's VM mapping path and
// was measured at 3-600 ms (cfg 2) up to 2 s (131,072 chains) per sampler on the public-API call
// (profiles/microbench/e2e_jitter3.py, cfg5_breakdown.py); buffers up to SCRATCH_KEEP_BYTES are
// therefore kept per process and handed out again (rsfm_trim releases them).  A buffer is only
// returned to the cache after the work that used it has been synchronised, so the next user may be on
// any stream.
static const size_t SCRATCH_KEEP_BYTES = (size_t)2 << 30;
struct ScratchEntry { int device; void *ptr; size_t bytes; bool busy; };
static ScratchEntry g_scratch[6];
static std::mutex g_scratch_mu;

static int scratch_acquire(size_t bytes, void **out)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        ScratchEntry *best = nullptr;              // best fit: the smallest idle buffer that is large enough
        for (ScratchEntry &e : g_scratch)
            if (e.ptr && !e.busy && e.device == dev && e.bytes >= bytes && (!best || e.bytes < best->bytes)) best = &e;
        if (best) {
            best->busy = true;
            *out = best->ptr;
            return RSFM_OK;
        }
    }
    CUDA_TRY(cudaMalloc(out, bytes));
    return RSFM_OK;
}

// caller guarantees no work that touches `p` is still in flight
static void scratch_release(void *p, size_t bytes)
{
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    void *victim = p;
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        for (ScratchEntry &e : g_scratch)
            if (e.ptr == p) { e.busy = false; return; }
        if (bytes <= SCRATCH_KEEP_BYTES) {
            ScratchEntry *slot = nullptr;
            for (ScratchEntry &e : g_scratch) if (!e.ptr) { slot = &e; break; }
            if (!slot)                             // replace the smallest idle buffer if this one is larger
                for (ScratchEntry &e : g_scratch)
                    if (!e.busy && e.bytes < bytes && (!slot || e.bytes < slot->bytes)) slot = &e;
            if (slot) {
                victim = slot->ptr;
                int victim_dev = slot->device;
                slot->device = dev; slot->ptr = p; slot->bytes = bytes; slot->busy = false;
                if (victim && victim_dev != dev) {     // free on the owning device
                    cudaSetDevice(victim_dev); cudaFree(victim); cudaSetDevice(dev);
                    victim = nullptr;
                }
            }
        }
    }
    if (victim) cudaFree(victim);
}

extern "C" int rsfm_trim(void)
{
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        for (ScratchEntry &e : g_scratch)
            if (e.ptr && !e.busy) {
                cudaSetDevice(e.device); cudaFree(e.ptr);
                e.ptr = nullptr; e.bytes = 0;
            }
    }
    {
        std::lock_guard<std::mutex> lock(g_nom_mu);
        for (int i = 0; i < g_nom_n; i++) {
            if (g_nom[i].device < 0) continue;
            cudaSetDevice(g_nom[i].device);
            cudaDeviceSynchronize();
            cudaFree(g_nom[i].ptr);
            cudaEventDestroy(g_nom[i].ready);
        }
        g_nom_n = 0;
    }
    cudaSetDevice(dev);
    return RSFM_OK;
}

// ---------------------------------------------------------------------------
// forward batch
// ---------------------------------------------------------------------------
// MB = resident blocks per SM the register budget is sized for: 3 (<= 168 registers) pays at saturating
// batch sizes (+5 %), 1 (no cap) is 8 % faster when the batch is small and the kernel latency-bound.
// K1P: dc_in holds the per-chain k1 (RSFM_PARAM_K1), Dc = M.dc_fixed
template <int MB, bool VS, bool K1P = false>
__global__ void __launch_bounds__(128, VS ? 1 : MB)
rsf_forward_kernel(const __grid_constant__ ModelK M, int C, double a0, double b0, const double *__restrict__ dc_in,
                   const double *__restrict__ a_in, const double *__restrict__ b_in,
                   const double *__restrict__ data, double *__restrict__ acc_out,
                   double *__restrict__ t_out, double *__restrict__ sse_out, int32_t *__restrict__ status_out,
                   int32_t *__restrict__ filled_out, unsigned long long *__restrict__ nrhs_out,
                   unsigned long long *__restrict__ nstep_out, const double *__restrict__ nom)
{
    __shared__ __align__(128) double s_tile[2 * SERIES_TILE];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_ltab[(VS ? 1 : 4) * LTAB_STRIDE];       // VS kernels run one-warp blocks (STIFF_BLOCK)
    __shared__ double s_lpriv[11 * (VS ? 32 : 128)];
    LoadScratch lscr;
    lscr.tab = s_ltab; lscr.priv = s_lpriv; lscr.nom = nom;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = c < C;
    const int cc = active ? c : C - 1;
    const double dc = dc_in[cc];
    const double a = a_in ? a_in[cc] : a0;
    const double b = b_in ? b_in[cc] : b0;
    SeriesStage series;
    series.begin(s_tile, s_bar, data, M.n_out);
    series.start_solve();
    SolveOut o = rsf_solve<VS, false, K1P>(M, a, b, dc, active, series, lscr, acc_out ? acc_out + cc : nullptr, nullptr,
                                           (size_t)C, 1.0, nullptr, t_out ? t_out + cc : nullptr);
    if (active) {
        if (sse_out) sse_out[c] = o.sse;
        if (status_out) status_out[c] = o.status;
        if (filled_out) filled_out[c] = o.filled;
        if (nrhs_out) nrhs_out[c] = o.nrhs;
        if (nstep_out) nstep_out[c] = o.nstep;
    }
}

extern "C" int rsfm_forward_batch(const rsfm_cfg *cfg, int32_t C, const double *dc_dev, const double *a_dev,
                                  const double *b_dev, const double *data_dev, double *acc_out_dev,
                                  double *t_out_dev, double *sse_out_dev, int32_t *status_dev, int32_t *filled_dev,
                                  uint64_t *nrhs_dev, uint64_t *nstep_dev, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (C < 1 || !dc_dev) return set_err(RSFM_ERR_INVALID, "forward_batch: C < 1 or dc_dev NULL%s", "");
    if (sse_out_dev && !data_dev) return set_err(RSFM_ERR_INVALID, "forward_batch: sse_out needs data%s", "");
    if (data_dev && ((uintptr_t)data_dev & 15)) return set_err(RSFM_ERR_INVALID, "forward_batch: data_dev must be 16-byte aligned%s", "");
    rc = require_device();
    if (rc) return rc;
    const int block = pick_block(C, cfg);
    const int grid = (C + block - 1) / block;
    const ModelK M = make_model(cfg);
    // The nominal load table (fast interval) serves the aging law; a tabulated load is not cached per process (the
    // cache is keyed by the model constants, and the caller may rewrite the table behind the same pointer): those
    // batches run the state-by-state interval path, whose warp-shared stage table is built from the table on the fly.
    const double *nom = nullptr;
    if (cfg->state_law == RSFM_LAW_AGING && cfg->loading != RSFM_LOAD_TABLE) {
        rc = get_nominal_table(M, (cudaStream_t)stream, &nom);
        if (rc) return rc;
    }
    const bool vs = stiff_variant(cfg);
    const int vgrid = (C + STIFF_BLOCK - 1) / STIFF_BLOCK;
#define RSFM_FWD(MB, VS)                                                                                              \
    rsf_forward_kernel<MB, VS><<<VS ? vgrid : grid, VS ? STIFF_BLOCK : block, 0, (cudaStream_t)stream>>>(            \
        M, C, cfg->a, cfg->b, dc_dev, a_dev, b_dev, sse_out_dev ? data_dev : nullptr, acc_out_dev, t_out_dev,         \
        sse_out_dev, status_dev, filled_dev, (unsigned long long *)nrhs_dev, (unsigned long long *)nstep_dev, nom)
    if (cfg->sampled_param == RSFM_PARAM_K1)
        rsf_forward_kernel<1, false, true><<<grid, block, 0, (cudaStream_t)stream>>>(
            M, C, cfg->a, cfg->b, dc_dev, a_dev, b_dev, sse_out_dev ? data_dev : nullptr, acc_out_dev, t_out_dev,
            sse_out_dev, status_dev, filled_dev, (unsigned long long *)nrhs_dev, (unsigned long long *)nstep_dev, nom);
    else if (vs) RSFM_FWD(1, true);
    else if (C <= 148 * 4 * 32 * 2) RSFM_FWD(1, false);
    else RSFM_FWD(3, false);
#undef RSFM_FWD
    CUDA_TRY(cudaGetLastError());
    return RSFM_OK;
}

// ---------------------------------------------------------------------------
// sampler state
// ---------------------------------------------------------------------------
struct SamplerDev {
    double *q;          // [d][C]
    double *sse;        // [C]
    double *sigma2;     // [C]
    double *chol;       // [d(d+1)/2][C]; d = 1: proposal variance (as the reference stores it)
    double *ring;       // [adapt_interval][C] last samples, COMPAT adaptation (d = 1)
    double *suff;       // [d + d(d+1)/2][C] per-chain sums for POOLED adaptation
    double *fit;        // [SPEC_FIT or SPEC_FIT3][C] predictor state of the speculative kernel (few chains), else unused
    double *data;       // [n_out] padded to an even count
    double *nom;        // [n_out][NOM_STRIDE] nominal loading table (loading_table_kernel)
    unsigned int *accepted;        // [C]
    int *status;                   // [C] sticky OR of RSFM_CHAIN_* of all solves
    unsigned long long *nrhs;      // [C]
    unsigned long long *nstep;     // [C]
    unsigned long long *nsolve;    // [C] forward solves executed
    unsigned long long *nearly;    // [C] of which stopped early (rejection certain)
    unsigned long long *nexec;     // [C] solves executed, speculative ones included
    unsigned long long *urhs;      // [C] RHS evaluations of the solves that decided a proposal (no speculation waste)
    unsigned long long *ustep;     // [C] steps of those solves
};

struct rsfm_sampler {
    rsfm_cfg cfg;
    int C;
    uint64_t seed, chain_id0;
    int64_t iteration;
    int64_t suff_count;            // iterations accumulated in suff
    int initialised;
    int device;
    SamplerDev d;
    double *nom_buf;               // the sampler's nominal load table (d.nom points at it, or is NULL for the slip law)
    double *scratch;               // [n_out][C] base trajectory for rsfm_init
    size_t scratch_bytes;
    char *slab;                    // one device buffer behind every array of `d`, `totals` and `reduce_out`
    size_t slab_bytes;
    double *reduce_out;            // [16] device scratch for suffstats
    unsigned long long *totals;    // [8] device scratch for rsfm_get_totals
    // Chain groups (pooled adaptation, large batches): the chains are served by n_groups launches on the sampler's
    // own streams, which have no barrier in common -- the tail of one group's 10-iteration launch overlaps the next
    // launches of the others (a 65,536-chain launch is 1.15 waves and left the SMs idle 18 % of the time).
    int n_groups;                  // 1 = off
    cudaStream_t gstream[RSFM_MAX_GROUPS];
    cudaEvent_t gevent[RSFM_MAX_GROUPS];
    cudaEvent_t fork_event, fac_event[2];
    int need_fork;                 // state was written on the caller's stream since the groups last met it
    int in_flight;                 // group work the caller's stream has not been ordered behind yet
    double *fac_buf[2];            // proposal factor of the pooled update, double-buffered for the group installs
    int fac_parity;
    int staged;                    // slot of a factor formed by rsfm_pooled_update(install = 2) and not installed yet, or -1
    int fit_rows;                  // rows of d.fit (SPEC_FIT, SPEC_FIT3, or 0 when the sampler keeps no predictor state)
};

static int tri(int d) { return d * (d + 1) / 2; }

// Predictor state of rsf_mcmc_spec_kernel per chain: moments S0..S4 of x and T0..T2 of x^k y (0-7), coefficients
// c0..c2 (8-10), mean squared residual (11), "fit valid" (12), "residual known" (13), centre 1/q_c, SS_c and scale
// of the fit variables (14-16), state (17: 0 = not set up, 1 = in use, -1 = not usable for this chain).
static const int SPEC_FIT = 18;
// d = 3: the fit is a full quadratic in t = (1/a, b/a, 1/Dc) (centred, scaled): normal matrix P (55 entries, upper
// triangle by rows; 0-54), right-hand side R (55-64), coefficients (65-74), mean squared residual (75), "fit valid"
// (76), "residual known" (77), centre (78-80) and scale (81-83) of t, SS_c (84), state as above (85), padding.
static const int SPEC_FIT3 = 88;
static const int SPEC_MAX_CHAINS = 148 * 4 * 32 * 2 / 2;      // most chains the speculative kernel is chosen for

// ---------------------------------------------------------------------------
// chain groups
// ---------------------------------------------------------------------------
// Number of groups a sampler runs with: RSFM_ADAPT_POOLED only (there the launches are one adaptation interval
// long), the one-thread-per-chain default kernel, group boundaries on multiples of RSFM_POOL_GROUP in the GLOBAL
// chain numbering (each group then owns whole rows of the pooled partial sums) and at least 8,192 chains per group.
static int pick_groups(const rsfm_cfg *c, int C, uint64_t chain_id0)
{
    if (c->adapt_mode != RSFM_ADAPT_POOLED || c->chain_groups == 1) return 1;
    if (stiff_variant(c) || c->n_out > 2 * SERIES_TILE) return 1;
    if (chain_id0 % RSFM_POOL_GROUP != 0) return 1;
    // samplers the speculative kernel serves (it runs on the caller's stream) have no groups
    if (c->spec_depth >= 2 || (c->spec_depth == 0 && c->n_params == 1 && C <= SPEC_MAX_CHAINS)) return 1;
    int g = c->chain_groups >= 2 ? c->chain_groups : 4;
    while (g > 1 && (C % (g * RSFM_POOL_GROUP) != 0 || C / g < 8192)) g--;
    return g;
}

// Order the caller's stream behind everything the groups have been given so far.
static int join_groups(rsfm_sampler *s, cudaStream_t stream)
{
    if (s->n_groups <= 1 || !s->in_flight) return RSFM_OK;
    for (int g = 0; g < s->n_groups; g++) {
        CUDA_TRY(cudaEventRecord(s->gevent[g], s->gstream[g]));
        CUDA_TRY(cudaStreamWaitEvent(stream, s->gevent[g], 0));
    }
    s->in_flight = 0;
    return RSFM_OK;
}

// Order the groups behind what the caller's stream holds (state written by rsfm_init / rsfm_set_*).
static int fork_groups(rsfm_sampler *s, cudaStream_t stream)
{
    if (s->n_groups <= 1 || !s->need_fork) return RSFM_OK;
    CUDA_TRY(cudaEventRecord(s->fork_event, stream));
    for (int g = 0; g < s->n_groups; g++) CUDA_TRY(cudaStreamWaitEvent(s->gstream[g], s->fork_event, 0));
    s->need_fork = 0;
    return RSFM_OK;
}

extern "C" rsfm_sampler *rsfm_create(const rsfm_cfg *cfg, int32_t C, uint64_t seed, uint64_t chain_id0)
{
    if (check_cfg(cfg)) return nullptr;
    if (C < 1) { set_err(RSFM_ERR_INVALID, "rsfm_create: C < 1%s", ""); return nullptr; }
    if (require_device()) return nullptr;
    rsfm_sampler *s = new (std::nothrow) rsfm_sampler();
    if (!s) { set_err(RSFM_ERR_INVALID, "rsfm_create: out of host memory%s", ""); return nullptr; }
    memset(s, 0, sizeof(*s));
    s->cfg = *cfg; s->C = C; s->seed = seed; s->chain_id0 = chain_id0;
    cudaGetDevice(&s->device);
    const int d = cfg->n_params;
    const size_t Cz = (size_t)C;
    // one slab for all per-chain state (sub-arrays 256-byte aligned): a single (cached) allocation
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_q = take(sizeof(double) * d * Cz), o_sse = take(sizeof(double) * Cz), o_s2 = take(sizeof(double) * Cz);
    const size_t o_chol = take(sizeof(double) * tri(d) * Cz), o_ring = take(sizeof(double) * (cfg->adapt_mode == RSFM_ADAPT_COMPAT ? cfg->adapt_interval : 1) * Cz);
    const size_t o_suff = take(sizeof(double) * (d + tri(d)) * Cz), o_data = take(sizeof(double) * ((size_t)cfg->n_out + 2));
    const bool want_fit = (d == 1 && C <= SPEC_MAX_CHAINS) || (d == 3 && C <= SPEC_MAX_CHAINS / 2);
    s->fit_rows = want_fit ? (d == 1 ? SPEC_FIT : SPEC_FIT3) : 0;
    const size_t o_fit = take(sizeof(double) * (want_fit ? s->fit_rows * Cz : 1));
    const size_t o_nom = take(sizeof(double) * NOM_STRIDE * (size_t)cfg->n_out);
    const size_t o_acc = take(sizeof(unsigned int) * Cz), o_status = take(sizeof(int) * Cz);
    size_t o_cnt[7];
    for (size_t &o : o_cnt) o = take(sizeof(unsigned long long) * Cz);
    const size_t o_totals = take(sizeof(unsigned long long) * 16), o_reduce = take(sizeof(double) * 16);
    const size_t o_fac2 = take(sizeof(double) * 16);
    s->slab_bytes = off;
    if (scratch_acquire(off, (void **)&s->slab) != RSFM_OK || cudaMemset(s->slab, 0, off) != cudaSuccess) {
        set_err(RSFM_ERR_CUDA, "rsfm_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        rsfm_destroy(s);
        return nullptr;
    }
    char *b = s->slab;
    s->d.q = (double *)(b + o_q); s->d.sse = (double *)(b + o_sse); s->d.sigma2 = (double *)(b + o_s2);
    s->d.chol = (double *)(b + o_chol); s->d.ring = (double *)(b + o_ring); s->d.suff = (double *)(b + o_suff);
    s->d.fit = want_fit ? (double *)(b + o_fit) : nullptr;
    s->d.data = (double *)(b + o_data); s->nom_buf = (double *)(b + o_nom); s->d.nom = s->nom_buf;
    s->d.accepted = (unsigned int *)(b + o_acc); s->d.status = (int *)(b + o_status);
    s->d.nrhs = (unsigned long long *)(b + o_cnt[0]); s->d.nstep = (unsigned long long *)(b + o_cnt[1]);
    s->d.nsolve = (unsigned long long *)(b + o_cnt[2]); s->d.nearly = (unsigned long long *)(b + o_cnt[3]);
    s->d.nexec = (unsigned long long *)(b + o_cnt[4]); s->d.urhs = (unsigned long long *)(b + o_cnt[5]);
    s->d.ustep = (unsigned long long *)(b + o_cnt[6]);
    s->totals = (unsigned long long *)(b + o_totals); s->reduce_out = (double *)(b + o_reduce);
    s->fac_buf[0] = s->reduce_out; s->fac_buf[1] = (double *)(b + o_fac2);
    s->n_groups = pick_groups(cfg, C, chain_id0);
    s->need_fork = 1;
    s->staged = -1;
    {
        bool ok = cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&s->fac_event[0], cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&s->fac_event[1], cudaEventDisableTiming) == cudaSuccess;
        for (int g = 0; g < s->n_groups && ok && s->n_groups > 1; g++)
            ok = cudaStreamCreateWithFlags(&s->gstream[g], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&s->gevent[g], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            set_err(RSFM_ERR_CUDA, "rsfm_create: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            rsfm_destroy(s);
            return nullptr;
        }
    }
    return s;
}

extern "C" void rsfm_destroy(rsfm_sampler *s)
{
    if (!s) return;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != s->device) cudaSetDevice(s->device);
    if (s->slab || s->scratch) cudaDeviceSynchronize();      // nothing may still be using the buffers
    for (int g = 0; g < RSFM_MAX_GROUPS; g++) {
        if (s->gstream[g]) cudaStreamDestroy(s->gstream[g]);
        if (s->gevent[g]) cudaEventDestroy(s->gevent[g]);
    }
    if (s->fork_event) cudaEventDestroy(s->fork_event);
    if (s->fac_event[0]) cudaEventDestroy(s->fac_event[0]);
    if (s->fac_event[1]) cudaEventDestroy(s->fac_event[1]);
    if (s->scratch) scratch_release(s->scratch, s->scratch_bytes);     // only after a failed rsfm_init
    if (s->slab) scratch_release(s->slab, s->slab_bytes);
    if (cur != s->device) cudaSetDevice(cur);
    delete s;
}

extern "C" int64_t rsfm_iteration(const rsfm_sampler *s) { return s ? s->iteration : -1; }

// ---------------------------------------------------------------------------
// init: compute_initial_covariance + SSqcalc(qstart)   (MCMC.py:245-266, 468)
// ---------------------------------------------------------------------------
// pass 0: base solve at q0, trajectory to scratch, SSE -> sse, sigma2_0
// pass j = 1..d: solve at q0 with parameter j-1 scaled by (1 + 1e-6); accumulates
//                the column products needed for X'X.  d = 1 keeps everything in
//                one pass; d = 3 stores the three sensitivity columns.
template <int D, bool VS, bool K1P = false>
__global__ void __launch_bounds__(128)
rsf_init_kernel(const __grid_constant__ ModelK M, int C, int pass, double a0, double b0, int n_prior_len, SamplerDev S,
                double *__restrict__ scratch /* [(1+ (D>1?D:0))][n_out][C] */)
{
    __shared__ __align__(128) double s_tile[2 * SERIES_TILE];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_ltab[(VS ? 1 : 4) * LTAB_STRIDE];       // VS kernels run one-warp blocks (STIFF_BLOCK)
    __shared__ double s_lpriv[11 * (VS ? 32 : 128)];
    LoadScratch lscr;
    lscr.tab = s_ltab; lscr.priv = s_lpriv; lscr.nom = S.nom;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = c < C;
    const int cc = active ? c : C - 1;
    double q[D];
#pragma unroll
    for (int j = 0; j < D; j++) q[j] = S.q[(size_t)j * C + cc];
    double fd_den = 1.0;
    if (pass > 0) {
        q[pass - 1] *= (1 + 1e-6);                    // MCMC.py:251
        fd_den = q[pass - 1] * 1e-6;                  // :264
    }
    const double a = (D == 3) ? q[0] : a0;
    const double b = (D == 3) ? q[1] : b0;
    const double dc = q[D - 1];
    SeriesStage series;
    series.begin(s_tile, s_bar, pass == 0 ? S.data : nullptr, M.n_out);
    series.start_solve();
    const size_t plane = (size_t)M.n_out * C;
    double xtx = 0.0;
    // one call site (one copy of the solver): what is written / compared is chosen by pointers
    //   pass 0          : trajectory -> scratch plane 0, SSE against the data
    //   pass 1, d = 1   : compared on the fly with plane 0 (X'X accumulates inside the solve)
    //   pass j, d = 3   : trajectory -> scratch plane j (combined by rsf_init_finish_kernel)
    double *wr = (pass == 0) ? scratch + cc : (D == 1 ? nullptr : scratch + (size_t)pass * plane + cc);
    const double *rd = (pass > 0 && D == 1) ? scratch + cc : nullptr;
    SolveOut o = rsf_solve<VS, false, K1P>(M, a, b, dc, active, series, lscr, wr, rd, (size_t)C, fd_den, &xtx);
    if (active) {
        if (pass == 0) {
            S.sse[c] = o.sse;
            S.sigma2[c] = o.sse / (double)(M.n_out - n_prior_len);     // :261
        } else if (D == 1) {
            S.chol[c] = S.sigma2[c] * (1.0 / xtx);                     // Vstart, :265-266
        }
        S.nrhs[c] += o.nrhs; S.nstep[c] += o.nstep; S.status[c] |= o.status; S.nsolve[c] += 1; S.nexec[c] += 1; S.urhs[c] += o.nrhs; S.ustep[c] += o.nstep;
    }
}

// d = 3: X'X from the stored trajectories, Vstart = sigma2_0 (X'X)^-1, chol(Vstart)
struct Box3 { double lo[3], hi[3]; };
__global__ void rsf_init_finish_kernel(int C, int n_out, SamplerDev S, const double *__restrict__ scratch, Box3 box)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const size_t plane = (size_t)n_out * C;
    double den[3];
    for (int j = 0; j < 3; j++) den[j] = S.q[(size_t)j * C + c] * (1 + 1e-6) * 1e-6;
    double g00 = 0, g10 = 0, g11 = 0, g20 = 0, g21 = 0, g22 = 0;
    for (int k = 0; k < n_out; k++) {
        const double base = scratch[(size_t)k * C + c];
        const double x0 = (scratch[plane + (size_t)k * C + c] - base) / den[0];
        const double x1 = (scratch[2 * plane + (size_t)k * C + c] - base) / den[1];
        const double x2 = (scratch[3 * plane + (size_t)k * C + c] - base) / den[2];
        g00 += x0 * x0; g10 += x1 * x0; g11 += x1 * x1; g20 += x2 * x0; g21 += x2 * x1; g22 += x2 * x2;
    }
    // Cholesky G = R R^T, then G^-1 = R^-T R^-1; V = s2 G^-1; L = chol(V)
    const double s2 = S.sigma2[c];
    double r00 = sqrt(g00), r10 = g10 / r00, r20 = g20 / r00;
    double r11 = sqrt(g11 - r10 * r10), r21 = (g21 - r20 * r10) / r11;
    double r22 = sqrt(g22 - r20 * r20 - r21 * r21);
    // W = R^-1 (lower)
    double w00 = 1 / r00, w11 = 1 / r11, w22 = 1 / r22;
    double w10 = -r10 * w00 * w11, w21 = -r21 * w11 * w22;
    double w20 = -(r20 * w00 + r21 * w10) * w22;
    // V = s2 * W^T W
    double v00 = s2 * (w00 * w00 + w10 * w10 + w20 * w20);
    double v10 = s2 * (w10 * w11 + w20 * w21);
    double v11 = s2 * (w11 * w11 + w21 * w21);
    double v20 = s2 * (w20 * w22);
    double v21 = s2 * (w21 * w22);
    double v22 = s2 * (w22 * w22);
    // (a, b) are nearly collinear in X, so Vstart can put most proposals outside the prior box.  No
    // reference behaviour exists for d = 3; the covariance is shrunk by one common factor (shape and
    // correlations kept) until every marginal s.d. is at most 1/20 of its prior width.
    {
        double gam = 1.0;
        const double w0 = (box.hi[0] - box.lo[0]) / 20.0, w1 = (box.hi[1] - box.lo[1]) / 20.0,
                     w2 = (box.hi[2] - box.lo[2]) / 20.0;
        if (v00 > w0 * w0) gam = fmin(gam, w0 * w0 / v00);
        if (v11 > w1 * w1) gam = fmin(gam, w1 * w1 / v11);
        if (v22 > w2 * w2) gam = fmin(gam, w2 * w2 / v22);
        v00 *= gam; v10 *= gam; v11 *= gam; v20 *= gam; v21 *= gam; v22 *= gam;
    }
    double l00 = sqrt(v00), l10 = v10 / l00, l20 = v20 / l00;
    double l11 = sqrt(v11 - l10 * l10), l21 = (v21 - l20 * l10) / l11;
    double l22 = sqrt(v22 - l20 * l20 - l21 * l21);
    S.chol[0 * (size_t)C + c] = l00; S.chol[1 * (size_t)C + c] = l10; S.chol[2 * (size_t)C + c] = l11;
    S.chol[3 * (size_t)C + c] = l20; S.chol[4 * (size_t)C + c] = l21; S.chol[5 * (size_t)C + c] = l22;
}

__global__ void fill_ring_kernel(int C, int W, SamplerDev S)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    // slot (iteration % W) holds the sample of that iteration; iteration 0 is the start value
    S.ring[c] = S.q[c];
}

extern "C" int rsfm_init(rsfm_sampler *s, const double *q0_dev, const double *data_dev, void *stream_)
{
    if (!s || !q0_dev || !data_dev) return set_err(RSFM_ERR_INVALID, "rsfm_init: NULL argument%s", "");
    cudaStream_t stream = (cudaStream_t)stream_;
    { int rc_ = join_groups(s, stream); if (rc_) return rc_; s->need_fork = 1; }
    const int d = s->cfg.n_params, C = s->C, n = s->cfg.n_out;
    CUDA_TRY(cudaMemcpyAsync(s->d.q, q0_dev, sizeof(double) * d * (size_t)C, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(s->d.data, data_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    const size_t planes = (d == 1) ? 1 : (size_t)(1 + d);
    const size_t scratch_bytes = sizeof(double) * planes * n * (size_t)C;
    if (!s->scratch) {
        int rc = scratch_acquire(scratch_bytes, (void **)&s->scratch);
        if (rc) return rc;
        s->scratch_bytes = scratch_bytes;
    }
    CUDA_TRY(cudaMemsetAsync(s->d.accepted, 0, sizeof(unsigned int) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.status, 0, sizeof(int) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.nrhs, 0, sizeof(unsigned long long) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.nstep, 0, sizeof(unsigned long long) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.nsolve, 0, sizeof(unsigned long long) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.nearly, 0, sizeof(unsigned long long) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.nexec, 0, sizeof(unsigned long long) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.urhs, 0, sizeof(unsigned long long) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.ustep, 0, sizeof(unsigned long long) * C, stream));
    CUDA_TRY(cudaMemsetAsync(s->d.suff, 0, sizeof(double) * (d + tri(d)) * (size_t)C, stream));
    if (s->d.fit) CUDA_TRY(cudaMemsetAsync(s->d.fit, 0, sizeof(double) * s->fit_rows * (size_t)C, stream));
    const int block = pick_block(C, &s->cfg), grid = (C + block - 1) / block;
    const ModelK M = make_model(&s->cfg);
    // nominal load table of this sampler (a tabulated load is read HERE: the table must not change afterwards); the
    // slip law has no fast interval and runs without it
    s->d.nom = s->cfg.state_law == RSFM_LAW_AGING ? s->nom_buf : nullptr;
    if (s->d.nom) {
        loading_table_kernel<<<1, 1024, 0, stream>>>(M, s->d.nom);
        CUDA_TRY(cudaGetLastError());
    }
    for (int pass = 0; pass <= d; pass++) {
#define RSFM_INIT(D, VS)                                                                                              \
    rsf_init_kernel<D, VS><<<VS ? (C + STIFF_BLOCK - 1) / STIFF_BLOCK : grid, VS ? STIFF_BLOCK : block, 0, stream>>>( \
        M, C, pass, s->cfg.a, s->cfg.b, s->cfg.n_prior_len, s->d, s->scratch)
        if (s->cfg.sampled_param == RSFM_PARAM_K1)
            rsf_init_kernel<1, false, true><<<grid, block, 0, stream>>>(M, C, pass, s->cfg.a, s->cfg.b, s->cfg.n_prior_len,
                                                                        s->d, s->scratch);
        else if (d == 1) { if (stiff_variant(&s->cfg)) RSFM_INIT(1, true); else RSFM_INIT(1, false); }
        else { if (stiff_variant(&s->cfg)) RSFM_INIT(3, true); else RSFM_INIT(3, false); }
#undef RSFM_INIT
        CUDA_TRY(cudaGetLastError());
    }
    if (d == 3) {
        Box3 box;
        for (int i = 0; i < 3; i++) { box.lo[i] = s->cfg.lo[i]; box.hi[i] = s->cfg.hi[i]; }
        rsf_init_finish_kernel<<<(C + 127) / 128, 128, 0, stream>>>(C, n, s->d, s->scratch, box);
        CUDA_TRY(cudaGetLastError());
    } else {
        fill_ring_kernel<<<(C + 127) / 128, 128, 0, stream>>>(C, s->cfg.adapt_interval, s->d);
        CUDA_TRY(cudaGetLastError());
    }
    // the base trajectory is only needed during init; hand it back (small ones are cached, GB-sized
    // ones freed)
    CUDA_TRY(cudaStreamSynchronize(stream));
    scratch_release(s->scratch, scratch_bytes);
    s->scratch = nullptr;
    s->iteration = 0;
    s->suff_count = 0;
    s->initialised = 1;
    return RSFM_OK;
}

// ---------------------------------------------------------------------------
// fused MCMC iterations
// ---------------------------------------------------------------------------
struct RunArgs {
    int n_iters;
    long long iter0;
    unsigned long long seed, chain_id0;
    // outputs
    double *samples;        // [n_iters][d][C]
    double *sigma2_out;     // [n_iters][C]
    unsigned char *accept;  // [n_iters][C]
    double *draws;          // [n_iters][d+2][C]
    // deterministic inputs
    const double *proposals;  // [n_iters][d][C]
    const double *uniforms;   // [n_iters][C]
    const double *gammas;     // [n_iters][C]
    int proposals_are_z;
    int deterministic;
    // config
    double a0, b0, n0;
    double lo[RSFM_MAX_PARAMS], hi[RSFM_MAX_PARAMS];
    int adapt_mode, adapt_interval;
    int c_begin, c_end;     // chains [c_begin, c_end) of the sampler are served by this launch (chain groups)
};

// (128, 3): three resident blocks per SM (<= 168 registers); measured +5 % at saturating sizes
// PACK (d = 3, Philox-driven, 128-thread blocks): at the start of every solve round the in-bounds proposals of the
// BLOCK are packed into its lowest threads, so that a warp's round is spent on up to 32 solves whichever chains they
// belong to; warps left without a solve skip the round.  Without it a warp runs as many rounds as its lane with the
// most in-bounds proposals has (8.2 of the 10 iterations of a launch, against 5 on average at cfg 3).  The solver is
// untouched: a thread solves for the proposal it was handed and the result goes back to the owner through shared
// memory, so chains are bit-identical with and without packing.
// K1P (d = 1): the chain's scalar is k1 (RSFM_PARAM_K1), Dc = M.dc_fixed; everything but the solve is unchanged.
template <int D, bool DET, bool VS, bool PACK = false, bool K1P = false>
__global__ void __launch_bounds__(128, VS ? 1 : 3)
rsf_mcmc_kernel(const __grid_constant__ ModelK M, int C, SamplerDev S, RunArgs A)
{
    static_assert(!PACK || (D == 3 && !DET && !VS), "packing: the Philox-driven d = 3 kernel of the default variant");
    static_assert(!K1P || (D == 1 && !VS), "k1 chains: d = 1, default variant");
    __shared__ double s_job[PACK ? 4 * 128 : 1];        // (a, b, Dc, rejection threshold) of the packed proposals
    __shared__ double s_res[PACK ? 128 : 1];            // sum of squares of the finished solve, by owner
    __shared__ int s_own[PACK ? 128 : 1], s_resi[PACK ? 3 * 128 : 1], s_wc[PACK ? 4 : 1];
    __shared__ __align__(128) double s_tile[2 * SERIES_TILE];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_ltab[(VS ? 1 : 4) * LTAB_STRIDE];       // VS kernels run one-warp blocks (STIFF_BLOCK)
    __shared__ double s_lpriv[11 * (VS ? 32 : 128)];
    LoadScratch lscr;
    lscr.tab = s_ltab; lscr.priv = s_lpriv; lscr.nom = S.nom;
    constexpr int T = D * (D + 1) / 2;
    const int c = A.c_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = c < A.c_end;
    const int cc = active ? c : A.c_end - 1;
    const size_t Cz = (size_t)C;

    double q[D], L[T], sq[D], sqq[T];
#pragma unroll
    for (int j = 0; j < D; j++) { q[j] = S.q[j * Cz + cc]; sq[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < T; j++) { L[j] = S.chol[j * Cz + cc]; sqq[j] = 0.0; }
    double ss = S.sse[cc], s2 = S.sigma2[cc];
    unsigned int n_acc = 0;
    unsigned long long nrhs = 0, nstep = 0;
    unsigned int nsolve = 0, nearly = 0;
    int status = 0;
    const unsigned long long gid = A.chain_id0 + (unsigned long long)cc;
    const PhiloxKey key = philox_key(A.seed);
    const double gshape = 0.5 * (A.n0 + (double)M.n_out);          // MCMC.py:158
    SeriesStage series;
    series.begin(s_tile, s_bar, S.data, M.n_out);

    // d = 3, Philox-driven: a proposal outside the prior box is rejected without a solve (MCMC.py:318-320),
    // and with a wide posterior that is every other proposal (cfg 3).  Instead of idling through the solve of
    // its warp-mates, a lane walks on through its own out-of-bounds iterations (sigma^2 draw, outputs, pooled
    // sums -- everything but a solve) until it holds an in-bounds proposal; the warp then solves together.
    // Lanes get out of step in iteration count, not in results: every draw is keyed by (chain, iteration).
    // Needs a single-tile resident series (no block barriers after the first solve).
    constexpr bool SKIP_T = (D == 3) && !DET;
    if constexpr (SKIP_T) {
        const bool skip_ok = M.n_out <= SERIES_TILE;
        auto propose = [&](int it_, double (&qn_)[D]) {
            const unsigned int giter = (unsigned int)(A.iter0 + it_);
            double z0, z1, z2, zz;
            philox_normal2(key, gid, giter, 0u, z0, z1);
            philox_normal2(key, gid, giter, 1u, z2, zz);
            qn_[0] = q[0] + L[0] * z0;
            qn_[1] = q[1] + L[1] * z0 + L[2] * z1;
            qn_[2] = q[2] + L[3] * z0 + L[4] * z1 + L[5] * z2;
            bool inb_ = true;
#pragma unroll
            for (int j = 0; j < D; j++) inb_ = inb_ && (qn_[j] > A.lo[j]) && (qn_[j] < A.hi[j]);
            return inb_;
        };
        // everything of an iteration after the accept / reject decision (sigma^2 draw, outputs, pooled sums)
        auto finish = [&](int it_, bool acc_, const double (&qn_)[D], double u_) {
            const unsigned int giter = (unsigned int)(A.iter0 + it_);
            const double g0 = philox_gamma(key, gid, giter, gshape);
            {
                const double bval = 0.5 * (A.n0 * s2 + ss);
                const double scale = 1.0 / bval;
                s2 = 1.0 / (g0 * scale);
            }
            if (A.samples) {
#pragma unroll
                for (int j = 0; j < D; j++) A.samples[((size_t)it_ * D + j) * Cz + c] = q[j];
            }
            if (A.sigma2_out) A.sigma2_out[(size_t)it_ * Cz + c] = s2;
            if (A.accept) A.accept[(size_t)it_ * Cz + c] = acc_ ? 1 : 0;
            if (A.draws) {
#pragma unroll
                for (int j = 0; j < D; j++) A.draws[((size_t)it_ * (D + 2) + j) * Cz + c] = qn_[j];
                A.draws[((size_t)it_ * (D + 2) + D) * Cz + c] = u_;
                A.draws[((size_t)it_ * (D + 2) + D + 1) * Cz + c] = g0;
            }
            if (A.adapt_mode == RSFM_ADAPT_POOLED) {
#pragma unroll
                for (int j = 0; j < D; j++) sq[j] += q[j];
                int t = 0;
#pragma unroll
                for (int i = 0; i < D; i++)
#pragma unroll
                    for (int j = 0; j <= i; j++) sqq[t++] += q[i] * q[j];
            }
        };
        int it = 0;
#pragma unroll 1
        for (int trip = 0;; trip++) {
            bool live = active && it < A.n_iters;
            if (skip_ok) {
                // (PACK: the round contains block barriers, so the block leaves the loop together)
                if (PACK) { if (trip > 0 && !__syncthreads_or(live)) break; }
                else if (trip > 0 && !__any_sync(FULL_MASK, live)) break;
            }
            else if (trip >= A.n_iters) break;
            double qn[D] = {q[0], q[1], q[2]};
            bool inb = false;
            while (live) {
                inb = propose(it, qn);
                if (inb || !skip_ok) break;
                finish(it, false, qn, nan(""));
                it++;
                live = it < A.n_iters;
            }
            const bool solve = live && inb;
            double u = nan("");
            double lnu = 0.0;
            if (solve) {
                u = philox_uniform(key, gid, (unsigned int)(A.iter0 + it), 2u);
                lnu = log(u);
            }
            const double sse_limit = solve ? ss - 2.0 * s2 * lnu : INFINITY;
            series.start_solve();
            SolveOut o;
            if constexpr (PACK) {
                // rank of this proposal among the in-bounds proposals of the block
                const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
                const unsigned m = __ballot_sync(FULL_MASK, solve);
                if (lane == 0) s_wc[wid] = __popc(m);
                __syncthreads();
                int base = 0, total = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { const int n = s_wc[i]; base += (i < wid) ? n : 0; total += n; }
                if (solve) {
                    const int r = base + __popc(m & ((1u << lane) - 1u));
                    s_job[r] = qn[0]; s_job[128 + r] = qn[1]; s_job[256 + r] = qn[2]; s_job[384 + r] = sse_limit;
                    s_own[r] = threadIdx.x;
                }
                __syncthreads();
                // thread t solves proposal t; a warp without a proposal skips the round
                const bool work = (int)threadIdx.x < total;
                const int t = threadIdx.x;
                SolveOut w;
                w.sse = 0.0; w.status = 0; w.filled = 0; w.nrhs = 0; w.nstep = 0;
                if ((int)(threadIdx.x & ~31u) < total)
                    w = rsf_solve<VS>(M, work ? s_job[t] : q[0], work ? s_job[128 + t] : q[1], work ? s_job[256 + t] : q[2], work,
                                      series, lscr, nullptr, nullptr, Cz, 1.0, nullptr, nullptr, work ? s_job[384 + t] : INFINITY);
                if (work) {
                    const int ow = s_own[t];
                    s_res[ow] = w.sse; s_resi[ow] = w.status; s_resi[128 + ow] = (int)w.nrhs; s_resi[256 + ow] = (int)w.nstep;
                }
                __syncthreads();
                o.sse = s_res[t]; o.status = s_resi[t]; o.filled = 0;
                o.nrhs = (uint32_t)s_resi[128 + t]; o.nstep = (uint32_t)s_resi[256 + t];
            } else {
                o = rsf_solve<VS>(M, solve ? qn[0] : q[0], solve ? qn[1] : q[1], solve ? qn[2] : q[2], solve, series, lscr,
                                  nullptr, nullptr, Cz, 1.0, nullptr, nullptr, sse_limit);
            }
            bool acc = false;
            if (solve) {
                nrhs += o.nrhs; nstep += o.nstep; status |= (o.status & ~RSFM_CHAIN_EARLY); nsolve++;
                if (o.status & RSFM_CHAIN_EARLY) nearly++;
                double la = 0.5 * (ss - o.sse) / s2;
                if (la > 0.0) la = 0.0;
                acc = la > lnu;
                if (acc) {
#pragma unroll
                    for (int j = 0; j < D; j++) q[j] = qn[j];
                    ss = o.sse;
                    n_acc++;
                }
            }
            if (live) {
                finish(it, acc, qn, u);
                it++;
            }
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < A.n_iters; it++) {
            const unsigned int giter = (unsigned int)(A.iter0 + it);
            // ---- proposal (MCMC.py:497) ----
            double qn[D];
            if (DET && !A.proposals_are_z) {
#pragma unroll
                for (int j = 0; j < D; j++) qn[j] = A.proposals[((size_t)it * D + j) * Cz + cc];
            } else {
                double z[D];
                if (DET) {
#pragma unroll
                    for (int j = 0; j < D; j++) z[j] = A.proposals[((size_t)it * D + j) * Cz + cc];
                } else {
                    double z0, z1;
                    philox_normal2(key, gid, giter, 0u, z0, z1);
                    z[0] = z0;
                    if (D > 1) z[1] = z1;
                    if (D > 2) { philox_normal2(key, gid, giter, 1u, z0, z1); z[2] = z0; }
                }
                if (D == 1) {
                    qn[0] = q[0] + sqrt(L[0]) * z[0];       // L[0] is the proposal variance for d = 1
                } else {
                    qn[0] = q[0] + L[0] * z[0];
                    qn[1] = q[1] + L[1] * z[0] + L[2] * z[1];
                    qn[2] = q[2] + L[3] * z[0] + L[4] * z[1] + L[5] * z[2];
                }
            }
            // ---- strict box prior (MCMC.py:318-320) ----
            bool inb = true;
#pragma unroll
            for (int j = 0; j < D; j++) inb = inb && (qn[j] > A.lo[j]) && (qn[j] < A.hi[j]);
            const bool solve = active && inb;
            // ---- acceptance uniform first (MCMC.py:331): it fixes the rejection threshold ----
            // accept  <=>  min(0, 0.5 (SS - SS')/s2) > ln U  <=>  SS' < SS - 2 s2 ln U   (ln U < 0)
            double u = nan("");
            double lnu = 0.0;
            if (solve) {
                if (DET) u = A.uniforms[(size_t)it * Cz + cc];
                else u = philox_uniform(key, gid, giter, 2u);
                lnu = log(u);
            }
            const double sse_limit = solve ? ss - 2.0 * s2 * lnu : INFINITY;
            // ---- forward solve + SSE (MCMC.py:324, 381-387), stopped early once rejection is certain ----
            const double pa = (D == 3) ? qn[0] : A.a0;
            const double pb = (D == 3) ? qn[1] : A.b0;
            const double pdc = solve ? qn[D - 1] : q[D - 1];
            series.start_solve();
            SolveOut o = rsf_solve<VS, false, K1P>(M, solve ? pa : ((D == 3) ? q[0] : A.a0), solve ? pb : ((D == 3) ? q[1] : A.b0),
                                                   pdc, solve, series, lscr, nullptr, nullptr, Cz, 1.0, nullptr, nullptr, sse_limit);
            // ---- accept / reject (MCMC.py:327-331) ----
            bool acc = false;
            if (solve) {
                nrhs += o.nrhs; nstep += o.nstep; status |= (o.status & ~RSFM_CHAIN_EARLY); nsolve++;
                if (o.status & RSFM_CHAIN_EARLY) nearly++;
                double la = 0.5 * (ss - o.sse) / s2;
                if (la > 0.0) la = 0.0;
                acc = la > lnu;
                if (acc) {
#pragma unroll
                    for (int j = 0; j < D; j++) q[j] = qn[j];
                    ss = o.sse;
                    n_acc++;
                }
            }
            // ---- sigma^2 Gibbs draw (MCMC.py:158-160) ----
            double g0;
            if (DET) g0 = A.gammas[(size_t)it * Cz + cc];
            else g0 = philox_gamma(key, gid, giter, gshape);
            {
                const double bval = 0.5 * (A.n0 * s2 + ss);
                const double scale = 1.0 / bval;
                s2 = 1.0 / (g0 * scale);
            }
            // ---- outputs ----
            if (active) {
                if (A.samples) {
#pragma unroll
                    for (int j = 0; j < D; j++) A.samples[((size_t)it * D + j) * Cz + c] = q[j];
                }
                if (A.sigma2_out) A.sigma2_out[(size_t)it * Cz + c] = s2;
                if (A.accept) A.accept[(size_t)it * Cz + c] = acc ? 1 : 0;
                if (A.draws) {
#pragma unroll
                    for (int j = 0; j < D; j++) A.draws[((size_t)it * (D + 2) + j) * Cz + c] = qn[j];
                    A.draws[((size_t)it * (D + 2) + D) * Cz + c] = u;
                    A.draws[((size_t)it * (D + 2) + D + 1) * Cz + c] = g0;
                }
            }
            // ---- adaptation ----
            if (A.adapt_mode == RSFM_ADAPT_POOLED) {
#pragma unroll
                for (int j = 0; j < D; j++) sq[j] += q[j];
                int t = 0;
#pragma unroll
                for (int i = 0; i < D; i++)
#pragma unroll
                    for (int j = 0; j <= i; j++) sqq[t++] += q[i] * q[j];
            }
            if (D == 1 && A.adapt_mode == RSFM_ADAPT_COMPAT && active) {
                // MCMC.py:523-527 + 200-204: V <- chol(2.38^2/len(keys) * cov(last W samples)), W = adapt_interval,
                // len(qpriors.keys()) = 2 for the dict form.  The Cholesky FACTOR is then used as a covariance (q3).
                const int W = A.adapt_interval;
                const long long gi = A.iter0 + it + 1;              // chain index of the sample just appended
                S.ring[(size_t)(gi % W) * Cz + c] = q[0];
                if (gi % W == 0) {
                    double mean = 0.0;
                    for (int w = 1; w <= W; w++) mean += S.ring[(size_t)((gi - W + w) % W) * Cz + c];
                    mean /= W;
                    double v = 0.0;
                    for (int w = 1; w <= W; w++) {
                        const double dlt = S.ring[(size_t)((gi - W + w) % W) * Cz + c] - mean;
                        v += dlt * dlt;
                    }
                    v /= (W - 1);
                    const double vnew = 2.38 * 2.38 / 2.0 * v;
                    if (vnew > 0.0) L[0] = sqrt(vnew);             // cov = 0 -> cholesky raises -> unchanged (q4)
                }
            }
        }
    }

    if (active) {
#pragma unroll
        for (int j = 0; j < D; j++) S.q[j * Cz + c] = q[j];
#pragma unroll
        for (int j = 0; j < T; j++) S.chol[j * Cz + c] = L[j];
        S.sse[c] = ss; S.sigma2[c] = s2;
        S.accepted[c] += n_acc;
        S.nrhs[c] += nrhs; S.nstep[c] += nstep; S.status[c] |= status; S.nsolve[c] += nsolve; S.nearly[c] += nearly;
        S.nexec[c] += nsolve; S.urhs[c] += nrhs; S.ustep[c] += nstep;
        if (A.adapt_mode == RSFM_ADAPT_POOLED) {
#pragma unroll
            for (int j = 0; j < D; j++) S.suff[j * Cz + c] += sq[j];
#pragma unroll
            for (int j = 0; j < T; j++) S.suff[(D + j) * Cz + c] += sqq[j];
        }
    }
}


// ---------------------------------------------------------------------------
// speculative ("prefetching") Metropolis for small chain counts, with a predictor
// ---------------------------------------------------------------------------
// With C chains on a 148-SM part, one thread per chain leaves most SMSPs idle while every chain
// waits a full solve latency per iteration (the loop of MCMC.py:494 is strictly sequential).  Here
// G = 2^g lanes serve one chain: every lane evaluates one NODE of the binary tree of the chain's
// possible futures.  A node at depth l is reached by a path of l-1 accept / reject outcomes of the
// round's earlier iterations and evaluates the proposal iteration it+l-1 would make after them
// (its accept child proposes from that proposal, its reject child from the state before).  All
// solves of the tree run concurrently; afterwards every lane of the group walks the realised path
// with the tree's sums of squares, applying exactly the reference's accept rule and sigma^2 update in
// order.  Random draws are Philox values keyed by (chain, iteration), so the chain is the one the
// sequential kernel produces, bit for bit; only the wall time per iteration changes.
//
// WHICH G nodes are evaluated is chosen anew every round, best first by the probability that the
// realised path reaches them (round 2; "predictive prefetching").  The root has probability 1, the
// children of a node with estimated acceptance probability p have reach p and 1 - p times the
// node's; the open child with the largest reach is added G - 1 times (one argmax over the group per
// node; the lane that receives the node computes its proposal from the parent's state handed over by
// shuffles).  The estimate p never enters a result -- a wrong guess only ends the round earlier:
//   * d = 1 (Dc): the sum of squares is, to ~1e-2 sigma^2 over the whole posterior, a quadratic in
//     1/Dc (the response scales with k' = 0.1/Dc, RateStateModel.py:324), so a three-coefficient
//     least-squares fit through the completed solves of the last few rounds -- every node of every
//     tree is an exact sample of SS(q), used or not -- predicts SS' of a proposal, and with the
//     acceptance uniform and the gamma draws of the coming iterations known in advance (Philox) the
//     decision SS' < SS - 2 s2 ln U itself: p = Phi((threshold - SS'_fit) / rms residual of the fit).
//     On the bench workload 99 % of the decisions are predicted and a 16-lane group advances ~15
//     iterations per round where the balanced tree of round 1 advanced 4.
//   * d = 3 (a, b, Dc): the same with a full quadratic (ten coefficients) in (1/a, b/a, 1/Dc), whose
//     normal equations one lane of the group keeps in shared memory and solves by Cholesky.
//   * no fit yet: p = the running acceptance rate of the launch; p = 0.5 gives back the balanced
//     tree.
//
// Early stopping: a node whose path holds no accept is judged against the state the round started
// with, so its threshold SS - 2 s2 ln U is known exactly (s2 follows from the gamma draws alone)
// and is its stopping bound, like the root's.  Any other node stops at a generous bound
// H = SS + 100 s2; at resolution time a stopped node counts as rejected only if its bound is at
// least the true threshold (then rejection is certain), otherwise it is unresolved and the round
// ends before it -- that iteration becomes the root of the next round.  Exactness never depends on
// the bounds.
//
// COMPAT = true (d = 1, the reference's dict-prior adaptation): the proposal scale changes after every
// adapt_interval-th sample, so a round never looks past that boundary; the writer lane keeps the
// reference's sample ring and hands the new scale to its group when the boundary is reached.
template <int D, bool COMPAT, bool VS>
__global__ void __launch_bounds__(128)
rsf_mcmc_spec_kernel(const __grid_constant__ ModelK M, int C, SamplerDev S, RunArgs A, int g)
{
    __shared__ __align__(128) double s_tile[2 * SERIES_TILE];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_ltab[4 * LTAB_STRIDE];
    __shared__ double s_lpriv[11 * 128];
    // the fit of SS(q), one slot per group (one-warp blocks: <= 16 groups per block): moments S0..S4 of x and T0..T2 of
    // x^k y, coefficients c0..c2, mean squared residual, "fit valid", "residual known", and the centre / scale of the
    // fit variables (kept here, not in registers: nothing of the predictor is live across the solve)
    __shared__ double s_sur[(D == 3) ? 8 * SPEC_FIT3 : 32 * SPEC_FIT];    // (d = 3 runs >= 4 lanes per chain: <= 8 groups)
    LoadScratch lscr;
    lscr.tab = s_ltab; lscr.priv = s_lpriv; lscr.nom = S.nom;
    constexpr int T = D * (D + 1) / 2;
    const int G = 1 << g;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int chain = tid >> g;
    const bool chain_ok = chain < C;
    const int cc = chain_ok ? chain : C - 1;
    const int lane = threadIdx.x & 31;
    const int gbase = lane & ~(G - 1);
    const int li = lane - gbase;                       // this lane's slot in its group: it receives the li-th node
    const size_t Cz = (size_t)C;
    const bool writer = chain_ok && li == 0;
    double *const sur = s_sur + (threadIdx.x >> g) * ((D == 3) ? SPEC_FIT3 : SPEC_FIT);

    double q[D], L[T], sq[D], sqq[T];
#pragma unroll
    for (int j = 0; j < D; j++) { q[j] = S.q[j * Cz + cc]; sq[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < T; j++) { L[j] = S.chol[j * Cz + cc]; sqq[j] = 0.0; }
    double ss = S.sse[cc], s2 = S.sigma2[cc];
    unsigned int n_acc = 0, nsolve = 0, nearly = 0, nexec = 0;
    unsigned long long nrhs = 0, nstep = 0, urhs = 0, ustep = 0;
    int status = 0;
    const unsigned long long gid = A.chain_id0 + (unsigned long long)cc;
    const PhiloxKey key = philox_key(A.seed);
    const double gshape = 0.5 * (A.n0 + (double)M.n_out);
    SeriesStage series;
    series.begin(s_tile, s_bar, S.data, M.n_out);

    // fit variable x = (1/q' - 1/q_c) q_c^2 / (proposal s.d.), y = SS' - SS_c, centred on the state at launch
    // (the fit lives in the sampler between launches, SamplerDev::fit, and in shared memory during one)
    if (D == 3 && li == 0) {
        // d = 3: t = ((1/a' - 1/a_c) a_c^2/sd_a, (b'/a' - b_c/a_c) a_c/sd_b, (1/Dc' - 1/Dc_c) Dc_c^2/sd_Dc) with the
        // marginal proposal s.d. of the factor in force at the first launch
        const bool kept = S.fit && S.fit[85 * Cz + cc] != 0.0;
        if (kept) {
#pragma unroll 1
            for (int k = 0; k < SPEC_FIT3; k++) sur[k] = S.fit[k * Cz + cc];
        } else {
#pragma unroll 1
            for (int k = 0; k < SPEC_FIT3; k++) sur[k] = 0.0;
            const double sda = fabs(L[0]), sdb = sqrt(L[1 % T] * L[1 % T] + L[2 % T] * L[2 % T]);
            const double sdd = sqrt(L[3 % T] * L[3 % T] + L[4 % T] * L[4 % T] + L[5 % T] * L[5 % T]);
            const double a_c = q[0], b_c = q[1 % D], d_c = q[2 % D];
            sur[78] = 1.0 / a_c; sur[79] = b_c / a_c; sur[80] = 1.0 / d_c;
            sur[81] = a_c * a_c / sda; sur[82] = a_c / sdb; sur[83] = d_c * d_c / sdd; sur[84] = ss;
            const bool ok = a_c > 0.0 && d_c > 0.0 && sur[81] > 0.0 && sur[81] < 1e300 && sur[82] > 0.0 && sur[82] < 1e300 &&
                            sur[83] > 0.0 && sur[83] < 1e300;
            sur[85] = ok ? 1.0 : -1.0;
        }
    }
    if (D != 3 && li == 0) {
        const bool kept = (D == 1) && S.fit && S.fit[17 * Cz + cc] != 0.0;
        if (kept) {
#pragma unroll
            for (int k = 0; k < SPEC_FIT; k++) sur[k] = S.fit[k * Cz + cc];
        } else {
#pragma unroll
            for (int k = 0; k < 14; k++) sur[k] = 0.0;
            const double xs = (D == 1) ? q[0] * q[0] / sqrt(L[0]) : 0.0;
            sur[14] = 1.0 / q[0]; sur[15] = ss; sur[16] = xs;
            sur[17] = ((D == 1) && q[0] > 0.0 && xs > 0.0 && xs < 1e300) ? 1.0 : -1.0;
        }
    }
    __syncwarp();

    int it = 0;
    while (__any_sync(FULL_MASK, chain_ok && it < A.n_iters)) {
        const bool live = chain_ok && it < A.n_iters;
        int rmax = live ? min(G, A.n_iters - it) : 0;
        if (COMPAT) rmax = min(rmax, A.adapt_interval - (int)((A.iter0 + it) % A.adapt_interval));
        // ---- draws of iteration it + li (slot li of the group): the tree below reads them by shuffle ----
        double dz[D], dlnu = 0.0, dgam = 1.0;
#pragma unroll
        for (int j = 0; j < D; j++) dz[j] = 0.0;
        if (li < rmax) {
            const unsigned int giter = (unsigned int)(A.iter0 + it + li);
            double z0, z1;
            philox_normal2(key, gid, giter, 0u, z0, z1);
            dz[0] = z0;
            if (D > 1) dz[1] = z1;
            if (D > 2) { philox_normal2(key, gid, giter, 1u, z0, z1); dz[2] = z0; }
            dlnu = log(philox_uniform(key, gid, giter, 2u));
            dgam = philox_gamma(key, gid, giter, gshape);
        }
        // ---- predictor of this round ----
        const bool have_fit = (D == 3) ? (sur[85] > 0.0 && sur[76] != 0.0 && sur[77] != 0.0)
                                       : (sur[17] > 0.0 && sur[12] != 0.0 && sur[13] != 0.0);
        const double fc0 = sur[8], fc1 = sur[9], fc2 = sur[10];
        const double rqc = sur[14], ss0 = sur[15], xs = sur[16];
        const double rtau = have_fit ? rsqrt(2.0 * fmax((D == 3) ? sur[75] : sur[11], 1e-300)) : 0.0;
        const double p0 = fmin(fmax(((double)n_acc + 1.0) / ((double)nsolve + 2.0), 0.05), 0.95);
        // ---- the tree: slot r of the group receives the r-th node, best first by reach probability ----
        bool have = false, n_inb = false, n_exact = false;
        int n_depth = 0, cA = -1, cR = -1;
        double n_cur[D], qn[D], n_ssc = ss, n_s2c = s2, n_reach = 0.0, n_p = 0.0, n_sshat = ss, n_gam = 1.0;
        double vA = -1.0, vR = -1.0, limit = INFINITY;
#pragma unroll
        for (int j = 0; j < D; j++) { n_cur[j] = q[j]; qn[j] = q[j]; }
#pragma unroll 1
        for (int r = 0; r < G; r++) {
            // the open child with the largest reach (ties: lowest lane)
            double bv = have ? fmax(vA, vR) : -1.0;
            int bl = lane;
            for (int off = 1; off < G; off <<= 1) {
                const double ov = __shfl_xor_sync(FULL_MASK, bv, off);
                const int ol = __shfl_xor_sync(FULL_MASK, bl, off);
                if (ov > bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
            }
            const bool root_step = r == 0;
            const bool take = root_step ? rmax >= 1 : bv > 0.0;
            // what every lane would hand to its better child; the winner's offer is read by shuffle
            const bool pickA = vA >= vR;
            const double o_ssc = pickA ? n_sshat : n_ssc;
            double o_s2;
            {
                const double bval = 0.5 * (A.n0 * n_s2c + o_ssc);
                const double scale = 1.0 / bval;
                o_s2 = 1.0 / (n_gam * scale);
            }
            double t_cur[D];
#pragma unroll
            for (int j = 0; j < D; j++) t_cur[j] = __shfl_sync(FULL_MASK, pickA ? qn[j] : n_cur[j], bl);
            double t_ssc = __shfl_sync(FULL_MASK, o_ssc, bl);
            double t_s2 = __shfl_sync(FULL_MASK, o_s2, bl);
            int t_depth = __shfl_sync(FULL_MASK, n_depth + 1, bl);
            bool t_exact = __shfl_sync(FULL_MASK, (!pickA && n_exact) ? 1 : 0, bl) != 0;
            if (root_step) {
#pragma unroll
                for (int j = 0; j < D; j++) t_cur[j] = q[j];
                t_ssc = ss; t_s2 = s2; t_depth = 1; t_exact = true; bv = 1.0;
            }
            if (!root_step && take && lane == bl) {
                if (pickA) { vA = -1.0; cA = gbase + r; } else { vR = -1.0; cR = gbase + r; }
            }
            // draws of the new node's iteration
            const int dsrc = gbase + min(max(t_depth, 1), G) - 1;
            double zz[D];
#pragma unroll
            for (int j = 0; j < D; j++) zz[j] = __shfl_sync(FULL_MASK, dz[j], dsrc);
            const double lnu_m = __shfl_sync(FULL_MASK, dlnu, dsrc);
            const double gam_m = __shfl_sync(FULL_MASK, dgam, dsrc);
            if (take && li == r) {
                have = true;
#pragma unroll
                for (int j = 0; j < D; j++) n_cur[j] = t_cur[j];
                n_ssc = t_ssc; n_s2c = t_s2; n_depth = t_depth; n_reach = bv; n_exact = t_exact; n_gam = gam_m;
                if (D == 1) {
                    qn[0] = n_cur[0] + sqrt(L[0]) * zz[0];       // L[0] is the proposal variance for d = 1
                } else {
                    qn[0] = n_cur[0] + L[0] * zz[0];
                    qn[1] = n_cur[1] + L[1] * zz[0] + L[2] * zz[1];
                    qn[2] = n_cur[2] + L[3] * zz[0] + L[4] * zz[1] + L[5] * zz[2];
                }
                n_inb = true;
#pragma unroll
                for (int j = 0; j < D; j++) n_inb = n_inb && (qn[j] > A.lo[j]) && (qn[j] < A.hi[j]);
                const double thr = n_ssc - 2.0 * n_s2c * lnu_m;
                n_sshat = n_ssc;
                if (!n_inb) {
                    n_p = 0.0;                                   // an out-of-bounds proposal is never accepted
                } else if (D == 3 && have_fit && qn[0] > 0.0 && qn[D - 1] > 0.0) {
                    const double ra = 1.0 / qn[0];
                    const double t0 = (ra - sur[78]) * sur[81], t1 = (qn[1 % D] * ra - sur[79]) * sur[82];
                    const double t2 = (1.0 / qn[D - 1] - sur[80]) * sur[83];
                    n_sshat = sur[84] + sur[65] + t0 * (sur[66] + t0 * sur[69] + t1 * sur[70] + t2 * sur[71]) +
                              t1 * (sur[67] + t1 * sur[72] + t2 * sur[73]) + t2 * (sur[68] + t2 * sur[74]);
                    n_p = fmin(fmax(0.5 * erfc(-(thr - n_sshat) * rtau), 0.02), 0.98);
                } else if (D == 1 && have_fit && qn[0] > 0.0) {
                    const double x = (1.0 / qn[0] - rqc) * xs;
                    n_sshat = ss0 + fma(x, fma(x, fc2, fc1), fc0);
                    n_p = fmin(fmax(0.5 * erfc(-(thr - n_sshat) * rtau), 0.02), 0.98);
                } else {
                    n_p = p0;
                }
                const bool kids = n_depth < rmax;
                vA = kids ? n_reach * n_p : -1.0;
                vR = kids ? n_reach * (1.0 - n_p) : -1.0;
                limit = n_inb ? (n_exact ? thr : ss + 100.0 * s2) : INFINITY;
            }
        }
        const bool inb = have && n_inb;
        const bool solve = inb;
        const double pa = (D == 3) ? qn[0] : A.a0;
        const double pb = (D == 3) ? qn[1] : A.b0;
        series.start_solve();
        // ISO: the general interval path out of line (255 registers available here; the 168-register kernels
        // would spill in the fast interval around the call)
        SolveOut o = rsf_solve<VS, true>(M, solve ? pa : A.a0, solve ? pb : A.b0, solve ? qn[D - 1] : q[D - 1], solve, series, lscr,
                               nullptr, nullptr, Cz, 1.0, nullptr, nullptr, limit);
        // executed work of this lane (speculative or not) is accounted by the writer after a group sum
        unsigned int w_rhs = solve ? o.nrhs : 0u, w_step = solve ? o.nstep : 0u, w_exec = solve ? 1u : 0u;
        for (int off = 1; off < G; off <<= 1) {
            w_rhs += __shfl_xor_sync(FULL_MASK, w_rhs, off);
            w_step += __shfl_xor_sync(FULL_MASK, w_step, off);
            w_exec += __shfl_xor_sync(FULL_MASK, w_exec, off);
        }
        nrhs += w_rhs; nstep += w_step; nexec += w_exec;
        const int oflags = (inb ? 1 : 0) | (solve ? 2 : 0) | ((o.status & RSFM_CHAIN_EARLY) ? 4 : 0) |
                           ((o.status & ~RSFM_CHAIN_EARLY) << 4);

        // ---- the fit learns from every solve of the tree that ran to the end ----
        // (two groups of one warp may be in different states: every shuffle and warp barrier below is executed by all
        //  lanes, only the arithmetic in between is conditional)
        if (D == 1) {
            const bool use = sur[17] > 0.0;
            const bool pt = use && solve && o.status == 0 && qn[0] > 0.0 && o.sse < 1e300;
            const bool had_fit = sur[12] != 0.0;
            const double x = pt ? (1.0 / qn[0] - sur[14]) * sur[16] : 0.0, y = pt ? o.sse - sur[15] : 0.0;
            const double res = (pt && had_fit) ? y - fma(x, fma(x, sur[10], sur[9]), sur[8]) : 0.0;
            double mo[10] = {pt ? 1.0 : 0.0, x, x * x, x * x * x, (x * x) * (x * x), y, x * y, (x * x) * y,
                             res * res, (pt && had_fit) ? 1.0 : 0.0};
            for (int off = 1; off < G; off <<= 1) {
#pragma unroll
                for (int k = 0; k < 10; k++) mo[k] += __shfl_xor_sync(FULL_MASK, mo[k], off);
            }
            // (the sums are the same bits in every lane of the group: each butterfly level adds the same two numbers)
            const double n_round = mo[0];
#pragma unroll
            for (int k = 0; k < 8; k++) mo[k] += 0.7 * sur[k];                   // the fit forgets older rounds
            double tau2 = sur[11], has_tau = sur[13];
            if (mo[9] > 0.0) {
                const double msr = mo[8] / mo[9];
                tau2 = has_tau != 0.0 ? 0.7 * tau2 + 0.3 * msr : msr;
                has_tau = 1.0;
            }
            // normal equations [[S0 S1 S2] [S1 S2 S3] [S2 S3 S4]] c = (T0 T1 T2) by elimination
            const double l10 = mo[1] / mo[0], l20 = mo[2] / mo[0];
            const double a11 = mo[2] - l10 * mo[1], a12 = mo[3] - l10 * mo[2], b1 = mo[6] - l10 * mo[5];
            const double a22 = mo[4] - l20 * mo[2], b2 = mo[7] - l20 * mo[5];
            const double l21 = a12 / a11;
            const double a22p = a22 - l21 * a12, b2p = b2 - l21 * b1;
            const double c2 = b2p / a22p, c1 = (b1 - a12 * c2) / a11, c0 = (mo[5] - mo[1] * c1 - mo[2] * c2) / mo[0];
            const bool valid = mo[0] >= 6.0 && a11 > 1e-9 * mo[2] && a22p > 1e-9 * mo[4] && c2 > 0.0 &&
                               fabs(c0) < 1e300 && fabs(c1) < 1e300 && c2 < 1e300;
            // first fit: no residual against an earlier fit exists; take the in-sample one (n - 3 degrees of freedom)
            const double r0 = (pt && valid) ? y - fma(x, fma(x, c2, c1), c0) : 0.0;
            double rr = r0 * r0;
            for (int off = 1; off < G; off <<= 1) rr += __shfl_xor_sync(FULL_MASK, rr, off);
            if (has_tau == 0.0 && valid && n_round > 3.5) { tau2 = rr / (n_round - 3.0); has_tau = 1.0; }
            __syncwarp();
            if (use && li == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) sur[k] = mo[k];
                sur[8] = c0; sur[9] = c1; sur[10] = c2; sur[11] = tau2; sur[12] = valid ? 1.0 : 0.0; sur[13] = has_tau;
            }
            __syncwarp();
        }

        if (D == 3) {
            // Ten-coefficient fit: features f = (1, t0, t1, t2, t0^2, t0 t1, t0 t2, t1^2, t1 t2, t2^2).  The group sums
            // f f^T and f y of this round's completed solves entry by entry (rolled loops, f on the stack), slot 0 of
            // the group keeps the normal equations in shared memory (older rounds forgotten by 0.7) and solves them by
            // Cholesky; the other lanes wait at the warp barrier.  Every shuffle is executed by all lanes.
            const bool use = sur[85] > 0.0;
            const bool pt = use && solve && o.status == 0 && qn[0] > 0.0 && qn[D - 1] > 0.0 && o.sse < 1e300;
            double f[10];
            {
                const double ra = pt ? 1.0 / qn[0] : 0.0;
                const double t0 = pt ? (ra - sur[78]) * sur[81] : 0.0, t1 = pt ? (qn[1 % D] * ra - sur[79]) * sur[82] : 0.0;
                const double t2 = pt ? (1.0 / qn[D - 1] - sur[80]) * sur[83] : 0.0;
                f[0] = pt ? 1.0 : 0.0; f[1] = t0; f[2] = t1; f[3] = t2; f[4] = t0 * t0; f[5] = t0 * t1; f[6] = t0 * t2;
                f[7] = t1 * t1; f[8] = t1 * t2; f[9] = t2 * t2;
            }
            const double y = pt ? o.sse - sur[84] : 0.0;
            const bool had_fit = sur[76] != 0.0;
            double pred = 0.0;
#pragma unroll 1
            for (int k = 0; k < 10; k++) pred += f[k] * sur[65 + k];
            const double res = (pt && had_fit) ? y - pred : 0.0;
            double rs = res * res, rn = (pt && had_fit) ? 1.0 : 0.0, n_round = f[0];
            for (int off = 1; off < G; off <<= 1) {
                rs += __shfl_xor_sync(FULL_MASK, rs, off);
                rn += __shfl_xor_sync(FULL_MASK, rn, off);
                n_round += __shfl_xor_sync(FULL_MASK, n_round, off);
            }
            __syncwarp();                                   // the old coefficients have been read by every lane
            int idx = 0;
#pragma unroll 1
            for (int i = 0; i < 10; i++) {
#pragma unroll 1
                for (int j = i; j <= 10; j++) {             // j = 10: the right-hand side f_i y
                    double v = f[i] * (j < 10 ? f[j] : y);
                    for (int off = 1; off < G; off <<= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
                    const int at = j < 10 ? idx++ : 55 + i;
                    if (use && li == 0) sur[at] = 0.7 * sur[at] + v;
                }
            }
            if (use && li == 0) {
                // P c = R by Cholesky (P = U^T U, upper triangle by rows), a relative ridge of 1e-10 on the diagonal
                double U[55], c[10];
                bool ok = sur[0] >= 20.0;
                int p0i = 0;
#pragma unroll 1
                for (int i = 0; i < 10 && ok; i++) {
                    // row i of U: U_ii = sqrt(P_ii - sum_k<i U_ki^2), U_ij = (P_ij - sum_k<i U_ki U_kj) / U_ii
#pragma unroll 1
                    for (int j = i; j < 10; j++) {
                        double a = sur[p0i + (j - i)] * (j == i ? 1.0 + 1e-10 : 1.0);
                        int pk = 0;
#pragma unroll 1
                        for (int k = 0; k < i; k++) { a -= U[pk + (i - k)] * U[pk + (j - k)]; pk += 10 - k; }
                        if (j == i) {
                            if (!(a > 1e-12 * sur[p0i]) || !(a < 1e300)) { ok = false; break; }
                            U[p0i] = sqrt(a);
                        } else {
                            U[p0i + (j - i)] = a / U[p0i];
                        }
                    }
                    p0i += 10 - i;
                }
                if (ok) {
                    // U^T w = R, then U c = w
                    int pi = 0;
#pragma unroll 1
                    for (int i = 0; i < 10; i++) {
                        double a = sur[55 + i];
                        int pk = 0;
#pragma unroll 1
                        for (int k = 0; k < i; k++) { a -= U[pk + (i - k)] * c[k]; pk += 10 - k; }
                        c[i] = a / U[pi];
                        pi += 10 - i;
                    }
#pragma unroll 1
                    for (int i = 9; i >= 0; i--) {
                        int pi2 = 0;
                        for (int k = 0; k < i; k++) pi2 += 10 - k;
                        double a = c[i];
#pragma unroll 1
                        for (int j = i + 1; j < 10; j++) a -= U[pi2 + (j - i)] * c[j];
                        c[i] = a / U[pi2];
                        if (!(fabs(c[i]) < 1e300)) ok = false;
                    }
                }
                if (ok) {
#pragma unroll 1
                    for (int k = 0; k < 10; k++) sur[65 + k] = c[k];
                }
                sur[76] = ok ? 1.0 : 0.0;
            }
            __syncwarp();
            // residual scale: against the previous fit when there was one, else in-sample (n - 10 degrees of freedom)
            const bool valid = sur[76] != 0.0;
            double pred2 = 0.0;
#pragma unroll 1
            for (int k = 0; k < 10; k++) pred2 += f[k] * sur[65 + k];
            const double r0 = (pt && valid) ? y - pred2 : 0.0;
            double rr = r0 * r0;
            for (int off = 1; off < G; off <<= 1) rr += __shfl_xor_sync(FULL_MASK, rr, off);
            if (use && li == 0) {
                double tau2 = sur[75], has_tau = sur[77];
                if (rn > 0.0) { const double msr = rs / rn; tau2 = has_tau != 0.0 ? 0.7 * tau2 + 0.3 * msr : msr; has_tau = 1.0; }
                if (has_tau == 0.0 && valid && n_round > 10.5) { tau2 = rr / (n_round - 10.0); has_tau = 1.0; }
                sur[75] = tau2; sur[77] = has_tau;
            }
            __syncwarp();
        }

        // ---- resolution: every lane of the group walks the realised path (identical arithmetic) ----
        // the acceptance uniform and the gamma draw of iteration it + li once more, one iteration per lane (not kept
        // across the solve: registers), read by shuffle below instead of being drawn by every lane at every level
        double du = 0.5, dlnu2 = 0.0, dgam2 = 1.0;
        if (li < rmax) {
            const unsigned int giter = (unsigned int)(A.iter0 + it + li);
            du = philox_uniform(key, gid, giter, 2u);
            dlnu2 = log(du);
            dgam2 = philox_gamma(key, gid, giter, gshape);
        }
        int cl = gbase, ndone = 0;
        bool stopped = !live || rmax < 1;
        for (int m = 1; m <= G; m++) {
            const int src = cl >= 0 ? cl : lane;
            const int dsrc = gbase + m - 1;
            const double u_m = __shfl_sync(FULL_MASK, du, dsrc), lnu_m = __shfl_sync(FULL_MASK, dlnu2, dsrc);
            const double g0 = __shfl_sync(FULL_MASK, dgam2, dsrc);
            double pq[D];
#pragma unroll
            for (int jj = 0; jj < D; jj++) pq[jj] = __shfl_sync(FULL_MASK, qn[jj], src);
            const double psse = __shfl_sync(FULL_MASK, o.sse, src);
            const double plim = __shfl_sync(FULL_MASK, limit, src);
            const int pf = __shfl_sync(FULL_MASK, oflags, src);
            const unsigned int prhs = __shfl_sync(FULL_MASK, o.nrhs, src), pstep = __shfl_sync(FULL_MASK, o.nstep, src);
            const int pcA = __shfl_sync(FULL_MASK, cA, src), pcR = __shfl_sync(FULL_MASK, cR, src);
            if (stopped || m > rmax) { stopped = true; continue; }
            const bool p_inb = pf & 1, p_early = pf & 4;
            bool acc = false;
            double u = nan("");
            if (p_inb) {
                u = u_m;
                const double lnu = lnu_m;
                const double thr = ss - 2.0 * s2 * lnu;
                // a node stopped at its bound is decided only if that bound is at least the threshold
                // (the root's bound IS its threshold; m > 1 also guarantees progress if the state is NaN)
                if (p_early && m > 1 && !(plim >= thr)) { stopped = true; continue; }
                double la = 0.5 * (ss - psse) / s2;
                if (la > 0.0) la = 0.0;
                acc = la > lnu;
                nsolve++;
                urhs += prhs; ustep += pstep;
                if (p_early) nearly++;
                status |= (pf >> 4);
                if (acc) {
#pragma unroll
                    for (int jj = 0; jj < D; jj++) q[jj] = pq[jj];
                    ss = psse;
                    n_acc++;
                }
            }
            {
                const double bval = 0.5 * (A.n0 * s2 + ss);
                const double scale = 1.0 / bval;
                s2 = 1.0 / (g0 * scale);
            }
            if (writer) {
                const size_t row = (size_t)(it + m - 1);
                if (A.samples) {
#pragma unroll
                    for (int jj = 0; jj < D; jj++) A.samples[(row * D + jj) * Cz + chain] = q[jj];
                }
                if (A.sigma2_out) A.sigma2_out[row * Cz + chain] = s2;
                if (A.accept) A.accept[row * Cz + chain] = acc ? 1 : 0;
                if (A.draws) {
#pragma unroll
                    for (int jj = 0; jj < D; jj++) A.draws[(row * (D + 2) + jj) * Cz + chain] = pq[jj];
                    A.draws[(row * (D + 2) + D) * Cz + chain] = u;
                    A.draws[(row * (D + 2) + D + 1) * Cz + chain] = g0;
                }
            }
            if (A.adapt_mode == RSFM_ADAPT_POOLED) {
#pragma unroll
                for (int jj = 0; jj < D; jj++) sq[jj] += q[jj];
                int t = 0;
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int b = 0; b <= a; b++) sqq[t++] += q[a] * q[b];
            }
            if (COMPAT && writer) S.ring[(size_t)((A.iter0 + it + m) % A.adapt_interval) * Cz + chain] = q[0];
            cl = acc ? pcA : pcR;                           // the node the realised outcome leads to, if it is in the tree
            ndone++;
            if (cl < 0) stopped = true;
        }
        it += ndone;
        if (COMPAT) {
            // MCMC.py:523-527 + 200-204 at the boundary the round just reached (same arithmetic as
            // rsf_mcmc_kernel); every lane of the group then proposes with the new scale
            double lnew = L[0];
            const int W = A.adapt_interval;
            const long long gi = A.iter0 + it;
            if (writer && ndone > 0 && gi % W == 0) {
                double mean = 0.0;
                for (int w = 1; w <= W; w++) mean += S.ring[(size_t)((gi - W + w) % W) * Cz + chain];
                mean /= W;
                double v = 0.0;
                for (int w = 1; w <= W; w++) {
                    const double dlt = S.ring[(size_t)((gi - W + w) % W) * Cz + chain] - mean;
                    v += dlt * dlt;
                }
                v /= (W - 1);
                const double vnew = 2.38 * 2.38 / 2.0 * v;
                if (vnew > 0.0) lnew = sqrt(vnew);
            }
            L[0] = __shfl_sync(FULL_MASK, lnew, gbase);
        }
    }

    if (writer) {
        const int c = chain;
        if (D == 1 && S.fit) {
#pragma unroll
            for (int k = 0; k < SPEC_FIT; k++) S.fit[k * Cz + c] = sur[k];
        }
        if (D == 3 && S.fit) {
#pragma unroll 1
            for (int k = 0; k < SPEC_FIT3; k++) S.fit[k * Cz + c] = sur[k];
        }
#pragma unroll
        for (int j = 0; j < D; j++) S.q[j * Cz + c] = q[j];
        S.sse[c] = ss; S.sigma2[c] = s2;
        if (COMPAT) S.chol[c] = L[0];
        S.accepted[c] += n_acc;
        S.nrhs[c] += nrhs; S.nstep[c] += nstep; S.status[c] |= status; S.nsolve[c] += nsolve; S.nearly[c] += nearly;
        S.nexec[c] += nexec; S.urhs[c] += urhs; S.ustep[c] += ustep;
        if (A.adapt_mode == RSFM_ADAPT_POOLED) {
#pragma unroll
            for (int j = 0; j < D; j++) S.suff[j * Cz + c] += sq[j];
#pragma unroll
            for (int j = 0; j < T; j++) S.suff[(D + j) * Cz + c] += sqq[j];
        }
    }
}

// lanes per chain (2^g) of the speculative kernel for C chains; 0 = the one-thread-per-chain kernel
static int pick_spec_depth(const rsfm_sampler *s, const RunArgs &A)
{
    if (A.deterministic) return 0;
    if (A.n_iters == 1) return 0;                          // nothing to look ahead to (rsfm_spec_depth asks with 0)
    if (s->cfg.sampled_param == RSFM_PARAM_K1) return 0;   // k1 chains: one-thread-per-chain kernel only
    if (s->cfg.adapt_mode == RSFM_ADAPT_COMPAT && s->cfg.n_params != 1) return 0;
    // (a streamed series, n_out > 1,024, is fine: the kernel runs one-warp blocks, its tile barriers are warp-wide)
    const int want = s->cfg.spec_depth;
    if (want == 1) return 0;
    if (want >= 2 && want <= 5) return want;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    // The largest tree that leaves at most two warps per SM sub-partition (all the 255-register kernel can hold).
    // Round 1 (balanced trees, a round advanced g iterations with 2^g lanes) stopped at one warp: the second warp of a
    // sub-partition only shares its FP64 pipe.  With the predictor a round advances nearly as many iterations as the
    // group has lanes, so the second warp's lanes pay (profiles/microbench/spec_predict.py: 1,024 chains 16.6 -> 17.9 M
    // chain-iterations/s from near starts, 10.4 -> 16.7 M from the prior's width; 2,048: 18.1 -> 20.9 M; 4,096: 19.0 -> 22.9 M).
    // (d = 3 has its own fit, ten coefficients in (1/a, b/a, 1/Dc), and follows the same rule:
    //  profiles/microbench/spec_predict3.py)
    const long long one = (long long)sms * 4 * 32;
    const long long cap = 2 * one;
    int g = 0;
    while (g < 5 && ((long long)s->C << (g + 1)) <= cap) g++;
    // d = 1: two lanes per chain (root + its likelier child) still pay -- the one-thread-per-chain kernel is latency
    // bound at one warp per sub-partition (9,473 .. 18,944 chains; cfg 4: 16,384)
    return (g >= 2 || (g == 1 && s->cfg.n_params == 1)) ? g : 0;
}

extern "C" int rsfm_chain_groups(const rsfm_sampler *s) { return s ? s->n_groups : -1; }

extern "C" int rsfm_join(rsfm_sampler *s, void *stream_)
{
    if (!s) return set_err(RSFM_ERR_INVALID, "rsfm_join: NULL sampler%s", "");
    int rc = join_groups(s, (cudaStream_t)stream_);
    if (rc) return rc;
    s->need_fork = 1;
    return RSFM_OK;
}

extern "C" int rsfm_spec_depth(const rsfm_sampler *s)
{
    if (!s) return -1;
    RunArgs A;
    memset(&A, 0, sizeof(A));
    return pick_spec_depth(s, A);
}

static int launch_run(rsfm_sampler *s, RunArgs &A, cudaStream_t stream)
{
    if (!s->initialised) return set_err(RSFM_ERR_STATE, "sampler used before rsfm_init%s", "");
    if (A.n_iters < 1) return set_err(RSFM_ERR_INVALID, "n_iters < 1%s", "");
    const int C = s->C, block = pick_block(C, &s->cfg), grid = (C + block - 1) / block;
    A.iter0 = s->iteration; A.seed = s->seed; A.chain_id0 = s->chain_id0;
    A.a0 = s->cfg.a; A.b0 = s->cfg.b; A.n0 = s->cfg.n0;
    for (int i = 0; i < RSFM_MAX_PARAMS; i++) { A.lo[i] = s->cfg.lo[i]; A.hi[i] = s->cfg.hi[i]; }
    A.adapt_mode = s->cfg.adapt_mode; A.adapt_interval = s->cfg.adapt_interval;
    A.c_begin = 0; A.c_end = C;
    const ModelK M = make_model(&s->cfg);
    const int g = pick_spec_depth(s, A);
    const bool vs = stiff_variant(&s->cfg);
    // Chain groups: Philox-driven launches of the one-thread-per-chain kernel go to the sampler's own streams, one
    // launch per group, and are NOT joined here -- the caller's stream is ordered behind them by the next call on
    // this sampler that needs their results (rsfm_pooled_partials and every getter / setter; rsfm_join).
    const bool k1p = s->cfg.sampled_param == RSFM_PARAM_K1;
    if (s->n_groups > 1 && g < 1 && !A.deterministic && !vs) {
        int rc = fork_groups(s, stream);
        if (rc) return rc;
        const int per = C / s->n_groups, ggrid = (per + block - 1) / block;
        const bool pack = s->cfg.n_params == 3 && block == 128 && s->cfg.round_packing != 1;
        for (int k = 0; k < s->n_groups; k++) {
            A.c_begin = k * per; A.c_end = (k + 1) * per;
            if (k1p) rsf_mcmc_kernel<1, false, false, false, true><<<ggrid, block, 0, s->gstream[k]>>>(M, C, s->d, A);
            else if (s->cfg.n_params == 1) rsf_mcmc_kernel<1, false, false><<<ggrid, block, 0, s->gstream[k]>>>(M, C, s->d, A);
            else if (pack) rsf_mcmc_kernel<3, false, false, true><<<ggrid, block, 0, s->gstream[k]>>>(M, C, s->d, A);
            else rsf_mcmc_kernel<3, false, false><<<ggrid, block, 0, s->gstream[k]>>>(M, C, s->d, A);
        }
        CUDA_TRY(cudaGetLastError());
        s->in_flight = 1;
        s->iteration += A.n_iters;
        s->suff_count += A.n_iters;
        return RSFM_OK;
    }
    {
        int rc = join_groups(s, stream);
        if (rc) return rc;
        s->need_fork = 1;                 // this launch writes the state on the caller's stream
    }
    if (g >= 1) {
        const long long threads = (long long)C << g;
        // one-warp blocks: up to 1,184 warps (two per sub-partition) are spread evenly over the SMs -- 128-thread
        // blocks put two blocks on 108 SMs and one on 40 at 1,024 warps, and the launch waits for the full ones
        const int sblock = 32;
        const int sgrid = (int)((threads + sblock - 1) / sblock);
#define RSFM_SPEC(D, CO)                                                                                              \
    { if (vs) rsf_mcmc_spec_kernel<D, CO, true><<<sgrid, sblock, 0, stream>>>(M, C, s->d, A, g);                      \
      else rsf_mcmc_spec_kernel<D, CO, false><<<sgrid, sblock, 0, stream>>>(M, C, s->d, A, g); }
#define RSFM_SEQ(D, DET)                                                                                              \
    { if (vs) rsf_mcmc_kernel<D, DET, true><<<(C + STIFF_BLOCK - 1) / STIFF_BLOCK, STIFF_BLOCK, 0, stream>>>(M, C, s->d, A); \
      else rsf_mcmc_kernel<D, DET, false><<<grid, block, 0, stream>>>(M, C, s->d, A); }
        if (s->cfg.n_params == 1 && s->cfg.adapt_mode == RSFM_ADAPT_COMPAT) RSFM_SPEC(1, true)
        else if (s->cfg.n_params == 1) RSFM_SPEC(1, false)
        else RSFM_SPEC(3, false)
    } else if (k1p) {
        if (A.deterministic) rsf_mcmc_kernel<1, true, false, false, true><<<grid, block, 0, stream>>>(M, C, s->d, A);
        else rsf_mcmc_kernel<1, false, false, false, true><<<grid, block, 0, stream>>>(M, C, s->d, A);
    } else if (s->cfg.n_params == 1) {
        if (A.deterministic) RSFM_SEQ(1, true) else RSFM_SEQ(1, false)
    } else {
        const bool pack = !A.deterministic && !vs && block == 128 && s->cfg.n_out <= 2 * SERIES_TILE && s->cfg.round_packing != 1;
        if (A.deterministic) RSFM_SEQ(3, true)
        else if (pack) rsf_mcmc_kernel<3, false, false, true><<<grid, block, 0, stream>>>(M, C, s->d, A);
        else RSFM_SEQ(3, false)
    }
#undef RSFM_SPEC
#undef RSFM_SEQ
    CUDA_TRY(cudaGetLastError());
    s->iteration += A.n_iters;
    if (s->cfg.adapt_mode == RSFM_ADAPT_POOLED) s->suff_count += A.n_iters;
    return RSFM_OK;
}

extern "C" int rsfm_run(rsfm_sampler *s, int32_t n_iters, double *samples_out_dev, double *sigma2_out_dev,
                        uint8_t *accept_out_dev, double *draws_out_dev, void *stream)
{
    if (!s) return set_err(RSFM_ERR_INVALID, "rsfm_run: NULL sampler%s", "");
    RunArgs A;
    memset(&A, 0, sizeof(A));
    A.n_iters = n_iters; A.samples = samples_out_dev; A.sigma2_out = sigma2_out_dev; A.accept = accept_out_dev;
    A.draws = draws_out_dev; A.deterministic = 0;
    return launch_run(s, A, (cudaStream_t)stream);
}

extern "C" int rsfm_run_deterministic(rsfm_sampler *s, int32_t n_iters, const double *proposals_dev,
                                      int32_t proposals_are_z, const double *uniforms_dev, const double *gammas_dev,
                                      double *samples_out_dev, double *sigma2_out_dev, uint8_t *accept_out_dev,
                                      void *stream)
{
    if (!s || !proposals_dev || !uniforms_dev || !gammas_dev)
        return set_err(RSFM_ERR_INVALID, "rsfm_run_deterministic: NULL argument%s", "");
    RunArgs A;
    memset(&A, 0, sizeof(A));
    A.n_iters = n_iters; A.samples = samples_out_dev; A.sigma2_out = sigma2_out_dev; A.accept = accept_out_dev;
    A.proposals = proposals_dev; A.uniforms = uniforms_dev; A.gammas = gammas_dev;
    A.proposals_are_z = proposals_are_z; A.deterministic = 1;
    return launch_run(s, A, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// state access
// ---------------------------------------------------------------------------
extern "C" int rsfm_get_state(rsfm_sampler *s, double *q_dev, double *sse_dev, double *sigma2_dev, double *chol_dev,
                              uint32_t *accepted_dev, int32_t *status_dev, uint64_t *nrhs_dev, uint64_t *nstep_dev,
                              void *stream_)
{
    if (!s) return set_err(RSFM_ERR_INVALID, "rsfm_get_state: NULL sampler%s", "");
    cudaStream_t st = (cudaStream_t)stream_;
    { int rc_ = join_groups(s, st); if (rc_) return rc_; s->need_fork = 1; }      // chain groups meet the caller's stream
    const size_t C = s->C; const int d = s->cfg.n_params;
    if (q_dev) CUDA_TRY(cudaMemcpyAsync(q_dev, s->d.q, sizeof(double) * d * C, cudaMemcpyDeviceToDevice, st));
    if (sse_dev) CUDA_TRY(cudaMemcpyAsync(sse_dev, s->d.sse, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
    if (sigma2_dev) CUDA_TRY(cudaMemcpyAsync(sigma2_dev, s->d.sigma2, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
    if (chol_dev) CUDA_TRY(cudaMemcpyAsync(chol_dev, s->d.chol, sizeof(double) * tri(d) * C, cudaMemcpyDeviceToDevice, st));
    if (accepted_dev) CUDA_TRY(cudaMemcpyAsync(accepted_dev, s->d.accepted, sizeof(uint32_t) * C, cudaMemcpyDeviceToDevice, st));
    if (status_dev) CUDA_TRY(cudaMemcpyAsync(status_dev, s->d.status, sizeof(int32_t) * C, cudaMemcpyDeviceToDevice, st));
    if (nrhs_dev) CUDA_TRY(cudaMemcpyAsync(nrhs_dev, s->d.nrhs, sizeof(uint64_t) * C, cudaMemcpyDeviceToDevice, st));
    if (nstep_dev) CUDA_TRY(cudaMemcpyAsync(nstep_dev, s->d.nstep, sizeof(uint64_t) * C, cudaMemcpyDeviceToDevice, st));
    return RSFM_OK;
}

extern "C" int rsfm_set_state(rsfm_sampler *s, const double *q_dev, const double *sse_dev, const double *sigma2_dev,
                              const double *chol_dev, int64_t iteration, void *stream_)
{
    if (!s) return set_err(RSFM_ERR_INVALID, "rsfm_set_state: NULL sampler%s", "");
    if (!s->initialised) return set_err(RSFM_ERR_STATE, "rsfm_set_state before rsfm_init (the data series lives in the sampler)%s", "");
    cudaStream_t st = (cudaStream_t)stream_;
    { int rc_ = join_groups(s, st); if (rc_) return rc_; s->need_fork = 1; }      // chain groups meet the caller's stream
    const size_t C = s->C; const int d = s->cfg.n_params;
    if (q_dev) CUDA_TRY(cudaMemcpyAsync(s->d.q, q_dev, sizeof(double) * d * C, cudaMemcpyDeviceToDevice, st));
    if (sse_dev) CUDA_TRY(cudaMemcpyAsync(s->d.sse, sse_dev, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
    if (sigma2_dev) CUDA_TRY(cudaMemcpyAsync(s->d.sigma2, sigma2_dev, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
    if (chol_dev) CUDA_TRY(cudaMemcpyAsync(s->d.chol, chol_dev, sizeof(double) * tri(d) * C, cudaMemcpyDeviceToDevice, st));
    // the predictor of the speculative kernel starts over around the new state (it never affects results)
    if (s->d.fit) CUDA_TRY(cudaMemsetAsync(s->d.fit, 0, sizeof(double) * s->fit_rows * C, st));
    if (iteration >= 0) s->iteration = iteration;
    return RSFM_OK;
}

static int ring_copy(rsfm_sampler *s, double *dst, const double *src, cudaStream_t st, const char *who)
{
    if (!s || !dst || !src) return set_err(RSFM_ERR_INVALID, "%s: NULL argument", who);
    if (s->cfg.adapt_mode != RSFM_ADAPT_COMPAT) return set_err(RSFM_ERR_INVALID, "%s: the sampler keeps a ring only with RSFM_ADAPT_COMPAT", who);
    { int rc_ = join_groups(s, st); if (rc_) return rc_; s->need_fork = 1; }      // chain groups meet the caller's stream
    CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)s->cfg.adapt_interval * (size_t)s->C, cudaMemcpyDeviceToDevice, st));
    return RSFM_OK;
}

extern "C" int rsfm_get_ring(rsfm_sampler *s, double *ring_dev, void *stream)
{
    return ring_copy(s, ring_dev, s ? s->d.ring : nullptr, (cudaStream_t)stream, "rsfm_get_ring");
}

extern "C" int rsfm_set_ring(rsfm_sampler *s, const double *ring_dev, void *stream)
{
    if (s && !s->initialised) return set_err(RSFM_ERR_STATE, "rsfm_set_ring before rsfm_init%s", "");
    return ring_copy(s, s ? s->d.ring : nullptr, ring_dev, (cudaStream_t)stream, "rsfm_set_ring");
}

// ---------------------------------------------------------------------------
// work totals (forward solves, RHS evaluations, steps, accepted moves, failed chains)
// ---------------------------------------------------------------------------
__global__ void totals_kernel(int C, SamplerDev S, unsigned long long *__restrict__ out)
{
    unsigned long long v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        v[0] += S.nsolve[c]; v[1] += S.nrhs[c]; v[2] += S.nstep[c]; v[3] += S.accepted[c];
        v[4] += S.status[c] != 0 ? 1 : 0; v[5] += S.nearly[c]; v[6] += S.nexec[c]; v[7] += S.urhs[c]; v[8] += S.ustep[c];
    }
#pragma unroll
    for (int j = 0; j < 9; j++) {
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_down_sync(FULL_MASK, v[j], o);
        if ((threadIdx.x & 31) == 0 && v[j]) atomicAdd(&out[j], v[j]);
    }
}

extern "C" int rsfm_get_totals(rsfm_sampler *s, uint64_t *out_host, void *stream_)
{
    if (!s || !out_host) return set_err(RSFM_ERR_INVALID, "rsfm_get_totals: NULL argument%s", "");
    cudaStream_t st = (cudaStream_t)stream_;
    { int rc_ = join_groups(s, st); if (rc_) return rc_; s->need_fork = 1; }      // chain groups meet the caller's stream
    CUDA_TRY(cudaMemsetAsync(s->totals, 0, sizeof(unsigned long long) * 16, st));
    const int grid = (s->C + 255) / 256 < 592 ? (s->C + 255) / 256 : 592;
    totals_kernel<<<grid, 256, 0, st>>>(s->C, s->d, s->totals);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_host, s->totals, sizeof(uint64_t) * 9, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return RSFM_OK;
}

// ---------------------------------------------------------------------------
// pooled sufficient statistics: sum over chains with warp shuffles
// ---------------------------------------------------------------------------
__global__ void suffstats_kernel(int C, int rows, double n, const double *__restrict__ suff, double *__restrict__ out)
{
    // one block per row; grid-stride not needed: blockDim = 1024.  out[0] = n, out[1 + r] = sum of row r
    const int r = blockIdx.x;
    if (r == 0 && threadIdx.x == 0) out[0] = n;
    out += 1;
    double acc = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) acc += suff[(size_t)r * C + c];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL_MASK, acc, o);
    __shared__ double warp_sum[32];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? warp_sum[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
        if (threadIdx.x == 0) out[r] = v;
    }
    (void)rows;
}

extern "C" int rsfm_get_suffstats(rsfm_sampler *s, double *out_dev, int32_t reset, void *stream_)
{
    if (!s || !out_dev) return set_err(RSFM_ERR_INVALID, "rsfm_get_suffstats: NULL argument%s", "");
    cudaStream_t st = (cudaStream_t)stream_;
    { int rc_ = join_groups(s, st); if (rc_) return rc_; s->need_fork = 1; }      // chain groups meet the caller's stream
    const int d = s->cfg.n_params, rows = d + tri(d);
    const double n = (double)s->suff_count * (double)s->C;
    suffstats_kernel<<<rows, 1024, 0, st>>>(s->C, rows, n, s->d.suff, out_dev);
    CUDA_TRY(cudaGetLastError());
    if (reset) {
        CUDA_TRY(cudaMemsetAsync(s->d.suff, 0, sizeof(double) * rows * (size_t)s->C, st));
        s->suff_count = 0;
    }
    return RSFM_OK;
}

__global__ void broadcast_chol_kernel(int C, int T, const double *__restrict__ src, double *__restrict__ chol)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int j = 0; j < T; j++) chol[(size_t)j * C + c] = src[j];
}

extern "C" int rsfm_set_proposal_chol(rsfm_sampler *s, const double *chol_host, void *stream_)
{
    if (!s || !chol_host) return set_err(RSFM_ERR_INVALID, "rsfm_set_proposal_chol: NULL argument%s", "");
    cudaStream_t st = (cudaStream_t)stream_;
    { int rc_ = join_groups(s, st); if (rc_) return rc_; s->need_fork = 1; }      // chain groups meet the caller's stream
    const int T = tri(s->cfg.n_params);
    CUDA_TRY(cudaMemcpyAsync(s->reduce_out, chol_host, sizeof(double) * T, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    broadcast_chol_kernel<<<(s->C + 255) / 256, 256, 0, st>>>(s->C, T, s->reduce_out, s->d.chol);
    CUDA_TRY(cudaGetLastError());
    return RSFM_OK;
}


// ---------------------------------------------------------------------------
// pooled adaptation on the device: sharding-invariant partial sums, moments, Cholesky, install
// ---------------------------------------------------------------------------
// One block per group of RSFM_POOL_GROUP chains aligned on the GLOBAL chain id; every reduction runs in a fixed
// order (shuffle tree inside a warp, then the 32 warp sums through the same tree), so a group's row is the same
// bits on whichever rank, block or launch it is computed.
__global__ void __launch_bounds__(RSFM_POOL_GROUP)
pooled_partials_kernel(int C, int rows, unsigned long long id0, double iters, double *__restrict__ suff,
                       double *__restrict__ out, int block0, int reset)
{
    const int blk = block0 + (int)blockIdx.x;                   // row of the output = 1,024-chain group of the GLOBAL numbering
    const unsigned long long gfirst = (id0 / RSFM_POOL_GROUP + blk) * (unsigned long long)RSFM_POOL_GROUP;
    const long long c = (long long)(gfirst + threadIdx.x) - (long long)id0;
    const bool in = c >= 0 && c < (long long)C;
    __shared__ double ws[32];
    double *o = out + (size_t)blk * RSFM_POOL_ROWS;
    for (int r = -1; r < rows; r++) {
        double v = in ? (r < 0 ? iters : suff[(size_t)r * C + c]) : 0.0;       // r = -1: the sample count
        if (reset && in && r >= 0) suff[(size_t)r * C + c] = 0.0;
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(FULL_MASK, v, off);
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            double w = ws[threadIdx.x];
            for (int off = 16; off > 0; off >>= 1) w += __shfl_down_sync(FULL_MASK, w, off);
            if (threadIdx.x == 0) o[r + 1] = w;
        }
        __syncthreads();
    }
    if (threadIdx.x > rows && threadIdx.x < RSFM_POOL_ROWS) o[threadIdx.x] = 0.0;
}

extern "C" int rsfm_pooled_groups(const rsfm_sampler *s)
{
    if (!s) return -1;
    const unsigned long long first = s->chain_id0 / RSFM_POOL_GROUP, last = (s->chain_id0 + (unsigned long long)s->C - 1) / RSFM_POOL_GROUP;
    return (int)(last - first + 1);
}

extern "C" int rsfm_pooled_partials(rsfm_sampler *s, double *out_dev, int32_t reset, void *stream_)
{
    if (!s || !out_dev) return set_err(RSFM_ERR_INVALID, "rsfm_pooled_partials: NULL argument%s", "");
    cudaStream_t st = (cudaStream_t)stream_;
    const int d = s->cfg.n_params, rows = d + tri(d);
    const int nrow = rsfm_pooled_groups(s);
    if (s->n_groups > 1 && s->in_flight) {
        // every group sums its own rows on its own stream, right behind its iterations; the caller's stream then waits
        // for all of them: this is where the groups' work of the interval meets the caller's stream
        const int per = nrow / s->n_groups;
        for (int g = 0; g < s->n_groups; g++)
            pooled_partials_kernel<<<per, RSFM_POOL_GROUP, 0, s->gstream[g]>>>(s->C, rows, s->chain_id0, (double)s->suff_count,
                                                                               s->d.suff, out_dev, g * per, reset ? 1 : 0);
        CUDA_TRY(cudaGetLastError());
        int rc = join_groups(s, st);
        if (rc) return rc;
    } else {
        pooled_partials_kernel<<<nrow, RSFM_POOL_GROUP, 0, st>>>(s->C, rows, s->chain_id0, (double)s->suff_count, s->d.suff,
                                                                 out_dev, 0, reset ? 1 : 0);
        CUDA_TRY(cudaGetLastError());
        if (reset) s->need_fork = 1;
    }
    if (reset) s->suff_count = 0;
    return RSFM_OK;
}

// moments += sum of the gathered rows (sequential over parts: the order is the global chain order), then the
// Haario proposal (2.38^2/d) cov and its closed-form Cholesky factor (d <= 3).  fac[0] = 1 when a factor was
// formed (finite, positive definite), fac[1..] = the factor (d = 1: the variance).
__global__ void pooled_update_kernel(int d, int n_parts, const double *__restrict__ parts, double *__restrict__ moments,
                                     int accumulate, int install, double *__restrict__ fac)
{
    const int T = d * (d + 1) / 2, rows = 1 + d + T;
    const int r = threadIdx.x;
    if (accumulate && parts != nullptr && r < rows) {
        double a = moments[r];
        for (int p = 0; p < n_parts; p++) a += parts[(size_t)p * RSFM_POOL_ROWS + r];
        moments[r] = a;
    }
    __syncthreads();
    if (r != 0) return;
    fac[0] = 0.0;
    if (!install) return;
    const double n = moments[0];
    if (!(n > (double)(d + 1))) return;
    const double sc = 2.38 * 2.38 / (double)d;
    double mean[3], v[6];
    for (int i = 0; i < d; i++) mean[i] = moments[1 + i] / n;
    bool ok = true;
    int t = 0;
    for (int i = 0; i < d; i++)
        for (int j = 0; j <= i; j++, t++) {
            double c = (moments[1 + d + t] - n * (mean[i] * mean[j])) / (n - 1.0);
            c *= sc;
            if (i == j) { ok = ok && (c > 0.0); c = c + 1e-10 * c; }
            ok = ok && isfinite(c);
            v[t] = c;
        }
    if (!ok) return;
    if (d == 1) { fac[1] = v[0]; fac[0] = 1.0; return; }
    // d = 3: v = (v00, v10, v11, v20, v21, v22)
    const double l00 = sqrt(v[0]), l10 = v[1] / l00, l20 = v[3] / l00;
    const double a11 = v[2] - l10 * l10;
    if (!(a11 > 0.0)) return;
    const double l11 = sqrt(a11), l21 = (v[4] - l20 * l10) / l11;
    const double a22 = v[5] - l20 * l20 - l21 * l21;
    if (!(a22 > 0.0)) return;
    const double l22 = sqrt(a22);
    if (!(isfinite(l10) && isfinite(l20) && isfinite(l21) && isfinite(l22))) return;
    fac[1] = l00; fac[2] = l10; fac[3] = l11; fac[4] = l20; fac[5] = l21; fac[6] = l22;
    fac[0] = 1.0;
}

__global__ void install_chol_kernel(int C, int T, const double *__restrict__ fac, double *__restrict__ chol, int c0, int c1)
{
    if (fac[0] == 0.0) return;                      // no valid factor: the proposal stays (cf. quirk q4)
    const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    for (int j = 0; j < T; j++) chol[(size_t)j * C + c] = fac[1 + j];
}

// The factor in slot `slot` (formed behind fac_event[slot]) replaces every chain's proposal factor.  With chain groups
// each group takes it over on its own stream, between two of its launches (the kernel keeps the factor in registers
// and writes it back when it ends); the wait is for the event of the FORMING of the factor, not for the caller's
// stream as it stands now, so a group is never held up by the other groups' running launches.
static int install_factor(rsfm_sampler *s, int slot, cudaStream_t st)
{
    const int T = tri(s->cfg.n_params);
    if (s->n_groups > 1) {
        // iterations that ran on the caller's stream (host-supplied draws) must be over before a group installs
        { int rc_ = fork_groups(s, st); if (rc_) return rc_; }
        const int per = s->C / s->n_groups;
        for (int g = 0; g < s->n_groups; g++) {
            CUDA_TRY(cudaStreamWaitEvent(s->gstream[g], s->fac_event[slot], 0));
            install_chol_kernel<<<(per + 255) / 256, 256, 0, s->gstream[g]>>>(s->C, T, s->fac_buf[slot], s->d.chol, g * per, (g + 1) * per);
        }
        CUDA_TRY(cudaGetLastError());
        s->in_flight = 1;
    } else {
        CUDA_TRY(cudaStreamWaitEvent(st, s->fac_event[slot], 0));
        install_chol_kernel<<<(s->C + 255) / 256, 256, 0, st>>>(s->C, T, s->fac_buf[slot], s->d.chol, 0, s->C);
        CUDA_TRY(cudaGetLastError());
        s->need_fork = 1;
    }
    return RSFM_OK;
}

extern "C" int rsfm_pooled_update(rsfm_sampler *s, const double *parts_dev, int32_t n_parts, double *moments_dev,
                                  int32_t accumulate, int32_t install, double *factor_out_dev, void *stream_)
{
    if (!s || !moments_dev) return set_err(RSFM_ERR_INVALID, "rsfm_pooled_update: NULL argument%s", "");
    if (parts_dev && n_parts < 0) return set_err(RSFM_ERR_INVALID, "rsfm_pooled_update: n_parts < 0%s", "");
    cudaStream_t st = (cudaStream_t)stream_;
    const int d = s->cfg.n_params, T = tri(d);
    // (double-buffered: the install of a factor may still be queued behind a group's running launch, or not have
    // been asked for yet, when the next factor is formed)
    const int slot = s->fac_parity;
    double *fac = s->fac_buf[slot];
    s->fac_parity ^= 1;
    pooled_update_kernel<<<1, 32, 0, st>>>(d, n_parts, parts_dev, moments_dev, accumulate, install, fac);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->fac_event[slot], st));
    if (factor_out_dev)
        CUDA_TRY(cudaMemcpyAsync(factor_out_dev, fac, sizeof(double) * (1 + T), cudaMemcpyDeviceToDevice, st));
    if (install == 2) { s->staged = slot; return RSFM_OK; }       // formed now, installed by rsfm_pooled_install
    if (install) return install_factor(s, slot, st);
    return RSFM_OK;
}

extern "C" int rsfm_pooled_install(rsfm_sampler *s, void *stream_)
{
    if (!s) return set_err(RSFM_ERR_INVALID, "rsfm_pooled_install: NULL sampler%s", "");
    if (s->staged < 0) return RSFM_OK;                             // nothing staged: the proposal stays
    const int slot = s->staged;
    s->staged = -1;
    return install_factor(s, slot, (cudaStream_t)stream_);
}

// ---------------------------------------------------------------------------
// test hooks: the device's own Philox / samplers / RHS, element by element (tests compare them with the CPU oracle)
// ---------------------------------------------------------------------------
__global__ void philox_raw_kernel(int n, const uint32_t *__restrict__ in, uint32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PhiloxKey k; k.k0 = in[6 * i + 4]; k.k1 = in[6 * i + 5];
    const uint4 r = philox4x32_10(make_uint4(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3]), k);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

extern "C" int rsfm_philox_raw(const uint32_t *in_dev, uint32_t *out_dev, int32_t n, void *stream)
{
    if (!in_dev || !out_dev || n < 1) return set_err(RSFM_ERR_INVALID, "rsfm_philox_raw: bad argument%s", "");
    int rc = require_device();
    if (rc) return rc;
    philox_raw_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, in_dev, out_dev);
    CUDA_TRY(cudaGetLastError());
    return RSFM_OK;
}

__global__ void philox_draws_kernel(unsigned long long seed, unsigned long long id0, int C, unsigned int iter0,
                                    int n_iters, double shape, double *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const PhiloxKey key = philox_key(seed);
    const unsigned long long gid = id0 + (unsigned long long)c;
    for (int i = 0; i < n_iters; i++) {
        const unsigned int it = iter0 + (unsigned int)i;
        double z0, z1, z2, zz;
        philox_normal2(key, gid, it, 0u, z0, z1);
        philox_normal2(key, gid, it, 1u, z2, zz);
        const double u = philox_uniform(key, gid, it, 2u);
        uint32_t tries = 0;
        const double g = philox_gamma(key, gid, it, shape, &tries);
        double *o = out + (size_t)i * 6 * C + c;
        o[0] = z0; o[(size_t)C] = z1; o[2 * (size_t)C] = z2; o[3 * (size_t)C] = u; o[4 * (size_t)C] = g;
        o[5 * (size_t)C] = (double)tries;
    }
}

extern "C" int rsfm_philox_draws(uint64_t seed, uint64_t chain_id0, int32_t C, uint32_t iter0, int32_t n_iters,
                                 double gamma_shape, double *out_dev, void *stream)
{
    if (!out_dev || C < 1 || n_iters < 1 || !(gamma_shape > 1.0))
        return set_err(RSFM_ERR_INVALID, "rsfm_philox_draws: bad argument%s", "");
    int rc = require_device();
    if (rc) return rc;
    philox_draws_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, chain_id0, C, iter0, n_iters, gamma_shape, out_dev);
    CUDA_TRY(cudaGetLastError());
    return RSFM_OK;
}

__global__ void rhs_eval_kernel(const __grid_constant__ ModelK M, int n, double a0, double b0, const double *__restrict__ t,
                                const double *__restrict__ mu, const double *__restrict__ th,
                                const double *__restrict__ dc, const double *__restrict__ a, const double *__restrict__ b,
                                int general, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ChainConst cc = make_chain_const(M, a ? a[i] : a0, b ? b[i] : b0, dc[i]);
    const double L = loading_of(M, t[i]);
    double rth = 1.0 / th[i], dmu, dth, dV;
    if (general) {
        bool unused = false;
        rsf_rhs<false>(cc, L, mu[i], th[i], rth, dmu, dth, dV, unused);
    } else {
        rsf_rhs_checked(cc, L, mu[i], th[i], rth, dmu, dth, dV);
    }
    out[i] = dmu; out[(size_t)n + i] = dth; out[2 * (size_t)n + i] = dV;
}

extern "C" int rsfm_rhs_eval(const rsfm_cfg *cfg, int32_t n, const double *t_dev, const double *mu_dev,
                             const double *theta_dev, const double *dc_dev, const double *a_dev, const double *b_dev,
                             int32_t general, double *out_dev, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (n < 1 || !t_dev || !mu_dev || !theta_dev || !dc_dev || !out_dev)
        return set_err(RSFM_ERR_INVALID, "rsfm_rhs_eval: bad argument%s", "");
    rc = require_device();
    if (rc) return rc;
    const ModelK M = make_model(cfg);
    rhs_eval_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(M, n, cfg->a, cfg->b, t_dev, mu_dev, theta_dev, dc_dev,
                                                                        a_dev, b_dev, general, out_dev);
    CUDA_TRY(cudaGetLastError());
    return RSFM_OK;
}

// ---------------------------------------------------------------------------
// per-chain diagnostics: mean, variance, ESS (Geyer initial positive sequence)
// ---------------------------------------------------------------------------
__global__ void chain_diag_kernel(const double *__restrict__ x, int n, int d, int C, int p, int max_lag,
                                  double *__restrict__ mean_out, double *__restrict__ var_out,
                                  double *__restrict__ ess_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const size_t stride = (size_t)d * C;
    const double *xc = x + (size_t)p * C + c;
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean += xc[i * stride];
    mean /= n;
    double c0 = 0.0;
    for (int i = 0; i < n; i++) { const double e = xc[i * stride] - mean; c0 += e * e; }
    if (mean_out) mean_out[c] = mean;
    if (var_out) var_out[c] = n > 1 ? c0 / (n - 1) : 0.0;
    if (!ess_out) return;
    if (!(c0 > 0.0)) { ess_out[c] = (double)n; return; }      // constant chain: no information, report n
    // rho_t = c_t / c_0 with c_t = sum_{i<n-t} (x_i - m)(x_{i+t} - m);  tau = -1 + 2 sum_{pairs} P_k,
    // P_k = rho_{2k} + rho_{2k+1}, truncated at the first non-positive pair and made monotone.
    double tau = 0.0, prev_pair = 1e300;
    const int L = max_lag < n ? max_lag : n - 1;
    // Eight lags (four pairs) per pass over the series: a register window e_{i+t0} .. e_{i+t0+7} slides along, so a
    // pass costs two loads per draw for eight lags instead of three per pair.  Every c_t is the same left-to-right
    // sum as before (same bits); the pairs are then judged in order and the first non-positive one ends the sum.
    bool done = false;
    for (int t0 = 0; t0 + 1 <= L && !done; t0 += 8) {
        double cs[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        double w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = (t0 + j < n) ? xc[(size_t)(t0 + j) * stride] - mean : 0.0;
        for (int i = 0; i + t0 < n; i++) {
            const double ei = xc[(size_t)i * stride] - mean;
#pragma unroll
            for (int j = 0; j < 8; j++) cs[j] += ei * w[j];           // w[j] = e_{i+t0+j}, 0 beyond the end
#pragma unroll
            for (int j = 0; j < 7; j++) w[j] = w[j + 1];
            w[7] = (i + t0 + 8 < n) ? xc[(size_t)(i + t0 + 8) * stride] - mean : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            if (done || t0 + j + 1 > L) { done = true; continue; }
            double pair = (cs[j] + cs[j + 1]) / c0;
            if (pair <= 0.0) { done = true; continue; }
            if (pair > prev_pair) pair = prev_pair;
            prev_pair = pair;
            tau += 2.0 * pair;
        }
    }
    tau -= 1.0;
    if (tau < 1.0 / n) tau = 1.0 / n;
    ess_out[c] = n / tau;
}

extern "C" int rsfm_chain_diagnostics(const double *samples_dev, int32_t n, int32_t d, int32_t C, int32_t p,
                                      int32_t max_lag, double *mean_dev, double *var_dev, double *ess_dev,
                                      void *stream)
{
    if (!samples_dev || n < 2 || d < 1 || C < 1 || p < 0 || p >= d)
        return set_err(RSFM_ERR_INVALID, "rsfm_chain_diagnostics: bad argument%s", "");
    int rc = require_device();
    if (rc) return rc;
    chain_diag_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(samples_dev, n, d, C, p, max_lag, mean_dev,
                                                                          var_dev, ess_dev);
    CUDA_TRY(cudaGetLastError());
    return RSFM_OK;
}

// ---------------------------------------------------------------------------
// Gaussian kernel density estimate on a grid (posterior post-processing, RSF.py:734-737)
// ---------------------------------------------------------------------------
// pdf(x_g) = 1/(n bw sqrt(2 pi)) sum_i exp(-(x_g - x_i)^2 / (2 bw^2)); one block per grid point.
__global__ void __launch_bounds__(256) kde_grid_kernel(const double *__restrict__ x, long long n,
                                                        const double *__restrict__ grid, int G, double bw,
                                                        double *__restrict__ pdf)
{
    const int gidx = blockIdx.x;
    if (gidx >= G) return;
    const double xg = grid[gidx];
    const double inv2 = -0.5 / (bw * bw);
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = xg - x[i];
        acc += exp(d * d * inv2);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL_MASK, acc, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < 8 ? ws[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
        if (threadIdx.x == 0) pdf[gidx] = v / ((double)n * bw * 2.5066282746310002);
    }
}

extern "C" int rsfm_kde_grid(const double *samples_dev, int64_t n, const double *grid_dev, int32_t G, double bandwidth,
                             double *pdf_out_dev, void *stream)
{
    if (!samples_dev || !grid_dev || !pdf_out_dev || n < 1 || G < 1 || !(bandwidth > 0.0))
        return set_err(RSFM_ERR_INVALID, "rsfm_kde_grid: bad argument%s", "");
    int rc = require_device();
    if (rc) return rc;
    kde_grid_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(samples_dev, (long long)n, grid_dev, G, bandwidth, pdf_out_dev);
    CUDA_TRY(cudaGetLastError());
    return RSFM_OK;
}

// ---------------------------------------------------------------------------
// FP64 peak: 8 independent DFMA chains per thread, all SMs full
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double *sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 12345.678) sink[0] = r;     // never true; keeps the chains alive
}

extern "C" int rsfm_measure_fp64_peak(double millis, double *flops_out)
{
    if (!flops_out) return set_err(RSFM_ERR_INVALID, "rsfm_measure_fp64_peak: NULL%s", "");
    int rc = require_device();
    if (rc) return rc;
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *sink = nullptr;
    CUDA_TRY(cudaMalloc((void **)&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int grid = sms * 8, block = 256;
    int iters = 2000;
    float ms = 0.f;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        dfma_peak_kernel<<<grid, block>>>(iters, 1.0, sink);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * (double)block / (ms * 1e-3);
        if (rep > 0 && flops > best) best = flops;
        if (ms < millis) iters = (int)fmin(2.0e6, iters * fmax(1.5, millis / fmax(ms, 1e-3)));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *flops_out = best;
    return RSFM_OK;
}

#ifdef RSFM_DEBUG_COUNT
extern "C" int rsfm_debug_counters(unsigned long long *out16, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, rsfm::g_dbg, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(rsfm::g_dbg, z, sizeof(z)); }
    return 0;
}
extern "C" int rsfm_debug_records(double *out512, unsigned int *n_out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out512, rsfm::g_dbgrec, sizeof(double) * 512);
    cudaMemcpyFromSymbol(n_out, rsfm::g_dbgrec_n, sizeof(unsigned int));
    if (reset) { unsigned int z = 0; cudaMemcpyToSymbol(rsfm::g_dbgrec_n, &z, sizeof(z)); }
    return 0;
}
#endif
