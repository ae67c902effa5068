// rsfm_device.cuh -- device-side RSF spring-slider solver for sm_100a.
//
// One thread integrates one chain's 3-state ODE (mu, theta, V) in FP64 with
// DOP853 and per-chain adaptive steps, following the step logic of SciPy's
// ode('dop853') as the reference drives it (RateStateModel.py:374-389, SURVEY.md
// Appendix B): a fresh call per output interval, hinit, err/accept, controller.
//
// B200-specific structure (not present in any form in the reference):
//   * V (y[2]) never feeds back into the RHS (RateStateModel.py:336-346 use only
//     mu, theta), so its stage values are not stored: the three linear forms that
//     need them (8th-order sum, 5th- and 3rd-order error forms) are accumulated
//     as the stages are produced.  Stage storage is 10 x (mu, theta) registers.
//   * the load-point velocity V_l(t) depends on time only.  Whenever all lanes of
//     a warp are at the same (t, h) -- always, outside the stiff regime -- lanes
//     0..10 evaluate the eleven stage abscissae once and every lane fetches its
//     stage value with a shuffle: one exp + one sin per step instead of eleven.
//   * the observed series streams through shared memory in 4 KB tiles fetched by
//     the TMA bulk-copy engine (cp.async.bulk + mbarrier), double-buffered.
//   * restarting k1 = f(t, y) at every interval is skipped: it is bit-identical
//     to the FSAL evaluation that closed the previous interval.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "dop853_coeffs.h"
#include "rsfm.h"

namespace rsfm {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int SERIES_TILE = 512;          // doubles per staged tile (4 KB)

struct ModelK {
    double mu_ref, V_ref, k1, t_start, delta_t, mu_t_zero;
    double rtol, atol, vstep_period, vstep_factor;
    int n_out, nmax, damping, loading, integ_mode;
};

struct ChainConst {
    double b, inv_a, inv_dc, kprime;
};

__device__ __forceinline__ ChainConst make_chain_const(double a, double b, double dc)
{
    ChainConst c;
    c.b = b;
    c.inv_a = 1.0 / a;
    c.inv_dc = 1.0 / dc;
    c.kprime = 1e-2 * 10 / dc;            // RateStateModel.py:324
    return c;
}

// ---------------------------------------------------------------------------
// TMA bulk-copy staging of the observed series
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Double-buffered tile stream over data[0 .. n).  Every thread of the block calls
// begin() once per kernel, start_solve() before each pass over the series and
// at(k) for k = 0, 1, 2, ... in order (block-uniform k).  A series of at most two
// tiles (n <= 1024, every reference configuration with N = 500) is fetched once
// and stays resident for all later passes of the same kernel.
struct SeriesStage {
    double *buf;        // [2][SERIES_TILE] shared
    uint64_t *bar;      // [2] shared
    const double *g;    // global, 16-byte aligned
    int n;
    uint32_t par0, par1;   // phase parity to wait for next, per barrier
    bool resident, loaded;

    __device__ __forceinline__ void issue(int tile)
    {
        // single thread
        const int k0 = tile * SERIES_TILE;
        int cnt = n - k0;
        if (cnt > SERIES_TILE) cnt = SERIES_TILE;
        const int even = cnt & ~1;              // bulk copies move multiples of 16 bytes
        double *dst = buf + (tile & 1) * SERIES_TILE;
        uint64_t *b = bar + (tile & 1);
        if (even > 0) {
            mbar_expect_tx(b, (uint32_t)even * 8u);
            bulk_g2s(dst, g + k0, (uint32_t)even * 8u, b);
        } else {
            mbar_arrive(b);
        }
        if (cnt & 1) dst[cnt - 1] = g[k0 + cnt - 1];   // odd tail element, ordered by the next barrier
    }

    __device__ __forceinline__ void begin(double *smem_buf, uint64_t *smem_bar, const double *data, int n_)
    {
        buf = smem_buf; bar = smem_bar; g = data; n = n_;
        par0 = 0; par1 = 0;
        resident = n_ <= 2 * SERIES_TILE;
        loaded = false;
        if (g == nullptr) return;
        if (threadIdx.x == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    __device__ __forceinline__ void start_solve()
    {
        if (g == nullptr || (resident && loaded)) return;
        __syncthreads();                               // readers of the previous pass are done
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(0);
            if (n > SERIES_TILE) issue(1);
        }
        __syncthreads();                               // odd-tail plain stores visible
    }

    // value data[k]; in streaming mode entering a tile recycles the buffer two tiles back
    __device__ __forceinline__ double at(int k)
    {
        const int tile = k / SERIES_TILE;
        const int off = k - tile * SERIES_TILE;
        if (off == 0 && !(resident && loaded)) {
            if (!resident && tile > 0) {
                __syncthreads();                       // everyone is done with tile-1
                if (threadIdx.x == 0 && (tile + 1) * SERIES_TILE < n) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(tile + 1);                   // overwrites the buffer of tile-1
                }
                __syncthreads();                       // tail element of this tile visible
            }
            if (tile & 1) { while (!mbar_try_wait(&bar[1], par1)) { } par1 ^= 1u; }
            else          { while (!mbar_try_wait(&bar[0], par0)) { } par0 ^= 1u; }
            if (resident && (tile + 1) * SERIES_TILE >= n) loaded = true;
        }
        return buf[(tile & 1) * SERIES_TILE + off];
    }
};

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
// load-point velocity, RateStateModel.py:327-329 (SINE_DECAY) / VSTEP extension
__device__ __noinline__ double loading_velocity(int loading, double V_ref, double t_start, double period,
                                                double factor, double t)
{
    if (loading == RSFM_LOAD_VSTEP) {
        const double ph = floor((t - t_start) / period);
        const long long i = (long long)ph;
        return (i & 1) ? factor * V_ref : V_ref;
    }
    return V_ref * (1.0 + exp(-t / 20.0) * sin(10.0 * t));
}

__device__ __forceinline__ double loading_of(const ModelK &M, double t)
{
    return loading_velocity(M.loading, M.V_ref, M.t_start, M.vstep_period, M.vstep_factor, t);
}

// friction(t, y), RateStateModel.py:336-353; V_l is passed in (time-only term)
__device__ __forceinline__ void rsf_rhs(const ModelK &M, const ChainConst &c, double vl, double mu, double th,
                                        double &dmu, double &dth, double &dV)
{
    const double temp = c.inv_a * (mu - M.mu_ref - c.b * log(M.V_ref * th * c.inv_dc));   // :336
    const double v = M.V_ref * exp(temp);                                                // :337
    dth = 1.0 - v * th * c.inv_dc;                                                       // :340
    dmu = c.kprime * vl - c.kprime * v;                                                  // :343
    const double voa = v * c.inv_a;
    const double s = (c.b / th) * dth;
    dV = voa * (dmu - s);                                                                // :346
    if (M.damping) {                                                                     // :349-353
        dmu = dmu - M.k1 * dV;
        dV = voa * (dmu - s);
    }
}

__device__ __forceinline__ double root8(double x) { return sqrt(sqrt(sqrt(x))); }

__device__ __forceinline__ double stage_c(int i)   // abscissa of stage i+2, i = 0..10
{
    switch (i) {
        case 0: return DP_C2;
        case 1: return DP_C3;
        case 2: return DP_C4;
        case 3: return DP_C5;
        case 4: return DP_C6;
        case 5: return DP_C7;
        case 6: return DP_C8;
        case 7: return DP_C9;
        case 8: return DP_C10;
        case 9: return DP_C11;
        default: return 1.0;
    }
}

struct SolveOut {
    double sse;
    int status;
    int filled;
    uint32_t nrhs, nstep;
};

// Integrate one chain over the whole output grid.  All 32 lanes of a warp must
// call this together (it contains warp collectives) and all threads of a block
// must call it together when `series.g != nullptr` (block barriers at tile
// boundaries).  `active` = this lane owns a real chain.
// acc_out / t_out (optional) are written time-major with stride `acc_stride` (= C).
// acc_ref (optional, same layout) and xtx: if acc_ref != nullptr the solve also
// accumulates xtx += ((acc - acc_ref[k]) / fd_den)^2  (MCMC.py:264-265).
__device__ __forceinline__ SolveOut rsf_solve(const ModelK &M, double a, double b, double dc, bool active,
                                               SeriesStage &series, double *acc_out, const double *acc_ref,
                                               size_t acc_stride, double fd_den, double *xtx_out,
                                               double *t_out = nullptr)
{
    const int lane = threadIdx.x & 31;
    const ChainConst cc = make_chain_const(a, b, dc);
    const bool have_data = series.g != nullptr;
    const double uround = 2.3e-16, safe = 0.9;
    const double facc1 = 1.0 / 0.3, facc2 = 1.0 / 6.0;

    double t = M.t_start;
    double mu = M.mu_t_zero, th = dc / M.V_ref, V = M.V_ref;       // :367-370,377
    double k1m, k1t, k1v;
    SolveOut out;
    out.status = RSFM_CHAIN_OK;
    out.filled = M.n_out;
    out.nrhs = 0; out.nstep = 0;
    double sse = 0.0, xtx = 0.0;
    if (have_data) { const double d0 = series.at(0); sse = d0 * d0; }   // acc[0] = 0  (:371)
    if (acc_out && active) acc_out[0] = 0.0;
    if (t_out && active) t_out[0] = t;
    if (acc_ref && active) { const double x0 = (0.0 - acc_ref[0]) / fd_den; xtx = x0 * x0; }

    rsf_rhs(M, cc, loading_of(M, t), mu, th, k1m, k1t, k1v);
    out.nrhs++;
    bool failed = false;
    double vprev = V;
    double h_carry = 0.0;

    for (int k = 1; k < M.n_out; k++) {
        const double dk = have_data ? series.at(k) : 0.0;
        const bool running = active && !failed;
        const double xend = t + M.delta_t;                         // :382
        const double hmax = fabs(xend - t);
        int nstep_call = 0;
        bool reject = false, last = false;
        double h;

        // ---- hinit (dop853.f HINIT, iord = 8) or carried step ----
        if (M.integ_mode == RSFM_INTEG_PARITY || k == 1) {
            const double i0 = 1.0 / (M.atol + M.rtol * fabs(mu));
            const double i1 = 1.0 / (M.atol + M.rtol * fabs(th));
            const double i2 = 1.0 / (M.atol + M.rtol * fabs(V));
            const double dnf = (k1m * i0) * (k1m * i0) + (k1t * i1) * (k1t * i1) + (k1v * i2) * (k1v * i2);
            const double dny = (mu * i0) * (mu * i0) + (th * i1) * (th * i1) + (V * i2) * (V * i2);
            double h0 = (dnf <= 1e-10 || dny <= 1e-10) ? 1.0e-6 : sqrt(dny / dnf) * 0.01;
            h0 = fmin(h0, hmax);
            // probe loading at t + h0: shared when the running lanes agree on (t, h0)
            const unsigned m = __ballot_sync(FULL_MASK, running);
            double vlp = 0.0;
            if (m != 0) {
                const int src = __ffs(m) - 1;
                const double ts = __shfl_sync(FULL_MASK, t, src), hs = __shfl_sync(FULL_MASK, h0, src);
                const bool uni = __all_sync(FULL_MASK, !running || (t == ts && h0 == hs));
                if (uni) {
                    if (lane == src) vlp = loading_of(M, t + h0);
                    vlp = __shfl_sync(FULL_MASK, vlp, src);
                } else if (running) {
                    vlp = loading_of(M, t + h0);
                }
            }
            double f1m, f1t, f1v;
            rsf_rhs(M, cc, vlp, mu + h0 * k1m, th + h0 * k1t, f1m, f1t, f1v);
            if (running) out.nrhs++;
            const double e0 = (f1m - k1m) * i0, e1 = (f1t - k1t) * i1, e2 = (f1v - k1v) * i2;
            const double der2 = sqrt(e0 * e0 + e1 * e1 + e2 * e2) / h0;
            const double der12 = fmax(fabs(der2), sqrt(dnf));
            const double h1 = (der12 <= 1e-15) ? fmax(1.0e-6, fabs(h0) * 1.0e-3) : root8(0.01 / der12);
            h = fmin(fmin(100.0 * fabs(h0), h1), hmax);
        } else {
            h = fmin(h_carry, hmax);
        }

        // ---- dp86co step loop ----
        bool done = !running;
        for (;;) {
            if (!done) {
                if (nstep_call > M.nmax) { failed = true; done = true; out.status = RSFM_CHAIN_NMAX; }
                else if (0.1 * fabs(h) <= fabs(t) * uround) { failed = true; done = true; out.status = RSFM_CHAIN_HSMALL; }
                else {
                    if ((t + 1.01 * h - xend) > 0.0) { h = xend - t; last = true; }
                    nstep_call++;
                }
            }
            const unsigned m = __ballot_sync(FULL_MASK, !done);
            if (m == 0) break;
            const bool stepping = !done;

            // stage-time loading table, shared across the warp when (t, h) agree
            const int src = __ffs(m) - 1;
            const double ts = __shfl_sync(FULL_MASK, t, src), hs = __shfl_sync(FULL_MASK, h, src);
            const bool uni = __all_sync(FULL_MASK, !stepping || (t == ts && h == hs));
            double vl_tab = 0.0;
            if (uni && lane < 11) vl_tab = loading_of(M, __dadd_rn(ts, __dmul_rn(stage_c(lane), hs)));
#define RSFM_VL(i)                                   \
    (uni ? __shfl_sync(FULL_MASK, vl_tab, (i))      \
         : (stepping ? loading_of(M, __dadd_rn(t, __dmul_rn(stage_c(i), h))) : 0.0))

            double k2m, k2t, k3m, k3t, k4m, k4t, k5m, k5t, k6m, k6t, k7m, k7t, k8m, k8t, k9m, k9t, k10m, k10t;
            double kv, k9v, k12v, bV, eV;
            // stage 2..5 (their V-derivatives carry zero weight everywhere)
            rsf_rhs(M, cc, RSFM_VL(0), mu + h * DP_A2_1 * k1m, th + h * DP_A2_1 * k1t, k2m, k2t, kv);
            rsf_rhs(M, cc, RSFM_VL(1), mu + h * (DP_A3_1 * k1m + DP_A3_2 * k2m),
                    th + h * (DP_A3_1 * k1t + DP_A3_2 * k2t), k3m, k3t, kv);
            rsf_rhs(M, cc, RSFM_VL(2), mu + h * (DP_A4_1 * k1m + DP_A4_3 * k3m),
                    th + h * (DP_A4_1 * k1t + DP_A4_3 * k3t), k4m, k4t, kv);
            rsf_rhs(M, cc, RSFM_VL(3), mu + h * (DP_A5_1 * k1m + DP_A5_3 * k3m + DP_A5_4 * k4m),
                    th + h * (DP_A5_1 * k1t + DP_A5_3 * k3t + DP_A5_4 * k4t), k5m, k5t, kv);
            bV = DP_B1 * k1v;
            eV = DP_ER1 * k1v;
            rsf_rhs(M, cc, RSFM_VL(4), mu + h * (DP_A6_1 * k1m + DP_A6_4 * k4m + DP_A6_5 * k5m),
                    th + h * (DP_A6_1 * k1t + DP_A6_4 * k4t + DP_A6_5 * k5t), k6m, k6t, kv);
            bV += DP_B6 * kv; eV += DP_ER6 * kv;
            rsf_rhs(M, cc, RSFM_VL(5), mu + h * (DP_A7_1 * k1m + DP_A7_4 * k4m + DP_A7_5 * k5m + DP_A7_6 * k6m),
                    th + h * (DP_A7_1 * k1t + DP_A7_4 * k4t + DP_A7_5 * k5t + DP_A7_6 * k6t), k7m, k7t, kv);
            bV += DP_B7 * kv; eV += DP_ER7 * kv;
            rsf_rhs(M, cc, RSFM_VL(6),
                    mu + h * (DP_A8_1 * k1m + DP_A8_4 * k4m + DP_A8_5 * k5m + DP_A8_6 * k6m + DP_A8_7 * k7m),
                    th + h * (DP_A8_1 * k1t + DP_A8_4 * k4t + DP_A8_5 * k5t + DP_A8_6 * k6t + DP_A8_7 * k7t),
                    k8m, k8t, kv);
            bV += DP_B8 * kv; eV += DP_ER8 * kv;
            rsf_rhs(M, cc, RSFM_VL(7),
                    mu + h * (DP_A9_1 * k1m + DP_A9_4 * k4m + DP_A9_5 * k5m + DP_A9_6 * k6m + DP_A9_7 * k7m +
                              DP_A9_8 * k8m),
                    th + h * (DP_A9_1 * k1t + DP_A9_4 * k4t + DP_A9_5 * k5t + DP_A9_6 * k6t + DP_A9_7 * k7t +
                              DP_A9_8 * k8t),
                    k9m, k9t, k9v);
            bV += DP_B9 * k9v; eV += DP_ER9 * k9v;
            rsf_rhs(M, cc, RSFM_VL(8),
                    mu + h * (DP_A10_1 * k1m + DP_A10_4 * k4m + DP_A10_5 * k5m + DP_A10_6 * k6m + DP_A10_7 * k7m +
                              DP_A10_8 * k8m + DP_A10_9 * k9m),
                    th + h * (DP_A10_1 * k1t + DP_A10_4 * k4t + DP_A10_5 * k5t + DP_A10_6 * k6t + DP_A10_7 * k7t +
                              DP_A10_8 * k8t + DP_A10_9 * k9t),
                    k10m, k10t, kv);
            bV += DP_B10 * kv; eV += DP_ER10 * kv;
            // stage 11 -> k2 slot
            rsf_rhs(M, cc, RSFM_VL(9),
                    mu + h * (DP_A11_1 * k1m + DP_A11_4 * k4m + DP_A11_5 * k5m + DP_A11_6 * k6m + DP_A11_7 * k7m +
                              DP_A11_8 * k8m + DP_A11_9 * k9m + DP_A11_10 * k10m),
                    th + h * (DP_A11_1 * k1t + DP_A11_4 * k4t + DP_A11_5 * k5t + DP_A11_6 * k6t + DP_A11_7 * k7t +
                              DP_A11_8 * k8t + DP_A11_9 * k9t + DP_A11_10 * k10t),
                    k2m, k2t, kv);
            bV += DP_B11 * kv; eV += DP_ER11 * kv;
            // stage 12 -> k3 slot, at xph = t + h
            const double vl12 = RSFM_VL(10);
            const double xph = t + h;
            rsf_rhs(M, cc, vl12,
                    mu + h * (DP_A12_1 * k1m + DP_A12_4 * k4m + DP_A12_5 * k5m + DP_A12_6 * k6m + DP_A12_7 * k7m +
                              DP_A12_8 * k8m + DP_A12_9 * k9m + DP_A12_10 * k10m + DP_A12_11 * k2m),
                    th + h * (DP_A12_1 * k1t + DP_A12_4 * k4t + DP_A12_5 * k5t + DP_A12_6 * k6t + DP_A12_7 * k7t +
                              DP_A12_8 * k8t + DP_A12_9 * k9t + DP_A12_10 * k10t + DP_A12_11 * k2t),
                    k3m, k3t, k12v);
            bV += DP_B12 * k12v; eV += DP_ER12 * k12v;
#undef RSFM_VL
            // 8th-order increment and new state
            const double bM = DP_B1 * k1m + DP_B6 * k6m + DP_B7 * k7m + DP_B8 * k8m + DP_B9 * k9m + DP_B10 * k10m +
                              DP_B11 * k2m + DP_B12 * k3m;
            const double bT = DP_B1 * k1t + DP_B6 * k6t + DP_B7 * k7t + DP_B8 * k8t + DP_B9 * k9t + DP_B10 * k10t +
                              DP_B11 * k2t + DP_B12 * k3t;
            const double muN = mu + h * bM, thN = th + h * bT, VN = V + h * bV;
            // error estimate
            const double s0 = 1.0 / (M.atol + M.rtol * fmax(fabs(mu), fabs(muN)));
            const double s1 = 1.0 / (M.atol + M.rtol * fmax(fabs(th), fabs(thN)));
            const double s2 = 1.0 / (M.atol + M.rtol * fmax(fabs(V), fabs(VN)));
            const double e3m = (bM - DP_BHH1 * k1m - DP_BHH2 * k9m - DP_BHH3 * k3m) * s0;
            const double e3t = (bT - DP_BHH1 * k1t - DP_BHH2 * k9t - DP_BHH3 * k3t) * s1;
            const double e3v = (bV - DP_BHH1 * k1v - DP_BHH2 * k9v - DP_BHH3 * k12v) * s2;
            const double e5m = (DP_ER1 * k1m + DP_ER6 * k6m + DP_ER7 * k7m + DP_ER8 * k8m + DP_ER9 * k9m +
                                DP_ER10 * k10m + DP_ER11 * k2m + DP_ER12 * k3m) * s0;
            const double e5t = (DP_ER1 * k1t + DP_ER6 * k6t + DP_ER7 * k7t + DP_ER8 * k8t + DP_ER9 * k9t +
                                DP_ER10 * k10t + DP_ER11 * k2t + DP_ER12 * k3t) * s1;
            const double e5v = eV * s2;
            const double err2 = e3m * e3m + e3t * e3t + e3v * e3v;
            double err = e5m * e5m + e5t * e5t + e5v * e5v;
            double deno = err + 0.01 * err2;
            if (deno <= 0.0) deno = 1.0;
            err = fabs(h) * err * sqrt(1.0 / (3.0 * deno));

            if (stepping) {
                out.nstep++;
                out.nrhs += 11;
                if (err <= 1.0) {
                    // accepted: FSAL evaluation f(xph, ynew) becomes the next k1
                    rsf_rhs(M, cc, vl12, muN, thN, k1m, k1t, k1v);
                    out.nrhs++;
                    mu = muN; th = thN; V = VN; t = xph;
                    const bool need_hnew = !last || M.integ_mode == RSFM_INTEG_CARRY;
                    double hnew = h;
                    if (need_hnew) {
                        const double fac = fmax(facc2, fmin(facc1, root8(err) / safe));
                        hnew = h / fac;
                        if (fabs(hnew) > hmax) hnew = hmax;
                        if (reject) hnew = fmin(fabs(hnew), fabs(h));
                    }
                    reject = false;
                    if (last) { done = true; h_carry = hnew; }
                    h = hnew;
                } else {
                    // rejected (also err = NaN).  SciPy 1.18.1's dop853 shrinks by 1/facc1 here.
                    h = h / facc1;
                    reject = true;
                    last = false;
                }
            }
        }

        // ---- output point k: RateStateModel.py:384-388, MCMC.py:387 ----
        double accv = 0.0;
        if (running) {
            accv = (V - vprev) / M.delta_t;
            vprev = V;
            if (failed) out.filled = k + 1;
        }
        if (have_data) { const double e = accv - dk; sse += e * e; }
        if (active) {
            if (acc_out) acc_out[(size_t)k * acc_stride] = accv;
            if (t_out) t_out[(size_t)k * acc_stride] = running ? t : 0.0;
            if (acc_ref) { const double x = (accv - acc_ref[(size_t)k * acc_stride]) / fd_den; xtx += x * x; }
        }
    }
    out.sse = sse;
    if (xtx_out) *xtx_out = xtx;
    return out;
}

}  // namespace rsfm
