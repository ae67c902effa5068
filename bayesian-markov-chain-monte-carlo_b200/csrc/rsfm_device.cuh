// rsfm_device.cuh -- device-side RSF spring-slider solver for sm_100a.
//
// One thread integrates one chain's 3-state ODE (mu, theta, V) in FP64 with
// DOP853 and per-chain adaptive steps, following the step logic of SciPy's
// ode('dop853') as the reference drives it (RateStateModel.py:374-389, SURVEY.md
// Appendix B): a fresh call per output interval, hinit, err/accept, controller.
//
// B200-specific structure (not present in any form in the reference; details at each definition):
//   * V (y[2]) never feeds back into the RHS (RateStateModel.py:336-346 use only mu, theta): its stage
//     values are not stored, the three linear forms that need them are accumulated on the fly.
//   * the load-point term L(t) depends on time only: it is tabulated once per sampler on the nominal time
//     grid (loading_table_kernel), prefetched per interval with cp.async into a per-warp shared table;
//     lanes off that grid share an on-the-fly table per warp, or fill a private column (stiff regime).
//   * the fast interval: hinit's probe + the twelve stages + FSAL as one branch-free block verified by a
//     single warp vote (dop853_step_fast: short dependency chain, mu / theta series evaluated in parallel);
//     a state-by-state general path (dop853_step_impl<false>, reference formulas) covers everything else.
//   * the observed series streams through shared memory in 4 KB tiles fetched by the TMA bulk-copy engine
//     (cp.async.bulk + mbarrier), double-buffered; series of <= 1024 points stay resident.
//   * restarting k1 = f(t, y) at every interval is skipped: it is bit-identical to the FSAL evaluation that
//     closed the previous interval.
//   * sqrt- and division-free accept / hinit decisions; exact early rejection against an SSE bound.
//
// Layout of this file: TMA staging of the series; the RHS (rsf_rhs) and the two DOP853 step forms (dop853_step_impl:
// general range; dop853_step_fast: short dependency chain); the general interval
// path in its two variants -- rsf_interval_plain (SciPy's controller as written; inline in the 168-register kernels,
// out of line in the speculative one) and rsf_interval_general (stiff variant for velocity-step loading: re-based
// friction law, exploding trial steps not scored, SFU-seeded controller root) -- and rsf_solve_mode, the output loop
// with the fast interval.  Kernels pick the variant with the template flags VS / ISO (rsfm_kernels.cu).
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
    double vstep_lnf, vstep_rfac;      // log(vstep_factor), 1/vstep_factor: host-computed (re-basing, see rsf_interval_general)
    double vstep_rper;                 // 1/vstep_period (vstep_index)
    int stiff_exact;                   // stiff variant: score every step that left the fast ranges (tuning / tests)
    int observable;                    // RSFM_OBS_ACC (reference) / RSFM_OBS_MU
    int n_out, nmax, damping, loading, integ_mode;
    int state_law;                     // RSFM_LAW_AGING (reference, :340) / RSFM_LAW_SLIP
    int load_n;                        // RSFM_LOAD_TABLE: entries of load_tab
    double load_dt;                    //   spacing of the table (first entry at t_start)
    const double *load_tab;            //   device pointer, V_l/V_ref - 1
    double dc_fixed;                   // RSFM_PARAM_K1: the per-chain scalar is k1 and Dc is this constant
};

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

// 8-byte asynchronous global->shared copy (LDGSTS), used to prefetch the next interval's stage values
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src_global)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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
        if (resident) {
            // A resident series (n <= 1024: one or two tiles, fetched once per kernel) is waited for here, by every
            // thread, before the first pass: a warp may leave the output loop of any pass early (all its lanes
            // rejected / inactive), so no later point of the loop is reached by every warp of the block.  Each
            // barrier completes exactly one phase in the kernel's lifetime.
            while (!mbar_try_wait(&bar[0], 0u)) { }
            if (n > SERIES_TILE) { while (!mbar_try_wait(&bar[1], 0u)) { } }
            loaded = true;
        }
    }

    // value data[k]; in streaming mode entering a tile recycles the buffer two tiles back
    __device__ __forceinline__ double at(int k)
    {
        if (resident && loaded) return buf[k];             // both tiles are in place and contiguous
        return at_staging(k);
    }

    // streamed series (n > 1024): out of line, so that the output loop of the solver stays compact (every thread
    // of the block calls it with the same k: it contains block barriers)
    __device__ __noinline__ double at_staging(int k)
    {
        const int tile = k / SERIES_TILE;
        const int off = k - tile * SERIES_TILE;
        if (off == 0) {
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
        }
        return buf[(tile & 1) * SERIES_TILE + off];
    }
};

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
// Butcher tableau and series coefficients live in the constant bank (loaded into uniform registers
// by the compiler; FP64 instructions on sm_100 take register or uniform-register operands).
struct Dop853Tab {
    double c[12];                                   // abscissae of stages 2..12 (+ pad)
    double a21, a31, a32, a41, a43, a51, a53, a54, a61, a64, a65, a71, a74, a75, a76;
    double a81, a84, a85, a86, a87, a91, a94, a95, a96, a97, a98;
    double a101, a104, a105, a106, a107, a108, a109;
    double a111, a114, a115, a116, a117, a118, a119, a1110;
    double a121, a124, a125, a126, a127, a128, a129, a1210, a1211;
    double b1, b6, b7, b8, b9, b10, b11, b12;
    double e1, e6, e7, e8, e9, e10, e11, e12;
    double bhh1, bhh2, bhh3;
    double lp[6];      // log1p series (Horner, highest first): (f + lp0) lp1, then + lp2.. ; see rsf_rhs
    double ex[6];      // expm1 series
};
__constant__ Dop853Tab TB = {
    {DP_C2, DP_C3, DP_C4, DP_C5, DP_C6, DP_C7, DP_C8, DP_C9, DP_C10, DP_C11, 1.0, 0.0},
    DP_A2_1, DP_A3_1, DP_A3_2, DP_A4_1, DP_A4_3, DP_A5_1, DP_A5_3, DP_A5_4, DP_A6_1, DP_A6_4, DP_A6_5,
    DP_A7_1, DP_A7_4, DP_A7_5, DP_A7_6,
    DP_A8_1, DP_A8_4, DP_A8_5, DP_A8_6, DP_A8_7, DP_A9_1, DP_A9_4, DP_A9_5, DP_A9_6, DP_A9_7, DP_A9_8,
    DP_A10_1, DP_A10_4, DP_A10_5, DP_A10_6, DP_A10_7, DP_A10_8, DP_A10_9,
    DP_A11_1, DP_A11_4, DP_A11_5, DP_A11_6, DP_A11_7, DP_A11_8, DP_A11_9, DP_A11_10,
    DP_A12_1, DP_A12_4, DP_A12_5, DP_A12_6, DP_A12_7, DP_A12_8, DP_A12_9, DP_A12_10, DP_A12_11,
    DP_B1, DP_B6, DP_B7, DP_B8, DP_B9, DP_B10, DP_B11, DP_B12,
    DP_ER1, DP_ER6, DP_ER7, DP_ER8, DP_ER9, DP_ER10, DP_ER11, DP_ER12,
    DP_BHH1, DP_BHH2, DP_BHH3,
    {-6.0 / 5.0, -1.0 / 6.0, -0.25, 0.3333333333333333, -0.5, 0.0},
    {7.0, 1.0 / 5040.0, 0.008333333333333333, 0.041666666666666664, 0.16666666666666666, 0.5},
};

// Relative load-point perturbation L(t) with V_l = V_ref (1 + L):
// SINE_DECAY  L = exp(-t/20) sin(10 t)            RateStateModel.py:327-329
// VSTEP       L = factor - 1 on odd periods, 0 on even periods (extension)
__device__ __noinline__ double loading_rel(int loading, double t_start, double period, double factor, double t)
{
    if (loading == RSFM_LOAD_VSTEP) {
        const double ph = floor((t - t_start) / period);
        const long long i = (long long)ph;
        return (i & 1) ? factor - 1.0 : 0.0;
    }
    return exp(-t / 20.0) * sin(10.0 * t);
}

// Period index floor((t - t_start)/period) of the VSTEP load, as the correctly rounded division gives it.
// The product with the host-computed reciprocal differs from the rounded quotient by a few ulp at most, so its
// floor is the same unless the quotient lies within 1e-6 of an integer (indices stay far below 1e9); only then is the division done
// (output times that are multiples of the period: one interval end per period).
__device__ __forceinline__ double vstep_index(const ModelK &M, double t)
{
    const double x = t - M.t_start;
    const double q = x * M.vstep_rper;
    const double n = rint(q);
    if (fabs(q - n) <= 1e-6) return floor(x / M.vstep_period);
    return floor(q);
}

__device__ __forceinline__ double loading_of(const ModelK &M, double t)
{
    if (M.loading == RSFM_LOAD_VSTEP) {            // piecewise constant: cheap enough to inline
        const long long i = (long long)vstep_index(M, t);
        return (i & 1) ? M.vstep_factor - 1.0 : 0.0;
    }
    if (M.loading == RSFM_LOAD_TABLE) {
        // piecewise linear through the table, constant beyond its ends: one rounded division for the position and
        // T[i] + fr (T[i+1] - T[i]) without contraction, the same operations, in the same order, as the CPU restatement the parity tests compare with
        const double x = __ddiv_rn(t - M.t_start, M.load_dt);
        double fi = floor(x);
        fi = fmin(fmax(fi, 0.0), (double)(M.load_n - 2));
        double fr = x - fi;
        fr = fmin(fmax(fr, 0.0), 1.0);
        const int i = (int)fi;
        const double t0 = M.load_tab[i], t1 = M.load_tab[i + 1];
        return __dadd_rn(t0, __dmul_rn(fr, __dadd_rn(t1, -t0)));
    }
    return loading_rel(M.loading, M.t_start, M.vstep_period, M.vstep_factor, t);
}

struct ChainConst {
    double b, inv_a, w, kV, voa0, k1e, mu_ref, th_eq;
    int slip;                  // state law: 0 aging (reference), 1 slip -- general-range formulas only, see make_chain_const
    double cq[6], qscale;      // (1+f)^q - 1 = f (cq0 + cq1 f + ... + cq5 f^5), q = -b/a; qscale = 1 + |q|
    // w = V_ref/Dc, kV = k' V_ref with k' = 0.1/Dc (RateStateModel.py:324), voa0 = V_ref/a,
    // k1e = k1 when RadiationDamping else 0 (:349); th_eq = Dc/V_ref, the sliding steady state AND the reference's
    // start value (:367-370), see state_excess
};

// f = V_ref theta/Dc - 1 as w (theta - theta_eq): the subtraction is exact near theta_eq, so f is accurate to 2 ulp
// OF ITSELF and EXACTLY zero at the start value theta_eq = Dc/V_ref.  The reference evaluates V_ref*theta/Dc (:336,
// :340), which is exactly 1 there (V_ref = 1) and carries 1e-16 of absolute rounding anywhere else, so the two forms
// agree to the reference's own noise -- but only this one reproduces a derivative of EXACTLY zero while the slider
// rests in steady state under a constant load (velocity-step loading before the first step): there SciPy's hinit
// takes its `der12 <= 1e-15` branch (h = 1e-6, eight steps per call), and fma(w, theta, -1) = 400 fl(1/400) - 1 =
// 5e-17 would send it down the other one (h = 1e-4, five steps) and the trajectory after the jump onto another
// accepted-step sequence (5e-6 of max|acc| at Dc = 400; found by the round-2 parity gate).
__device__ __forceinline__ double state_excess(const ChainConst &c, double th) { return c.w * (th - c.th_eq); }

__device__ __forceinline__ ChainConst make_chain_const(const ModelK &M, double a, double b, double dc)
{
    ChainConst c;
    c.b = b;
    c.inv_a = 1.0 / a;
    c.w = M.V_ref / dc;
    c.th_eq = dc / M.V_ref;
    c.kV = (1e-2 * 10 / dc) * M.V_ref;
    c.voa0 = M.V_ref * c.inv_a;
    c.k1e = M.damping ? M.k1 : 0.0;
    c.mu_ref = M.mu_ref;
    const double q = -b * c.inv_a;
    c.cq[0] = q;
    c.cq[1] = c.cq[0] * (q - 1.0) / 2.0;
    c.cq[2] = c.cq[1] * (q - 2.0) / 3.0;
    c.cq[3] = c.cq[2] * (q - 3.0) / 4.0;
    c.cq[4] = c.cq[3] * (q - 4.0) / 5.0;
    c.cq[5] = c.cq[4] * (q - 5.0) / 6.0;
    c.qscale = 1.0 + fabs(q);
    // Slip law (extension): theta' = -(v theta/Dc) ln(v theta/Dc) has no short-series form here; an
    // infinite range scale makes every fast stage report "out of range" (inf * |f| < limit is false, inf * 0 = NaN
    // too), so steps are scored by the general-range stages, which carry the law (rsf_rhs<false>), without a single
    // extra instruction in the fast stage of the aging law.
    c.slip = M.state_law == RSFM_LAW_SLIP ? 1 : 0;
    if (c.slip) c.qscale = INFINITY;
    return c;
}

// friction(t, y), RateStateModel.py:336-353, in a cancellation-free form.  With x = V_ref theta/Dc =
// 1 + f, v = V_ref e^temp = V_ref (1 + E) and V_l = V_ref (1 + L):
//     theta' = 1 - v theta/Dc = -(E + f + E f)          (:340)
//     mu'    = k'(V_l - v)    = k' V_ref (L - E)         (:343)
//     V'     = (v/a)(mu' - (b/theta) theta')             (:346)
//     damping (:349-353): mu' <- mu' - k1 V' ; V' <- (v/a)(mu' - ...) = V' (1 - k1 v/a)
// The reference evaluates the same expressions with ~1e-16 absolute rounding noise in the
// differences; this form is algebraically identical and differs from it at that noise level only.
//
// FAST = true: branch-free.  log1p and expm1 are the Taylor series truncated for |f| < 2^-9 (terms
// to f^5, truncation < 1e-17 relative) and |temp| < 2^-6 (terms to t^6), each a Horner chain with one
// constant per FMA; 1/theta is refreshed from the previous evaluation by two Newton steps (theta
// moves ~1e-6 relative between stages).  The three range conditions are OR-ed into `bad`; the
// caller recomputes with FAST = false (libdevice log1p / expm1, true division) when it is set.
// Along every trajectory with Dc >= 120 of the reference's loading the fast ranges always hold.
template <bool FAST>
__device__ __forceinline__ void rsf_rhs(const ChainConst &c, double L, double mu, double th, double &rth,
                                        double &dmu, double &dth, double &dV, bool &bad)
{
    const double f = state_excess(c, th);
    double lg, E, r;
    const double dmu0 = mu - c.mu_ref;
    if (FAST) {
        double p = (f + TB.lp[0]) * TB.lp[1];                  // -1/6 f + 1/5
        p = fma(p, f, TB.lp[2]);
        p = fma(p, f, TB.lp[3]);
        p = fma(p, f, TB.lp[4]);
        lg = fma(f * f, p, f);                                  // log(V_ref theta / Dc)
        const double temp = c.inv_a * fma(-c.b, lg, dmu0);      // :336
        double q = (temp + TB.ex[0]) * TB.ex[1];                // t/5040 + 1/720
        q = fma(q, temp, TB.ex[2]);
        q = fma(q, temp, TB.ex[3]);
        q = fma(q, temp, TB.ex[4]);
        q = fma(q, temp, TB.ex[5]);
        E = fma(temp * temp, q, temp);                          // v / V_ref - 1, :337
        const double e0 = fma(-th, rth, 1.0);
        r = fma(rth, e0, rth);
        r = fma(r, fma(-th, r, 1.0), r);
        bad = bad || !(fabs(f) < 0.001953125 && fabs(temp) < 0.015625 && fabs(e0) < 6.0e-5);
    } else {
        // general range (stiff regime, large excursions): the reference's formulas as written,
        // v = V_ref exp((mu - mu_ref - b log(V_ref theta/Dc))/a), RateStateModel.py:336-346
        const double x = 1.0 + f;                                   // V_ref theta / Dc
        const double temp = c.inv_a * fma(-c.b, log(x), dmu0);
        const double ev = exp(temp);                                // v / V_ref
        r = 1.0 / th;
        dth = fma(-ev, x, 1.0);                                      // (one rounding, as the compiler contracted it before
                                                                     //  the slip law shared the product below)
        if (c.slip) { const double z = __dmul_rn(ev, x); dth = -z * log(z); }      // Ruina slip law
        const double voa_g = c.voa0 * ev;
        const double d0_g = c.kV * ((L + 1.0) - ev);
        const double s_g = (c.b * r) * dth;
        const double v0_g = voa_g * (d0_g - s_g);
        dmu = fma(-c.k1e, v0_g, d0_g);
        dV = fma(-(voa_g * c.k1e), v0_g, v0_g);
        rth = r;
        return;
    }
    dth = -(fma(E, f, E) + f);
    const double voa = fma(c.voa0, E, c.voa0);                  // v / a
    const double d0 = c.kV * (L - E);
    const double s = (c.b * r) * dth;
    const double v0 = voa * (d0 - s);
    dmu = fma(-c.k1e, v0, d0);
    dV = fma(-(voa * c.k1e), v0, v0);
    rth = r;
}

__device__ __forceinline__ double root8(double x) { return sqrt(sqrt(sqrt(x))); }

// (q/den)^(-1/16) for positive, normal, finite q and den, to ~5e-12 relative: exponents by integer arithmetic,
// log2 / exp2 of the mantissas on the SFU in single precision (z0, ~2^-21), then one Newton step in double,
// z = z0 (1 + (1 - (q/den) z0^16)/16), whose small residual only needs the approximate reciprocal of den.
// Replaces sqrt + division + three sqrt + two divisions (~1,000 dependent cycles) in the step-size controller
// of the stiff variant by ~200.
__device__ __forceinline__ double inv_root16(double q, double den)
{
    const int hq = __double2hiint(q), hd = __double2hiint(den);
    const int eq = (hq >> 20) - 1023, ed = (hd >> 20) - 1023;
    const float mq = (float)__hiloint2double((hq & 0x000fffff) | 0x3ff00000, __double2loint(q));
    const float md = (float)__hiloint2double((hd & 0x000fffff) | 0x3ff00000, __double2loint(den));
    const float l = (float)(eq - ed) + (__log2f(mq) - __log2f(md));
    float z0f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(z0f) : "f"(-0.0625f * l));
    const double z0 = (double)z0f;
    const double z2 = z0 * z0, z4 = z2 * z2, z8 = z4 * z4, z16 = z8 * z8;
    const double r = fma(-q, z16, den);
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(den));
    return fma(z0, (r * rc) * 0.0625, z0);
}

// One attempted DOP853 step (stages 2..12, 8th-order solution, error forms).  Lp points at the
// eleven stage values of L (stride ls doubles, shared memory).
struct StepIn { double h, mu, th, V, k1m, k1t, k1v, rth, atol, rtol; };
// err = |h| ||e5||^2 / sqrt(3 (||e5||^2 + 0.01 ||e3||^2)) = |h| errA / sqrt(den3)   (dop853.f, n = 3)
struct StepOut { double muN, thN, VN, errA, den3, L12, rth; };

// General-range step: the reference's formulas (libdevice log / exp, true division) at every stage.
// Lc, rl: the stage values of L are re-based on the fly, L' = (L - Lc) rl (see rsf_interval_general; 0, 1 = none).
__device__ __forceinline__ void dop853_step_impl(const ChainConst &cc, const StepIn &I, const double *Lp, int ls,
                                                 double Lc, double rl, StepOut &O)
{
    bool bad = false;
    const double h = I.h, mu = I.mu, th = I.th, k1m = I.k1m, k1t = I.k1t, k1v = I.k1v;
    double rth = I.rth;
    double k2m, k2t, k3m, k3t, k4m, k4t, k5m, k5t, k6m, k6t, k7m, k7t, k8m, k8t, k9m, k9t, k10m, k10t;
    double kv, k9v, k12v, bV, eV;
    // stages 2..5: their V-derivatives carry zero weight everywhere
    rsf_rhs<false>(cc, (Lp[0 * ls] - Lc) * rl, mu + h * (TB.a21 * k1m), th + h * (TB.a21 * k1t), rth, k2m, k2t, kv, bad);
    rsf_rhs<false>(cc, (Lp[1 * ls] - Lc) * rl, mu + h * (TB.a31 * k1m + TB.a32 * k2m), th + h * (TB.a31 * k1t + TB.a32 * k2t),
                  rth, k3m, k3t, kv, bad);
    rsf_rhs<false>(cc, (Lp[2 * ls] - Lc) * rl, mu + h * (TB.a41 * k1m + TB.a43 * k3m), th + h * (TB.a41 * k1t + TB.a43 * k3t),
                  rth, k4m, k4t, kv, bad);
    rsf_rhs<false>(cc, (Lp[3 * ls] - Lc) * rl, mu + h * (TB.a51 * k1m + TB.a53 * k3m + TB.a54 * k4m),
                  th + h * (TB.a51 * k1t + TB.a53 * k3t + TB.a54 * k4t), rth, k5m, k5t, kv, bad);
    rsf_rhs<false>(cc, (Lp[4 * ls] - Lc) * rl, mu + h * (TB.a61 * k1m + TB.a64 * k4m + TB.a65 * k5m),
                  th + h * (TB.a61 * k1t + TB.a64 * k4t + TB.a65 * k5t), rth, k6m, k6t, kv, bad);
    bV = TB.b1 * k1v + TB.b6 * kv;
    eV = TB.e1 * k1v + TB.e6 * kv;
    rsf_rhs<false>(cc, (Lp[5 * ls] - Lc) * rl, mu + h * (TB.a71 * k1m + TB.a74 * k4m + TB.a75 * k5m + TB.a76 * k6m),
                  th + h * (TB.a71 * k1t + TB.a74 * k4t + TB.a75 * k5t + TB.a76 * k6t), rth, k7m, k7t, kv, bad);
    bV += TB.b7 * kv; eV += TB.e7 * kv;
    rsf_rhs<false>(cc, (Lp[6 * ls] - Lc) * rl,
                  mu + h * (TB.a81 * k1m + TB.a84 * k4m + TB.a85 * k5m + TB.a86 * k6m + TB.a87 * k7m),
                  th + h * (TB.a81 * k1t + TB.a84 * k4t + TB.a85 * k5t + TB.a86 * k6t + TB.a87 * k7t), rth,
                  k8m, k8t, kv, bad);
    bV += TB.b8 * kv; eV += TB.e8 * kv;
    rsf_rhs<false>(cc, (Lp[7 * ls] - Lc) * rl,
                  mu + h * (TB.a91 * k1m + TB.a94 * k4m + TB.a95 * k5m + TB.a96 * k6m + TB.a97 * k7m + TB.a98 * k8m),
                  th + h * (TB.a91 * k1t + TB.a94 * k4t + TB.a95 * k5t + TB.a96 * k6t + TB.a97 * k7t + TB.a98 * k8t),
                  rth, k9m, k9t, k9v, bad);
    bV += TB.b9 * k9v; eV += TB.e9 * k9v;
    rsf_rhs<false>(cc, (Lp[8 * ls] - Lc) * rl,
                  mu + h * (TB.a101 * k1m + TB.a104 * k4m + TB.a105 * k5m + TB.a106 * k6m + TB.a107 * k7m +
                            TB.a108 * k8m + TB.a109 * k9m),
                  th + h * (TB.a101 * k1t + TB.a104 * k4t + TB.a105 * k5t + TB.a106 * k6t + TB.a107 * k7t +
                            TB.a108 * k8t + TB.a109 * k9t),
                  rth, k10m, k10t, kv, bad);
    bV += TB.b10 * kv; eV += TB.e10 * kv;
    // stage 11 -> k2 slot
    rsf_rhs<false>(cc, (Lp[9 * ls] - Lc) * rl,
                  mu + h * (TB.a111 * k1m + TB.a114 * k4m + TB.a115 * k5m + TB.a116 * k6m + TB.a117 * k7m +
                            TB.a118 * k8m + TB.a119 * k9m + TB.a1110 * k10m),
                  th + h * (TB.a111 * k1t + TB.a114 * k4t + TB.a115 * k5t + TB.a116 * k6t + TB.a117 * k7t +
                            TB.a118 * k8t + TB.a119 * k9t + TB.a1110 * k10t),
                  rth, k2m, k2t, kv, bad);
    bV += TB.b11 * kv; eV += TB.e11 * kv;
    // stage 12 -> k3 slot, at x + h
    const double L12 = Lp[10 * ls];                 // returned as stored (the caller re-bases it for the FSAL evaluation)
    rsf_rhs<false>(cc, (L12 - Lc) * rl,
                  mu + h * (TB.a121 * k1m + TB.a124 * k4m + TB.a125 * k5m + TB.a126 * k6m + TB.a127 * k7m +
                            TB.a128 * k8m + TB.a129 * k9m + TB.a1210 * k10m + TB.a1211 * k2m),
                  th + h * (TB.a121 * k1t + TB.a124 * k4t + TB.a125 * k5t + TB.a126 * k6t + TB.a127 * k7t +
                            TB.a128 * k8t + TB.a129 * k9t + TB.a1210 * k10t + TB.a1211 * k2t),
                  rth, k3m, k3t, k12v, bad);
    bV += TB.b12 * k12v; eV += TB.e12 * k12v;
    // 8th-order solution, 5th/3rd-order error forms
    const double bM = TB.b1 * k1m + TB.b6 * k6m + TB.b7 * k7m + TB.b8 * k8m + TB.b9 * k9m + TB.b10 * k10m +
                      TB.b11 * k2m + TB.b12 * k3m;
    const double bT = TB.b1 * k1t + TB.b6 * k6t + TB.b7 * k7t + TB.b8 * k8t + TB.b9 * k9t + TB.b10 * k10t +
                      TB.b11 * k2t + TB.b12 * k3t;
    const double muN = mu + h * bM, thN = th + h * bT, VN = I.V + h * bV;
    // error norms over the common denominator D = (sk0 sk1 sk2)^2 (no divisions):
    //   ||e5/sk||^2 = A/D, ||e3/sk||^2 = B/D  with  A = sum (e5_i p_i)^2,  p_i = prod_{j != i} sk_j
    const double k0 = I.atol + I.rtol * fmax(fabs(mu), fabs(muN));
    const double k1 = I.atol + I.rtol * fmax(fabs(th), fabs(thN));
    const double k2 = I.atol + I.rtol * fmax(fabs(I.V), fabs(VN));
    const double p0 = k1 * k2, p1 = k0 * k2, p2 = k0 * k1;
    const double dd = k0 * p0;
    const double e3m = (bM - TB.bhh1 * k1m - TB.bhh2 * k9m - TB.bhh3 * k3m) * p0;
    const double e3t = (bT - TB.bhh1 * k1t - TB.bhh2 * k9t - TB.bhh3 * k3t) * p1;
    const double e3v = (bV - TB.bhh1 * k1v - TB.bhh2 * k9v - TB.bhh3 * k12v) * p2;
    const double e5m = (TB.e1 * k1m + TB.e6 * k6m + TB.e7 * k7m + TB.e8 * k8m + TB.e9 * k9m + TB.e10 * k10m +
                        TB.e11 * k2m + TB.e12 * k3m) * p0;
    const double e5t = (TB.e1 * k1t + TB.e6 * k6t + TB.e7 * k7t + TB.e8 * k8t + TB.e9 * k9t + TB.e10 * k10t +
                        TB.e11 * k2t + TB.e12 * k3t) * p1;
    const double e5v = eV * p2;
    const double B = e3m * e3m + e3t * e3t + e3v * e3v;
    const double A = e5m * e5m + e5t * e5t + e5v * e5v;
    O.muN = muN; O.thN = thN; O.VN = VN; O.errA = A; O.den3 = 3.0 * (A + 0.01 * B) * (dd * dd); O.L12 = L12;
    O.rth = rth;
}

// ---------------------------------------------------------------------------
// Fast step: the same DOP853 step as dop853_step_impl, organised for a short dependency chain.
// ---------------------------------------------------------------------------
// One warp per SM sub-partition is latency-bound: a stage costs (dependent FP64 ops) x 8 cycles.  The
// textbook order  theta_s -> f -> log1p -> temp -> expm1 -> theta'  is a 20-op chain.  Here
//     v/V_ref = e^temp = e^A (1+f)^q,   A = (mu_s - mu_ref)/a,   q = -b/a,
// so with u = expm1(A) (a series in A, the mu path) and g = (1+f)^q - 1 (a binomial series in f, the
// theta path), E = v/V_ref - 1 = u + g + u g: the two series are independent and overlap.  Stage
// arguments are formed in the scaled variables directly,
//     f_s = f_0 + h w  sum_j a_sj theta'_j ,   A_s = A_0 + (h/a) sum_j a_sj mu'_j ,
// with everything except the newest stage's term pre-accumulated while that stage is still being
// evaluated, and the newest mu' split into its early part d0 = k'V_ref (L - E) and the radiation-
// damping correction c = k1 V' (RateStateModel.py:351), which arrives later and enters through
// e^(A + delta) - 1 = u + delta (1 + u) (exact to O(delta^2) ~ 1e-20).  Chain per stage: 13 ops.
// Values agree with the textbook order to rounding; both are the same function of (mu_s, theta_s).
struct StageOut { double km, kt, kv, d0, c; };

// STIFF: the stiff variant's form (third Newton step for 1/theta).  The range test stays one expression OR-ed into
// `bad`: ptxas schedules the 1,000-instruction fast-interval block ~25 % longer when a per-stage flag is materialised.
template <bool STIFF>
__device__ __forceinline__ void rsf_stage_fast(const ChainConst &cc, double fs, double As, double dl, double th1,
                                               double kVL, double &rth, StageOut &o, bool &bad)
{
    double p = fma(cc.cq[5], fs, cc.cq[4]);
    p = fma(p, fs, cc.cq[3]);
    p = fma(p, fs, cc.cq[2]);
    p = fma(p, fs, cc.cq[1]);
    p = fma(p, fs, cc.cq[0]);
    const double g = p * fs;                                    // (1+f)^q - 1
    double qq = (As + TB.ex[0]) * TB.ex[1];
    qq = fma(qq, As, TB.ex[2]);
    qq = fma(qq, As, TB.ex[3]);
    qq = fma(qq, As, TB.ex[4]);
    qq = fma(qq, As, TB.ex[5]);
    const double u = fma(As * As, qq, As);                      // expm1(A)
    const double E0 = fma(u, g, u) + g;
    const double E = fma(dl, E0, E0) + dl;                      // v / V_ref - 1
    o.kt = -(fma(E, fs, E) + fs);                               // theta'
    o.d0 = fma(-cc.kV, E, kVL);                                 // k' V_ref (L - E)
    const double voa = fma(cc.voa0, E, cc.voa0);                // v / a
    const double e0 = fma(-th1, rth, 1.0);
    double r = fma(rth, e0, rth);
    r = fma(r, fma(-th1, r, 1.0), r);
    // stiff variant: a third Newton step (error e0^8) admits |e0| < 1e-2.  Near the stability limit theta moves
    // by more than 6e-5 between stages in 40 % of the trial steps that leave the ranges, and in nothing else.
    if (STIFF) r = fma(r, fma(-th1, r, 1.0), r);
    rth = r;
    const double sb = (cc.b * r) * o.kt;
    const double v0 = voa * (o.d0 - sb);
    o.c = cc.k1e * v0;
    o.km = o.d0 - o.c;                                          // mu' with radiation damping
    o.kv = fma(-(voa * cc.k1e), v0, v0);                        // V'
    bad = bad || !(fabs(fs) * cc.qscale < 0.001953125 && fabs(As) < 0.015625 && fabs(e0) < (STIFF ? 1.0e-2 : 6.0e-5));
}

// RB = true: re-based frame (see Rebase): cc holds the re-based constants and the stage values of L are taken
// relative to the base level, k'V_ref' L' = kVb (L - Lc) with kVb the chain's own k'V_ref.
template <bool RB>
__device__ __forceinline__ void dop853_step_fast(const ChainConst &cc, const StepIn &I, const double *Lp, int ls,
                                                 double kVb, double Lc, StepOut &O, bool &bad)
{
    const double h = I.h, mu = I.mu, th = I.th, k1m = I.k1m, k1t = I.k1t, k1v = I.k1v;
    double rth = I.rth;
    const double hw = h * cc.w, ih = h * cc.inv_a;
    const double f0 = state_excess(cc, th), A0 = (mu - cc.mu_ref) * cc.inv_a;
    double k2m, k2t, k3m, k3t, k4m, k4t, k5m, k5t, k6m, k6t, k7m, k7t, k8m, k8t, k9m, k9t, k10m, k10t;
    double k9v, k12v, bV, eV;
    StageOut so;
    // PT / PM: sums over the already-known stages; (NEW) the coefficient of the stage still in flight
#define RSFM_STAGE(IDX, PT, PM, ANEW)                                                                     \
    {                                                                                                      \
        const double pt__ = (PT), pm__ = (PM);                                                             \
        const double fs__ = fma(hw * (ANEW), so.kt, fma(hw, pt__, f0));                                    \
        const double th__ = fma(h * (ANEW), so.kt, fma(h, pt__, th));                                      \
        const double As__ = fma(ih * (ANEW), so.d0, fma(ih, pm__, A0));                                    \
        const double dl__ = -(ih * (ANEW)) * so.c;                                                         \
        rsf_stage_fast<RB>(cc, fs__, As__, dl__, th__,                                                     \
                           RB ? kVb * (Lp[(IDX) * ls] - Lc) : cc.kV * Lp[(IDX) * ls], rth, so, bad);       \
    }
    // stage 2: the "stage in flight" is k1 itself (complete: no damping split)
    so.kt = k1t; so.d0 = k1m; so.c = 0.0;
    RSFM_STAGE(0, 0.0, 0.0, TB.a21);
    k2m = so.km; k2t = so.kt;
    RSFM_STAGE(1, TB.a31 * k1t, TB.a31 * k1m, TB.a32);
    k3m = so.km; k3t = so.kt;
    RSFM_STAGE(2, TB.a41 * k1t, TB.a41 * k1m, TB.a43);
    k4m = so.km; k4t = so.kt;
    RSFM_STAGE(3, TB.a51 * k1t + TB.a53 * k3t, TB.a51 * k1m + TB.a53 * k3m, TB.a54);
    k5m = so.km; k5t = so.kt;
    RSFM_STAGE(4, TB.a61 * k1t + TB.a64 * k4t, TB.a61 * k1m + TB.a64 * k4m, TB.a65);
    k6m = so.km; k6t = so.kt;
    bV = TB.b1 * k1v + TB.b6 * so.kv;
    eV = TB.e1 * k1v + TB.e6 * so.kv;
    RSFM_STAGE(5, TB.a71 * k1t + TB.a74 * k4t + TB.a75 * k5t, TB.a71 * k1m + TB.a74 * k4m + TB.a75 * k5m, TB.a76);
    k7m = so.km; k7t = so.kt;
    bV += TB.b7 * so.kv; eV += TB.e7 * so.kv;
    RSFM_STAGE(6, TB.a81 * k1t + TB.a84 * k4t + TB.a85 * k5t + TB.a86 * k6t,
               TB.a81 * k1m + TB.a84 * k4m + TB.a85 * k5m + TB.a86 * k6m, TB.a87);
    k8m = so.km; k8t = so.kt;
    bV += TB.b8 * so.kv; eV += TB.e8 * so.kv;
    RSFM_STAGE(7, TB.a91 * k1t + TB.a94 * k4t + TB.a95 * k5t + TB.a96 * k6t + TB.a97 * k7t,
               TB.a91 * k1m + TB.a94 * k4m + TB.a95 * k5m + TB.a96 * k6m + TB.a97 * k7m, TB.a98);
    k9m = so.km; k9t = so.kt; k9v = so.kv;
    bV += TB.b9 * k9v; eV += TB.e9 * k9v;
    RSFM_STAGE(8, TB.a101 * k1t + TB.a104 * k4t + TB.a105 * k5t + TB.a106 * k6t + TB.a107 * k7t + TB.a108 * k8t,
               TB.a101 * k1m + TB.a104 * k4m + TB.a105 * k5m + TB.a106 * k6m + TB.a107 * k7m + TB.a108 * k8m, TB.a109);
    k10m = so.km; k10t = so.kt;
    bV += TB.b10 * so.kv; eV += TB.e10 * so.kv;
    // stage 11 -> k2 slot
    RSFM_STAGE(9,
               TB.a111 * k1t + TB.a114 * k4t + TB.a115 * k5t + TB.a116 * k6t + TB.a117 * k7t + TB.a118 * k8t + TB.a119 * k9t,
               TB.a111 * k1m + TB.a114 * k4m + TB.a115 * k5m + TB.a116 * k6m + TB.a117 * k7m + TB.a118 * k8m + TB.a119 * k9m,
               TB.a1110);
    k2m = so.km; k2t = so.kt;
    bV += TB.b11 * so.kv; eV += TB.e11 * so.kv;
    // stage 12 -> k3 slot, at x + h
    const double L12 = Lp[10 * ls];
    RSFM_STAGE(10,
               TB.a121 * k1t + TB.a124 * k4t + TB.a125 * k5t + TB.a126 * k6t + TB.a127 * k7t + TB.a128 * k8t + TB.a129 * k9t +
                   TB.a1210 * k10t,
               TB.a121 * k1m + TB.a124 * k4m + TB.a125 * k5m + TB.a126 * k6m + TB.a127 * k7m + TB.a128 * k8m + TB.a129 * k9m +
                   TB.a1210 * k10m,
               TB.a1211);
    k3m = so.km; k3t = so.kt; k12v = so.kv;
    bV += TB.b12 * k12v; eV += TB.e12 * k12v;
#undef RSFM_STAGE
    const double bM = TB.b1 * k1m + TB.b6 * k6m + TB.b7 * k7m + TB.b8 * k8m + TB.b9 * k9m + TB.b10 * k10m +
                      TB.b11 * k2m + TB.b12 * k3m;
    const double bT = TB.b1 * k1t + TB.b6 * k6t + TB.b7 * k7t + TB.b8 * k8t + TB.b9 * k9t + TB.b10 * k10t +
                      TB.b11 * k2t + TB.b12 * k3t;
    const double muN = mu + h * bM, thN = th + h * bT, VN = I.V + h * bV;
    const double k0 = I.atol + I.rtol * fmax(fabs(mu), fabs(muN));
    const double k1 = I.atol + I.rtol * fmax(fabs(th), fabs(thN));
    const double k2 = I.atol + I.rtol * fmax(fabs(I.V), fabs(VN));
    const double p0 = k1 * k2, p1 = k0 * k2, p2 = k0 * k1;
    const double dd = k0 * p0;
    const double e3m = (bM - TB.bhh1 * k1m - TB.bhh2 * k9m - TB.bhh3 * k3m) * p0;
    const double e3t = (bT - TB.bhh1 * k1t - TB.bhh2 * k9t - TB.bhh3 * k3t) * p1;
    const double e3v = (bV - TB.bhh1 * k1v - TB.bhh2 * k9v - TB.bhh3 * k12v) * p2;
    const double e5m = (TB.e1 * k1m + TB.e6 * k6m + TB.e7 * k7m + TB.e8 * k8m + TB.e9 * k9m + TB.e10 * k10m +
                        TB.e11 * k2m + TB.e12 * k3m) * p0;
    const double e5t = (TB.e1 * k1t + TB.e6 * k6t + TB.e7 * k7t + TB.e8 * k8t + TB.e9 * k9t + TB.e10 * k10t +
                        TB.e11 * k2t + TB.e12 * k3t) * p1;
    const double e5v = eV * p2;
    const double B = e3m * e3m + e3t * e3t + e3v * e3v;
    const double A = e5m * e5m + e5t * e5t + e5v * e5v;
    O.muN = muN; O.thN = thN; O.VN = VN; O.errA = A; O.den3 = 3.0 * (A + 0.01 * B) * (dd * dd); O.L12 = L12;
    O.rth = rth;
}

// Re-based copy of the chain constants: reference velocity lam V_ref, mu_ref' = mu_ref_r (see rsf_solve_mode).
// lam = 1 reproduces the constants bit for bit.
__device__ __forceinline__ ChainConst rebase_const(const ChainConst &c, double lam, double rl, double mu_ref_r)
{
    ChainConst r = c;
    r.w = c.w * lam; r.kV = c.kV * lam; r.voa0 = c.voa0 * lam; r.mu_ref = mu_ref_r; r.th_eq = c.th_eq * rl;   // rl = 1/lam
    return r;
}

// out-of-line general-range step (one copy; taken only when a fast range condition failed).  The base
// constants are passed by address and re-based inside, so that the caller keeps no second copy in memory.
__device__ __noinline__ void dop853_step_general(const ChainConst *cc, double lam, double mu_ref_r, const StepIn *I,
                                                 const double *Lp, int ls, double Lc, double rl, StepOut *O)
{
    const ChainConst cr = rebase_const(*cc, lam, rl, mu_ref_r);
    dop853_step_impl(cr, *I, Lp, ls, Lc, rl, *O);
}

// fast step in a re-based frame (rsf_interval_general); returns whether a stage left the fast ranges
__device__ __forceinline__ bool dop853_step_fast_rb(const ChainConst *cc, double lam, double rl, double mu_ref_r,
                                                    const StepIn *I, const double *Lp, int ls, double Lc, StepOut *O)
{
    const ChainConst cr = rebase_const(*cc, lam, rl, mu_ref_r);
    bool bad = false;
    dop853_step_fast<true>(cr, *I, Lp, ls, cc->kV, Lc, *O, bad);
    return bad;
}

__device__ __noinline__ void rsf_rhs_general(const ChainConst *c, double lam, double rl, double mu_ref_r, double L,
                                             double mu, double th, double *res)
{
    bool unused = false;
    double rth = 0.0;
    const ChainConst cr = rebase_const(*c, lam, rl, mu_ref_r);
    rsf_rhs<false>(cr, L, mu, th, rth, res[0], res[1], res[2], unused);
    res[3] = rth;
}

// single evaluation with fallback (start value, hinit probe, FSAL).  c = the (possibly re-based) constants
// the fast evaluation uses; c0, lam, mu_ref_r = what they were re-based from; L is relative to c's frame.
__device__ __forceinline__ void rsf_rhs_checked(const ChainConst &c, const ChainConst &c0, double lam, double rl,
                                                double mu_ref_r, double L, double mu, double th, double &rth, double &dmu,
                                                double &dth, double &dV)
{
    bool bad = false;
    rsf_rhs<true>(c, L, mu, th, rth, dmu, dth, dV, bad);
    if (bad || c.slip) {
        double res[4];
        rsf_rhs_general(&c0, lam, rl, mu_ref_r, L, mu, th, res);
        dmu = res[0]; dth = res[1]; dV = res[2]; rth = res[3];
    }
}

// Plain (un-re-based) forms with the short signatures the non-VSTEP solver uses.
__device__ __noinline__ void dop853_step_general(const ChainConst *cc, const StepIn *I, const double *Lp, int ls,
                                                 StepOut *O)
{
    dop853_step_impl(*cc, *I, Lp, ls, 0.0, 1.0, *O);
}

__device__ __noinline__ void rsf_rhs_general(const ChainConst *c, double L, double mu, double th, double *res)
{
    bool unused = false;
    double rth = 0.0;
    rsf_rhs<false>(*c, L, mu, th, rth, res[0], res[1], res[2], unused);
    res[3] = rth;
}

__device__ __forceinline__ void rsf_rhs_checked(const ChainConst &c, double L, double mu, double th, double &rth,
                                                double &dmu, double &dth, double &dV)
{
    bool bad = false;
    rsf_rhs<true>(c, L, mu, th, rth, dmu, dth, dV, bad);
    if (bad || c.slip) {
        double res[4];
        rsf_rhs_general(&c, L, mu, th, res);
        dmu = res[0]; dth = res[1]; dV = res[2]; rth = res[3];
    }
}

__device__ __forceinline__ void dop853_step_fast(const ChainConst &cc, const StepIn &I, const double *Lp, int ls,
                                                 StepOut &O, bool &bad)
{
    dop853_step_fast<false>(cc, I, Lp, ls, cc.kV, 0.0, O, bad);
}

#ifdef RSFM_DEBUG_COUNT
// tuning builds only (profiles/microbench/stiff_paths.py): where the step loop spends its trips
__device__ unsigned long long g_dbg[16];
__device__ double g_dbgrec[64 * 8];
__device__ unsigned int g_dbgrec_n;
#define RSFM_DBGREC(pred, a0, a1, a2, a3, a4, a5, a6, a7)                                  \
    if (pred) {                                                                            \
        const unsigned int i__ = atomicAdd(&g_dbgrec_n, 1u);                               \
        if (i__ < 64) { double *r__ = g_dbgrec + 8 * i__; r__[0] = a0; r__[1] = a1; r__[2] = a2; r__[3] = a3; \
                        r__[4] = a4; r__[5] = a5; r__[6] = a6; r__[7] = a7; }              \
    }
#define RSFM_DBG(i, pred)                                                                  \
    {                                                                                      \
        const unsigned m__ = __ballot_sync(FULL_MASK, (pred));                             \
        if ((threadIdx.x & 31) == 0 && m__) atomicAdd(&g_dbg[i], (unsigned long long)__popc(m__)); \
    }
#define RSFM_DBGW(i, pred)                                                                 \
    {                                                                                      \
        const unsigned m__ = __ballot_sync(FULL_MASK, (pred));                             \
        if ((threadIdx.x & 31) == 0 && m__) atomicAdd(&g_dbg[i], 1ull);                    \
    }
#else
#define RSFM_DBG(i, pred)
#define RSFM_DBGW(i, pred)
#define RSFM_DBGREC(pred, a0, a1, a2, a3, a4, a5, a6, a7)
#endif

struct SolveOut {
    double sse;
    int status;
    int filled;
    uint32_t nrhs, nstep;
};

constexpr int LTAB_STRIDE = 32;     // doubles per warp: two buffers of 16 (11 stage values used in each)
constexpr int NOM_STRIDE = 16;      // doubles per interval in the precomputed nominal table

// Per-block shared scratch for the stage values of L: one shared table per warp (filled by lanes
// 0..10, one exp + one sin per warp and step) and a private column per thread for lanes whose
// (t, h) differs from the table's key (stiff regime).
struct LoadScratch {
    double *tab;        // [warps_per_block][LTAB_STRIDE]
    double *priv;       // [11][blockDim.x]
    const double *nom;  // global [n_out][NOM_STRIDE] or nullptr: nominal table built by loading_table_kernel.
                        // Entry k = stage values of L for the single step that covers output interval k on the
                        // nominal time grid (t_k, h_k = (t_k + delta_t) - t_k), then the key t_k, h_k.  Outside the
                        // stiff regime every chain walks exactly this grid, so no exp/sin is evaluated in the solve.
};

// Per-lane solver state that crosses an output-interval boundary (what rsf_interval_general reads and updates).
struct IvState {
    double t, mu, th, V, rth, k1m, k1t, k1v, h_carry, tab_t, tab_h;
    int status, failed;
    uint32_t nrhs, nstep;
};

// One output interval on the general path: hinit + the dp86co step loop (everything the fast interval of
// rsf_solve_mode cannot do in one verified block: the start-up interval, the stiff regime, rejected steps,
// arguments outside the fast ranges).  Out of line on purpose: the fast interval is the hot path of every
// non-stiff configuration and keeps the kernel's register budget and instruction schedule to itself; this
// function is compiled with its own.  All 32 lanes of the warp call it together.
template <bool PARITY>
__device__ __forceinline__ void rsf_interval_general(const ModelK *Mp, const ChainConst *ccp, IvState *S, double *wtab,
                                                  double *ptab, int k, int running_i)
{
    const ModelK &M = *Mp;
    const ChainConst &cc = *ccp;
    constexpr bool parity = PARITY;
    const int lane = threadIdx.x & 31;
    const int nthr = blockDim.x;
    const double uround = 2.3e-16, safe = 0.9;
    const double facc1 = 1.0 / 0.3, facc2 = 1.0 / 6.0;
    // Re-basing (VSTEP loading).  The friction law is invariant under a change of reference velocity:
    // with V_ref' = lam V_ref and mu_ref' = mu_ref + (a - b) ln lam it reads the same in (mu, theta, v), so
    // the RHS may be evaluated relative to the level the load point currently moves at, lam = 1 + L(t):
    //     w' = lam w,  k'V_ref' = lam kV,  V_ref'/a = lam voa0,  1 + L' = (1 + L)/lam  =>  L' = (L - Lc)/lam.
    // Around the sliding steady state OF THAT LEVEL (v = lam V_ref, theta = Dc/(lam V_ref)) f' and A' are small,
    // so the short series of the fast step hold there too.  Without it every step of a raised-velocity
    // period (f = 1/lam - 1 = -0.9 at cfg 4) runs the general-range step (libdevice log / exp, division).
    // The base is chosen per output interval from L at its start; a load that switches inside the
    // interval simply shows up as L' != 0 (any L' is valid; the range flags still guard the series).
    const bool vstep = M.loading == RSFM_LOAD_VSTEP && M.vstep_factor != 1.0;
    double t = S->t, mu = S->mu, th = S->th, V = S->V, rth = S->rth, k1m = S->k1m, k1t = S->k1t, k1v = S->k1v;
    double h_carry = S->h_carry, tab_t = S->tab_t, tab_h = S->tab_h;
    struct { int status; uint32_t nrhs, nstep; } out = {S->status, S->nrhs, S->nstep};
    bool failed = S->failed != 0;
    const bool running = running_i != 0;
    const double xend = t + M.delta_t;                         // :382
    const double hmax = fabs(xend - t);
    int nstep_call = 0;
    bool reject = false, last = false;
    double h;
    // Key (tab_t, tab_h) of the warp's shared table.  It is rebuilt when some lane misses and at
    // least two lanes would share the new key (always, outside the stiff regime).
    auto ensure_table = [&](bool want, double hk) {
        const bool miss = want && !(t == tab_t && hk == tab_h);
        const unsigned mm = __ballot_sync(FULL_MASK, miss);
        if (mm != 0) {
            const int src = __ffs(mm) - 1;
            const double nt = __shfl_sync(FULL_MASK, t, src), nh = __shfl_sync(FULL_MASK, hk, src);
            const unsigned share = __ballot_sync(FULL_MASK, want && t == nt && hk == nh);
            if (__popc(share) >= 2) {
                tab_t = nt; tab_h = nh;
                __syncwarp();                                      // readers of the old table are done
                if (lane < 11) wtab[lane] = loading_of(M, __dadd_rn(nt, __dmul_rn(TB.c[lane], nh)));
                __syncwarp();
            }
        }
    };

    double Lc = 0.0, rl = 1.0, lam = 1.0, mu_ref_r = cc.mu_ref;
    // [pw_lo, pw_hi]: times at which the piecewise-constant load certainly has the value pw_L (the period of t,
    // shrunk by a guard far above the rounding of the index computation); a step inside it needs no index
    // arithmetic at all, any other one takes the exact per-stage route.
    double pw_lo = 1.0, pw_hi = 0.0, pw_L = 0.0;
    if (M.loading == RSFM_LOAD_VSTEP) {
        const double ia = vstep_index(M, t);
        const double guard = 1e-6 * M.vstep_period;
        pw_lo = fma(ia, M.vstep_period, M.t_start) + guard;
        pw_hi = fma(ia + 1.0, M.vstep_period, M.t_start) - guard;
        pw_L = ((long long)ia & 1) ? M.vstep_factor - 1.0 : 0.0;
        if (vstep && pw_L != 0.0) {
            Lc = pw_L; rl = M.vstep_rfac; lam = M.vstep_factor;
            mu_ref_r = fma(1.0 / cc.inv_a - cc.b, M.vstep_lnf, cc.mu_ref);     // (a - b) ln lam: a, b not kept live
        }
    }
    const ChainConst cr = rebase_const(cc, lam, rl, mu_ref_r);
    if (parity || k == 1) {
        // ---- HINIT (dop853.f, iord = 8).  The common outcome h0 = h = hmax is recognised by
        // comparisons in the squared / 16th-power domain, without sqrt or division. ----
        // Norms over the common denominator D = (sk0 sk1 sk2)^2:  ||f/sk||^2 = Nf/D etc.
        const double k0 = M.atol + M.rtol * fabs(mu), k1 = M.atol + M.rtol * fabs(th), k2 = M.atol + M.rtol * fabs(V);
        const double p0 = k1 * k2, p1 = k0 * k2, p2 = k0 * k1;
        const double dd = k0 * p0, D = dd * dd;
        const double Nf = (k1m * p0) * (k1m * p0) + (k1t * p1) * (k1t * p1) + (k1v * p2) * (k1v * p2);
        const double Ny = (mu * p0) * (mu * p0) + (th * p1) * (th * p1) + (V * p2) * (V * p2);
        // h0 = min(0.01 sqrt(dny/dnf), hmax)  (1e-6 when either norm^2 <= 1e-10)
        const bool tiny = (Nf <= 1e-10 * D) || (Ny <= 1e-10 * D);
        const bool h0max = !tiny && (Ny >= Nf * (1.0e4 * hmax * hmax));
        double h0 = hmax;
        if (!h0max) h0 = tiny ? fmin(1.0e-6, hmax) : fmin(sqrt(Ny / Nf) * 0.01, hmax);
        // the probe point t + h0 is stage 12 of the step (t, hmax): speculate on that table
        double Lp;
        if (M.loading == RSFM_LOAD_VSTEP) {
            Lp = loading_of(M, t + h0);                   // piecewise constant: nothing to share
        } else {
            ensure_table(running, hmax);
            Lp = wtab[10];
            if (running && !(t == tab_t && h0 == tab_h)) Lp = loading_of(M, t + h0);
        }
        double f1m, f1t, f1v, rprobe = rth;
        rsf_rhs_checked(cr, cc, lam, rl, mu_ref_r, (Lp - Lc) * rl, mu + h0 * k1m, th + h0 * k1t, rprobe, f1m, f1t, f1v);
        if (running) out.nrhs++;
        const double e0 = (f1m - k1m) * p0, e1 = (f1t - k1t) * p1, e2 = (f1v - k1v) * p2;
        const double Ne = e0 * e0 + e1 * e1 + e2 * e2;                      // ||(f1-f0)/sk||^2 = Ne/D
        // h = min(100 h0, (0.01/der12)^(1/8), hmax) with der12^2 = max(Ne/(D h0^2), Nf/D).
        // (0.01/der12)^(1/8) >= hmax  <=>  der12^2 hmax^16 <= 1e-4: decided without sqrt / division
        const double hm2 = hmax * hmax, hm4 = hm2 * hm2, hm8 = hm4 * hm4, hm16 = hm8 * hm8;
        const double lim = 1.0e-4 * D;
        if (h0max && Ne * hm16 <= lim * hm2 && Nf * hm16 <= lim && fmax(Ne, Nf * hm2) > 1e-30 * D * hm2) {
            h = hmax;
        } else {
            const double d12sq = fmax(Ne / (D * h0 * h0), Nf / D);
            double h1;
            if (d12sq <= 1e-30) h1 = fmax(1.0e-6, fabs(h0) * 1.0e-3);
            else h1 = root8(0.01 / sqrt(d12sq));
            h = fmin(fmin(100.0 * fabs(h0), h1), hmax);
        }
    } else {
        h = fmin(h_carry, hmax);
    }

    // ---- dp86co step loop: every iteration is one attempted step of all unfinished lanes ----
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
        if (__ballot_sync(FULL_MASK, !done) == 0) break;
        const bool stepping = !done;
        const bool pwc = M.loading == RSFM_LOAD_VSTEP;      // piecewise-constant load: no table to share
        if (!pwc) ensure_table(stepping, h);
        const bool hit = !pwc && (t == tab_t && h == tab_h);
        int lstride = hit ? 1 : nthr;
        bool on_base = false;                                 // the load stays at the frame's base level over the step
        if (stepping && !hit) {
            // private stage values (this lane is not on the warp's (t, h)).  A piecewise-constant load
            // that does not switch between t and t + h has one value for the whole step.
            bool flat = t >= pw_lo && t + h <= pw_hi;
            double La = pw_L;
            if (!flat && M.loading == RSFM_LOAD_VSTEP) {
                const double ia = vstep_index(M, t), ib = vstep_index(M, t + h);
                flat = ia == ib;
                La = ((long long)ia & 1) ? M.vstep_factor - 1.0 : 0.0;
            }
            if (flat) {
                ptab[0] = La;                                 // one value for every stage (stride 0)
                lstride = 0;
                on_base = pwc && La == Lc;
            } else {
#pragma unroll 1
                for (int i = 0; i < 11; i++) ptab[i * nthr] = loading_of(M, __dadd_rn(t, __dmul_rn(TB.c[i], h)));
            }
        }
        const double *Lsrc = hit ? wtab : ptab;

        StepIn in;
        in.h = h; in.mu = mu; in.th = th; in.V = V; in.k1m = k1m; in.k1t = k1t; in.k1v = k1v; in.rth = rth;
        in.atol = M.atol; in.rtol = M.rtol;
        StepOut so;
        bool bad = false;
        // The fast step is tried when some stepping lane STARTS inside the fast ranges (with a margin).
        // A step that starts inside and leaves them at a late stage (the trial step the controller grew
        // past the stability limit) says nothing about the state: the next one is tried again.
        const double f0s = state_excess(cr, th), A0s = (mu - cr.mu_ref) * cr.inv_a;
        const bool start_in = fabs(f0s) * cr.qscale < 0.5 * 0.001953125 && fabs(A0s) < 0.5 * 0.015625;
        const bool try_fast = __any_sync(FULL_MASK, stepping && start_in);
        if (try_fast) bad = dop853_step_fast_rb(&cc, lam, rl, mu_ref_r, &in, Lsrc, lstride, Lc, &so);
        else bad = true;
        // A step under a CONSTANT load at the frame's base level that STARTS well inside the fast ranges (i.e. close to
        // the sliding steady state of that very level) and leaves them at an internal stage is the trial step the
        // controller grew past the stability limit: it explodes and SciPy rejects it.  It is taken as rejected without
        // scoring it with the general-range stages (a rejected step shrinks by exactly 0.3 whatever its error was).
        // The load condition matters: the accumulated output times put a velocity jump a few ulp INSIDE an interval
        // whose frame is still the old level; the first steps after the jump start at the old steady state, leave the
        // ranges because the state really moves, and are accepted (err ~ 1e-13) -- those are scored
        // (profiles/microbench/stiff_wild_probe.py: one such step per downward/upward jump, 1e-6 of the trajectory).
        // CPU-oracle count (tests/test_stiff_rule.py; profiles/microbench/forward_stiff_r1b.txt): of ~49,000 steps that
        // meet the condition none is accepted by the exact arithmetic; were one ever, the retry at 0.3 h only costs a
        // step.  cfg.stiff_exact scores them with the general-range step instead.
        const bool presumed_wild = bad && start_in && on_base && !M.stiff_exact;
        if (stepping && bad && !presumed_wild) dop853_step_general(&cc, lam, mu_ref_r, &in, Lsrc, lstride, Lc, rl, &so);
        RSFM_DBG(0, stepping) RSFM_DBG(1, stepping && try_fast) RSFM_DBG(2, stepping && bad)
        RSFM_DBGW(3, stepping) RSFM_DBGW(4, stepping && bad) RSFM_DBGW(5, stepping && try_fast)
        RSFM_DBG(8, stepping && lam != 1.0) RSFM_DBG(9, stepping && bad && start_in)
        // err <= 1   <=>   h^2 errA^2 <= den3   (no sqrt, no division; NaN rejects; 0 <= 0 accepts).
        // errA < 1e140 keeps the squares finite: an unstable step whose error norm overflows must be
        // rejected (dop853.f gets inf * 0 = NaN there), not pass as inf <= inf.
        const bool accept = !presumed_wild && so.errA < 1e140 && (h * h) * (so.errA * so.errA) <= so.den3;

        RSFM_DBG(6, stepping && accept) RSFM_DBG(10, stepping && !accept && bad)
        RSFM_DBG(11, stepping && accept && bad && start_in)
        RSFM_DBGREC(stepping && accept && bad && start_in, t, h, f0s * cr.qscale / 0.001953125, A0s / 0.015625, so.errA, so.den3,
                    cc.w, lam)
        if (stepping) {
            out.nstep++;
            out.nrhs += 11;
            if (accept) {
                // FSAL: f(x + h, y_new) is k1 of the next step and of the next interval's restart
                rth = so.rth;
                rsf_rhs_checked(cr, cc, lam, rl, mu_ref_r, (so.L12 - Lc) * rl, so.muN, so.thN, rth, k1m, k1t, k1v);
                out.nrhs++;
                const double hold = h;
                mu = so.muN; th = so.thN; V = so.VN; t = t + h;
                // the controller's h_new is only consumed when the step does not end the
                // interval, or when the step size is carried across output points
                // (h_new = h / max(1/6, min(1/0.3, err^(1/8)/0.9)) >= h when err <= 0.9^8; with h = hmax it is
                //  clamped back to hmax, so the 8th root is skipped)
                const bool keeps_hmax = (hold == hmax) && !reject &&
                                        (hold * hold) * (so.errA * so.errA) <= 0.185302018885184 * so.den3;
                if ((!last || !parity) && !keeps_hmax) {
                    // h_new = h g,  g = 1/fac = clamp(0.9 err^(-1/8), 0.3, 6),  err^2 = q/den3  (accepted: err <= 1)
                    const double q = (hold * hold) * (so.errA * so.errA);
                    double g;
                    if (q <= 6.568408355712891e-14 * so.den3) {
                        g = 6.0;                                   // err <= 0.15^8 (and den3 = 0): growth limit
                    } else if (q > 1e-290 && so.den3 > 1e-290 && so.den3 < 1e290) {
                        g = fmin(6.0, fmax(0.3, safe * inv_root16(q, so.den3)));
                    } else {                                       // out of inv_root16's range: as written in dop853.f
                        const double err = fabs(hold) * so.errA / sqrt(so.den3);
                        g = 1.0 / fmax(facc2, fmin(facc1, root8(err) / safe));
                    }
                    double hnew = hold * g;
                    if (fabs(hnew) > hmax) hnew = hmax;
                    if (reject) hnew = fmin(fabs(hnew), fabs(hold));
                    h = hnew;
                }
                reject = false;
                if (last) { done = true; h_carry = h; }
            } else {
                // rejected (also NaN).  SciPy 1.18.1's dop853 shrinks by exactly 1/facc1 here.
                h = h / facc1;
                reject = true;
                last = false;
            }
        }
    }

    S->t = t; S->mu = mu; S->th = th; S->V = V; S->rth = rth; S->k1m = k1m; S->k1t = k1t; S->k1v = k1v;
    S->h_carry = h_carry; S->tab_t = tab_t; S->tab_h = tab_h;
    S->status = out.status; S->nrhs = out.nrhs; S->nstep = out.nstep; S->failed = failed ? 1 : 0;
}

// The general interval path of the default (non-stiff) variant, out of line: same arithmetic as ever (SciPy's
// controller as written, libdevice roots), but its ~35 KB of code no longer sit in the middle of the output loop,
// whose hot part (fast interval, FSAL, output) then spans ~22 KB instead of ~58 KB of addresses (instruction-fetch
// stalls at the loop head, profiles/README.md).  ISO kernels only (see rsf_solve_mode).
template <bool PARITY>
__device__ __forceinline__ void rsf_interval_plain_body(const ModelK *Mp, const ChainConst *ccp, IvState *S, double *wtab,
                                                        double *ptab, int k, int running_i, int *was_bad_io,
                                                        unsigned int *stiff_steps_io)
{
    const ModelK &M = *Mp;
    const ChainConst &cc = *ccp;
    constexpr bool parity = PARITY;
    const int lane = threadIdx.x & 31;
    const int nthr = blockDim.x;
    const double uround = 2.3e-16, safe = 0.9;
    const double facc1 = 1.0 / 0.3, facc2 = 1.0 / 6.0;
    double t = S->t, mu = S->mu, th = S->th, V = S->V, rth = S->rth, k1m = S->k1m, k1t = S->k1t, k1v = S->k1v;
    double h_carry = S->h_carry, tab_t = S->tab_t, tab_h = S->tab_h;
    struct { int status; uint32_t nrhs, nstep; } out = {S->status, S->nrhs, S->nstep};
    bool failed = S->failed != 0;
    const bool running = running_i != 0;
    bool was_bad = *was_bad_io != 0;
    unsigned int stiff_steps = *stiff_steps_io;
    const double xend = t + M.delta_t;                         // :382
    const double hmax = fabs(xend - t);
    int nstep_call = 0;
    bool reject = false, last = false;
    double h;
    // Key (tab_t, tab_h) of the warp's shared table.  It is rebuilt when some lane misses and at
    // least two lanes would share the new key (always, outside the stiff regime).
    auto ensure_table = [&](bool want, double hk) {
        const bool miss = want && !(t == tab_t && hk == tab_h);
        const unsigned mm = __ballot_sync(FULL_MASK, miss);
        if (mm != 0) {
            const int src = __ffs(mm) - 1;
            const double nt = __shfl_sync(FULL_MASK, t, src), nh = __shfl_sync(FULL_MASK, hk, src);
            const unsigned share = __ballot_sync(FULL_MASK, want && t == nt && hk == nh);
            if (__popc(share) >= 2) {
                tab_t = nt; tab_h = nh;
                __syncwarp();                                      // readers of the old table are done
                if (lane < 11) wtab[lane] = loading_of(M, __dadd_rn(nt, __dmul_rn(TB.c[lane], nh)));
                __syncwarp();
            }
        }
    };

    if (parity || k == 1) {
        // ---- HINIT (dop853.f, iord = 8).  The common outcome h0 = h = hmax is recognised by
        // comparisons in the squared / 16th-power domain, without sqrt or division. ----
        // Norms over the common denominator D = (sk0 sk1 sk2)^2:  ||f/sk||^2 = Nf/D etc.
        const double k0 = M.atol + M.rtol * fabs(mu), k1 = M.atol + M.rtol * fabs(th), k2 = M.atol + M.rtol * fabs(V);
        const double p0 = k1 * k2, p1 = k0 * k2, p2 = k0 * k1;
        const double dd = k0 * p0, D = dd * dd;
        const double Nf = (k1m * p0) * (k1m * p0) + (k1t * p1) * (k1t * p1) + (k1v * p2) * (k1v * p2);
        const double Ny = (mu * p0) * (mu * p0) + (th * p1) * (th * p1) + (V * p2) * (V * p2);
        // h0 = min(0.01 sqrt(dny/dnf), hmax)  (1e-6 when either norm^2 <= 1e-10)
        const bool tiny = (Nf <= 1e-10 * D) || (Ny <= 1e-10 * D);
        const bool h0max = !tiny && (Ny >= Nf * (1.0e4 * hmax * hmax));
        double h0 = hmax;
        if (!h0max) h0 = tiny ? fmin(1.0e-6, hmax) : fmin(sqrt(Ny / Nf) * 0.01, hmax);
        // the probe point t + h0 is stage 12 of the step (t, hmax): speculate on that table
        ensure_table(running, hmax);
        double Lp = wtab[10];
        if (running && !(t == tab_t && h0 == tab_h)) Lp = loading_of(M, t + h0);
        double f1m, f1t, f1v, rprobe = rth;
        rsf_rhs_checked(cc, Lp, mu + h0 * k1m, th + h0 * k1t, rprobe, f1m, f1t, f1v);
        if (running) out.nrhs++;
        const double e0 = (f1m - k1m) * p0, e1 = (f1t - k1t) * p1, e2 = (f1v - k1v) * p2;
        const double Ne = e0 * e0 + e1 * e1 + e2 * e2;                      // ||(f1-f0)/sk||^2 = Ne/D
        // h = min(100 h0, (0.01/der12)^(1/8), hmax) with der12^2 = max(Ne/(D h0^2), Nf/D).
        // (0.01/der12)^(1/8) >= hmax  <=>  der12^2 hmax^16 <= 1e-4: decided without sqrt / division
        const double hm2 = hmax * hmax, hm4 = hm2 * hm2, hm8 = hm4 * hm4, hm16 = hm8 * hm8;
        const double lim = 1.0e-4 * D;
        if (h0max && Ne * hm16 <= lim * hm2 && Nf * hm16 <= lim && fmax(Ne, Nf * hm2) > 1e-30 * D * hm2) {
            h = hmax;
        } else {
            const double d12sq = fmax(Ne / (D * h0 * h0), Nf / D);
            double h1;
            if (d12sq <= 1e-30) h1 = fmax(1.0e-6, fabs(h0) * 1.0e-3);
            else h1 = root8(0.01 / sqrt(d12sq));
            h = fmin(fmin(100.0 * fabs(h0), h1), hmax);
        }
    } else {
        h = fmin(h_carry, hmax);
    }

    // ---- dp86co step loop: every iteration is one attempted step of all unfinished lanes ----
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
        if (__ballot_sync(FULL_MASK, !done) == 0) break;
        const bool stepping = !done;
        ensure_table(stepping, h);
        const bool hit = (t == tab_t && h == tab_h);
        if (stepping && !hit) {
            // private stage values (this lane is not on the warp's (t, h)).  A piecewise-constant load
            // that does not switch between t and t + h has one value for the whole step.
            const double La = loading_of(M, t), Lb = loading_of(M, t + h);
            if (M.loading == RSFM_LOAD_VSTEP && La == Lb &&
                floor((t - M.t_start) / M.vstep_period) == floor((t + h - M.t_start) / M.vstep_period)) {
#pragma unroll
                for (int i = 0; i < 11; i++) ptab[i * nthr] = La;
            } else {
#pragma unroll 1
                for (int i = 0; i < 11; i++) ptab[i * nthr] = loading_of(M, __dadd_rn(t, __dmul_rn(TB.c[i], h)));
            }
        }
        const double *Lsrc = hit ? wtab : ptab;
        const int lstride = hit ? 1 : nthr;

        StepIn in;
        in.h = h; in.mu = mu; in.th = th; in.V = V; in.k1m = k1m; in.k1t = k1t; in.k1v = k1v; in.rth = rth;
        in.atol = M.atol; in.rtol = M.rtol;
        StepOut so;
        bool bad = false;
        // A warp whose stepping lanes all left the fast ranges on their previous step (stiff regime)
        // goes straight to the general step; the fast one is retried every 64 steps.
        const bool try_fast = !__all_sync(FULL_MASK, !stepping || was_bad) || (++stiff_steps & 63u) == 0u;
        if (try_fast) dop853_step_fast(cc, in, Lsrc, lstride, so, bad);
        else bad = true;
        if (stepping && bad) dop853_step_general(&cc, &in, Lsrc, lstride, &so);
        if (stepping) was_bad = bad;
        // err <= 1   <=>   h^2 errA^2 <= den3   (no sqrt, no division; NaN rejects; 0 <= 0 accepts).
        // errA < 1e140 keeps the squares finite: an unstable step whose error norm overflows must be
        // rejected (dop853.f gets inf * 0 = NaN there), not pass as inf <= inf.
        const bool accept = so.errA < 1e140 && (h * h) * (so.errA * so.errA) <= so.den3;

        if (stepping) {
            out.nstep++;
            out.nrhs += 11;
            if (accept) {
                // FSAL: f(x + h, y_new) is k1 of the next step and of the next interval's restart
                rth = so.rth;
                rsf_rhs_checked(cc, so.L12, so.muN, so.thN, rth, k1m, k1t, k1v);
                out.nrhs++;
                const double hold = h;
                mu = so.muN; th = so.thN; V = so.VN; t = t + h;
                // the controller's h_new is only consumed when the step does not end the
                // interval, or when the step size is carried across output points
                // (h_new = h / max(1/6, min(1/0.3, err^(1/8)/0.9)) >= h when err <= 0.9^8; with h = hmax it is
                //  clamped back to hmax, so the 8th root is skipped)
                const bool keeps_hmax = (hold == hmax) && !reject &&
                                        (hold * hold) * (so.errA * so.errA) <= 0.185302018885184 * so.den3;
                if ((!last || !parity) && !keeps_hmax) {
                    const double err = so.den3 > 0.0 ? fabs(hold) * so.errA / sqrt(so.den3) : 0.0;
                    const double fac = fmax(facc2, fmin(facc1, root8(err) / safe));
                    double hnew = hold / fac;
                    if (fabs(hnew) > hmax) hnew = hmax;
                    if (reject) hnew = fmin(fabs(hnew), fabs(hold));
                    h = hnew;
                }
                reject = false;
                if (last) { done = true; h_carry = h; }
            } else {
                // rejected (also NaN).  SciPy 1.18.1's dop853 shrinks by exactly 1/facc1 here.
                h = h / facc1;
                reject = true;
                last = false;
            }
        }
    }

    S->t = t; S->mu = mu; S->th = th; S->V = V; S->rth = rth; S->k1m = k1m; S->k1t = k1t; S->k1v = k1v;
    S->h_carry = h_carry; S->tab_t = tab_t; S->tab_h = tab_h;
    S->status = out.status; S->nrhs = out.nrhs; S->nstep = out.nstep; S->failed = failed ? 1 : 0;
    *was_bad_io = was_bad ? 1 : 0; *stiff_steps_io = stiff_steps;
}

template <bool PARITY>
__device__ __noinline__ void rsf_interval_plain(const ModelK *Mp, const ChainConst *ccp, IvState *S, double *wtab,
                                                double *ptab, int k, int running_i, int *was_bad_io,
                                                unsigned int *stiff_steps_io)
{
    rsf_interval_plain_body<PARITY>(Mp, ccp, S, wtab, ptab, k, running_i, was_bad_io, stiff_steps_io);
}

// Integrate one chain over the whole output grid.  All 32 lanes of a warp must call this together
// (it contains warp collectives) and all threads of a block must call it together when
// `series.g != nullptr` (block barriers at tile boundaries).  `active` = this lane owns a chain.
// acc_out / t_out (optional) are written time-major with stride `acc_stride` (= C).
// acc_ref (optional, same layout): the solve also accumulates xtx += ((acc - acc_ref[k]) / fd_den)^2
// (MCMC.py:264-265).
// sse_limit: the partial sums of squares only grow with k, so once they exceed the Metropolis
// threshold SS_cur - 2 sigma^2 ln U the proposal is certainly rejected (MCMC.py:327-331 with U drawn
// first); the lane stops integrating there (`out.status |= RSFM_CHAIN_EARLY`, sse = partial > limit).
// The accept/reject decision is exactly the one the full solve would give.  A warp leaves the
// output loop when all its lanes are finished (resident series only: no block barriers pending).
// K1P (RSFM_PARAM_K1, extension): the per-chain scalar passed as `dc` is the radiation-damping coefficient k1
// (RateStateModel.py:171, 349-353) and Dc is the model's constant M.dc_fixed -- separate instantiations, so that k1
// stays a kernel-uniform constant of the fast-interval block everywhere else.
template <bool PARITY, bool VS, bool ISO, bool K1P = false>
__device__ __forceinline__ SolveOut rsf_solve_mode(const ModelK &M, double a, double b, double dc, bool active,
                                                    SeriesStage &series, const LoadScratch &ls, double *acc_out,
                                                    const double *acc_ref, size_t acc_stride, double fd_den,
                                                    double *xtx_out, double *t_out, double sse_limit)
{
    const int lane = threadIdx.x & 31;
    double *const wbase = ls.tab + (threadIdx.x >> 5) * LTAB_STRIDE;
    double *wtab = wbase;                  // current buffer (wbase or wbase + 16)
    double *ptab = ls.priv + threadIdx.x;
    const double k1_chain = dc;                        // (K1P only)
    if constexpr (K1P) dc = M.dc_fixed;
    const ChainConst cc = [&] {
        ChainConst c_ = make_chain_const(M, a, b, dc);
        if constexpr (K1P) {
            c_.k1e = M.damping ? k1_chain : 0.0;
            // The fast stage takes the damping correction of the stage in flight to first order,
            // e^(A + delta) - 1 = u + delta (1 + u) with delta = (h/a) a_sj k1 V' (rsf_stage_fast): 1e-16 at the
            // reference's k1 = 1e-7, but delta^2/2 ~ 4e-10 at k1 = 1e-3.  Away from the reference's value the steps
            // are therefore scored by the general-range stages (the reference's formulas, exact in k1), selected the
            // way the slip law selects them: an infinite range scale, no instruction in the fast stage.
            if (!(fabs(c_.k1e) <= 4.0e-7)) c_.qscale = INFINITY;
        }
        return c_;
    }();
    const bool have_data = series.g != nullptr;
    constexpr bool parity = PARITY;      // compile-time: keeps the fast interval one branch-free block
    const double uround = 2.3e-16;

    double t = M.t_start;
    double mu = M.mu_t_zero, th = dc / M.V_ref, V = M.V_ref;       // :367-370,377
    double rth = 1.0 / th;
    double k1m, k1t, k1v;
    SolveOut out;
    out.status = RSFM_CHAIN_OK;
    out.filled = M.n_out;
    out.nrhs = 0; out.nstep = 0;
    double sse = 0.0, xtx = 0.0;
    // observable: acc (reference, acc[0] = 0, :371) or the friction series mu (mu[0] = mu_ref, :367)
    const bool obs_mu = M.observable == RSFM_OBS_MU;
    const double obs0 = obs_mu ? M.mu_ref : 0.0;
    if (have_data) { const double d0 = obs0 - series.at(0); sse = d0 * d0; }
    if (acc_out && active) acc_out[0] = obs0;
    if (t_out && active) t_out[0] = t;
    if (acc_ref && active) { const double x0 = (obs0 - acc_ref[0]) / fd_den; xtx = x0 * x0; }

    rsf_rhs_checked(cc, loading_of(M, t), mu, th, rth, k1m, k1t, k1v);
    out.nrhs++;
    bool failed = false;
    double vprev = V;
    double h_carry = 0.0;

    // Key (tab_t, tab_h) of the warp's shared table (rebuilt by the general interval path when lanes leave the
    // nominal grid).
    double tab_t = 0.0, tab_h = -1.0;
    // Nominal table (global, built by loading_table_kernel): entry k is fetched one interval ahead with
    // cp.async into the other half of the warp's double buffer, so no load latency is ever exposed.
    // Its key (t_k, h_k) is the recurrence every lane can run itself: t_{k+1} = t_k + ((t_k + dt) - t_k).
    const double *nom = ls.nom;
    double tn = M.t_start;
    if (nom != nullptr && M.n_out > 1) {
        if (lane < 11) cp_async8(wbase + 16 + lane, nom + NOM_STRIDE + lane);      // entry 1 -> buffer 1
        cp_async_commit();
    }
    const double inv_dt = 1.0 / M.delta_t;
    double acc_pending = 0.0, dk_pending = 0.0;      // output point whose SSE term is folded in one interval late
    bool have_pending = false;
    int fast_resume = 0;
    bool was_bad = false;
    unsigned int stiff_steps = 0;          // warp-uniform count of steps taken on the general path

    for (int k = 1; k < M.n_out; k++) {
        const double dk = have_data ? series.at(k) : 0.0;
        if (nom != nullptr) {
            // install the nominal entry of this interval (buffer k & 1) and prefetch the next one
            cp_async_wait_all();
            __syncwarp();
            wtab = wbase + ((k & 1) << 4);
            const double xn = tn + M.delta_t;
            tab_t = tn; tab_h = xn - tn;
            tn = tn + tab_h;
            if (k + 1 < M.n_out) {
                if (lane < 11) cp_async8(wbase + (((k + 1) & 1) << 4) + lane, nom + (size_t)(k + 1) * NOM_STRIDE + lane);
                cp_async_commit();
            }
        }
        // SSE term of the previous output point (its division overlaps this interval's first stages)
        if (have_pending) {
            const double e = acc_pending - dk_pending;
            sse += e * e;
            if (active && !failed && sse > sse_limit) { failed = true; out.status |= RSFM_CHAIN_EARLY; }
        }
        const bool running = active && !failed;
        const double xend = t + M.delta_t;                         // :382
        const double hmax = fabs(xend - t);

        // ---- fast interval: the whole warp is on the nominal grid ----
        // Outside the stiff regime every SciPy call of the reference does the same thing: hinit returns
        // hmax, dp86co clamps it to xend - t, the single step is accepted.  That case is executed as one
        // branch-free block (hinit's Euler probe and the twelve stages are independent given (t, y, k1)
        // and interleave), verified by ONE warp vote; any lane that deviates (h0 or h below hmax,
        // rejected step, argument outside the fast ranges) sends the warp through the general path
        // below for this interval, which recomputes it from the same state.
        bool fast_done = false;
        if (k > 1 && k >= fast_resume && nom != nullptr) {
            // (CARRY mode: no hinit; the carried step must already be hmax and must stay hmax)
            const bool cand = (t == tab_t) && (hmax == tab_h) && (parity || h_carry >= hmax);
            const unsigned rmask = __ballot_sync(FULL_MASK, running);
            if (rmask != 0 && __all_sync(FULL_MASK, !running || cand)) {
                const double k0 = M.atol + M.rtol * fabs(mu), k1 = M.atol + M.rtol * fabs(th), k2 = M.atol + M.rtol * fabs(V);
                const double p0 = k1 * k2, p1 = k0 * k2, p2 = k0 * k1;
                const double dd = k0 * p0, D = dd * dd;
                const double Nf = (k1m * p0) * (k1m * p0) + (k1t * p1) * (k1t * p1) + (k1v * p2) * (k1v * p2);
                const double Ny = (mu * p0) * (mu * p0) + (th * p1) * (th * p1) + (V * p2) * (V * p2);
                const bool h0max = !parity ||
                                   (!((Nf <= 1e-10 * D) || (Ny <= 1e-10 * D)) && (Ny >= Nf * (1.0e4 * hmax * hmax)));
                bool bad = false;
                double f1m = k1m, f1t = k1t, f1v = k1v, rprobe = rth;
                if (parity) rsf_rhs<true>(cc, wtab[10], mu + hmax * k1m, th + hmax * k1t, rprobe, f1m, f1t, f1v, bad);
                StepIn in;
                in.h = xend - t; in.mu = mu; in.th = th; in.V = V; in.k1m = k1m; in.k1t = k1t; in.k1v = k1v; in.rth = rth;
                in.atol = M.atol; in.rtol = M.rtol;
                StepOut so;
                dop853_step_fast(cc, in, wtab, 1, so, bad);
                const double e0 = (f1m - k1m) * p0, e1 = (f1t - k1t) * p1, e2 = (f1v - k1v) * p2;
                const double Ne = e0 * e0 + e1 * e1 + e2 * e2;
                const double hm2 = hmax * hmax, hm4 = hm2 * hm2, hm8 = hm4 * hm4, hm16 = hm8 * hm8;
                const double lim = 1.0e-4 * D;
                const bool h1max = !parity || (Ne * hm16 <= lim * hm2 && Nf * hm16 <= lim &&
                                               fmax(Ne, Nf * hm2) > 1e-30 * D * hm2);
                const bool entry = !(0.1 * hmax <= fabs(t) * uround) && ((t + 1.01 * hmax - xend) > 0.0);
                const double h2a2 = (in.h * in.h) * (so.errA * so.errA);
                const bool accept = so.errA < 1e140 && h2a2 <= so.den3;
                // CARRY: the controller must leave h at hmax (err <= 0.9^8), else the general path decides
                const bool keeps = parity || h2a2 <= 0.185302018885184 * so.den3;
                const bool ok = h0max && h1max && entry && accept && keeps && !bad;
                if (__all_sync(FULL_MASK, !running || ok)) {
                    if (running) {
                        rth = so.rth;
                        rsf_rhs_checked(cc, so.L12, so.muN, so.thN, rth, k1m, k1t, k1v);
                        mu = so.muN; th = so.thN; V = so.VN; t = t + in.h;
                        out.nrhs += parity ? 13 : 12; out.nstep++;
                        h_carry = hmax;
                    }
                    fast_done = true;
                    RSFM_DBGW(7, running)
                } else {
                    fast_resume = k + 8;       // back off: this warp is (partly) off the nominal regime
                }
            }
        }

        if constexpr (VS) {
        if (!fast_done) {
            IvState S;
            S.t = t; S.mu = mu; S.th = th; S.V = V; S.rth = rth; S.k1m = k1m; S.k1t = k1t; S.k1v = k1v;
            S.h_carry = h_carry; S.tab_t = tab_t; S.tab_h = tab_h;
            S.status = out.status; S.nrhs = out.nrhs; S.nstep = out.nstep; S.failed = failed ? 1 : 0;
            rsf_interval_general<PARITY>(&M, &cc, &S, wtab, ptab, k, running ? 1 : 0);
            t = S.t; mu = S.mu; th = S.th; V = S.V; rth = S.rth; k1m = S.k1m; k1t = S.k1t; k1v = S.k1v;
            h_carry = S.h_carry; tab_t = S.tab_t; tab_h = S.tab_h;
            out.status = S.status; out.nrhs = S.nrhs; out.nstep = S.nstep; failed = S.failed != 0;
        }
        (void)was_bad; (void)stiff_steps;
        } else {
        if (!fast_done) {
            IvState S;
            S.t = t; S.mu = mu; S.th = th; S.V = V; S.rth = rth; S.k1m = k1m; S.k1t = k1t; S.k1v = k1v;
            S.h_carry = h_carry; S.tab_t = tab_t; S.tab_h = tab_h;
            S.status = out.status; S.nrhs = out.nrhs; S.nstep = out.nstep; S.failed = failed ? 1 : 0;
            int wb = was_bad ? 1 : 0;
            if constexpr (ISO) rsf_interval_plain<PARITY>(&M, &cc, &S, wtab, ptab, k, running ? 1 : 0, &wb, &stiff_steps);
            else rsf_interval_plain_body<PARITY>(&M, &cc, &S, wtab, ptab, k, running ? 1 : 0, &wb, &stiff_steps);
            t = S.t; mu = S.mu; th = S.th; V = S.V; rth = S.rth; k1m = S.k1m; k1t = S.k1t; k1v = S.k1v;
            h_carry = S.h_carry; tab_t = S.tab_t; tab_h = S.tab_h;
            out.status = S.status; out.nrhs = S.nrhs; out.nstep = S.nstep; failed = S.failed != 0;
            was_bad = wb != 0;
        }
        }

        // ---- output point k: RateStateModel.py:384-388, MCMC.py:387 ----
        double accv = 0.0;
        if (running) {
            const double dv = V - vprev;
            const double qd = dv * inv_dt;                        // (V_k - V_{k-1}) / delta_t  (:388):
            accv = fma(fma(-qd, M.delta_t, dv), inv_dt, qd);      // reciprocal + one correction = correctly rounded
            vprev = V;
            if (obs_mu) accv = mu;                                // mu[k] = r.y[0]  (:385)
            if (failed && out.filled == M.n_out) out.filled = k + 1;
        }
        acc_pending = accv; dk_pending = dk; have_pending = have_data;
        if (series.resident && __all_sync(FULL_MASK, !active || (failed && (out.status & RSFM_CHAIN_EARLY)))) break;
        if (active) {
            if (acc_out) acc_out[(size_t)k * acc_stride] = accv;
            if (t_out) t_out[(size_t)k * acc_stride] = running ? t : 0.0;
            if (acc_ref) { const double x = (accv - acc_ref[(size_t)k * acc_stride]) / fd_den; xtx += x * x; }
        }
    }
    if (have_pending) { const double e = acc_pending - dk_pending; sse += e * e; }
    out.sse = sse;
    if (xtx_out) *xtx_out = xtx;
    return out;
}

// integration mode is a kernel-uniform run-time choice; each mode is its own instantiation
template <bool VS, bool ISO = false, bool K1P = false>
__device__ __forceinline__ SolveOut rsf_solve(const ModelK &M, double a, double b, double dc, bool active,
                                               SeriesStage &series, const LoadScratch &ls, double *acc_out,
                                               const double *acc_ref, size_t acc_stride, double fd_den,
                                               double *xtx_out, double *t_out = nullptr,
                                               double sse_limit = INFINITY)
{
    if (M.integ_mode == RSFM_INTEG_PARITY)
        return rsf_solve_mode<true, VS, ISO, K1P>(M, a, b, dc, active, series, ls, acc_out, acc_ref, acc_stride, fd_den, xtx_out, t_out,
                                    sse_limit);
    return rsf_solve_mode<false, VS, ISO, K1P>(M, a, b, dc, active, series, ls, acc_out, acc_ref, acc_stride, fd_den, xtx_out, t_out,
                                 sse_limit);
}

}  // namespace rsfm
