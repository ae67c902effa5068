/* rsf_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 * See rsf_oracle.h for scope, citations and the pinning statement.
 * Build: make -C oracle      (gcc -O2 -ffp-contract=off -pthread)
 */
#define _GNU_SOURCE   /* M_PI */
#include "rsf_oracle.h"
#include "dop853_coeffs.h"

#include <math.h>
#ifdef ORC_TRACE
#include <stdio.h>
#endif
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

/* RateStateModel.py:5-11,167-184 and the solver settings of :374 */
/* Optional step instrumentation (tests/test_stiff_rule.py; single-threaded use only).  After every attempted
 * DOP853 step the hook receives the step, its error estimate and, relative to the limits of the CUDA fast step
 * (|f'| (1 + b/a) < 2^-9, |A'| < 2^-6 in the frame re-based on the load level at the start of the step), the excess
 * of the series arguments at the start of the step and the largest excess at any of its stages. */
static orc_step_hook_t g_hook = 0;
static double g_hook_lam = 1.0, g_hook_max = 0.0;
void orc_set_step_hook(orc_step_hook_t hook) { g_hook = hook; }

static double hook_ratio(const orc_model *m, const double y[3])
{
    const double fp = g_hook_lam * m->V_ref * y[1] / m->Dc - 1.0;
    const double Ap = (y[0] - (m->mu_ref + (m->a - m->b) * log(g_hook_lam))) / m->a;
    const double r1 = fabs(fp) * (1.0 + m->b / m->a) / 0.001953125, r2 = fabs(Ap) / 0.015625;
    const double r = r1 > r2 ? r1 : r2;
    return r == r ? r : HUGE_VAL;
}

void orc_model_defaults(orc_model *m)
{
    m->a = 0.011; m->b = 0.014; m->mu_ref = 0.6; m->V_ref = 1.0; m->k1 = 1.0e-7;
    m->Dc = 1000.0;
    m->t_start = 0.0; m->t_final = 50.0; m->num_tsteps = 500;
    m->mu_t_zero = 0.6;
    m->radiation_damping = 1;
    m->loading = ORC_LOAD_SINE_DECAY;
    m->vstep_period = 1000.0; m->vstep_factor = 10.0;
    m->rtol = 1e-6; m->atol = 1e-10; m->nmax = 500;
    m->observable = ORC_OBS_ACC;
    m->state_law = ORC_LAW_AGING;
    m->load_table = 0; m->n_load_table = 0; m->load_dt = 0.1;
    m->sampled_param = ORC_PARAM_DC;
}

/* load-point velocity.  SINE_DECAY is the reference (RateStateModel.py:327-329);
 * VSTEP is the piecewise-constant extension of SURVEY.md D1: V_ref on even
 * periods, vstep_factor*V_ref on odd periods. */
static double loading_velocity(const orc_model *m, double t)
{
    if (m->loading == ORC_LOAD_VSTEP) {
        double ph = floor((t - m->t_start) / m->vstep_period);
        int odd = ((long long)ph) & 1;
        return odd ? m->vstep_factor * m->V_ref : m->V_ref;
    }
    if (m->loading == ORC_LOAD_TABLE) {
        /* piecewise linear through the tabulated relative perturbation; x = (t - t_start)/load_dt as ONE rounded
         * division, the interpolation T[i] + fr (T[i+1] - T[i]) as written (no fused multiply-add) */
        const int n = m->n_load_table;
        const double x = (t - m->t_start) / m->load_dt;
        double fi = floor(x);
        if (fi < 0.0) fi = 0.0;
        if (fi > (double)(n - 2)) fi = (double)(n - 2);
        double fr = x - fi;
        if (fr < 0.0) fr = 0.0;
        if (fr > 1.0) fr = 1.0;
        const int i = (int)fi;
        const double d = m->load_table[i + 1] - m->load_table[i];
        volatile double prod = fr * d;                          /* keep the product rounded on its own */
        return m->V_ref * (1.0 + (m->load_table[i] + prod));
    }
    const double a1 = 20.0, a2 = 10.0;
    return m->V_ref * (1.0 + exp(-t / a1) * sin(a2 * t));
}

/* friction(t, y): RateStateModel.py:318-355, same operation order */
void orc_rhs(const orc_model *m, double t, const double y[3], double dydt[3])
{
    const double V_ref = m->V_ref, a = m->a, b = m->b, dc = m->Dc;
    const double kprime = 1e-2 * 10 / dc;                       /* :324 */
    const double V_l = loading_velocity(m, t);                  /* :329 */
    if (g_hook) { const double r = hook_ratio(m, y); if (r > g_hook_max) g_hook_max = r; }
    const double temp = 1 / a * (y[0] - m->mu_ref - b * log(V_ref * y[1] / dc));  /* :336 */
    const double v = V_ref * exp(temp);                         /* :337 */
    dydt[1] = 1. - v * y[1] / dc;                               /* :340 */
    if (m->state_law == ORC_LAW_SLIP) {                         /* Ruina slip law (extension) */
        const double z = v * y[1] / dc;
        dydt[1] = -z * log(z);
    }
    dydt[0] = kprime * V_l - kprime * v;                        /* :343 */
    dydt[2] = v / a * (dydt[0] - b / y[1] * dydt[1]);           /* :346 */
    if (m->radiation_damping) {                                 /* :349-353 */
        dydt[0] = dydt[0] - m->k1 * dydt[2];
        dydt[2] = v / a * (dydt[0] - b / y[1] * dydt[1]);
    }
}

/* Hairer's HINIT with iord = 8, scalar tolerances */
static double dop_hinit(const orc_model *m, double x, const double y[3],
                        double posneg, const double f0[3], double hmax,
                        orc_stats *st)
{
    const int n = 3;
    double dnf = 0.0, dny = 0.0, f1[3], y1[3];
    for (int i = 0; i < n; i++) {
        double sk = m->atol + m->rtol * fabs(y[i]);
        dnf += (f0[i] / sk) * (f0[i] / sk);
        dny += (y[i] / sk) * (y[i] / sk);
    }
    double h;
    if (dnf <= 1e-10 || dny <= 1e-10) h = 1.0e-6;
    else h = sqrt(dny / dnf) * 0.01;
    h = fmin(h, hmax);
    h = copysign(h, posneg);
    for (int i = 0; i < n; i++) y1[i] = y[i] + h * f0[i];
    orc_rhs(m, x + h, y1, f1);
    st->nrhs++;
    double der2 = 0.0;
    for (int i = 0; i < n; i++) {
        double sk = m->atol + m->rtol * fabs(y[i]);
        double d = (f1[i] - f0[i]) / sk;
        der2 += d * d;
    }
    der2 = sqrt(der2) / h;
    double der12 = fmax(fabs(der2), sqrt(dnf));
    double h1;
    if (der12 <= 1e-15) h1 = fmax(1.0e-6, fabs(h) * 1.0e-3);
    else h1 = pow(0.01 / der12, 1.0 / 8.0);
    h = fmin(fmin(100 * fabs(h), h1), hmax);
    return copysign(h, posneg);
}

/* One fresh dopri853 call (dp86co core), scipy defaults: safe 0.9, fac1 0.3,
 * fac2 6, beta 0, hmax = xend - x, h = 0 -> hinit, nmax = 500, iout = 0.
 * Stiffness detection (nstiff = 1000) cannot trigger within nmax = 500 steps. */
int orc_dop853_call(const orc_model *m, double *t, double y[3], double xend,
                    orc_stats *st)
{
    const int n = 3;
    const double uround = 2.3e-16, safe = 0.9, fac1 = 0.3, fac2 = 6.0, beta = 0.0;
    double x = *t;
    double k1[3], k2[3], k3[3], k4[3], k5[3], k6[3], k7[3], k8[3], k9[3], k10[3], y1[3];
    double facold = 1.0e-4;
    const double expo1 = 1.0 / 8.0 - beta * 0.2;
    const double facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    const double posneg = copysign(1.0, xend - x);
    int last = 0, reject = 0;
    int nstep = 0, naccpt = 0;
    double hmax = fabs(xend - x);

    orc_rhs(m, x, y, k1);
    st->nrhs++;
    double h = dop_hinit(m, x, y, posneg, k1, hmax, st);

    for (;;) {
        if (nstep > m->nmax) { *t = x; st->istate = -2; return -2; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { *t = x; st->istate = -3; return -3; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
        nstep++;
        st->nstep++;
        double hook_start = 0.0;
        if (g_hook) { g_hook_lam = loading_velocity(m, x) / m->V_ref; hook_start = hook_ratio(m, y); g_hook_max = 0.0; }
        for (int i = 0; i < n; i++) y1[i] = y[i] + h * DP_A2_1 * k1[i];
        orc_rhs(m, x + DP_C2 * h, y1, k2);
        for (int i = 0; i < n; i++) y1[i] = y[i] + h * (DP_A3_1 * k1[i] + DP_A3_2 * k2[i]);
        orc_rhs(m, x + DP_C3 * h, y1, k3);
        for (int i = 0; i < n; i++) y1[i] = y[i] + h * (DP_A4_1 * k1[i] + DP_A4_3 * k3[i]);
        orc_rhs(m, x + DP_C4 * h, y1, k4);
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A5_1 * k1[i] + DP_A5_3 * k3[i] + DP_A5_4 * k4[i]);
        orc_rhs(m, x + DP_C5 * h, y1, k5);
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A6_1 * k1[i] + DP_A6_4 * k4[i] + DP_A6_5 * k5[i]);
        orc_rhs(m, x + DP_C6 * h, y1, k6);
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] + DP_A7_6 * k6[i]);
        orc_rhs(m, x + DP_C7 * h, y1, k7);
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] + DP_A8_6 * k6[i]
                                + DP_A8_7 * k7[i]);
        orc_rhs(m, x + DP_C8 * h, y1, k8);
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] + DP_A9_6 * k6[i]
                                + DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
        orc_rhs(m, x + DP_C9 * h, y1, k9);
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A10_1 * k1[i] + DP_A10_4 * k4[i] + DP_A10_5 * k5[i] + DP_A10_6 * k6[i]
                                + DP_A10_7 * k7[i] + DP_A10_8 * k8[i] + DP_A10_9 * k9[i]);
        orc_rhs(m, x + DP_C10 * h, y1, k10);
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A11_1 * k1[i] + DP_A11_4 * k4[i] + DP_A11_5 * k5[i] + DP_A11_6 * k6[i]
                                + DP_A11_7 * k7[i] + DP_A11_8 * k8[i] + DP_A11_9 * k9[i]
                                + DP_A11_10 * k10[i]);
        orc_rhs(m, x + DP_C11 * h, y1, k2);
        const double xph = x + h;
        for (int i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A12_1 * k1[i] + DP_A12_4 * k4[i] + DP_A12_5 * k5[i] + DP_A12_6 * k6[i]
                                + DP_A12_7 * k7[i] + DP_A12_8 * k8[i] + DP_A12_9 * k9[i]
                                + DP_A12_10 * k10[i] + DP_A12_11 * k2[i]);
        orc_rhs(m, xph, y1, k3);
        st->nrhs += 11;
        for (int i = 0; i < n; i++) {
            k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] + DP_B9 * k9[i]
                    + DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }
        /* error estimation */
        double err = 0.0, err2 = 0.0;
        for (int i = 0; i < n; i++) {
            double sk = m->atol + m->rtol * fmax(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
            err2 += (erri / sk) * (erri / sk);
            erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] + DP_ER9 * k9[i]
                   + DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
            err += (erri / sk) * (erri / sk);
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (n * deno));
        if (g_hook) g_hook(x, h, err, hook_start, g_hook_max);
        /* step-size controller */
        double fac11 = pow(err, expo1);
        double fac = fac11 / pow(facold, beta);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        double hnew = h / fac;
#ifdef ORC_TRACE
        fprintf(stderr, "step x=%.10g h=%.10g err=%.6g %s\n", x - ORC_TRACE_T0, h, err, err <= 1.0 ? "ACC" : "REJ");
#endif
        if (err <= 1.0) {
            facold = fmax(err, 1.0e-4);
            naccpt++;
            st->naccpt++;
            orc_rhs(m, xph, k5, k4);
            st->nrhs++;
            for (int i = 0; i < n; i++) { k1[i] = k4[i]; y[i] = k5[i]; }
            x = xph;
            if (last) { *t = x; st->istate = 1; return 1; }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            /* dop853.f has hnew = h/min(facc1, fac11/safe).  SciPy 1.18.1's C
             * translation -- the solver the reference actually runs here -- was
             * observed (oracle/probe_scipy_dop853.py) to shrink a rejected step by
             * exactly 1/facc1 = 0.3 whatever err is; the oracle follows SciPy. */
            hnew = h / facc1;
            reject = 1;
            if (naccpt >= 1) st->nrejct++;
            last = 0;
        }
        h = hnew;
    }
}

/* RateStateModel.evaluate(), :357-389 (noise draw of :392 is the caller's) */
int orc_forward(const orc_model *m, double *t_out, double *mu_out,
                double *theta_out, double *vel_out, double *acc_out,
                orc_stats *st)
{
    orc_stats local;
    if (!st) st = &local;
    memset(st, 0, sizeof(*st));
    const double delta_t = (m->t_final - m->t_start) / m->num_tsteps;   /* :177 */
    const int num_steps = (int)floor((m->t_final - m->t_start) / delta_t);  /* :358 */
    if (num_steps <= 0) return 0;
    if (t_out) memset(t_out, 0, sizeof(double) * num_steps);
    if (mu_out) memset(mu_out, 0, sizeof(double) * num_steps);
    if (theta_out) memset(theta_out, 0, sizeof(double) * num_steps);
    if (vel_out) memset(vel_out, 0, sizeof(double) * num_steps);
    if (acc_out) memset(acc_out, 0, sizeof(double) * num_steps);

    double t = m->t_start;
    double y[3] = { m->mu_t_zero, m->Dc / m->V_ref, m->V_ref };          /* :367-370,377 */
    if (t_out) t_out[0] = t;
    if (mu_out) mu_out[0] = m->mu_ref;
    if (theta_out) theta_out[0] = y[1];
    if (vel_out) vel_out[0] = y[2];
    double v_prev = y[2];
    int ok = 1, k = 1;
    st->istate = 1;
    while (ok && k < num_steps) {                                        /* :381 */
        int idid = orc_dop853_call(m, &t, y, t + delta_t, st);          /* :382 */
        ok = (idid > 0);
        if (t_out) t_out[k] = t;
        if (mu_out) mu_out[k] = y[0];
        if (theta_out) theta_out[k] = y[1];
        if (vel_out) vel_out[k] = y[2];
        if (acc_out) acc_out[k] = (y[2] - v_prev) / delta_t;             /* :388 */
        v_prev = y[2];
        k++;
    }
    st->filled = k;
    return num_steps;
}

/* the series the sampler scores: model.evaluate()[1] (MCMC.py:127) -- acc for the reference, mu for the
 * friction-series extension (mu[0] = mu_ref, :367; mu[k] = r.y[0], :385) */
static int orc_observe(const orc_model *m, double *obs, orc_stats *st)
{
    if (m->observable == ORC_OBS_MU) return orc_forward(m, NULL, obs, NULL, NULL, NULL, st);
    return orc_forward(m, NULL, NULL, NULL, NULL, obs, st);
}

/* sum((acc - data)**2) with numpy's pairwise reduction (blocks of 128, eight
 * partial sums), MCMC.py:387 */
static double pairwise_sq(const double *a, const double *d, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) { double e = a[i] - d[i]; res += e * e; }
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; j++) { double e = a[j] - d[j]; r[j] = e * e; }
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) { double e = a[i + j] - d[i + j]; r[j] += e * e; }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) { double e = a[i] - d[i]; res += e * e; }
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sq(a, d, n2) + pairwise_sq(a + n2, d + n2, n - n2);
    }
}

double orc_sse(const double *acc, const double *data, int n)
{
    return pairwise_sq(acc, data, n);
}

/* the scalar the sampler varies: model.Dc in the reference (MCMC.py:245, 251, 381), k1 with ORC_PARAM_K1 */
static void set_scalar(orc_model *m, double q)
{
    if (m->sampled_param == ORC_PARAM_K1) m->k1 = q; else m->Dc = q;
}
static double get_scalar(const orc_model *m)
{
    return m->sampled_param == ORC_PARAM_K1 ? m->k1 : m->Dc;
}

/* batch of independent solves over pthreads (gcc here has no libgomp) */
typedef struct {
    const orc_model *m; const double *dc; int C; const double *data; int n;
    double *sse_out, *acc_out; int64_t *nrhs_out;
    atomic_int next; atomic_int bad;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    double *scratch = j->acc_out ? NULL : (double *)malloc(sizeof(double) * j->n);
    for (;;) {
        int c = atomic_fetch_add(&j->next, 1);
        if (c >= j->C) break;
        orc_model mm = *j->m;
        set_scalar(&mm, j->dc[c]);
        orc_stats st;
        double *acc = j->acc_out ? j->acc_out + (size_t)c * j->n : scratch;
        int ns = orc_observe(&mm, acc, &st);
        if (ns != j->n) { atomic_fetch_add(&j->bad, 1); continue; }
        if (j->sse_out && j->data) j->sse_out[c] = orc_sse(acc, j->data, j->n);
        if (j->nrhs_out) j->nrhs_out[c] = st.nrhs;
    }
    free(scratch);
    return NULL;
}

int orc_forward_batch(const orc_model *m, const double *dc, int C,
                      const double *data, int n, double *sse_out,
                      double *acc_out, int64_t *nrhs_out, int nthreads)
{
    batch_job job = { m, dc, C, data, n, sse_out, acc_out, nrhs_out, 0, 0 };
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, batch_worker, &job);
    batch_worker(&job);
    for (int i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
    return atomic_load(&job.bad) ? -1 : 0;
}

/* np.cov of a 1-D window (ddof = 1): mean first, then sum of squared deviations */
static double window_var(const double *x, int n)
{
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean += x[i];
    mean /= n;
    double s = 0.0;
    for (int i = 0; i < n; i++) { double d = x[i] - mean; s += d * d; }
    return s / (n - 1);
}

/* MCMC.sample() under host-supplied randomness; SURVEY.md Appendix A */
int orc_chain_replay(const orc_model *m0, const double *data, int n,
                     double qstart, double lo, double hi, int n_prior_len,
                     int nsamples, int adapt_interval, int compat_adapt,
                     const double *proposals, const double *uniforms,
                     const double *gammas, double *chain, double *s2,
                     uint8_t *accept, double *vstart_out, int64_t *nsolves)
{
    orc_model m = *m0;
    const double n0 = 0.01;                                   /* MCMC.py:97 */
    double *acc = (double *)malloc(sizeof(double) * n);
    double *acc_dq = (double *)malloc(sizeof(double) * n);
    int64_t solves = 0;
    /* compute_initial_covariance, MCMC.py:245-266 */
    set_scalar(&m, qstart);
    if (orc_observe(&m, acc, 0) != n) { free(acc); free(acc_dq); return -1; }
    set_scalar(&m, get_scalar(&m) * (1 + 1e-6));
    orc_observe(&m, acc_dq, 0);
    solves += 2;
    s2[0] = orc_sse(acc, data, n) / (n - n_prior_len);        /* :261 */
    double xtx = 0.0;
    {
        /* X'X via np.dot on an (1,N)x(N,1) product; plain accumulation here */
        double den = get_scalar(&m) * 1e-6;
        for (int i = 0; i < n; i++) { double x = (acc_dq[i] - acc[i]) / den; xtx += x * x; }
    }
    double V = s2[0] * (1.0 / xtx);                           /* :265-266 */
    if (vstart_out) *vstart_out = V;
    /* SSqprev = SSqcalc(qstart), :468 */
    set_scalar(&m, qstart);
    orc_observe(&m, acc, 0);
    solves++;
    double ss = orc_sse(acc, data, n);
    double q = qstart;
    chain[0] = q;
    for (int i = 0; i < nsamples; i++) {
        double qn = compat_adapt ? q + sqrt(V) * proposals[i] : proposals[i];   /* :497 */
        int ok = (qn > lo) && (qn < hi);                      /* :318-320, strict */
        if (ok) {
            set_scalar(&m, qn);
            orc_observe(&m, acc, 0);              /* :324 */
            solves++;
            double ssn = orc_sse(acc, data, n);
            double la = 0.5 * (ss - ssn) / s2[i];             /* :327 */
            if (la > 0.0) la = 0.0;
            ok = la > log(uniforms[i]);                       /* :331 */
            if (ok) { q = qn; ss = ssn; }
        }
        accept[i] = (uint8_t)ok;
        chain[i + 1] = q;                                     /* :507-517 */
        /* update_standard_deviation :158-160.  gammas[i] is the unit-scale
         * Gamma(aval) draw; scipy's rvs scales it by `scale = 1/bval`. */
        {
            double bval = 0.5 * (n0 * s2[i] + ss);
            double scale = 1 / bval;
            s2[i + 1] = 1 / (gammas[i] * scale);
        }
        if (compat_adapt && (i + 1) % adapt_interval == 0) {  /* :523-527, 200-204 */
            int w = adapt_interval;
            if (i + 2 >= w) {
                double var = window_var(chain + (i + 2 - w), w);
                double vnew = 2.38 * 2.38 / 2.0 * var;        /* len(qpriors.keys()) = 2 */
                if (vnew > 0.0) V = sqrt(vnew);               /* cholesky of 1x1, used as a covariance */
            }
        }
    }
    if (nsolves) *nsolves = solves;
    free(acc); free(acc_dq);
    return 0;
}

/* d-parameter replay with absolute proposals: q = (Dc) or (a, b, Dc) */
static void set_params(orc_model *m, int d, const double *q)
{
    if (d == 3) { m->a = q[0]; m->b = q[1]; m->Dc = q[2]; }
    else set_scalar(m, q[0]);
}

int orc_chain_replay_nd(const orc_model *m0, const double *data, int n, int d,
                        const double *qstart, const double *lo, const double *hi,
                        int n_prior_len, int nsamples, const double *proposals,
                        const double *uniforms, const double *gammas, double *chain,
                        double *s2, uint8_t *accept, int64_t *nsolves)
{
    if (d != 1 && d != 3) return -2;
    orc_model m = *m0;
    const double n0 = 0.01;                                   /* MCMC.py:97 */
    double *obs = (double *)malloc(sizeof(double) * n);
    int64_t solves = 0;
    double q[3];
    for (int j = 0; j < d; j++) q[j] = qstart[j];
    set_params(&m, d, q);
    if (orc_observe(&m, obs, 0) != n) { free(obs); return -1; }
    solves++;
    double ss = orc_sse(obs, data, n);                        /* :468 */
    s2[0] = ss / (n - n_prior_len);                           /* :261 */
    for (int j = 0; j < d; j++) chain[j] = q[j];
    for (int i = 0; i < nsamples; i++) {
        const double *qn = proposals + (size_t)i * d;
        int ok = 1;
        for (int j = 0; j < d; j++) ok = ok && (qn[j] > lo[j]) && (qn[j] < hi[j]);   /* :318-320, strict */
        if (ok) {
            set_params(&m, d, qn);
            orc_observe(&m, obs, 0);                          /* :324 */
            solves++;
            double ssn = orc_sse(obs, data, n);
            double la = 0.5 * (ss - ssn) / s2[i];             /* :327 */
            if (la > 0.0) la = 0.0;
            ok = la > log(uniforms[i]);                       /* :331 */
            if (ok) { for (int j = 0; j < d; j++) q[j] = qn[j]; ss = ssn; }
        }
        accept[i] = (uint8_t)ok;
        for (int j = 0; j < d; j++) chain[(size_t)(i + 1) * d + j] = q[j];   /* :507-517 */
        {
            double bval = 0.5 * (n0 * s2[i] + ss);            /* :158-160 */
            double scale = 1 / bval;
            s2[i + 1] = 1 / (gammas[i] * scale);
        }
    }
    if (nsolves) *nsolves = solves;
    free(obs);
    return 0;
}

/* Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- samplers on Philox (stream layout: rsf_oracle.h) ---- */
static double philox_u53(uint32_t hi, uint32_t lo)
{
    const uint64_t k = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)k + 0.5) * 0x1.0p-53;
}

static void philox_block(uint64_t seed, uint64_t chain, uint32_t iter, uint32_t slot, uint32_t out[4])
{
    const uint32_t ctr[4] = { (uint32_t)chain, (uint32_t)(chain >> 32), iter, slot };
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    orc_philox4x32_10(ctr, key, out);
}

/* sin(pi x), cos(pi x) for x in [0, 2): exact reduction to |r| <= 1/4 of a quadrant, then libm */
static void sincospi_02(double x, double *s, double *c)
{
    const double q = floor(2.0 * x + 0.5);            /* nearest multiple of 1/2 */
    const double r = x - 0.5 * q;                     /* exact */
    const double sr = sin(M_PI * r), cr = cos(M_PI * r);
    switch (((int)q) & 3) {
        case 0: *s = sr;  *c = cr;  break;
        case 1: *s = cr;  *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}

static void philox_normal2(uint64_t seed, uint64_t chain, uint32_t iter, uint32_t slot, double *z0, double *z1)
{
    uint32_t r[4];
    philox_block(seed, chain, iter, slot, r);
    const double u1 = philox_u53(r[0], r[1]), u2 = philox_u53(r[2], r[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi_02(2.0 * u2, &s, &c);
    *z0 = rad * c;
    *z1 = rad * s;
}

void orc_philox_draws(uint64_t seed, uint64_t chain, uint32_t iter, double shape, double out[6])
{
    double zz;
    uint32_t r[4];
    philox_normal2(seed, chain, iter, 0u, &out[0], &out[1]);
    philox_normal2(seed, chain, iter, 1u, &out[2], &zz);
    philox_block(seed, chain, iter, 2u, r);
    out[3] = philox_u53(r[0], r[1]);
    /* Marsaglia & Tsang (2000), shape > 1: the algorithm of NumPy's legacy standard_gamma */
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    out[4] = d; out[5] = 64.0;
    for (uint32_t j = 0; j < 64; j++) {
        double x, unused;
        philox_normal2(seed, chain, iter, 4u + 2u * j, &x, &unused);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        philox_block(seed, chain, iter, 5u + 2u * j, r);
        const double u = philox_u53(r[0], r[1]);
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) {
            out[4] = d * v; out[5] = (double)(j + 1);
            return;
        }
    }
}
