/*
 * clair_oracle.c — plain C (OpenMP) CPU restatement of the per-pixel radiometric hot path of
 * samivout/clair-torch.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the fast twin of oracle/clair_oracle.py: the same closed forms,
 * per pixel, so that parity tests can run at the BASELINE sizes and so that bench.py has a multi-threaded
 * CPU implementation of the path to time (`cpu_baseline`, `--impl reference`).  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the product never does.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks every entry point against the fixtures the
 * unmodified reference produced (tests/golden/, tests/golden/make_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off; contraction is disabled because LUT
 * indices and fp32 table values must round exactly like the reference's separate ATen ops).
 *
 * Citations are file:line in the reference repository.  Layout: stacks are (N, C, P) with P = H*W fastest.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HDR_SCALE 30.0   /* training/losses.py:193 default, inference/hdr_merge.py:95 */
#define PAIR_SCALE 10.0f /* training/losses.py:212 default */

/* Row of the (C, L) table used by element (c, p) in LINEAR mode: flat NCHW index mod C
 * (models/base.py:173-176); n*C*P is a multiple of C so the frame index drops out. */
static inline int curve_row(int64_t c, int64_t p, int64_t plane, int64_t flat_offset, int C) {
    return (int)((c * plane + p + flat_offset) % C);
}

typedef struct {
    float f;   /* models/base.py:182 */
    float fp;  /* autograd derivative: (g1-g0)*(L-1) on the closed clamp interval, else 0 */
    float w;
    int x0;
} tap_t;

static inline tap_t icrf_linear(float x, const float *row, int L) {
    tap_t t;
    const float lm1 = (float)(L - 1);
    const float xs_raw = x * lm1;                                   /* :167 */
    float xs = xs_raw < 0.0f ? 0.0f : xs_raw;
    xs = xs > lm1 ? lm1 : xs;
    const float fl = floorf(xs);                                    /* :169 */
    t.x0 = (int)fl;
    const int x1 = t.x0 + 1 > L - 1 ? L - 1 : t.x0 + 1;            /* :170 */
    t.w = xs - fl;                                                  /* :171 */
    const float g0 = row[t.x0], g1 = row[x1];
    const float a = g0 * (1.0f - t.w);
    const float b = g1 * t.w;
    t.f = a + b;                                                    /* :182 */
    t.fp = (xs_raw >= 0.0f && xs_raw <= lm1) ? (g1 - g0) * lm1 : 0.0f;
    return t;
}

/* ---------------------------------------------------------------------------------------------------- */
void oracle_icrf_linear(const float *x, const float *theta, int n_frames, int C, int64_t plane, int L,
                        int64_t flat_offset, float *f_out, float *fp_out, int32_t *x0_out) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < (int64_t)n_frames * C; ++s) {
        const int64_t c = s % C;
        for (int64_t p = 0; p < plane; ++p) {
            const int u = curve_row(c, p, plane, flat_offset, C);
            const tap_t t = icrf_linear(x[s * plane + p], theta + (int64_t)u * L, L);
            f_out[s * plane + p] = t.f;
            if (fp_out) fp_out[s * plane + p] = t.fp;
            if (x0_out) x0_out[s * plane + p] = t.x0;
        }
    }
}

/* LOOKUP mode, models/base.py:138-158: round-half-even, true channel */
void oracle_icrf_lookup(const float *x, const float *theta, int n_frames, int C, int64_t plane, int L, float *y_out,
                        int32_t *idx_out) {
    const float lm1 = (float)(L - 1);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < (int64_t)n_frames * C; ++s) {
        const int64_t c = s % C;
        for (int64_t p = 0; p < plane; ++p) {
            float r = rintf(x[s * plane + p] * lm1);
            r = r < 0.0f ? 0.0f : (r > lm1 ? lm1 : r);
            const int k = (int)r;
            y_out[s * plane + p] = theta[c * L + k];
            if (idx_out) idx_out[s * plane + p] = k;
        }
    }
}

/* inference/linearization.py:94-106,132 */
void oracle_linearize(const float *val, const float *std, const float *theta, int n_frames, int C, int64_t plane, int L,
                      int64_t flat_offset, float *lin_out, float *sigma_out) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < (int64_t)n_frames * C; ++s) {
        const int64_t c = s % C;
        for (int64_t p = 0; p < plane; ++p) {
            const int u = curve_row(c, p, plane, flat_offset, C);
            const tap_t t = icrf_linear(val[s * plane + p], theta + (int64_t)u * L, L);
            lin_out[s * plane + p] = t.f;
            float g = std ? t.fp * std[s * plane + p] : 0.0f;
            g = g * g;
            sigma_out[s * plane + p] = sqrtf(g);
        }
    }
}

/* ----------------------------------------------------------------------------------------------------
 * compute_hdr_image, inference/hdr_merge.py:58-128,155, batches of `batch_size` frames.
 * Per batch (common/statistics.py:74-109 and the closed form of the autograd pass):
 *   w_n = exp(-30 (x_n-.5)^2) | 1,  v_n = f(x_n)/t_n,  W_B = sum w_n,  mean_B = sum w_n v_n/(W_B+1e-6)
 *   W = W_A + W_B,  mean = mean_A + (W_B/W)(mean_B - mean_A)
 *   g_n = (W_B/W)[w_n f'_n/t_n + w'_n (v_n-mean_B)]/(W_B+1e-6) + w'_n (W_A/W^2)(mean_B-mean_A)
 *   var += sum_n (g_n s_n)^2
 * radiance_out, sigma_out: (C, plane) float64; sigma_out may be NULL (then std must be NULL too).
 */
void oracle_hdr_merge(const float *val, const float *std, const double *exposure, int n_frames, int C, int64_t plane,
                      const float *theta, int L, int gaussian, int batch_size, int64_t flat_offset,
                      double *radiance_out, double *sigma_out) {
    if (batch_size <= 0) batch_size = n_frames;
    const int64_t frame_stride = (int64_t)C * plane;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < frame_stride; ++e) {
        const int64_t c = e / plane, p = e % plane;
        const int u = theta ? curve_row(c, p, plane, flat_offset, C) : 0;
        double mean_a = 0.0, w_a = 0.0, var = 0.0;
        double w[64], wp[64], v[64], fpt[64];
        for (int n0 = 0; n0 < n_frames; n0 += batch_size) {
            const int nb = n_frames - n0 < batch_size ? n_frames - n0 : batch_size;
            double w_b = 0.0, s_b = 0.0;
            for (int k = 0; k < nb; ++k) {
                const float x = val[(int64_t)(n0 + k) * frame_stride + e];
                double f = x, fp = 1.0;
                if (theta) {
                    const tap_t t = icrf_linear(x, theta + (int64_t)u * L, L);
                    f = t.f;
                    fp = t.fp;
                }
                if (gaussian) {
                    const float d = x - 0.5f;
                    const float wf = expf(-30.0f * (d * d));          /* losses.py:205 in fp32 */
                    w[k] = wf;
                    wp[k] = -2.0 * HDR_SCALE * ((double)x - 0.5) * wf;
                } else {
                    w[k] = 1.0;
                    wp[k] = 0.0;
                }
                v[k] = f / exposure[n0 + k];
                fpt[k] = fp / exposure[n0 + k];
                w_b += w[k];
                s_b += w[k] * v[k];
            }
            const double mean_b = s_b / (w_b + 1e-6);
            const double w_tot = w_a + w_b;
            const double frac = w_b / w_tot;
            if (std) {
                for (int k = 0; k < nb; ++k) {
                    const double s = std[(int64_t)(n0 + k) * frame_stride + e];
                    const double dmb = (w[k] * fpt[k] + wp[k] * (v[k] - mean_b)) / (w_b + 1e-6);
                    const double g = frac * dmb + wp[k] * (w_a / (w_tot * w_tot)) * (mean_b - mean_a);
                    var += (g * s) * (g * s);
                }
            }
            mean_a = mean_a + frac * (mean_b - mean_a);
            w_a = w_tot;
        }
        radiance_out[e] = mean_a;
        if (sigma_out) sigma_out[e] = sqrt(var);
    }
}

/* ----------------------------------------------------------------------------------------------------
 * Pair statistics: measure_linearity (inference/measure_linearity.py:41-74) and the forward half of a
 * train_icrf step (training/icrf_training.py:105-136).  Two passes like the reference's
 * weighted_mean_and_std (common/general_functions.py:153-164): means first, then weighted squared deviations.
 * Outputs (P, C) float64: mean, stddev, errmean (errmean untouched when std == NULL).
 */
typedef struct {
    double ell, err, wt;
    int valid;
    /* pieces the gradient needs */
    double a, b, es, bs, sa, sb, sgn;
} pair_elem_t;

typedef struct {
    float f, fp, sig, gw, x;
    int x0, valid;
    float w;
} frame_elem_t;

static inline frame_elem_t frame_elem(float x, float s, const float *row, int L, int has_std, float lo, float hi) {
    frame_elem_t o;
    o.x = x;
    if (row) {
        const tap_t t = icrf_linear(x, row, L);
        o.f = t.f; o.fp = t.fp; o.x0 = t.x0; o.w = t.w;
    } else {
        o.f = x; o.fp = 1.0f; o.x0 = 0; o.w = 0.0f;
    }
    o.sig = has_std ? fabsf(o.fp * s) : 0.0f;                         /* icrf_training.py:124 */
    const float d = x - 0.5f;
    o.gw = expf(-PAIR_SCALE * (d * d));                               /* losses.py:229-230 */
    o.valid = (x >= lo) && (x <= hi);                                 /* general_functions.py:305 */
    return o;
}

static inline pair_elem_t pair_elem(const frame_elem_t *fi, const frame_elem_t *fj, double r, int relative, int has_std,
                                    int unc) {
    pair_elem_t o;
    o.a = fi->f; o.b = fj->f;
    const double e = o.b * r;                                         /* losses.py:40 */
    const double d = o.a - e;
    o.es = e + 1e-6;                                                  /* :45 */
    const double q = relative ? d / o.es : d;
    o.ell = fabs(q);
    o.sgn = (q > 0) - (q < 0);
    o.err = 0.0;
    o.sa = fi->sig; o.sb = fj->sig;
    o.bs = fj->f > 1e-6f ? fj->f : 1e-6f;                             /* :55 clamp in fp32 */
    if (has_std) {
        if (relative) {
            const double num = (double)(fi->f * fj->sig);             /* :58 fp32 product */
            const double t1 = o.sa / o.es, t2 = num / (o.es * o.bs);
            o.err = sqrt(t1 * t1 + t2 * t2 + 1e-6);
        } else {
            const double s2 = (double)(fi->sig * fi->sig);            /* :62 std_i ** 2 in fp32 */
            o.err = sqrt(s2 + (r * o.sb) * (r * o.sb));
        }
    }
    o.wt = (double)(fi->gw + fj->gw);                                 /* losses.py:234 fp32 add */
    if (has_std && unc) o.wt += 1.0 / (o.err + 1e-6);                 /* losses.py:97 */
    o.valid = fi->valid && fj->valid;
    return o;
}

void oracle_pair_stats(const float *val, const float *std, int n_frames, int C, int64_t plane, const int32_t *pi,
                       const int32_t *pj, const double *pr, int P, const float *theta, int L, int64_t flat_offset,
                       float lo, float hi, int relative, int unc, double *mean_out, double *std_out,
                       double *errmean_out) {
    const int has_std = std != NULL;
    const int64_t frame_stride = (int64_t)C * plane;
    const int64_t PC = (int64_t)P * C;
    double *s0 = calloc(PC, sizeof(double)), *s1 = calloc(PC, sizeof(double)), *s3 = calloc(PC, sizeof(double));
    double *s4 = calloc(PC, sizeof(double)), *m2 = calloc(PC, sizeof(double));
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel
        {
            double *t0 = calloc(PC, sizeof(double)), *t1 = calloc(PC, sizeof(double)), *t3 = calloc(PC, sizeof(double));
            double *t4 = calloc(PC, sizeof(double));
            frame_elem_t fe[64];
#pragma omp for schedule(static)
            for (int64_t e = 0; e < frame_stride; ++e) {
                const int64_t c = e / plane, p = e % plane;
                const float *row = theta ? theta + (int64_t)curve_row(c, p, plane, flat_offset, C) * L : NULL;
                for (int n = 0; n < n_frames; ++n)
                    fe[n] = frame_elem(val[n * frame_stride + e], has_std ? std[n * frame_stride + e] : 0.0f, row, L, has_std,
                                       lo, hi);
                for (int k = 0; k < P; ++k) {
                    const pair_elem_t q = pair_elem(&fe[pi[k]], &fe[pj[k]], pr[k], relative, has_std, unc);
                    if (!q.valid) continue;
                    const int64_t o = (int64_t)k * C + c;
                    if (pass == 0) {
                        t0[o] += q.wt; t1[o] += q.wt * q.ell; t3[o] += q.err; t4[o] += 1.0;
                    } else {
                        const double dv = q.ell - mean_out[o];
                        t0[o] += q.wt * dv * dv;
                    }
                }
            }
#pragma omp critical
            for (int64_t o = 0; o < PC; ++o) {
                if (pass == 0) { s0[o] += t0[o]; s1[o] += t1[o]; s3[o] += t3[o]; s4[o] += t4[o]; }
                else m2[o] += t0[o];
            }
            free(t0); free(t1); free(t3); free(t4);
        }
        if (pass == 0) {
            for (int64_t o = 0; o < PC; ++o) {
                const double den = s0[o] > 1e-8 ? s0[o] : 1e-8;       /* general_functions.py:156 */
                mean_out[o] = s1[o] / den;
                if (errmean_out && has_std) errmean_out[o] = s3[o] / (s4[o] > 1e-8 ? s4[o] : 1e-8);
            }
            if (!std_out) break;
        } else {
            for (int64_t o = 0; o < PC; ++o) {
                const double den = s0[o] > 1e-8 ? s0[o] : 1e-8;
                /* masked-out pixels enter the reference's sum as (0 - mean)^2 * 0 = 0 */
                std_out[o] = sqrt(m2[o] / den);
            }
        }
    }
    free(s0); free(s1); free(s3); free(s4); free(m2);
}

/* ----------------------------------------------------------------------------------------------------
 * d(sum_c sqrt(sum_p mean[p,c]^2)) / d theta for one train_icrf step (training/icrf_training.py:136,148-149),
 * closed form.  linloss_out (C), mean_out (P,C), grad_out (C,L) float64.
 */
void oracle_train_grad(const float *val, const float *std, int n_frames, int C, int64_t plane, const int32_t *pi,
                       const int32_t *pj, const double *pr, int P, const float *theta, int L, int64_t flat_offset,
                       float lo, float hi, int relative, int unc, double *linloss_out, double *mean_out,
                       double *grad_out) {
    const int has_std = std != NULL;
    const int64_t frame_stride = (int64_t)C * plane;
    const int64_t PC = (int64_t)P * C;
    oracle_pair_stats(val, std, n_frames, C, plane, pi, pj, pr, P, theta, L, flat_offset, lo, hi, relative, unc, mean_out,
                      NULL, NULL);
    /* denominators again (cheap second sweep keeps oracle_pair_stats' signature simple) */
    double *den = calloc(PC, sizeof(double));
#pragma omp parallel
    {
        double *t0 = calloc(PC, sizeof(double));
        frame_elem_t fe[64];
#pragma omp for schedule(static)
        for (int64_t e = 0; e < frame_stride; ++e) {
            const int64_t c = e / plane, p = e % plane;
            const float *row = theta + (int64_t)curve_row(c, p, plane, flat_offset, C) * L;
            for (int n = 0; n < n_frames; ++n)
                fe[n] = frame_elem(val[n * frame_stride + e], has_std ? std[n * frame_stride + e] : 0.0f, row, L, has_std, lo, hi);
            for (int k = 0; k < P; ++k) {
                const pair_elem_t q = pair_elem(&fe[pi[k]], &fe[pj[k]], pr[k], relative, has_std, unc);
                if (q.valid) t0[(int64_t)k * C + c] += q.wt;
            }
        }
#pragma omp critical
        for (int64_t o = 0; o < PC; ++o) den[o] += t0[o];
        free(t0);
    }
    double *up = calloc(PC, sizeof(double));
    for (int c = 0; c < C; ++c) {
        double acc = 0.0;
        for (int k = 0; k < P; ++k) acc += mean_out[(int64_t)k * C + c] * mean_out[(int64_t)k * C + c];
        linloss_out[c] = sqrt(acc);
        for (int k = 0; k < P; ++k) {
            const int64_t o = (int64_t)k * C + c;
            const double d = den[o] > 1e-8 ? den[o] : 1e-8;
            up[o] = linloss_out[c] > 0 ? mean_out[o] / linloss_out[c] / d : 0.0;
        }
    }
    memset(grad_out, 0, sizeof(double) * C * L);
#pragma omp parallel
    {
        double *g_loc = calloc((size_t)C * L, sizeof(double));
        frame_elem_t fe[64];
        double g_frame[64];
#pragma omp for schedule(static)
        for (int64_t e = 0; e < frame_stride; ++e) {
            const int64_t c = e / plane, p = e % plane;
            const int u = curve_row(c, p, plane, flat_offset, C);
            const float *row = theta + (int64_t)u * L;
            for (int n = 0; n < n_frames; ++n) {
                fe[n] = frame_elem(val[n * frame_stride + e], has_std ? std[n * frame_stride + e] : 0.0f, row, L, has_std, lo, hi);
                g_frame[n] = 0.0;
            }
            for (int k = 0; k < P; ++k) {
                const pair_elem_t q = pair_elem(&fe[pi[k]], &fe[pj[k]], pr[k], relative, has_std, unc);
                if (!q.valid) continue;
                const int64_t o = (int64_t)k * C + c;
                const double r = pr[k], U = up[o];
                double ga, gb;
                if (relative) {
                    ga = q.wt * U * q.sgn / q.es;
                    gb = -q.wt * U * q.sgn * r * (q.a + 1e-6) / (q.es * q.es);
                    if (has_std && unc) {
                        const int clamped = den[o] < 1e-8;
                        const double dm_dwt = (q.ell - (clamped ? 0.0 : mean_out[o])) * U;
                        const double dwt_dt = -1.0 / ((q.err + 1e-6) * (q.err + 1e-6)) / (2.0 * q.err);
                        const double esbs = q.es * q.bs;
                        const double dt_da = 2.0 * q.a * q.sb * q.sb / (esbs * esbs);
                        const double dt_des = -2.0 * q.sa * q.sa / (q.es * q.es * q.es)
                                              - 2.0 * q.a * q.a * q.sb * q.sb / (q.es * q.es * q.es * q.bs * q.bs);
                        const double dt_dbs = -2.0 * q.a * q.a * q.sb * q.sb / (q.es * q.es * q.bs * q.bs * q.bs);
                        const double dbs_db = fe[pj[k]].f >= 1e-6f ? 1.0 : 0.0;
                        ga += dm_dwt * dwt_dt * dt_da;
                        gb += dm_dwt * dwt_dt * (dt_des * r + dt_dbs * dbs_db);
                    }
                } else {
                    ga = q.wt * U * q.sgn;
                    gb = -q.wt * U * q.sgn * r;
                }
                g_frame[pi[k]] += ga;
                g_frame[pj[k]] += gb;
            }
            for (int n = 0; n < n_frames; ++n) {
                if (g_frame[n] == 0.0) continue;
                const int x1 = fe[n].x0 + 1 > L - 1 ? L - 1 : fe[n].x0 + 1;
                g_loc[(int64_t)u * L + fe[n].x0] += g_frame[n] * (1.0 - (double)fe[n].w);
                g_loc[(int64_t)u * L + x1] += g_frame[n] * (double)fe[n].w;
            }
        }
#pragma omp critical
        for (int64_t o = 0; o < (int64_t)C * L; ++o) grad_out[o] += g_loc[o];
        free(g_loc);
    }
    free(den); free(up);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torch.distributed.run exports OMP_NUM_THREADS=1: bench.py's CPU legs ask for the host's cores explicitly */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
