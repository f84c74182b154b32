/*
 * oracle.c -- CPU restatement of DI-hpc's trajectory-return hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker* for the CUDA product path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (di_hpc_b200/) never links, imports or falls back to anything in oracle/.
 *
 * Every function restates, in plain C, the algorithm of the reference's pure-PyTorch oracle
 * `hpc_rll/origin/{gae,td,vtrace,upgo,ppo}.py` (cited per function as file:line relative to /root/reference), which is
 * what the reference's own tests compare its CUDA kernels against (tests/test_gae.py:25-26 ...).
 * Backward passes are closed-form adjoints of the same expressions that
 * origin differentiates with autograd.
 *
 * Parity pin: the reference holds NO golden vectors for this path; the restatement is pinned to
 * outputs of the reference itself, run in the build container and committed as
 * the tests/golden/ fixtures (.npz) by tests/golden/make_golden.py (tests/test_oracle_golden.py checks them).
 *
 * Compiled twice by oracle/Makefile:
 *   -DREAL=float   liboracle_f32.so  -- element-wise arithmetic in fp32, in origin's operation
 *                                       order, no FMA contraction (-ffp-contract=off), so
 *                                       element-wise outputs (e.g. GAE adv) are bit-identical to
 *                                       origin on CPU; reductions accumulate in double.
 *   -DREAL=double  liboracle_f64.so  -- fp64 "truth".
 * Loops over the independent batch axis are OpenMP-parallel (this is also the CPU baseline).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#ifdef ORACLE_F64
#define FN(name) CAT(name, _f64)
#define R_EXP exp
#define R_LOG log
#define R_SQRT sqrt
#define R_FABS fabs
#define R_FLOOR floor
#define R_CEIL ceil
#else
#define FN(name) CAT(name, _f32)
#define R_EXP expf
#define R_LOG logf
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_FLOOR floorf
#define R_CEIL ceilf
#endif

typedef REAL real;

/* ------------------------------------------------------------------------------------------
 * helpers
 * ---------------------------------------------------------------------------------------- */

/* torch.distributions.Categorical(logits=x): normalised logits x - logsumexp(x)
 * (used by origin/vtrace.py:74,107-109 and origin/ppo.py:54-57). Writes logp[N], returns
 * entropy -sum p*logp  (probs = softmax). */
static real categorical_row(const real* x, int64_t N, real* logp, real* prob) {
    real m = x[0];
    for (int64_t k = 1; k < N; ++k) m = x[k] > m ? x[k] : m;
    double s = 0.0;
    for (int64_t k = 0; k < N; ++k) s += (double)R_EXP(x[k] - m);
    real lse = m + R_LOG((real)s);
    double h = 0.0;
    for (int64_t k = 0; k < N; ++k) {
        logp[k] = x[k] - lse;
        prob[k] = R_EXP(logp[k]);
        h += (double)(logp[k] * prob[k]);
    }
    return (real)(-h);
}

/* F.cross_entropy / log_softmax form: (x - max) - log(sum exp(x - max))  (origin/upgo.py:16) */
static void log_softmax_row(const real* x, int64_t N, real* logp) {
    real m = x[0];
    for (int64_t k = 1; k < N; ++k) m = x[k] > m ? x[k] : m;
    double s = 0.0;
    for (int64_t k = 0; k < N; ++k) s += (double)R_EXP(x[k] - m);
    real ls = R_LOG((real)s);
    for (int64_t k = 0; k < N; ++k) logp[k] = (x[k] - m) - ls;
}

/* n-step discounted reward  R_b = sum_i factor_i * r[i,b],  factor_0 = 1, factor_i = gamma*factor_{i-1}
 * built in the tensor dtype (origin/td.py:349-352: reward_factor tensor, then matmul). */
static real nstep_reward(const real* reward, int64_t T, int64_t B, int64_t b, real gamma_r) {
    real factor = (real)1;
    real acc = (real)0;
    for (int64_t i = 0; i < T; ++i) {
        acc = acc + factor * reward[i * B + b];
        factor = gamma_r * factor;
    }
    return acc;
}

static real sgn(real x) { return (real)((x > 0) - (x < 0)); }

/* origin/td.py:9-14 */
static real value_transform(real x, real eps) { return sgn(x) * (R_SQRT(R_FABS(x) + 1) - 1) + eps * x; }
/* origin/td.py:17-22 */
static real value_inv_transform(real x, real eps) {
    real t = (R_SQRT(1 + 4 * eps * (R_FABS(x) + 1 + eps)) - 1) / (2 * eps);
    return sgn(x) * (t * t - 1);
}

/* ------------------------------------------------------------------------------------------
 * GAE  -- origin/gae.py:28-37.  value (T+1,B), reward (T,B), adv (T,B).
 *   delta = reward + gamma*value[1:] - value[:-1]          (gae.py:29)
 *   denom = 1 + lambda*denom   (Python double, gae.py:34)  -> cast to tensor dtype when used
 *   gae_item = denom*delta[t] + factor*gae_item            (gae.py:35), factor=gamma*lambda (double)
 *   adv[t] += gae_item/denom                               (gae.py:36)
 * ---------------------------------------------------------------------------------------- */
static void gae_denoms(int64_t T, double lambda, real* d) {
    double den = 0.0;
    for (int64_t t = T - 1; t >= 0; --t) {
        den = 1.0 + lambda * den;
        d[t] = (real)den;
    }
}

/* column range [b0,b1) of thread `tid` of `nt`: contiguous, 16-float aligned (streams well, vectorises) */
static void col_range(int64_t B, int tid, int nt, int64_t* b0, int64_t* b1) {
    int64_t chunk = ((B + nt - 1) / nt + 15) / 16 * 16;
    *b0 = (int64_t)tid * chunk;
    *b1 = *b0 + chunk;
    if (*b0 > B) *b0 = B;
    if (*b1 > B) *b1 = B;
}

/* parallel first-touch copy with the same thread->column mapping as the kernels below (NUMA placement
 * for the CPU-baseline timing; not part of the algorithm) */
void FN(orc_place_copy)(real* dst, const real* src, int64_t rows, int64_t B) {
#pragma omp parallel
    {
        int64_t b0, b1;
        col_range(B, omp_get_thread_num(), omp_get_num_threads(), &b0, &b1);
        for (int64_t t = 0; t < rows; ++t)
            for (int64_t b = b0; b < b1; ++b) dst[t * B + b] = src[t * B + b];
    }
}

void FN(orc_gae_forward)(const real* value, const real* reward, real* adv, int64_t T, int64_t B, double gamma,
                         double lambda) {
    if (T <= 0 || B <= 0) return;
    real* d = (real*)malloc(sizeof(real) * (size_t)T);
    gae_denoms(T, lambda, d);
    const real g_r = (real)gamma, factor = (real)(gamma * lambda);
#pragma omp parallel
    {
        int64_t b0, b1;
        col_range(B, omp_get_thread_num(), omp_get_num_threads(), &b0, &b1);
        const int64_t w = b1 - b0;
        real* g = (real*)calloc((size_t)(w > 0 ? w : 1), sizeof(real));
        for (int64_t t = T - 1; t >= 0 && w > 0; --t) {
            const real* v0 = value + t * B + b0;
            const real* v1 = v0 + B;
            const real* r = reward + t * B + b0;
            real* a = adv + t * B + b0;
            const real dt = d[t];
            for (int64_t b = 0; b < w; ++b) {
                real delta = (r[b] + g_r * v1[b]) - v0[b];
                real gi = dt * delta + factor * g[b];
                g[b] = gi;
                a[b] = gi / dt;
            }
        }
        free(g);
    }
    free(d);
}

/* Adjoint of the above (the reference's GAEFunction.backward returns None,
 * hpc_rll/rl_utils/gae.py:16-18; origin/gae.py is differentiated by autograd).  SURVEY.md A.1:
 *   ghat_t = G_t/d_t + gamma*lambda*ghat_{t-1};  ddelta_t = d_t*ghat_t
 *   grad_reward_t = ddelta_t;  grad_value_t = -ddelta_t[t<T] + gamma*ddelta_{t-1}[t>=1]. */
void FN(orc_gae_backward)(const real* grad_adv, real* grad_value, real* grad_reward, int64_t T, int64_t B,
                          double gamma, double lambda) {
    if (T <= 0 || B <= 0) return;
    real* d = (real*)malloc(sizeof(real) * (size_t)T);
    gae_denoms(T, lambda, d);
    const real g_r = (real)gamma, factor = (real)(gamma * lambda);
#pragma omp parallel
    {
        int64_t b0, b1;
        col_range(B, omp_get_thread_num(), omp_get_num_threads(), &b0, &b1);
        const int64_t w = b1 - b0;
        real* gh = (real*)calloc((size_t)(w > 0 ? w : 1) * 2, sizeof(real));
        real* prev = gh + (w > 0 ? w : 1);
        for (int64_t t = 0; t < T && w > 0; ++t) {
            const real* G = grad_adv + t * B + b0;
            real* gr = grad_reward + t * B + b0;
            real* gv = grad_value + t * B + b0;
            const real dt = d[t];
            for (int64_t b = 0; b < w; ++b) {
                real h = G[b] / dt + factor * gh[b];
                gh[b] = h;
                real dd = dt * h;
                gr[b] = dd;
                gv[b] = g_r * prev[b] - dd;
                prev[b] = dd;
            }
        }
        for (int64_t b = 0; b < w; ++b) grad_value[T * B + b0 + b] = g_r * prev[b];
        free(gh);
    }
    free(d);
}

/* ------------------------------------------------------------------------------------------
 * TD(lambda)  -- origin/td.py:148-176 (td_lambda_error), 179-244 (lambda returns).
 *   ret[T-1] = r[T-1] + gamma*v[T];  ret[t] = r[t] + disc*ret[t+1] + (gamma-disc)*v[t+1]   (td.py:238-243)
 *   disc = gammas*lambda_ in tensor dtype (td.py:239);  loss = 0.5*mean(w*(ret-v[:-1])^2)    (td.py:174-175)
 *   ret is no_grad (td.py:171) => dloss/dv_t = -w_t(ret_t-v_t)/(TB), t<T; row T gets 0.
 * weight may be NULL (ones).  ret_out / grad_value may be NULL.
 * ---------------------------------------------------------------------------------------- */
void FN(orc_td_lambda)(const real* value, const real* reward, const real* weight, int64_t T, int64_t B,
                       double gamma, double lambda, double coef_loss, real* loss_out, real* ret_out,
                       real* grad_value) {
    const real g_r = (real)gamma, l_r = (real)lambda;
    const real disc = g_r * l_r;
    const real gmd = g_r - disc;
    const double inv_n = 1.0 / ((double)T * (double)B);
    double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t b = 0; b < B; ++b) {
        real ret = (real)0;
        double acc = 0.0;
        for (int64_t t = T - 1; t >= 0; --t) {
            real r = reward[t * B + b], v1 = value[(t + 1) * B + b], v0 = value[t * B + b];
            if (t == T - 1)
                ret = r + g_r * v1;
            else
                ret = (r + disc * ret) + gmd * v1;
            real w = weight ? weight[t * B + b] : (real)1;
            real diff = ret - v0;
            acc += (double)((diff * diff) * w);
            if (ret_out) ret_out[t * B + b] = ret;
            if (grad_value) grad_value[t * B + b] = (real)(-coef_loss * (double)w * (double)diff * inv_n);
        }
        if (grad_value) grad_value[T * B + b] = (real)0;
        total += acc;
    }
    *loss_out = (real)(0.5 * total * inv_n);
}

/* ------------------------------------------------------------------------------------------
 * V-trace  -- origin/vtrace.py:63-79 (vtrace_error), 5-13 (nstep return), 16-17 (advantage),
 * 81-111 (importance weights).  target/behaviour (T,B,N), action (T,B) int64, value (T+1,B),
 * reward (T,B), weight (T,B) or NULL.  losses_out[3] = {policy, value, entropy}.
 * grads use upstream coefficients coef[3] for the three scalars.
 * ---------------------------------------------------------------------------------------- */
void FN(orc_vtrace)(const real* target, const real* behaviour, const int64_t* action, const real* value,
                    const real* reward, const real* weight, int64_t T, int64_t B, int64_t N, double gamma,
                    double lambda, double rho_clip, double c_clip, double rho_pg_clip, const double* coef,
                    real* losses_out, real* ret_out, real* adv_out, real* grad_target, real* grad_value) {
    const real g_r = (real)gamma, factor = (real)(gamma * lambda);
    const real rc = (real)rho_clip, cc = (real)c_clip, pc = (real)rho_pg_clip;
    const double inv_n = 1.0 / ((double)T * (double)B);
    double s_pg = 0.0, s_v = 0.0, s_e = 0.0;
#pragma omp parallel reduction(+ : s_pg, s_v, s_e)
    {
        real* lp = (real*)malloc(sizeof(real) * (size_t)N * 3);
        real* pr = lp + N;
        real* lb = pr + N;
        real* is_col = (real*)malloc(sizeof(real) * (size_t)T * 4);
        real* logp_col = is_col + T;
        real* ent_col = logp_col + T;
        real* adv_col = ent_col + T;
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            for (int64_t t = 0; t < T; ++t) {
                int64_t a = action[t * B + b];
                real h = categorical_row(target + (t * B + b) * N, N, lp, pr);
                (void)categorical_row(behaviour + (t * B + b) * N, N, lb, pr);
                logp_col[t] = lp[a];
                ent_col[t] = h;
                is_col[t] = R_EXP(lp[a] - lb[a]); /* vtrace.py:109-110 */
            }
            real item = (real)0, ret_next = value[T * B + b];
            for (int64_t t = T - 1; t >= 0; --t) {
                real is = is_col[t];
                real rho = is < rc ? is : rc, c = is < cc ? is : cc, rpg = is < pc ? is : pc;
                real r = reward[t * B + b], v1 = value[(t + 1) * B + b], v0 = value[t * B + b];
                real delta = rho * ((r + g_r * v1) - v0);      /* vtrace.py:6 */
                item = delta + (factor * c) * item;            /* vtrace.py:11 */
                real ret = v0 + item;                          /* vtrace.py:8,12 */
                real adv = rpg * ((r + g_r * ret_next) - v0);  /* vtrace.py:17,70-71 */
                ret_next = ret;
                real w = weight ? weight[t * B + b] : (real)1;
                s_pg += (double)((logp_col[t] * adv) * w);
                real dv = v0 - ret;
                s_v += (double)((dv * dv) * w);
                s_e += (double)(ent_col[t] * w);
                adv_col[t] = adv;
                if (ret_out) ret_out[t * B + b] = ret;
                if (adv_out) adv_out[t * B + b] = adv;
                if (grad_value) grad_value[t * B + b] = (real)(coef[1] * 2.0 * (double)dv * (double)w * inv_n);
            }
            if (grad_value) grad_value[T * B + b] = (real)0;
            if (grad_target) {
                for (int64_t t = 0; t < T; ++t) {
                    int64_t a = action[t * B + b];
                    real h = categorical_row(target + (t * B + b) * N, N, lp, pr);
                    real w = weight ? weight[t * B + b] : (real)1;
                    double c1 = coef[0] * (-(double)adv_col[t] * (double)w * inv_n);
                    double c2 = coef[2] * ((double)w * inv_n);
                    for (int64_t k = 0; k < N; ++k) {
                        double oh = k == a ? 1.0 : 0.0;
                        double g = c1 * (oh - (double)pr[k]) + c2 * (-(double)pr[k] * ((double)lp[k] + (double)h));
                        grad_target[(t * B + b) * N + k] = (real)g;
                    }
                }
            }
        }
        free(lp);
        free(is_col);
    }
    losses_out[0] = (real)(-s_pg * inv_n);
    losses_out[1] = (real)(s_v * inv_n);
    losses_out[2] = (real)(s_e * inv_n);
}

/* ------------------------------------------------------------------------------------------
 * UPGO  -- origin/upgo.py:21-38 (returns), 40-70 (loss), 7-18 (tb_cross_entropy).
 *   lambda'_t = [r_{t+1}+v_{t+2} >= v_{t+1}] (t<T-1)   (upgo.py:36-37)
 *   ret_{T-1} = r + v_T; ret_t = r_t + l*ret_{t+1} + (1-l)*v_{t+1}   (td.py:238-243 with gamma=1)
 *   adv = rho*(ret - v[:-1]) (no_grad, upgo.py:64-65); loss = -mean(adv*logp[a]) (upgo.py:66-70)
 * ---------------------------------------------------------------------------------------- */
void FN(orc_upgo)(const real* target, const real* rhos, const int64_t* action, const real* rewards,
                  const real* values, int64_t T, int64_t B, int64_t N, double coef_loss, real* loss_out,
                  real* ret_out, real* grad_target) {
    const double inv_n = 1.0 / ((double)T * (double)B);
    double total = 0.0;
#pragma omp parallel reduction(+ : total)
    {
        real* lp = (real*)malloc(sizeof(real) * (size_t)N);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            real ret = (real)0;
            for (int64_t t = T - 1; t >= 0; --t) {
                real r = rewards[t * B + b], v1 = values[(t + 1) * B + b], v0 = values[t * B + b];
                if (t == T - 1) {
                    ret = r + v1;
                } else {
                    real r1 = rewards[(t + 1) * B + b], v2 = values[(t + 2) * B + b];
                    real l = (r1 + v2) >= v1 ? (real)1 : (real)0;
                    ret = (r + l * ret) + ((real)1 - l) * v1;
                }
                real adv = rhos[t * B + b] * (ret - v0);
                log_softmax_row(target + (t * B + b) * N, N, lp);
                int64_t a = action[t * B + b];
                total += (double)(adv * lp[a]);
                if (ret_out) ret_out[t * B + b] = ret;
                if (grad_target) {
                    double c = coef_loss * (-(double)adv * inv_n);
                    for (int64_t k = 0; k < N; ++k) {
                        double oh = k == a ? 1.0 : 0.0;
                        grad_target[(t * B + b) * N + k] = (real)(c * (oh - (double)R_EXP(lp[k])));
                    }
                }
            }
        }
        free(lp);
    }
    *loss_out = (real)(-total * inv_n);
}

/* ------------------------------------------------------------------------------------------
 * PPO  -- origin/ppo.py:51-80.  logits (B,N), action (B) int64, the rest (B).  weight NULL = ones.
 * dual_clip <= 0 means None.  out[5] = {policy, value, entropy, approx_kl, clipfrac}.
 * Tie rules follow autograd: torch.min/max split the gradient evenly on exact ties and clamp
 * passes it on the closed interval.
 * ---------------------------------------------------------------------------------------- */
void FN(orc_ppo)(const real* logits_new, const real* logits_old, const int64_t* action, const real* value_new,
                 const real* value_old, const real* adv, const real* return_, const real* weight, int64_t B,
                 int64_t N, double clip_ratio, int use_value_clip, double dual_clip, const double* coef, real* out,
                 real* grad_logits, real* grad_value) {
    const real lo = (real)(1.0 - clip_ratio), hi = (real)(1.0 + clip_ratio), eps = (real)clip_ratio;
    const real dc = (real)dual_clip;
    const double inv_n = 1.0 / (double)B;
    double s_p = 0, s_v = 0, s_e = 0, s_kl = 0, s_cf = 0;
#pragma omp parallel reduction(+ : s_p, s_v, s_e, s_kl, s_cf)
    {
        real* lp = (real*)malloc(sizeof(real) * (size_t)N * 4);
        real* pr = lp + N;
        real* lo_ = pr + N;
        real* prtmp = lo_ + N;
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            int64_t a = action[b];
            real h = categorical_row(logits_new + b * N, N, lp, pr);
            real logp_new = lp[a];
            (void)categorical_row(logits_old + b * N, N, lo_, prtmp);
            real logp_old = lo_[a];
            real w = weight ? weight[b] : (real)1;
            real ratio = R_EXP(logp_new - logp_old);
            real ad = adv[b];
            real s1 = ratio * ad;
            real rcl = ratio < lo ? lo : (ratio > hi ? hi : ratio);
            real s2 = rcl * ad;
            real m = s1 < s2 ? s1 : s2;
            /* d m / d ratio */
            double in_range = (ratio >= lo && ratio <= hi) ? 1.0 : 0.0;
            double g1 = s1 < s2 ? 1.0 : (s1 == s2 ? 0.5 : 0.0);
            double dm = g1 * (double)ad + (1.0 - g1) * (double)ad * in_range;
            real pol = m;
            if (dual_clip > 0.0) {
                real d = dc * ad;
                double gm = m > d ? 1.0 : (m == d ? 0.5 : 0.0);
                pol = m > d ? m : d;
                dm *= gm;
            }
            s_p += (double)(-pol * w);
            s_kl += (double)(logp_old - logp_new);
            s_cf += (ratio > hi || ratio < lo) ? 1.0 : 0.0;
            real vn = value_new[b], vo = value_old[b], rt = return_[b];
            real e1 = rt - vn;
            real v1 = e1 * e1;
            double dval;
            real vl;
            if (use_value_clip) {
                real dvv = vn - vo;
                real cl = dvv < -eps ? -eps : (dvv > eps ? eps : dvv);
                real vclip = vo + cl;
                real e2 = rt - vclip;
                real v2 = e2 * e2;
                double inr = (dvv >= -eps && dvv <= eps) ? 1.0 : 0.0;
                double k1 = v1 > v2 ? 1.0 : (v1 == v2 ? 0.5 : 0.0);
                vl = v1 > v2 ? v1 : v2;
                dval = k1 * (-2.0 * (double)e1) + (1.0 - k1) * (-2.0 * (double)e2) * inr;
            } else {
                vl = v1;
                dval = -2.0 * (double)e1;
            }
            s_v += (double)(vl * w);
            s_e += (double)(h * w);
            if (grad_value) grad_value[b] = (real)(coef[1] * 0.5 * dval * (double)w * inv_n);
            if (grad_logits) {
                /* d policy_loss / d logp_new[a] = -dm*ratio*w/B */
                double c1 = coef[0] * (-dm * (double)ratio * (double)w * inv_n);
                double c2 = coef[2] * ((double)w * inv_n);
                for (int64_t k = 0; k < N; ++k) {
                    double oh = k == a ? 1.0 : 0.0;
                    grad_logits[b * N + k] =
                        (real)(c1 * (oh - (double)pr[k]) + c2 * (-(double)pr[k] * ((double)lp[k] + (double)h)));
                }
            }
        }
        free(lp);
    }
    out[0] = (real)(s_p * inv_n);
    out[1] = (real)(0.5 * s_v * inv_n);
    out[2] = (real)(s_e * inv_n);
    out[3] = (real)(s_kl * inv_n);
    out[4] = (real)(s_cf * inv_n);
}

/* ------------------------------------------------------------------------------------------
 * q n-step TD (+ value rescale)  -- origin/td.py:252-291, 294-340, 345-354.
 *   target = R + gamma^n * next_q[a'] * (1-done);  td = (q[a]-target)^2;  loss = mean(td*w)
 *   rescale: target = h(R + gamma^n * h^-1(next_q[a']) * (1-done)), eps = 1e-2.
 * q,next_n_q (B,N); reward (T,B); done (B) float; weight NULL = ones.
 * ---------------------------------------------------------------------------------------- */
void FN(orc_q_nstep_td)(const real* q, const real* next_n_q, const int64_t* action, const int64_t* next_action,
                        const real* reward, const real* done, const real* weight, int64_t T, int64_t B,
                        int64_t N, double gamma, int rescale, double coef_loss, real* loss_out, real* td_err,
                        real* grad_q) {
    const real g_r = (real)gamma, gn = (real)pow(gamma, (double)T), eps = (real)1e-2;
    const double inv_n = 1.0 / (double)B;
    double total = 0.0;
    if (grad_q) memset(grad_q, 0, sizeof(real) * (size_t)B * (size_t)N);
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t b = 0; b < B; ++b) {
        real qa = q[b * N + action[b]];
        real tq = next_n_q[b * N + next_action[b]];
        if (rescale) tq = value_inv_transform(tq, eps);
        real R = nstep_reward(reward, T, B, b, g_r);
        real target = R + (gn * tq) * ((real)1 - done[b]);
        if (rescale) target = value_transform(target, eps);
        real diff = qa - target;
        real td = diff * diff;
        real w = weight ? weight[b] : (real)1;
        td_err[b] = td;
        total += (double)(td * w);
        if (grad_q) grad_q[b * N + action[b]] = (real)(coef_loss * 2.0 * (double)diff * (double)w * inv_n);
    }
    *loss_out = (real)(total * inv_n);
}

/* ------------------------------------------------------------------------------------------
 * C51 distributional n-step TD  -- origin/td.py:56-143.
 *   support = linspace(v_min,v_max,n_atom); Tz = clamp(R + (1-done)*gamma^n*support)
 *   b = (Tz-v_min)/dz; l=floor(b), u=ceil(b); proj[l]+=p'(u-b); proj[u]+=p'(b-l)   (td.py:96-117)
 *   (when l==u both weights are 0 and the mass is dropped -- origin does the same)
 *   td_err = -sum log p[a]*proj;  loss = -mean(sum log p*proj*w)                     (td.py:129-137)
 * dist,next_n_dist (B,N,n_atom).
 * ---------------------------------------------------------------------------------------- */
void FN(orc_dist_nstep_td)(const real* dist, const real* next_n_dist, const int64_t* action,
                           const int64_t* next_action, const real* reward, const real* done, const real* weight,
                           int64_t T, int64_t B, int64_t N, int64_t n_atom, double gamma, double v_min,
                           double v_max, double coef_loss, real* loss_out, real* td_err, real* grad_dist) {
    const real g_r = (real)gamma, gn = (real)pow(gamma, (double)T);
    const real vmin = (real)v_min, vmax = (real)v_max;
    const real dz = (real)((v_max - v_min) / (double)(n_atom - 1));
    const double inv_n = 1.0 / (double)B;
    /* torch.linspace (CPU): step=(end-start)/(steps-1); i<steps/2 ? start+step*i : end-step*(steps-1-i) */
    real* support = (real*)malloc(sizeof(real) * (size_t)n_atom);
    {
        real step = (vmax - vmin) / (real)(n_atom - 1);
        int64_t half = n_atom / 2;
        for (int64_t i = 0; i < n_atom; ++i)
            support[i] = i < half ? vmin + step * (real)i : vmax - step * (real)(n_atom - 1 - i);
    }
    double total = 0.0;
    if (grad_dist) memset(grad_dist, 0, sizeof(real) * (size_t)B * (size_t)N * (size_t)n_atom);
#pragma omp parallel reduction(+ : total)
    {
        real* proj = (real*)malloc(sizeof(real) * (size_t)n_atom);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            const real* pn = next_n_dist + (b * N + next_action[b]) * n_atom;
            const real* pd = dist + (b * N + action[b]) * n_atom;
            real R = nstep_reward(reward, T, B, b, g_r);
            real sc = ((real)1 - done[b]) * gn;
            for (int64_t k = 0; k < n_atom; ++k) proj[k] = (real)0;
            for (int pass = 0; pass < 2; ++pass) { /* two index_add_ calls, td.py:116-117 */
                for (int64_t j = 0; j < n_atom; ++j) {
                    real tz = R + sc * support[j];
                    tz = tz < vmin ? vmin : (tz > vmax ? vmax : tz);
                    real bb = (tz - vmin) / dz;
                    real l = R_FLOOR(bb), u = R_CEIL(bb);
                    if (pass == 0)
                        proj[(int64_t)l] += pn[j] * (u - bb);
                    else
                        proj[(int64_t)u] += pn[j] * (bb - l);
                }
            }
            real w = weight ? weight[b] : (real)1;
            double acc = 0.0, accw = 0.0;
            for (int64_t k = 0; k < n_atom; ++k) {
                real lp = R_LOG(pd[k]);
                acc += (double)(lp * proj[k]);
                accw += (double)((lp * proj[k]) * w);
                if (grad_dist)
                    grad_dist[(b * N + action[b]) * n_atom + k] =
                        (real)(-coef_loss * (double)w * (double)proj[k] / (double)pd[k] * inv_n);
            }
            td_err[b] = (real)(-acc);
            total += accw;
        }
        free(proj);
    }
    free(support);
    *loss_out = (real)(-total * inv_n);
}

/* ------------------------------------------------------------------------------------------
 * QR-DQN n-step TD  -- origin/td.py:480-517.  q,next_n_q (B,N,tau).
 *   target_j = R + vg*next_q[a',j]*(1-done); e_ij = target_j - q[a,i]
 *   u = smooth_l1(e) (beta=1, quadratic for |e|<1);  weight = |tau - 1{e<=0}|, `tau` is the
 *   INTEGER quantile count exactly as the reference wrapper/test pass it
 *   (hpc_rll/rl_utils/td.py:538, tests/test_qrdqn_nstep_td_error.py:57)
 *   td_b = mean_i sum_j u*weight; loss = mean(td*w).  value_gamma NULL => gamma^n.
 * ---------------------------------------------------------------------------------------- */
void FN(orc_qrdqn_nstep_td)(const real* q, const real* next_n_q, const int64_t* action, const int64_t* next_action,
                            const real* reward, const real* done, const real* weight, const real* value_gamma,
                            int64_t tau, int64_t T, int64_t B, int64_t N, double gamma, double coef_loss,
                            real* loss_out, real* td_err, real* grad_q) {
    const real g_r = (real)gamma, gn = (real)pow(gamma, (double)T), tauf = (real)tau;
    const double inv_n = 1.0 / (double)B;
    double total = 0.0;
    if (grad_q) memset(grad_q, 0, sizeof(real) * (size_t)B * (size_t)N * (size_t)tau);
#pragma omp parallel reduction(+ : total)
    {
        real* tg = (real*)malloc(sizeof(real) * (size_t)tau);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            const real* qa = q + (b * N + action[b]) * tau;
            const real* nq = next_n_q + (b * N + next_action[b]) * tau;
            real R = nstep_reward(reward, T, B, b, g_r);
            real vg = value_gamma ? value_gamma[b] : gn;
            real nd = (real)1 - done[b];
            for (int64_t j = 0; j < tau; ++j) tg[j] = R + (vg * nq[j]) * nd;
            real w = weight ? weight[b] : (real)1;
            double acc = 0.0;
            for (int64_t i = 0; i < tau; ++i) {
                double row = 0.0, grow = 0.0;
                for (int64_t j = 0; j < tau; ++j) {
                    real e = tg[j] - qa[i];
                    real ae = R_FABS(e);
                    real u = ae < (real)1 ? ((real)0.5 * ae) * ae : ae - (real)0.5;
                    real du = ae < (real)1 ? e : sgn(e);
                    real wt = R_FABS(tauf - (e <= (real)0 ? (real)1 : (real)0));
                    row += (double)(u * wt);
                    grow += (double)(du * wt);
                }
                acc += row;
                if (grad_q)
                    grad_q[(b * N + action[b]) * tau + i] =
                        (real)(-coef_loss * (double)w * inv_n * grow / (double)tau);
            }
            real td = (real)(acc / (double)tau);
            td_err[b] = td;
            total += (double)(td * w);
        }
        free(tg);
    }
    *loss_out = (real)(total * inv_n);
}

/* ------------------------------------------------------------------------------------------
 * IQN n-step TD  -- origin/td.py:391-448.  q (tau,B,N), next_n_q (tau',B,N), replay_quantiles (tau,B).
 *   e_ji = target_j - q_i;  huber_k (quadratic for |e|<=kappa);  weight |rq[i,b] - 1{e<0}| / kappa
 *   td_b = mean_j sum_i;  loss = mean(td*w).
 * ---------------------------------------------------------------------------------------- */
void FN(orc_iqn_nstep_td)(const real* q, const real* next_n_q, const int64_t* action, const int64_t* next_action,
                          const real* reward, const real* done, const real* replay_quantiles, const real* weight,
                          const real* value_gamma, int64_t tau, int64_t tau_p, int64_t T, int64_t B, int64_t N,
                          double gamma, double kappa, double coef_loss, real* loss_out, real* td_err,
                          real* grad_q) {
    const real g_r = (real)gamma, gn = (real)pow(gamma, (double)T), kp = (real)kappa;
    const double inv_n = 1.0 / (double)B;
    double total = 0.0;
    if (grad_q) memset(grad_q, 0, sizeof(real) * (size_t)tau * (size_t)B * (size_t)N);
#pragma omp parallel reduction(+ : total)
    {
        real* tg = (real*)malloc(sizeof(real) * (size_t)tau_p);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            real R = nstep_reward(reward, T, B, b, g_r);
            real vg = value_gamma ? value_gamma[b] : gn;
            real nd = (real)1 - done[b];
            for (int64_t j = 0; j < tau_p; ++j) tg[j] = R + (vg * next_n_q[(j * B + b) * N + next_action[b]]) * nd;
            real w = weight ? weight[b] : (real)1;
            double acc = 0.0;
            for (int64_t i = 0; i < tau; ++i) {
                real qi = q[(i * B + b) * N + action[b]];
                real rq = replay_quantiles[i * B + b];
                double col = 0.0, gcol = 0.0;
                for (int64_t j = 0; j < tau_p; ++j) {
                    real e = tg[j] - qi;
                    real ae = R_FABS(e);
                    real hub = ae <= kp ? ((real)0.5 * e) * e : kp * (ae - (real)0.5 * kp);
                    real dh = ae <= kp ? e : kp * sgn(e);
                    real wt = R_FABS(rq - (e < (real)0 ? (real)1 : (real)0));
                    col += (double)((wt * hub) / kp);
                    gcol += (double)((wt * dh) / kp);
                }
                acc += col;
                if (grad_q)
                    grad_q[(i * B + b) * N + action[b]] =
                        (real)(-coef_loss * (double)w * inv_n * gcol / (double)tau_p);
            }
            real td = (real)(acc / (double)tau_p);
            td_err[b] = td;
            total += (double)(td * w);
        }
        free(tg);
    }
    *loss_out = (real)(total * inv_n);
}
