/* 80-bit long-double evaluation of the EKLD hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Purpose: an arbiter that is ~1000x more accurate than any FP64 route, used by tests/ to decide conditioning
 * disputes between the GPy-route oracle (oracle/ekld_oracle.py, explicit inverse via dpotri) and the CUDA path
 * (triangular factors).  It evaluates the same mathematics as reference bode/_sampler.py:391-410, 464-481 and
 * bode/_core.py:86-101 with direct-difference ARD-RBF distances, a long-double Cholesky and triangular solves.
 * The EKLD is evaluated in its analytically cancelled form  -1/2 log(1 - v^2 / (sigma_1 k_xx))  which equals the
 * reference expression (_sampler.py:480) exactly in real arithmetic (sigma_2/(2 sigma_1) + v^2/(2 sigma_1 k_xx)
 * - 1/2 == 0 because sigma_2 = sigma_1 - v^2/k_xx).
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC).  x86-64 long double = x87 extended, eps 1.08e-19.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long double ld;

static ld xik_ld(const double *x, const ld *ell, ld ss, int d) {
    /* _core.py:86-92 */
    const ld s2 = sqrtl(2.0L), c = sqrtl(acosl(-1.0L)) / s2;
    ld p = 1.0L;
    for (int k = 0; k < d; k++)
        p *= c * ell[k] * (erfl((1.0L - (ld)x[k]) / (s2 * ell[k])) - erfl((0.0L - (ld)x[k]) / (s2 * ell[k])));
    return ss * p;
}

static ld bk_ld(const ld *ell, ld ss, int d) {
    /* _core.py:95-101 */
    const ld sp = sqrtl(acosl(-1.0L));
    ld p = 1.0L;
    for (int k = 0; k < d; k++) {
        ld z = 1.0L / (sqrtl(2.0L) * ell[k]);
        p *= 2.0L * sp * ell[k] * ell[k] * (z * erfl(z) + expl(-z * z) / sp - 1.0L / sp);
    }
    return ss * p;
}

static ld rbf_ld(const double *a, const double *b, const ld *ell, ld ss, int d) {
    ld r2 = 0.0L;
    for (int k = 0; k < d; k++) {
        ld t = ((ld)a[k] - (ld)b[k]) / ell[k];
        r2 += t * t;
    }
    return ss * expl(-0.5L * r2);
}

/* Returns 0, or s+1 if the Cholesky of sample s met a non-positive pivot.
 * hyp: (S, ndim) rows [variance, ell_1..ell_d (, noise_std)]; noise variance = hyp[-1]^2 if ndim==d+2 else nugget_var.
 * Outputs (double, any may be NULL): kld (S*Nd), sigx (S*Nd, latent predictive variance), v (S*Nd),
 * mu1 (S), sig1 (S), loglik (S). */
/* (weighted) mean over the quadrature points of k(x, x_q): quadrature stand-in for xik (north-star quad mode) */
static ld quad_mean_ld(const double *x, const double *Xq, const double *wq, long Nq, const ld *ell, ld ss, int d) {
    ld acc = 0.0L, wsum = 0.0L;
    for (long q = 0; q < Nq; q++) {
        ld w = wq ? (ld)wq[q] : 1.0L;
        acc += w * rbf_ld(x, Xq + (size_t)q * d, ell, ss, d);
        wsum += w;
    }
    return acc / wsum;
}

int ekld_truth(const double *X, const double *Y, int n, int d, const double *hyp, int S, int ndim,
               double nugget_var, double jitter, const double *Xd, long Nd,
               double *kld, double *sigx, double *v_out, double *mu1, double *sig1, double *loglik,
               const double *Xq, const double *wq, long Nq) {
    ld *A = (ld *)malloc(sizeof(ld) * (size_t)n * n);
    ld *e = (ld *)malloc(sizeof(ld) * n), *a = (ld *)malloc(sizeof(ld) * n), *w = (ld *)malloc(sizeof(ld) * n);
    ld *al = (ld *)malloc(sizeof(ld) * n), *k = (ld *)malloc(sizeof(ld) * n), *ell = (ld *)malloc(sizeof(ld) * d);
    int rc = 0;
    for (int s = 0; s < S && rc == 0; s++) {
        const double *h = hyp + (size_t)s * ndim;
        ld ss = h[0];
        for (int q = 0; q < d; q++) ell[q] = h[1 + q];
        ld nv = (ndim == d + 2) ? (ld)h[d + 1] * (ld)h[d + 1] : (ld)nugget_var;
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) {
                ld kij = (i == j) ? ss : rbf_ld(X + (size_t)i * d, X + (size_t)j * d, ell, ss, d);
                if (i == j) kij += nv + (ld)jitter;
                A[(size_t)i * n + j] = kij;
            }
        /* Cholesky, lower, in place */
        for (int j = 0; j < n && rc == 0; j++) {
            ld dj = A[(size_t)j * n + j];
            for (int q = 0; q < j; q++) dj -= A[(size_t)j * n + q] * A[(size_t)j * n + q];
            if (!(dj > 0.0L)) { rc = s + 1; break; }
            dj = sqrtl(dj);
            A[(size_t)j * n + j] = dj;
            for (int i = j + 1; i < n; i++) {
                ld t = A[(size_t)i * n + j];
                for (int q = 0; q < j; q++) t -= A[(size_t)i * n + q] * A[(size_t)j * n + q];
                A[(size_t)i * n + j] = t / dj;
            }
        }
        if (rc) break;
        for (int i = 0; i < n; i++)
            e[i] = Xq ? quad_mean_ld(X + (size_t)i * d, Xq, wq, Nq, ell, ss, d) : xik_ld(X + (size_t)i * d, ell, ss, d);
        /* a = L^-1 e ; w = L^-T a ; al = A^-1 Y */
        for (int i = 0; i < n; i++) {
            ld t = e[i];
            for (int q = 0; q < i; q++) t -= A[(size_t)i * n + q] * a[q];
            a[i] = t / A[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            ld t = a[i];
            for (int q = i + 1; q < n; q++) t -= A[(size_t)q * n + i] * w[q];
            w[i] = t / A[(size_t)i * n + i];
        }
        for (int i = 0; i < n; i++) {
            ld t = Y[i];
            for (int q = 0; q < i; q++) t -= A[(size_t)i * n + q] * al[q];
            al[i] = t / A[(size_t)i * n + i];
        }
        ld quad = 0.0L, logdet = 0.0L;
        for (int i = 0; i < n; i++) { quad += al[i] * al[i]; logdet += 2.0L * logl(A[(size_t)i * n + i]); }
        for (int i = n - 1; i >= 0; i--) {
            ld t = al[i];
            for (int q = i + 1; q < n; q++) t -= A[(size_t)q * n + i] * al[q];
            al[i] = t / A[(size_t)i * n + i];
        }
        ld s1, m1 = 0.0L;
        if (Xq) {
            ld acc = 0.0L, wsum = 0.0L;
            for (long q = 0; q < Nq; q++) {
                ld w = wq ? (ld)wq[q] : 1.0L;
                acc += w * quad_mean_ld(Xq + (size_t)q * d, Xq, wq, Nq, ell, ss, d);
                wsum += w;
            }
            s1 = acc / wsum;
        } else {
            s1 = bk_ld(ell, ss, d);
        }
        for (int i = 0; i < n; i++) { s1 -= a[i] * a[i]; m1 += al[i] * e[i]; }
        if (mu1) mu1[s] = (double)m1;
        if (sig1) sig1[s] = (double)s1;
        if (loglik) loglik[s] = (double)(0.5L * (-(ld)n * logl(2.0L * acosl(-1.0L)) - logdet - quad));
        for (long c = 0; c < Nd; c++) {
            const double *x = Xd + (size_t)c * d;
            ld vv = Xq ? quad_mean_ld(x, Xq, wq, Nq, ell, ss, d) : xik_ld(x, ell, ss, d), tt = 0.0L;
            for (int i = 0; i < n; i++) { k[i] = rbf_ld(X + (size_t)i * d, x, ell, ss, d); vv -= w[i] * k[i]; }
            for (int i = 0; i < n; i++) {
                ld t = k[i];
                for (int q = 0; q < i; q++) t -= A[(size_t)i * n + q] * k[q];
                t /= A[(size_t)i * n + i];
                k[i] = t;
                tt += t * t;
            }
            ld sx = ss - tt, kxx = sx + nv;
            ld rho = vv * vv / (s1 * kxx);
            if (kld) kld[(size_t)s * Nd + c] = (double)(-0.5L * log1pl(-rho));
            if (sigx) sigx[(size_t)s * Nd + c] = (double)sx;
            if (v_out) v_out[(size_t)s * Nd + c] = (double)vv;
        }
    }
    free(A); free(e); free(a); free(w); free(al); free(k); free(ell);
    return rc;
}
