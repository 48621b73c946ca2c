/* ransac_common.h -- helpers shared by the CPU oracles (ransac_f.c, ransac_h.c, pose.c).
 *
 * TEST INFRASTRUCTURE ONLY (see ransac_f.c).  Operation for operation the same as
 * sfm-project_b200/csrc/ransac_common.cuh; built with -ffp-contract=off.
 */
#ifndef SFM_ORACLE_RANSAC_COMMON_H
#define SFM_ORACLE_RANSAC_COMMON_H
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sfm_b200.h"

/* hypotheses evaluated before the next termination check: 32, 32, 64, then 128 at a time */
static __attribute__((unused)) int ransac_batch(int done) { return done < 64 ? 32 : (done < 128 ? 64 : 128); }
#define SFM_RANSAC_LANES 256      /* virtual reduction lanes of the LO refit    */
#define SFM_LO_ROUNDS 2

typedef struct { double s, cx, cy; } norm2d;

/* ---------------------------------------------------------------- sampling */
static __attribute__((unused)) uint32_t rng_u32(uint64_t seed, uint32_t pair, uint32_t hyp, uint32_t ctr)
{
    uint64_t x = seed + 0x9E3779B97F4A7C15ULL * ((((uint64_t)pair) << 32) | (uint64_t)hyp);
    x ^= 0xD1B54A32D192ED03ULL * (uint64_t)(ctr + 1u);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return (uint32_t)(x >> 32);
}

static __attribute__((unused)) void draw_sample(uint64_t seed, uint32_t pair, uint32_t hyp, int m, int M, int* idx)
{
    for (int k = 0; k < m; ++k) {
        int v = 0;
        for (int attempt = 0; attempt < 16; ++attempt) {
            uint32_t r = rng_u32(seed, pair, hyp, (uint32_t)(k * 16 + attempt));
            v = (int)(((uint64_t)r * (uint64_t)(uint32_t)M) >> 32);
            int dup = 0;
            for (int j = 0; j < k; ++j) dup |= (idx[j] == v);
            if (!dup) break;
        }
        idx[k] = v;
    }
}

/* cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 9).
 * A is destroyed (diagonal = eigenvalues), V gets eigenvectors in columns. */
static __attribute__((unused)) void jacobi_eig(double* A, double* V, int n, int sweeps)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int s = 0; s < sweeps; ++s) {
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double app = A[p * n + p], aqq = A[q * n + q];
                double theta = (aqq - app) / (2.0 * apq);
                double at = fabs(theta);
                double t = 1.0 / (at + sqrt(theta * theta + 1.0));
                if (theta < 0.0) t = -t;
                double c = 1.0 / sqrt(t * t + 1.0);
                double sn = t * c;
                for (int k = 0; k < n; ++k) {        /* columns p,q */
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - sn * akq;
                    A[k * n + q] = sn * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {        /* rows p,q */
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - sn * aqk;
                    A[q * n + k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - sn * vkq;
                    V[k * n + q] = sn * vkp + c * vkq;
                }
            }
        }
    }
}

/* stop when (1 - w^m)^done <= 1 - confidence, IEEE basic ops only */
static __attribute__((unused)) int should_stop(int best, int M, int m, int done, double confidence)
{
    if (confidence >= 1.0 || best <= 0) return 0;
    double w = (double)best / (double)M;
    double wm = 1.0;
    for (int k = 0; k < m; ++k) wm *= w;
    double q = 1.0 - wm;
    if (!(q > 0.0)) return 1;
    double res = 1.0, base = q;
    int e = done;
    while (e) { if (e & 1) res *= base; base *= base; e >>= 1; }
    return res <= (1.0 - confidence);
}

static __attribute__((unused)) double lane_tree(double* v /* [SFM_RANSAC_LANES] partials */)
{
    /* per 32-lane warp: shfl_down tree; then warps summed in order */
    double total = 0.0;
    for (int w = 0; w < SFM_RANSAC_LANES / 32; ++w) {
        double* l = v + 32 * w;
        for (int off = 16; off >= 1; off >>= 1)
            for (int i = 0; i < off; ++i) l[i] += l[i + off];
        total = (w == 0) ? l[0] : total + l[0];
    }
    return total;
}


/* Eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 9 x 9 matrix: Cholesky factor of
 * A + eps*I (eps = 1e-13 * trace), then inverse iterations from a fixed start vector until the normalised iterate stops
 * moving (at most 16).  Operation for operation
 * smallest_eigvec9 of csrc/ransac_common.cuh. */
static __attribute__((unused)) int smallest_eigvec9(const double* A, double* xout)
{
    double L[45], invd[9], x[9], y[9], z[9];
    double tr = 0.0;
    for (int i = 0; i < 9; ++i) tr += A[i * 9 + i];
    if (!(tr > 0.0) || !(tr < 1e300)) return 0;
    const double eps = tr * 1e-13;
    for (int j = 0; j < 9; ++j) {
        double sj = A[j * 9 + j] + eps;
        for (int k = 0; k < j; ++k) sj -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
        if (!(sj > 0.0)) return 0;
        const double d = sqrt(sj);
        L[j * (j + 1) / 2 + j] = d;
        invd[j] = 1.0 / d;
        for (int i = j + 1; i < 9; ++i) {
            double v = A[i * 9 + j];
            for (int k = 0; k < j; ++k) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            L[i * (i + 1) / 2 + j] = v * invd[j];
        }
    }
    for (int i = 0; i < 9; ++i) x[i] = 1.0 + 0.125 * (double)i;
    for (int it = 0; it < 16; ++it) {
        for (int i = 0; i < 9; ++i) {
            double v = x[i];
            for (int k = 0; k < i; ++k) v -= L[i * (i + 1) / 2 + k] * y[k];
            y[i] = v * invd[i];
        }
        for (int i = 8; i >= 0; --i) {
            double v = y[i];
            for (int k = i + 1; k < 9; ++k) v -= L[k * (k + 1) / 2 + i] * z[k];
            z[i] = v * invd[i];
        }
        double ss = 0.0;
        for (int i = 0; i < 9; ++i) ss += z[i] * z[i];
        if (!(ss > 0.0) || !(ss < 1e300)) return 0;
        const double inv = 1.0 / sqrt(ss);
        double diff = 0.0;
        for (int i = 0; i < 9; ++i) {
            const double xi = z[i] * inv;
            diff = fmax(diff, fabs(xi - x[i]));
            x[i] = xi;
        }
        if (it > 0 && diff < 1e-13) break;
    }
    for (int i = 0; i < 9; ++i) xout[i] = x[i];
    return 1;
}
#endif
