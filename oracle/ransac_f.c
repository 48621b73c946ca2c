/* ransac_f.c -- CPU oracle for batched RANSAC fundamental-matrix verification.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sfm-project_b200/ links, imports or
 * executes this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may.
 *
 * What it restates.  The reference's code/geometric_verification.py is a 0-byte
 * file (placeholder comment at code/pipeline.py:60), so there is no reference
 * algorithm to follow line by line.  The conventions (F with x2^T F x1 = 0,
 * F[8] normalised to 1, uint8 mask, inlier iff max(d1^2,d2^2) <= thr^2) are those
 * of cv2.findFundamentalMat(FM_RANSAC) -- OpenCV 4.13.0 (third party, not under
 * /root/reference, version unpinned by the reference) -- as characterised in
 * SURVEY.md Appendix A.4 and pinned by tests/test_oracle_pinned.py against cv2
 * run in this image.  Sampling and termination are this build's own: a
 * counter-based RNG, hypotheses evaluated in batches of 32, 32, 64, then 128 (ransac_batch), and a
 * stop rule that uses only IEEE + - * / sqrt so that the CUDA kernel and this
 * file produce bit-identical masks and counts.  PARITY UNPINNED by the reference
 * (it has no tests); pinned against cv2 statistically (tests/).
 *
 * Build:  gcc -O2 -ffp-contract=off -mfma -fPIC -shared  (oracle/Makefile)
 * Every floating-point operation is written in the order the CUDA kernel
 * (sfm-project_b200/csrc/ransac_f.cu) performs it; no contraction is allowed
 * on either side except the explicit fma()/fmaf() calls.
 */
#include "ransac_common.h"

/* ------------------------------------------------------------ small algebra */
static double det3(const double* r0, const double* r1, const double* r2)
{
    double a = r1[1] * r2[2] - r1[2] * r2[1];
    double b = r1[0] * r2[2] - r1[2] * r2[0];
    double c = r1[0] * r2[1] - r1[1] * r2[0];
    return r0[0] * a - r0[1] * b + r0[2] * c;
}

/* F <- closest rank-2 matrix: remove the component along the right singular
 * vector of the smallest singular value. */
static void enforce_rank2(double* F)
{
    double G[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            G[i * 3 + j] = F[0 + i] * F[0 + j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
    jacobi_eig(G, V, 3, 6);
    int k = 0;
    if (G[4] < G[k * 4]) k = 1;
    if (G[8] < G[k * 4]) k = 2;
    double v0 = V[0 + k], v1 = V[3 + k], v2 = V[6 + k];
    for (int r = 0; r < 3; ++r) {
        double w = F[r * 3 + 0] * v0 + F[r * 3 + 1] * v1 + F[r * 3 + 2] * v2;
        F[r * 3 + 0] -= w * v0;
        F[r * 3 + 1] -= w * v1;
        F[r * 3 + 2] -= w * v2;
    }
}


/* F = T2^T Fh T1, then scale to unit Frobenius norm.  Returns 0 if unusable. */
static int denormalise(const double* Fh, norm2d n1, norm2d n2, double* F)
{
    double G[9];
    for (int r = 0; r < 3; ++r) {
        double f0 = Fh[r * 3 + 0], f1 = Fh[r * 3 + 1], f2 = Fh[r * 3 + 2];
        G[r * 3 + 0] = n1.s * f0;
        G[r * 3 + 1] = n1.s * f1;
        G[r * 3 + 2] = f2 - n1.s * (n1.cx * f0 + n1.cy * f1);
    }
    for (int c = 0; c < 3; ++c) {
        double g0 = G[0 + c], g1 = G[3 + c], g2 = G[6 + c];
        F[0 + c] = n2.s * g0;
        F[3 + c] = n2.s * g1;
        F[6 + c] = g2 - n2.s * (n2.cx * g0 + n2.cy * g1);
    }
    double ss = 0.0;
    for (int i = 0; i < 9; ++i) ss += F[i] * F[i];
    if (!(ss > 0.0) || !(ss < 1e300)) return 0;
    double inv = 1.0 / sqrt(ss);
    for (int i = 0; i < 9; ++i) F[i] *= inv;
    return 1;
}

/* Real roots of c3 x^3 + c2 x^2 + c1 x + c0 using only + - * / sqrt. */
static int solve_cubic(double c3, double c2, double c1, double c0, double* roots)
{
    double mx = fmax(fabs(c2), fmax(fabs(c1), fabs(c0)));
    int n = 0;
    if (!(fabs(c3) > 1e-14 * mx)) {
        /* degenerate: quadratic c2 x^2 + c1 x + c0 */
        if (!(fabs(c2) > 1e-14 * fmax(fabs(c1), fabs(c0)))) {
            if (c1 != 0.0) roots[n++] = -c0 / c1;
            return n;
        }
        double disc = c1 * c1 - 4.0 * c2 * c0;
        if (disc < 0.0) return 0;
        double sq = sqrt(disc);
        double q = -0.5 * (c1 + (c1 >= 0.0 ? sq : -sq));
        roots[n++] = q / c2;
        if (q != 0.0) roots[n++] = c0 / q;
        return n;
    }
    double b = c2 / c3, c = c1 / c3, d = c0 / c3;
    double R = 1.0 + fmax(fabs(b), fmax(fabs(c), fabs(d)));
    double lo = -R, hi = R;
    for (int it = 0; it < 80; ++it) {
        double mid = 0.5 * (lo + hi);
        double fm = ((mid + b) * mid + c) * mid + d;
        if (fm < 0.0) lo = mid; else hi = mid;
    }
    double r = 0.5 * (lo + hi);
    for (int it = 0; it < 2; ++it) {
        double f = ((r + b) * r + c) * r + d;
        double fp = (3.0 * r + 2.0 * b) * r + c;
        if (fp != 0.0) r = r - f / fp;
    }
    roots[n++] = r;
    double B = b + r;
    double C = c + r * B;
    double disc = B * B - 4.0 * C;
    if (disc >= 0.0) {
        double sq = sqrt(disc);
        double q = -0.5 * (B + (B >= 0.0 ? sq : -sq));
        double r2 = q;
        double r3 = (q != 0.0) ? C / q : q;
        for (int it = 0; it < 2; ++it) {
            double f = ((r2 + b) * r2 + c) * r2 + d;
            double fp = (3.0 * r2 + 2.0 * b) * r2 + c;
            if (fp != 0.0) r2 = r2 - f / fp;
            f = ((r3 + b) * r3 + c) * r3 + d;
            fp = (3.0 * r3 + 2.0 * b) * r3 + c;
            if (fp != 0.0) r3 = r3 - f / fp;
        }
        roots[n++] = r2;
        roots[n++] = r3;
    }
    return n;
}

/* Minimal solver: m = 7 or 8 sample points -> up to 3 unit-norm F (row-major). */
static int solve_minimal(const float* corr, const int* idx, int m, double* Fout)
{
    double x1[8], y1[8], x2[8], y2[8];
    for (int k = 0; k < m; ++k) {
        const float* c = corr + 4 * (size_t)idx[k];
        x1[k] = (double)c[0]; y1[k] = (double)c[1];
        x2[k] = (double)c[2]; y2[k] = (double)c[3];
    }
    norm2d n1, n2;
    {
        double sx = 0.0, sy = 0.0, tx = 0.0, ty = 0.0;
        for (int k = 0; k < m; ++k) { sx += x1[k]; sy += y1[k]; tx += x2[k]; ty += y2[k]; }
        double inv = 1.0 / (double)m;
        n1.cx = sx * inv; n1.cy = sy * inv; n2.cx = tx * inv; n2.cy = ty * inv;
        double d1 = 0.0, d2 = 0.0;
        for (int k = 0; k < m; ++k) {
            double ax = x1[k] - n1.cx, ay = y1[k] - n1.cy;
            double bx = x2[k] - n2.cx, by = y2[k] - n2.cy;
            d1 += sqrt(ax * ax + ay * ay);
            d2 += sqrt(bx * bx + by * by);
        }
        d1 *= inv; d2 *= inv;
        if (!(d1 > 1e-9) || !(d2 > 1e-9)) return 0;
        n1.s = 1.4142135623730951 / d1;
        n2.s = 1.4142135623730951 / d2;
    }
    double A[8][9];
    for (int k = 0; k < m; ++k) {
        double u1 = (x1[k] - n1.cx) * n1.s, v1 = (y1[k] - n1.cy) * n1.s;
        double u2 = (x2[k] - n2.cx) * n2.s, v2 = (y2[k] - n2.cy) * n2.s;
        A[k][0] = u2 * u1; A[k][1] = u2 * v1; A[k][2] = u2;
        A[k][3] = v2 * u1; A[k][4] = v2 * v1; A[k][5] = v2;
        A[k][6] = u1;      A[k][7] = v1;      A[k][8] = 1.0;
    }
    /* Gauss-Jordan with complete pivoting */
    int perm[9];
    for (int j = 0; j < 9; ++j) perm[j] = j;
    for (int k = 0; k < m; ++k) {
        int pi = k, pj = k;
        double best = -1.0;
        for (int i = k; i < m; ++i)
            for (int j = k; j < 9; ++j) {
                double v = fabs(A[i][j]);
                if (v > best) { best = v; pi = i; pj = j; }
            }
        if (!(best > 1e-12)) return 0;
        if (pi != k)
            for (int j = 0; j < 9; ++j) { double t = A[k][j]; A[k][j] = A[pi][j]; A[pi][j] = t; }
        if (pj != k) {
            for (int i = 0; i < m; ++i) { double t = A[i][k]; A[i][k] = A[i][pj]; A[i][pj] = t; }
            int t = perm[k]; perm[k] = perm[pj]; perm[pj] = t;
        }
        double inv = 1.0 / A[k][k];
        for (int j = k; j < 9; ++j) A[k][j] *= inv;
        for (int i = 0; i < m; ++i) {
            if (i == k) continue;
            double f = A[i][k];
            for (int j = k; j < 9; ++j) A[i][j] -= f * A[k][j];
        }
    }
    double N[2][9];
    int nfree = 9 - m;
    for (int c = 0; c < nfree; ++c) {
        for (int j = 0; j < 9; ++j) N[c][j] = 0.0;
        N[c][perm[m + c]] = 1.0;
        for (int k = 0; k < m; ++k) N[c][perm[k]] = -A[k][m + c];
    }
    if (m == 8) {
        double Fh[9];
        for (int i = 0; i < 9; ++i) Fh[i] = N[0][i];
        enforce_rank2(Fh);
        return denormalise(Fh, n1, n2, Fout);
    }
    /* 7 point: det(l*F1 + F2) = 0 */
    const double* F1 = N[0];
    const double* F2 = N[1];
    double c3 = det3(F1, F1 + 3, F1 + 6);
    double c0 = det3(F2, F2 + 3, F2 + 6);
    double c2 = det3(F2, F1 + 3, F1 + 6) + det3(F1, F2 + 3, F1 + 6) + det3(F1, F1 + 3, F2 + 6);
    double c1 = det3(F1, F2 + 3, F2 + 6) + det3(F2, F1 + 3, F2 + 6) + det3(F2, F2 + 3, F1 + 6);
    double roots[3];
    int nr = solve_cubic(c3, c2, c1, c0, roots);
    int nm = 0;
    for (int r = 0; r < nr; ++r) {
        double Fh[9];
        for (int i = 0; i < 9; ++i) Fh[i] = roots[r] * F1[i] + F2[i];
        if (denormalise(Fh, n1, n2, Fout + 9 * nm)) ++nm;
    }
    return nm;
}

/* ------------------------------------------------------------------ scoring */
static int is_inlier(const float* Ff, const float* c, float thr2, int score)
{
    float x1 = c[0], y1 = c[1], x2 = c[2], y2 = c[3];
    float a = fmaf(Ff[0], x1, fmaf(Ff[1], y1, Ff[2]));
    float b = fmaf(Ff[3], x1, fmaf(Ff[4], y1, Ff[5]));
    float cc = fmaf(Ff[6], x1, fmaf(Ff[7], y1, Ff[8]));
    float num = fmaf(x2, a, fmaf(y2, b, cc));
    float a2 = fmaf(Ff[0], x2, fmaf(Ff[3], y2, Ff[6]));
    float b2 = fmaf(Ff[1], x2, fmaf(Ff[4], y2, Ff[7]));
    float bb = b * b;
    float s1 = fmaf(a, a, bb);
    float bb2 = b2 * b2;
    float s2 = fmaf(a2, a2, bb2);
    float n2 = num * num;
    float lim = (score == SFM_SCORE_SAMPSON) ? thr2 * (s1 + s2) : thr2 * fminf(s1, s2);
    return n2 <= lim;
}

static int count_inliers(const double* F, const float* corr, int M, float thr2, int score, uint8_t* mask)
{
    float Ff[9];
    for (int i = 0; i < 9; ++i) Ff[i] = (float)F[i];
    int n = 0;
    for (int i = 0; i < M; ++i) {
        int in = is_inlier(Ff, corr + 4 * (size_t)i, thr2, score);
        if (mask) mask[i] = (uint8_t)in;
        n += in;
    }
    return n;
}

/* -------------------------------------------------------------- LO refit */
static int lo_refit(const float* corr, int M, const uint8_t* mask, double* F)
{
    double part[SFM_RANSAC_LANES];
    double mom[5];
    for (int q = 0; q < 5; ++q) {
        for (int t = 0; t < SFM_RANSAC_LANES; ++t) {
            double s = 0.0;
            for (int i = t; i < M; i += SFM_RANSAC_LANES)
                if (mask[i]) s += (q == 4) ? 1.0 : (double)corr[4 * (size_t)i + q];
            part[t] = s;
        }
        mom[q] = lane_tree(part);
    }
    if (!(mom[4] >= 8.0)) return 0;
    norm2d n1, n2;
    double inv = 1.0 / mom[4];
    n1.cx = mom[0] * inv; n1.cy = mom[1] * inv; n2.cx = mom[2] * inv; n2.cy = mom[3] * inv;
    double dd[2];
    for (int q = 0; q < 2; ++q) {
        double cx = q ? n2.cx : n1.cx, cy = q ? n2.cy : n1.cy;
        for (int t = 0; t < SFM_RANSAC_LANES; ++t) {
            double s = 0.0;
            for (int i = t; i < M; i += SFM_RANSAC_LANES)
                if (mask[i]) {
                    double ax = (double)corr[4 * (size_t)i + 2 * q] - cx;
                    double ay = (double)corr[4 * (size_t)i + 2 * q + 1] - cy;
                    s += sqrt(ax * ax + ay * ay);
                }
            part[t] = s;
        }
        dd[q] = lane_tree(part) * inv;
    }
    if (!(dd[0] > 1e-9) || !(dd[1] > 1e-9)) return 0;
    n1.s = 1.4142135623730951 / dd[0];
    n2.s = 1.4142135623730951 / dd[1];
    double AtA[81];
    for (int a = 0; a < 9; ++a)
        for (int b = a; b < 9; ++b) {
            for (int t = 0; t < SFM_RANSAC_LANES; ++t) {
                double s = 0.0;
                for (int i = t; i < M; i += SFM_RANSAC_LANES)
                    if (mask[i]) {
                        const float* c = corr + 4 * (size_t)i;
                        double u1 = ((double)c[0] - n1.cx) * n1.s, v1 = ((double)c[1] - n1.cy) * n1.s;
                        double u2 = ((double)c[2] - n2.cx) * n2.s, v2 = ((double)c[3] - n2.cy) * n2.s;
                        double r[9] = {u2 * u1, u2 * v1, u2, v2 * u1, v2 * v1, v2, u1, v1, 1.0};
                        s += r[a] * r[b];
                    }
                part[t] = s;
            }
            double v = lane_tree(part);
            AtA[a * 9 + b] = v;
            AtA[b * 9 + a] = v;
        }
    double Fh[9];
    if (!smallest_eigvec9(AtA, Fh)) return 0;
    enforce_rank2(Fh);
    return denormalise(Fh, n1, n2, F);
}

/* --------------------------------------------------------------- driver */
int sfm_oracle_ransac_f(const float* corr, int M, const sfm_ransac_params* prm, uint32_t pair_id,
                        const uint32_t* samples, double* out_F, int32_t* out_ninl,
                        uint8_t* out_mask, int32_t* out_iters)
{
    const int m = prm->solver;
    const float thr2 = prm->threshold * prm->threshold;
    for (int i = 0; i < 9; ++i) out_F[i] = 0.0;
    if (out_mask) memset(out_mask, 0, (size_t)(M > 0 ? M : 0));
    *out_ninl = 0;
    if (out_iters) *out_iters = 0;
    if ((m != 7 && m != 8) || M < m) return 0;

    double bestF[9];
    int best = 0, done = 0;
    while (done < prm->max_iters) {
        int nb = prm->max_iters - done;
        if (nb > ransac_batch(done)) nb = ransac_batch(done);
        for (int h = 0; h < nb; ++h) {
            int idx[8];
            if (samples) {
                for (int k = 0; k < m; ++k) {
                    uint32_t v = samples[(size_t)(done + h) * 8 + k];
                    idx[k] = (int)(v % (uint32_t)M);
                }
            } else {
                draw_sample(prm->seed, pair_id, (uint32_t)(done + h), m, M, idx);
            }
            double Fm[27];
            int nm = solve_minimal(corr, idx, m, Fm);
            for (int r = 0; r < nm; ++r) {
                int cnt = count_inliers(Fm + 9 * r, corr, M, thr2, prm->score, NULL);
                if (cnt > best) { best = cnt; memcpy(bestF, Fm + 9 * r, sizeof bestF); }
            }
        }
        done += nb;
        if (should_stop(best, M, m, done, prm->confidence)) break;
    }
    if (out_iters) *out_iters = done;
    if (best < m) return 0;

    uint8_t* mask = out_mask ? out_mask : (uint8_t*)malloc((size_t)M);
    best = count_inliers(bestF, corr, M, thr2, prm->score, mask);
    if (prm->lo_refit) {
        uint8_t* trial = (uint8_t*)malloc((size_t)M);
        for (int round = 0; round < SFM_LO_ROUNDS; ++round) {
            double Fr[9];
            if (!lo_refit(corr, M, mask, Fr)) break;
            int cnt = count_inliers(Fr, corr, M, thr2, prm->score, trial);
            if (cnt <= best) break;
            best = cnt;
            memcpy(bestF, Fr, sizeof bestF);
            memcpy(mask, trial, (size_t)M);
        }
        free(trial);
    }
    if (prm->min_inliers > 0 && best < prm->min_inliers) {
        memset(mask, 0, (size_t)M);
        if (!out_mask) free(mask);
        return 0;
    }
    if (!out_mask) free(mask);
    /* cv2 convention F[2,2] == 1.0 exactly: divide (x / x == 1), never multiply by a reciprocal */
    const double s = (fabs(bestF[8]) > 1.1920928955078125e-07) ? bestF[8] : 1.0;
    for (int i = 0; i < 9; ++i) out_F[i] = bestF[i] / s;
    *out_ninl = best;
    return 0;
}

/* Expose the pieces so tests can pin them individually. */
int sfm_oracle_solve_minimal(const float* corr, const int32_t* idx, int m, double* Fout)
{
    int id[8];
    for (int k = 0; k < m; ++k) id[k] = idx[k];
    return solve_minimal(corr, id, m, Fout);
}

int sfm_oracle_count_inliers(const double* F, const float* corr, int M, float thr, int score, uint8_t* mask)
{
    return count_inliers(F, corr, M, thr * thr, score, mask);
}

void sfm_oracle_draw_sample(uint64_t seed, uint32_t pair, uint32_t hyp, int m, int M, int32_t* idx)
{
    int id[8];
    draw_sample(seed, pair, hyp, m, M, id);
    for (int k = 0; k < m; ++k) idx[k] = id[k];
}

int sfm_oracle_smallest_eigvec9(const double* A, double* x) { return smallest_eigvec9(A, x); }
