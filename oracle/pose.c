/* pose.c -- CPU oracle for the batched two-view initialisation (E from F, pose
 * recovery by cheirality vote, DLT triangulation).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sfm-project_b200/ links, imports or
 * executes this file; only tests/ and __graft_entry__.smoke() may.
 *
 * What it restates.  code/3d_reconstruction.py in the reference is a 0-byte file
 * (imported at code/pipeline.py:4); SURVEY.md section 8f rank 4 names two-view
 * initialisation as the consumer of the hot path's (F, inlier mask).  Conventions
 * are those of cv2.recoverPose(E, pts1, pts2, K, distanceThresh) and
 * cv2.triangulatePoints (OpenCV 4.13.0, third party, unpinned by the reference):
 * x2 ~ R x1 + t, |t| = 1, candidates (R1,t) (R2,t) (R1,-t) (R2,-t), ties to the
 * earlier, a point votes iff its depth is in (0, dist) in both cameras.  Pinned
 * against cv2 run in this image by tests/test_oracle_pinned.py (R, t within
 * 1e-6, points within 1e-5 relative).  PARITY UNPINNED by the reference.
 *
 * Operation for operation the same as sfm-project_b200/csrc/pose.cu; built with
 * -ffp-contract=off so R, t, votes, masks and points are compared bit for bit.
 */
#include "ransac_common.h"

static void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

static int decompose_essential(const double* E, double* R1, double* R2, double* t)
{
    double G[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i * 3 + j] = E[0 + i] * E[0 + j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    jacobi_eig(G, V, 3, 8);
    int i0 = 0;
    if (G[4] > G[i0 * 4]) i0 = 1;
    if (G[8] > G[i0 * 4]) i0 = 2;
    int i2 = 0;
    if (G[4] < G[i2 * 4]) i2 = 1;
    if (G[8] < G[i2 * 4]) i2 = 2;
    if (i0 == i2) return 0;
    int i1 = 3 - i0 - i2;
    double v0[3] = {V[0 + i0], V[3 + i0], V[6 + i0]};
    double v1[3] = {V[0 + i1], V[3 + i1], V[6 + i1]};
    double v2[3];
    cross3(v0, v1, v2);
    double u0[3], u1[3], u2[3];
    for (int r = 0; r < 3; ++r) {
        u0[r] = E[r * 3 + 0] * v0[0] + E[r * 3 + 1] * v0[1] + E[r * 3 + 2] * v0[2];
        u1[r] = E[r * 3 + 0] * v1[0] + E[r * 3 + 1] * v1[1] + E[r * 3 + 2] * v1[2];
    }
    double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
    if (!(n0 > 1e-12)) return 0;
    for (int r = 0; r < 3; ++r) u0[r] = u0[r] / n0;
    double d = u1[0] * u0[0] + u1[1] * u0[1] + u1[2] * u0[2];
    for (int r = 0; r < 3; ++r) u1[r] = u1[r] - d * u0[r];
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (!(n1 > 1e-9)) return 0;
    for (int r = 0; r < 3; ++r) u1[r] = u1[r] / n1;
    cross3(u0, u1, u2);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double a = u1[r] * v0[c] - u0[r] * v1[c];
            double b = u2[r] * v2[c];
            R1[r * 3 + c] = b + a;
            R2[r * 3 + c] = b - a;
        }
    t[0] = u2[0]; t[1] = u2[1]; t[2] = u2[2];
    return 1;
}

static int triangulate(const double* R, const double* t, double x1, double y1, double x2, double y2, double dist, double* X)
{
    double A[4][4];
    A[0][0] = -1.0; A[0][1] = 0.0;  A[0][2] = x1; A[0][3] = 0.0;
    A[1][0] = 0.0;  A[1][1] = -1.0; A[1][2] = y1; A[1][3] = 0.0;
    for (int c = 0; c < 3; ++c) {
        A[2][c] = x2 * R[6 + c] - R[0 + c];
        A[3][c] = y2 * R[6 + c] - R[3 + c];
    }
    A[2][3] = x2 * t[2] - t[0];
    A[3][3] = y2 * t[2] - t[1];
    double G[16], V[16];
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            double s = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j] + A[3][i] * A[3][j];
            G[i * 4 + j] = s;
            G[j * 4 + i] = s;
        }
    jacobi_eig(G, V, 4, 6);
    int k = 0;
    for (int j = 1; j < 4; ++j)
        if (G[j * 5] < G[k * 5]) k = j;
    double w = V[12 + k];
    X[0] = 0.0; X[1] = 0.0; X[2] = 0.0;
    if (!(fabs(w) > 1e-300)) return 0;
    X[0] = V[0 + k] / w;
    X[1] = V[4 + k] / w;
    X[2] = V[8 + k] / w;
    double z2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    return (X[2] > 0.0 && X[2] < dist && z2 > 0.0 && z2 < dist) ? 1 : 0;
}

/* corr [M,4], in_mask [M] or NULL, F [9], cam [8] = fx1 fy1 cx1 cy1 fx2 fy2 cx2 cy2.
 * out_R [9], out_t [3], out_E [9] (may be NULL), out_mask [M], out_X float [M,3] (may be NULL).
 * Returns the vote count of the chosen candidate (0 = no pose). */
int sfm_oracle_two_view_pose(const float* corr, int M, const uint8_t* in_mask, const double* F, const double* cam, double dist,
                             double* out_R, double* out_t, double* out_E, uint8_t* out_mask, float* out_X)
{
    const double fx1 = cam[0], fy1 = cam[1], cx1 = cam[2], cy1 = cam[3];
    const double fx2 = cam[4], fy2 = cam[5], cx2 = cam[6], cy2 = cam[7];
    for (int i = 0; i < 9; ++i) { out_R[i] = 0.0; if (out_E) out_E[i] = 0.0; }
    for (int i = 0; i < 3; ++i) out_t[i] = 0.0;
    if (M > 0) memset(out_mask, 0, (size_t)M);
    if (out_X && M > 0) memset(out_X, 0, sizeof(float) * 3 * (size_t)M);
    double G[9], E[9];
    for (int r = 0; r < 3; ++r) {
        G[r * 3 + 0] = F[r * 3 + 0] * fx1;
        G[r * 3 + 1] = F[r * 3 + 1] * fy1;
        G[r * 3 + 2] = F[r * 3 + 0] * cx1 + F[r * 3 + 1] * cy1 + F[r * 3 + 2];
    }
    for (int c = 0; c < 3; ++c) {
        E[0 + c] = fx2 * G[0 + c];
        E[3 + c] = fy2 * G[3 + c];
        E[6 + c] = cx2 * G[0 + c] + cy2 * G[3 + c] + G[6 + c];
    }
    double ss = 0.0;
    for (int i = 0; i < 9; ++i) ss += E[i] * E[i];
    if (!((ss > 0.0) && (ss < 1e300) && M > 0)) return 0;
    double inv = 1.0 / sqrt(ss);
    for (int i = 0; i < 9; ++i) E[i] *= inv;
    double Rc[2][9], tp[3], tn[3];
    if (!decompose_essential(E, Rc[0], Rc[1], tp)) return 0;
    for (int i = 0; i < 3; ++i) tn[i] = -tp[i];
    int votes[4] = {0, 0, 0, 0};
    for (int i = 0; i < M; ++i) {
        if (in_mask && !in_mask[i]) continue;
        const float* c = corr + 4 * (size_t)i;
        double x1 = ((double)c[0] - cx1) / fx1, y1 = ((double)c[1] - cy1) / fy1;
        double x2 = ((double)c[2] - cx2) / fx2, y2 = ((double)c[3] - cy2) / fy2;
        double X[3];
        votes[0] += triangulate(Rc[0], tp, x1, y1, x2, y2, dist, X);
        votes[1] += triangulate(Rc[1], tp, x1, y1, x2, y2, dist, X);
        votes[2] += triangulate(Rc[0], tn, x1, y1, x2, y2, dist, X);
        votes[3] += triangulate(Rc[1], tn, x1, y1, x2, y2, dist, X);
    }
    int best = 0;
    for (int k = 1; k < 4; ++k)
        if (votes[k] > votes[best]) best = k;
    const double* R = Rc[best & 1];
    const double* t = (best & 2) ? tn : tp;
    for (int i = 0; i < M; ++i) {
        if (in_mask && !in_mask[i]) continue;
        const float* c = corr + 4 * (size_t)i;
        double x1 = ((double)c[0] - cx1) / fx1, y1 = ((double)c[1] - cy1) / fy1;
        double x2 = ((double)c[2] - cx2) / fx2, y2 = ((double)c[3] - cy2) / fy2;
        double X[3];
        int good = triangulate(R, t, x1, y1, x2, y2, dist, X);
        out_mask[i] = (uint8_t)good;
        if (out_X && good) { out_X[3 * i + 0] = (float)X[0]; out_X[3 * i + 1] = (float)X[1]; out_X[3 * i + 2] = (float)X[2]; }
    }
    for (int i = 0; i < 9; ++i) { out_R[i] = R[i]; if (out_E) out_E[i] = E[i]; }
    for (int i = 0; i < 3; ++i) out_t[i] = t[i];
    return votes[best];
}

/* exposed so that tests can pin the mirror property the CUDA kernel's vote relies on */
int sfm_oracle_triangulate(const double* R, const double* t, const double* xy4, double dist, double* X)
{
    return triangulate(R, t, xy4[0], xy4[1], xy4[2], xy4[3], dist, X);
}
