/* oracle/qr_colpiv.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * fp32 column-pivoted Householder QR solve of a small dense system, restating the
 * published algorithm of Eigen 3.4.0 `ColPivHouseholderQR` (the third-party
 * dependency the reference calls at correlation_class.cpp:742-747:
 *   Eigen::Map<MatrixXf>(mat_A,n,n).colPivHouseholderQr().solve(Map<VectorXf>(vec_B,n))
 * Eigen 3.4.0 is pinned only by prose, README.md:22, and is NOT vendored under
 * /root/reference nor installed in this image).
 *
 * Restated steps (Eigen/src/QR/ColPivHouseholderQR.h, Eigen/src/Householder/Householder.h):
 *   - column norms (direct + updated tables), pivot = largest updated norm,
 *   - nonzero-pivot bookkeeping with threshold_helper = (maxnorm*eps)^2 / rows,
 *   - makeHouseholderInPlace / applyHouseholderOnTheLeft,
 *   - LAPACK-style (LAWN 176) norm downdate with sqrt(eps) recompute threshold,
 *   - solve: c = Q^T b over the first `nonzero_pivots` reflectors, back-substitution
 *     on the leading triangle, zero for the discarded columns, un-permute.
 *
 * PARITY UNPINNED at this boundary: the reference holds no test vector for the
 * solve, and Eigen's internal SIMD summation order is not reproduced (results
 * agree to a few fp32 ulps * cond(A), which the next Gauss-Newton step absorbs).
 * The same header backs the `Dense` shim used to compile the unmodified
 * reference (oracle/shim/Dense) and the C restatement (oracle/dic_oracle.c), so
 * those two agree bit-for-bit with each other.
 *
 * Storage: column-major n x n (the reference maps its row-major symmetric
 * matrix through a column-major Eigen::Map; A is symmetric so it is the same).
 */
#ifndef ORACLE_QR_COLPIV_H
#define ORACLE_QR_COLPIV_H

#include <float.h>
#include <math.h>

#define ORACLE_QR_MAXN 16

/* Solves A x = b. A (column-major, n x n) is copied, not modified. */
static inline void oracle_qr_colpiv_solve(const float *A_in, const float *b_in,
                                          float *x_out, int n) {
  float qr[ORACLE_QR_MAXN * ORACLE_QR_MAXN];
  float hcoef[ORACLE_QR_MAXN];
  float norms_upd[ORACLE_QR_MAXN], norms_dir[ORACLE_QR_MAXN];
  float tmp[ORACLE_QR_MAXN];
  float c[ORACLE_QR_MAXN];
  int perm[ORACLE_QR_MAXN];
  const float eps = FLT_EPSILON;
  int rows = n, cols = n, size = n;
  int i, j, k;

#define QR(r, cc) qr[(cc) * n + (r)]
  for (i = 0; i < n * n; ++i) qr[i] = A_in[i];
  for (i = 0; i < n; ++i) perm[i] = i;

  float maxnorm = 0.f;
  for (k = 0; k < cols; ++k) {
    float s = 0.f;
    for (i = 0; i < rows; ++i) s += QR(i, k) * QR(i, k);
    norms_dir[k] = sqrtf(s);
    norms_upd[k] = norms_dir[k];
    if (k == 0 || norms_upd[k] > maxnorm) maxnorm = norms_upd[k];
  }
  float th = maxnorm * eps;
  float threshold_helper = (th * th) / (float)rows;
  float norm_downdate_threshold = sqrtf(eps);
  int nonzero_pivots = size;

  for (k = 0; k < size; ++k) {
    int big = k;
    float bigv = norms_upd[k];
    for (j = k + 1; j < cols; ++j)
      if (norms_upd[j] > bigv) { bigv = norms_upd[j]; big = j; }
    float big_sq = bigv * bigv;
    if (nonzero_pivots == size && big_sq < threshold_helper * (float)(rows - k))
      nonzero_pivots = k;
    if (big != k) {
      for (i = 0; i < rows; ++i) { float t = QR(i, k); QR(i, k) = QR(i, big); QR(i, big) = t; }
      { float t = norms_upd[k]; norms_upd[k] = norms_upd[big]; norms_upd[big] = t; }
      { float t = norms_dir[k]; norms_dir[k] = norms_dir[big]; norms_dir[big] = t; }
      { int t = perm[k]; perm[k] = perm[big]; perm[big] = t; }
    }
    /* makeHouseholderInPlace on qr(k:rows-1, k) */
    float tail_sq = 0.f;
    for (i = k + 1; i < rows; ++i) tail_sq += QR(i, k) * QR(i, k);
    float c0 = QR(k, k);
    float tau, beta;
    if (tail_sq <= FLT_MIN) {
      tau = 0.f;
      beta = c0;
      for (i = k + 1; i < rows; ++i) QR(i, k) = 0.f;
    } else {
      beta = sqrtf(c0 * c0 + tail_sq);
      if (c0 >= 0.f) beta = -beta;
      float denom = c0 - beta;
      for (i = k + 1; i < rows; ++i) QR(i, k) = QR(i, k) / denom;
      tau = (beta - c0) / beta;
    }
    hcoef[k] = tau;
    QR(k, k) = beta;
    /* applyHouseholderOnTheLeft on the bottom-right corner */
    if (k + 1 < cols) {
      if (rows - k == 1) {
        for (j = k + 1; j < cols; ++j) QR(k, j) *= (1.f - tau);
      } else if (tau != 0.f) {
        for (j = k + 1; j < cols; ++j) {
          float s = 0.f;
          for (i = k + 1; i < rows; ++i) s += QR(i, k) * QR(i, j);
          tmp[j] = s + QR(k, j);
        }
        for (j = k + 1; j < cols; ++j) QR(k, j) -= tau * tmp[j];
        for (j = k + 1; j < cols; ++j)
          for (i = k + 1; i < rows; ++i) QR(i, j) -= tau * QR(i, k) * tmp[j];
      }
    }
    /* norm downdate */
    for (j = k + 1; j < cols; ++j) {
      if (norms_upd[j] != 0.f) {
        float t = fabsf(QR(k, j)) / norms_upd[j];
        t = (1.f + t) * (1.f - t);
        t = t < 0.f ? 0.f : t;
        float r = norms_upd[j] / norms_dir[j];
        float t2 = t * (r * r);
        if (t2 <= norm_downdate_threshold) {
          float s = 0.f;
          for (i = k + 1; i < rows; ++i) s += QR(i, j) * QR(i, j);
          norms_dir[j] = sqrtf(s);
          norms_upd[j] = norms_dir[j];
        } else {
          norms_upd[j] *= sqrtf(t);
        }
      }
    }
  }

  /* solve */
  if (nonzero_pivots == 0) {
    for (i = 0; i < n; ++i) x_out[i] = 0.f;
    return;
  }
  for (i = 0; i < n; ++i) c[i] = b_in[i];
  for (k = 0; k < nonzero_pivots; ++k) {
    float tau = hcoef[k];
    if (rows - k == 1) {
      c[k] *= (1.f - tau);
    } else if (tau != 0.f) {
      float s = 0.f;
      for (i = k + 1; i < rows; ++i) s += QR(i, k) * c[i];
      s += c[k];
      c[k] -= tau * s;
      for (i = k + 1; i < rows; ++i) c[i] -= tau * QR(i, k) * s;
    }
  }
  for (i = nonzero_pivots - 1; i >= 0; --i) {
    float s = c[i];
    for (j = i + 1; j < nonzero_pivots; ++j) s -= QR(i, j) * c[j];
    c[i] = s / QR(i, i);
  }
  for (i = 0; i < nonzero_pivots; ++i) x_out[perm[i]] = c[i];
  for (i = nonzero_pivots; i < cols; ++i) x_out[perm[i]] = 0.f;
#undef QR
}

#endif /* ORACLE_QR_COLPIV_H */
