"""Kalman filter of the reference tracker, restated as four independent 2x2 blocks
with explicit float32 rounding (oracle; test infrastructure).

Follows ``/root/reference/src/tracker/core/kalman_filter.py``:
  initiate         :55-83
  predict          :85-120
  project          :122-151
  update           :153-204
  gating_distance  :206-249

The reference multiplies dense 8x8 / 4x8 float32 matrices through BLAS and
solves with LAPACK/BLAS triangular solves.  The only entries of the covariance
that are ever non-zero are, per coordinate i in (cx, cy, a, h):
    a = P[i,i]   b = P[i,i+4]   c = P[i+4,i]   d = P[i+4,i+4]
so every BLAS sum has at most two non-zero terms and the result is independent
of summation order.  This file spells the surviving float32 operations out one
by one (numpy elementwise ufuncs never contract a*b+c into an FMA), which makes
the oracle reproducible on any host CPU, unlike a BLAS call.

Scalar-promotion semantics are those of numpy >= 2 (NEP 50), the numpy the
oracle is pinned under: ``python_float * np.float32`` is a float32 product.

State layout used here and by the CUDA kernels:
    mean[8]  float32   (cx, cy, a, h, vcx, vcy, va, vh)
    cov[16]  float32   (a0..a3, b0..b3, c0..c3, d0..d3)
"""
import numpy as np

F32 = np.float32

# python doubles rounded once to float32 when they meet a float32 operand
WP = F32(1.0 / 20)          # _std_weight_position            kalman_filter.py:52
WV = F32(1.0 / 160)         # _std_weight_velocity            kalman_filter.py:53
WP2 = F32(2 * (1.0 / 20))   # 2 * _std_weight_position        kalman_filter.py:73
WV10 = F32(10 * (1.0 / 160))  # 10 * _std_weight_velocity     kalman_filter.py:77

# constants that stay python doubles inside the std list and are squared in float64
SQ_1E2 = F32(1e-2 * 1e-2)   # aspect-ratio position std, initiate/predict   :75,:102
SQ_1E5 = F32(1e-5 * 1e-5)   # aspect-ratio velocity std                     :79,:108
SQ_1E1 = F32(1e-1 * 1e-1)   # aspect-ratio measurement std, project         :139

CHI2_GATE = 9.487729036781154


def _sq(x):
    """np.square on a float64 array holding float32 values, then astype(float32)."""
    x = np.asarray(x, dtype=F32)
    return (x.astype(np.float64) * x.astype(np.float64)).astype(F32)


def initiate(z):
    """kalman_filter.py:55-83.  z = (cx, cy, a, h) float32 -> (mean[8], cov[16])."""
    z = np.asarray(z, dtype=F32)
    mean = np.zeros(8, F32)
    mean[:4] = z
    p = WP2 * z[3]
    v = WV10 * z[3]
    cov = np.zeros(16, F32)
    sp, sv = _sq(p), _sq(v)
    cov[0:4] = (sp, sp, SQ_1E2, sp)     # a
    cov[12:16] = (sv, sv, SQ_1E5, sv)   # d
    return mean, cov


def predict(mean, cov):
    """kalman_filter.py:85-120 for a batch: mean (T,8), cov (T,16) -> new (mean, cov).

    P' = F (P F^T) + Q with F = [[I, I],[0, I]] (np.linalg.multi_dot associates
    right-first for three equal-cost operands)."""
    mean = np.asarray(mean, F32).reshape(-1, 8)
    cov = np.asarray(cov, F32).reshape(-1, 16)
    a, b, c, d = cov[:, 0:4], cov[:, 4:8], cov[:, 8:12], cov[:, 12:16]
    h = mean[:, 3]
    sp = _sq(WP * h)
    sv = _sq(WV * h)
    qp = np.stack([sp, sp, np.full_like(sp, SQ_1E2), sp], axis=1)
    qv = np.stack([sv, sv, np.full_like(sv, SQ_1E5), sv], axis=1)
    new_mean = mean.copy()
    new_mean[:, :4] = mean[:, :4] + mean[:, 4:]
    y00 = a + b
    y10 = c + d
    na = (y00 + y10) + qp
    nb = b + d
    nc = y10
    nd = d + qv
    return new_mean, np.concatenate([na, nb, nc, nd], axis=1).astype(F32)


def innovation_diag(mean, cov):
    """Diagonal of S = H P H^T + R, kalman_filter.py:122-151.  (T,4) float32."""
    mean = np.asarray(mean, F32).reshape(-1, 8)
    cov = np.asarray(cov, F32).reshape(-1, 16)
    r = _sq(WP * mean[:, 3])
    rr = np.stack([r, r, np.full_like(r, SQ_1E1), r], axis=1)
    return cov[:, 0:4] + rr


def gating_distance(mean, cov, measurements):
    """kalman_filter.py:206-249 for ONE track against N measurements (N,4).

    S is diagonal, so the Cholesky factor is sqrt of the diagonal.  With N >= 2
    right-hand sides OpenBLAS strsm multiplies by a float32 reciprocal of the
    diagonal; with exactly one right-hand side the solve divides (verified
    against the reference in tests/golden/make_golden.py)."""
    z = np.asarray(measurements, F32).reshape(-1, 4)
    s = innovation_diag(mean, cov)[0]
    L = np.sqrt(s)
    delta = z - np.asarray(mean, F32).reshape(-1, 8)[0, :4]
    if z.shape[0] >= 2:
        y = delta * (F32(1.0) / L)
    else:
        y = delta / L
    q = y * y
    return ((q[:, 0] + q[:, 1]) + q[:, 2]) + q[:, 3]


def update(mean, cov, z):
    """kalman_filter.py:153-204 for ONE track.  Returns (mean[8], cov[16])."""
    mean = np.asarray(mean, F32).reshape(8).copy()
    cov = np.asarray(cov, F32).reshape(16)
    z = np.asarray(z, F32).reshape(4)
    a, b, c, d = cov[0:4], cov[4:8], cov[8:12], cov[12:16]
    s = innovation_diag(mean, cov)[0]
    inv = F32(1.0) / np.sqrt(s)
    k0 = (a * inv) * inv
    k1 = (c * inv) * inv
    e = z - mean[:4]
    new_mean = mean.copy()
    new_mean[:4] = mean[:4] + k0 * e
    new_mean[4:] = mean[4:] + k1 * e
    s0 = s * k0
    s1 = s * k1
    na = a - k0 * s0
    nb = b - k0 * s1
    nc = c - k1 * s0
    nd = d - k1 * s1
    return new_mean.astype(F32), np.concatenate([na, nb, nc, nd]).astype(F32)


def cov_to_dense(cov):
    """(16,) block storage -> dense 8x8, for comparisons with the reference."""
    cov = np.asarray(cov, F32).reshape(16)
    P = np.zeros((8, 8), F32)
    for i in range(4):
        P[i, i] = cov[i]
        P[i, i + 4] = cov[4 + i]
        P[i + 4, i] = cov[8 + i]
        P[i + 4, i + 4] = cov[12 + i]
    return P


def dense_to_cov(P):
    P = np.asarray(P, F32)
    cov = np.zeros(16, F32)
    for i in range(4):
        cov[i] = P[i, i]
        cov[4 + i] = P[i, i + 4]
        cov[8 + i] = P[i + 4, i]
        cov[12 + i] = P[i + 4, i + 4]
    return cov
