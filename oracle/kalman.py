"""oracle/kalman.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Exact float64 Kalman filter / RTS smoother for linear-Gaussian state-space models, replacing the
pykalman dependency of the reference's tests (test/models/lgssm.py:75-88, test/test_inference.py:
146-375; pykalman is not installed in this image).  Convention follows the reference's infer():
x_0 ~ N(m0, P0) and y_0 is emitted from x_0 (no transition before the first observation,
inference.py:85-98); x_t = A x_{t-1} + a + N(0, Q); y_t = C x_t + c + N(0, R).

All functions are batched over the leading axis of ``obs`` ([T, B] for the 1-D model,
[T, B, Dy] for the matrix model).
"""
import numpy as np


def lgssm1d_log_evidence(obs, m0, P0, A, Q, C, R, a=0.0, c=0.0):
    """log p(y_{0:T-1}) per batch row for the scalar LGSSM.  obs: [T, B] -> [B] float64."""
    y = np.asarray(obs, dtype=np.float64)
    T, B = y.shape
    m = np.full(B, float(m0))
    P = np.full(B, float(P0))
    ll = np.zeros(B)
    for t in range(T):
        if t > 0:
            m = A * m + a
            P = A * A * P + Q
        S = C * C * P + R
        r = y[t] - (C * m + c)
        ll += -0.5 * (np.log(2.0 * np.pi * S) + r * r / S)
        G = P * C / S
        m = m + G * r
        P = (1.0 - G * C) * P
    return ll


def lgssm1d_smooth(obs, m0, P0, A, Q, C, R, a=0.0, c=0.0):
    """RTS smoother: returns (means [T, B], variances [T, B])."""
    y = np.asarray(obs, dtype=np.float64)
    T, B = y.shape
    mf = np.zeros((T, B)); Pf = np.zeros((T, B)); mp = np.zeros((T, B)); Pp = np.zeros((T, B))
    m = np.full(B, float(m0)); P = np.full(B, float(P0))
    for t in range(T):
        if t > 0:
            m = A * m + a
            P = A * A * P + Q
        mp[t], Pp[t] = m, P
        S = C * C * P + R
        G = P * C / S
        m = m + G * (y[t] - (C * m + c))
        P = (1.0 - G * C) * P
        mf[t], Pf[t] = m, P
    ms, Ps = mf.copy(), Pf.copy()
    for t in range(T - 2, -1, -1):
        J = Pf[t] * A / Pp[t + 1]
        ms[t] = mf[t] + J * (ms[t + 1] - mp[t + 1])
        Ps[t] = Pf[t] + J * J * (Ps[t + 1] - Pp[t + 1])
    return ms, Ps


def lgssm_log_evidence(obs, m0, P0, A, Q, C, R):
    """Matrix LGSSM.  obs: [T, B, Dy]; m0 [Dx]; P0, A, Q [Dx, Dx]; C [Dy, Dx]; R [Dy, Dy] -> [B]."""
    y = np.asarray(obs, dtype=np.float64)
    T, B, Dy = y.shape
    A, Q, C, R = (np.asarray(v, dtype=np.float64) for v in (A, Q, C, R))
    m = np.tile(np.asarray(m0, dtype=np.float64), (B, 1))
    P = np.asarray(P0, dtype=np.float64).copy()
    ll = np.zeros(B)
    for t in range(T):
        if t > 0:
            m = m @ A.T
            P = A @ P @ A.T + Q
        S = C @ P @ C.T + R
        Sinv = np.linalg.inv(S)
        r = y[t] - m @ C.T
        _, logdet = np.linalg.slogdet(S)
        ll += -0.5 * (Dy * np.log(2.0 * np.pi) + logdet + np.einsum("bi,ij,bj->b", r, Sinv, r))
        G = P @ C.T @ Sinv
        m = m + r @ G.T
        P = (np.eye(P.shape[0]) - G @ C) @ P
    return ll
