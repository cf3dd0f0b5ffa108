"""Prototype (numpy, test infrastructure) of the pose-graph step solve planned for the next kernel revision: conjugate gradients preconditioned
with the BLOCK-TRIDIAGONAL part of the damped normal equations (diagonal blocks + the odometry chain), factorised once per solve as
M = L D L^T with unit block-bidiagonal L.  Written operation by operation the way the CUDA kernel will do it (block Thomas
factorisation, forward sweep, block-diagonal solve, backward sweep, fixed node held at zero) so that it can serve as that kernel's checker.

    chain_blocks(ids, H_diag, H_off, radius, fixed)  -> damped diagonal blocks Ad [n,6,6], lower chain blocks C [n-1,6,6] (= A[i+1][i])
    factorise(Ad, C)                                 -> Sinv [n,6,6], W [n-1,6,6]
    apply(Sinv, W, r)                                -> M^-1 r
    solve(ids, H_diag, H_off, g, ...)                -> delta, iterations, relative residual
"""
import numpy as np


def chain_blocks(ids, Hd, Ho, radius, fixed):
    n = len(Hd)
    idx = np.arange(6)
    Ad = np.array(Hd, np.float64).reshape(n, 6, 6).copy()
    Ad[:, idx, idx] += np.clip(Ad[:, idx, idx], 1e-6, 1e32) / radius
    C = np.zeros((max(n - 1, 0), 6, 6))
    for c, (a, b, _t) in enumerate(np.asarray(ids).reshape(-1, 3)):
        if abs(int(a) - int(b)) == 1:                      # A[begin][end] = Ho_c, A[end][begin] = Ho_c^T
            lo = min(a, b)
            C[lo] += Ho[c].reshape(6, 6) if a > b else Ho[c].reshape(6, 6).T
    Ad[fixed] = np.eye(6)                                  # the fixed node is not a variable: identity block, no coupling
    if fixed > 0:
        C[fixed - 1] = 0.0
    if fixed < n - 1:
        C[fixed] = 0.0
    return Ad, C


def factorise(Ad, C):
    """Block Thomas: S_0 = Ad_0; W_i = C_{i-1} S_{i-1}^-1; S_i = Ad_i - W_i C_{i-1}^T."""
    n = len(Ad)
    Sinv, W = np.zeros((n, 6, 6)), np.zeros((max(n - 1, 0), 6, 6))
    S = Ad[0]
    Sinv[0] = np.linalg.inv(S)
    for i in range(1, n):
        W[i - 1] = C[i - 1] @ Sinv[i - 1]
        S = Ad[i] - W[i - 1] @ C[i - 1].T
        Sinv[i] = np.linalg.inv(S)
    return Sinv, W


def apply(Sinv, W, r):
    n = len(Sinv)
    y = np.array(r, np.float64).reshape(n, 6).copy()
    for i in range(1, n):                                  # L y = r
        y[i] -= W[i - 1] @ y[i - 1]
    z = np.einsum("nij,nj->ni", Sinv, y)                   # D w = y
    for i in range(n - 2, -1, -1):                         # L^T z = w
        z[i] -= W[i].T @ z[i + 1]
    return z


def matvec(ids, Hd, Ho, radius, fixed, x):
    n = len(Hd)
    idx = np.arange(6)
    Hd = np.asarray(Hd, np.float64).reshape(n, 6, 6)
    y = np.einsum("nij,nj->ni", Hd, x) + np.clip(Hd[:, idx, idx], 1e-6, 1e32) / radius * x
    ids = np.asarray(ids).reshape(-1, 3)
    if len(ids):
        B = np.asarray(Ho, np.float64).reshape(-1, 6, 6)[:len(ids)]
        np.add.at(y, ids[:, 0], np.einsum("cij,cj->ci", B, x[ids[:, 1]]))
        np.add.at(y, ids[:, 1], np.einsum("cji,cj->ci", B, x[ids[:, 0]]))
    y[fixed] = 0.0
    return y


def solve(ids, Hd, Ho, g, fixed=0, radius=1e4, max_iters=2000, rel_tol=1e-12):
    n = len(Hd)
    Sinv, W = factorise(*chain_blocks(ids, Hd, Ho, radius, fixed))
    x = np.zeros((n, 6))
    r = -np.array(g, np.float64).reshape(n, 6)
    r[fixed] = 0.0
    z = apply(Sinv, W, r)
    z[fixed] = 0.0
    p = z.copy()
    rz, bnorm = float(np.sum(r * z)), float(np.linalg.norm(r))
    rel, it = (1.0 if bnorm > 0 else 0.0), 0
    while it < max_iters and rel > rel_tol:
        q = matvec(ids, Hd, Ho, radius, fixed, p)
        pq = float(np.sum(p * q))
        if not pq > 0.0:
            break
        alpha = rz / pq
        x += alpha * p
        r -= alpha * q
        z = apply(Sinv, W, r)
        z[fixed] = 0.0
        rz_new = float(np.sum(r * z))
        rel = float(np.linalg.norm(r)) / bnorm
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it, rel
