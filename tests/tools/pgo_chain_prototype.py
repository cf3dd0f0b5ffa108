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


# ---- the same preconditioner by PARALLEL CYCLIC REDUCTION: the fully parallel form of the two sweeps (log2 n levels) ------------------------------
# Row i of M z = r is  L_i z_{i-1} + D_i z_i + U_i z_{i+1} = r_i  with L_i = C_{i-1}, U_i = C_i^T.  One PCR level with stride s eliminates the
# neighbours at distance s from every row at once:
#     alpha_i = -L_i D_{i-s}^-1,  gamma_i = -U_i D_{i+s}^-1
#     D_i <- D_i + alpha_i U_{i-s} + gamma_i L_{i+s};   L_i <- alpha_i L_{i-s};   U_i <- gamma_i U_{i+s};   r_i <- r_i + alpha_i r_{i-s} + gamma_i r_{i+s}
# After ceil(log2 n) levels every row is decoupled: z_i = D_i^-1 r_i.  The matrix part does not depend on r: `pcr_factorise` keeps alpha, gamma of
# every level and the final D^-1 (once per solve), `pcr_apply` replays the right-hand-side part (once per CG iteration; every level is one fully
# parallel pass over the nodes).
def pcr_factorise(Ad, C):
    n = len(Ad)
    D = np.array(Ad, np.float64)
    L = np.zeros((n, 6, 6)); U = np.zeros((n, 6, 6))
    if n > 1:
        L[1:] = C
        U[:-1] = np.transpose(C, (0, 2, 1))
    levels = []
    s = 1
    while s < n:
        Dinv = np.linalg.inv(D)
        alpha, gamma = np.zeros((n, 6, 6)), np.zeros((n, 6, 6))
        alpha[s:] = -np.einsum("nij,njk->nik", L[s:], Dinv[:-s])
        gamma[:-s] = -np.einsum("nij,njk->nik", U[:-s], Dinv[s:])
        Dn = D.copy()
        Dn[s:] += np.einsum("nij,njk->nik", alpha[s:], U[:-s])
        Dn[:-s] += np.einsum("nij,njk->nik", gamma[:-s], L[s:])
        Ln, Un = np.zeros_like(L), np.zeros_like(U)
        Ln[s:] = np.einsum("nij,njk->nik", alpha[s:], L[:-s])
        Un[:-s] = np.einsum("nij,njk->nik", gamma[:-s], U[s:])
        levels.append((s, alpha, gamma))
        D, L, U = Dn, Ln, Un
        s *= 2
    return levels, np.linalg.inv(D)


def pcr_apply(levels, Dinv, r):
    r = np.array(r, np.float64).reshape(len(Dinv), 6).copy()
    for s, alpha, gamma in levels:
        rn = r.copy()
        rn[s:] += np.einsum("nij,nj->ni", alpha[s:], r[:-s])
        rn[:-s] += np.einsum("nij,nj->ni", gamma[:-s], r[s:])
        r = rn
    return np.einsum("nij,nj->ni", Dinv, r)
