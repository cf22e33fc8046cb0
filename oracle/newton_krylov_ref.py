"""TEST INFRASTRUCTURE — CPU restatement of the Newton-Krylov caller of the momentum residual.

The reference solves its implicit momentum step with PETSc 3.1 (Source/makefile:26-28 names the
version; PETSc is an external dependency, absent from /root/reference): Implicit_MatrixFree,
Source/implicitsolver.c:4203-4299, sets up SNESTR (:4251) with Eisenstat-Walker version 3 (:4251-4252),
a matrix-free Jacobian (MatCreateSNESMF, :4242), GMRES without preconditioner (:4260,4273) and the
tolerances of :4254 / :4277-4279.  The published PETSc 3.1 algorithms are restated here in numpy:

  SNESSolve_TR           src/snes/impls/tr/tr.c      trust region, More' step-length test on the Krylov iterates
  MatMFFD "wp"           src/mat/impls/mffd/wp.c     h = error_rel * sqrt(1 + |u|) / |a|, error_rel = sqrt(eps)
  KSPSolve_GMRES         src/ksp/ksp/impls/gmres     classical Gram-Schmidt, Givens rotations, restart
  SNESKSPEW_PreSolve     src/snes/interface/snes.c   version 3 forcing term
  SNES/KSPDefaultConverged

`residual(x) -> F` is any callable; tests pass the ORACLE's FormFunction_SNES (oracle/refdrv.py), so the
device solver (vfs_momentum_solve, driving the CUDA residual) is compared with the same algorithm driving
the reference's own residual.  Parity against PETSc's iterates themselves is UNPINNED (PETSc cannot be
built here); only tests/ import this module.
"""
import math
import numpy as np

ERROR_REL = 1.490116119384766e-08


def gmres(matvec, b, restart, rtol, atol, dtol, maxits, delta=0.0):
    """Solve J y = b from y = 0.  Returns (y, its, reason, rnorm)."""
    n = b.size
    y_sol = np.zeros(n)
    have_base = False
    its, reason, res = 0, 0, 0.0
    m = restart
    rnorm0 = ttol = 0.0
    cycle = 0
    while True:
        if cycle == 0 or not have_base or np.linalg.norm(y_sol) == 0:
            r = b.copy()
        else:
            r = b - matvec(y_sol)
        beta = math.sqrt(float(r @ r))
        if cycle == 0:
            rnorm0 = beta
            ttol = max(rtol * rnorm0, atol)
            if beta <= ttol:
                reason = 3 if beta <= atol else 2
                break
        V = [r / beta]
        H = np.zeros((m + 1, m))
        cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
        g[0] = beta
        k = 0
        res = beta
        kk = 0
        while k < m:
            w = matvec(V[k], unit=True)
            h = np.array([float(V[q] @ w) for q in range(k + 1)])
            for q in range(k + 1):
                w = w - h[q] * V[q] if q == 0 else w - h[q] * V[q]
            hn = math.sqrt(float(w @ w))
            col = np.concatenate([h, [hn]])
            for q in range(k):
                t = cs[q] * col[q] + sn[q] * col[q + 1]
                col[q + 1] = -sn[q] * col[q] + cs[q] * col[q + 1]
                col[q] = t
            den = math.sqrt(col[k] ** 2 + col[k + 1] ** 2)
            if den == 0:
                reason = -5
                kk = k
                break
            cs[k], sn[k] = col[k] / den, col[k + 1] / den
            col[k] = den
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            H[:k + 1, k] = col[:k + 1]
            res = abs(g[k + 1])
            its += 1
            if hn != 0:
                V.append(w / hn)
            kk = k + 1
            if res <= ttol:
                reason = 3 if res <= atol else 2
                break
            if res >= dtol * rnorm0:
                reason = -4
                break
            if its >= maxits:
                reason = -3
                break
            if delta > 0 and not have_base:
                yk = np.linalg.solve(np.triu(H[:k + 1, :k + 1]), g[:k + 1])
                if math.sqrt(float(yk @ yk)) >= delta:
                    reason = 6
                    break
            if hn == 0:
                reason = 5
                break
            k += 1
        if kk > 0:
            yk = np.zeros(kk)
            for q in range(kk - 1, -1, -1):
                t = g[q]
                for r_ in range(q + 1, kk):
                    t -= H[q, r_] * yk[r_]
                yk[q] = t / H[q, q]
            for q in range(kk):
                y_sol = y_sol + yk[q] * V[q]
            have_base = True
        if reason:
            break
        if delta > 0 and math.sqrt(float(y_sol @ y_sol)) >= delta:
            reason = 6
            break
        cycle += 1
    return y_sol, its, reason, res


def snes_tr(residual, u0, max_newton=50, restart=30, max_krylov=1000, snes_atol=1e-50, snes_rtol=1e-8, snes_stol=1e-8,
            ksp_rtol=1e-5, ksp_atol=1e-50, ksp_dtol=1e5, use_ew=True, trust_region=True):
    """Returns (u, info) with info = dict(fnorm_history, ksp_its_history, reason, residual_evals)."""
    shape = u0.shape
    evals = [0]

    def F(x):
        evals[0] += 1
        return np.asarray(residual(x.reshape(shape)), dtype=float).ravel().copy()
    U = np.asarray(u0, dtype=float).ravel().copy()
    mu, eta, delta0, delta1, delta2, delta3, sigma, deltatol = 0.25, 0.75, 0.2, 0.3, 0.75, 2.0, 1e-4, 1e-12
    Fu = F(U)
    fnorm = math.sqrt(float(Fu @ Fu))
    xnorm = math.sqrt(float(U @ U))
    delta = delta0 * xnorm
    hist, kits = [fnorm], []
    ttol = fnorm * snes_rtol
    reason = 0
    if fnorm != fnorm:
        reason = -4
    elif fnorm < snes_atol:
        reason = 2
    rtol0, rtolmax, gamma, alpha = 0.3, 0.9, 1.0, 0.5 * (1.0 + math.sqrt(5.0))
    rtol_last = norm_last = 0.0
    ynorm = 0.0
    it = 0
    while it < max_newton and not reason:
        rtol = ksp_rtol
        if use_ew:
            if it == 0:
                rtol = rtol0
            else:
                rtol = gamma * (fnorm / norm_last) ** alpha
                stol = gamma * rtol_last ** alpha
                stol = max(rtol, stol)
                rtol = min(rtol0, stol)
                stol = gamma * ttol / fnorm
                stol = max(rtol, stol)
                rtol = min(rtol0, stol)
            rtol = min(rtol, rtolmax)
            rtol_last, norm_last = rtol, fnorm
        ufact = math.sqrt(1.0 + xnorm)

        def matvec(a, unit=False, U=U, Fu=Fu, ufact=ufact):
            na = 1.0 if unit else math.sqrt(float(a @ a))        # GMRES basis vectors are unit vectors (to rounding)
            h = ERROR_REL * ufact / na
            return (1.0 / h) * F(U + h * a) + (-1.0 / h) * Fu
        Yt, lits, kreason, _ = gmres(matvec, Fu, restart, rtol, ksp_atol, ksp_dtol, max_krylov, delta if trust_region else 0.0)
        kits.append(lits)
        nrm1 = math.sqrt(float(Yt @ Yt))
        breakout = False
        if not trust_region:
            Y = U - Yt
            G = F(Y)
            gnorm = math.sqrt(float(G @ G))
            ynorm = nrm1
        else:
            while True:
                nrm, scale = nrm1, 1.0
                if nrm >= delta:
                    nrm = delta / nrm
                    gpnorm = (1.0 - nrm) * fnorm
                    scale = nrm
                    ynorm = delta
                else:
                    gpnorm = 0.0
                    ynorm = nrm
                Y = U + (-scale) * Yt
                G = F(Y)
                gnorm = math.sqrt(float(G @ G))
                rho = 0.0 if fnorm == gpnorm else (fnorm * fnorm - gnorm * gnorm) / (fnorm * fnorm - gpnorm * gpnorm)
                if rho < mu:
                    delta *= delta1
                elif rho < eta:
                    delta *= delta2
                else:
                    delta *= delta3
                if rho > sigma:
                    break
                if delta < xnorm * deltatol:
                    reason, breakout = -8, True
                    break
        if breakout:
            break
        fnorm, Fu, U = gnorm, G, Y
        hist.append(fnorm)
        xnorm = math.sqrt(float(U @ U))
        if trust_region and delta < xnorm * deltatol:
            reason = 7
        elif fnorm != fnorm:
            reason = -4
        elif fnorm < snes_atol:
            reason = 2
        elif fnorm <= ttol:
            reason = 3
        elif ynorm < snes_stol * xnorm:
            reason = 4
        it += 1
    if not reason:
        reason = -5
    return U.reshape(shape), dict(fnorm_history=hist, ksp_its_history=kits, reason=reason, residual_evals=evals[0], delta=delta)
