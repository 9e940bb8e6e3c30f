"""Developer prototype (CPU, numpy, vectorised over fits) of the `fast` solver that the CUDA kernel
fsq_lmfast.cu implements: MINPACK's bounded trust-region LM (mpfit semantics for pegging, alpha
scaling, snapping, ratio / delta / par updates and termination) driven by the ANALYTIC Jacobian
through column-scaled normal equations + Cholesky instead of a finite-difference Jacobian + QR.

Not product code and not the oracle: it exists to choose the arithmetic (which quantities need
float64) before writing the kernel, by measuring agreement with the reference goldens.

    python tools/fast_proto.py [--dtype f32|f64] [--acc f32|f64]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEG = np.pi / 180.


def model_jac(x, rr, cc, dt):
    """x [N,7] -> E [N,P], J [N,P,7] = d(model)/dp, in dtype dt. rr/cc pixel row/col [P]."""
    x = x.astype(dt)
    H, A, cy, cx, wx, wy, th = [x[:, i:i + 1] for i in range(7)]
    rota = (th * dt(DEG))
    cs, sn = np.cos(rota), np.sin(rota)
    dx = cx - rr[None, :].astype(dt)
    dy = cy - cc[None, :].astype(dt)
    u = dx * cs - dy * sn
    v = dx * sn + dy * cs
    iwx, iwy = dt(1) / wx, dt(1) / wy
    a, b = u * iwx, v * iwy
    E = np.exp(dt(-0.5) * (a * a + b * b))
    AE = A * E
    J = np.empty(E.shape + (7,), dtype=dt)
    J[..., 0] = 1
    J[..., 1] = E
    J[..., 2] = AE * (a * sn * iwx - b * cs * iwy)
    J[..., 3] = -AE * (a * cs * iwx + b * sn * iwy)
    J[..., 4] = AE * a * a * iwx
    J[..., 5] = AE * b * b * iwy
    J[..., 6] = AE * a * b * (wy * iwx - wx * iwy) * dt(DEG)
    return E, J


def model_only(x, rr, cc, dt):
    x = x.astype(dt)
    H, A, cy, cx, wx, wy, th = [x[:, i:i + 1] for i in range(7)]
    rota = (th * dt(DEG))
    cs, sn = np.cos(rota), np.sin(rota)
    dx = cx - rr[None, :].astype(dt)
    dy = cy - cc[None, :].astype(dt)
    a = (dx * cs - dy * sn) / wx
    b = (dx * sn + dy * cs) / wy
    return H + A * np.exp(dt(-0.5) * (a * a + b * b))


def chol_solve(C, rhs, rank_eps):
    """Batched Cholesky of symmetric [N,7,7] with singular-pivot skipping; solves C y = rhs.
    Returns y, L, ok-mask per column (False = skipped)."""
    N, n = rhs.shape
    L = np.zeros_like(C)
    okc = np.ones((N, n), dtype=bool)
    for j in range(n):
        d = C[:, j, j] - np.einsum("nk,nk->n", L[:, j, :j], L[:, j, :j])
        good = d > rank_eps
        okc[:, j] = good
        dj = np.sqrt(np.where(good, d, 1))
        L[:, j, j] = np.where(good, dj, 0)
        for i in range(j + 1, n):
            s = C[:, i, j] - np.einsum("nk,nk->n", L[:, i, :j], L[:, j, :j])
            L[:, i, j] = np.where(good, s / dj, 0)
    y = tri_solve(L, okc, rhs)
    return y, L, okc


def fwd(L, okc, rhs):
    N, n = rhs.shape
    z = np.zeros_like(rhs)
    for j in range(n):
        s = rhs[:, j] - np.einsum("nk,nk->n", L[:, j, :j], z[:, :j])
        z[:, j] = np.where(okc[:, j], s / np.where(okc[:, j], L[:, j, j], 1), 0)
    return z


def tri_solve(L, okc, rhs):
    N, n = rhs.shape
    z = fwd(L, okc, rhs)
    y = np.zeros_like(rhs)
    for j in range(n - 1, -1, -1):
        s = z[:, j] - np.einsum("nk,nk->n", L[:, j + 1:, j], y[:, j + 1:])
        y[:, j] = np.where(okc[:, j], s / np.where(okc[:, j], L[:, j, j], 1), 0)
    return y


LMPAR_EXTRA, LMPAR_TICKS = [], []       # instrumentation: extra lmpar iterations per damped fit-tick; fits per tick


def fast_fit(data, p0, lo, hi, lim_lo, lim_hi, dt=np.float64, acc=np.float64, fdt=None, maxiter=200, tr='lm',
             ftol=1e-10, xtol=1e-10, gtol=1e-10, factor=100.0, rank_eps=None, verbose=False):
    """data [N,win,win]; returns dict(params, status, niter, nfev, chi2)."""
    N, win, _ = data.shape
    P = win * win
    rr, cc = [a.ravel() for a in np.indices((win, win))]
    fdt = fdt or dt                                        # dtype of model / residual / chi^2
    d = data.reshape(N, P).astype(fdt)
    x = p0.astype(np.float64).copy()
    lo, hi = lo.astype(np.float64), hi.astype(np.float64)
    ql, qu = lim_lo.astype(bool), lim_hi.astype(bool)
    eps = float(np.finfo(dt).eps)
    MACHEP = float(np.finfo(np.float64).eps)
    if rank_eps is None:
        rank_eps = 16 * eps
    status = np.zeros(N, dtype=np.int32)
    niter = np.ones(N, dtype=np.int32)
    nfev = np.ones(N, dtype=np.int32)
    par = np.zeros(N)
    delta = np.zeros(N)
    xnorm = np.zeros(N)
    diag = np.ones((N, 7))
    need = np.ones(N, dtype=bool)
    m0 = model_only(x, rr, cc, fdt)
    f = d - m0
    facc = np.float64 if fdt == np.float64 else acc
    fnorm = np.sqrt((f.astype(facc) ** 2).sum(axis=1)).astype(np.float64)
    fnorm1 = np.full(N, -1.0)
    A = np.zeros((N, 7, 7))
    g = np.zeros((N, 7))
    acn = np.zeros((N, 7))
    lpeg = np.zeros((N, 7), dtype=bool)
    upeg = np.zeros((N, 7), dtype=bool)
    gnorm = np.zeros(N)
    for it in range(100000):
        act = status == 0
        if not act.any():
            break
        idx = np.nonzero(act & need)[0]
        if len(idx):
            E, J = model_jac(x[idx], rr, cc, dt)
            J = -J                                         # d(residual)/dp
            fi = f[idx].astype(dt)
            gi = np.einsum("npk,np->nk", J.astype(acc), fi.astype(acc)).astype(np.float64)
            lp = ql[idx] & (x[idx] == lo[idx])
            up = qu[idx] & (x[idx] == hi[idx])
            zero = (lp & (gi > 0)) | (up & (gi < 0))       # mpfit.py:1073-1091
            J = np.where(zero[:, None, :], 0, J)
            gi = np.where(zero, 0.0, gi)
            Ai = np.einsum("npk,npl->nkl", J.astype(acc), J.astype(acc)).astype(np.float64)
            A[idx], g[idx], lpeg[idx], upeg[idx] = Ai, gi, lp, up
            an = np.sqrt(np.einsum("nkk->nk", Ai))
            acn[idx] = an
            first = niter[idx] == 1
            dg = np.where(an == 0, 1.0, an)
            diag[idx] = np.where(first[:, None], dg, diag[idx])
            xn = np.sqrt(((diag[idx] * x[idx]) ** 2).sum(axis=1))
            xnorm[idx] = np.where(first, xn, xnorm[idx])
            dl = factor * xn
            dl = np.where(dl == 0, factor, dl)
            delta[idx] = np.where(first, dl, delta[idx])
            # scaled gradient norm, mpfit.py:1142-1148: max_j |J_j . f| / (|f| |J_j|)
            with np.errstate(divide="ignore", invalid="ignore"):
                gn = np.where(an > 0, np.abs(gi) / (fnorm[idx][:, None] * np.where(an > 0, an, 1)), 0).max(axis=1)
            gn = np.where(fnorm[idx] == 0, 0.0, gn)
            gnorm[idx] = gn
            status[idx] = np.where(gn <= gtol, 4, status[idx])
            diag[idx] = np.maximum(diag[idx], an)
            need[idx] = False
        idx = np.nonzero(status == 0)[0]
        if not len(idx):
            continue
        # ---------------- lmpar on normal equations (scaled by current column norms)
        Ai, gi, Di, dl = A[idx], g[idx], diag[idx], delta[idx]
        S = np.where(acn[idx] > 0, acn[idx], 1.0)
        C = (Ai / (S[:, :, None] * S[:, None, :])).astype(dt).astype(np.float64)
        gs = (gi / S)
        T = (Di / S) ** 2                                   # damping weights in scaled variables
        y, L, okc = chol_solve(C, -gs, rank_eps)
        p = y / S
        dxn = np.sqrt(((Di * p) ** 2).sum(axis=1))
        fp = dxn - dl
        pr = np.zeros(len(idx))
        todo = fp > 0.1 * dl
        dog = np.zeros(len(idx), dtype=bool)
        if tr == 'dogleg' and todo.any():
            # Powell dogleg in the scaled variables y = S p, trust region |W y| <= delta, W^2 = T
            Wd2 = 1.0 / T                                        # W^-2
            sdir = -Wd2 * gs                                        # steepest descent in the W metric
            Cd = np.einsum("nkl,nl->nk", C, sdir)
            dCd = (sdir * Cd).sum(axis=1)
            gWg = (gs * Wd2 * gs).sum(axis=1)
            tau = gWg / np.where(dCd > 0, dCd, 1)
            yc = tau[:, None] * sdir
            nyc = np.sqrt((T * yc * yc).sum(axis=1))
            nd = np.sqrt((T * sdir * sdir).sum(axis=1))
            ygn = y
            v = ygn - yc
            a_ = (T * v * v).sum(axis=1)
            b_ = 2 * (T * yc * v).sum(axis=1)
            c_ = nyc ** 2 - dl ** 2
            disc = np.maximum(b_ * b_ - 4 * a_ * c_, 0)
            beta = (-b_ + np.sqrt(disc)) / np.where(a_ > 0, 2 * a_, 1)
            beta = np.clip(beta, 0, 1)
            ydl = np.where((nyc >= dl)[:, None], (dl / np.where(nd > 0, nd, 1))[:, None] * sdir, yc + beta[:, None] * v)
            ydl = np.where((dCd > 0)[:, None], ydl, (dl / np.where(nd > 0, nd, 1))[:, None] * sdir)
            y = np.where(todo[:, None], ydl, y)
            p = y / S
            dog = todo.copy()
            pr = np.where(todo, 1.0, 0.0)
            todo = np.zeros_like(todo)
        if todo.any():
            full = okc.all(axis=1)
            # parl (only when full rank): phi(0)/ -phi'(0)
            u = (Di ** 2 * p / np.where(dxn > 0, dxn, 1)[:, None]) / S
            w = fwd(L, okc, u)
            t2 = (w ** 2).sum(axis=1)
            parl = np.where(full & (t2 > 0), (fp / dl) / np.where(t2 > 0, t2, 1), 0.0)
            gsn = np.sqrt(((gi / Di) ** 2).sum(axis=1))
            paru = gsn / dl
            paru = np.where(paru == 0, 2.2e-308 / np.minimum(dl, 0.1), paru)
            prr = np.minimum(np.maximum(par[idx], parl), paru)
            prr = np.where(prr == 0, gsn / np.where(dxn > 0, dxn, 1), prr)
            run = todo.copy()
            fp_run = fp.copy()
            cnt = np.zeros(len(idx), dtype=np.int64)            # extra lmpar iterations of this tick, per damped fit
            for k in range(10):
                cnt += run
                prr = np.where(run & (prr == 0), np.maximum(2.2e-308, paru * 0.001), prr)
                M = C + prr[:, None, None] * (T[:, :, None] * np.eye(7)[None])
                y2, L2, ok2 = chol_solve(M.astype(dt).astype(np.float64), -gs, rank_eps * 0)
                p2 = y2 / S
                dx2 = np.sqrt(((Di * p2) ** 2).sum(axis=1))
                temp = fp_run.copy()
                fp2 = dx2 - dl
                p = np.where(run[:, None], p2, p)
                dxn = np.where(run, dx2, dxn)
                fp_run = np.where(run, fp2, fp_run)
                stop = (np.abs(fp2) <= 0.1 * dl) | ((parl == 0) & (fp2 <= temp) & (temp < 0)) | (k == 9)
                pr = np.where(run, prr, pr)
                run = run & ~stop
                if not run.any():
                    break
                u = (Di ** 2 * p2 / np.where(dx2 > 0, dx2, 1)[:, None]) / S
                w = fwd(L2, ok2, u)
                t2 = (w ** 2).sum(axis=1)
                parc = (fp2 / dl) / np.where(t2 > 0, t2, 1)
                parl = np.where(run & (fp2 > 0), np.maximum(parl, prr), parl)
                paru = np.where(run & (fp2 < 0), np.minimum(paru, prr), paru)
                prr = np.where(run, np.maximum(parl, prr + parc), prr)
            LMPAR_EXTRA.append(cnt[todo])
        LMPAR_TICKS.append(len(idx))
        par[idx] = pr
        # ---------------- bounds, mpfit.py:1184-1231
        xi = x[idx]
        lp, up = lpeg[idx], upeg[idx]
        mx = p.max(axis=1, keepdims=True)
        mn = p.min(axis=1, keepdims=True)
        p = np.where(lp, np.clip(p, 0, np.maximum(mx, 0)), p)
        p = np.where(up, np.clip(p, np.minimum(mn, 0), 0), p)
        big = np.abs(p) > MACHEP
        with np.errstate(divide="ignore", invalid="ignore"):
            tl = np.where(big & ql[idx] & (xi + p < lo[idx]), (lo[idx] - xi) / p, 1.0)
            tu = np.where(big & qu[idx] & (xi + p > hi[idx]), (hi[idx] - xi) / p, 1.0)
        alpha = np.minimum(1.0, np.minimum(tl.min(axis=1), tu.min(axis=1)))
        p = p * alpha[:, None]
        xnew = xi + p
        sgnu = np.where(hi[idx] >= 0, 1.0, -1.0)
        sgnl = np.where(lo[idx] >= 0, 1.0, -1.0)
        ulim1 = hi[idx] * (1 - sgnu * MACHEP) - (hi[idx] == 0) * MACHEP
        llim1 = lo[idx] * (1 + sgnl * MACHEP) + (lo[idx] == 0) * MACHEP
        xnew = np.where(qu[idx] & (xnew >= ulim1), hi[idx], xnew)
        xnew = np.where(ql[idx] & (xnew <= llim1), lo[idx], xnew)
        pnorm = np.sqrt(((Di * p) ** 2).sum(axis=1))
        dl = np.where(niter[idx] == 1, np.minimum(dl, pnorm), dl)
        f1 = d[idx] - model_only(xnew, rr, cc, fdt)
        fn1 = np.sqrt((f1.astype(facc) ** 2).sum(axis=1)).astype(np.float64)
        nfev[idx] += 1
        fn = fnorm[idx]
        actred = np.where(0.1 * fn1 < fn, 1.0 - (fn1 / fn) ** 2, -1.0)
        pAp = np.einsum("nk,nkl,nl->n", p, Ai, p)          # |J p|^2 (p already scaled by alpha)
        t1sq = alpha ** 2 * np.maximum(pAp, 0) / fn ** 2   # mpfit applies alpha once more here (:1265)
        t2sq = alpha * pr * pnorm ** 2 / fn ** 2
        prered = t1sq + 2 * t2sq
        dirder = -(t1sq + t2sq)
        if tr == 'dogleg':
            gp = (gi * p).sum(axis=1)
            prered = np.where(dog, -(2 * gp + pAp) / fn ** 2, prered)
            dirder = np.where(dog, gp / fn ** 2, dirder)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(prered != 0, actred / prered, 0.0)
        low = ratio <= 0.25
        with np.errstate(divide="ignore", invalid="ignore"):
            tmp = np.where(actred >= 0, 0.5, 0.5 * dirder / (dirder + 0.5 * actred))
        tmp = np.where((0.1 * fn1 >= fn) | (tmp < 0.1), 0.1, tmp)
        dl_low = tmp * np.minimum(dl, pnorm / 0.1)
        grow = (~low) & ((pr == 0) | (ratio >= 0.75))
        dl = np.where(low, dl_low, np.where(grow, pnorm / 0.5, dl))
        pr = np.where(low, pr / tmp, np.where(grow, 0.5 * pr, pr))
        acc_ = ratio >= 1e-4
        x[idx] = np.where(acc_[:, None], xnew, xi)
        f[idx] = np.where(acc_[:, None], f1, f[idx])
        xnorm[idx] = np.where(acc_, np.sqrt(((Di * xnew) ** 2).sum(axis=1)), xnorm[idx])
        fnorm[idx] = np.where(acc_, fn1, fn)
        fnorm1[idx] = fn1
        niter[idx] += acc_
        delta[idx], par[idx] = dl, pr
        c1 = (np.abs(actred) <= ftol) & (prered <= ftol) & (0.5 * ratio <= 1)
        st = np.where(c1, 1, 0)
        st = np.where(dl <= xtol * xnorm[idx], 2, st)
        st = np.where(c1 & (st == 2), 3, st)
        st = np.where((st == 0) & (niter[idx] >= maxiter), 5, st)
        st = np.where((st == 0) & (dl <= MACHEP * xnorm[idx]), 7, st)
        bad = ~np.isfinite(ratio) | ~np.isfinite(xnew).all(axis=1)
        st = np.where((st == 0) & ~acc_ & bad, -16, st)
        status[idx] = st
        need[idx] = acc_
    chi2 = np.maximum(fnorm, fnorm1) ** 2
    return dict(params=x, status=status, niter=niter, nfev=nfev + (status > 0), chi2=chi2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--acc", default="f64")
    ap.add_argument("--fdt", default="")
    ap.add_argument("--ftol", type=float, default=1e-10)
    ap.add_argument("--xtol", type=float, default=1e-10)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--tr", default="lm")
    a = ap.parse_args()
    from fluorosequencingimageanalysis_b200 import synth, pflib
    from oracle.stability import agree
    G = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(G, "fits5_seed0.npz"))
    st = np.load(os.path.join(G, "stable5_seed0.npz"))
    img = synth.synth_frame(0)
    cands = g["cands"]
    n = a.n or len(cands)
    subs = np.stack([img[h - 2:h + 3, w - 2:w + 3].astype(np.int64) for h, w in cands[:n]])
    p0, lo, hi, ll, lh = pflib._pflib_limits(subs)
    dt = {"f32": np.float32, "f64": np.float64}[a.dtype]
    acc = {"f32": np.float32, "f64": np.float64}[a.acc]
    t = time.time()
    fdt = {"f32": np.float32, "f64": np.float64, "": None}[a.fdt]
    r = fast_fit(subs, p0, lo, hi, ll, lh, dt=dt, acc=acc, fdt=fdt, ftol=a.ftol, xtol=a.xtol, tr=a.tr)
    print("time %.1fs" % (time.time() - t))
    P = r["params"]
    for key in ("clean", "ref"):
        ok = agree(P, g[key + "_params"][:n])
        stab = st["stable_" + key][:n]
        print("%s: agree all %.4f, on stable %.4f (n=%d); chi2<=ref(1+1e-6): %.4f; on stable %.4f" % (
            key, ok.mean(), ok[stab].mean(), stab.sum(), np.mean(r["chi2"] <= g[key + "_fnorm"][:n] * (1 + 1e-6)),
            np.mean((r["chi2"] <= g[key + "_fnorm"][:n] * (1 + 1e-6))[stab])))
    print("status", dict(zip(*np.unique(r["status"], return_counts=True))), "mean niter %.2f nfev %.2f" % (r["niter"].mean(), r["nfev"].mean()))
    ex = np.concatenate(LMPAR_EXTRA) if LMPAR_EXTRA else np.zeros(0, dtype=np.int64)
    ticks = int(np.sum(LMPAR_TICKS))
    print("lmpar: %d fit-ticks, %.3f damped, extra iterations per damped tick: mean %.2f, histogram 1..10 %s" % (
        ticks, len(ex) / max(ticks, 1), ex.mean() if len(ex) else 0, np.bincount(ex, minlength=11)[1:11].tolist()))
    # what a warp pays: the maximum over its 32 lanes (fits drawn at random per tick)
    rng = np.random.default_rng(0)
    per_lane = np.zeros(ticks, dtype=np.int64)
    per_lane[:len(ex)] = ex
    rng.shuffle(per_lane)
    w = per_lane[:ticks // 32 * 32].reshape(-1, 32)
    print("       per warp-tick (32 random lanes): mean of the maximum %.2f, mean over lanes %.2f" % (w.max(axis=1).mean(), w.mean()))
    np.savez("/tmp/fast_proto_%s_%s.npz" % (a.dtype, a.acc), **r)


if __name__ == "__main__":
    main()
