"""
TEST INFRASTRUCTURE -- CPU oracle, not product code.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module.

Restatement (Python 3 / numpy) of the bounded Levenberg-Marquardt solver the reference
runs for every candidate window: ``class mpfit`` of /root/reference/agpy/mpfit/mpfit.py,
as exercised by ``gaussfitter.gaussfit`` (all parameters free, box limits, finite-difference
Jacobian, quiet, default tolerances).  Written as plain functions over explicit state, not
as a transliteration of the class; every block cites the reference lines it follows.

Two flavours:

  faithful=True   ``qrsolv`` works on a *view* of R's diagonal exactly as the reference does
                  on numpy >= 1.9 (mpfit.py:1915, 1956, 1976-1977): the "restore the diagonal"
                  store is a no-op and the solution vector is scattered into R's diagonal.
  faithful=False  clean MINPACK: the diagonal is saved in a copy and restored.

Parity pin: tests/test_oracle_pins.py checks this module bit-for-bit against the reference
itself (oracle/_ref, built by oracle/build_ref.py) wherever /root/reference is available,
and against the committed golden vectors in tests/golden/ (generated from oracle/_ref by
oracle/make_golden.py) everywhere else.  Known-answer vector KAT-1: SURVEY.md App. D.

Unsupported on purpose (never reached from gaussfit): tied parameters, user step sizes,
two-sided derivatives, mpmaxstep, damping, analytic derivatives, iterfunct.
"""
import numpy as np

MACHEP = float(np.finfo(np.float64).eps)       # mpfit.py:2345 (machar, double)
DWARF = float(np.finfo(np.float64).tiny)       # mpfit.py:2347


def enorm(v):
    """mpfit.py:1504-1509 -- sqrt(dot(v, v)), no scaling."""
    return np.sqrt(np.dot(v.T, v))


class LMResult(object):
    __slots__ = ("params", "perror", "covar", "status", "niter", "nfev", "fnorm", "dof",
                 "errmsg", "n_qrsolv", "n_reject", "trace")

    def __repr__(self):
        return "LMResult(status=%r, niter=%r, nfev=%r, fnorm=%r, params=%r)" % (
            self.status, self.niter, self.nfev, self.fnorm, self.params)


# --------------------------------------------------------------------------------------
# Jacobian by forward differences -- mpfit.py:1512-1612 (one-sided branch only)
# --------------------------------------------------------------------------------------
def fd_jacobian(resid, x, fvec, has_hi, hi, counter):
    eps = np.sqrt(np.max([MACHEP, MACHEP]))                 # :1529 (epsfcn defaults to machep)
    m, n = len(fvec), len(x)
    J = np.zeros([m, n], dtype=float)                       # :1555
    h = eps * np.abs(x)                                     # :1557
    h[h == 0] = eps                                         # :1576
    flip = (has_hi != 0) & (x > hi - h)                     # :1582-1584 (dside == 0 everywhere)
    h[flip] = -h[flip]                                      # :1587
    for j in range(n):                                      # :1589-1599
        xp = x.copy()
        xp[j] = xp[j] + h[j]
        counter[0] += 1
        fp = resid(xp)
        J[0:, j] = (fp - fvec) / h[j]
    return J


# --------------------------------------------------------------------------------------
# Householder QR with column pivoting -- mpfit.py:1748-1822
# --------------------------------------------------------------------------------------
def qr_pivot(a):
    m, n = a.shape
    acnorm = np.zeros(n, dtype=float)
    for j in range(n):
        acnorm[j] = enorm(a[:, j])                          # :1758-1759
    rdiag = acnorm.copy()
    wa = rdiag.copy()
    ipvt = np.arange(n)
    for j in range(min(m, n)):
        # pivot: first position holding the largest remaining norm (:1769-1783)
        rmax = np.max(rdiag[j:])
        kmax = np.nonzero(rdiag[j:] == rmax)[0]
        if len(kmax) > 0:
            kmax = kmax[0] + j
            if kmax != j:
                ipvt[j], ipvt[kmax] = ipvt[kmax], ipvt[j]
                rdiag[kmax] = rdiag[j]
                wa[kmax] = wa[j]
        lj = ipvt[j]
        ajj = a[j:, lj]
        ajnorm = enorm(ajj)                                 # :1789
        if ajnorm == 0:
            break                                           # :1790-1791
        if a[j, lj] < 0:
            ajnorm = -ajnorm
        ajj = ajj / ajnorm
        ajj[0] = ajj[0] + 1
        a[j:, lj] = ajj                                     # :1795-1798
        for k in range(j + 1, n):                           # :1806-1820
            lk = ipvt[k]
            ajk = a[j:, lk]
            if a[j, lj] != 0:
                a[j:, lk] = ajk - ajj * sum(ajk * ajj) / a[j, lj]
                if rdiag[k] != 0:
                    temp = a[j, lk] / rdiag[k]
                    rdiag[k] = rdiag[k] * np.sqrt(np.max([(1. - temp ** 2), 0.]))
                    temp = rdiag[k] / wa[k]
                    if (0.05 * temp * temp) <= MACHEP:
                        rdiag[k] = enorm(a[j + 1:, lk])
                        wa[k] = rdiag[k]
        rdiag[j] = -ajnorm                                  # :1821
    return a, ipvt, rdiag, acnorm


# --------------------------------------------------------------------------------------
# qrsolv -- mpfit.py:1903-1978, including the diagonal-view behaviour
# --------------------------------------------------------------------------------------
def qrsolv(r, ipvt, diag, qtb, sdiag, faithful):
    n = r.shape[1]
    for j in range(n):
        r[j:n, j] = r[j, j:n]                               # :1913-1914
    if faithful:
        x = np.diagonal(r)                                  # :1915 -- a VIEW on numpy>=1.9
    else:
        x = np.diagonal(r).copy()
    wa = qtb.copy()
    for j in range(n):                                      # :1919-1956
        l = ipvt[j]
        if diag[l] == 0:
            break
        sdiag[j:] = 0
        sdiag[j] = diag[l]
        qtbpj = 0.
        for k in range(j, n):
            if sdiag[k] == 0:
                break
            if np.abs(r[k, k]) < np.abs(sdiag[k]):
                cotan = r[k, k] / sdiag[k]
                sine = 0.5 / np.sqrt(.25 + .25 * cotan * cotan)
                cosine = sine * cotan
            else:
                tang = sdiag[k] / r[k, k]
                cosine = 0.5 / np.sqrt(.25 + .25 * tang * tang)
                sine = cosine * tang
            r[k, k] = cosine * r[k, k] + sine * sdiag[k]
            temp = cosine * wa[k] + sine * qtbpj
            qtbpj = -sine * wa[k] + cosine * qtbpj
            wa[k] = temp
            if n > k + 1:
                temp = cosine * r[k + 1:n, k] + sine * sdiag[k + 1:n]
                sdiag[k + 1:n] = -sine * r[k + 1:n, k] + cosine * sdiag[k + 1:n]
                r[k + 1:n, k] = temp
        sdiag[j] = r[j, j]
        r[j, j] = x[j]                                      # :1956 -- no-op when x is a view
    nsing = n                                               # :1960-1971
    wh = np.nonzero(sdiag == 0)[0]
    if len(wh) > 0:
        nsing = wh[0]
        wa[nsing:] = 0
    if nsing >= 1:
        wa[nsing - 1] = wa[nsing - 1] / sdiag[nsing - 1]
        for j in range(nsing - 2, -1, -1):
            sum0 = sum(r[j + 1:nsing, j] * wa[j + 1:nsing])
            wa[j] = (wa[j] - sum0) / sdiag[j]
    if faithful:
        x.setflags(write=True)                              # :1976
    x[ipvt] = wa                                            # :1977 -- lands in diag(r) when a view
    return r, x, sdiag


# --------------------------------------------------------------------------------------
# lmpar -- mpfit.py:2077-2190
# --------------------------------------------------------------------------------------
def lmpar(r, ipvt, diag, qtb, delta, x, sdiag, par, faithful, stats):
    n = r.shape[1]
    nsing = n
    wa1 = qtb.copy()
    rthresh = np.max(np.abs(np.diagonal(r))) * MACHEP       # :2091
    wh = np.nonzero(np.abs(np.diagonal(r)) < rthresh)[0]
    if len(wh) > 0:
        nsing = wh[0]
        wa1[wh[0]:] = 0
    if nsing >= 1:
        for j in range(nsing - 1, -1, -1):                  # :2098-2101
            wa1[j] = wa1[j] / r[j, j]
            if j - 1 >= 0:
                wa1[0:j] = wa1[0:j] - r[0:j, j] * wa1[j]
    x[ipvt] = wa1                                           # :2104
    it = 0
    wa2 = diag * x
    dxnorm = enorm(wa2)
    fp = dxnorm - delta
    if fp <= 0.1 * delta:                                   # :2112-2113 Gauss-Newton accepted
        return r, 0., x, sdiag
    parl = 0.                                               # :2119-2128
    if nsing >= n:
        wa1 = diag[ipvt] * wa2[ipvt] / dxnorm
        wa1[0] = wa1[0] / r[0, 0]
        for j in range(1, n):
            sum0 = sum(r[0:j, j] * wa1[0:j])
            wa1[j] = (wa1[j] - sum0) / r[j, j]
        temp = enorm(wa1)
        parl = ((fp / delta) / temp) / temp
    for j in range(n):                                      # :2131-2137
        sum0 = sum(r[0:j + 1, j] * qtb[0:j + 1])
        wa1[j] = sum0 / diag[ipvt[j]]
    gnorm = enorm(wa1)
    paru = gnorm / delta
    if paru == 0:
        paru = DWARF / np.min([delta, 0.1])
    par = np.max([par, parl])                               # :2142-2145
    par = np.min([par, paru])
    if par == 0:
        par = gnorm / dxnorm
    while True:                                             # :2148-2187
        it += 1
        if par == 0:
            par = np.max([DWARF, paru * 0.001])
        temp = np.sqrt(par)
        wa1 = temp * diag
        stats["n_qrsolv"] += 1
        r, x, sdiag = qrsolv(r, ipvt, wa1, qtb, sdiag, faithful)
        wa2 = diag * x
        dxnorm = enorm(wa2)
        temp = fp
        fp = dxnorm - delta
        if (np.abs(fp) <= 0.1 * delta) or ((parl == 0) and (fp <= temp) and (temp < 0)) or (it == 10):
            break
        wa1 = diag[ipvt] * wa2[ipvt] / dxnorm               # :2168-2176
        for j in range(n - 1):
            wa1[j] = wa1[j] / sdiag[j]
            wa1[j + 1:n] = wa1[j + 1:n] - r[j + 1:n, j] * wa1[j]
        wa1[n - 1] = wa1[n - 1] / sdiag[n - 1]
        temp = enorm(wa1)
        parc = ((fp / delta) / temp) / temp
        if fp > 0:
            parl = np.max([parl, par])
        if fp < 0:
            paru = np.min([paru, par])
        par = np.max([parl, par + parc])
    return r, par, x, sdiag


# --------------------------------------------------------------------------------------
# covariance -- mpfit.py:2274-2336
# --------------------------------------------------------------------------------------
def covar_from_r(rr, ipvt, tol=1.e-14):
    n = rr.shape[0]
    r = rr.copy()
    l = -1
    tolr = tol * np.abs(r[0, 0])
    for k in range(n):                                      # :2295-2303 inverse of R
        if np.abs(r[k, k]) <= tolr:
            break
        r[k, k] = 1. / r[k, k]
        for j in range(k):
            temp = r[k, k] * r[j, k]
            r[j, k] = 0.
            r[0:j + 1, k] = r[0:j + 1, k] - temp * r[0:j + 1, j]
        l = k
    if l >= 0:                                              # :2307-2313 (R^T R)^-1 upper triangle
        for k in range(l + 1):
            for j in range(k):
                temp = r[j, k]
                r[0:j + 1, j] = r[0:j + 1, j] + temp * r[0:j + 1, k]
            temp = r[k, k]
            r[0:k + 1, k] = temp * r[0:k + 1, k]
    wa = np.repeat([r[0, 0]], n)                            # :2317-2329 un-pivot
    for j in range(n):
        jj = ipvt[j]
        sing = j > l
        for i in range(j + 1):
            if sing:
                r[i, j] = 0.
            ii = ipvt[i]
            if ii > jj:
                r[ii, jj] = r[i, j]
            if ii < jj:
                r[jj, ii] = r[i, j]
        wa[jj] = r[j, j]
    for j in range(n):                                      # :2332-2334 symmetrise
        r[0:j + 1, j] = r[j, 0:j + 1]
        r[j, j] = wa[j]
    return r


# --------------------------------------------------------------------------------------
# driver -- mpfit.py:600-1388 restricted to what gaussfit uses
# --------------------------------------------------------------------------------------
def lm_solve(resid, x0, lim_lo, lim_hi, lo, hi, faithful=True,
             ftol=1.e-10, xtol=1.e-10, gtol=1.e-10, maxiter=200, factor=100., trace=None):
    """
    resid(p) -> residual vector (data - model), float64.
    x0      start values (n,), lim_lo/lim_hi bool (n,), lo/hi limits (n,).
    Returns LMResult with the attributes the reference's mpfit object exposes
    (params, perror, covar, status, niter, nfev, fnorm(=chi^2), dof, errmsg) plus the
    diagnostic counters n_qrsolv / n_reject (the robust set is
    n_qrsolv == 0).
    """
    res = LMResult()
    res.trace = trace
    res.niter = 0
    res.params = None
    res.covar = None
    res.perror = None
    res.status = 0
    res.errmsg = ''
    res.nfev = 0
    res.dof = 0
    res.n_qrsolv = 0
    res.n_reject = 0
    res.fnorm = -1.
    stats = {"n_qrsolv": 0}
    counter = [0]

    xall = np.asarray(x0)
    if xall.dtype.kind != 'f' or xall.dtype.itemsize <= 4:   # :901-902
        xall = xall.astype(float)
    n = len(xall)
    fnorm1 = -1.
    qllim = np.asarray(lim_lo).astype(int)                    # parinfo() -> int arrays (:952, :1478-1479)
    qulim = np.asarray(lim_hi).astype(int)
    llim = np.asarray(lo, dtype=float)
    ulim = np.asarray(hi, dtype=float)
    params = xall.copy()
    x = params.copy()
    if np.any((qllim & (xall < llim)) | (qulim & (xall > ulim))):      # :956-959
        res.errmsg = 'ERROR: parameters are not within PARINFO limits'
        res.params = params
        return res
    if np.any((qllim & qulim) & (llim >= ulim)):              # :960-964
        res.errmsg = 'ERROR: PARINFO parameter limits are not consistent'
        res.params = params
        return res
    qanylim = 1 if np.any((qulim != 0.) | (qllim != 0.)) else 0

    counter[0] += 1
    fvec = resid(params)                                      # :999
    m = len(fvec)
    if m < n:
        res.errmsg = 'ERROR: number of parameters must not exceed data'
        res.params = params
        return res
    res.dof = m - n
    fnorm = enorm(fvec)                                       # :1019
    par = 0.
    niter = 1
    qtf = x * 0.
    status = 0
    diag = None
    delta = xnorm = 0.
    nlpeg = nupeg = 0
    whlpeg = whupeg = None
    gnorm = 0.

    while True:                                               # outer loop :1030
        params[:] = x
        fjac = fd_jacobian(resid, x, fvec, qulim, ulim, counter)      # :1064
        if qanylim:                                           # :1073-1091
            whlpeg = np.nonzero(qllim & (x == llim))[0]
            nlpeg = len(whlpeg)
            whupeg = np.nonzero(qulim & (x == ulim))[0]
            nupeg = len(whupeg)
            for i in range(nlpeg):
                sum0 = sum(fvec * fjac[:, whlpeg[i]])
                if sum0 > 0:
                    fjac[:, whlpeg[i]] = 0
            for i in range(nupeg):
                sum0 = sum(fvec * fjac[:, whupeg[i]])
                if sum0 < 0:
                    fjac[:, whupeg[i]] = 0
        fjac, ipvt, wa1, wa2 = qr_pivot(fjac)                 # :1094
        if niter == 1:                                        # :1099-1110
            diag = wa2.copy()
            diag[diag == 0] = 1.
            wa3 = diag * x
            xnorm = enorm(wa3)
            delta = factor * xnorm
            if delta == 0.:
                delta = factor
        wa4 = fvec.copy()                                     # :1114-1124  Q^T f
        for j in range(n):
            lj = ipvt[j]
            temp3 = fjac[j, lj]
            if temp3 != 0:
                fj = fjac[j:, lj]
                wj = wa4[j:]
                wa4[j:] = wj - fj * sum(fj * wj) / temp3
            fjac[j, lj] = wa1[j]
            qtf[j] = wa4[j]
        fjac = fjac[0:n, 0:n]                                 # :1127-1132 R in pivot order
        temp = fjac.copy()
        for i in range(n):
            temp[:, i] = fjac[:, ipvt[i]]
        fjac = temp.copy()
        gnorm = 0.                                            # :1142-1148
        if fnorm != 0:
            for j in range(n):
                l = ipvt[j]
                if wa2[l] != 0:
                    sum0 = sum(fjac[0:j + 1, j] * qtf[0:j + 1]) / fnorm
                    gnorm = np.max([gnorm, np.abs(sum0 / wa2[l])])
        if trace is not None:
            trace.append(dict(kind="outer", niter=niter, x=x.copy(), fnorm=float(fnorm),
                              gnorm=float(gnorm), ipvt=ipvt.copy(), R=fjac.copy(),
                              qtf=qtf.copy(), acnorm=wa2.copy(), diag=diag.copy(),
                              delta=float(delta), par=float(par)))
        if gnorm <= gtol:                                     # :1151-1153
            status = 4
            break
        if maxiter == 0:
            status = 5
            break
        diag = np.choose(diag > wa2, (wa2, diag))             # :1160

        while True:                                           # inner loop :1163
            fjac, par, wa1, wa2 = lmpar(fjac, ipvt, diag, qtf, delta, wa1, wa2, par,
                                        faithful, stats)      # :1167
            wa1 = -wa1
            alpha = 1.
            if qanylim:                                       # :1184-1202
                if nlpeg > 0:
                    wa1[whlpeg] = np.clip(wa1[whlpeg], 0., np.max(wa1))
                if nupeg > 0:
                    wa1[whupeg] = np.clip(wa1[whupeg], np.min(wa1), 0.)
                dwa1 = np.abs(wa1) > MACHEP
                whl = np.nonzero(((dwa1 != 0.) & qllim) & ((x + wa1) < llim))[0]
                if len(whl) > 0:
                    t = ((llim[whl] - x[whl]) / wa1[whl])
                    alpha = np.min([alpha, np.min(t)])
                whu = np.nonzero(((dwa1 != 0.) & qulim) & ((x + wa1) > ulim))[0]
                if len(whu) > 0:
                    t = ((ulim[whu] - x[whu]) / wa1[whu])
                    alpha = np.min([alpha, np.min(t)])
            wa1 = wa1 * alpha                                 # :1215-1216
            wa2 = x + wa1
            if qanylim:
                sgnu = (ulim >= 0) * 2. - 1.                  # :1220-1231 snap to the bound
                sgnl = (llim >= 0) * 2. - 1.
                ulim1 = ulim * (1 - sgnu * MACHEP) - (ulim == 0) * MACHEP
                llim1 = llim * (1 + sgnl * MACHEP) + (llim == 0) * MACHEP
                wh = np.nonzero((qulim != 0) & (wa2 >= ulim1))[0]
                if len(wh) > 0:
                    wa2[wh] = ulim[wh]
                wh = np.nonzero((qllim != 0.) & (wa2 <= llim1))[0]
                if len(wh) > 0:
                    wa2[wh] = llim[wh]
            wa3 = diag * wa1
            pnorm = enorm(wa3)
            if niter == 1:
                delta = np.min([delta, pnorm])                # :1237-1238
            params[:] = wa2
            counter[0] += 1
            wa4 = resid(params)                               # :1245
            fnorm1 = enorm(wa4)
            actred = -1.                                      # :1253-1255
            if (0.1 * fnorm1) < fnorm:
                actred = - (fnorm1 / fnorm) ** 2 + 1.
            for j in range(n):                                # :1259-1261  R * (P^T p)
                wa3[j] = 0
                wa3[0:j + 1] = wa3[0:j + 1] + fjac[0:j + 1, j] * wa1[ipvt[j]]
            temp1 = enorm(alpha * wa3) / fnorm                # :1265-1268
            temp2 = (np.sqrt(alpha * par) * pnorm) / fnorm
            prered = temp1 * temp1 + (temp2 * temp2) / 0.5
            dirder = -(temp1 * temp1 + temp2 * temp2)
            ratio = 0.
            if prered != 0:
                ratio = actred / prered
            if ratio <= 0.25:                                 # :1276-1288
                if actred >= 0:
                    temp = .5
                else:
                    temp = .5 * dirder / (dirder + .5 * actred)
                if ((0.1 * fnorm1) >= fnorm) or (temp < 0.1):
                    temp = 0.1
                delta = temp * np.min([delta, pnorm / 0.1])
                par = par / temp
            else:
                if (par == 0) or (ratio >= 0.75):
                    delta = pnorm / .5
                    par = .5 * par
            accepted = ratio >= 0.0001
            if accepted:                                      # :1291-1298
                x = wa2
                wa2 = diag * x
                fvec = wa4
                xnorm = enorm(wa2)
                fnorm = fnorm1
                niter = niter + 1
            else:
                res.n_reject += 1
            if trace is not None:
                trace.append(dict(kind="inner", niter=niter, accepted=bool(accepted),
                                  ratio=float(ratio), actred=float(actred), prered=float(prered),
                                  delta=float(delta), par=float(par), alpha=float(alpha),
                                  pnorm=float(pnorm), fnorm1=float(fnorm1), step=wa1.copy(),
                                  xnew=params.copy(), n_qrsolv=stats["n_qrsolv"]))
            if (np.abs(actred) <= ftol) and (prered <= ftol) and (0.5 * ratio <= 1):   # :1301-1310
                status = 1
            if delta <= xtol * xnorm:
                status = 2
            if (np.abs(actred) <= ftol) and (prered <= ftol) and (0.5 * ratio <= 1) and (status == 2):
                status = 3
            if status != 0:
                break
            if niter >= maxiter:                              # :1313-1323
                status = 5
            if (np.abs(actred) <= MACHEP) and (prered <= MACHEP) and (0.5 * ratio <= 1):
                status = 6
            if delta <= MACHEP * xnorm:
                status = 7
            if gnorm <= MACHEP:
                status = 8
            if status != 0:
                break
            if accepted:                                      # :1326-1327
                break
            if (not np.all(np.isfinite(wa1) & np.isfinite(wa2) & np.isfinite(x))) \
                    or (not np.isfinite(ratio)):              # :1330-1335
                status = -16
                break
        if status != 0:
            break

    params[:] = x                                             # :1350
    if status > 0:                                            # :1351-1355
        counter[0] += 1
        fvec = resid(params)
        fnorm = enorm(fvec)
    fnorm = np.max([fnorm, fnorm1])                           # :1357-1359
    fnorm = fnorm ** 2.
    res.params = params
    res.status = status
    res.niter = niter
    res.nfev = counter[0]
    res.fnorm = fnorm
    res.n_qrsolv = stats["n_qrsolv"]
    if status > 0:                                            # :1364-1387
        cv = covar_from_r(fjac[0:n, 0:n], ipvt[0:n])
        res.covar = cv
        res.perror = np.zeros(n, dtype=float)
        d = np.diagonal(cv)
        wh = np.nonzero(d >= 0)[0]
        if len(wh) > 0:
            res.perror[wh] = np.sqrt(d[wh])
    return res
