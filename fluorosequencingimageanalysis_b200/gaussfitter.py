"""
Drop-in mirror of the reference's ``gaussfitter`` 2-D entry points (agpy/gaussfitter.py:29-255):
``gaussfit``, ``twodgaussian``, ``moments`` -- same signatures and return selection.  The fit
itself (the ``mpfit`` call of gaussfitter.py:243) runs in the CUDA LM kernel of libfsq.so.

What cannot be dropped in: ``mpfit`` with an arbitrary Python
callable.  The GPU path covers every form of ``gaussfit``: the 7-parameter model and its reduced
forms (``circle``, ``rotate=0``, ``vheight=0``), ``fixed`` parameters, ``err`` weights, and the
``perror`` / ``covar`` of the mpfit object (fsq_gaussfit_batch_ex).
"""
import numpy as np
from numpy import pi
from numpy.ma import median

from . import engine


def moments(data, circle, rotate, vheight, estimator=median, **kwargs):
    """agpy/gaussfitter.py:29-61 -> [height,] amplitude, x, y, width_x, width_y[, 0.] (or one mean
    width for circle=1).  The statistics come from the device kernel behind ``fsq_moments``; a
    non-default ``estimator`` (e.g. numpy.ma.median) replaces the height only, as in the reference."""
    data = np.asarray(data)
    m = engine.moments_batch(data[None])[0].cpu().numpy()
    height, amplitude, x, y, width_x, width_y = (float(v) for v in m[:6])
    if estimator is not median:
        height = estimator(data.ravel())
        amplitude = data.max() - height
    if any(np.isnan(v) for v in (width_y, width_x, height, amplitude)):
        raise ValueError("something is nan")
    head = [height] if vheight == 1 else []
    if circle != 0:
        return head + [amplitude, int(x), int(y), (width_x + width_y) / 2.]
    tail = [0.] if rotate == 1 else []
    return head + [amplitude, int(x), int(y), width_x, width_y] + tail


def twodgaussian(inpars, circle=False, rotate=True, vheight=True, shape=None):
    """agpy/gaussfitter.py:63-140.  NOTE the third parameter is the centre along axis 1 and
    the fourth along axis 0 of numpy.indices (gaussfitter.py:100)."""
    inpars_old = inpars
    inpars = list(inpars)
    height = float(inpars.pop(0)) if vheight == 1 else float(0)
    amplitude, center_y, center_x = float(inpars.pop(0)), float(inpars.pop(0)), float(inpars.pop(0))
    if circle == 1:
        width_x = width_y = float(inpars.pop(0))
        rotate = 0
    else:
        width_x, width_y = float(inpars.pop(0)), float(inpars.pop(0))
    if rotate == 1:
        rota = pi / 180. * float(inpars.pop(0))
        rcen_x = center_x * np.cos(rota) - center_y * np.sin(rota)
        rcen_y = center_x * np.sin(rota) + center_y * np.cos(rota)
    else:
        rota = 0.
        rcen_x, rcen_y = center_x, center_y
    if len(inpars) > 0:
        raise ValueError("There are still input parameters:" + str(inpars) +
                         " and you've input: " + str(inpars_old) +
                         " circle=%d, rotate=%d, vheight=%d" % (circle, rotate, vheight))

    def rotgauss(x, y):
        if rotate == 1:
            xp = x * np.cos(rota) - y * np.sin(rota)
            yp = x * np.sin(rota) + y * np.cos(rota)
        else:
            xp, yp = x, y
        return height + amplitude * np.exp(-(((rcen_x - xp) / width_x) ** 2 +
                                             ((rcen_y - yp) / width_y) ** 2) / 2.)
    if shape is not None:
        return rotgauss(*np.indices(shape))
    return rotgauss


class MpfitResult(object):
    """The attributes of the reference's mpfit object that gaussfit callers read
    (agpy/mpfit/mpfit.py:749-838)."""
    _ERRMSG = {0: 'ERROR: parameters are not within PARINFO limits',
               -16: "ERROR: parameter or function value(s) have become infinite; "
                    "check model function for over- and underflow"}

    def __init__(self, params, perror, status, niter, nfev, fnorm, dof, n_qrsolv, covar=None):
        self.params = params
        self.perror = perror
        self.covar = covar           # mpfit.py:1361-1388 (None unless the fit converged)
        self.status = status
        self.niter = niter
        self.nfev = nfev
        self.fnorm = fnorm
        self.dof = dof
        self.errmsg = self._ERRMSG.get(status, '')
        self.n_qrsolv = n_qrsolv     # extra: 0 <=> robust set


FAITHFUL = True
#: "minpack" (the reference's algorithm, default) or "fast" (the production fitter) -- see pflib.SOLVER
SOLVER = "minpack"


def gaussfit(data, err=None, params=(), autoderiv=True, return_all=False, circle=False,
             fixed=np.repeat(False, 7), limitedmin=[False, False, False, False, True, True, True],
             limitedmax=[False, False, False, False, False, False, True],
             usemoment=np.array([], dtype='bool'),
             minpars=np.repeat(0, 7), maxpars=[0, 0, 0, 0, 0, 0, 360],
             rotate=1, vheight=1, quiet=True, returnmp=False,
             returnfitimage=False, **kwargs):
    """agpy/gaussfitter.py:142-255."""
    data = np.asarray(data)
    usemoment = np.array(usemoment, dtype='bool')
    params = np.array(params, dtype='float')
    if autoderiv == 0:
        raise ValueError("I'm sorry, I haven't implemented this feature yet.")   # gaussfitter.py:239
    if usemoment.any() and len(params) == len(usemoment):
        moment = np.array(moments(data, circle, rotate, vheight, **kwargs), dtype='float')
        params[usemoment] = moment[usemoment]
    elif len(params) == 0:
        params = np.array(moments(data, circle, rotate, vheight, **kwargs), dtype='float')
    # ---- the parameter layouts of gaussfitter.py:195-232, mapped onto the kernel's 7 slots
    #      (height, amplitude, p2, p3, width_x, width_y, rota):
    #   vheight=0: a 0 is put in front and fixed[0] is set -- in the caller's array when the default is used, like the
    #              reference's mutable default (gaussfitter.py:195-198)
    #   circle=1 : parinfo has no entries 5 / 6 (:224-232), the model ties width_y to width_x (:104-107)
    #   rotate=0 : parinfo has no entry 6; rota = 0 makes the rotated model the unrotated one bit for bit
    params = np.array(params, dtype='float')
    if vheight == 0:
        vheight = 1
        params = np.concatenate([[0], params])
        fixed[0] = 1
    n_par = 5 if circle else (7 if rotate == 1 else 6)
    if len(params) < n_par:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (len(params), len(params)))   # parinfo, :224-232
    lmin = np.zeros(7, dtype=bool)
    lmax = np.zeros(7, dtype=bool)
    mn = np.zeros(7)
    mx = np.zeros(7)
    fx = np.zeros(7, dtype=np.uint8)
    for i in range(len(params)):                                             # gaussfitter.py:202-204
        if params[i] > maxpars[i] and limitedmax[i]:
            params[i] = maxpars[i]
        if params[i] < minpars[i] and limitedmin[i]:
            params[i] = minpars[i]
    p7 = np.zeros(7)
    for i in range(n_par):
        p7[i] = params[i]
        lmin[i], lmax[i], mn[i], mx[i] = bool(limitedmin[i]), bool(limitedmax[i]), float(minpars[i]), float(maxpars[i])
        fx[i] = 1 if fixed[i] else 0
    if circle:
        p7[5] = p7[4]
    elif rotate != 1:
        fx[6] = 1                                                            # rota stays 0
    general = bool(circle) or bool(fx.any()) or err is not None
    win = data[None].astype(np.int64) if data.dtype.kind in "iub" else data[None].astype(np.float64)
    err_a = None
    if err is not None:
        err_a = np.broadcast_to(np.asarray(err, dtype=np.float64), data.shape)[None]
    # SOLVER = "fast" runs the production fitter where it applies (5x5 / 11x11 windows, the plain 7-parameter model,
    # no perror / covar asked for)
    fast_ok = (SOLVER == "fast" and win.shape[1] == win.shape[2] and win.shape[1] in (5, 11)
               and not (return_all or returnmp) and not general)
    want_cv = bool(returnmp)
    r = engine.gaussfit_batch(win, p7[None], mn[None], mx[None], lmin[None].astype(np.uint8),
                              lmax[None].astype(np.uint8), faithful=FAITHFUL, solver="fast" if fast_ok else "minpack",
                              want_perror=bool(return_all or returnmp), want_fit_img=bool(returnfitimage),
                              fixed=fx[None] if general else None, err=err_a, circle=bool(circle), want_covar=want_cv)
    p = r.params[0].cpu().numpy()[:n_par]
    status = int(r.status[0].item())
    perror = covar = None
    if r.perror is not None and status > 0:
        perror = r.perror[0].cpu().numpy()[:n_par]
    if r.covar is not None and status > 0:
        covar = r.covar[0].cpu().numpy()[:n_par, :n_par]
    if returnmp:
        n_free = n_par - int(fx[:n_par].sum())
        returns = MpfitResult(p, perror, status, int(r.niter[0].item()), int(r.nfev[0].item()),
                              float(r.chi2[0].item()), data.size - n_free, int(r.n_qrsolv[0].item()), covar)
    elif return_all == 0:
        returns = p
    else:
        returns = p, perror
    if returnfitimage:
        returns = (returns, r.fit_img[0].cpu().numpy())
    return returns
