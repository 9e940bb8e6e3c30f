"""GPU parity table: BOTH solvers against reference-generated fits for every BASELINE spot density.

Golden cases (oracle/make_golden.py FITS5_CASES / fits11: the reference itself, oracle/_ref, run on seeded frames):
  seed0      configs[0]/[1]: 512x512, 500 spots, all 4988 candidates
  seed3      a second seed, all 5036 candidates
  dense1000  configs[2]/[4] density: 512x512, 1000 spots, all 7978 candidates
  d2048      configs[3]: the 2048x2048 / 20 000-spot frame, the 3341 candidates of a 320x320 region
  fits11_d2048  configs[3]: 400 windows 11x11 cut from that dense frame, default gaussfit arguments

For each case and for the production solver (FAST) and the parity instrument (MINPACK, faithful = 1) the table holds the
per-fit agreement with the reference (north_star tolerances: H, A, widths 1e-4 relative, centres 1e-3 px, theta
excluded, converged flags equal) on
  * SURVEY.md 8(c)'s robust set (golden n_qrsolv == 0: the reference never left the Gauss-Newton branch of lmpar),
  * the fits the reference accepts (R^2 >= 0.7 -- the ones that become PSFs),
  * all candidates,
  * the reference-stable set (oracle/stability.py: the reference's own answer survives +-1 ulp of exp()),
next to the agreement of the R^2 gate and of the final PSF keys.  Every figure is an asserted floor (measured on
B200, round 2, minus a small margin), so a regression of any of them fails the suite instead of hiding in -s output.

What the numbers mean: the pflib call starts every fit with theta = 0 ON its lower bound and
width_x == width_y, where the model does not depend on theta.  The reference's finite-difference Jacobian column for
theta is therefore rounding noise of exp() and its first step a coin flip (oracle/stability.py) -- only 27 % of its
answers survive a 1-ulp perturbation of exp().  Per-fit equality with the reference is attainable, and asserted, on the
fits whose reference answer is reproducible; elsewhere the table reports where the solvers land."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, relerr
from test_gpu_fit import agree

pytestmark = pytest.mark.gpu


def _mods():
    from fluorosequencingimageanalysis_b200 import engine, pflib, synth
    return engine, pflib, synth


def _frame(case):
    _, _, synth = _mods()
    if case == "seed0":
        return synth.synth_frame(0)
    if case == "seed3":
        return synth.synth_frame(3)
    if case == "dense1000":
        return synth.synth_frame(11, n_spots=1000)
    if case == "d2048":
        return synth.synth_frame(4, H=2048, W=2048, n_spots=20000)
    raise KeyError(case)


def model25(P):
    """twodgaussian (agpy/gaussfitter.py:100-136) for n parameter vectors [n,7] on the 5x5 grid -> [n,25], numpy"""
    r, c = np.indices((5, 5))
    r, c = r.reshape(1, 25).astype(float), c.reshape(1, 25).astype(float)
    th = P[:, 6:7] * (np.pi / 180.0)
    cs, sn = np.cos(th), np.sin(th)
    rp, cp = r * cs - c * sn, r * sn + c * cs
    R0 = P[:, 3:4] * cs - P[:, 2:3] * sn
    C0 = P[:, 3:4] * sn + P[:, 2:3] * cs
    return P[:, 0:1] + P[:, 1:2] * np.exp(-(((R0 - rp) / P[:, 4:5]) ** 2 + ((C0 - cp) / P[:, 5:6]) ** 2) / 2.0)


_CACHE = {}


def fits(case, solver):
    """-> dict(P [n,7] window parameters, status, r2, fit [n,12], hw, subs [n,25], g = golden)"""
    key = (case, solver)
    if key in _CACHE:
        return _CACHE[key]
    engine, _, _ = _mods()
    import torch
    g = golden("fits5_%s.npz" % case)
    img = _frame(case)
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == str(g["img_sha"]), \
        "synthetic generator drifted from the golden frame"
    hw = np.ascontiguousarray(g["cands"].astype(np.int32))
    if (case, "det") not in _CACHE:                      # the candidate list itself: bit-exact against the reference's
        det = engine.detect_batch(img)
        got = det.cand_hw[:det.total].cpu().numpy()
        assert det.total == int(g["n_cands_frame"]) if "n_cands_frame" in g.files else det.total == len(hw)
        if case == "d2048":
            m = (got[:, 0] >= 800) & (got[:, 0] < 1120) & (got[:, 1] >= 800) & (got[:, 1] < 1120)
            got = got[m]
        assert np.array_equal(got, hw), "candidate list differs from the reference's"
        _CACHE[(case, "det")] = True
    name, faithful = {"fast": ("fast", False), "minpack": ("minpack", True), "minpack-clean": ("minpack", False)}[solver]
    hw_d = torch.from_numpy(hw).cuda()
    fr_d = torch.zeros(len(hw), dtype=torch.int32, device="cuda")
    fit, ints, _ = engine.fit_candidates(img, hw_d, fr_d, len(hw), solver=name, faithful=faithful)
    fit, ints = fit.cpu().numpy(), ints.cpu().numpy()
    P = fit[:, [2, 3, 0, 1, 4, 5, 6]].copy()
    P[:, 2] = fit[:, 0] - hw[:, 0] + 2.5
    P[:, 3] = fit[:, 1] - hw[:, 1] + 2.5
    subs = np.stack([img[h - 2:h + 3, w - 2:w + 3].astype(np.float64).reshape(25) for h, w in hw])
    out = dict(P=P, status=ints[:, 0], r2=fit[:, 8], fit=fit, hw=hw, hw_d=hw_d, fr_d=fr_d, subs=subs, g=g, shape=img.shape)
    _CACHE[key] = out
    return out


def final_keys(f):
    """R^2 gate + consolidation + re-key on the device for the fits of `f` -> set of (h, w) keys"""
    engine, _, _ = _mods()
    import torch
    n = len(f["hw"])
    c = engine.consolidate_batch(f["hw_d"], f["fr_d"], torch.from_numpy(f["fit"]).cuda(), n, 1)
    c.check()
    st = c.state.cpu().numpy()[:n]
    return set(map(tuple, c.key.cpu().numpy()[:n][st >= engine.PSF_FINAL].tolist()))


# floors: (robust, accepted, all, R^2 gate agreement, identical final keys / reference keys) = measured on B200 (round 2)
# minus ~0.01.  Measured:            fast                                      minpack (faithful)
#   seed0      0.9765 0.5938 0.3488 0.8955 0.8248          0.9403 0.7644 0.5956 0.9541 0.8889
#   seed3      0.9738 0.5876 0.3527 0.9009 0.8414          0.9417 0.7563 0.6005 0.9565 0.8879
#   dense1000  0.9762 0.5735 0.3541 0.8936 0.8133          0.9380 0.7562 0.6024 0.9544 0.8797
#   d2048      0.9849 0.5836 0.3717 0.8949 0.7907          0.9428 0.7683 0.6184 0.9533 0.9018
# The reference's own reproducibility under +-1 ulp of exp() (seed0, K = 8): robust 0.852, accepted 0.694, all 0.538.
FLOORS = {
    "seed0":     ((0.966, 0.583, 0.338, 0.885, 0.810), (0.930, 0.754, 0.585, 0.944, 0.875)),
    "seed3":     ((0.963, 0.577, 0.342, 0.890, 0.825), (0.931, 0.746, 0.590, 0.946, 0.875)),
    "dense1000": ((0.966, 0.563, 0.344, 0.883, 0.800), (0.928, 0.746, 0.592, 0.944, 0.865)),
    "d2048":     ((0.970, 0.570, 0.360, 0.883, 0.775), (0.930, 0.755, 0.605, 0.942, 0.885)),
}


@pytest.mark.parametrize("case", ["seed0", "seed3", "dense1000", "d2048"])
def test_parity_table(case):
    rows = []
    for si, solver in enumerate(("fast", "minpack")):
        f = fits(case, solver)
        g = f["g"]
        ok = agree(f["P"], g["ref_params"]) & (f["status"] > 0) & (g["ref_status"] > 0)
        robust = g["n_qrsolv"] == 0
        accepted = g["r_2"] >= 0.7
        gate = ((f["r2"] >= 0.7) == accepted).mean()
        keys = final_keys(f)
        want = set(map(tuple, g["final_keys"].tolist()))
        kfrac = len(keys & want) / max(len(want), 1)
        a = np.array(sorted(want), dtype=float)
        b = np.array(sorted(keys), dtype=float)
        near = (np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)).min(axis=1) <= 1.5).mean()
        row = (ok[robust].mean(), ok[accepted].mean(), ok.mean(), gate, kfrac)
        line = ("parity[%s] %-7s robust(n_qrsolv==0, n=%d) %.4f | reference-accepted (n=%d) %.4f | all (n=%d) %.4f | "
                "R^2 gate agrees %.4f | final PSFs: ref %d ours %d identical keys %.4f, ref PSFs with ours within 1.5 px %.4f"
                % (case, solver, robust.sum(), row[0], accepted.sum(), row[1], len(ok), row[2], gate, len(want), len(keys), kfrac, near))
        sfile = os.path.join(GOLDEN, "stable5_%s.npz" % case)
        if os.path.exists(sfile):
            st = np.load(sfile)
            stable = st["stable_ref"]
            ens = st["ens_agree_ref"]                      # [K, n]: perturbed reference run k reproduces the reference
            line += (" | reference-stable (n=%d) %.4f, robust & stable (n=%d) %.4f; the reference's own reproducibility under "
                     "1 ulp of exp(): robust %.4f accepted %.4f all %.4f"
                     % (stable.sum(), ok[stable].mean(), (robust & stable).sum(), ok[robust & stable].mean(),
                        ens[:, robust].mean(), ens[:, accepted].mean(), ens.mean()))
            # per-fit equality wherever the reference's answer is reproducible AND its trajectory is a clean one.  The
            # stable flags of seed0 come from K = 8 perturbed runs; the other cases have K = 3 (a weaker filter: some
            # fits that a further perturbation would flip still count as stable), hence the looser bar there
            strict = int(st["k"]) >= 8
            assert ok[robust & stable].mean() >= ((0.999 if strict else 0.985) if solver == "fast" else 0.96), line
            if solver == "fast":
                # ... and every miss on the survey's robust set is a fit whose REFERENCE answer is not reproducible
                miss = robust & ~ok
                assert (miss & stable).sum() <= max(1, int((0.001 if strict else 0.012) * robust.sum())), line
        print(line)
        rows.append(row)
        assert near >= 0.95, line
        for v, floor, what in zip(row, FLOORS[case][si], ("robust", "accepted", "all", "gate", "keys")):
            assert v >= floor, "%s: %s %.4f below its floor %.4f\n%s" % (solver, what, v, floor, line)


@pytest.mark.parametrize("case", ["seed0", "seed3", "dense1000", "d2048"])
def test_chi2_against_the_clean_oracle_both_directions(case):
    """chi^2 at the returned parameters (recomputed on the host from the parameters, same model code for both
    sides) of FAST and of the GPU clean-MINPACK kernel against the clean (defect-free) oracle.  Where the two differ by
    more than 1e-3 it is the theta peg: the pflib call starts theta = 0 on its lower bound with equal widths, the first
    steps decide whether theta leaves the bound, and a solver that stays pegged ends in the axis-aligned stationary
    point of the box-constrained problem (chi^2 ~2 % higher) -- in either direction, about equally often."""
    out = []
    for solver in ("fast", "minpack-clean"):
        f = fits(case, solver)
        g = f["g"]
        chi = ((f["subs"] - model25(f["P"])) ** 2).sum(axis=1)
        chc = ((f["subs"] - model25(g["clean_params"])) ** 2).sum(axis=1)
        worse, better = chi > chc * (1 + 1e-3), chc > chi * (1 + 1e-3)
        peg = (f["P"][:, 6] == 0) | (f["P"][:, 6] == 360)
        pegc = (g["clean_params"][:, 6] == 0) | (g["clean_params"][:, 6] == 360)
        diff = worse | better
        expl_w = (peg & ~pegc)[worse].mean() if worse.any() else 1.0       # we stayed pegged, the oracle did not
        expl_b = (~peg & pegc)[better].mean() if better.any() else 1.0     # the oracle stayed pegged, we did not
        same = ~diff
        line = ("chi2[%s] %-13s vs clean oracle: worse by >1e-3 %.4f (theta pegged only on our side: %.3f of them), better by "
                ">1e-3 %.4f (theta pegged only on the oracle's side: %.3f), equal %.4f; theta pegged overall: ours %.3f oracle %.3f; "
                "median chi2 ratio where worse %.4f, where better %.4f"
                % (case, solver, worse.mean(), expl_w, better.mean(), expl_b, same.mean(), peg.mean(), pegc.mean(),
                   np.median((chi / chc)[worse]) if worse.any() else 1.0, np.median((chi / chc)[better]) if better.any() else 1.0))
        print(line)
        out.append((worse.mean(), better.mean(), expl_w, expl_b))
        assert worse.mean() <= 0.10 and better.mean() >= worse.mean() - 0.03, line
        assert expl_w >= 0.90 and expl_b >= 0.88, line
        # same peg state on both sides => same stationary point
        agree_peg = peg == pegc
        assert (same[agree_peg]).mean() >= 0.985, line


def test_fast_drop_in_final_psf_keys_as_a_set(fits5, frame0):
    """pflib.SOLVER = 'fast' through the drop-in entry point: the final PSF dictionary against the reference's 468
    keys as SET OVERLAP (identical (h, w) keys), not as a distance."""
    _, pflib, _ = _mods()
    old = pflib.SOLVER
    pflib.SOLVER = "fast"
    try:
        out = pflib.find_peptides(frame0)
    finally:
        pflib.SOLVER = old
    want = set(tuple(k) for k in fits5["final_keys"].tolist())
    got = set(out.keys())
    assert got == final_keys(fits("seed0", "fast"))            # drop-in == packed path
    frac = len(want & got) / len(want)
    print("fast drop-in: ref %d PSFs, ours %d, identical keys %d (%.4f)" % (len(want), len(got), len(want & got), frac))
    assert frac >= FLOORS["seed0"][0][4]
    out_m = pflib.find_peptides(frame0)
    got_m = set(out_m.keys())
    frac_m = len(want & got_m) / len(want)
    print("faithful drop-in: ref %d PSFs, ours %d, identical keys %d (%.4f)" % (len(want), len(got_m), len(want & got_m), frac_m))
    assert frac_m >= FLOORS["seed0"][1][4]


@pytest.mark.parametrize("name", ["seed0", "d2048"])
def test_parity_table_11x11(name):
    """11x11 windows, gaussfit's default arguments (moments start, gaussfitter.py:142-148): isolated spots (configs[0])
    and windows cut from the dense 2048x2048 frame (configs[3], neighbouring spots inside most windows)."""
    engine, _, _ = _mods()
    g = golden("fits11_%s.npz" % name)
    n = len(g["windows"])
    lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
    t = lambda v: np.tile(v, (n, 1))
    p0 = engine.moments_batch(g["windows"], lo, hi, lmin, lmax).cpu().numpy()
    assert np.array_equal(p0, g["p0"]) or np.allclose(p0, g["p0"], rtol=0, atol=0, equal_nan=True), "moments start differs"
    robust = g["n_qrsolv"] == 0
    sfile = os.path.join(GOLDEN, "stable11_%s.npz" % name)
    stable = np.load(sfile)["stable_ref"] if os.path.exists(sfile) else None
    # (robust-set agreement, chi^2 not worse than the reference); measured: fast 1.0000 / 0.9900 and 1.0000 / 0.9875,
    # minpack 1.0000 / 0.9150 and 1.0000 / 0.9525 (the faithful kernel reproduces the reference's premature exits)
    floors = {"seed0": {"fast": (0.99, 0.98), "minpack": (0.99, 0.90)}, "d2048": {"fast": (0.99, 0.975), "minpack": (0.99, 0.94)}}[name]
    for solver, faithful in (("fast", False), ("minpack", True)):
        r = engine.gaussfit_batch(g["windows"], g["p0"], t(lo), t(hi), t(lmin), t(lmax), solver=solver, faithful=faithful)
        P, s, chi = r.params.cpu().numpy(), r.status.cpu().numpy(), r.chi2.cpu().numpy()
        ok = agree(P, g["ref_params"]) & (s > 0) & (g["ref_status"] > 0)
        nw = chi <= g["ref_fnorm"] * (1 + 1e-6)
        line = ("parity11[%s] %-7s robust (n=%d) %.4f | all (n=%d) %.4f | chi2 not worse than the reference %.4f | status>0 %.4f"
                % (name, solver, robust.sum(), ok[robust].mean(), n, ok.mean(), nw.mean(), (s > 0).mean()))
        if stable is not None:
            line += " | robust & stable (n=%d) %.4f" % ((robust & stable).sum(), ok[robust & stable].mean())
            assert ok[robust & stable].mean() >= 0.99, line
        print(line)
        assert (s > 0).all(), line
        assert ok[robust].mean() >= floors[solver][0], line
        assert nw.mean() >= floors[solver][1], line
