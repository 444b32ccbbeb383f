"""Comparison helpers shared by the parity tests (tie- and conditioning-aware)."""
import numpy as np

ODF_TOL = 1e-4        # north_star: ODF amplitudes within 1e-4 relative (to the voxel's max |ODF|)
SCALAR_TOL = 1e-4     # FA / MD / eigenvalues within 1e-4 relative
V1_DOT = 0.9999       # abs(dot) >= 0.9999 for V1
TIE_TAU = 1e-5        # a peak-index mismatch is "explained" iff the competing values differ by
                      # <= TIE_TAU * max|odf| in the float64 oracle


def flat(vol, nfr=None):
    v = np.asarray(vol)
    if nfr is None:
        return v.reshape(-1, order="F")
    return v.reshape((-1, nfr), order="F")


def odf_rel_err(got, ref64):
    """max over voxels of max_v |got - ref| / max_v |ref|  (voxels with ref == 0 must be exactly 0)."""
    g = flat(got, got.shape[-1]).astype(np.float64); r = flat(ref64, ref64.shape[-1]).astype(np.float64)
    scale = np.abs(r).max(axis=1)
    z = scale == 0
    assert np.all(g[z] == 0), "voxels the reference leaves untouched must stay exactly zero"
    err = np.abs(g - r).max(axis=1)[~z] / scale[~z]
    return float(err.max()) if err.size else 0.0


def peak_mismatch_report(idx_got, oracle64, nbr, tau=TIE_TAU):
    """Compare peak indices [.., 3] with the float64 oracle.  Returns (n_mismatch, n_unexplained).

    A mismatching voxel is explained iff re-running the peak rule on the float64 ODF perturbed
    within +-tau*max can reproduce the GPU answer: we accept it when every index the two answers
    disagree on belongs to a vertex whose candidate status or rank is decided by a margin
    <= tau*max|odf| in float64."""
    ig = flat(idx_got, 3).astype(np.int64); io = flat(oracle64["peak_idx"], 3).astype(np.int64)
    odf = flat(oracle64["odf"], oracle64["odf"].shape[-1]).astype(np.float64)
    bad = np.nonzero((ig != io).any(axis=1))[0]
    unexplained = 0
    for v in bad:
        o = odf[v]; m = np.abs(o).max(); eps = tau * m
        # margin of each vertex: how far it is from changing candidate status
        nb = nbr
        pad = np.concatenate([o, [-np.inf]])
        nmax = pad[np.where(nb < 0, o.shape[0], nb)].max(axis=1)
        margin = np.minimum(o - nmax, o)            # > 0 : candidate
        cand_lo = margin > eps                      # certainly candidates
        cand_hi = margin > -eps                     # possibly candidates
        sure = np.nonzero(cand_lo)[0]
        maybe = np.nonzero(cand_hi)[0]
        got = [i for i in ig[v] if i >= 0]
        ok = all(i in maybe for i in got)
        # ranking: got must be sorted by value desc up to eps, and no sure candidate with a value
        # more than eps above the smallest reported may be missing (if fewer than 3 reported: none missing)
        vals = o[got]
        ok = ok and all(vals[i] >= vals[i + 1] - eps for i in range(len(got) - 1))
        thresh = vals[-1] if len(got) == 3 else -np.inf
        missing = [i for i in sure if i not in got and o[i] > thresh + eps]
        ok = ok and not missing
        if len(got) < 3:
            # fewer peaks reported: every certain candidate must be reported
            ok = ok and all(i in got for i in sure)
        unexplained += (not ok)
    return len(bad), unexplained


def rel_err(got, ref, where=None):
    g = flat(got).astype(np.float64); r = flat(ref).astype(np.float64)
    if where is not None:
        g = g[where]; r = r[where]
    d = np.abs(g - r) / np.maximum(np.abs(r), 1e-30)
    return float(np.nanmax(d)) if d.size else 0.0
