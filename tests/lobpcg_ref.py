"""TEST INFRASTRUCTURE: NumPy restatement of the reference's LOBPCG loop (src/lobpcg/algorithm.rs:119-431) for a dense
operator, no preconditioner, no constraints -- the loop whose dense blocks tests/test_gpu_lobpcg_resident.py keeps on the
device.  Third-party arithmetic (ndarray `dot`, the crate's own eigh / cholesky / triangular solve) is restated with
NumPy / LAPACK by definition; what is pinned here is the SEQUENCE of block operations and the Rayleigh-Ritz bookkeeping."""
import numpy as np


def sorted_eig(a, b, size, largest=True):                                   # algorithm.rs:16-44
    if b is not None:
        vb, qb = np.linalg.eigh(b)
        qb = qb * (1.0 / np.sqrt(np.maximum(vb, np.float32(1e-10))))
        va, qa = np.linalg.eigh(qb.T @ (a @ qb))
        vals, vecs = va, qb @ qa
    else:
        vals, vecs = np.linalg.eigh(a)
    idx = np.argsort(-vals if largest else vals, kind="stable")
    vals, vecs = vals[idx], vecs[:, idx]
    vecs = vecs * np.where(np.signbit(vecs[0, :]), -1.0, 1.0)
    return vals[:size], vecs[:, :size]


def orthonormalize(v):                                                       # algorithm.rs:81-97
    l = np.linalg.cholesky(v.T @ v)
    return np.linalg.solve(l, v.T).T, l


def lobpcg(a_mat, x, tol, maxiter, largest=True):
    """Returns (lambda history, final lambda, final x, residual norms history)."""
    n, size_x = x.shape
    it = min(n * 10, maxiter)
    x, _ = orthonormalize(x)                                                 # :163
    ax = a_mat @ x                                                           # :166
    lam, eig_block = sorted_eig(x.T @ ax, None, size_x, largest)             # :167-171
    x, ax = x @ eig_block, ax @ eig_block                                    # :174-175
    active = np.ones(size_x, dtype=bool)
    prev_p_ap = None
    hist, rhist = [lam.copy()], []
    while True:
        r = ax - x * lam[None, :]                                            # :197-201
        rn = np.linalg.norm(r, axis=0)                                       # :204-208
        rhist.append(rn.copy())
        active = (rn > tol) & active                                         # :222-226
        cur = int(active.sum())
        if cur == 0 or it == 0:                                              # :237-239
            break
        ar_ = r[:, active].copy()                                            # :242 (identity preconditioner, no constraints)
        ar_ -= x @ (x.T @ ar_)                                               # :251-257
        r, _ = orthonormalize(ar_)                                           # :259-262
        ar = a_mat @ r                                                       # :264
        xar, rar = x.T @ ar, r.T @ ar                                        # :278-279
        rar = (rar + rar.T) / 2                                              # :285 (explicit_gram_flag starts true and stays true)
        xax = x.T @ ax
        xax, xx, rr, xr = (xax + xax.T) / 2, x.T @ x, r.T @ r, x.T @ r       # :286-293
        p_ap = None
        if prev_p_ap is not None:                                            # :305-321
            p, ap = prev_p_ap
            try:
                act_p, p_r = orthonormalize(p[:, active])
                p_ap = (act_p, np.linalg.solve(p_r, ap[:, active].T).T)
            except np.linalg.LinAlgError:
                p_ap = None
        res = None
        if p_ap is not None:                                                 # :327-358
            act_p, act_ap = p_ap
            xap, rap, pap = x.T @ act_ap, r.T @ act_ap, act_p.T @ act_ap
            xp, rp = x.T @ act_p, r.T @ act_p
            pap, pp = (pap + pap.T) / 2, act_p.T @ act_p
            ga = np.block([[xax, xar, xap], [xar.T, rar, rap], [xap.T, rap.T, pap]])
            gb = np.block([[xx, xr, xp], [xr.T, rr, rp], [xp.T, rp.T, pp]])
            try:
                res = sorted_eig(ga, gb, size_x, largest)
            except np.linalg.LinAlgError:
                res = None
        if res is None:                                                      # :359-378
            p_ap = None
            ga = np.block([[xax, xar], [xar.T, rar]])
            gb = np.block([[xx, xr], [xr.T, rr]])
            res = sorted_eig(ga, gb, size_x, largest)
        lam, ev = res                                                        # :383-389
        tau = ev[:size_x]
        if p_ap is not None:                                                 # :392-417
            act_p, act_ap = p_ap
            alpha, gamma = ev[size_x:size_x + cur], ev[size_x + cur:]
            p, ap = r @ alpha + act_p @ gamma, ar @ alpha + act_ap @ gamma
        else:
            alpha = ev[size_x:]
            p, ap = r @ alpha, ar @ alpha
        x, ax = x @ tau + p, ax @ tau + ap                                   # :420-421
        prev_p_ap = (p, ap)
        hist.append(lam.copy())
        it -= 1
    return hist, lam, x, rhist
