"""NumPy model of the *sweep structure* the CUDA UVd kernels use (psgd_tf_b200/csrc/uvd.cu) -- test infrastructure.

The reference (psgd.py:554-627) evaluates every r-sized quantity with its own pass over the N x r arrays.  The CUDA
path gets all of them from ONE Gram table  G = [U V dh w]^T [U V dh w]  (dh = d*h, w = v/d) by re-association, so that
the rank-2 update of U (or V) can ride in the same sweep that forms a, b and nablaD, and -- in the fused
update+apply call -- the Gram quantities of the *updated* factors needed by the apply are accumulated in that sweep
(those that involve g) or expanded once more through the table (U'^T U').
This file restates that algebra on the host (float32 per-row arithmetic, float64 for the r-sized algebra, as on the
device) so that it can be checked against the oracle without a GPU (tests/test_uvd_pipeline_model.py).
"""
import numpy as np

F = np.float32
TINY = F(2.0 ** -126)


def gram_table(U, V, d, h, v):
    """Sweep 1: upper triangle of Z^T Z, Z = [U V dh w] (float32 products, float64 sums like the two-stage reduction)."""
    dh = (d * h).astype(F)
    w = (v / d).astype(F)
    Z = np.concatenate([U, V, dh, w], axis=1).astype(np.float64)
    return Z.T @ Z


def small1(G, r, update_U, step, tiny=TINY):
    """r x r algebra after sweep 1 (uvd_small1_kernel): p, t, s1, s2 and -- new -- the rank-2 coefficients."""
    W = 2 * r
    UtU, VtV, UtV = G[:r, :r], G[r:W, r:W], G[:r, r:W]
    Utdh, Vtdh, Utw, Vtw = G[:r, W], G[r:W, W], G[:r, W + 1], G[r:W, W + 1]
    dhdh, dhw, ww = G[W, W], G[W, W + 1], G[W + 1, W + 1]
    IpVtU = np.eye(r) + UtV.T                                                   # psgd.py:574-575
    p = Vtdh                                                                    # Qh = dh + U p            :569
    t = Utdh + UtU @ p                                                          # U^T Qh                   :570
    s1 = np.linalg.solve(IpVtU.T, Utw)                                          # invQtv = w - V s1        :577
    s2 = np.linalg.solve(IpVtU, Vtw - VtV @ s1)                                 # invPv = (b - U s2)/d     :578
    # a = dh + U p, b = w - V s1: every reduction over a, b expands through G
    aa = dhdh + 2 * p @ Utdh + p @ UtU @ p
    bb = ww - 2 * s1 @ Vtw + s1 @ VtV @ s1
    ab = dhw - s1 @ Vtdh + p @ Utw - p @ UtV @ s1
    if update_U:
        atX = Vtdh + p @ UtV                                                    # a^T V                    :589
        btX = Vtw - s1 @ VtV                                                    # b^T V                    :591
        XtX = VtV
    else:
        atX = t                                                                 # a^T U                    :603
        btX = Utw - UtV @ s1                                                    # b^T U                    :604
        XtX = UtU
    qaa, qbb, qab = atX @ XtX @ atX, btX @ XtX @ btX, atX @ XtX @ btX
    norm = F(np.sqrt(np.abs(F(aa * qaa + bb * qbb - 2 * ab * qab))))            # :594-596 / :608-610
    mu = F(step) / (norm + tiny)
    if update_U:
        c1, c2 = (mu * (atX @ IpVtU).astype(F)).astype(F), (mu * (btX @ IpVtU).astype(F)).astype(F)   # :600-601
    else:
        c1, c2 = atX.astype(F), btX.astype(F)
    # kept for the fused update+apply: U'^T U' of the updated U is itself an expansion (no per-row accumulators needed)
    Uta, Utb = t, Utw - UtV @ s1
    return dict(p=p.astype(F), t=t.astype(F), s1=s1.astype(F), s2=s2.astype(F), c1=c1, c2=c2, mu=mu,
                UtU=UtU, Uta=Uta, Utb=Utb, aa=aa, bb=bb, ab=ab)


def updated_UtU(k, update_U):
    """U'^T U' after the rank-2 step U' = U - a c1^T + b c2^T (c1, c2 pre-scaled by mu); U' = U on the V branch."""
    if not update_U:
        return k["UtU"]
    c1, c2 = k["c1"].astype(np.float64), k["c2"].astype(np.float64)
    o = np.outer
    return (k["UtU"] - o(k["Uta"], c1) - o(c1, k["Uta"]) + o(k["Utb"], c2) + o(c2, k["Utb"])
            + k["aa"] * o(c1, c1) - k["ab"] * (o(c1, c2) + o(c2, c1)) + k["bb"] * o(c2, c2))


def fused_map(U, V, d, h, v, k, update_U):
    """Sweep 2 (fused): per row a, b, nablaD and the rank-2 update of U (or V); returns U', V', nablaD."""
    dh = d * h
    a = dh + U @ k["p"][:, None]
    Ph = d * (a + V @ k["t"][:, None])
    w = v / d
    b = w - V @ k["s1"][:, None]
    invPv = (b - U @ k["s2"][:, None]) / d
    nd = (Ph * h - v * invPv).astype(F)
    if update_U:
        Un = (U - (a * k["c1"][None, :] - b * k["c2"][None, :])).astype(F)
        Vn = V
    else:
        sa = a + V @ k["c1"][:, None]
        sb = b + V @ k["c2"][:, None]
        Vn = (V - k["mu"] * (sa * k["c1"][None, :] - sb * k["c2"][None, :])).astype(F)
        Un = U
    return Un, Vn, nd


def update(U, V, d, v, h, step, update_U, tiny=TINY):
    """update_precond_UVd_math_ (no balance) in the two-sweep form: Gram sweep, fused map sweep, d update."""
    r = U.shape[1]
    k = small1(gram_table(U, V, d, h, v), r, update_U, step, tiny)
    Un, Vn, nd = fused_map(U, V, d, h, v, k, update_U)
    mu_d = F(step) / (np.max(np.abs(nd)) + tiny)
    dn = (d - (mu_d * d) * nd).astype(F)
    return Un, Vn, dn


def update_apply(U, V, d, v, h, g, step, update_U, tiny=TINY):
    """Fused update + apply: the apply's Gram quantities of the UPDATED factors are accumulated in the map sweep, with
    d' g = d g - mu_d (d nablaD g) split into two columns because mu_d is only known after the sweep."""
    r = U.shape[1]
    k = small1(gram_table(U, V, d, h, v), r, update_U, step, tiny)
    Un, Vn, nd = fused_map(U, V, d, h, v, k, update_U)
    x0 = (d * g).astype(F)
    x1 = (x0 * nd).astype(F)
    U64, V64 = Un.astype(np.float64), Vn.astype(np.float64)
    UtU = updated_UtU(k, update_U)
    assert np.allclose(UtU, U64.T @ U64, rtol=1e-5, atol=1e-7 * np.abs(UtU).max())
    UtX0, UtX1, VtX0, VtX1 = U64.T @ x0, U64.T @ x1, V64.T @ x0, V64.T @ x1
    mu_d = F(step) / (np.max(np.abs(nd)) + tiny)
    p = (VtX0 - np.float64(mu_d) * VtX1)[:, 0]                                  # V'^T (d' g)              :625
    t = (UtX0 - np.float64(mu_d) * UtX1)[:, 0] + UtU @ p                        # U'^T (d' g + U' p)       :626
    p, t = p.astype(F), t.astype(F)
    dn = (d - (mu_d * d) * nd).astype(F)
    dg = dn * g
    y = dg + Un @ p[:, None]
    pre = (dn * (y + Vn @ t[:, None])).astype(F)
    return Un, Vn, dn, pre
