"""CPU study behind DESIGN.md section 8: PCG iteration counts on the cfg4 normal equations for (a) the preconditioner the
device uses (3x3 block Jacobi + rigid-body-mode coarse space over G contiguous aggregates), (b) the same coarse space with
exact aggregate-local solves (additive Schwarz), also rounded to fp32.  Uses the oracle's problem set-up (test infrastructure).
    python tools/precond_study.py            # a few minutes on one core
"""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, scipy.linalg as sla
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolbox_b200 import synth
from oracle import posegraph as PG

g = synth.make_pose_graph(0, 10000, 40000, sigma_xy=0.03, sigma_th=0.01)
U = np.stack([PG.sqrt_information(c) for c in g["cov"]])
x = np.array(g["init"], float)
pb = PG.Problem(x, g["edge_a"], g["edge_b"], g["z"], U, 0, "none", 0.7)
N = len(x)
free = pb.free
print("free nodes", len(free))

def system(xx, radius):
    r = pb.residuals(xx)
    J = pb.jacobian(xx)
    scale = 1.0 / (1.0 + np.sqrt(np.asarray(J.multiply(J).sum(axis=0)).reshape(-1)))
    J = J @ sp.diags(scale)
    diag = np.asarray(J.multiply(J).sum(axis=0)).reshape(-1)
    diag = np.minimum(np.maximum(diag, 1e-6), 1e32)
    H = (J.T @ J + sp.diags(diag / radius)).tocsr()
    return H, J.T @ r, scale

def pcg(H, b, Minv, tol=1e-10, maxit=20000):
    xk = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; bb = b @ b
    for it in range(1, maxit + 1):
        q = H @ p; a = rz / (p @ q); xk += a * p; r -= a * q
        if r @ r <= tol * tol * bb: return xk, it
        z = Minv(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return xk, maxit

def make_precond(H, xx, scale, G, mode, f32=False, sub=1):
    n = H.shape[0]; nn = n // 3
    per = (nn + G - 1) // G
    starts = list(range(0, nn, per)) + [nn]
    Gu = len(starts) - 1
    # coarse space
    rows, cols, vals = [], [], []
    fx = xx[free]
    for a in range(Gu):
        lo, hi = starts[a], starts[a + 1]
        cx, cy = fx[lo:hi, 0].mean(), fx[lo:hi, 1].mean()
        for i in range(lo, hi):
            sx, sy, st = 1 / scale[3 * i], 1 / scale[3 * i + 1], 1 / scale[3 * i + 2]
            rows += [3 * i, 3 * i + 1, 3 * i, 3 * i + 1, 3 * i + 2]
            cols += [3 * a, 3 * a + 1, 3 * a + 2, 3 * a + 2, 3 * a + 2]
            vals += [sx, sy, -(fx[i, 1] - cy) * sx, (fx[i, 0] - cx) * sy, st]
    P = sp.csr_matrix((vals, (rows, cols)), shape=(n, 3 * Gu))
    Ac = (P.T @ H @ P).toarray()
    Aci = np.linalg.inv(Ac)
    Hc = H.tocsc()
    if mode == "jacobi":
        blocks = [(3 * i, 3 * i + 3) for i in range(nn)]
    else:
        blocks = []
        for a in range(Gu):
            lo, hi = starts[a], starts[a + 1]
            m = (hi - lo + sub - 1) // sub
            for s in range(lo, hi, m): blocks.append((3 * s, 3 * min(s + m, hi)))
    invs = []
    for lo, hi in blocks:
        B = H[lo:hi, lo:hi].toarray()
        Bi = np.linalg.inv(B)
        Bi = 0.5 * (Bi + Bi.T)
        if f32: Bi = Bi.astype(np.float32).astype(np.float64)
        invs.append(Bi)
    def Minv(r):
        z = np.empty_like(r)
        for (lo, hi), Bi in zip(blocks, invs): z[lo:hi] = Bi @ r[lo:hi]
        return z + P @ (Aci @ (P.T @ r))
    return Minv, max(np.linalg.cond(np.linalg.inv(b)) for b in invs[:3])

xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
print("oracle trace", so.trace)
for label, xx, radius in (("init r=1e4", x, 1e4), ("init r=3e4", x, 3e4), ("final r=1e7", xo, 1e7), ("final r=1e10", xo, 1e10)):
    H, b, scale = system(xx, radius)
    for mode, f32, sub in (("jacobi", False, 1), ("schwarz", False, 1), ("schwarz", True, 1), ("schwarz", False, 2), ("schwarz", False, 4)):
        t = time.time()
        Minv, cnd = make_precond(H, xx, scale, 148, mode, f32, sub)
        y, its = pcg(H, b, Minv)
        print(f"{label:14s} {mode:8s} f32={f32!s:5s} sub={sub} its={its:5d} cond(block)~{cnd:.2e} ({time.time()-t:.1f}s)", flush=True)

H, b, scale = system(x, 3e4)
for G in (148, 296, 592, 1184):
    for mode in ("jacobi", "schwarz"):
        Minv, _ = make_precond(H, x, scale, G, mode)
        print(f"G={G:5d} {mode:8s} its={pcg(H, b, Minv)[1]:5d}", flush=True)
