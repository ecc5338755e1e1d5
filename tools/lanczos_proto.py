"""Prototype (numpy, float32 vectors / float64 tridiagonal) of the GPU eigensolver's algorithm.

Not product code: used to choose the tolerance / check cadence / max basis size before writing eigsh.cu and to
document the algorithm. Lanczos with full re-orthogonalisation (CGS2) on S = D^-1/2 W D^-1/2 with the known top
eigenvector D^1/2 1 deflated; Ritz values by bisection, Ritz vectors of T by twisted factorisation.
"""
import sys
import numpy as np

f32 = np.float32


def sturm_count(alpha, beta2, x):
    """number of eigenvalues of T strictly less than x"""
    cnt = 0
    q = 1.0
    for i in range(len(alpha)):
        q = (alpha[i] - x) - (beta2[i - 1] / q if i > 0 else 0.0)
        if q == 0.0:
            q = -1e-300
        if q < 0:
            cnt += 1
    return cnt


def bisect_kth_largest(alpha, beta, k, lo, hi, rounds=12, ways=32):
    """k-th largest eigenvalue (k=0 largest) of tridiagonal T by multi-section."""
    n = len(alpha)
    beta2 = beta * beta
    target = n - 1 - k  # index in ascending order: count(x) <= target  <=> lambda_target >= x
    for _ in range(rounds):
        xs = lo + (hi - lo) * (np.arange(1, ways + 1) / (ways + 1))
        cnts = np.array([sturm_count(alpha, beta2, x) for x in xs])
        # eigenvalue #target (ascending, 0-based) lies in interval where count goes from <=target to >target
        below = np.nonzero(cnts <= target)[0]
        i = below[-1] if len(below) else -1
        nlo = xs[i] if i >= 0 else lo
        nhi = xs[i + 1] if i + 1 < ways else hi
        lo, hi = nlo, nhi
    return 0.5 * (lo + hi)


def tri_eigvec(alpha, beta, theta):
    n = len(alpha)
    if n == 1:
        return np.ones(1)
    dp = np.zeros(n); dm = np.zeros(n); l = np.zeros(n - 1); u = np.zeros(n - 1)
    tiny = 1e-300
    dp[0] = alpha[0] - theta
    for i in range(n - 1):
        if dp[i] == 0: dp[i] = tiny
        l[i] = beta[i] / dp[i]
        dp[i + 1] = (alpha[i + 1] - theta) - l[i] * beta[i]
    dm[n - 1] = alpha[n - 1] - theta
    for i in range(n - 2, -1, -1):
        if dm[i + 1] == 0: dm[i + 1] = tiny
        u[i] = beta[i] / dm[i + 1]
        dm[i] = (alpha[i] - theta) - u[i] * beta[i]
    gamma = dp + dm - (alpha - theta)
    r = int(np.argmin(np.abs(gamma)))
    z = np.zeros(n)
    z[r] = 1.0
    for i in range(r - 1, -1, -1):
        z[i] = -l[i] * z[i + 1]
    for i in range(r, n - 1):
        z[i + 1] = -u[i] * z[i]
    return z / np.linalg.norm(z)


def lanczos_topk(W, K, tol=1e-6, mmax=300, check_every=4, seed=0, verbose=False):
    """W float32 (N,N) symmetric nonneg. Returns (lambda (K,), V (K,N) D-orthonormal, n_matvec)."""
    N = W.shape[0]
    deg = W @ np.ones(N, f32)
    deg[deg < 1e-12] = 1.0
    dinv = (1.0 / np.sqrt(deg.astype(np.float64))).astype(f32)
    sumd = float(deg.astype(np.float64).sum())
    u0 = (np.sqrt(deg.astype(np.float64)) / np.sqrt(sumd)).astype(f32)
    Kw = K - 1
    rng = np.random.default_rng(seed)
    V = np.zeros((mmax + 1, N), f32)
    v = rng.uniform(-1, 1, N).astype(f32)
    v -= f32(v @ u0) * u0
    v -= f32(v @ u0) * u0
    v /= f32(np.linalg.norm(v))
    V[0] = v
    alpha = []; beta = []
    theta = None; S = None
    for j in range(mmax):
        x = dinv * V[j]
        w = dinv * (W @ x)
        # CGS2 against u0, V[0..j]
        c0 = f32(w @ u0); c = V[:j + 1] @ w
        w = w - c0 * u0 - V[:j + 1].T @ c
        a = float(c[j])
        c0 = f32(w @ u0); c = V[:j + 1] @ w
        w = w - c0 * u0 - V[:j + 1].T @ c
        a += float(c[j])
        b = float(np.linalg.norm(w.astype(np.float64)))
        alpha.append(a); beta.append(b)
        V[j + 1] = (w / f32(b)).astype(f32)
        n = j + 1
        if n >= max(Kw + 2, 8) and ((n - Kw) % check_every == 0 or n == mmax):
            al = np.array(alpha); be = np.array(beta[:-1])
            gl = min(al - np.abs(np.r_[0, be]) - np.abs(np.r_[be, 0])); gh = max(al + np.abs(np.r_[0, be]) + np.abs(np.r_[be, 0]))
            theta = np.array([bisect_kth_largest(al, be, k, gl, gh) for k in range(Kw)])
            S = np.stack([tri_eigvec(al, be, t) for t in theta])  # (Kw, n)
            resid = np.abs(beta[-1] * S[:, -1])
            if verbose:
                print(n, theta, resid)
            if np.all(resid <= tol):
                break
    U = (S.astype(f32) @ V[:n])  # (Kw, N)
    U /= np.linalg.norm(U, axis=1, keepdims=True)
    vecs = np.concatenate([(np.ones(N, f32) / f32(np.sqrt(sumd)))[None], U * dinv[None]], 0)
    lam = np.concatenate([[0.0], 1.0 - theta])
    return lam.astype(f32), vecs, n


if __name__ == "__main__":
    import importlib, torch, time
    sys.path.insert(0, "/root/repo")
    from oracle import eigs_ref, dino_vit
    synth = importlib.import_module("deep-spectral-segmentation_b200.synth")
    torch.set_grad_enabled(False)
    cases = []
    for seed in range(3):
        cases.append((f"struct900_s{seed}", synth.structured_features(900, 384, 6, seed), 5))
    cases.append(("struct900_r4", synth.structured_features(900, 384, 4, 7), 5))
    m = dino_vit.build("dino_vits16", 0)
    for seed in range(3):
        img = synth.blobs_image(480, 480, seed)
        cases.append((f"blobs480_s{seed}", m.forward_k(dino_vit.preprocess_u8(img, 16))[0], 5))
    g = torch.Generator().manual_seed(0)
    noise = (torch.rand(480, 480, 3, generator=g) * 255).to(torch.uint8)
    cases.append(("noise480", m.forward_k(dino_vit.preprocess_u8(noise, 16))[0], 5))
    cases.append(("struct900_K15", synth.structured_features(900, 384, 6, 3), 15))
    for name, feats, K in cases:
        W, D = eigs_ref.affinity_matrices(feats)
        t = time.time(); lam, vecs, n = lanczos_topk(W, K); dt = time.time() - t
        ev64, vec64 = eigs_ref.eigh_f64(feats, K)
        evr, vecr = eigs_ref.extract_eig(feats, K, rng_seed=0)
        def relerr(a, b):
            out = []
            for k in range(K):
                s = np.sign(np.dot(a[k], b[k]))
                out.append(np.linalg.norm(a[k] - s * b[k]) / np.linalg.norm(b[k]))
            return np.array(out)
        print(f"{name}: n_matvec={n} lam={np.round(ev64[:6],4)}")
        print("   proto vs f64 :", relerr(vecs, vec64).max(), " ref(scipy f32) vs f64:", relerr(vecr.numpy(), vec64).max(),
              " proto vs ref:", relerr(vecs, vecr.numpy()).max(), " |dlam|:", np.abs(lam - evr.numpy()).max())
