"""Synthetic inputs of the BASELINE.json configs; seed fixed."""
import numpy as np

SEED = 20260


def two_peak_lnpi(n=1001, noise=1e-3, scale=1.0, seed=SEED):
    """config 2/3/4 ln(PI): two Gaussians in log space + Gaussian noise."""
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.float64)
    lnpi = np.logaddexp(-(i - 150.0 * scale) ** 2 / (2 * (40.0 * scale) ** 2),
                        -(i - 600.0 * scale) ** 2 / (2 * (60.0 * scale) ** 2) - 1.0)
    return lnpi + noise * rng.standard_normal(n)


def one_comp_moments(n=1001, max_order=2):
    """config 2 moment tensor (1, mo+1, 1, mo+1, mo+1, n): mom[0,j,0,m,p] = N^(j+m) g_p(N)."""
    i = np.arange(n, dtype=np.float64)
    u = -2.0 * i - 0.002 * i * i
    g = [np.ones(n), u, u * u + 0.5 * i, u ** 3 + 1.5 * i * u]
    mo = max_order
    mom = np.zeros((1, mo + 1, 1, mo + 1, mo + 1, n))
    for j in range(mo + 1):
        for m in range(mo + 1):
            for p in range(mo + 1):
                mom[0, j, 0, m, p] = i ** (j + m) * g[p]
    return mom


def two_comp_moments(n=1001):
    """config 3 moment tensor (2,3,2,3,3,n) from means + covariances linear in N (max_order 2)."""
    i = np.arange(n, dtype=np.float64)
    x = 0.3 + 0.2 * np.sin(i / 50.0)
    mean = {"N1": x * i, "N2": i - x * i, "U": -2.0 * i + 0.01 * i ** 1.5}
    cov = {("N1", "N1"): 0.1 * i, ("N2", "N2"): 0.1 * i, ("N1", "N2"): -0.05 * i, ("U", "U"): 0.5 * i,
           ("N1", "U"): -0.2 * i, ("N2", "U"): -0.1 * i}

    def c(a, b):
        return cov[(a, b)] if (a, b) in cov else cov[(b, a)]

    mom = np.zeros((2, 3, 2, 3, 3, n))
    sp = ["N1", "N2"]
    for a in range(2):
        for j in range(3):
            for b in range(2):
                for m in range(3):
                    for p in range(3):
                        facs = [sp[a]] * j + [sp[b]] * m + ["U"] * p
                        if len(facs) == 0:
                            mom[a, j, b, m, p] = 1.0
                        elif len(facs) == 1:
                            mom[a, j, b, m, p] = mean[facs[0]]
                        elif len(facs) == 2:
                            mom[a, j, b, m, p] = mean[facs[0]] * mean[facs[1]] + c(facs[0], facs[1])
    return mom


def joint_2d(n1=512, n2=512, cut=640):
    """config 5: two anisotropic 2-D Gaussians in log space; N1+N2 > cut set to -inf, matching bounds."""
    a = np.arange(n1, dtype=np.float64)[:, None]
    b = np.arange(n2, dtype=np.float64)[None, :]
    g1 = -((a - 80.0) ** 2 / (2 * 30.0 ** 2) + (b - 60.0) ** 2 / (2 * 20.0 ** 2))
    g2 = -((a - 300.0) ** 2 / (2 * 50.0 ** 2) + (b - 250.0) ** 2 / (2 * 70.0 ** 2)) - 2.0
    lnpi = np.logaddexp(g1, g2)
    mask = (a + b) > cut
    lnpi[mask] = -np.inf
    bounds = np.zeros((n1, 2), dtype=np.int32)
    for i in range(n1):
        hi = int(min(n2, max(0, cut - i + 1)))
        bounds[i] = (0, hi)
    return lnpi, bounds


def n1_two_comp_moments(n=201, max_order=3):
    """Moment tensor (2, mo+1, 2, mo+1, mo+1, n) for an N_1 order parameter: N_1 = n is sharp in every bin and
    (N_2, U) follow a discrete two-variable distribution (8x8 Gauss-Hermite nodes of a correlated Gaussian), so every
    stored moment <N_i^j N_k^m U^p> is an exact moment of one distribution and all index symmetries hold."""
    i = np.arange(n, dtype=np.float64)
    z, w = np.polynomial.hermite_e.hermegauss(8)
    w = w / w.sum()
    z1, z2 = np.meshgrid(z, z, indexing="ij")
    ww = (w[:, None] * w[None, :]).ravel()
    z1, z2 = z1.ravel(), z2.ravel()
    rho = -0.4
    n2 = (0.4 * i + 5.0)[:, None] + np.sqrt(0.1 * i + 1.0)[:, None] * z1[None, :]
    u = (-2.0 * i + 0.01 * i ** 1.5)[:, None] + np.sqrt(0.5 * i + 1.0)[:, None] * (rho * z1 + np.sqrt(1 - rho * rho) * z2)[None, :]
    x = [np.broadcast_to(i[:, None], n2.shape), n2]
    mo = max_order
    mom = np.zeros((2, mo + 1, 2, mo + 1, mo + 1, n))
    for a in range(2):
        for j in range(mo + 1):
            for b in range(2):
                for m in range(mo + 1):
                    for p in range(mo + 1):
                        mom[a, j, b, m, p] = np.sum(ww[None, :] * x[a] ** j * x[b] ** m * u ** p, axis=1)
    return mom
