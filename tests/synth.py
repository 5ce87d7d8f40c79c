"""Deterministic synthetic instances for the dual-evaluation kernel.

Counter-based generator: u01(k, j) = top 53 bits of mix64((0x5EED0000 + k) * GOLDEN + j) * 2^-53,
mix64 = the splitmix64 finaliser.  The CUDA side (nlopt_b200/csrc/synth.cuh) evaluates the very
same integer hash and the same un-fused floating-point expressions, so host and device arrays
are bit-identical.
"""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
SEED0 = 0x5EED0000


def mix64(z):
    with np.errstate(over="ignore"):
        z = z + GOLDEN
        z = (z ^ (z >> np.uint64(30))) * M1
        z = (z ^ (z >> np.uint64(27))) * M2
        return z ^ (z >> np.uint64(31))


def u01(k, n, seed=SEED0, j0=0):
    with np.errstate(over="ignore"):
        base = np.uint64(seed + k) * GOLDEN
        j = np.arange(j0, j0 + n, dtype=np.uint64)
        z = mix64(base + j)
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def kernel_instance(n, m, seed=SEED0, special_lanes=True):
    """Arrays + scalars of one dual evaluation: dict with x, lb, ub, sigma, grad_f, grad_c (m x n),
    f0, rho, c0, rhoc, y."""
    cls = u01(99, n, seed)
    lb = np.full(n, -2.0)
    ub = np.full(n, 2.0)
    sigma = (0.05 + 0.95 * u01(0, n, seed)) * 2.0
    x = lb + (0.25 + 0.5 * u01(1, n, seed)) * (ub - lb)
    if special_lanes:
        fixed = cls < 0.001
        free_inf = (cls >= 0.001) & (cls < 0.002)
        lb[fixed] = x[fixed]
        ub[fixed] = x[fixed]
        sigma[fixed] = 0.0
        lb[free_inf] = -np.inf
        ub[free_inf] = np.inf
    grad_f = (2.0 * u01(2, n, seed) - 1.0) * 10.0
    grad_c = np.empty((m, n))
    for i in range(m):
        grad_c[i] = 2.0 * u01(3 + i, n, seed) - 1.0
    i = np.arange(m, dtype=np.float64)
    return dict(n=n, m=m, x=x, lb=lb, ub=ub, sigma=sigma, grad_f=grad_f, grad_c=grad_c,
                f0=1.0, rho=1.0, c0=-0.1 * (i + 1.0), rhoc=1.0 + 0.1 * i, y=0.5 * (i + 1.0))
