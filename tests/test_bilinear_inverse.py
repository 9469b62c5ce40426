"""The inverse of the bilinear map used by the queue kernel for quadrilateral faces that are no parallelograms
(csrc/rthx_kernels.cu::lattice_cell_bilinear), restated in numpy line by line and checked against the forward map
P(s,t) = A + s E + t F + s t G of meshQuad.jl:116-136 on random convex quadrilaterals, trapezoids (a2 = 0 along one axis)
and near-parallelograms (G -> 0)."""
import numpy as np


def cross(a, b):
    return a[0] * b[1] - a[1] * b[0]


def inverse_bilinear(A, B, C, D, p):
    """Line-by-line restatement of lattice_cell_bilinear (before the floor): returns (s, t) or None."""
    E, F, G, H = B - A, D - A, A - B + C - D, p - A
    a2 = cross(E, G)
    inv_a2 = 1.0 / a2 if a2 != 0.0 else np.inf                     # CoarseDev::hw[0], computed on the host
    a1 = cross(E, F) - cross(H, G)
    a0 = H[1] * F[0] - H[0] * F[1]
    disc = a1 * a1 - 4.0 * a2 * a0
    if not disc >= 0.0:
        return None
    q = -0.5 * (a1 + np.copysign(np.sqrt(disc), a1))
    with np.errstate(all="ignore"):
        sb = a0 / q
        sa = q * inv_a2
    s = sb if (-1e-9 <= sb <= 1.0 + 1e-9) else sa
    T = F + s * G
    t = ((H - s * E) @ T) / (T @ T)
    return s, t


def forward(A, B, C, D, s, t):
    return A + s * (B - A) + t * (D - A) + s * t * (A - B + C - D)


def _check(P, rng, n=25):
    A, B, C, D = P
    worst = 0.0
    for _ in range(n):
        s, t = rng.random(2)
        r = rng.random()
        if r < 0.1:
            s = 1e-13
        elif r < 0.2:
            t = 1.0 - 1e-13
        got = inverse_bilinear(A, B, C, D, forward(A, B, C, D, s, t))
        assert got is not None
        worst = max(worst, abs(got[0] - s), abs(got[1] - t))
    return worst


def test_inverse_of_the_bilinear_map():
    rng = np.random.default_rng(5)
    worst, n_quads = 0.0, 0
    while n_quads < 3000:
        ang = np.sort(rng.random(4) * 2 * np.pi)
        r = 0.3 + rng.random(4)
        P = np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
        if not all(cross(P[(i + 1) % 4] - P[i], P[(i + 2) % 4] - P[(i + 1) % 4]) > 1e-3 for i in range(4)):
            continue
        n_quads += 1
        worst = max(worst, _check(P, rng))
    for _ in range(500):
        w = rng.random() * 0.8 + 0.1
        worst = max(worst, _check(np.array([[0, 0], [1, 0], [0.5 + w / 2, 1], [0.5 - w / 2, 1.0]]), rng))          # trapezoid
        worst = max(worst, _check(np.array([[0, 0], [0.3, 1.0], [-0.5, 1.2], [-0.4, 0.2]]) + rng.random(2), rng))   # rotated order
        worst = max(worst, _check(np.array([[0, 0], [1, 0], [1 + 1e-9 * rng.random(), 1], [0, 1.0]]), rng))         # almost a square
    assert worst < 1e-11
