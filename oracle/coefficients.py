"""Reconstruction coefficient tables (TEST INFRASTRUCTURE -- part of the CPU oracle).

Restates src/Advection/reconstruction_coefficients.jl:62-77 (`stencil_coefficients`, evaluated
there in BigFloat; here in exact rationals), weno_interpolants.jl:76-103 (optimal weights C*) and
:169-192 (uniform smoothness coefficients).  The last coefficient of every stencil is
`1 - sum(others)` evaluated *in FT* exactly as the reference does.
"""
from fractions import Fraction as Fr

import numpy as np

MAXBUF = 6


def _num_prod(i, m, l, r, order):
    p = Fr(1)
    for q in range(order + 1):
        if q != m and q != l:
            p *= Fr(i) - Fr(i - (r - q + 1))
    return p


def stencil_coefficients_exact(r, order, i=50):
    """reconstruction_coefficients.jl:62-77 on the uniform grid xr = xi = 1:100."""
    coeffs = [Fr(0)] * order
    for j in range(order):
        for m in range(j + 1, order + 1):
            num = sum(_num_prod(i, m, l, r, order) for l in range(order + 1) if l != m)
            den = Fr(1)
            for l in range(order + 1):
                if l != m:
                    den *= Fr(i - (r - m + 1)) - Fr(i - (r - l + 1))
            coeffs[j] += num / den * (Fr(i - (r - j)) - Fr(i - (r - j + 1)))
    return coeffs


def _to_ft(fracs, ft):
    """`coeffs = FT.(coeffs)[1:end-1]; (coeffs..., 1 - sum(coeffs))` in FT arithmetic.

    float(Fraction) is correctly rounded to Float64; the further rounding to Float32 could differ from a
    direct rounding only on an exact tie, which none of these rationals produces."""
    ft = np.dtype(ft).type
    head = [ft(float(f)) for f in fracs[:-1]]
    s = ft(0)
    for h in head:
        s = ft(s + h)
    return head + [ft(ft(1) - s)]


def weno_coeff_table(ft):
    """coeff_p(WENO{buffer}, stencil) -> array [MAXBUF+1, MAXBUF, MAXBUF] (weno_interpolants.jl:117-118)."""
    t = np.zeros((MAXBUF + 1, MAXBUF, MAXBUF), dtype=ft)
    for n in range(2, MAXBUF + 1):
        for s in range(n):
            t[n, s, :n] = _to_ft(stencil_coefficients_exact(s, n), ft)
    return t


def centered_coeff_table(ft):
    """Centered{n} coefficients in *application order* (reconstruction_coefficients.jl:146-149:
    C = coeff[order - idx + 1]) -> array [MAXBUF+1, 2*MAXBUF]."""
    t = np.zeros((MAXBUF + 1, 2 * MAXBUF), dtype=ft)
    for n in range(1, MAXBUF + 1):
        c = _to_ft(stencil_coefficients_exact(n - 1, 2 * n), ft)
        t[n, : 2 * n] = c[::-1]
    return t


CSTAR = {
    2: (Fr(2, 3), Fr(1, 3)),
    3: (Fr(3, 10), Fr(3, 5), Fr(1, 10)),
    4: (Fr(4, 35), Fr(18, 35), Fr(12, 35), Fr(1, 35)),
    5: (Fr(5, 126), Fr(20, 63), Fr(10, 21), Fr(10, 63), Fr(1, 126)),
    6: (Fr(1, 77), Fr(25, 154), Fr(100, 231), Fr(25, 77), Fr(5, 77), Fr(1, 462)),
}


def cstar_table(ft):
    t = np.zeros((MAXBUF + 1, MAXBUF), dtype=ft)
    for n, vals in CSTAR.items():
        t[n, :n] = [np.dtype(ft).type(float(v)) for v in vals]
    return t


# weno_interpolants.jl:169-192 (decimal literals converted with FT.(...))
SMOOTHNESS = {
    (2, 0): (1, -2, 1),
    (2, 1): (1, -2, 1),
    (3, 0): (10, -31, 11, 25, -19, 4),
    (3, 1): (4, -13, 5, 13, -13, 4),
    (3, 2): (4, -19, 11, 25, -31, 10),
    (4, 0): (2.107, -9.402, 7.042, -1.854, 11.003, -17.246, 4.642, 7.043, -3.882, 0.547),
    (4, 1): (0.547, -2.522, 1.922, -0.494, 3.443, -5.966, 1.602, 2.843, -1.642, 0.267),
    (4, 2): (0.267, -1.642, 1.602, -0.494, 2.843, -5.966, 1.922, 3.443, -2.522, 0.547),
    (4, 3): (0.547, -3.882, 4.642, -1.854, 7.043, -17.246, 7.042, 11.003, -9.402, 2.107),
    (5, 0): (1.07918, -6.49501, 7.58823, -4.11487, 0.86329, 10.20563, -24.62076, 13.58458, -2.88007, 15.21393,
             -17.04396, 3.64863, 4.82963, -2.08501, 0.22658),
    (5, 1): (0.22658, -1.40251, 1.65153, -0.88297, 0.18079, 2.42723, -6.11976, 3.37018, -0.70237, 4.06293,
             -4.64976, 0.99213, 1.38563, -0.60871, 0.06908),
    (5, 2): (0.06908, -0.51001, 0.67923, -0.38947, 0.08209, 1.04963, -2.99076, 1.79098, -0.38947, 2.31153,
             -2.99076, 0.67923, 1.04963, -0.51001, 0.06908),
    (5, 3): (0.06908, -0.60871, 0.99213, -0.70237, 0.18079, 1.38563, -4.64976, 3.37018, -0.88297, 4.06293,
             -6.11976, 1.65153, 2.42723, -1.40251, 0.22658),
    (5, 4): (0.22658, -2.08501, 3.64863, -2.88007, 0.86329, 4.82963, -17.04396, 13.58458, -4.11487, 15.21393,
             -24.62076, 7.58823, 10.20563, -6.49501, 1.07918),
    (6, 0): (0.6150211, -4.7460464, 7.6206736, -6.3394124, 2.7060170, -0.4712740, 9.4851237, -31.1771244,
             26.2901672, -11.3206788, 1.9834350, 26.0445372, -44.4003904, 19.2596472, -3.3918804, 19.0757572,
             -16.6461044, 2.9442256, 3.6480687, -1.2950184, 0.1152561),
    (6, 1): (0.1152561, -0.9117992, 1.4742480, -1.2183636, 0.5134574, -0.0880548, 1.9365967, -6.5224244,
             5.5053752, -2.3510468, 0.4067018, 5.6662212, -9.7838784, 4.2405032, -0.7408908, 4.3093692,
             -3.7913324, 0.6694608, 0.8449957, -0.3015728, 0.0271779),
    (6, 2): (0.0271779, -0.2380800, 0.4086352, -0.3462252, 0.1458762, -0.0245620, 0.5653317, -2.0427884,
             1.7905032, -0.7727988, 0.1325006, 1.9510972, -3.5817664, 1.5929912, -0.2792660, 1.7195652,
             -1.5880404, 0.2863984, 0.3824847, -0.1429976, 0.0139633),
    (6, 3): (0.0139633, -0.1429976, 0.2863984, -0.2792660, 0.1325006, -0.0245620, 0.3824847, -1.5880404,
             1.5929912, -0.7727988, 0.1458762, 1.7195652, -3.5817664, 1.7905032, -0.3462252, 1.9510972,
             -2.0427884, 0.4086352, 0.5653317, -0.2380800, 0.0271779),
    (6, 4): (0.0271779, -0.3015728, 0.6694608, -0.7408908, 0.4067018, -0.0880548, 0.8449957, -3.7913324,
             4.2405032, -2.3510468, 0.5134574, 4.3093692, -9.7838784, 5.5053752, -1.2183636, 5.6662212,
             -6.5224244, 1.4742480, 1.9365967, -0.9117992, 0.1152561),
    (6, 5): (0.1152561, -1.2950184, 2.9442256, -3.3918804, 1.9834350, -0.4712740, 3.6480687, -16.6461044,
             19.2596472, -11.3206788, 2.7060170, 19.0757572, -44.4003904, 26.2901672, -6.3394124, 26.0445372,
             -31.1771244, 7.6206736, 9.4851237, -4.7460464, 0.6150211),
}


def smoothness_table(ft):
    t = np.zeros((MAXBUF + 1, MAXBUF, 21), dtype=ft)
    for (n, s), vals in SMOOTHNESS.items():
        t[n, s, : len(vals)] = np.asarray(vals, dtype=np.float64).astype(ft)
    return t


def weno_eps(ft):
    """`const ϵ = 1f-8` (weno_interpolants.jl:71): a Float32 literal promoted to FT."""
    return np.dtype(ft).type(np.float32(1e-8))
