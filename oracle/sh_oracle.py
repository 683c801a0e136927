"""oracle/sh_oracle.py -- TEST INFRASTRUCTURE ONLY.

Independent float64 construction of the real spherical-harmonics polynomials the reference
tabulates (shencoder/src/shencoder.cu:50-120) and of their partial derivatives (:131-349), via
numpy polynomial algebra:  Y_l^{+-m} = K_lm * (d^m P_l / dz^m)(z) * {Re, Im}(x + i y)^m.
The first 9 outputs are also written out explicitly from the reference table as a pinned check
(tests/test_oracle_selfcheck.py)."""
from math import factorial, pi, sqrt

import numpy as np
from numpy.polynomial import legendre as Leg
from numpy.polynomial import polynomial as Pol


def _K(l, m):
    k = sqrt((2 * l + 1) / (4 * pi) * factorial(l - m) / factorial(l + m))
    return k * (sqrt(2) * (-1) ** m if m > 0 else 1.0)


def _Q(l, m):
    """power-series coefficients of d^m/dz^m P_l(z)."""
    c = Leg.leg2poly([0] * l + [1])
    return Pol.polyder(c, m) if m > 0 else c


def sh_encode(inputs, degree, calc_grad=False):
    """inputs [B,3] -> outputs [B, degree^2] (+ dy_dx [B, 3*degree^2], blocks dx|dy|dz)."""
    v = np.asarray(inputs, np.float64)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    B = len(v)
    C2 = degree * degree
    cz = (x + 1j * y)
    out = np.zeros((B, C2))
    g = np.zeros((B, 3, C2)) if calc_grad else None
    for l in range(degree):
        for m in range(l + 1):
            q = Pol.polyval(z, _Q(l, m)) * _K(l, m)
            qz = Pol.polyval(z, Pol.polyder(_Q(l, m))) * _K(l, m) if l - m >= 1 else np.zeros(B)
            pw = cz ** m
            pw1 = m * cz ** (m - 1) if m > 0 else np.zeros(B, complex)
            ip, im = l * l + l + m, l * l + l - m
            out[:, ip] = q * pw.real
            if m > 0:
                out[:, im] = q * pw.imag
            if calc_grad:
                # d/dx (x+iy)^m = m (x+iy)^(m-1); d/dy = i m (x+iy)^(m-1)
                g[:, 0, ip] = q * pw1.real
                g[:, 1, ip] = q * (1j * pw1).real
                g[:, 2, ip] = qz * pw.real
                if m > 0:
                    g[:, 0, im] = q * pw1.imag
                    g[:, 1, im] = q * (1j * pw1).imag
                    g[:, 2, im] = qz * pw.imag
    return out, (g.reshape(B, 3 * C2) if calc_grad else None)
