"""oracle/sh_oracle.py -- TEST INFRASTRUCTURE ONLY.

Two independent statements of what encoder/shencoder/src/shencoder.cu:45-123 computes:
 (1) `sh_closed_form`: the nine degree<=3 polynomials written from the mathematical forms in the reference's
     own comments (e.g. ":52 // -sqrt(3)*y/(2*sqrt(pi))"), valid for arbitrary (non-unit) xyz;
 (2) `sh_scipy`: all 64 coefficients on UNIT vectors from scipy's complex spherical harmonics,
     real form sqrt(2) Re/Im Y_l^|m| (Condon-Shortley phase included), index l*l + l + m.
Derivatives are checked by central differences of (2)/(1) in float64."""
import numpy as np


def sh_closed_form(p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    pi = np.pi
    c1, c2 = np.sqrt(3 / (4 * pi)), np.sqrt(15 / (4 * pi))
    return np.stack([np.full_like(x, 0.5 / np.sqrt(pi)), -c1 * y, c1 * z, -c1 * x, c2 * x * y, -c2 * y * z,
                     np.sqrt(5 / (16 * pi)) * (3 * z * z - 1), -c2 * x * z, np.sqrt(15 / (16 * pi)) * (x * x - y * y)], -1)


def sh_scipy(p, degree):
    from scipy import special
    p = p / np.linalg.norm(p, axis=1, keepdims=True)
    theta, phi = np.arccos(np.clip(p[:, 2], -1, 1)), np.arctan2(p[:, 1], p[:, 0])
    out = np.zeros((p.shape[0], degree * degree))
    for l in range(degree):
        for m in range(0, l + 1):
            if hasattr(special, "sph_harm_y"):
                Y = special.sph_harm_y(l, m, theta, phi)
            else:
                Y = special.sph_harm(m, l, phi, theta)
            if m == 0:
                out[:, l * l + l] = Y.real
            else:
                out[:, l * l + l + m] = np.sqrt(2) * Y.real
                out[:, l * l + l - m] = np.sqrt(2) * Y.imag
    return out
