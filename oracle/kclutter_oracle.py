"""CPU oracle of the correlated K-distributed clutter field — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy / scipy (float64) restatement of /root/reference/source_code/K_distributed_SeaClutter_Simulation_20210919.py:
    mnlt                                     :83-91   Gaussian sample -> Gamma(v, 1) quantile (memoryless non-linear transform)
    hermite_polynomials / coeff_acf_polyn    :93-139  three polynomial coefficients from the samples themselves
    solve_acf_polyn                          :141-164 per element: np.roots(quadratic)[0]
    generate_correlated_Gaussian_via_expdecay:270-297 complex speckle field, power spectrum |f|^-0.6
    generate_K_distributed_noise             :469-526 the pipeline
with the two white-noise fields the reference draws with np.random.normal (:483 and :287, in this order) as INPUTS, so that
a test can replay the reference's own draws.  `np.roots(c)[0]` of a quadratic is restated in closed form (the root of larger
magnitude when the roots are real, the one with positive imaginary part otherwise; checked against np.roots in
tests/test_oracle_golden.py).  Pinned by tests/golden/make_kclutter_golden.py -> tests/golden/kclutter.npz."""
import math

import numpy as np
import scipy.special as ss
from numpy.fft import fft2, ifft2


def mnlt(x, v):                                                          # :83-91
    return ss.gammaincinv(v, 1 - ss.erfc(x / np.sqrt(2)) / 2)


def acf_coefficients(x, g):                                              # coeff_acf_polyn :121-139 + normalisation :488
    """[alpha_2, alpha_1, alpha_0] / alpha_0 with H_2 = 4x^2 - 2, H_1 = 2x, H_0 = 1 (physicists' Hermite polynomials)."""
    e = np.exp(-x ** 2) * g
    sums = [np.sum(e * (4 * x ** 2 - 2)), np.sum(e * (2 * x)), np.sum(e)]
    co = np.array([s ** 2 / (np.pi * math.factorial(n) * 2 ** n) for s, n in zip(sums, (2, 1, 0))])
    return co / co[-1]


def first_root(a, b, c):
    """np.roots([a, b, c])[0] for arrays c (a, b scalars), complex128."""
    c = np.asarray(c, dtype=np.float64)
    disc = b * b - 4 * a * c
    s = np.sqrt(np.abs(disc))
    real = (-b - np.sign(b) * s) / (2 * a) + 0j
    cplx = (-b + 1j * s) / (2 * a)
    return np.where(disc >= 0, real, cplx)


def gamma_field_acf(height, width, v):                                   # :477-484, eq. (69) of Tough & Ward
    xs = np.linspace(10, height, num=width, endpoint=True)
    ys = np.linspace(10, height, num=height, endpoint=True)
    XS, YS = np.meshgrid(xs, ys)
    return 1 + np.exp(-(XS + YS) / 10) * np.cos(np.pi * YS / 8) / v


def speckle_spectrum(M):                                                 # :276-295: sqrt of the power spectrum |f|^-0.6
    fs = M / 10.0
    f = np.linspace(0.1, fs, num=M, endpoint=True)
    Fx, Fy = np.meshgrid(f, f)
    return np.sqrt(np.sqrt(Fx ** 2 + Fy ** 2) ** (-0.6))


def k_field(white_texture, white_speckle, v=5):                          # generate_K_distributed_noise :469-526
    """Returns (amplitude |speckle * sqrt(texture)|, texture) for square fields, float64."""
    h, w = white_texture.shape
    g = mnlt(white_texture, v)
    co = acf_coefficients(white_texture, g)
    gauss_acf = first_root(co[0], co[1], co[2] - gamma_field_acf(h, w, v))
    gcn = np.real(ifft2(fft2(white_texture) * np.sqrt(fft2(gauss_acf))))
    texture = mnlt(gcn, v)
    speckle = ifft2(fft2(white_speckle) * speckle_spectrum(h))
    return np.abs(speckle * np.sqrt(texture)), texture
