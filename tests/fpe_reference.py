"""Exact (continuum) ensemble statistics of ONE uniaxial macrospin in a field along its easy axis: test infrastructure.

The stochastic LLG equation the reference integrates (lib/llg.cpp:14-29, 92-106 in the reduced units of
lib/simulation.cpp:498-549) has, for an axially symmetric problem, the Fokker-Planck equation in x = cos(theta)

    dW/dt = 1/(2 tau_N) d/dx [ (1 - x^2) ( dW/dx + W dU/dx ) ],      U = -sigma x^2 - 2 sigma h(t) x,
    tau_N = sigma / alpha   (reduced time),   sigma = K V / (k_B T),   h = H / H_k

(Brown 1963).  Expanding in Legendre polynomials, f_n = <P_n(cos theta)>, gives the pentadiagonal hierarchy

    2 tau_N / (n (n + 1)) df_n/dt = - f_n + xi / (2n + 1) (f_{n-1} - f_{n+1})
          + 2 sigma / (2n + 1) { [n f_n + (n - 1) f_{n-2}] / (2n - 1) - [(n + 2) f_{n+2} + (n + 1) f_n] / (2n + 3) },   xi = 2 sigma h

derived in DESIGN.md section 10 by integrating P_n against the equation above ((1-x^2) P_n' = n(n+1)/(2n+1) (P_{n-1} - P_{n+1}),
x P_m = ((m+1) P_{m+1} + m P_{m-1}) / (2m+1)).  f_1(t) is the ensemble mean <m_z>(t) of an infinite sLLG ensemble at
dt -> 0 — the comparator that is neither the two-state approximation of lib/dom.cpp (which drops the intra-well motion and
uses the high-barrier rate asymptote) nor a second Monte-Carlo run.  The same matrix gives the exact Neel relaxation
rate (smallest odd-mode eigenvalue at zero field)."""
import numpy as np
from scipy.integrate import solve_ivp

KB, MU0, GYROMAG = 1.38064852e-23, 1.25663706e-6, 1.76086e11   # include/constants.hpp:10-12


def reduced(radius, K, Ms, alpha, T, H0=0.0, f=0.0):
    V = 4.0 / 3.0 * np.pi * radius ** 3
    H_k = 2 * K / MU0 / Ms
    tau = GYROMAG * MU0 * H_k / (1 + alpha * alpha)
    return dict(sigma=K * V / (KB * T), H_k=H_k, time_factor=tau, h0=H0 / H_k, f_red=f / tau, V=V)


def hierarchy(sigma, alpha, nmax):
    """df/dt = (A0 + h A1) f + (c0 + h c1) for f = (f_1 ... f_nmax) in reduced time; f_0 = 1 sits in c0, c1."""
    A0 = np.zeros((nmax, nmax)); A1 = np.zeros((nmax, nmax)); c0 = np.zeros(nmax); c1 = np.zeros(nmax)
    for n in range(1, nmax + 1):
        pre = n * (n + 1) * alpha / (2.0 * sigma)
        i = n - 1

        def add(M, c, m, val):
            if m == 0:
                c[i] += pre * val
            elif 1 <= m <= nmax:
                M[i, m - 1] += pre * val
        add(A0, c0, n, -1.0 + 2.0 * sigma / ((2 * n - 1) * (2 * n + 3)))
        add(A0, c0, n - 2, 2.0 * sigma * (n - 1) / ((2 * n + 1) * (2 * n - 1)))
        add(A0, c0, n + 2, -2.0 * sigma * (n + 2) / ((2 * n + 1) * (2 * n + 3)))
        add(A1, c1, n - 1, 2.0 * sigma / (2 * n + 1))
        add(A1, c1, n + 1, -2.0 * sigma / (2 * n + 1))
    return A0, A1, c0, c1


def neel_rate(sigma, alpha, nmax=120):
    """Smallest relaxation rate of <m_z> at zero field (reduced time): the Neel-Brown rate without the asymptote."""
    A0 = hierarchy(sigma, alpha, nmax)[0]
    odd = A0[0::2, 0::2]                      # odd n couple to odd n only when h = 0
    ev = np.linalg.eigvals(odd)
    return float(np.min(-ev.real))


def mean_mz(sigma, alpha, h_of_t, t_eval, f_init=None, nmax=64, rtol=1e-10, atol=1e-13):
    """<m_z>(t) = f_1(t) (and f_2) at the reduced times `t_eval` for the field h_of_t(t); f_init = initial moments
    f_1..f_nmax (default: all members at m = +z, f_n = 1)."""
    A0, A1, c0, c1 = hierarchy(sigma, alpha, nmax)
    y0 = np.ones(nmax) if f_init is None else np.asarray(f_init, dtype=float)

    def rhs(t, y):
        h = h_of_t(t)
        return (A0 + h * A1) @ y + c0 + h * c1

    def banded(M):   # LSODA's packed form: ab[uband + i - j, j] = M[i, j], lband = uband = 2
        ab = np.zeros((5, nmax))
        for d in range(-2, 3):
            diag = np.diagonal(M, offset=-d)      # M[j + d, j]
            if d >= 0:
                ab[2 + d, :nmax - d] = diag
            else:
                ab[2 + d, -d:] = diag
        return ab
    B0, B1 = banded(A0), banded(A1)

    def jac(t, y):
        return B0 + h_of_t(t) * B1
    sol = solve_ivp(rhs, (0.0, float(t_eval[-1])), y0, method='LSODA', t_eval=t_eval, jac=jac, lband=2, uband=2,
                    rtol=rtol, atol=atol)
    assert sol.success, sol.message
    return sol.y[0], sol.y[1]


def equilibrium_moments(sigma, xi, nmax):
    """f_n of the Boltzmann distribution exp(sigma x^2 + xi x) by quadrature."""
    x, w = np.polynomial.legendre.leggauss(400)
    p = np.exp(sigma * x * x + xi * x - (sigma + abs(xi)))
    Z = np.sum(w * p)
    P = np.polynomial.legendre.legvander(x, nmax)       # [len(x), nmax + 1]
    return (P[:, 1:] * (w * p)[:, None]).sum(axis=0) / Z


def bonferroni_z(m, alpha=0.0027):
    """Two-sided z bound for m simultaneous comparisons at the family-wise level of one 3-sigma test."""
    from scipy.stats import norm
    return float(norm.isf(alpha / (2.0 * m)))
