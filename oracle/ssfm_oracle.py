"""CPU oracle for the split-step Fourier hot path (FIBER / DBP).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this file;
nothing under ``opticomlib_b200/`` does.  It is the checker, never the product.

What it restates
----------------
``opticomlib.devices.FIBER`` (reference ``opticomlib/devices.py:1137-1196``)
and ``DBP`` (``devices.py:1280-1283``) as plain NumPy, with the real dtype as a
parameter:

* ``real=np.float32``  -> the algorithm exactly as shipped (every scalar cast
  to float32, field complex64, ``devices.py:1137-1147``).  Under the NumPy 2.x
  of this image the shipped code stays complex64 end to end, and this
  restatement reproduces it with rel-L2 == 0.0 (pinned by
  ``tests/test_oracle_vs_reference.py`` and the fixtures in ``tests/golden``).
* ``real=np.float64``  -> the "dtype-lifted" oracle: the same statements with
  float64 / complex128.  It is cross-checked against the reference's own
  float64 loop ``animated_fiber_propagation_with_phase``
  (``devices.py:2440-2486``) in the same test file.

Parity status: PINNED.  The reference ships no golden vectors for this path
(SURVEY.md §4), so the pins are outputs of the reference itself executed in the
build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

Third-party arithmetic the reference delegates to and that is NOT under
/root/reference: ``numpy.fft`` (pocketfft; requirements.txt pins numpy==1.26.4,
the image has 2.3.5 -- see SURVEY.md F2) and NumPy's exp/abs ufuncs.  The oracle
calls the same NumPy functions, so on one machine it is the reference bit for bit.
"""
from __future__ import annotations

import numpy as np

_DB_PER_NEPER = 4.343  # literal used at devices.py:1137 (not 10/ln 10)


def _ctype(real):
    return np.complex64 if np.dtype(real) == np.float32 else np.complex128


def omega_rad_per_ps(n: int, dt: float, real=np.float32) -> np.ndarray:
    """Angular-frequency grid of the linear operator, in rad/ps.

    Follows ``optical_signal.w()`` (typing.py:1641: ``fftfreq(size, gv.dt)*2*pi``)
    and the cast at devices.py:1144 (``* 1e-12`` in float64, then rounded).
    """
    w64 = np.fft.fftfreq(n, dt) * 2 * np.pi
    return np.asarray(w64 * 1e-12, dtype=real)


def linear_operator(n, dt, alpha_db_km, beta_2, beta_3, real=np.float32):
    """D~(w) = -alpha/2 + j/2 b2 w^2 + j/6 b3 w^3   (devices.py:1137-1145)."""
    a = np.array(alpha_db_km / _DB_PER_NEPER, dtype=real)
    b2 = np.array(beta_2, dtype=real)
    b3 = np.array(beta_3, dtype=real)
    w = omega_rad_per_ps(n, dt, real)
    return -a / 2 + 1j / 2 * b2 * w**2 + 1j / 6 * b3 * w**3


def oracle_fiber(
    field,
    dt,
    length,
    alpha=0.0,
    beta_2=0.0,
    beta_3=0.0,
    gamma=0.0,
    phi_max=0.01,
    h=None,
    real=np.float32,
    return_steps=False,
    max_steps=None,
):
    """Split-step propagation of ``field`` ((N,) or (P,N) complex).

    Returns a dict: ``out`` (complex, same shape), ``z`` (positions after each
    step, dtype ``real``), ``h`` (step sizes actually used, dtype ``real``),
    ``steps`` (int) and, with ``return_steps``, ``traj`` ((steps+1, ...)).

    Statement-for-statement correspondence with devices.py:
      1137-1142 scalar casts   1144-1145 operator   1147 field cast
      1155-1161 first step     1172-1181 the split step
      1193-1196 step-size controller.
    """
    R = np.dtype(real).type
    C = _ctype(real)

    a_lin = np.array(alpha / _DB_PER_NEPER, dtype=R)
    b2 = np.array(beta_2, dtype=R)
    b3 = np.array(beta_3, dtype=R)
    g = np.array(gamma, dtype=R)
    L = np.array(length, dtype=R)
    pm = np.array(phi_max, dtype=R)

    A = np.asarray(field, dtype=C)
    n = A.shape[-1]
    w = omega_rad_per_ps(n, dt, R)
    D = -a_lin / 2 + 1j / 2 * b2 * w**2 + 1j / 6 * b3 * w**3

    traj = [A.copy()] if return_steps else None

    with np.errstate(divide="ignore", invalid="ignore"):
        if h is None:
            if (b2 == 0 and b3 == 0) or g == 0:
                hk = L
            else:
                hk = pm / (np.abs(g) * (np.abs(A) ** 2)).max()
        else:
            hk = np.array(h, dtype=R)
        hk = np.array(min(hk, L), dtype=R)

        z = np.array(0, dtype=R)
        z_log, h_log = [], []
        while z < L:
            z += hk
            nl = 1j * g * np.abs(A) ** 2  # frozen for both half steps (F4)
            A = A * np.exp(hk / 2 * nl)
            A = np.fft.fft(A)
            A = A * np.exp(D * hk)
            A = np.fft.ifft(A)
            A = A * np.exp(hk / 2 * nl)

            z_log.append(z.copy())
            h_log.append(hk.copy())
            if return_steps:
                traj.append(A.copy())

            if h is None:
                hk = pm / (np.abs(g) * (np.abs(A) ** 2)).max()
            hk = np.array(min(hk, L - z), dtype=R)
            if max_steps is not None and len(z_log) >= max_steps:
                break

    res = {
        "out": A,
        "z": np.array(z_log, dtype=R),
        "h": np.array(h_log, dtype=R),
        "steps": len(z_log),
    }
    if return_steps:
        res["traj"] = np.array(traj)
    return res


def oracle_dbp(field, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, **kw):
    """devices.py:1280-1283: FIBER with alpha, beta_2, beta_3, gamma negated."""
    return oracle_fiber(field, dt, length, alpha=-alpha, beta_2=-beta_2,
                        beta_3=-beta_3, gamma=-gamma, **kw)


def rel_l2(a, b) -> float:
    """||a-b||_2 / ||b||_2 in float64 (the parity metric of BASELINE.json)."""
    a = np.asarray(a, dtype=np.complex128).ravel()
    b = np.asarray(b, dtype=np.complex128).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
