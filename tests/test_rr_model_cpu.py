"""The small eigen-solve of the eigen-solver's two-stage Rayleigh-Ritz step, restated in numpy
(tools/lobpcg_study.py rqi3_lowest; the device routine geig3_lowest_rqi in csrc/mac.cu follows it line
by line): Rayleigh-quotient iteration from e_0 with adjugate solves, a lowestness check through leading
minors, one deflation + closed-form 2 x 2 solve when the iteration lands on another pair.  Whatever it
returns must be the LOWEST pair of the pencil; None (caller takes the Jacobi path) must stay rare."""
import importlib.util
import os

import numpy as np
from scipy.linalg import eigh

_spec = importlib.util.spec_from_file_location(
    "lobpcg_study", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "lobpcg_study.py"))
study = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(study)


def _pencil(rng, k, close=False):
    """3 x 3 pencil with unit-diagonal-ish B whose eigenvector number k is close to e_0."""
    Rm = rng.normal(size=(3, 3))
    B = np.eye(3) + 0.2 * (Rm + Rm.T) / 2
    lam = np.sort(rng.uniform(1e-6, 1.0, 3)) * (10.0 ** rng.integers(-6, 1))
    if close:
        lam[1] = lam[0] * (1 + 1e-4)
    Lc = np.linalg.cholesky(B)
    V = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    V[:, 0] = Lc.T @ np.array([1.0, 0.0, 0.0]) + 0.05 * rng.normal(size=3)
    V = np.linalg.qr(V)[0]
    Y = np.linalg.solve(Lc.T, V)
    order = [k] + [q for q in range(3) if q != k]
    Yi = np.linalg.inv(Y)
    A = Yi.T @ np.diag(lam[order]) @ Yi
    return 0.5 * (A + A.T), B


def test_rqi_model_returns_the_lowest_pair():
    rng = np.random.default_rng(5)
    study.RQI_STATS.update(calls=0, fallback=0, steps=0, deflated=0)
    for trial in range(300):
        A, B = _pencil(rng, trial % 3, close=trial % 10 == 9)
        got = study.rqi3_lowest(A, B)
        lam, Y = eigh(A, B)
        if got is None:
            continue
        th, y = got
        scale = np.abs(np.diag(A)).sum()
        # the lowest value, or (close pair) a value within the check's delta of it
        assert th <= lam[0] + 2e-9 * scale + 2e-4 * lam[0] * (trial % 10 == 9)
        assert abs(y @ B @ y - 1.0) < 1e-10
        assert abs(y @ A @ y - th) <= 1e-9 * scale
    assert study.RQI_STATS["fallback"] <= 0.05 * study.RQI_STATS["calls"]
    assert study.RQI_STATS["deflated"] >= 100       # the e_0-next-to-another-eigenvector cases took the deflation step
