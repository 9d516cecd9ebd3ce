"""The native optimizers of csrc/eincm_opt.h (host-only C++) against scipy.optimize.minimize on classic test functions:
same minima, scipy's status codes (0 converged, 1 maxiter).  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.optimize

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'native', 'opt_harness.cpp')
SO = os.path.join(HERE, 'native', '_opt_harness.so')


class Out(C.Structure):
    _fields_ = [('fun', C.c_double), ('nit', C.c_int), ('nfev', C.c_int), ('status', C.c_int), ('pad', C.c_int)]


@pytest.fixture(scope='module')
def lib():
    hdr = os.path.join(HERE, '..', 'edge-informed-contrast-maximization_b200', 'csrc', 'eincm_opt.h')
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', SO, SRC], check=True)
    return C.CDLL(SO)


def _rosen(x):
    return scipy.optimize.rosen(x), scipy.optimize.rosen_der(x)


@pytest.mark.parametrize('n', [2, 8, 32])
def test_bfgs_rosenbrock_matches_scipy(lib, n):
    x0 = np.linspace(-1.2, 1.0, n)
    ref = scipy.optimize.minimize(_rosen, x0, jac=True, method='BFGS', options={'gtol': 1e-7, 'maxiter': 2000})
    x = x0.copy()
    out = Out()
    assert lib.bfgs_rosenbrock(n, x.ctypes.data_as(C.POINTER(C.c_double)), 2000, C.c_double(1e-7), C.byref(out)) == 0
    assert out.status == 0 and ref.status == 0
    assert out.fun <= 1e-12 and np.abs(x - 1.0).max() <= 1e-5
    assert np.abs(x - ref.x).max() <= 1e-5
    assert out.nfev <= 2 * ref.nfev + 20                       # comparable work, not the same iterates


def test_bfgs_status_codes(lib):
    x = np.linspace(-1.2, 1.0, 8)
    out = Out()
    lib.bfgs_rosenbrock(8, x.ctypes.data_as(C.POINTER(C.c_double)), 5, C.c_double(1e-7), C.byref(out))
    assert out.status == 1 and out.nit == 5                    # maxiter, as scipy reports it
    f0 = scipy.optimize.rosen(np.linspace(-1.2, 1.0, 8))
    assert out.fun < f0
    x = np.ones(8)
    lib.bfgs_rosenbrock(8, x.ctypes.data_as(C.POINTER(C.c_double)), 5, C.c_double(1e-7), C.byref(out))
    assert out.status == 0 and out.nit == 0 and out.nfev == 1  # already converged


def test_bfgs_quartic_matches_scipy(lib):
    n = 16

    def f(v):
        c = 1.0 + 10.0 * np.arange(n); d = v - 0.1 * np.arange(n)
        return float((0.5 * c * d * d + 0.25 * d ** 4).sum()), c * d + d ** 3

    x0 = np.full(n, 3.0)
    ref = scipy.optimize.minimize(f, x0, jac=True, method='BFGS', options={'gtol': 1e-7, 'maxiter': 500})
    x = x0.copy()
    out = Out()
    assert lib.bfgs_quartic(n, x.ctypes.data_as(C.POINTER(C.c_double)), 500, C.c_double(1e-7), C.byref(out)) == 0
    assert out.status == 0
    np.testing.assert_allclose(x, 0.1 * np.arange(n), atol=1e-6)
    np.testing.assert_allclose(x, ref.x, atol=1e-6)


@pytest.mark.parametrize('m,lo,hi,a0', [(0.3, 0.0, 1.0, 0.5), (1.7, 0.0, 1.0, 0.5), (-0.4, 0.0, 1.0, 0.5), (0.6, 0.0, 1.0, 0.0)])
def test_bounded_scalar_matches_lbfgsb(lib, m, lo, hi, a0):
    def f(a):
        return float((a[0] - m) ** 2 + 0.3 * np.sin(5 * a[0])), np.array([2 * (a[0] - m) + 1.5 * np.cos(5 * a[0])])

    ref = scipy.optimize.minimize(f, np.array([a0]), jac=True, method='L-BFGS-B', bounds=[(lo, hi)], options={'gtol': 1e-6, 'maxiter': 20})
    a = C.c_double(a0)
    out = Out()
    assert lib.bounded_scalar_wavy(C.c_double(m), C.byref(a), C.c_double(lo), C.c_double(hi), 20, C.c_double(1e-6), C.byref(out)) == 0
    assert lo <= a.value <= hi
    assert out.fun <= ref.fun + 1e-7                            # at least as good a point as L-BFGS-B finds
    assert abs(a.value - ref.x[0]) <= 1e-3 or out.fun < ref.fun - 1e-9


@pytest.mark.parametrize('n', [1, 4])
@pytest.mark.parametrize('start', [0.0, 0.5, 0.9, 0.99, -3.0])
def test_bfgs_falls_back_to_wolfe2_like_scipy(lib, n, start):
    """scipy's _line_search_wolfe12: when dcsrch fails (here: a trial step leaves the domain of a log barrier, the objective is NaN)
    scalar_search_wolfe2 takes over - it bisects back into the domain, or, like scipy, walks on into the NaN region and ends with
    status 2.  Same status, iterations, final point and evaluations as scipy either way (the first version stopped with
    status 2 at the first dcsrch failure)."""
    def f(v):
        c = 1.0 + np.arange(n)
        with np.errstate(invalid='ignore', divide='ignore'):
            return float((c * (v - 2.0) ** 2 - np.log(1.0 - v)).sum()), 2.0 * c * (v - 2.0) + 1.0 / (1.0 - v)

    x0 = np.full(n, start)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = scipy.optimize.minimize(f, x0, jac=True, method='BFGS', options={'gtol': 1e-7, 'maxiter': 200})
    x = x0.copy()
    out = Out()
    assert lib.bfgs_barrier(n, x.ctypes.data_as(C.POINTER(C.c_double)), 200, C.c_double(1e-7), C.byref(out)) == 0
    assert (out.status, out.nit) == (ref.status, ref.nit)
    assert abs(out.nfev - ref.nfev) <= 1                       # one more evaluation of a NaN point on some failing searches
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-12)
    assert (np.isnan(out.fun) and np.isnan(ref.fun)) or abs(out.fun - ref.fun) <= 1e-12 * abs(ref.fun)


@pytest.mark.parametrize('n', [2, 8])
def test_bfgs_same_schedule_as_scipy(lib, n):
    """Same line-search policy and parameters: the native BFGS takes the iterations and evaluations scipy takes."""
    x0 = np.linspace(-1.2, 1.0, n)
    ref = scipy.optimize.minimize(_rosen, x0, jac=True, method='BFGS', options={'gtol': 1e-7, 'maxiter': 2000})
    x = x0.copy()
    out = Out()
    lib.bfgs_rosenbrock(n, x.ctypes.data_as(C.POINTER(C.c_double)), 2000, C.c_double(1e-7), C.byref(out))
    assert abs(out.nit - ref.nit) <= max(1, ref.nit // 20) and abs(out.nfev - ref.nfev) <= max(2, ref.nfev // 20)


@pytest.mark.parametrize('m,lo,hi,a0', [(0.3, 0.0, 1.0, 0.5), (1.7, 0.0, 1.0, 0.5), (-0.4, 0.0, 1.0, 0.5), (0.6, 0.0, 1.0, 0.0), (0.45, 0.0, 1.0, 0.9)])
def test_bounded_scalar_same_schedule_as_lbfgsb(lib, m, lo, hi, a0):
    def f(a):
        return float((a[0] - m) ** 2 + 0.3 * np.sin(5 * a[0])), np.array([2 * (a[0] - m) + 1.5 * np.cos(5 * a[0])])

    ref = scipy.optimize.minimize(f, np.array([a0]), jac=True, method='L-BFGS-B', bounds=[(lo, hi)], options={'gtol': 1e-6, 'maxiter': 20})
    a = C.c_double(a0)
    out = Out()
    assert lib.bounded_scalar_wavy(C.c_double(m), C.byref(a), C.c_double(lo), C.c_double(hi), 20, C.c_double(1e-6), C.byref(out)) == 0
    assert abs(a.value - ref.x[0]) <= 1e-6
    assert (out.nit, out.nfev) == (ref.nit, ref.nfev)


# ---------------------------------------------------------------------------------------------------------------------------------
# on the EINCM objective itself (evaluated by the CPU oracle): what the 'native' and 'graph' solver backends run inside the library,
# against the scipy calls jaxopt makes for the reference (src/eincm/solver.py:165-183)
# ---------------------------------------------------------------------------------------------------------------------------------
OBJ = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


@pytest.fixture(scope='module')
def window():
    import eincm_b200.synth as S
    rs = np.random.default_rng(5)
    return S.make_window(32, 48, 4000, seed=300, n_segments=40, scene_seed=77, truth_theta=rs.uniform(-4, 4, size=(2, 2, 2)))


@pytest.mark.parametrize('shape,lvl,maxiter', [((1, 1), 2, 8), ((2, 2), 1, 6), ((4, 4), 0, 5)])
def test_native_bfgs_follows_scipy_on_the_eincm_objective(lib, window, shape, lvl, maxiter):
    from oracle import eincm_oracle as O
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=lvl, n_pyr_lvls=3, sensor_size=window.sensor_size,
              scale_to_sensor_size_method='bilinear')
    n = shape[0] * shape[1] * 2
    calls = []

    def fun(x):
        v, g = O.value_and_grad(np.asarray(x, dtype=np.float64).reshape(shape + (2,)), *window.args(), **kw)
        calls.append(v)
        return v, g.ravel()

    @OBJ
    def cfun(xp, fp, gp):
        v, g = fun(np.ctypeslib.as_array(xp, shape=(n,)).copy())
        fp[0] = v
        np.ctypeslib.as_array(gp, shape=(n,))[:] = g
        return 0

    x0 = np.zeros(n)
    ref = scipy.optimize.minimize(fun, x0, jac=True, method='BFGS', options={'gtol': 1e-7, 'maxiter': maxiter})
    scipy_calls, calls[:] = list(calls), []          # what jaxopt's scipy_fun is really called with (scipy's nfev / njev count separately)
    x = x0.copy()
    out = Out()
    assert lib.bfgs_callback(n, x.ctypes.data_as(C.POINTER(C.c_double)), maxiter, C.c_double(1e-7), cfun, C.byref(out)) == 0
    assert (out.nit, out.status) == (ref.nit, ref.status)
    assert out.nfev == len(calls)
    print(f'evaluations: native {len(calls)}, scipy {len(scipy_calls)} (nfev {ref.nfev}, njev {ref.njev})')
    assert abs(len(calls) - len(scipy_calls)) <= max(3, len(scipy_calls) // 20)
    k = min(len(calls), len(scipy_calls), 10)
    np.testing.assert_allclose(calls[:k], scipy_calls[:k], rtol=1e-9)          # the same trial points from the start
    assert out.fun == pytest.approx(ref.fun, rel=1e-8)          # same algorithm, other summation order in the dot products
    np.testing.assert_allclose(x, ref.x, rtol=1e-5, atol=1e-6)


def test_native_bounded_scalar_follows_scipy_on_the_handover_objective(lib, window):
    from oracle import eincm_oracle as O
    import eincm_b200.synth as S
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=0, n_pyr_lvls=3, sensor_size=window.sensor_size,
              scale_to_sensor_size_method='bilinear')
    pts = S.theta_test_points(window, (4, 4))
    prev, cur = pts['truth'], pts['perturbed']

    def fun(a):
        v, da = O.handover_value_and_grad(float(a[0]), prev, cur, *window.args(), **kw)
        return v, np.array([da])

    @OBJ
    def cfun(xp, fp, gp):
        v, g = fun([xp[0]])
        fp[0], gp[0] = v, g[0]
        return 0

    ref = scipy.optimize.minimize(fun, np.array([0.5]), jac=True, method='L-BFGS-B', bounds=[(0.0, 1.0)], options={'maxiter': 6, 'gtol': 1e-6})
    a = C.c_double(0.5)
    out = Out()
    assert lib.bounded_scalar_callback(C.byref(a), C.c_double(0.0), C.c_double(1.0), 6, C.c_double(1e-6), cfun, C.byref(out)) == 0
    assert a.value == pytest.approx(float(ref.x[0]), abs=1e-8)
    assert out.fun == pytest.approx(float(ref.fun), rel=1e-10)
    assert (out.nit, out.nfev) == (ref.nit, ref.nfev)
