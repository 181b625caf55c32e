"""ctypes wrapper around the C oracle (oracle/sph_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  The product package
(python-sph_b200/) never imports this module.

The oracle restates the reference's O(N^2) algorithm (pyfluid.py) in plain C; it is pinned
against fixtures generated from the reference itself (tests/golden/make_golden.py).
Function names mirror the reference's: ``density(j, pos, h_j)`` is pyfluid.py:218 etc.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# status bits (same values as include/sphb200.h SPHB_ST_*)
NEG_H, NEG_P, NEG_PI, NEG_VISC, NEG_E, BISECT_NONE = 1, 2, 4, 8, 16, 32

# pyfluid.py:5-41 defaults
DEFAULTS = dict(
    PARTICLE_MASS=1.0, COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
    BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
    INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=39.4227, ADIABATIC_INDEX=5 / 3,
)


class Consts(C.Structure):
    _fields_ = [
        ("particle_mass", C.c_double), ("coupling_const", C.c_double), ("h_tol", C.c_double),
        ("newton_limit", C.c_int32), ("bisect_limit", C.c_int32),
        ("alpha", C.c_double), ("beta", C.c_double), ("epsilon", C.c_double),
        ("bisect_tol", C.c_double), ("bisect_a", C.c_double), ("bisect_b", C.c_double),
        ("G", C.c_double), ("gamma", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_distance.restype = C.c_double
        _lib.oracle_M4.restype = C.c_double
        _lib.oracle_M4_d1.restype = C.c_double
        _lib.oracle_M4_dh.restype = C.c_double
        for f in ("oracle_M4", "oracle_M4_d1", "oracle_M4_dh"):
            getattr(_lib, f).argtypes = [C.c_double, C.c_double]
        _lib.oracle_neighbours.restype = C.c_int64
        _lib.oracle_neighbours_sym.restype = C.c_int64
        _lib.oracle_newton_arr.restype = C.c_int32
        _lib.oracle_leapfrog.restype = C.c_int32
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The reference's functions for given module constants (read at construction)."""

    def __init__(self, **overrides):
        k = dict(DEFAULTS)
        k.update(overrides)
        self.k = k
        self.c = Consts(
            float(k["PARTICLE_MASS"]), float(k["COUPLING_CONST"]), float(k["SMOOTHLENGTH_VARIATION_TOLERANCE"]),
            int(k["NEWTON_ITERATION_LIMIT"]), int(k["BISECTION_ITERATION_LIMIT"]),
            float(k["ALPHA_SPH"]), float(k["BETA_SPH"]), float(k["EPSILON"]), float(k["BISECTION_TOLERANCE"]),
            float(k["INITIAL_BISECTION_SMOOTHLENGTH_A"]), float(k["INITIAL_BISECTION_SMOOTHLENGTH_B"]),
            float(k["G"]), float(k["ADIABATIC_INDEX"]),
        )
        self.L = lib()

    # -- scalar kernel functions: pyfluid.py:43,56,64,85 -----------------------------------
    def distance(self, rj, ri):
        rj, ri = _f64(rj), _f64(ri)
        return self.L.oracle_distance(_p(rj), _p(ri))

    def M4(self, dist, h):
        return self.L.oracle_M4(float(dist), float(h))

    def M4_d1(self, dist, h):
        return self.L.oracle_M4_d1(float(dist), float(h))

    def M4_smoothlength_derivative(self, dist, h):
        return self.L.oracle_M4_dh(float(dist), float(h))

    # -- list forms -------------------------------------------------------------------------
    @staticmethod
    def _jl(j, n):
        if j is None:
            return np.arange(n, dtype=np.int64)
        return np.ascontiguousarray(np.atleast_1d(j), dtype=np.int64)

    def density(self, j, pos, h_j):
        """pyfluid.py:218 for each j in the list (h_j: same length)."""
        pos = _f64(pos); jl = self._jl(j, len(pos)); h_j = _f64(np.broadcast_to(h_j, jl.shape))
        out = np.empty(len(jl))
        self.L.oracle_density_list(C.byref(self.c), _p(pos), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(h_j), _p(out))
        return out

    def density_arr(self, pos, h):
        """pyfluid.py:226"""
        return self.density(None, pos, h)

    def var_density_arr(self, h):
        """pyfluid.py:240 (same pow() the C side uses)"""
        h = _f64(h)
        m = getattr(self, "_masses", None)
        return np.array([(self.k["PARTICLE_MASS"] if m is None else float(m[i])) * (self.k["COUPLING_CONST"] / float(x)) ** 3
                         for i, x in enumerate(h)])

    def omega_j(self, j, pos, rho_j, h_j):
        """pyfluid.py:91 for each j in the list."""
        pos = _f64(pos); jl = self._jl(j, len(pos))
        rho_j = _f64(np.broadcast_to(rho_j, jl.shape)); h_j = _f64(np.broadcast_to(h_j, jl.shape))
        out = np.empty(len(jl))
        self.L.oracle_omega_list(C.byref(self.c), _p(pos), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(rho_j), _p(h_j), _p(out))
        return out

    def omega_arr(self, pos, rho, h):
        """pyfluid.py:99"""
        return self.omega_j(None, pos, rho, h)

    def xi_arr(self, rho, h, pos, j=None):
        """pyfluid.py:412 (or pyfluid.py:401 for a list of j)"""
        pos = _f64(pos); jl = self._jl(j, len(pos))
        rho_j = _f64(rho)[jl] if j is not None else _f64(rho); h_j = _f64(h)[jl] if j is not None else _f64(h)
        out = np.empty(len(jl))
        self.L.oracle_xi_list(C.byref(self.c), _p(pos), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(_f64(rho_j)), _p(_f64(h_j)), _p(out))
        return out

    def newton_smoothlength_while(self, j, pos, h_guess_j):
        """pyfluid.py:165 for each j; returns (h, iterations, code, out_of_bounds_prints);
        code 0 = Newton converged, 1 = bisection fallback, 2 = bisection returned None."""
        pos = _f64(pos); jl = self._jl(j, len(pos)); hg = _f64(np.broadcast_to(h_guess_j, jl.shape))
        h = np.empty(len(jl)); it = np.empty(len(jl), np.int32); code = np.empty(len(jl), np.int32); oob = np.empty(len(jl), np.int32)
        self.L.oracle_newton_list(C.byref(self.c), _p(pos), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(hg), _p(h), _p(it), _p(code), _p(oob))
        return h, it, code, oob

    def bisection_h(self, j, pos):
        """pyfluid.py:123 for each j; returns (h, code, oob)."""
        pos = _f64(pos); jl = self._jl(j, len(pos))
        h = np.empty(len(jl)); code = np.empty(len(jl), np.int32); oob = np.empty(len(jl), np.int32)
        self.L.oracle_bisection_list(C.byref(self.c), _p(pos), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(h), _p(code), _p(oob))
        return h, code, oob

    def newton_smoothlength_arr(self, pos, h_guess, details=False):
        """pyfluid.py:203.  A bisection `None` (:153) becomes nan when stored into the float64
        array at :206 (numpy >= 2 behaviour, measured by make_golden.py case F), so only the
        negative-h RuntimeError (:209-210) can be raised."""
        pos = _f64(pos); hg = _f64(np.broadcast_to(h_guess, (len(pos),)))
        h = np.empty(len(pos)); it = np.empty(len(pos), np.int32); code = np.empty(len(pos), np.int32)
        st = self.L.oracle_newton_arr(C.byref(self.c), _p(pos), C.c_int64(len(pos)), _p(hg), _p(h), _p(it), _p(code))
        if st & NEG_H:
            raise RuntimeError("Obtained negative smoothing length.")
        return (h, it, code) if details else h

    def pressure_arr(self, e, rho):
        """pyfluid.py:251"""
        e, rho = _f64(e), _f64(rho)
        P = (self.k["ADIABATIC_INDEX"] - 1) * e * rho
        if np.any(P < 0):
            raise RuntimeError("Obtained negative pressure.")
        return P

    def neighbours(self, pos, j, h_j):
        """NB(j) = {i : distance(pos[j],pos[i]) / h_j < 2}, ascending."""
        pos = _f64(pos); out = np.empty(len(pos), np.int64)
        n = self.L.oracle_neighbours(_p(pos), C.c_int64(len(pos)), C.c_int64(int(j)), C.c_double(float(h_j)), _p(out), C.c_int64(len(pos)))
        return out[:n].copy()

    def neighbours_sym(self, pos, h, j):
        """Force support: i in NB(j) or j in NB(i)."""
        pos, h = _f64(pos), _f64(h); out = np.empty(len(pos), np.int64)
        n = self.L.oracle_neighbours_sym(_p(pos), _p(h), C.c_int64(len(pos)), C.c_int64(int(j)), _p(out), C.c_int64(len(pos)))
        return out[:n].copy()

    @staticmethod
    def _raise(st):
        if st & NEG_P:
            raise RuntimeError("Obtained negative pressure.")
        if st & NEG_PI:
            raise RuntimeError("Obtained negative value for Pi")
        if st & NEG_VISC:
            raise RuntimeError("Obtained negative energy rate due to viscosity.")
        if st & NEG_E:
            raise RuntimeError("Obtained negative energy.")
        if st & NEG_H:
            raise RuntimeError("Obtained negative smoothing length.")

    def energy_rate_arr(self, pos, vel, P, rho, h, omega, j=None):
        """pyfluid.py:355 (or :316 for a list of j)"""
        pos, vel, P, rho, h, omega = map(_f64, (pos, vel, P, rho, h, omega)); jl = self._jl(j, len(pos))
        out = np.empty(len(jl)); st = C.c_int32(0)
        self.L.oracle_energy_rate_list(C.byref(self.c), _p(pos), _p(vel), _p(P), _p(rho), _p(h), _p(omega), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(out), C.byref(st))
        self._raise(st.value)
        return out

    def acceleration_arr(self, pos, rho, vel, P, h, omega, xi, j=None):
        """pyfluid.py:487 (or :477 for a list of j); includes gravity when G != 0"""
        pos, rho, vel, P, h, omega, xi = map(_f64, (pos, rho, vel, P, h, omega, xi)); jl = self._jl(j, len(pos))
        out = np.empty((len(jl), 3)); st = C.c_int32(0)
        self.L.oracle_acceleration_list(C.byref(self.c), _p(pos), _p(rho), _p(vel), _p(P), _p(h), _p(omega), _p(xi), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(out), C.byref(st))
        self._raise(st.value)
        return out

    def var_smoothlength_leapfrog(self, pos0, vel0, e0, h0, dt):
        """pyfluid.py:498"""
        pos0, vel0, e0, h0 = map(_f64, (pos0, vel0, e0, h0)); n = len(pos0)
        pos1 = np.empty((n, 3)); vel1 = np.empty((n, 3)); e1 = np.empty(n); h1 = np.empty(n)
        st = self.L.oracle_leapfrog(C.byref(self.c), _p(pos0), _p(vel0), _p(e0), _p(h0), C.c_double(float(dt)), C.c_int64(n), _p(pos1), _p(vel1), _p(e1), _p(h1))
        if st < 0:
            raise MemoryError("oracle_leapfrog: allocation failed")
        self._raise(st)
        return pos1, vel1, e1, h1

    def var_smoothlength_sim(self, time_arr, pos0, vel0, e0, h_approx):
        """pyfluid.py:535 (without the per-step print)"""
        dt = time_arr[1] - time_arr[0]; Nt = len(time_arr); N = len(e0)
        P = np.zeros((Nt, N, 3)); V = np.zeros((Nt, N, 3)); E = np.zeros((Nt, N)); H = np.zeros((Nt, N))
        P[0], V[0], E[0] = pos0, vel0, e0
        H[0] = self.newton_smoothlength_arr(P[0], h_approx)
        for t in range(Nt - 1):
            P[t + 1], V[t + 1], E[t + 1], H[t + 1] = self.var_smoothlength_leapfrog(P[t], V[t], E[t], H[t], dt)
        return P, V, E, H

    # -- static-h path, REPAIRED: small-N numpy restatement over the C primitives --------
    def energy_evolve_arr(self, pos, vel, e, P, rho, h, dt):
        """pyfluid.py:292-314: e_j + P_j/rho_j^2 * sum_i m (v_j - v_i).dellM4(r_j, r_i, h_j) * dt, index-order sum,
        dellM4 = M4_d1(distance) * direction_i_j (:50-54, :81-82; zero for coincident positions)"""
        pos, vel, e, P, rho, h = map(_f64, (pos, vel, e, P, rho, h)); n = len(pos)
        m = float(self.k["PARTICLE_MASS"])
        out = np.empty(n)
        for j in range(n):
            change = 0.0
            for i in range(n):
                if np.array_equal(pos[j], pos[i]):
                    grad = np.zeros(3)
                else:
                    d = self.distance(pos[j], pos[i])
                    grad = self.M4_d1(d, h[j]) * ((pos[j] - pos[i]) / d)
                change += m * np.dot(vel[j] - vel[i], grad)
            out[j] = e[j] + P[j] / rho[j] ** 2 * change * dt
        return out

    def static_smoothlength_leapfrog(self, pos0, vel0, e0, dt, h_static=2.0):
        """pyfluid.py:607-621 with the minimal repair of the two broken calls (acceleration_arr with Omega = 1, xi = 0;
        energy_evolve_arr with its unused 7th argument); h_static = DEFAULT_SMOOTHLENGTH (:6)"""
        pos0, vel0, e0 = map(_f64, (pos0, vel0, e0)); n = len(pos0)
        hs, one, zero = np.full(n, float(h_static)), np.ones(n), np.zeros(n)
        den0 = self.density_arr(pos0, hs)
        press0 = self.pressure_arr(e0, den0)
        acc0 = self.acceleration_arr(pos0, den0, vel0, press0, hs, one, zero)
        pos1 = pos0 + vel0 * dt + 0.5 * acc0 * dt ** 2
        e1 = self.energy_evolve_arr(pos0, vel0, e0, press0, den0, hs, dt)
        den1 = self.density_arr(pos1, hs)
        press1 = self.pressure_arr(e1, den1)
        acc1 = self.acceleration_arr(pos1, den1, vel0, press1, hs, one, zero)
        vel1 = vel0 + 0.5 * (acc0 + acc1) * dt
        return pos1, vel1, e1

    def step_work_sample(self, pos, vel, e, h, jlist):
        """Per-particle share of one step for the sampled j (bench.py cpu_baseline)."""
        pos, vel, e, h = map(_f64, (pos, vel, e, h)); jl = self._jl(jlist, len(pos)); sink = np.empty(len(jl))
        self.L.oracle_step_work_sample(C.byref(self.c), _p(pos), _p(vel), _p(e), _p(h), C.c_int64(len(pos)), _p(jl), C.c_int64(len(jl)), _p(sink))
        return sink

    def set_wide(self, on):
        """bench.py only: split each per-particle O(N) loop over the OpenMP threads"""
        self.L.oracle_set_wide(C.c_int(1 if on else 0))

    def set_masses(self, m):
        """per-particle masses (the reference has one scalar PARTICLE_MASS, pyfluid.py:3-5): every
        PARTICLE_MASS inside a sum over i becomes m[i], the one of var_density / zetaprime the target's m[j].
        None: back to the reference's scalar.  The array is process-wide state of the oracle library, like its
        thread count; keep `m` alive until it is unset."""
        if m is None:
            self._masses = None
            self.L.oracle_set_masses(None)
            return
        self._masses = np.ascontiguousarray(m, dtype=np.float64)
        self.L.oracle_set_masses(self._masses.ctypes.data_as(C.c_void_p))

    def num_threads(self):
        return int(self.L.oracle_num_threads())

    def set_num_threads(self, n):
        self.L.oracle_set_num_threads(C.c_int(int(n)))
