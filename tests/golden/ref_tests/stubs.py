"""Stand-ins for the two third-party packages the reference's tests import but this image lacks.

matplotlib.pyplot -> objects that accept every call (the tests only draw and save plots with it);
pykalman          -> a scalar-state Kalman filter / RTS smoother / EM in numpy float64 with pykalman's
                     defaults (identity matrices, zero offsets, em_vars = transition_covariance,
                     observation_covariance, initial_state_mean, initial_state_covariance; n_iter = 10),
                     which is all test_inference.py:TestInfer and models/lgssm.py use.
Test infrastructure only."""
import sys
import types
from unittest import mock

import numpy as np


def _subplots(nrows=1, ncols=1, *args, **kwargs):
    fig = mock.MagicMock(name="figure")
    n = nrows * ncols
    if n == 1:
        return fig, mock.MagicMock(name="axes")
    return fig, [mock.MagicMock(name="axes%d" % i) for i in range(n)]


class KalmanFilter:
    def __init__(self, transition_matrices=None, observation_matrices=None, transition_covariance=None,
                 observation_covariance=None, transition_offsets=None, observation_offsets=None,
                 initial_state_mean=None, initial_state_covariance=None, n_dim_state=None, n_dim_obs=None,
                 em_vars=("transition_covariance", "observation_covariance", "initial_state_mean",
                          "initial_state_covariance")):
        def scalar(v, default):
            return float(np.asarray(default if v is None else v, dtype=np.float64).reshape(-1)[0])
        self._A, self._C = scalar(transition_matrices, 1.0), scalar(observation_matrices, 1.0)
        self._Q, self._R = scalar(transition_covariance, 1.0), scalar(observation_covariance, 1.0)
        self._b, self._d = scalar(transition_offsets, 0.0), scalar(observation_offsets, 0.0)
        self._m0, self._P0 = scalar(initial_state_mean, 0.0), scalar(initial_state_covariance, 1.0)
        self.em_vars = tuple(em_vars)

    # pykalman's attribute names, array-shaped like the real thing
    transition_matrices = property(lambda s: np.array([[s._A]]))
    observation_matrices = property(lambda s: np.array([[s._C]]))
    transition_covariance = property(lambda s: np.array([[s._Q]]))
    observation_covariance = property(lambda s: np.array([[s._R]]))
    transition_offsets = property(lambda s: np.array([s._b]))
    observation_offsets = property(lambda s: np.array([s._d]))
    initial_state_mean = property(lambda s: np.array([s._m0]))
    initial_state_covariance = property(lambda s: np.array([[s._P0]]))

    def _filter_smooth(self, y):
        T = len(y)
        mp, Pp, mf, Pf = (np.zeros(T) for _ in range(4))
        for t in range(T):
            if t == 0:
                mp[t], Pp[t] = self._m0, self._P0
            else:
                mp[t], Pp[t] = self._A * mf[t - 1] + self._b, self._A * Pf[t - 1] * self._A + self._Q
            S = self._C * Pp[t] * self._C + self._R
            G = Pp[t] * self._C / S
            mf[t] = mp[t] + G * (y[t] - self._C * mp[t] - self._d)
            Pf[t] = Pp[t] - G * self._C * Pp[t]
        ms, Ps, J = mf.copy(), Pf.copy(), np.zeros(T)
        for t in range(T - 2, -1, -1):
            J[t] = Pf[t] * self._A / Pp[t + 1]
            ms[t] = mf[t] + J[t] * (ms[t + 1] - mp[t + 1])
            Ps[t] = Pf[t] + J[t] * (Ps[t + 1] - Pp[t + 1]) * J[t]
        return ms, Ps, J

    def smooth(self, observations):
        y = np.asarray(observations, dtype=np.float64).reshape(-1)
        ms, Ps, _ = self._filter_smooth(y)
        return ms.reshape(-1, 1), Ps.reshape(-1, 1, 1)

    def em(self, observations, n_iter=10):
        y = np.asarray(observations, dtype=np.float64).reshape(-1)
        T = len(y)
        for _ in range(n_iter):
            ms, Ps, J = self._filter_smooth(y)
            pair = J[:-1] * Ps[1:]  # Cov(x_t, x_{t+1} | y)
            if "observation_covariance" in self.em_vars:
                err = y - self._C * ms - self._d
                R = float(np.mean(err ** 2 + self._C * Ps * self._C))
            if "transition_covariance" in self.em_vars:
                err = ms[1:] - self._A * ms[:-1] - self._b
                Q = float(np.mean(err ** 2 + self._A * Ps[:-1] * self._A + Ps[1:] - 2 * self._A * pair)) if T > 1 else self._Q
            if "observation_covariance" in self.em_vars:
                self._R = R
            if "transition_covariance" in self.em_vars:
                self._Q = Q
            if "initial_state_mean" in self.em_vars:
                self._m0 = float(ms[0])
            if "initial_state_covariance" in self.em_vars:
                self._P0 = float(Ps[0] + (ms[0] - self._m0) ** 2)
        return self


def install():
    """Register the stubs for whichever of matplotlib / pykalman cannot be imported; returns the names added."""
    added = []
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.subplots = _subplots
        plt.__getattr__ = lambda name: mock.MagicMock(name="pyplot." + name)
        mpl.pyplot = plt
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
        added += ["matplotlib", "matplotlib.pyplot"]
    try:
        import pykalman  # noqa: F401
    except Exception:
        pk = types.ModuleType("pykalman")
        pk.KalmanFilter = KalmanFilter
        sys.modules["pykalman"] = pk
        added.append("pykalman")
    return added
