"""EnsembleAnalysis (re-exported by the reference: /root/reference/test/qa/qa.jl:211) -- statistics over the
trajectories of an EnsembleSolution, with upstream's function names.  "timestep" functions take the index of a save
point (1-based like Julia), "timepoint" functions a time; "timeseries_steps_*" return one value per save point.

They work on the gathered arrays of an EnsembleSolution ([N, n_save, n_state], host numpy).  When only mean and
variance are needed, prefer solve(...; summary=True): the device reduces them and the trajectories never leave the GPU
(b200ens_solve_moments)."""
import numpy as np


def _arr(sim):
    a = getattr(sim, "u_array", None)
    if a is None:
        raise TypeError("EnsembleAnalysis needs an EnsembleSolution with gathered arrays (no output_func/reduction)")
    if np.ndim(sim.t) != 1:
        raise TypeError("EnsembleAnalysis needs a common time grid: solve the ensemble with saveat (not save_everystep)")
    return a.astype(np.float64, copy=False)


def _squeeze(sim, v):
    return v[..., 0] if getattr(sim, "_scalar", False) else v


def _step(sim, i):
    n = _arr(sim).shape[1]
    if not 1 <= int(i) <= n:
        raise IndexError(f"timestep {i} outside 1..{n}")
    return int(i) - 1


def get_timestep(sim, i):
    """Iterator over the states of all trajectories at save point i."""
    return iter(_squeeze(sim, _arr(sim)[:, _step(sim, i)]))


def componentwise_vectors_timestep(sim, i):
    a = _arr(sim)[:, _step(sim, i)]
    return [a[:, k] for k in range(a.shape[1])]


def timestep_mean(sim, i):
    return _squeeze(sim, _arr(sim)[:, _step(sim, i)].mean(axis=0))


def timestep_median(sim, i):
    return _squeeze(sim, np.median(_arr(sim)[:, _step(sim, i)], axis=0))


def timestep_quantile(sim, q, i):
    return _squeeze(sim, np.quantile(_arr(sim)[:, _step(sim, i)], q, axis=0))


def timestep_meanvar(sim, i):
    a = _arr(sim)[:, _step(sim, i)]
    return _squeeze(sim, a.mean(axis=0)), _squeeze(sim, a.var(axis=0, ddof=1))


def timestep_meancov(sim, i, j):
    a, b = _arr(sim)[:, _step(sim, i)], _arr(sim)[:, _step(sim, j)]
    ma, mb = a.mean(axis=0), b.mean(axis=0)
    cov = ((a - ma) * (b - mb)).sum(axis=0) / max(a.shape[0] - 1, 1)
    return _squeeze(sim, ma), _squeeze(sim, mb), _squeeze(sim, cov)


def timestep_meancor(sim, i, j):
    ma, mb, cov = timestep_meancov(sim, i, j)
    sa = np.sqrt(_arr(sim)[:, _step(sim, i)].var(axis=0, ddof=1))
    sb = np.sqrt(_arr(sim)[:, _step(sim, j)].var(axis=0, ddof=1))
    return ma, mb, cov / _squeeze(sim, sa * sb)


def timeseries_steps_mean(sim):
    return _squeeze(sim, _arr(sim).mean(axis=0))


def timeseries_steps_median(sim):
    return _squeeze(sim, np.median(_arr(sim), axis=0))


def timeseries_steps_quantile(sim, q):
    return _squeeze(sim, np.quantile(_arr(sim), q, axis=0))


def timeseries_steps_meanvar(sim):
    a = _arr(sim)
    return _squeeze(sim, a.mean(axis=0)), _squeeze(sim, a.var(axis=0, ddof=1))


def _point_index(sim, t):
    k = np.nonzero(np.asarray(sim.t, dtype=np.float64) == float(t))[0]
    return int(k[0]) + 1 if k.size else None


def _at_time(sim, t):
    """[N, n_state] at time t: a save point, or through every trajectory's dense output (solve(...; dense=True))."""
    k = _point_index(sim, t)
    if k is not None:
        return _arr(sim)[:, k - 1]
    if getattr(sim, "_dense", None) is None:
        raise ValueError(f"t = {t} is not a save point and the ensemble was solved without dense=True")
    return np.stack([sim._dense(i, np.array([float(t)]))[0] for i in range(len(sim))])


def get_timepoint(sim, t):
    return iter(_squeeze(sim, _at_time(sim, t)))


def timepoint_mean(sim, t):
    return _squeeze(sim, _at_time(sim, t).mean(axis=0))


def timepoint_median(sim, t):
    return _squeeze(sim, np.median(_at_time(sim, t), axis=0))


def timepoint_quantile(sim, q, t):
    return _squeeze(sim, np.quantile(_at_time(sim, t), q, axis=0))


def timepoint_meanvar(sim, t):
    a = _at_time(sim, t)
    return _squeeze(sim, a.mean(axis=0)), _squeeze(sim, a.var(axis=0, ddof=1))


def timeseries_point_mean(sim, ts):
    return np.stack([timepoint_mean(sim, t) for t in ts])


def timeseries_point_meanvar(sim, ts):
    mv = [timepoint_meanvar(sim, t) for t in ts]
    return np.stack([m for m, _ in mv]), np.stack([v for _, v in mv])
