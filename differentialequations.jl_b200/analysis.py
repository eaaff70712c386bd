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
    return _sq3(sim, _meancov(_arr(sim)[:, _step(sim, i)], _arr(sim)[:, _step(sim, j)]))


def timestep_meancor(sim, i, j):
    return _sq3(sim, _meancor(_arr(sim)[:, _step(sim, i)], _arr(sim)[:, _step(sim, j)]))


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


def timeseries_point_median(sim, ts):
    return np.stack([timepoint_median(sim, t) for t in ts])


def timeseries_point_quantile(sim, q, ts):
    return np.stack([timepoint_quantile(sim, q, t) for t in ts])


def timeseries_point_meanvar(sim, ts):
    mv = [timepoint_meanvar(sim, t) for t in ts]
    return np.stack([m for m, _ in mv]), np.stack([v for _, v in mv])


# ---- covariance / correlation between two save points or times, and the weighted variants --------------------------
def _meancov(a, b):
    ma, mb = a.mean(axis=0), b.mean(axis=0)
    return ma, mb, ((a - ma) * (b - mb)).sum(axis=0) / max(a.shape[0] - 1, 1)


def _meancor(a, b):
    ma, mb, cov = _meancov(a, b)
    return ma, mb, cov / np.sqrt(a.var(axis=0, ddof=1) * b.var(axis=0, ddof=1))


def _weighted_meancov(a, b, W, weight_type="reliability"):
    """Componentwise weighted means and covariance (upstream componentwise_weighted_meancov): weights W[k] per
    trajectory; normalisation 'reliability' (default) sum_w / (sum_w^2 - sum_w2), 'frequency' 1 / (sum_w - 1),
    anything else 1 / sum_w."""
    w = np.asarray(W, dtype=np.float64)
    if w.shape != (a.shape[0],):
        raise ValueError(f"need one weight per trajectory ({a.shape[0]}), got shape {w.shape}")
    sw, sw2 = w.sum(), (w * w).sum()
    ma, mb = (w[:, None] * a).sum(axis=0) / sw, (w[:, None] * b).sum(axis=0) / sw
    c = (w[:, None] * (a - ma) * (b - mb)).sum(axis=0)
    if weight_type == "reliability":
        c = c * (sw / (sw * sw - sw2))
    elif weight_type == "frequency":
        c = c / (sw - 1.0)
    else:
        c = c / sw
    return ma, mb, c


def _sq3(sim, t3):
    return tuple(_squeeze(sim, x) for x in t3)


def componentwise_vectors_timepoint(sim, t):
    a = _at_time(sim, t)
    return [a[:, k] for k in range(a.shape[1])]


def timestep_weighted_meancov(sim, W, i, j, weight_type="reliability"):
    return _sq3(sim, _weighted_meancov(_arr(sim)[:, _step(sim, i)], _arr(sim)[:, _step(sim, j)], W, weight_type))


def timepoint_meancov(sim, t1, t2):
    return _sq3(sim, _meancov(_at_time(sim, t1), _at_time(sim, t2)))


def timepoint_meancor(sim, t1, t2):
    return _sq3(sim, _meancor(_at_time(sim, t1), _at_time(sim, t2)))


def timepoint_weighted_meancov(sim, W, t1, t2, weight_type="reliability"):
    return _sq3(sim, _weighted_meancov(_at_time(sim, t1), _at_time(sim, t2), W, weight_type))


def _steps(sim):
    return range(1, _arr(sim).shape[1] + 1)


def timeseries_steps_meancov(sim):
    """Matrix [i][j] of timestep_meancov(sim, i, j) over all pairs of save points (upstream returns the same matrix of
    (mean_i, mean_j, cov) tuples)."""
    return [[timestep_meancov(sim, i, j) for j in _steps(sim)] for i in _steps(sim)]


def timeseries_steps_meancor(sim):
    return [[timestep_meancor(sim, i, j) for j in _steps(sim)] for i in _steps(sim)]


def timeseries_steps_weighted_meancov(sim, W, weight_type="reliability"):
    return [[timestep_weighted_meancov(sim, W, i, j, weight_type) for j in _steps(sim)] for i in _steps(sim)]


def timeseries_point_meancov(sim, ts1, ts2):
    return [[timepoint_meancov(sim, t1, t2) for t2 in ts2] for t1 in ts1]


def timeseries_point_meancor(sim, ts1, ts2):
    return [[timepoint_meancor(sim, t1, t2) for t2 in ts2] for t1 in ts1]


def timeseries_point_weighted_meancov(sim, W, ts1, ts2, weight_type="reliability"):
    return [[timepoint_weighted_meancov(sim, W, t1, t2, weight_type) for t2 in ts2] for t1 in ts1]


class HostEnsembleSummary:
    """EnsembleSummary(sim, t = sim.t; quantiles = [0.05, 0.95]) (qa.jl:54) built from gathered trajectories: mean `u`,
    sample variance `v`, median `med` and the quantile band `qlow` / `qhigh` per time point, each [len(t), n_state].
    (solve(...; summary=True) returns the device-reduced mean/variance summary without gathering anything.)"""

    def __init__(self, sim, t=None, quantiles=(0.05, 0.95)):
        on_grid = t is None
        self.t = np.asarray(sim.t if on_grid else t, dtype=np.float64)
        if on_grid:
            self.u, self.v = timeseries_steps_meanvar(sim)
            self.med = timeseries_steps_median(sim)
            self.qlow, self.qhigh = (timeseries_steps_quantile(sim, q) for q in quantiles)
        else:
            self.u, self.v = timeseries_point_meanvar(sim, self.t)
            self.med = timeseries_point_median(sim, self.t)
            self.qlow, self.qhigh = (timeseries_point_quantile(sim, q, self.t) for q in quantiles)
        self.num_monte = len(sim)
        self.elapsedTime = getattr(sim, "elapsedTime", 0.0)
        self.converged = getattr(sim, "converged", True)
