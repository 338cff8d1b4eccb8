"""Result containers with the reference's field names (``Src/results.py``), so pickled results and
driver code written against the reference keep working.  Only the parts the BO loop fills are kept
(SURVEY.md §2 row 7): ``result_opt``, ``hp_result``, ``hp_series``, ``result_diversity_trial``.
"""
import random
import string

import numpy as np


def _key():
    return "".join(random.choice(string.ascii_letters) for _ in range(8))


class result_opt:
    """One optimisation run (results.py:94-148). ``yall`` is the best-so-far curve of length
    max_iter + 2, ``xall`` the queried points, ``nit`` the iterations run."""

    def __init__(self):
        self.success = None
        self.nit = None
        self.xopt = None
        self.yopt = None
        self.minstate = None
        self._opt = None
        self.xall = []
        self.yall = []
        self.traindata = []
        self.hpdata = None
        self.status = 0
        self.key = _key()

    def __repr__(self):
        return str(self.__dict__)

    @property
    def opt_gap(self):
        """results.py:131-138."""
        if self._opt is None:
            raise Exception("Enter a thoeretical min/max for this result in self._opt")
        y = np.asarray(self.yall)
        return y - self._opt if self.minstate else self._opt - y

    def cum_opt_gap(self):
        """results.py:143-145: trapezoid area of opt_gap[1:]."""
        from utils import cumoptgap
        return cumoptgap(self.opt_gap[1:])


class hp_result:
    """Hyper-parameters of one fitted model (results.py:193-244); any object with
    ``named_parameters()`` is accepted (the reference insists on a botorch SingleTaskGP)."""

    def __init__(self, *args, **kwargs):
        self.hyperparameters = {}

    def addhp(self, model):
        for name, param in model.named_parameters():
            self.hyperparameters.setdefault(name, []).append(param.item())


def reject_outliers(data, m=4):
    """Keep the values within m population standard deviations of the mean (results.py:57-58)."""
    data = np.asarray(data)
    return data[abs(data - np.mean(data)) < m * np.std(data)]


def num_der(x, y):
    """Forward differences dy/dx, one value shorter than the input (results.py:89-90)."""
    x = np.asarray(x); y = np.asarray(y)
    return (y[1:] - y[:-1]) / (x[1:] - x[:-1])


def _flat_mask(peak, x, dens, peak_locs, biggest):
    """Where the density is numerically flat (results.py:61-70): the squared slope, rounded to the number of decimals
    of the squared slope at the peak itself (one fewer for secondary peaks), is zero; peaks are never flat."""
    slope2 = num_der(x, dens) ** 2
    digits = np.log(slope2[peak]) / np.log(10)
    digits = -int(digits) if biggest else -int(digits + 1)
    flat = np.append(np.round(slope2, digits) == 0, True)
    flat[peak_locs] = False
    return flat


def peak_extremes(peak_arg_loc, peak_num, x, density_function, peak_locs):
    """(start, end) indices of the hump around one KDE peak (results.py:73-86): the nearest flat positions at least
    two samples away on either side, else the ends of the mesh."""
    dens = density_function(x) if callable(density_function) else np.asarray(density_function)
    flat = _flat_mask(peak_arg_loc, x, dens, peak_locs, biggest=(peak_num == 0))
    after = np.nonzero(flat[peak_arg_loc + 2:])[0]
    end = peak_arg_loc + 2 + after[0] if len(after) and peak_arg_loc + 2 < len(x) else len(x)
    before = np.nonzero(flat[:max(peak_arg_loc - 2, 0)])[0]
    start = before[-1] if len(before) else 0
    return (int(start), int(end))


class hp_series(hp_result):
    """Series of hp_result over the iterations / training-set sizes of one run (results.py:247-446, data part):
    ``data`` is a list of hp_result until ``updateres()`` turns it into {parameter name: array}, and
    ``extract_hyperparam()`` finds the modes of each parameter's distribution (the "fixed" hyper-parameters that
    Experiment3-2-3 feeds to BOloop_fixparam)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.data = []
        self.key = _key()
        self.optdict = None
        self.opthp_alldata = None
        self.seed = None
        self.wwargs = None

    def updateres(self):
        """results.py:322-340. Failed fits (``hyperparameters is None``) take the values of the nearest successful
        entry, as ``replacetool(..., method='nearest-element')`` does."""
        if self.data is None:
            raise Exception("Data not collected yet")
        if isinstance(self.data, dict):
            return
        hps = [r.hyperparameters for r in self.data]
        good = [i for i, h in enumerate(hps) if h is not None]
        if not good:
            self.data = {}
            self.hyperparameters = {}
            return
        filled = [hps[min(good, key=lambda g: abs(g - i))] if h is None else h for i, h in enumerate(hps)]
        names = [k for k in filled[0] if all(k in h for h in filled)]
        merged = {k: np.array([h[k] for h in filled]).flatten() for k in names}
        self.data = merged
        self.hyperparameters = merged

    def getkeys(self):
        return list(self.data.keys())

    def extract_hyperparam(self, npoints=1000, rel_area_factor=0.333):
        """Modes of every hyper-parameter's distribution (results.py:342-392): jitter the values with N(0, 2e-3) noise
        (numpy's global generator, as the reference), drop 4-sigma outliers, Gaussian KDE on a 1000-point mesh over
        mean +- 4 sd, peaks sorted by height, each peak's hump delimited by flat density, and every peak whose hump
        area exceeds a third of the biggest hump's is kept. ``optdict[name]`` = peak locations, biggest first."""
        from scipy.signal import find_peaks
        from scipy.stats import gaussian_kde
        self.optdict = {}
        self.opthp_alldata = {}
        for col in self.data.keys():
            vals = np.asarray(self.data[col], dtype=np.float64)
            sample = reject_outliers(vals + np.random.normal(0, 2e-3, len(vals)))
            kde = gaussian_kde(sample)
            sd, mu = np.std(sample), np.mean(sample)
            x = np.arange(mu - 4 * sd, mu + 4 * sd, 8 * sd / npoints)
            dens = kde(x)
            peaks = find_peaks(dens)[0]
            peaks = peaks[np.argsort(dens[peaks])][::-1]
            humps = [peak_extremes(p, i, x, dens, peaks) for i, p in enumerate(peaks)]
            areas = np.array([np.sum((dens[a:b][1:] + dens[a:b][:-1]) * 0.5) for a, b in humps])
            keep = peaks[areas > rel_area_factor * areas[np.argmax(areas)]]
            self.opthp_alldata[col] = {"all_peaks": peaks, "peak_extremes": humps}
            self.optdict[col] = x[keep]
            if col == "mean_module.raw_constant":            # newer / older gpytorch names of the same parameter
                self.optdict["mean_module.constant"] = x[keep]
            elif col == "mean_module.constant":
                self.optdict["mean_module.raw_constant"] = x[keep]


class result_diversity_trial:
    """Results grouped by diversity percentile (results.py:681-861, container part only)."""

    def __init__(self):
        self.percentiles = []
        self.data = {}
        self.key = _key()

    def addresult(self, percentile, current_result):
        self.data.setdefault(percentile, {})[current_result.key] = current_result

    def yall(self, percentile):
        return np.stack([np.asarray(r.yall) for r in self.data[percentile].values()])

    def cum_opt_gap(self, percentile):
        return np.array([r.cum_opt_gap() for r in self.data[percentile].values()])

    def bootstrap(self, percentile):
        """Bootstrap mean / 2.5 / 97.5 band of the optimality-gap curves of one percentile (results.py:787-794)."""
        from utils import bootstrap
        return bootstrap([obj.opt_gap for obj in self.data[percentile].values()])

    def bootstrap_cumoptgap(self, percentile):
        """Same for the cumulative optimality gaps (results.py:796-802)."""
        from utils import bootstrap
        return bootstrap([obj.cum_opt_gap() for obj in self.data[percentile].values()])
