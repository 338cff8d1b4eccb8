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


class hp_series(hp_result):
    """Series of hp_result over the iterations of one run (results.py:247-446, data part only)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.data = []
        self.optdict = None

    def updateres(self):
        merged = {}
        for r in self.data:
            for name, vals in r.hyperparameters.items():
                merged.setdefault(name, []).extend(vals)
        self.hyperparameters = {k: np.asarray(v) for k, v in merged.items()}


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
