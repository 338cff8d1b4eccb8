"""Drop-in for the reference's ``Src/Optimizers.py``: the ``optimizer`` facade the drivers configure
(Scripts/Experiment2.py:214-225) and call as ``opt.optimize(f)``.

Attributes and their meaning are the reference's (Optimizers.py:9-26). Engine-only knobs live in
``options``: 'budget' (candidate lattice size, default 10 000 = 100x100), 'device'.
The scipy baselines CG / NM / SA (Optimizers.py:69-166) are outside the hot path and not provided.
"""
import numpy as np

from BO import BOloop, BOloop_fixparam


class optimizer:
    def __init__(self):
        self.max_iter = 100
        self.obj_func = None
        self.train = None
        self.opt = "BO"
        self.options = {"tol": 0.1, "nu": 2.5, "verbose": False, "budget": 10000, "device": "cuda:0"}
        self.optima = 100
        self.minimize = False
        self.paramset = None
        self.datagenmodule = None
        self.funcseed = None
        self.bounds = [(0, 100), (0, 100)]

    def optimize(self, f, method=None):
        if method is None:
            method = self.opt
        table = {"BO": self.BO, "BO-fixparam": self.BO_fixparam}
        if method not in table:
            raise Exception("Method {} is not available at this point in time, the current options are: {}".format(
                method, list(table.keys())))
        return table[method](f)

    def _prep(self, f):
        self.nu = 2.5
        self.covar = "matern"
        X = self.train
        f_c = (lambda x: -f(x)) if self.minimize else f
        if self.minimize and hasattr(f, "grid_table"):
            f_c.grid_table = lambda lat: -f.grid_table(lat)
        Y = np.asarray(f_c(X)).astype(np.float32)                 # Optimizers.py:52,66
        return X, Y, f_c

    def BO_fixparam(self, f):
        X, Y, f_c = self._prep(f)
        return BOloop_fixparam(f_c, self.max_iter, X, Y, self.paramset, bounds=self.bounds, tol=self.options["tol"],
                               nu=self.options["nu"], minstate=self.minimize, optimal_y=self.optima,
                               budget=self.options.get("budget", 10000), device=self.options.get("device", "cuda:0"))

    def BO(self, f):
        # the reference forgets to pass options['nu'] here (Optimizers.py:67) so Experiment3-1's nu sweep
        # silently runs nu = 2.5; the engine honours the option
        X, Y, f_c = self._prep(f)
        return BOloop(f_c, self.max_iter, X, Y, bounds=self.bounds, tol=self.options["tol"],
                      verbose=self.options["verbose"], nu=self.options["nu"], minstate=self.minimize,
                      optimal_y=self.optima, budget=self.options.get("budget", 10000),
                      device=self.options.get("device", "cuda:0"))
