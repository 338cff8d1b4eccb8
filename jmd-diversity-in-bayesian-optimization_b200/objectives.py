"""Drop-in for the reference's ``Src/objectives.py`` (Wildcat Wells part).

``objectives`` keeps the reference's attributes (args, bounds, seed, fun_name, minstate, budget,
parallel, optimal_y, dim) and methods ``construct_X``, ``wildcatwells`` and ``generate_cont``. The
10 201 x 3 pure-Python simplex-noise calls of the reference (objectives.py:199-207) run as one CUDA
kernel (csrc/wildcat.cu); only the seed hashing, the 4-entry gradient-table shuffle and the peak draw
stay on the host, with the reference's own generators.

Additions: ``wildcatwells_batch`` builds many surfaces in one launch; the callable returned by
``generate_cont`` exposes ``grid_table(lattice)`` so the BO engine can read the objective at every
candidate without leaving the device.  Rastrigin / Sphere / Rosenbrock are out of scope.
"""
import numpy as np
import torch

from b200bo import host, ops


class WildcatFunction:
    """Callable f(x) over one surface: exact node values, piecewise-linear on the triangles of scipy's Delaunay
    (Qhull) triangulation of the 101 x 101 nodes (per-cell diagonal from ops.qhull_diag_mask), NaN outside
    [0,100]^2 — the scipy LinearNDInterpolator the reference returns (objectives.py:349-350), bit-identical on
    half-node queries. Accepts numpy arrays, lists or torch tensors of shape (m,2) / (2,)."""

    def __init__(self, X, surf, device=None):
        self.X = X
        self.surf = np.asarray(surf, dtype=np.float32)
        self.device = torch.device(device or "cuda:0")
        self._dev = None

    @property
    def surf_device(self):
        if self._dev is None:
            self._dev = torch.as_tensor(self.surf, device=self.device).reshape(1, 101, 101).contiguous()
        return self._dev

    def __call__(self, x, *more):
        if more:                                        # f(x, y) call style
            x = np.stack([np.asarray(x, dtype=np.float64)] + [np.asarray(m, dtype=np.float64) for m in more], axis=-1)
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        pts = np.asarray(x, dtype=np.float64).reshape(-1, 2)
        out = ops.wildcat_eval(self.surf_device, torch.as_tensor(pts, device=self.device))
        return out.cpu().numpy()

    def grid_table(self, lattice: ops.Lattice) -> torch.Tensor:
        """(G,) f32 objective at every lattice candidate, on device."""
        pts = lattice.points(self.device)
        return ops.wildcat_eval(self.surf_device, pts).to(torch.float32)


class objectives:
    def __init__(self):
        self.args = None
        self.bounds = None
        self.seed = None
        self._dim = None
        self._fun_name = "wildcatwells"
        self.noise_based_funcs = ["wildcatwells"]
        self.minstate = False
        self.budget = None
        self.parallel = False
        self.optimal_y = 100
        self.device = "cuda:0"

    @property
    def fun_name(self):
        return self._fun_name

    @fun_name.setter
    def fun_name(self, func_name):
        if func_name == "wildcatwells":
            self.minstate = False
            self._fun_name = func_name
            self.optimal_y = 100
        elif func_name in ("Sphere", "Rastrigin", "Rosenbrock"):
            raise NotImplementedError(f"{func_name}: closed-form N-D objectives are outside this engine's hot path")
        else:
            print(f"Choose a different function_name {func_name} is not available. \n Available functions are: wildcatwells")

    @property
    def dim(self):
        if self.bounds is None:
            raise Exception("Specify bounds for each dimension")
        if self._dim is None:
            self._dim = len(self.bounds)
        return self._dim

    def construct_X(self, with_noise_coords=False):
        """Node coordinates (objectives.py:85-104): (101,101,2) f32 for the default budget, X[i,j] = (j,i);
        optionally the noise coordinates arange(0,100,100/101) per axis."""
        budget = 10000 if self.budget is None else self.budget
        axes, _ = host.axis_grid(self.bounds, budget, endpoint=True, dtype=np.float32)
        X = np.stack(np.meshgrid(*axes), axis=-1)
        if with_noise_coords:
            naxes = [np.arange(0, 100, 100 / len(a)) for a in axes]
            return X, np.stack(np.meshgrid(*naxes), axis=-1)
        return X

    def _check_supported(self):
        if self.bounds is None or len(self.bounds) != 2 or any(tuple(b) != (0, 100) for b in self.bounds):
            raise NotImplementedError("the Wildcat Wells kernel covers the reference's 2-D [(0,100),(0,100)] domain")
        if self.args is None or int(self.args.get("N", 1)) != 1:
            raise NotImplementedError("only N = 1 peak is supported (every shipped driver uses N = 1)")
        if self.budget is not None:
            raise TypeError("wildcatwells() with budget set fails in the reference too (objectives.py:171)")

    def wildcatwells(self):
        """(X (101,101,2) f32, surf (101,101) f32) for args {N:1, Smoothness, rug_freq, rug_amp}
        (objectives.py:107-218); surf.max() == 100."""
        self._check_supported()
        surf = self.wildcatwells_batch([self.seed], [self.args["Smoothness"]], [self.args["rug_amp"]])
        self.minstate = False
        return self.construct_X(), surf[0].cpu().numpy()

    def wildcatwells_batch(self, seeds, smoothness, rug_amp, device=None) -> torch.Tensor:
        """(S,101,101) f32 surfaces on device for S (seed, smoothness, rug_amp) triples in one launch."""
        dev = torch.device(device or self.device)
        S = len(seeds)
        smoothness = np.broadcast_to(np.asarray(smoothness, dtype=np.float64), (S,))
        rug_amp = np.broadcast_to(np.asarray(rug_amp, dtype=np.float64), (S,))
        hs = np.zeros(S, dtype=np.uint64)
        vecs = np.zeros((S, 4, 2), dtype=np.int8)
        peaks = np.zeros((S, 2), dtype=np.float64)
        cache = {}
        for i, sd in enumerate(seeds):
            if sd not in cache:
                h, v = host.simplex_setup(sd)
                cache[sd] = (h, v, host.wildcat_peak(sd))
            h, v, pk = cache[sd]
            hs[i] = h
            vecs[i] = v
            peaks[i] = pk
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
        return ops.wildcat_surface(t(hs.view(np.int64)), t(vecs), t(peaks), t(smoothness.copy()), t(rug_amp.copy()))

    def generate_cont(self, from_saved=True, local_dir=None, budget=10000):
        """Continuous objective (objectives.py:322-351)."""
        if self.fun_name not in self.noise_based_funcs:
            raise NotImplementedError(self.fun_name)
        if from_saved and local_dir:
            raise NotImplementedError("CSV surface cache (Windows paths) is outside the hot path; surfaces are regenerated")
        X, surf = self.wildcatwells()
        return WildcatFunction(X, surf, device=self.device)
