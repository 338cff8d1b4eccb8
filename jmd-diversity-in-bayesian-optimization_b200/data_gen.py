"""Drop-in for the reference's ``Src/data_gen.py`` (ranked-DPP "subDPP" sampler).

``diverse_data_generator`` keeps the reference's attributes (X, gamma, training_size, options, data,
key, acquiredsets, mean, sd, tracker, rng) and methods (constructX, generate, sample, twoD_score,
compare_subDPP, iid_comb_locs).  What changes underneath:

* the element ranks are still drawn by CPython's / numpy's generators exactly as the reference does
  (data_gen.py:853-880) — they decide which sets exist, so parity of *chosen sets* depends on them;
* un-ranking (93 % of the reference's wall time, Combination.py) runs on the GPU in 128-bit integers;
* scoring  log det exp(-gamma D)  (data_gen.py:484-497), the mean / std / z-score block (:406-408) run in
  CUDA kernels (csrc/dpp.cu); the argsort (:409) is a stable device sort.

Out of scope: the vanilla k-DPP eigendecomposition baseline and the on-disk CSV /
ray mode.
"""
import random
import string
import time

import numpy as np
import torch

from Combination import combination_finder
from b200bo import host, ops


def set_diff2d(A, B):
    """Rows of the larger of two (m,2) arrays that do not occur in the smaller one, in their original order
    (data_gen.py:59-66)."""
    A = np.asarray(A); B = np.asarray(B)
    if len(A) < len(B):
        A, B = B, A
    hit = (A[:, None] == B).all(-1).any(-1)
    return A[~hit]


class diverse_data_generator:
    def __init__(self):
        self.acquiredsets = []
        self.data = {}
        self.gamma = None
        self.X = None
        self._training_size = None
        self.rng = None
        self._key = None
        self._dim = None
        self.mean = None
        self.sd = None
        self.tracker = {"N_results": 0, "time-tracking": {}}
        self.options = {"GPU": True, "parallel-process": True, "seed": None, "N_samples": 10000, "bounds": None,
                        "time-debug": False, "local_dir": None}
        self.gamma_correlation_dict = {}
        self.combination_sampler = combination_finder()
        self.device = "cuda:0"
        self._X_dev = None
        self._lat = None
        self._lat_key = None
        self._tables = {}
        self._pos = self._pos_sorted = self._scores_dev = self._acq_mask = None
        self._acq_seen = 0

    # ------------------------------------------------------------------ bookkeeping (data_gen.py:110-171)
    @property
    def key(self):
        return self._key

    @key.setter
    def key(self, seed):
        if isinstance(seed, int):
            raise Exception("Seed for keygen can only be an int, not {}".format(type(seed)))
        random.seed(seed)
        self._key = "".join(random.choice(string.ascii_letters) for _ in range(8))
        self.create_data_dict()

    @property
    def training_size(self):
        return self._training_size

    @training_size.setter
    def training_size(self, value):
        self.gamma_correlation_dict = None
        self._training_size = int(value)

    @property
    def dim(self):
        if self.options["bounds"] is None:
            raise Exception("Specify bounds for each dimension")
        if self._dim is None:
            self._dim = len(self.options["bounds"])
        return self._dim

    def create_data_dict(self):
        self.data[self.key] = {"scores": None, "dist": None, "training_size": self.training_size,
                               "bounds": self.options["bounds"], "seed": self.options["seed"],
                               "N_samples": self.options["N_samples"], "gamma": self.gamma, "mean": None,
                               "standard-deviation": None}

    def update_data_dict(self, scored=False):
        self.data[self.key].update({"mean": self.mean, "standard-deviation": self.sd})

    def constructrng(self):
        if self.options["seed"] is None:
            print("There is no seed specified, so the generated results will be inconsistent")
        self.rng = {"sampler": np.random.default_rng(self.options["seed"])}

    def constructX(self, budget=10000):
        """Candidate pool (data_gen.py:294-303): (10000,2) f64 integer lattice on [0,99]^2, x fastest."""
        if self.options["bounds"] is None:
            raise Exception("Neither the sample space was given and can not be constructed because bounds have not been specified.")
        axes, _ = host.axis_grid(self.options["bounds"], budget, endpoint=False, dtype=np.float64)
        self.X = np.stack(np.meshgrid(*axes), axis=-1).reshape(-1, self.dim)
        self._X_dev = None

    def _pool_device(self):
        if self._X_dev is None or self._X_dev.shape[0] != self.X.shape[0]:
            self._X_dev = torch.as_tensor(np.ascontiguousarray(self.X, dtype=np.float64), device=self.device)
        return self._X_dev

    # ------------------------------------------------------------------ generate / sample (data_gen.py:305-470)
    def generate(self, method="subDPP"):
        if self.options["time-debug"]:
            self.start = time.perf_counter()
        self.acquiredsets = []
        self.key = None
        if self.rng is None:
            self.constructrng()
        if self.X is None:
            self.constructX()
        if method == "subDPP" and self.training_size is None or self.gamma is None:      # precedence as data_gen.py:327
            raise Exception("Training size or gamma has not been specified, so data can not be generated")
        if method != "subDPP":
            raise NotImplementedError("only the ranked-DPP ('subDPP') sampler is on the hot path")
        return self.sub_dpp()

    def sample(self, args=None, method="subDPP"):
        if method != "subDPP":
            raise NotImplementedError("only the ranked-DPP ('subDPP') sampler is on the hot path")
        return self.sub_dpp_sample(args)

    def iid_comb_locs(self, combination=None, sample_size=None):
        if combination is None:
            combination = self.X.shape[0], self.training_size
        if sample_size is None:
            sample_size = self.options["N_samples"]
        return host.iid_comb_locs(combination[0], combination[1], sample_size, self.options["seed"])

    def score_sets(self, idx: torch.Tensor, gamma=None) -> torch.Tensor:
        """log det for index sets idx (N,k) int32 on device -> (N,) f64 on device. On the constructX lattice the
        L-ensemble entries come from a per-gamma table (same values bit for bit, csrc/dpp.cu lattice variant)."""
        gamma = self.gamma if gamma is None else gamma
        pool = self._pool_device()
        if self._lat_key is not pool:
            self._lat = ops.dpp_lattice_of(self.X)
            self._lat_key = pool
            self._tables = {}
        if self._lat is not None and idx.shape[1] <= 12:
            key = float(gamma)
            if key not in self._tables:
                self._tables[key] = ops.dpp_table(self._lat, key, pool.device)
            return ops.dpp_logdet_score_lattice(self._lat, self._tables[key], idx)
        return ops.dpp_logdet_score(pool, idx, gamma)

    def sub_dpp(self):
        if self.gamma is None and self.training_size == 10:
            self.gamma = 1e-6
        elif self.gamma is None:
            raise Exception("self.gamma has not been fixed, look for a range of functional gammas using self.gammacalc")
        if self.options["local_dir"]:
            raise NotImplementedError("on-disk CSV mode is outside the hot path")
        if self.options["time-debug"]:
            t_setup = time.perf_counter()
        import math
        total = math.comb(self.X.shape[0], self.training_size)
        if np.iinfo(np.int64).max < total < 1 << 127:
            # the C(n,k) > int64 branch of iid_comb_locs (data_gen.py:865-871) entirely on the device: per-element
            # seeded Mersenne Twister + randint (csrc/mt_rank.cu), then the 128-bit un-ranker — same ranks bit for bit
            seeds = host.iid_comb_seeds(self.options["N_samples"], self.options["seed"])
            m_lo, m_hi = ops.seeded_randint_u128(seeds, total + 1, self.device)
            pos = ops.unrank_u128(self.X.shape[0], self.training_size, m_lo, m_hi)
        else:
            elements = self.iid_comb_locs()
            pos = self.combination_sampler.find_elms(self.X.shape[0], self.training_size, elements, device=self.device)
        if bool((pos < 0).any()):
            raise Exception("check previous")            # rank == C(n,k): the reference raises (Combination.py:150-153)
        if self.options["time-debug"]:
            torch.cuda.synchronize()
            t_datagen = time.perf_counter()
        pool = self._pool_device()
        if self.training_size > 1:
            logdet = self.score_sets(pos)
            stats, z = ops.dpp_zscore(logdet)
            # ascending (score, set index): data_gen.py:409 made deterministic — device radix sort + gather
            z_sorted, _, self._pos_sorted, dist = ops.dpp_rank(z, pos, pool)
            self.mean, self.sd = (float(v) for v in stats.cpu())
            self._pos, self._scores_dev = pos, z_sorted
            self._acq_mask, self._acq_seen = None, 0
            self.data[self.key].update({"scores": z_sorted.cpu().numpy(), "dist": dist.cpu().numpy()})
            self.update_data_dict()
        else:
            self.data[self.key].update({"scores": "Can not be calculated for just a sample_set of size 1.",
                                        "dist": pool[pos.long()].cpu().numpy()})
        self.tracker["N_results"] += 1
        if self.options["time-debug"]:
            torch.cuda.synchronize()
            t_score = time.perf_counter()
            total = t_score - self.start
            self.tracker["time-tracking"] = {"Total": total, "setting-up": (t_setup - self.start) / total,
                                             "sampling": (t_datagen - t_setup) / total,
                                             "scoring": (t_score - t_datagen) / total}
            del self.start

    @staticmethod
    def _window(percentile, n):
        """Rank window of one percentile (data_gen.py:445-450)."""
        if 2.5 < percentile < 97.5:
            return int((percentile - 2.5) / 100 * n), int(percentile / 100 * n)
        if percentile > 97.5:
            return int(95 / 100 * n), n
        return 0, int(5 / 100 * n)

    def _acquired_mask(self):
        """Device bit mask of self.acquiredsets (bit = rank), rebuilt whenever the public list was changed from outside."""
        n = self.options["N_samples"]
        if self._acq_mask is None or self._acq_seen != len(self.acquiredsets) or self._acq_mask.numel() != (n + 31) // 32:
            words = np.zeros((n + 31) // 32, dtype=np.uint32)
            for loc in self.acquiredsets:
                words[int(loc) >> 5] |= np.uint32(1) << np.uint32(int(loc) & 31)
            self._acq_mask = torch.as_tensor(words.view(np.int32), device=self.device)
            self._acq_seen = len(self.acquiredsets)
        return self._acq_mask

    def _pick(self, percentile):
        """One draw from a percentile's rank window minus the acquired ranks (data_gen.py:445-466). The reference draws
        Generator.choice(list(possible), 1) over possible = np.setdiff1d(window, acquiredsets): that is possible[r] with
        r the Generator's draw from range(len(possible)) — same stream position, same r; the r-th free rank of the window
        and its set are found on the device (csrc/dpp_rank.cu:window_pick_kernel)."""
        lo, hi = self._window(percentile, self.options["N_samples"])
        free = (hi - lo) - len({int(a) for a in self.acquiredsets if lo <= a < hi})
        try:
            r = int(self.rng["sampler"].choice(free, 1)[0])
        except ValueError:
            raise Exception("Constructed batch-DPP space exhausted above its {}th percentile. Re-generate the distribution "
                            "or sample from a smaller percentile.".format(percentile))
        loc, pts = ops.dpp_window_pick(self._acquired_mask(), self.options["N_samples"], lo, hi, r, self._pos_sorted,
                                       self._pool_device())
        self.acquiredsets.append(loc)
        self._acq_seen += 1
        return loc, pts

    def sub_dpp_sample(self, percentiles):
        try:
            self.data[self.key]
        except Exception:
            raise Exception("Data needs to be generated first using the generator method")
        if isinstance(percentiles, (list, np.ndarray)):
            return np.stack([self._pick(p)[1] for p in percentiles])
        return self._pick(percentiles)[1]

    # ------------------------------------------------------------------ scoring API (data_gen.py:472-497)
    def twoD_score(self, subset, gamma):
        """Diversity score of one (k,2) point set: log det exp(-gamma * squared distances)."""
        if not isinstance(subset, (list, np.ndarray)):
            raise TypeError("A list or np.array was expected but a {} was entered as the subset".format(type(subset)))
        first = subset[0][0]
        if isinstance(first, (list, np.ndarray)):
            raise ValueError("Elements of subset need to be a (numpy int,numpy int) but are ({},{})".format(type(first), type(first)))
        if not isinstance(first, (np.int32, np.float64, np.float32, np.int64)):
            raise TypeError("A list or np.array was expected but a {} was entered as the subset".format(type(subset)))
        pts = torch.as_tensor(np.asarray(subset, dtype=np.float64), device=self.device)
        k = pts.shape[0]
        idx = torch.arange(k, dtype=torch.int32, device=self.device).reshape(1, k)
        return float(ops.dpp_logdet_score(pts.contiguous(), idx, gamma)[0])

    def compare_subDPP(self, resultkey, subset):
        """Rank and z-score of an arbitrary set inside a stored distribution (data_gen.py:472-482)."""
        d = self.data[resultkey]
        score = self.twoD_score(subset, d["gamma"])
        z = (score - d["mean"]) / d["standard-deviation"]
        if resultkey == self.key and getattr(self, "_scores_dev", None) is not None:
            scores = self._scores_dev                                     # still resident from generate()
        else:
            scores = torch.as_tensor(np.ascontiguousarray(d["scores"], dtype=np.float64), device=self.device)
        return ops.dpp_count_less(scores, z), z

    # ------------------------------------------------------------------ gamma selection (data_gen.py:758-842)
    def gammacorrelation(self, g1, g2, min_samples):
        import scipy.stats
        self.options["N_samples"] = min_samples
        self.options["parallel-process"] = False
        self.gamma = 10 ** g1
        self.generate()
        score1 = self.data[self.key]["scores"]
        self.gamma = 10 ** g2
        s2 = self.score_sets(self._pos_sorted).cpu().numpy()
        s2 = (s2 - s2.mean()) / np.std(s2)
        x, y = np.argsort(score1), np.argsort(s2)
        if np.isnan(s2).any():
            np.random.shuffle(y)
        return scipy.stats.linregress(x, y).rvalue

    def _gamma_scores(self, exponents, min_samples):
        """What the 256 gammacorrelation calls of gammacalc recompute, once: the SAME min_samples sets (generate() with a
        fixed seed draws identical sets whatever gamma is) scored under every gamma of the range in one launch
        (csrc/dpp_rank.cu:dpp_multigamma_kernel), then per gamma the z-scores and the stable rank order on the device.
        Returns (raw (n_gamma, N) host log-dets, per-gamma sorted z-scores, per-gamma order)."""
        import scipy.stats  # noqa: F401  (linregress, used by the caller)
        self.options["N_samples"] = min_samples
        self.options["parallel-process"] = False
        self.gamma = 10.0 ** exponents[0]
        self.generate()
        pos, pool = self._pos, self._pool_device()
        raw = ops.dpp_logdet_score_multigamma(pool, pos, [10.0 ** g for g in exponents])
        zs, orders = [], []
        for row in raw:
            _, z = ops.dpp_zscore(row.contiguous())
            z_sorted, order, _, _ = ops.dpp_rank(z)
            zs.append(z_sorted.cpu().numpy())
            orders.append(order.cpu().numpy())
        self.gamma = 10.0 ** exponents[-1]                       # where the reference's last gammacorrelation leaves it
        return raw.cpu().numpy(), zs, orders

    @staticmethod
    def _gammacorrelation_batched(batched, i, j):
        """gammacorrelation(g_i, g_j) (data_gen.py:758-779) from the batched scores: same host arithmetic as there."""
        import scipy.stats
        raw, zs, orders = batched
        score1 = zs[i]
        s2 = raw[j][orders[i]]
        s2 = (s2 - s2.mean()) / np.std(s2)
        x, y = np.argsort(score1), np.argsort(s2)
        if np.isnan(s2).any():
            np.random.shuffle(y)
        return scipy.stats.linregress(x, y).rvalue

    def gammacalc(self, grange=None, plot=False, annotations=False, min_samples=200):
        if plot:
            raise NotImplementedError("plotting is outside the hot path")
        gamma_range = range(-12, 4) if grange is None else np.arange(*grange)
        if self.gamma_correlation_dict is None:
            self.gamma_correlation_dict = {}
        m = np.zeros((len(gamma_range), len(gamma_range)))
        batched = None
        for i, g1 in enumerate(gamma_range):
            for j, g2 in enumerate(gamma_range):
                try:
                    m[i, j] = self.gamma_correlation_dict[g1][g2]
                except KeyError:
                    if batched is None:
                        batched = self._gamma_scores(list(gamma_range), min_samples)
                    m[i, j] = self._gammacorrelation_batched(batched, i, j)
                    self.gamma_correlation_dict.setdefault(g1, {})[g2] = m[i, j]
                    self.gamma_correlation_dict.setdefault(g2, {})[g1] = m[i, j]
        np.fill_diagonal(m, 0)
        pairs, min_cor = [], 0.95
        while len(pairs) < 2:
            rows, cols = np.where(m > min_cor)
            pairs = np.stack([[gamma_range[i] for i in rows], [gamma_range[i] for i in cols]]).T
            min_cor -= 0.05
            if min_cor < 0.5:
                print("Only Poorly coorelated gammas could be found for the given range of gamma(s).")
                return
        return pairs[np.abs([np.abs(p)[0] - np.abs(p)[1] for p in pairs]).argmax()]


    def optgammarand(self, current_set, set_size):
        """`set_size` random pool points not in `current_set` (data_gen.py:844-851): the growing training sets of the
        hyper-parameter extraction experiment (Experiment3-2-1.py:296-305)."""
        random.seed(self.options["seed"])
        if self.X is None:
            self.constructX()
        return np.array(random.sample(list(set_diff2d(current_set, self.X)), set_size))


class random_data_generator:
    """Uniformly random k-subsets of the pool. The reference's drivers call
    ``data_gen.random_data_generator()`` (Experiment2.py:172) but the class is not defined in
    Src/data_gen.py; this provides the obvious behaviour behind the same generate()/sample() calls."""

    def __init__(self):
        self.options = {"seed": None, "bounds": None, "N_samples": 10000}
        self.training_size = None
        self.X = None
        self.acquiredsets = []
        self.rng = None

    def generate(self):
        if self.X is None:
            g = diverse_data_generator()
            g.options["bounds"] = self.options["bounds"]
            g.constructX()
            self.X = g.X
        self.rng = np.random.default_rng(self.options["seed"])

    def sample(self, args=None):
        if self.rng is None:
            self.generate()
        return self.X[self.rng.choice(self.X.shape[0], int(self.training_size), replace=False)]
