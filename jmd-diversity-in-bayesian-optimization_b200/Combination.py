"""Drop-in for the reference's ``Src/Combination.py``: lexicographic un-ranking of k-subsets.

``combination_finder.find_elm(n, k, m)`` keeps the reference signature (Combination.py:131-170).
Single ranks are un-ranked on the host with Python integers; ``find_elms`` un-ranks a whole batch
on the GPU (128-bit kernel, include/b200bo.h: b200bo_unrank_u128) when C(n,k) < 2^128.
"""
import math

from b200bo import host


class combination_finder:
    def __init__(self):
        self.explored_combinations_real = dict()

    def reduced_comb(self, n, k):
        key = (n, k)
        if key not in self.explored_combinations_real:
            self.explored_combinations_real[key] = math.comb(n, k)
        return self.explored_combinations_real[key]

    def LargestV(self, a, b, x):
        """Largest v < a with C(v,b) <= x (Combination.py:95-117)."""
        v = a - 1
        while self.reduced_comb(v, b) > x:
            v -= 1
        return v

    LargestV_binary = LargestV      # same value, the reference only differs in search order (:11-54)

    def find_elm(self, n, k, m):
        return host.unrank_host(n, k, m)

    def find_elms(self, n, k, ranks, device=None):
        """Batch un-ranking -> (N,k) int32 tensor on ``device`` (default cuda:0)."""
        import torch
        from b200bo import ops
        dev = torch.device(device or "cuda:0")
        if math.comb(n, k) < 1 << 128:
            return ops.unrank(n, k, ranks, dev)
        return torch.tensor([host.unrank_host(n, k, int(r)) for r in ranks], dtype=torch.int32, device=dev)
