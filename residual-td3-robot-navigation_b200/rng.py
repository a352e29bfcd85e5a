"""Device bank of numpy-legacy MT19937 streams, one per env (include/rtd3.h, `rtd3_mt_bank`).

The reference seeds numpy's global legacy generator once (robot-learning.py:19) and every random
draw of Environment / Robot / ReplayBuffer comes from it.  Batched envs get one stream each, seeded
`seed + i`; a single-env object can instead mirror numpy's *global* stream (`sync_from_numpy` /
`sync_to_numpy`) so that `np.random.seed(s); Environment()` behaves exactly like the reference.
"""
import ctypes

import numpy as np
import torch

from . import _lib

MT_N = 624


class MtBank:
    def __init__(self, n, device):
        self.n = int(n)
        self.device = torch.device(device)
        self.mt = torch.zeros((MT_N, self.n), dtype=torch.int32, device=self.device)   # uint32 bit patterns
        self.pos = torch.full((self.n,), MT_N, dtype=torch.int32, device=self.device)
        self.has_gauss = torch.zeros((self.n,), dtype=torch.int32, device=self.device)
        self.gauss = torch.zeros((self.n,), dtype=torch.float64, device=self.device)
        self._struct = _lib.MtBankStruct(self.mt.data_ptr(), self.pos.data_ptr(), self.has_gauss.data_ptr(),
                                         self.gauss.data_ptr(), self.n)

    @property
    def ref(self):
        return ctypes.byref(self._struct)

    def seed(self, seeds):
        """np.random.seed(seeds[i]) per stream; `seeds` int or array of uint32 values."""
        if np.isscalar(seeds):
            seeds = np.asarray([seeds], dtype=np.int64).repeat(self.n) if self.n == 1 else int(seeds) + np.arange(self.n, dtype=np.int64)
        seeds = np.asarray(seeds, dtype=np.int64)
        if seeds.shape != (self.n,):
            raise ValueError("need %d seeds" % self.n)
        if (seeds < 0).any() or (seeds > 0xFFFFFFFF).any():
            raise ValueError("seeds must fit in 32 bits (numpy's integer-seed path)")
        s = torch.from_numpy(seeds.astype(np.uint32).view(np.int32)).to(self.device)
        _lib.check(_lib.lib().rtd3_mt_seed(self.ref, _lib.ptr(s), _lib.stream_ptr(self.device)), "mt_seed")
        return self

    def draw_u32(self, k):
        out = torch.empty((k, self.n), dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().rtd3_mt_draw_u32(self.ref, _lib.ptr(out), k, _lib.stream_ptr(self.device)), "mt_draw_u32")
        return out

    def draw_gauss(self, k, where=None, equals=0):
        """k legacy normals per stream `[k,n]`; `where` (int8 `[n]`) restricts the draw to the streams with where == equals (the
        others are not advanced and get zeros)."""
        out = torch.empty((k, self.n), dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().rtd3_mt_draw_gauss_where(self.ref, _lib.ptr(out), k, _lib.ptr(where), int(equals), _lib.stream_ptr(self.device)),
                   "mt_draw_gauss_where")
        return out

    # ---- numpy global-state mirroring (single stream) ----
    def sync_from_numpy(self, stream=0):
        name, keys, pos, has_gauss, cached = np.random.get_state()
        assert name == "MT19937"
        self.mt[:, stream] = torch.from_numpy(keys.astype(np.uint32).view(np.int32)).to(self.device)
        self.pos[stream] = int(pos)
        self.has_gauss[stream] = int(has_gauss)
        self.gauss[stream] = float(cached)

    def sync_to_numpy(self, stream=0):
        keys = self.mt[:, stream].cpu().numpy().view(np.uint32)
        np.random.set_state(("MT19937", keys, int(self.pos[stream].item()), int(self.has_gauss[stream].item()),
                             float(self.gauss[stream].item())))
