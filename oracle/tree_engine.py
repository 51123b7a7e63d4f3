"""Checker-side engine for poyd_b200.tree: the three batch calls answered by the CPU checker (compiled reference /
port).  Test infrastructure like everything under oracle/: used by tests/ and by bench.py's cpu_baseline leg only, never
imported by the package."""
import numpy as np

from oracle import oracle
from poyd_b200 import sequence as S


class OracleEngine:
    def __init__(self, cm, nthreads: int = 8):
        self.cm = cm
        self.chk = oracle.best_checker(cm)
        self.nthreads = nthreads
        self.calls = 0
        self.pairs = 0
        self.log = []

    def close(self):
        pass

    def _pool(self, store, pairs):
        used, inv = np.unique(pairs.reshape(-1), return_inverse=True)
        return S.SeqPool([store[i] for i in used]), inv.reshape(-1, 2).astype(np.int32)

    def _deltaw(self, pool, pp, hint=None):
        cnt = pool.count(self.cm.gap)
        la, lb = pool.len[pp[:, 0]].astype(np.int64), pool.len[pp[:, 1]].astype(np.int64)
        gaps = np.maximum(cnt[pp[:, 0]], cnt[pp[:, 1]])
        return (gaps + S.deltaw_calc(np.maximum(la, lb), np.minimum(la, lb), hint)).astype(np.int32)

    def _run(self, mode, store, pairs, hint=None):
        pool, pp = self._pool(store, pairs)
        self.calls += 1
        self.pairs += len(pp)
        dw = None if self.cm.cost_model_type == 1 else self._deltaw(pool, pp, hint)
        self.log.append((str(mode), pool.len[pp[:, 0]].copy(), pool.len[pp[:, 1]].copy(), dw))
        return self.chk.batch(mode, pool.pool, pool.off, pool.len, pp, deltaw=dw, nthreads=self.nthreads)

    def median(self, store, pairs):
        o = self._run(3 if self.cm.cost_model_type == 1 else 1, store, pairs)
        return o["cost"], [o["median"][p, :o["lens"][p, 0]].copy() for p in range(len(pairs))]

    def align(self, store, pairs):
        o = self._run(3 if self.cm.cost_model_type == 1 else 1, store, pairs)
        return [(o["ra"][p, :o["lens"][p, 2]].copy(), o["rb"][p, :o["lens"][p, 3]].copy()) for p in range(len(pairs))]

    def distance(self, store, pairs):
        pool, pp = self._pool(store, pairs)
        la, lb = pool.len[pp[:, 0]].astype(np.int64), pool.len[pp[:, 1]].astype(np.int64)
        hint = np.maximum(np.abs(la - lb), 8)
        return self._run(2 if self.cm.cost_model_type == 1 else 0, store, pairs, hint)["cost"]
