"""CandidateTable — the candidate-edge dictionary of AlgebraicConnectivityMaximization with a
columnar mirror (incremental graph maintenance).

The reference keeps `candidate_edges` as a plain dict `key4 -> EdgeInterRobot`
(cslam/algebraic_connectivity_maximization.py:58,150,174) and walks it edge by edge in Python
every time a selection runs (:312-335 rekey, :391-417 inclusion, :205-218 weights, :178-190
removal).  At a million candidates those walks cost seconds while the solver costs
milliseconds.  This mapping behaves like that dict for every caller (same keys, same values,
same insertion order, `len`, `in`, `.values()`, `del`, `.pop`) and additionally keeps the
five fields of every live edge in numpy columns, updated in O(1) per insert / removal, so a
selection can be set up with a handful of vectorised passes and bulk inserts never create
Python objects.

Invariant: live slots in ascending slot order == the dict's iteration order (an overwritten
key keeps its slot, a removed and re-inserted key goes to the end — exactly what a dict does).
"""
from collections.abc import MutableMapping

import numpy as np


class CandidateTable(MutableMapping):

    def __init__(self, edge_type, capacity=1024):
        self._edge_type = edge_type
        self._slot = {}                                   # key -> slot (insertion ordered)
        self._ends = np.empty((capacity, 4), dtype=np.int64)   # r0, k0, r1, k1 as spelled
        self._w = np.empty(capacity, dtype=np.float64)
        self._alive = np.zeros(capacity, dtype=bool)
        self._obj = [None] * capacity                     # materialised EdgeInterRobot or None
        self._top = 0                                     # slots handed out so far
        self._dead = 0

    # ------------------------------------------------------------------ storage
    def _reserve(self, extra):
        need = self._top + extra
        cap = len(self._w)
        if need <= cap:
            return
        if self._dead and need - self._dead <= cap // 2:
            self._compact()
            if self._top + extra <= cap:
                return
        while cap < need:
            cap *= 2
        self._ends = np.concatenate([self._ends, np.empty((cap - len(self._ends), 4), np.int64)])
        self._w = np.concatenate([self._w, np.empty(cap - len(self._w))])
        self._alive = np.concatenate([self._alive, np.zeros(cap - len(self._alive), bool)])
        self._obj.extend([None] * (cap - len(self._obj)))

    def _compact(self):
        """Squeeze the dead slots out, keeping the order."""
        live = np.flatnonzero(self._alive[:self._top])
        n = len(live)
        self._ends[:n] = self._ends[live]
        self._w[:n] = self._w[live]
        self._alive[:n] = True
        self._alive[n:self._top] = False
        obj = self._obj
        obj[:n] = [obj[s] for s in live.tolist()]
        for s in range(n, self._top):
            obj[s] = None
        for new, key in enumerate(self._slot):            # dict order == slot order
            self._slot[key] = new
        self._top, self._dead = n, 0

    # ------------------------------------------------------------------ mapping protocol
    def __len__(self):
        return len(self._slot)

    def __iter__(self):
        return iter(self._slot)

    def __contains__(self, key):
        return key in self._slot

    def __getitem__(self, key):
        s = self._slot[key]
        e = self._obj[s]
        if e is None:
            r0, k0, r1, k1 = self._ends[s].tolist()
            e = self._obj[s] = self._edge_type(r0, k0, r1, k1, float(self._w[s]))
        return e

    def __setitem__(self, key, edge):
        s = self._slot.get(key)
        if s is None:
            self._reserve(1)
            s = self._slot[key] = self._top
            self._top += 1
            self._alive[s] = True
        self._ends[s] = edge[:4]
        self._w[s] = edge[4]
        self._obj[s] = edge

    def __delitem__(self, key):
        s = self._slot.pop(key)
        self._alive[s] = False
        self._obj[s] = None
        self._dead += 1

    def __repr__(self):
        return "CandidateTable(%d edges)" % len(self)

    def __eq__(self, other):
        if isinstance(other, (dict, MutableMapping)):
            return dict(self.items()) == dict(other.items())
        return NotImplemented

    __hash__ = None

    def clear(self):
        self.__init__(self._edge_type)

    # ------------------------------------------------------------------ columnar access
    def columns(self):
        """(ends int64 [m, 4] as spelled = r0,k0,r1,k1; weight float64 [m]) of the live edges
        in dictionary order.  Views when nothing was removed, copies otherwise."""
        if self._dead and self._dead * 2 > self._top:
            self._compact()
        if self._dead == 0:
            return self._ends[:self._top], self._w[:self._top]
        live = self._alive[:self._top]
        return self._ends[:self._top][live], self._w[:self._top][live]

    def raw(self):
        """(ends [top, 4], weight [top], alive [top] or None): every slot handed out so far, dead
        ones included (`alive` is None when there are none) — for callers that filter once at
        the end instead of copying the live rows first."""
        if self._dead and self._dead * 2 > self._top:
            self._compact()
        top = self._top
        return self._ends[:top], self._w[:top], (self._alive[:top] if self._dead else None)

    def robots_present(self):
        """Sorted robot ids that appear in a live edge (either end)."""
        ends, _, alive = self.raw()
        if len(ends) == 0:
            return []
        if alive is None:
            seen = np.bincount(ends[:, 0]) > 0, np.bincount(ends[:, 2]) > 0
        else:
            seen = np.bincount(ends[:, 0], weights=alive) > 0, np.bincount(ends[:, 2], weights=alive) > 0
        return sorted(set(np.flatnonzero(seen[0]).tolist()) | set(np.flatnonzero(seen[1]).tolist()))

    def weight_of(self, key):
        """Stored weight for `key`, or None — without materialising the edge object."""
        s = self._slot.get(key)
        return None if s is None else float(self._w[s])

    def remove_keys(self, keys):
        """pop(key, None) for many keys."""
        pop = self._slot.pop
        gone = [s for s in (pop(k, None) for k in keys) if s is not None]
        if gone:
            self._alive[gone] = False
            obj = self._obj
            for s in gone:
                obj[s] = None
            self._dead += len(gone)

    def put_rows(self, keys, ends, weights):
        """Bulk `self[key] = edge` from columns, no edge objects created: `keys` a list of
        4-tuples (distinct), `ends` int64 [n, 4] as spelled, `weights` float64 [n].  New keys are
        appended in the given order, stored keys are overwritten in place."""
        n = len(keys)
        if n == 0:
            return
        self._reserve(n)
        found = list(map(self._slot.get, keys))
        slots = np.array([-1 if s is None else s for s in found], dtype=np.int64) \
            if len(self._slot) else np.full(n, -1, dtype=np.int64)
        fresh = np.flatnonzero(slots < 0)
        if len(fresh):
            slots[fresh] = self._top + np.arange(len(fresh))
            new_keys = keys if len(fresh) == n else [keys[i] for i in fresh.tolist()]
            self._slot.update(zip(new_keys, range(self._top, self._top + len(fresh))))
            self._top += len(fresh)
        self._ends[slots] = ends
        self._w[slots] = weights
        self._alive[slots] = True
        if len(fresh) < n:                                # overwritten entries: drop stale objects
            obj = self._obj
            for s in slots[np.setdiff1d(np.arange(n), fresh)].tolist():
                obj[s] = None
