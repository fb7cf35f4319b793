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

Two ways to find the slot of a key:
  host mode    a Python dict key -> slot (every key is hashed as a 4-tuple: 1.3 us per pair, 1.3 s
               per million bulk inserts)
  device mode  (`use_device_index`) the key is packed into 64 bits and looked up in an open
               addressing hash table in HBM (csrc/keymap.cu, `cslam_keymap_*`): a batch of keys is
               one kernel, the tuple-keyed view is materialised lazily, only for callers that
               iterate.  Needs robot ids < 256 and keyframe ids < 2^24; a key outside that range
               moves the table back to host mode.
"""
import ctypes
from collections.abc import MutableMapping

import numpy as np

_R_BITS, _K_BITS = 8, 24


def pack_keys(r0, k0, r1, k1):
    """uint64 key of the NORMALISED 4-tuples (r0 < r1): r0 | k0 | r1 | k1 in 8 + 24 + 8 + 24 bits."""
    r0, k0, r1, k1 = (np.asarray(a, dtype=np.uint64) for a in (r0, k0, r1, k1))
    return (r0 << np.uint64(56)) | (k0 << np.uint64(32)) | (r1 << np.uint64(24)) | k1


def keys_packable(r0, k0, r1, k1):
    hi_r, hi_k = 1 << _R_BITS, 1 << _K_BITS
    return all(len(a) == 0 or (int(np.min(a)) >= 0 and int(np.max(a)) < lim)
               for a, lim in ((r0, hi_r), (k0, hi_k), (r1, hi_r), (k1, hi_k)))


def unpack_key(key):
    key = int(key)
    return (key >> 56, (key >> 32) & 0xFFFFFF, (key >> 24) & 0xFF, key & 0xFFFFFF)


class DeviceKeyMap(object):
    """ctypes face of `cslam_keymap_*`: uint64 key -> int32 slot, batches in, batches out."""

    def __init__(self, device=0, capacity_hint=1024):
        from . import _lib
        self._lib = _lib
        h = ctypes.c_void_p()
        _lib.check(_lib.load().cslam_keymap_create(int(capacity_hint), int(device), ctypes.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                self._lib.load().cslam_keymap_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.load().cslam_keymap_size(self._h))

    def lookup(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.empty(len(keys), dtype=np.int32)
        self._lib.check(self._lib.load().cslam_keymap_lookup(self._h, self._lib.ptr(keys), len(keys),
                                                             self._lib.ptr(out)))
        return out

    def insert(self, keys, values):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        values = np.ascontiguousarray(values, dtype=np.int32)
        assert len(keys) == len(values)
        self._lib.check(self._lib.load().cslam_keymap_insert(self._h, self._lib.ptr(keys),
                                                             self._lib.ptr(values), len(keys)))

    def erase(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.empty(len(keys), dtype=np.int32)
        self._lib.check(self._lib.load().cslam_keymap_erase(self._h, self._lib.ptr(keys), len(keys),
                                                            self._lib.ptr(out)))
        return out


class CandidateTable(MutableMapping):

    def __init__(self, edge_type, capacity=1024):
        self._edge_type = edge_type
        self._slot = {}                                   # host mode: key -> slot (insertion ordered)
        self._index = None                                # device mode: DeviceKeyMap key64 -> slot
        self._device = 0
        self._ends = np.empty((capacity, 4), dtype=np.int64)   # r0, k0, r1, k1 as spelled
        self._w = np.empty(capacity, dtype=np.float64)
        self._alive = np.zeros(capacity, dtype=bool)
        self._key64 = np.zeros(capacity, dtype=np.uint64)  # device mode: packed normalised key per slot
        self._obj = [None] * capacity                     # materialised EdgeInterRobot or None
        self._top = 0                                     # slots handed out so far
        self._dead = 0

    # ------------------------------------------------------------------ modes
    @property
    def device_mode(self):
        return self._index is not None

    def use_device_index(self, device=0):
        """Switch to device mode: the key -> slot map moves into the HBM hash table.  Returns False
        (and stays in host mode) when a stored key does not fit the 64-bit packing."""
        if self._index is not None:
            return True
        keys = list(self._slot)
        cols = [np.array([k[c] for k in keys], dtype=np.int64) for c in range(4)] if keys \
            else [np.zeros(0, np.int64)] * 4
        if not keys_packable(*cols):
            return False
        self._device = int(device)
        index = DeviceKeyMap(self._device, max(1024, 2 * len(keys)))
        if keys:
            packed = pack_keys(*cols)
            slots = np.fromiter(self._slot.values(), dtype=np.int64, count=len(keys))
            index.insert(packed, slots.astype(np.int32))
            self._key64[slots] = packed
        self._index = index
        self._slot = {}
        return True

    def _to_host_mode(self):
        """Back to the Python dict (a key outside the packable range arrived)."""
        if self._index is None:
            return
        live = np.flatnonzero(self._alive[:self._top])
        self._slot = {unpack_key(k): int(s_) for k, s_ in zip(self._key64[live].tolist(), live.tolist())}
        self._index = None

    @staticmethod
    def _pack_one(key):
        r0, k0, r1, k1 = key
        if not (0 <= r0 < (1 << _R_BITS) and 0 <= r1 < (1 << _R_BITS) and
                0 <= k0 < (1 << _K_BITS) and 0 <= k1 < (1 << _K_BITS)):
            return None
        return np.array([(int(r0) << 56) | (int(k0) << 32) | (int(r1) << 24) | int(k1)], dtype=np.uint64)

    def _find(self, key):
        """slot of `key` or None (either mode)."""
        if self._index is None:
            return self._slot.get(key)
        try:
            packed = self._pack_one(key)
        except (TypeError, ValueError):
            return None
        if packed is None:
            return None
        s = int(self._index.lookup(packed)[0])
        return None if s < 0 else s

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
        self._key64 = np.concatenate([self._key64, np.zeros(cap - len(self._key64), np.uint64)])
        self._obj.extend([None] * (cap - len(self._obj)))

    def _compact(self):
        """Squeeze the dead slots out, keeping the order."""
        live = np.flatnonzero(self._alive[:self._top])
        n = len(live)
        self._ends[:n] = self._ends[live]
        self._w[:n] = self._w[live]
        self._key64[:n] = self._key64[live]
        self._alive[:n] = True
        self._alive[n:self._top] = False
        obj = self._obj
        obj[:n] = [obj[s] for s in live.tolist()]
        for s in range(n, self._top):
            obj[s] = None
        if self._index is None:
            for new, key in enumerate(self._slot):        # dict order == slot order
                self._slot[key] = new
        else:                                             # every slot moved: a fresh index
            self._index = DeviceKeyMap(self._device, max(1024, 2 * n))
            if n:
                self._index.insert(self._key64[:n], np.arange(n, dtype=np.int32))
        self._top, self._dead = n, 0

    # ------------------------------------------------------------------ mapping protocol
    def __len__(self):
        return self._top - self._dead

    def __iter__(self):
        if self._index is None:
            return iter(self._slot)
        live = np.flatnonzero(self._alive[:self._top])
        return iter([unpack_key(k) for k in self._key64[live].tolist()])

    def __contains__(self, key):
        return self._find(key) is not None

    def __getitem__(self, key):
        s = self._find(key)
        if s is None:
            raise KeyError(key)
        e = self._obj[s]
        if e is None:
            r0, k0, r1, k1 = self._ends[s].tolist()
            e = self._obj[s] = self._edge_type(r0, k0, r1, k1, float(self._w[s]))
        return e

    def __setitem__(self, key, edge):
        if self._index is not None and self._pack_one(key) is None:
            self._to_host_mode()
        s = self._find(key)
        if s is None:
            self._reserve(1)
            s = self._top
            self._top += 1
            if self._index is None:
                self._slot[key] = s
            else:
                packed = self._pack_one(key)
                self._index.insert(packed, np.array([s], dtype=np.int32))
                self._key64[s] = packed[0]
            self._alive[s] = True
        self._ends[s] = edge[:4]
        self._w[s] = edge[4]
        self._obj[s] = edge

    def __delitem__(self, key):
        if self._index is None:
            s = self._slot.pop(key)
        else:
            s = self._find(key)
            if s is None:
                raise KeyError(key)
            self._index.erase(self._pack_one(key))
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
        device = self._device if self._index is not None else None
        self.__init__(self._edge_type)
        if device is not None:
            self.use_device_index(device)

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
        s = self._find(key)
        return None if s is None else float(self._w[s])

    def remove_keys(self, keys):
        """pop(key, None) for many keys."""
        keys = list(keys)
        if self._index is None:
            pop = self._slot.pop
            gone = [s for s in (pop(k, None) for k in keys) if s is not None]
        else:
            cols = [np.array([k[c] for k in keys], dtype=np.int64) for c in range(4)] if keys \
                else [np.zeros(0, np.int64)] * 4
            ok = np.ones(len(keys), dtype=bool)
            for a, lim in zip(cols, (1 << _R_BITS, 1 << _K_BITS, 1 << _R_BITS, 1 << _K_BITS)):
                ok &= (a >= 0) & (a < lim)                # an unpackable key cannot be stored
            packed = np.unique(pack_keys(*[a[ok] for a in cols]))
            erased = self._index.erase(packed) if len(packed) else np.zeros(0, np.int32)
            gone = erased[erased >= 0].astype(np.int64).tolist()
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
        if self._index is not None:
            cols = [np.array([k[c] for k in keys], dtype=np.int64) for c in range(4)]
            if keys_packable(*cols):
                return self.put_rows_packed(pack_keys(*cols), ends, weights)
            self._to_host_mode()
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

    # ------------------------------------------------------------------ device mode, packed keys
    def lookup_packed(self, key64):
        """slots (int64, -1 = absent) of packed normalised keys; device mode only."""
        return self._index.lookup(key64).astype(np.int64)

    def put_rows_packed(self, key64, ends, weights, slots=None):
        """`put_rows` with packed normalised keys (distinct): one lookup kernel + one insert kernel,
        no Python object per key.  `slots`: result of `lookup_packed(key64)` if the caller has it."""
        n = len(key64)
        if n == 0:
            return
        key64 = np.ascontiguousarray(key64, dtype=np.uint64)
        moved = self._top + n > len(self._w) and self._dead > 0
        self._reserve(n)                                  # may compact: slots change
        if slots is None or moved:
            slots = self.lookup_packed(key64)
        else:
            slots = np.array(slots, dtype=np.int64)
        fresh = np.flatnonzero(slots < 0)
        if len(fresh):
            slots[fresh] = self._top + np.arange(len(fresh))
            self._index.insert(key64[fresh], slots[fresh].astype(np.int32))
            self._key64[slots[fresh]] = key64[fresh]
            self._top += len(fresh)
        self._ends[slots] = ends
        self._w[slots] = weights
        self._alive[slots] = True
        if len(fresh) < n:
            obj = self._obj
            stale = np.ones(n, dtype=bool)
            stale[fresh] = False
            for s_ in slots[stale].tolist():
                obj[s_] = None
