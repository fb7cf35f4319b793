"""Cycles per grid barrier (csrc/mac.cu grid_barrier) on an otherwise empty co-resident grid, per
memory-ordering recipe, grid size and number of global stores a thread publishes before the barrier."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cslam_b200 import _lib

lib = _lib.load()
for ctas in (1, 32, 109, 148):
    for stores in (0, 8):
        row = []
        for variant in (0, 1, 2, 3, 4, 5, 6):
            c = ctypes.c_int64()
            _lib.check(lib.cslam_debug_grid_barrier(ctas, 256, 2000, stores, variant, 0, ctypes.byref(c)))
            row.append(c.value)
        print(f"ctas {ctas:4d} stores/thread {stores}: cycles per barrier  acquire-poll {row[0]:6d}  relaxed-poll+fence {row[1]:6d}  "
              f"atom.acq_rel+relaxed-poll {row[2]:6d}  [no fences {row[3]:6d}  release only {row[4]:6d}]  atom + flag broadcast {row[5]:6d}  8 counters {row[6]:6d}")
