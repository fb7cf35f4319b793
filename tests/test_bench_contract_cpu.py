"""bench.py contract, as far as it can be checked without a GPU: the reference arm (the oracle port
of the reference's CPU path on the host cores) prints ONE JSON line with the keys the driver reads,
and the product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "keyframes/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("keyframes/s descriptor+NNS+sparsify")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "keyframes/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)


def test_nns_parity_block_with_a_stand_in_pool():
    """The recall check bench.py attaches to its line (`parity`), driven with a CPU stand-in for
    the GPU pool: an exact pool gives recall 1 and identical ranked lists, a pool that returns a
    wrong neighbour is caught."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from oracle.nns import NNSOracle

    class Pool:
        def __init__(self, rows, spoil=False):
            self.rows, self.n, self.spoil = rows.astype(np.float32), len(rows), spoil

        def read_rows(self, start, count):
            return self.rows[start:start + count].copy()

        def search_batch_device(self, q, k):
            orc = NNSOracle(self.rows.shape[1])
            orc.data, orc.n = self.rows, self.n
            idx, sims = [], []
            for row in q.numpy():
                full = orc.similarities_vec(row)
                order = np.argsort(full)[::-1][:k].copy()
                if self.spoil:
                    order[-1] = np.argsort(full)[0]           # the worst row instead of the k-th best
                idx.append(order)
                sims.append(full[order])
            return torch.tensor(np.array(idx), dtype=torch.int32), torch.tensor(np.array(sims))

    rng = np.random.default_rng(0)
    rows = rng.random((2500, 32))
    good = bench.nns_parity(Pool(rows), 32, 10, torch.device("cpu"), own_rows=5, nq_fresh=3, chunk=1000)
    assert good["ok"] and good["nns_recall_at_k"] == 1.0 and good["nns_ranked_lists_identical"] == 8
    assert good["nns_max_abs_dsim"] == 0.0 and good["pool_rows_scored"] == 2500 and good["nns_top_k"] == 10
    assert good["last_step_rows_find_themselves"] == 5
    bad = bench.nns_parity(Pool(rows, spoil=True), 32, 10, torch.device("cpu"), own_rows=5, nq_fresh=3, chunk=1000)
    assert not bad["ok"] and bad["nns_recall_at_k"] == pytest.approx(0.9) and bad["nns_ranked_lists_identical"] == 0
    # what round 1's benchmark actually produced: zero / non-finite rows in the pool (descriptors
    # read before they were written).  The checker must fail, not report "identical lists".
    broken = rows.copy()
    broken[-3:] = 0.0
    z = bench.nns_parity(Pool(broken), 32, 10, torch.device("cpu"), own_rows=5, nq_fresh=3, chunk=1000)
    assert not z["ok"] and (z["pool_zero_rows"] == 3 or not z["pool_and_scores_finite"])
    broken[-3:] = np.inf
    z = bench.nns_parity(Pool(broken), 32, 10, torch.device("cpu"), own_rows=5, nq_fresh=3, chunk=1000)
    assert not z["ok"] and not z["pool_and_scores_finite"]
    # exact duplicates in the pool (the rotating synthetic batches): recall counts ties as hits
    dup = np.concatenate([rows, rows[-5:]])
    d = bench.nns_parity(Pool(dup), 32, 10, torch.device("cpu"), own_rows=5, nq_fresh=3, chunk=1000)
    assert d["ok"] and d["nns_recall_at_k"] == 1.0


def test_lists_match_modulo_ties_rejects_non_finite_scores():
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle.nns import lists_match_modulo_ties
    sims = np.array([0.9, 0.8, 0.7, 0.6])
    assert lists_match_modulo_ties([0, 1], [0, 1], sims)
    assert not lists_match_modulo_ties([0, 2], [0, 1], sims)
    sims_nan = np.array([np.nan, np.nan, 0.7, 0.6])
    assert not lists_match_modulo_ties([0, 2], [1, 3], sims_nan)
