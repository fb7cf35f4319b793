"""Seeded input generators shared by oracle/make_golden.py and the tests, so that the
golden files only need to store OUTPUTS of the reference (inputs are re-created from
the seed).  TEST INFRASTRUCTURE ONLY."""
import numpy as np

NNS_CASES = {"a": (300, 64, 10, np.float64), "b": (1000, 256, 30, np.float64),
             "c": (300, 64, 10, np.float32)}


def nns_case(tag):
    n, d, k, qdtype = NNS_CASES[tag]
    rng = np.random.default_rng(1234 + ord(tag))
    pool = rng.random((n, d))
    pool /= np.linalg.norm(pool, axis=1, keepdims=True)
    qs = rng.random((6, d))
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    items = [7 * i + 1 for i in range(n)]
    return pool, qs.astype(qdtype), items, k


MAC_CASES = {"g0": (3, 20, 60, 8, 0), "g1": (4, 250, 1500, 40, 1), "g2": (5, 60, 400, 20, 2),
             "g3": (2, 10, 12, 4, 3)}


def multi_robot_graph(R, P, m, seed):
    """Reference-test style graph (tests/test_algebraic_connectivity.py:38-74 of the
    reference): R chains of P poses (implicit odometry), R-1 fixed edges joining
    consecutive robots' last poses, m distinct random inter-robot candidates with
    U(0,1) weights.  Tuples are (robot0, kf0, robot1, kf1, weight)."""
    rng = np.random.default_rng(seed)
    fixed = [(r, P - 1, r + 1, P - 1, 1.0) for r in range(R - 1)]
    seen = set()
    cand = []
    while len(cand) < m:
        r0, r1 = (int(x) for x in rng.choice(R, size=2, replace=False))
        k0, k1 = int(rng.integers(0, P)), int(rng.integers(0, P))
        key = (r0, k0, r1, k1) if r0 < r1 else (r1, k1, r0, k0)
        if key in seen:
            continue
        seen.add(key)
        cand.append((r0, k0, r1, k1, float(rng.random())))
    return fixed, cand


def mac_scale_graph(R=8, P=12500, m=1000000, seed=0):
    """BASELINE.json configs[4]: R robots x P poses (odometry
    chains), R-1 fixed bridges between consecutive robots' last poses, m inter-robot
    candidates with uniform endpoints and U(0,1) weights.  Returns rekeyed arrays
    (fixed_i, fixed_j, fixed_w), (cand_i, cand_j, cand_w), n."""
    rng = np.random.default_rng(seed)
    n = R * P
    fi = np.concatenate([np.arange(r * P, r * P + P - 1) for r in range(R)] +
                        [np.array([(r + 1) * P - 1 for r in range(R - 1)])]).astype(np.int32)
    fj = np.concatenate([np.arange(r * P + 1, r * P + P) for r in range(R)] +
                        [np.array([(r + 2) * P - 1 for r in range(R - 1)])]).astype(np.int32)
    fw = np.ones(len(fi))
    r0 = rng.integers(0, R, m)
    r1 = (r0 + rng.integers(1, R, m)) % R
    ci = (r0 * P + rng.integers(0, P, m)).astype(np.int32)
    cj = (r1 * P + rng.integers(0, P, m)).astype(np.int32)
    cw = rng.random(m)
    return (fi, fj, fw), (ci, cj, cw), n


# ---- descriptor heads ----------------------------------------------------------------
def keyframe_image(seed=7, h=480, w=640):
    """Synthetic RGB keyframe (SURVEY.md section 8d, C2 inputs)."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def vlad_case(seed=0, n=2, c=512, hw=14, k=64):
    """Feature map + NetVLADLayer parameters (conv.weight [k,c], centroids [k,c])."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, c, hw, hw)).astype(np.float32)
    conv_w = (rng.standard_normal((k, c)) * 0.2).astype(np.float32)
    centroids = rng.random((k, c)).astype(np.float32)
    return x, conv_w, centroids


def gem_case(seed=1, n=3, c=512, hw=7, d=64):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, c, hw, hw)).astype(np.float32)
    w = (rng.standard_normal((d, c)) / np.sqrt(c)).astype(np.float32)
    b = (rng.standard_normal(d) * 0.01).astype(np.float32)
    return x, 3.0, 1e-6, w, b


def pca_case(seed=2, n=5, din=2048, dout=96, whiten=True):
    """Synthetic whitening PCA in the spirit of SURVEY.md section 8d (C2), small."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, din)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    comp = (rng.standard_normal((dout, din)) / np.sqrt(din)).astype(np.float32)
    mean = (0.01 * rng.standard_normal(din)).astype(np.float32)
    ev = rng.uniform(0.5, 1.5, dout)
    return x, comp, mean, ev, whiten


def subsample(a, step):
    return np.ascontiguousarray(np.asarray(a).reshape(-1)[::step])


# ---- front-end scenario (A7/A17): a seeded stream of local and remote keyframe descriptors ----
FRONTEND_PARAMS = {
    'robot_id': 0, 'max_nb_robots': 3, 'frontend.sensor_type': 'stereo',
    'frontend.similarity_threshold': 0.8, 'frontend.enable_sparsification': True,
    'evaluation.enable_sparsification_comparison': False, 'frontend.nb_best_matches': 10,
    'frontend.intra_loop_min_inbetween_keyframes': 5,
    'frontend.enable_intra_robot_loop_closures': True,
    'frontend.inter_robot_loop_closure_budget': 5,
}


def frontend_scenario(seed=11, dim=32, n_local=48, n_remote=40, block=8):
    """Events in arrival order: ('local', [kf ids], desc [b, dim]) blocks of `block` local
    keyframes alternating with ('remote', robot, [kf ids], desc) blocks of remote
    descriptors from robots 1 and 2.  Descriptors are unit-norm float32 drawn around a
    few shared 'places' so that a useful fraction of similarities clears the 0.8
    threshold and some do not."""
    rng = np.random.default_rng(seed)
    places = rng.random((12, dim))

    def draw(m):
        base = places[rng.integers(0, len(places), m)]
        x = base + 0.35 * rng.random((m, dim))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        return x.astype(np.float32)

    events = []
    next_local, next_remote = 0, {1: 0, 2: 0}
    while next_local < n_local or any(v < n_remote for v in next_remote.values()):
        if next_local < n_local:
            ids = list(range(next_local, min(n_local, next_local + block)))
            events.append(('local', ids, draw(len(ids))))
            next_local += len(ids)
        for r in (1, 2):
            if next_remote[r] < n_remote:
                ids = list(range(next_remote[r], min(n_remote, next_remote[r] + block)))
                events.append(('remote', r, ids, draw(len(ids))))
                next_remote[r] += len(ids)
    return events


# ---------------------------------------------------------------- lidar / Scan Context (f4)
SC_CASES = {"p300": (300, 40, 11), "p4": (4, 6, 12), "p1": (1, 3, 13), "p1100": (1100, 12, 14)}


def sc_case(tag, shape=(20, 60)):
    """Synthetic scan contexts: non-negative heights with empty cells and empty sectors
    (columns), as a lidar sweep with occlusions produces.  Queries: rotated (column-rolled)
    noisy copies of pool entries, exact copies, unrelated scans, an all-zero scan.
    Returns (pool [n, rings*sectors] float64, items, queries [q, rings*sectors] float64)."""
    n, nq, seed = SC_CASES[tag]
    rng = np.random.default_rng(seed)
    R, S = shape

    def scan():
        sc = rng.random((R, S)) * 6.0
        sc[rng.random((R, S)) < 0.3] = 0.0
        sc[:, rng.random(S) < 0.12] = 0.0
        return sc

    pool = np.stack([scan() for _ in range(n)])
    queries = []
    for t in range(nq):
        kind = t % 4
        if kind == 0 or kind == 1:                       # a revisit under another heading
            sc = np.roll(pool[int(rng.integers(0, n))], int(rng.integers(0, S)), axis=1).copy()
            sc[sc > 0] += rng.normal(0, 0.3 if kind == 0 else 1.5, size=int((sc > 0).sum()))
            sc = np.maximum(sc, 0.0)
        elif kind == 2:
            sc = pool[int(rng.integers(0, n))].copy()    # exact copy
        else:
            sc = scan()                                  # a place never seen
        queries.append(sc)
    queries[-1] = np.zeros((R, S))                       # nothing engaged: similarity 0, item 0
    items = [3 * i + 5 for i in range(n)]
    return pool.reshape(n, -1), items, np.stack(queries).reshape(nq, -1)
