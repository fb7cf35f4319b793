"""Edge container and Laplacian helpers with the names of cslam/mac/utils.py.

The hot path never calls the scipy builders below (the Laplacian is assembled and kept
on the GPU by libcslam_b200, see csrc/mac.cu); they exist so that code written against
the reference API (`weight_graph_lap_from_edge_list`, `MAC.combined_laplacian`) still
gets a scipy CSR matrix when it asks for one.
"""
from collections import namedtuple

import numpy as np

# cslam/mac/utils.py:13
Edge = namedtuple('Edge', ['i', 'j', 'weight'])


def edges_to_arrays(edges):
    """list[Edge] -> (i int32[], j int32[], w float64[])"""
    n = len(edges)
    i = np.fromiter((e.i for e in edges), dtype=np.int32, count=n)
    j = np.fromiter((e.j for e in edges), dtype=np.int32, count=n)
    w = np.fromiter((e.weight for e in edges), dtype=np.float64, count=n)
    return i, j, w


def _laplacian(i, j, w, n):
    from scipy.sparse import coo_matrix, csr_matrix
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    rows = np.stack([i, j, i, j], axis=1).ravel()
    cols = np.stack([i, j, j, i], axis=1).ravel()
    data = np.stack([w, w, -w, -w], axis=1).ravel()
    return csr_matrix(coo_matrix((data, (rows, cols)), shape=[n, n]))


def weight_graph_lap_from_edge_list(edges, num_vars):
    """Weighted graph Laplacian (scipy CSR) of a list of Edge (cslam/mac/utils.py:47-83)."""
    i, j, w = edges_to_arrays(edges)
    return _laplacian(i, j, w, num_vars)


def weight_graph_lap_from_edges(edges, weights, num_poses):
    """Same from an [m, 2] index array and a weight vector (cslam/mac/utils.py:86-126)."""
    edges = np.asarray(edges).reshape(-1, 2)
    return _laplacian(edges[:, 0], edges[:, 1], weights, num_poses)


def nx_to_mac(G):
    """Unit-weight Edge list of a networkx graph (cslam/mac/utils.py:16-29)."""
    return [Edge(u, v, 1.0) for u, v in G.edges()]


def mac_to_nx(edges):
    """networkx graph of an Edge list (cslam/mac/utils.py:32-44)."""
    import networkx as nx
    G = nx.Graph()
    for e in edges:
        G.add_edge(e.i, e.j, weight=e.weight)
    return G


def split_measurements(measurements):
    """(odometry, loop closures) by |i - j| == 1 (cslam/mac/utils.py:129-145)."""
    odom, lc = [], []
    for m in measurements:
        (lc if abs(m.j - m.i) > 1 else odom).append(m)
    return odom, lc


def select_measurements(measurements, w):
    """Edges whose selection weight is exactly 1.0 (cslam/mac/utils.py:148-158)."""
    assert len(measurements) == len(w)
    return [m for m, wi in zip(measurements, w) if wi == 1.0]
