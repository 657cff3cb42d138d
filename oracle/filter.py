"""Cone density filter restated from the reference (TEST INFRASTRUCTURE ONLY):
/root/reference/examples/beam_topo_opt/pre_processor/general_filter_model.py:67-90 --
W_ij = (R - d_ij) / sum_k (R - d_ik) over the points within R = beta*h_avg (cKDTree ball query);
the reference rebuilds the tree per point, one tree gives the same neighbours."""
import numpy as np
import scipy.sparse
from scipy import spatial


def weight_matrix(coords, h_avg, beta=2.0):
    coords = np.asarray(coords, dtype=np.float64)
    nel = coords.shape[0]
    radius = beta * h_avg
    tree = spatial.cKDTree(coords)
    rows, cols, vals = [], [], []
    for i in range(nel):
        idx = tree.query_ball_point(list(coords[i]), radius)
        d = np.linalg.norm(coords[i] - coords[idx], axis=1)
        w = (radius - d) / np.sum(radius - d)
        rows += [i] * len(idx)
        cols += list(idx)
        vals += list(w)
    return scipy.sparse.csr_matrix((vals, (rows, cols)), shape=(nel, nel))
