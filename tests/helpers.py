"""Shared comparison helpers for the parity tests."""
import numpy as np


def canonical_ties(idx: np.ndarray, d2: np.ndarray):
    """Re-order runs of bit-equal distances by ascending index (nanoflann's order among
    equal distances is KD-traversal order; the build's is (d2, idx))."""
    idx = idx.copy()
    for r in range(idx.shape[0]):
        order = np.lexsort((idx[r], d2[r]))
        idx[r] = idx[r][order]
    return idx


def boundary_tie_rows(d2_full_sorted_k1: np.ndarray, k: int):
    """Rows where the k-th and (k+1)-th smallest distances are bit-equal (the reference may
    legitimately return either row)."""
    return d2_full_sorted_k1[:, k - 1] == d2_full_sorted_k1[:, k]


def assert_knn_equal(idx, d2, ref_idx, ref_d2):
    assert np.array_equal(np.asarray(d2).view(np.uint32), np.asarray(ref_d2).view(np.uint32)), \
        "squared distances are not bit-equal"
    assert np.array_equal(np.asarray(idx).astype(np.uint64), np.asarray(ref_idx).astype(np.uint64)), \
        "neighbour indices differ"
