"""NumPy restatement of the multi-GPU softmax exchange -- TEST INFRASTRUCTURE.

The reference has no counterpart (it is single process); this restates what replaces
``MPPI.update`` (``autompc/control/mppi.py:110-118``) when the K samples are sharded over ranks
(SURVEY.md 8(e)): every shard emits the record ``[m, s, W]`` with ``m = min c``,
``s = sum_k exp(-(c_k - m)/lmda)``, ``W[h,j] = sum_k exp(-(c_k - m)/lmda) * eps[h,k,j]``; the records are
all-gathered and merged with log-sum-exp rescaling.  ``merge_records`` mirrors ``ampc_merge_records``
(autompc_b200/csrc/ampc_common.cuh)."""
import numpy as np


def shard_record(costs, eps, lmda):
    """costs (K_local,), eps (H, K_local, nu) clipped noise -> float64 record of 2 + H*nu numbers."""
    m = np.min(costs)
    w = np.exp(-(costs - m) / lmda)
    W = np.einsum("hkj,k->hj", eps, w)
    return np.concatenate([[m, w.sum()], W.ravel()])


def merge_records(records, lmda, H, nu):
    """records (n, 2 + H*nu) -> the update  sum_k w_k eps_k / sum_k w_k  of mppi.py:115-117, shape (H, nu)."""
    records = np.asarray(records, dtype=np.float64)
    m = records[:, 0].min()
    scale = np.exp(-(records[:, 0] - m) / lmda)
    s = np.sum(records[:, 1] * scale)
    W = np.einsum("r,re->e", scale, records[:, 2:])
    return (W / s).reshape(H, nu)
