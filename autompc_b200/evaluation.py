"""k-step model evaluation on the device -- SURVEY.md 8(f) row 3.

Mirror of ``autompc.evaluation.model_metrics.get_model_rmse`` (autompc/evaluation/model_metrics.py:12-43) for the
engine's MLP model: all windows of all trajectories are rolled ``horizon`` steps with their recorded controls in ONE
kernel launch (``B200MLP.rollout_batch`` -> ``ampc_mlp_rollout_batch``) instead of ``horizon`` ``pred_batch`` calls per
trajectory; the error statistics are the reference's (float64, same normalisation).  No CPU fallback: a model without
``rollout_batch`` raises.
"""
import numpy as np

from .mlp import B200MLP


def _as_b200(model):
    if isinstance(model, B200MLP):
        return model
    if hasattr(model, "net") and hasattr(model, "xu_means"):      # a trained reference autompc.sysid.mlp.MLP
        return B200MLP(model.system, model)
    raise ValueError("get_model_rmse runs on the B200 engine for MLP models only (got %s)" % type(model).__name__)


def get_model_rmse(model, trajs, horizon=1):
    """(Unnormalised) RMSE at a fixed prediction horizon; same value as the reference function.

    ``trajs``: objects with ``.obs`` (L,nx) and ``.ctrls`` (L,nu) (reference ``Trajectory``), or (obs, ctrls) pairs.
    """
    horizon = int(horizon)
    if horizon < 1:
        raise ValueError("horizon must be >= 1")
    m = _as_b200(model)
    starts, ctrl_cols, actual = [], [], []
    for tr in trajs:
        obs, ctrls = (tr if isinstance(tr, (tuple, list)) else (tr.obs, tr.ctrls))
        obs, ctrls = np.asarray(obs, dtype=np.float64), np.asarray(ctrls, dtype=np.float64)
        n = obs.shape[0] - horizon                                      # windows of this trajectory (:33)
        if n <= 0:
            continue
        starts.append(obs[:n])
        ctrl_cols.append(np.stack([ctrls[k:k + n] for k in range(horizon)]))   # ctrls[k:-(horizon-k)]  (:35)
        actual.append(obs[horizon:])                                    # (:38)
    if not starts:
        raise ValueError("no trajectory is longer than the horizon")
    pred = m.rollout_batch(np.concatenate(starts), np.concatenate(ctrl_cols, axis=1))
    sqerrs = (pred - np.concatenate(actual)) ** 2                       # (:39-41)
    return float(np.sqrt(np.mean(sqerrs, axis=None) * pred.shape[1]))   # (:42)
