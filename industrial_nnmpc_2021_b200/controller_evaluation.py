"""Host-side mirror of the hot-path pieces of /root/reference/lib/controller_evaluation.py.

* ``sample_prbs_like``          (:21-47)   scenario input definition, bit-reproducible
* ``_get_data_for_training``    (:254-271) state scaling of a generated dataset
* ``_post_process_data``        (:273-295) merge of the per-(task, process) dataset files
* ``H5pyTool``                  (lib/python_utils.py:41-58) dataset container (h5py, or ``.npz`` with the same keys)
* ``NeuralNetworkController``   (:780-892) deployment form of the structured network; here a
  *batched* evaluator on the GPU (``control_input_batch``) with the same weight-list layout
  ``[W1,b1,W2,b2,W3,b3,Wout]`` (W stored (in,out)), ``x/xscale`` scaling and output clip.
"""
from __future__ import annotations

import itertools

import numpy as np


def _sample_repeats(num_change, num_simulation_steps, mean_change, sigma_change):
    """How long each sampled level is held (controller_evaluation.py:21-29).

    Uses the legacy global NumPy stream on purpose: the reference seeds ``np.random.seed`` and
    draws ``rand`` then ``randn`` from it, and that MT19937 stream is frozen.
    """
    hold = np.floor(sigma_change * np.random.randn(num_change - 1) + mean_change)
    hold = np.where(hold <= 0.0, 0.0, hold)
    hold = np.append(hold, num_simulation_steps - int(np.sum(hold)))
    return hold.astype(int)


def sample_prbs_like(*, num_change, num_steps, lb, ub, mean_change, sigma_change, seed=1):
    """PRBS-like piecewise-constant signal, (num_steps, dim) (controller_evaluation.py:31-47)."""
    dim = lb.shape[0]
    lb = np.squeeze(lb)
    ub = np.squeeze(ub)
    np.random.seed(seed)
    levels = (ub - lb) * np.random.rand(num_change, dim) + lb
    hold = _sample_repeats(num_change, num_steps, mean_change, sigma_change)
    return np.repeat(levels, hold, axis=0)


def _get_data_for_training(*, data, num_samples, scale=True):
    """First ``num_samples`` rows, states divided by xscale = (max-min)/2 (:254-271)."""
    out = {k: np.asarray(data[k])[0:num_samples, :] for k in ("x", "uprev", "xs", "us", "u")}
    if not scale:
        return out
    xscale = 0.5 * (np.max(out["x"], axis=0) - np.min(out["x"], axis=0))
    out["x"] = out["x"] / xscale
    out["xs"] = out["xs"] / xscale
    return out, xscale


class H5pyTool:
    """lib/python_utils.py:41-58.  Same static methods and file keys; where h5py is not installed the keys go to
    ``<filename>.npz`` (see linearMPC._save_training_data)."""

    @staticmethod
    def load_training_data(filename):
        from .linearMPC import load_training_data
        return load_training_data(filename)

    @staticmethod
    def save_training_data(dictionary, filename):
        from .linearMPC import _save_training_data
        return _save_training_data(dictionary, filename)


def _post_process_data(*, data_filename, num_data_gen_task, num_process_per_task):
    """Compile the ``{task}-{process}-{data_filename}`` files written by ``OfflineSimulator.generate_data`` /
    ``simulate_offline`` into one dataset (controller_evaluation.py:273-295): arrays concatenated along axis 0 in
    ``(task, process)`` order, ``data_gen_time`` averaged; saved under ``data_filename`` and returned."""
    training_data = dict(x=[], uprev=[], xs=[], us=[], u=[], data_gen_time=[])
    for task, process in itertools.product(range(num_data_gen_task), range(num_process_per_task)):
        process_data = H5pyTool.load_training_data(f"{task}-{process}-{data_filename}")
        for key in process_data.keys():
            training_data[key].append(process_data[key])
    for key in training_data.keys():
        if key == "data_gen_time":
            training_data[key] = np.mean(np.asarray(training_data[key]))
        else:
            training_data[key] = np.concatenate(training_data[key], axis=0)
    H5pyTool.save_training_data(dictionary=training_data, filename=data_filename)
    return training_data


class NeuralNetworkController:
    """Batched deployment form of the structured network (controller_evaluation.py:780-892).

    ``regulator_weights`` is the Keras ``get_weights()`` list ``[W1,b1,W2,b2,W3,b3,Wout]``;
    ``xscale`` the state scaling returned by ``_get_data_for_training``.  ``control_input_batch``
    evaluates ``clip(us + NN(x/xscale,uprev,xs/xscale,us) - NN(xs/xscale,us,xs/xscale,us))`` for a
    whole batch on the GPU; ``_get_control_input`` keeps the reference's single-column signature
    (:868-875) except that it takes UNscaled x, xs and applies ``xscale`` itself.
    """

    def __init__(self, *, regulator_weights, xscale, nnwithuprev, ulb, uub, device=None, precision=None):
        from .LinearMPCLayers import RegulatorLayerWithUprev, RegulatorLayerWithoutUprev
        self.regulator_weights = regulator_weights
        self.xscale = np.asarray(xscale, dtype=np.float64).reshape(-1)
        self.nnwithuprev = nnwithuprev
        self.ulb, self.uub = ulb, uub
        dims = [w.shape[1] for w in regulator_weights[0::2]]
        cls = RegulatorLayerWithUprev if nnwithuprev else RegulatorLayerWithoutUprev
        self.layer = cls(layer_dims=dims, device=device, precision=precision)
        self.layer.set_weights(regulator_weights)

    def control_input_batch(self, x, uprev, xs, us):
        return self.layer.forward(x, uprev, xs, us, xscale=self.xscale, ulb=self.ulb, uub=self.uub)

    def _get_control_input(self, x, uprev, xs, us):
        row = lambda a: np.asarray(a, float).reshape(1, -1)
        return self.control_input_batch(row(x), row(uprev), row(xs), row(us)).reshape(-1, 1)
