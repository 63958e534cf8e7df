"""NumPy float64 restatement of the structured regulator network (TEST INFRASTRUCTURE).

Follows /root/reference/lib/LinearMPCLayers.py:40-61 (with uprev), :91-112 (without) and the
deployment form /root/reference/lib/controller_evaluation.py:863-892.  TensorFlow is absent here,
so the Keras layers cannot be imported; ``keras.layers.Dense`` is y = act(x @ W + b) with W (in,out),
which is what the reference's own NumPy controller evaluates column-wise as W.T @ u + b (:877-886).
"""
import numpy as np


def relu(x):
    """controller_evaluation.py:776-778."""
    return np.where(x < 0, 0.0, x)


def regulator_nn_output(weights, x, uprev, xs, us, with_uprev):
    """Column form (k,1) inputs, exactly controller_evaluation.py:877-886."""
    u = np.concatenate((x, uprev, xs, us), axis=0) if with_uprev else np.concatenate((x, xs, us), axis=0)
    for i in range(0, len(weights) - 1, 2):
        W, b = weights[i:i + 2]
        u = relu(W.T @ u + b[:, np.newaxis])
    return weights[-1].T @ u


def control_input(weights, x, uprev, xs, us, with_uprev, xscale=None, ulb=None, uub=None):
    """controller_evaluation.py:863-875, :888-892 for one state (column vectors)."""
    if xscale is not None:
        x, xs = x / xscale, xs / xscale
    u = regulator_nn_output(weights, x, uprev, xs, us, with_uprev)
    u = u - regulator_nn_output(weights, xs, us, xs, us, with_uprev)
    u = us + u
    if ulb is not None:
        u = np.where(u > uub, uub, u)
        u = np.where(u < ulb, ulb, u)
    return u


def layer_call(weights, inputs, with_uprev):
    """Batched (B,k) form of RegulatorLayerWith[out]Uprev.call (LinearMPCLayers.py:40-61, :91-112)."""
    def net(z):
        for i in range(0, len(weights) - 1, 2):
            z = relu(z @ weights[i] + weights[i + 1])
        return z @ weights[-1]
    if with_uprev:
        x, uprev, xs, us = inputs
        return us + net(np.concatenate((x, uprev, xs, us), axis=-1)) - net(np.concatenate((xs, us, xs, us), axis=-1))
    x, xs, us = inputs
    return us + net(np.concatenate((x, xs, us), axis=-1)) - net(np.concatenate((xs, xs, us), axis=-1))
