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


# ----------------------------------------------------------------------------- training step
def mse_loss_and_grads(weights, inputs, u, with_uprev):
    """Loss and gradients of keras 'mean_squared_error' (mean over batch and outputs) for the structured layer
    (LinearMPCLayers.py:40-61 / :91-112) by hand-written backpropagation; gradients in get_weights() order."""
    if with_uprev:
        x, uprev, xs, us = inputs
        a1 = np.concatenate((x, uprev, xs, us), axis=-1)
        a2 = np.concatenate((xs, us, xs, us), axis=-1)
    else:
        x, xs, us = inputs
        a1 = np.concatenate((x, xs, us), axis=-1)
        a2 = np.concatenate((xs, xs, us), axis=-1)
    nl = (len(weights) + 1) // 2

    def forward(a):
        acts = [a]
        for i in range(nl - 1):
            a = relu(a @ weights[2 * i] + weights[2 * i + 1])
            acts.append(a)
        return acts, a @ weights[-1]

    acts1, f1 = forward(a1)
    acts2, f2 = forward(a2)
    out = us + f1 - f2
    err = out - u
    B, nu = u.shape
    loss = float(np.mean(err ** 2))
    g = 2.0 * err / (B * nu)
    grads = [None] * len(weights)
    for acts, d in ((acts1, g), (acts2, -g)):
        gW = acts[-1].T @ d
        grads[-1] = gW if grads[-1] is None else grads[-1] + gW
        d = (d @ weights[-1].T) * (acts[-1] > 0)
        for i in range(nl - 2, -1, -1):
            gW, gb = acts[i].T @ d, d.sum(axis=0)
            grads[2 * i] = gW if grads[2 * i] is None else grads[2 * i] + gW
            grads[2 * i + 1] = gb if grads[2 * i + 1] is None else grads[2 * i + 1] + gb
            if i > 0:
                d = (d @ weights[2 * i].T) * (acts[i] > 0)
    return loss, grads


def adam_step(weights, grads, m, v, t, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    """keras.optimizers.Adam (non-amsgrad) as applied per variable: lr_t = lr sqrt(1-b2^t)/(1-b1^t),
    w -= lr_t m / (sqrt(v) + eps).  Returns new (weights, m, v)."""
    lr_t = lr * np.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t)
    nw, nm, nv = [], [], []
    for w, g, mi, vi in zip(weights, grads, m, v):
        mi = beta_1 * mi + (1.0 - beta_1) * g
        vi = beta_2 * vi + (1.0 - beta_2) * g * g
        nw.append(w - lr_t * mi / (np.sqrt(vi) + epsilon))
        nm.append(mi)
        nv.append(vi)
    return nw, nm, nv
