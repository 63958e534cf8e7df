"""Drop-in mirror of the structured-network layers of /root/reference/lib/LinearMPCLayers.py
(forward pass only), evaluated by the fused tensor-core kernels of libnnmpc.so (INT8 tcgen05 layers with
exact INT32 accumulation by default, FP64 DMMA on request).

    RegulatorLayerWithUprev    u = us + NN(x, uprev, xs, us) - NN(xs, us, xs, us)     (:15-64)
    RegulatorLayerWithoutUprev u = us + NN(x, xs, us)        - NN(xs, xs, us)         (:66-115)
    RegulatorModel             functional wrapper, input order [x, (uprev), xs, us]   (:117-133)

Same constructor arguments and call convention as the Keras classes (a list of (B,k) arrays in,
(B,Nu) out, float64 - the reference sets ``set_floatx('float64')``, :13).  Weights follow Keras:
created on first call (Glorot-uniform kernels, zero biases), ``get_weights()/set_weights()`` use
the list order ``[W1,b1,...,W_{L-1},b_{L-1},Wout]`` with ``W`` stored (in,out); the last Dense has
no bias (:31-32).  Training (backward) is out of scope of this path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class _StructuredRegulatorLayer:
    _with_uprev = True
    _MODES = {"f64": 0, "tc": 1, "fp16": 2}

    def __init__(self, layer_dims, trainable=True, name=None, *, device=None, seed=None, precision=None):
        """``precision`` (additive): "tc" - Dense layers on the INT8 tcgen05 tensor cores (digit-plane products with
        exact INT32 accumulation), outputs within ~1e-7 of the float64 Keras layer; "f64" - FP64 tensor-core GEMMs;
        "fp16" - split-fp16 tcgen05 GEMMs with fp32 accumulation (~2e-5, outside the 1e-5 tolerance: comparison only).
        Default: environment variable NNMPC_MLP, else "tc"."""
        import os
        self.precision = precision or os.environ.get("NNMPC_MLP", "tc")
        if self.precision not in self._MODES:
            raise ValueError(f"precision must be one of {sorted(self._MODES)}")
        self.layer_dims = [int(d) for d in layer_dims]
        self.trainable, self.name = trainable, name
        self._seed = seed
        self._device = device
        self._weights = None
        self._handle = None
        self._nx = self._nu = None

    # -- Keras-like weight handling ---------------------------------------------------------
    def _input_width(self, nx, nu):
        return 2 * nx + (2 if self._with_uprev else 1) * nu

    def build(self, nx, nu):
        """Create Glorot-uniform kernels / zero biases like keras.layers.Dense defaults."""
        if nu != self.layer_dims[-1]:
            raise ValueError(f"last layer width {self.layer_dims[-1]} must equal Nu={nu}")
        rng = np.random.default_rng(self._seed)
        widths = [self._input_width(nx, nu)] + self.layer_dims
        ws = []
        for i in range(len(self.layer_dims)):
            lim = np.sqrt(6.0 / (widths[i] + widths[i + 1]))
            ws.append(rng.uniform(-lim, lim, size=(widths[i], widths[i + 1])))
            if i < len(self.layer_dims) - 1:
                ws.append(np.zeros(widths[i + 1]))
        self._nx, self._nu = nx, nu
        self._install(ws)

    def get_weights(self):
        self._sync_weights()
        return [w.copy() for w in (self._weights or [])]

    def set_weights(self, weights):
        weights = [np.asarray(w, dtype=np.float64) for w in weights]
        if len(weights) != 2 * len(self.layer_dims) - 1:
            raise ValueError(f"expected {2 * len(self.layer_dims) - 1} arrays [W1,b1,...,Wout]")
        nu = weights[-1].shape[1]
        in_w = weights[0].shape[0]
        nx2 = in_w - (2 if self._with_uprev else 1) * nu
        if nx2 <= 0 or nx2 % 2:
            raise ValueError("first kernel height is not 2Nx+kNu")
        self._nx, self._nu = nx2 // 2, nu
        self._install(weights)

    def _install(self, weights):
        L = _lib.lib()
        nl = len(self.layer_dims)
        kernels = [_lib.host(weights[2 * i]) for i in range(nl - 1)] + [_lib.host(weights[-1])]
        biases = [_lib.host(weights[2 * i + 1]) for i in range(nl - 1)]
        dims = [kernels[0].shape[0]] + [k.shape[1] for k in kernels]
        for i, k in enumerate(kernels):
            if k.shape[0] != dims[i]:
                raise ValueError(f"kernel {i} has shape {k.shape}, expected ({dims[i]}, .)")
        if dims[1:] != self.layer_dims:
            raise ValueError(f"kernel widths {dims[1:]} do not match layer_dims {self.layer_dims}")
        import torch
        if not torch.cuda.is_available():
            raise _lib.NnmpcError("no CUDA device visible: this package has no CPU fallback")
        idx = None if self._device is None else torch.device(self._device).index
        dev = torch.cuda.current_device() if idx is None else idx
        self._free()
        warr = (C.c_void_p * nl)(*[k.ctypes.data for k in kernels])
        barr = (C.c_void_p * nl)(*([b.ctypes.data for b in biases] + [None]))
        darr = (C.c_int * (nl + 1))(*dims)
        hnd = C.c_void_p()
        rc = L.nnmpc_mlp_create(C.byref(hnd), self._nx, self._nu, int(self._with_uprev), nl, darr, warr, barr, dev)
        _lib.check(rc, "nnmpc_mlp_create")
        _lib.check(L.nnmpc_mlp_set_precision(hnd, self._MODES[self.precision]), "nnmpc_mlp_set_precision")
        self._handle, self._dev = hnd, dev
        self._weights = [np.array(w, dtype=np.float64) for w in weights]
        self._weights_stale, self._train_steps = False, 0      # a new handle starts with fresh Adam moments

    def _free(self):
        if getattr(self, "_handle", None):
            _lib.lib().nnmpc_mlp_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    # -- forward ----------------------------------------------------------------------------
    def forward(self, x, uprev, xs, us, *, xscale=None, ulb=None, uub=None):
        """Batched structured forward; optional state scaling and output clip reproduce the NumPy
        deployment form (controller_evaluation.py:863-892)."""
        L = _lib.lib()
        if self._handle is None:
            self.build(int(np.shape(x)[1]), int(np.shape(us)[1]))
        if isinstance(x, np.ndarray):
            x, xs, us = _lib.host(x), _lib.host(xs), _lib.host(us)
            uprev = _lib.host(uprev) if self._with_uprev else None
            B = x.shape[0]
            out = np.empty((B, self._nu))
            sc = None if xscale is None else _lib.host(np.ravel(xscale))
            lb = None if ulb is None else _lib.host(np.ravel(ulb))
            ub = None if uub is None else _lib.host(np.ravel(uub))
            rc = L.nnmpc_mlp_forward_host(self._handle, B, _lib.hptr(x), _lib.hptr(uprev), _lib.hptr(xs),
                                          _lib.hptr(us), _lib.hptr(sc), _lib.hptr(lb), _lib.hptr(ub), _lib.hptr(out))
            _lib.check(rc, "nnmpc_mlp_forward_host")
            return out
        import torch
        f64 = dict(dtype=torch.float64, device=x.device)
        x, xs, us = x.contiguous(), xs.contiguous(), us.contiguous()
        uprev = uprev.contiguous() if self._with_uprev else None
        sc = None if xscale is None else torch.as_tensor(np.ravel(xscale) if isinstance(xscale, np.ndarray)
                                                          else xscale, **f64).contiguous()
        lb = None if ulb is None else torch.as_tensor(np.ravel(ulb) if isinstance(ulb, np.ndarray) else ulb,
                                                      **f64).contiguous()
        ub = None if uub is None else torch.as_tensor(np.ravel(uub) if isinstance(uub, np.ndarray) else uub,
                                                      **f64).contiguous()
        out = torch.empty((x.shape[0], self._nu), **f64)
        dv = self._dev
        if tuple(x.shape) != tuple(xs.shape) or tuple(us.shape) != (x.shape[0], self._nu):
            raise ValueError("x, xs must be (B, Nx) and us (B, Nu)")
        rc = L.nnmpc_mlp_forward(self._handle, x.shape[0], _lib.dptr(x, device=dv), _lib.dptr(uprev, device=dv),
                                 _lib.dptr(xs, device=dv), _lib.dptr(us, device=dv), _lib.dptr(sc, device=dv),
                                 _lib.dptr(lb, device=dv), _lib.dptr(ub, device=dv), _lib.dptr(out, device=dv),
                                 _lib.stream_ptr(dv))
        _lib.check(rc, "nnmpc_mlp_forward")
        return out

    def get_config(self):
        return dict(layer_dims=list(self.layer_dims), trainable=self.trainable, name=self.name)

    # -- training (cdu_train.py:24-62: compile(optimizer='adam', loss='mean_squared_error') + fit) ------------------
    def _train_call(self, inputs, u, step, apply, lr, beta_1, beta_2, epsilon):
        import torch
        L = _lib.lib()
        if self._with_uprev:
            x, uprev, xs, us = inputs
        else:
            (x, xs, us), uprev = inputs, None
        if self._handle is None:
            self.build(int(np.shape(x)[1]), int(np.shape(us)[1]))
        dev = torch.device("cuda", self._dev)
        f64 = dict(dtype=torch.float64, device=dev)
        t = lambda a: None if a is None else torch.as_tensor(a, **f64).contiguous()
        xt, upt, xst, ust, ut = t(x), t(uprev), t(xs), t(us), t(u)
        Bn = xt.shape[0]
        if tuple(ut.shape) != (Bn, self._nu) or tuple(ust.shape) != (Bn, self._nu) or tuple(xst.shape) != tuple(xt.shape):
            raise ValueError("x, xs must be (B, Nx); us, u (B, Nu)")
        loss = C.c_double()
        dv = self._dev
        rc = L.nnmpc_mlp_train_step(self._handle, Bn, _lib.dptr(xt, device=dv), _lib.dptr(upt, device=dv),
                                    _lib.dptr(xst, device=dv), _lib.dptr(ust, device=dv), _lib.dptr(ut, device=dv),
                                    float(lr), float(beta_1), float(beta_2), float(epsilon), int(step), int(apply),
                                    C.byref(loss), _lib.stream_ptr(dv))
        _lib.check(rc, "nnmpc_mlp_train_step")
        if apply:
            self._weights_stale = True
        return loss.value

    def train_on_batch(self, inputs, u, *, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        """One Adam step on the mean-squared error against the MPC inputs ``u`` (Keras defaults).  Returns the loss
        of the batch before the update, like ``keras.Model.train_on_batch``."""
        self._train_steps = getattr(self, "_train_steps", 0) + 1
        return self._train_call(inputs, u, self._train_steps, 1, lr, beta_1, beta_2, epsilon)

    def evaluate(self, inputs, u, batch_size=8192):
        """Mean-squared error over the whole set (float64 forward)."""
        n = np.shape(u)[0]
        tot = 0.0
        for i in range(0, n, batch_size):
            sl = slice(i, min(n, i + batch_size))
            tot += self._train_call([a[sl] for a in inputs], u[sl], 1, 0, 0.0, 0.9, 0.999, 1e-7) * (sl.stop - sl.start)
        return tot / n

    def _sync_weights(self):
        """Trained kernels / biases back from the device (Keras get_weights() order)."""
        if not getattr(self, "_weights_stale", False):
            return
        nl = len(self.layer_dims)
        ws = [np.empty_like(self._weights[2 * i]) for i in range(nl - 1)] + [np.empty_like(self._weights[-1])]
        bs = [np.empty_like(self._weights[2 * i + 1]) for i in range(nl - 1)]
        warr = (C.c_void_p * nl)(*[w.ctypes.data for w in ws])
        barr = (C.c_void_p * nl)(*([b.ctypes.data for b in bs] + [None]))
        _lib.check(_lib.lib().nnmpc_mlp_get_weights(self._handle, warr, barr), "nnmpc_mlp_get_weights")
        out = []
        for i in range(nl - 1):
            out += [ws[i], bs[i]]
        self._weights = out + [ws[-1]]
        self._weights_stale = False


class RegulatorLayerWithUprev(_StructuredRegulatorLayer):
    """inputs = [x, uprev, xs, us]  (LinearMPCLayers.py:40-61)."""
    _with_uprev = True

    def call(self, inputs):
        x, uprev, xs, us = inputs
        return self.forward(x, uprev, xs, us)

    __call__ = call


class RegulatorLayerWithoutUprev(_StructuredRegulatorLayer):
    """inputs = [x, xs, us]  (LinearMPCLayers.py:91-112)."""
    _with_uprev = False

    def call(self, inputs):
        x, xs, us = inputs
        return self.forward(x, None, xs, us)

    __call__ = call


class RegulatorModel:
    """Keras-functional-model look-alike (LinearMPCLayers.py:117-133).  Note the reference drops
    ``regulator_dims[0]`` (``layer_dims = regulator_dims[1:]``, :128-131); so does this class."""

    def __init__(self, Nx, Nu, regulator_dims, nnwithuprev=True, *, device=None, seed=None, precision=None):
        self.Nx, self.Nu, self.nnwithuprev = Nx, Nu, nnwithuprev
        cls = RegulatorLayerWithUprev if nnwithuprev else RegulatorLayerWithoutUprev
        self.regulator = cls(layer_dims=regulator_dims[1:], device=device, seed=seed, precision=precision)
        self.regulator.build(Nx, Nu)
        self.input_names = ["x", "uprev", "xs", "us"] if nnwithuprev else ["x", "xs", "us"]

    def __call__(self, inputs):
        return self.regulator(inputs)

    predict = __call__

    def get_weights(self):
        return self.regulator.get_weights()

    def set_weights(self, weights):
        self.regulator.set_weights(weights)

    # -- Keras-style training surface used by cdu_train.py / cstrs_train.py (:33-35, :52-57) -----------------------
    def compile(self, optimizer="adam", loss="mean_squared_error"):
        if optimizer != "adam" or loss not in ("mean_squared_error", "mse"):
            raise NotImplementedError("only optimizer='adam', loss='mean_squared_error' (what the reference compiles)")
        self._compiled = True

    def train_on_batch(self, x, y, **adam):
        return self.regulator.train_on_batch(x, y[0] if isinstance(y, (list, tuple)) else y, **adam)

    def evaluate(self, x, y, batch_size=8192):
        return self.regulator.evaluate(x, y[0] if isinstance(y, (list, tuple)) else y, batch_size=batch_size)

    def fit(self, x, y, epochs=1, batch_size=32, validation_split=0.0, shuffle=True, seed=0, restore_best=True,
            verbose=0):
        """Minimal ``keras.Model.fit``: the LAST ``validation_split`` fraction of the samples is held out (Keras
        takes it before shuffling), mini-batches are reshuffled every epoch, and - the reference's ModelCheckpoint
        with monitor='val_loss', save_best_only=True followed by load_weights (cdu_train.py:44-49, :98) - the
        weights of the epoch with the best validation loss are restored at the end.  Returns
        ``{"loss": [...], "val_loss": [...]}``."""
        y = y[0] if isinstance(y, (list, tuple)) else y
        x = [np.asarray(a, dtype=np.float64) for a in x]
        y = np.asarray(y, dtype=np.float64)
        n = y.shape[0]
        nval = int(n * validation_split)
        ntr = n - nval
        xv, yv = [a[ntr:] for a in x], y[ntr:]
        rng = np.random.default_rng(seed)
        hist = {"loss": [], "val_loss": []}
        best, best_w = np.inf, None
        for ep in range(epochs):
            order = rng.permutation(ntr) if shuffle else np.arange(ntr)
            tot = 0.0
            for i in range(0, ntr, batch_size):
                idx = order[i:i + batch_size]
                tot += self.train_on_batch([a[idx] for a in x], y[idx]) * len(idx)
            hist["loss"].append(tot / ntr)
            if nval:
                vl = self.evaluate(xv, yv)
                hist["val_loss"].append(vl)
                if restore_best and vl < best:
                    best, best_w = vl, self.get_weights()
            if verbose:
                print(f"Epoch {ep + 1}/{epochs} - loss: {hist['loss'][-1]:.6e}" +
                      (f" - val_loss: {hist['val_loss'][-1]:.6e}" if nval else ""))
        if restore_best and best_w is not None:
            self.set_weights(best_w)
        return hist
