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
