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


class BatchedOnlineSimulation:
    """S closed-loop scenarios advanced in lock step on the GPU: the batched form of ``online_simulation`` with a
    ``LinearMPCController`` / ``NeuralNetworkController`` / ``SatDlqrController`` in the loop, as the reference's
    validation study runs them one scenario after the other (controller_evaluation.py:322-523; control laws
    lib/linearMPC.py:646-669, controller_evaluation.py:841-862, :975-993).

    Plant: ``x+ = A x + B u + Bp p``, ``y = C x + v`` (LinearPlantSimulator, linearMPC.py:87-131); the controller
    arguments are those of ``LinearMPCController`` (:525-528).  ``kind``: "mpc", "nn" (needs ``regulator_weights``,
    ``xscale``, ``nnwithuprev``) or "satdlqr".
    """
    KINDS = {"mpc": 0, "nn": 1, "satdlqr": 2}

    def __init__(self, *, kind, A, B, C, H, Qwx, Qwd, Rv, xprior, dprior, Rs, Qs, Bd, Cd, usp, uprev, Q, R, S, ulb, uub,
                 N=None, Bp=None, regulator_weights=None, xscale=None, nnwithuprev=True, device=None, precision=None,
                 **solver_kwargs):
        import ctypes as Ct
        from . import _lib
        from .linearMPC import LinearMPCController, dlqr, _device_index
        if kind not in self.KINDS:
            raise ValueError(f"kind must be one of {sorted(self.KINDS)}")
        self.kind = kind
        self.Nx, self.Nu, self.Ny, self.Nd = A.shape[0], B.shape[1], C.shape[0], Bd.shape[1]
        Bp = Bd if Bp is None else Bp
        self.Np = Bp.shape[1]
        self.xprior, self.dprior, self.uprev0 = xprior, dprior, uprev
        self.C, self.Rv = C, Rv
        self._dev = _device_index(device)
        self.filter = LinearMPCController.setup_filter(A=A, B=B, C=C, Bd=Bd, Cd=Cd, Qwx=Qwx, Qwd=Qwd, Rv=Rv,
                                                       xprior=xprior, dprior=dprior)
        f = self.filter
        ILC = np.eye(self.Nx + self.Nd) - f.L @ f.C
        Fkf = np.hstack([ILC @ f.A, ILC @ f.B, f.L])
        Fpl = np.block([[A, B, Bp], [C @ A, C @ B, C @ Bp]])
        self.target_selector = LinearMPCController.setup_target_selector(A=A, B=B, C=C, H=H, Bd=Bd, Cd=Cd, usp=usp, Qs=Qs,
                                                                         Rs=Rs, ulb=ulb, uub=uub, device=self._dev)
        Aaug, Baug, self.Qaug, self.Raug, self.Maug = LinearMPCController.get_augmented_matrices_for_regulator(A, B, Q, R, S)
        self.regulator = self.layer = None
        Kaug = None
        if kind == "mpc":
            self.regulator = LinearMPCController.setup_regulator(A=A, B=B, Q=Q, R=R, S=S, N=N, ulb=ulb, uub=uub,
                                                                 device=self._dev, **solver_kwargs)
        elif kind == "nn":
            ctl = NeuralNetworkController(regulator_weights=regulator_weights, xscale=xscale, nnwithuprev=nnwithuprev,
                                          ulb=ulb, uub=uub, device=self._dev, precision=precision)
            self.layer, self._nn = ctl.layer, ctl
        else:
            Kaug, _ = dlqr(Aaug, Baug, self.Qaug, self.Raug, self.Maug)      # controller_evaluation.py:958-960
        self.Kaug = Kaug
        L = _lib.lib()
        hnd = Ct.c_void_p()
        hp = lambda a: None if a is None else _lib.hptr(_lib.host(a))
        keep = [_lib.host(a) for a in (Fkf, Fpl, self.Qaug, self.Raug, self.Maug, ulb, uub)]
        kk = None if Kaug is None else _lib.host(Kaug)
        xsc = None if xscale is None or kind != "nn" else _lib.host(np.ravel(xscale))
        rc = L.nnmpc_online_create(Ct.byref(hnd), self.regulator._handle if self.regulator else None,
                                   self.target_selector._handle, self.layer._handle if self.layer else None,
                                   self.Nx, self.Nu, self.Ny, self.Nd, self.Np, _lib.hptr(keep[0]), _lib.hptr(keep[1]),
                                   None if kk is None else _lib.hptr(kk), _lib.hptr(keep[2]), _lib.hptr(keep[3]),
                                   _lib.hptr(keep[4]), _lib.hptr(keep[5]), _lib.hptr(keep[6]),
                                   None if xsc is None else _lib.hptr(xsc), self._dev)
        _lib.check(rc, "nnmpc_online_create")
        self._handle = hnd

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                from . import _lib
                _lib.lib().nnmpc_online_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def run(self, setpoints, disturbances, *, noise=None, x0=None, seed=None, tol=1e-9, max_iter=20000):
        """setpoints (S,T,Ny), disturbances (S,T,Np).  ``noise`` (S,T+1,Ny) measurement noise; default: drawn like
        the reference does - ``np.random.seed(seed)`` before every scenario, ``std * randn(Ny,1)`` at construction
        and after every step (controller_evaluation.py:347, linearMPC.py:103, :119).  Returns a dict of NumPy arrays:
        u (S,T,Nu), y (S,T+1,Ny), x (S,T+1,Nx), xhat, xs (S,T,Nx), us (S,T,Nu), average_stage_costs (S,T) and, for
        the MPC, iters / kkt (S,T)."""
        import torch
        from . import _lib
        L = _lib.lib()
        sp, ds = _lib.host(setpoints), _lib.host(disturbances)
        S_, T = sp.shape[0], sp.shape[1]
        if tuple(sp.shape) != (S_, T, self.Ny) or tuple(ds.shape) != (S_, T, self.Np):
            raise ValueError(f"setpoints / disturbances must have shapes (S,T,{self.Ny}) / (S,T,{self.Np})")
        if noise is None:
            std = np.sqrt(np.diag(self.Rv))
            noise = np.empty((S_, T + 1, self.Ny))
            for s in range(S_):
                if seed is not None:
                    np.random.seed(seed)
                for t in range(T + 1):
                    noise[s, t] = std * np.random.randn(self.Ny)
        noise = _lib.host(noise)
        x0 = np.zeros((S_, self.Nx)) if x0 is None else np.broadcast_to(np.asarray(x0, float).reshape(-1, self.Nx), (S_, self.Nx))
        dev = torch.device("cuda", self._dev)
        f64 = dict(dtype=torch.float64, device=dev)
        t_ = lambda a: torch.as_tensor(np.ascontiguousarray(a), **f64)
        x_io = t_(x0)
        xhat_io = t_(np.tile(np.vstack([self.xprior, self.dprior]).T, (S_, 1)))
        up_io = t_(np.tile(np.asarray(self.uprev0, float).reshape(1, -1), (S_, 1)))
        y = torch.zeros((S_, T + 1, self.Ny), **f64)
        y[:, 0] = t_(x0 @ self.C.T + noise[:, 0])
        out = dict(u=torch.empty((S_, T, self.Nu), **f64), x=torch.empty((S_, T + 1, self.Nx), **f64),
                   xhat=torch.empty((S_, T, self.Nx), **f64), xs=torch.empty((S_, T, self.Nx), **f64),
                   us=torch.empty((S_, T, self.Nu), **f64), average_stage_costs=torch.empty((S_, T), **f64))
        iters = torch.zeros((S_, T), dtype=torch.int32, device=dev)
        kkt = torch.zeros((S_, T), **f64)
        d = lambda a: _lib.dptr(a, device=self._dev)
        sp_d, ds_d, noise_d = t_(sp), t_(ds), t_(noise)       # named: the device copies must outlive the call
        rc = L.nnmpc_online_run(self._handle, self.KINDS[self.kind], S_, T, d(x_io), d(xhat_io), d(up_io), d(sp_d),
                                d(ds_d), d(noise_d), d(y), d(out["u"]), d(out["x"]), d(out["xhat"]), d(out["xs"]),
                                d(out["us"]), d(out["average_stage_costs"]), _lib.dptr_i32(iters, self._dev), d(kkt),
                                float(tol), int(max_iter), _lib.stream_ptr(self._dev))
        hit = _lib.check(rc, "nnmpc_online_run")
        res = {k: v.cpu().numpy() for k, v in out.items()}
        res.update(y=y.cpu().numpy(), xhat_final=xhat_io.cpu().numpy(), maxiter_hit=hit)
        if self.kind == "mpc":
            res.update(iters=iters.cpu().numpy(), kkt=kkt.cpu().numpy())
        return res

    @staticmethod
    def performance_loss(controller_result, mpc_result):
        """100 (Lambda_controller - Lambda_mpc) / Lambda_mpc per scenario (controller_evaluation.py:396-397), Lambda =
        the final average stage cost."""
        lc = controller_result["average_stage_costs"][:, -1]
        lm = mpc_result["average_stage_costs"][:, -1]
        return 100.0 * (lc - lm) / lm
