"""Container for one linear-MPC problem definition (matrices, tuning, bounds, scenarios)."""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np


@dataclass
class MPCProblem:
    """Everything ``LinearMPCController`` / ``OfflineSimulator`` take as keyword arguments
    (/root/reference/lib/linearMPC.py:525-528, :726-730), plus scenario signals."""
    name: str
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    H: np.ndarray
    Bd: np.ndarray
    Cd: np.ndarray
    Q: np.ndarray
    R: np.ndarray
    S: np.ndarray
    N: int
    Rs: np.ndarray
    Qs: np.ndarray
    usp: np.ndarray
    ulb: np.ndarray
    uub: np.ndarray
    xprior: np.ndarray
    uprev: np.ndarray
    setpoints: np.ndarray | None = None       # (Nsim, Ny)
    disturbances: np.ndarray | None = None    # (Nsim, Nd)
    extra: dict = field(default_factory=dict)

    @property
    def Nx(self): return self.A.shape[0]
    @property
    def Nu(self): return self.B.shape[1]
    @property
    def Ny(self): return self.C.shape[0]
    @property
    def Nd(self): return self.Bd.shape[1]

    def controller_kwargs(self):
        return dict(A=self.A, B=self.B, C=self.C, H=self.H, Rs=self.Rs, Qs=self.Qs, Bd=self.Bd,
                    Cd=self.Cd, usp=self.usp, uprev=self.uprev, Q=self.Q, R=self.R, S=self.S,
                    ulb=self.ulb, uub=self.uub, N=self.N)
