"""Plant / tuning / scenario builders (the *inputs* of the hot path).

``cstrs`` restates /root/reference/cstrs_parameters.py numerically (no casadi/mpctools);
``cdu`` is a documented synthetic stand-in for the crude-distillation model whose
``CDU_Model.mat`` is not shipped with the reference (cdu_parameters.py:200).
"""
from .cstrs import get_cstrs_problem            # noqa: F401
from .cdu import get_cdu_problem                # noqa: F401
from .problem import MPCProblem                 # noqa: F401
