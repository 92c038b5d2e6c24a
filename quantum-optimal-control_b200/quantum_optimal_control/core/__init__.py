from .problem import SystemParameters, choose_taylor_terms  # noqa: F401
from .defaults import Convergence, CONVERGENCE_DEFAULTS  # noqa: F401
from .engine import GrapeEngine, QocError, load_library  # noqa: F401
from .optimizer import run_session, TF1AdamState  # noqa: F401
