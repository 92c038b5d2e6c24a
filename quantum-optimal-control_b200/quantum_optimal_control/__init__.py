"""quantum_optimal_control -- B200-native GRAPE engine behind the reference package's import path
(``from quantum_optimal_control.main_grape.grape import Grape``)."""
from .main_grape import Grape  # noqa: F401
from .helper_functions import *  # noqa: F401,F403
