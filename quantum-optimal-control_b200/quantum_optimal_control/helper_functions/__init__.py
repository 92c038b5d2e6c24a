from .grape_functions import *  # noqa: F401,F403
