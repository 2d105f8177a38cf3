"""`custom/nonlinearities.py:4-16` mirror."""
from ..nonlinearities import *          # noqa: F401,F403
from ..nonlinearities import select_nonlinearity   # noqa: F401
