"""Drop-in alias: ``from bode import *`` as the reference drivers do (``tests/samp_ex1.py:19``)."""
from bode_b200 import *  # noqa: F401,F403
from bode_b200 import __all__  # noqa: F401
