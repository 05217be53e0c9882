"""emcee stand-in -- TEST INFRASTRUCTURE: the 2.x ``EnsembleSampler`` API slice of ``_sampler.py:270-274``."""
from bode_b200._emcee import EnsembleSampler  # noqa: F401

__version__ = "0-stub"
