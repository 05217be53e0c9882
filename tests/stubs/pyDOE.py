"""pyDOE stand-in -- TEST INFRASTRUCTURE: ``lhs`` with pyDOE's signature and RNG consumption (bode_b200/_lhs.py)."""
from bode_b200._lhs import lhs  # noqa: F401

__all__ = ["lhs"]
