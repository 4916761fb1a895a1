"""Drop-in for reference Static/inductive/load_data.py (DataLoader)."""
from ..data import InductiveLoader as DataLoader

__all__ = ["DataLoader"]
