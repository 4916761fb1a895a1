"""Drop-in for reference Static/transductive/load_data.py (DataLoader)."""
from ..data import TransductiveLoader as DataLoader

__all__ = ["DataLoader"]
