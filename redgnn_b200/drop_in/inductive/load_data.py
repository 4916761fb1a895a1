"""Put this directory FIRST on sys.path and the reference's Static/inductive/train.py
(`from load_data import DataLoader`) picks up the B200 implementation unchanged."""
from redgnn_b200.inductive.load_data import *  # noqa: F401,F403
