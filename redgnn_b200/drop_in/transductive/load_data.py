"""Put this directory FIRST on sys.path and the reference's Static/transductive/train.py
(`from load_data import DataLoader`) picks up the B200 implementation unchanged."""
from redgnn_b200.transductive.load_data import *  # noqa: F401,F403
