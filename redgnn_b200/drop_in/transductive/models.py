"""Put this directory FIRST on sys.path and the reference's Static/transductive/base_model.py
(`from models import ...`) picks up the B200 implementation unchanged."""
from redgnn_b200.transductive.models import *  # noqa: F401,F403
