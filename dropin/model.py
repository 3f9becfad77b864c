"""`from model import *` shim: put this directory on sys.path instead of the reference's `models/` and the
reference training scripts (experiments/new_vae.py:13, new_gan.py:22, new_betavaegan.py:18) pick up the
B200 kernel-backed classes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disentangle_mlp_b200.model import *  # noqa: F401,F403,E402
from disentangle_mlp_b200.model import __all__  # noqa: F401,E402
