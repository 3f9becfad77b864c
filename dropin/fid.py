"""`from fid import get_fid` shim (experiments/new_betavaegan.py:21): the GPU FID of disentangle_mlp_b200.fid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disentangle_mlp_b200.fid import get_fid  # noqa: F401,E402
