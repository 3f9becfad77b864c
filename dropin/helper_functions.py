"""`from helper_functions import *` shim (experiments/new_betavaegan.py:24, new_gan.py, new_vae.py): the reference's
utils/utils.py sampling / reconstruction helpers over the kernel-backed modules."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disentangle_mlp_b200.sampling import *  # noqa: F401,F403,E402
from disentangle_mlp_b200.sampling import __all__  # noqa: F401,E402
