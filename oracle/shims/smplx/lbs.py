import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.flame_oracle import lbs_torch as lbs  # noqa: E402,F401
