"""Drop-in for the reference's `agents/beam_search_agent.py`: same module path, same class name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from g2048_b200 import BeamSearchAgent  # noqa: E402,F401
from g2048_b200 import Game2048Env  # noqa: E402,F401  (the reference module imports it too, agent:4)
