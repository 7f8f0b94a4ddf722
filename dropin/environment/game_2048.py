"""Drop-in for the reference's `environment/game_2048.py`: same module path, same class name.

Put `dropin/` ahead of the reference checkout on sys.path (PYTHONPATH=/path/to/repo/dropin) and the
reference's own scripts (`train.py`, `evaluate_beam_search.py`, `agents/ppo_agent.py`, ...) import the
GPU engine instead of the NumPy environment, unchanged.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from g2048_b200 import Game2048Env  # noqa: E402,F401
