"""Puts the repository root on sys.path so that `g2048_b200` resolves from the drop-in shims."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
