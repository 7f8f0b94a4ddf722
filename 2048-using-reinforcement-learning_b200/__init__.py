"""B200-native batched 2048 engine: the env transition and beam-search hot path of
vivek-tiwari-vt/2048-Using-Reinforcement-Learning behind the reference's own Python API.

Importable as `g2048_b200` (see g2048_b200.py at the repo root; the directory name is not a
valid Python identifier).  Importing does not need a GPU; running anything does.
"""
from ._lib import G2048Error, build_library, launch_count, overflow_count, EXPORTS, LIB_PATH  # noqa: F401
from .packing import pack_board, pack_boards, unpack_board, unpack_boards  # noqa: F401
from .env import Game2048Env, BatchedGame2048Env  # noqa: F401
from .beam import BeamSearchAgent, BatchedBeamSearch  # noqa: F401
from .ppo import PPORewardShaper  # noqa: F401
from .hybrid import HybridBeamSearch  # noqa: F401
from .parallel import shard_range, all_reduce_stats, describe_stats  # noqa: F401
from .evaluation import run_evaluation, compile_results, write_overall_results, best_games  # noqa: F401

__all__ = ["Game2048Env", "BatchedGame2048Env", "BeamSearchAgent", "BatchedBeamSearch",
           "pack_board", "pack_boards", "unpack_board", "unpack_boards",
           "PPORewardShaper", "HybridBeamSearch", "shard_range", "all_reduce_stats", "describe_stats", "run_evaluation", "G2048Error", "build_library"]
