"""CPU oracle for the batched 2048 engine -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (2048-using-reinforcement-learning_b200/)
never does and fails loudly when its CUDA library is missing.
"""
