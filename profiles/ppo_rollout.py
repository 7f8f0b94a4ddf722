"""BASELINE config 3: PPO rollout collection -- 16,384 on-device envs x 128 steps feeding a PyTorch
transformer policy (same shape as the reference's models/transformer.py: Linear(1->64) per cell,
2 x TransformerEncoderLayer(d=64, 4 heads), FC 1024->128->64, actor/critic heads; random init).
The policy stays in PyTorch (out of scope); this script measures the env side and the loop.

usage: python profiles/ppo_rollout.py [envs] [steps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import g2048_b200 as G


class BoardTransformerPolicy(nn.Module):
    def __init__(self, d=64, heads=4, layers=2):
        super().__init__()
        self.embed = nn.Linear(1, d)
        self.encoder = nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model=d, nhead=heads, batch_first=True), layers)
        self.trunk = nn.Sequential(nn.Linear(16 * d, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU())
        self.actor, self.critic = nn.Linear(64, 4), nn.Linear(64, 1)

    def forward(self, obs):                      # obs float32[N,16]
        x = self.encoder(self.embed(obs.unsqueeze(-1)))
        h = self.trunk(x.flatten(1))
        return self.actor(h), self.critic(h)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = "cuda:0"
    torch.manual_seed(0)
    policy = BoardTransformerPolicy().to(dev).eval()
    env = G.BatchedGame2048Env(n, dev, seed=7)
    env.reset()

    def collect(with_policy):
        obs_buf = torch.empty(steps, n, 16, device=dev)
        rew_buf = torch.empty(steps, n, device=dev, dtype=torch.float32)
        legal = env.legal_masks()
        bits = torch.tensor([1, 2, 4, 8], device=dev, dtype=torch.uint8)
        for t in range(steps):
            obs = env.observe()
            obs_buf[t] = obs
            if with_policy:
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    logits, _ = policy(obs)
                mask = (legal.unsqueeze(1) & bits) == 0                      # illegal moves -> -inf (ppo_agent.py:214-219)
                logits = logits.float().masked_fill(mask & (legal.unsqueeze(1) != 0), float("-inf"))
                actions = torch.distributions.Categorical(logits=logits).sample().to(torch.uint8)
            else:
                actions = torch.randint(0, 4, (n,), device=dev, dtype=torch.uint8)
            _, _, done, info = env.step(actions, auto_reset=True)           # step + `if done: env.reset()` in one launch
            rew_buf[t] = info["reward32"]
            legal = info["legal_mask"]                                        # of the (possibly fresh) board
        return obs_buf, rew_buf

    out = {"envs": n, "steps": steps}
    for name, with_policy in (("env_only_random_actions", False), ("with_transformer_policy_bf16", True)):
        collect(with_policy)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); collect(with_policy); e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        out[name] = {"ms_per_update": ms, "board_steps_per_s": n * steps / ms * 1e3}
    # the env side of one rollout step as a CUDA graph: observe -> step(static actions) -> reset_done -> legal
    acts = torch.randint(0, 4, (n,), device=dev, dtype=torch.uint8)
    obs_static = torch.empty(n, 16, device=dev)

    def env_side():
        obs_static.copy_(env.observe())
        env.step(acts, auto_reset=True)

    graph = env.graph(env_side)
    for _ in range(10):
        graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        graph.replay()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    out["env_side_cuda_graph"] = {"ms_per_update": ms, "board_steps_per_s": n * steps / ms * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
