"""Batched driver of the hybrid agent's beam search: DQNAgent.beam_search
(agents/hybrid.py:814-907) for many boards at once.  Per level: ONE expansion launch pair over every
(game, beam entry, action) item (g2048_hybrid_expand_items), at most ONE Q-network call over every leaf
that needs a value, and a stable per-game top-k.  The Q-network stays in PyTorch (any module mapping
float32[N,16] tile values to float32[N,4]).

`reference_early_exit=True` (default) keeps hybrid.py:871 as written -- `all(done for _, _, _, done in beam)`
tests the probability field, which is always truthy, so the reference's loop ends after its FIRST level;
False runs the loop as it reads (`search_depth` levels, Q-values at the last level).
Arithmetic follows the reference: rewards, cumulative rewards, probabilities and sort keys are float64,
ties keep generation order (Python's stable sort), the action scores are summed in beam order.
"""
from __future__ import annotations

from . import _lib


class HybridBeamSearch:
    def __init__(self, model, beam_width=15, search_depth=30, gamma=0.99, beam_search_threshold=64,
                 device="cuda:0", seed=0, reference_early_exit=True):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.G2048Error("HybridBeamSearch needs a CUDA device (no CPU fallback)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)
        self.model = model.to(self.device).eval()
        self.beam_width, self.search_depth = int(beam_width), int(search_depth)
        self.gamma, self.threshold = float(gamma), int(beam_search_threshold)
        self.seed = int(seed) & (2**64 - 1)
        self.early_exit = bool(reference_early_exit)

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _values(self, boards):
        """packed int64[N] -> float32[N,16] tile values (what the reference feeds its Q-network)"""
        t = self.torch
        out = t.empty(boards.numel(), 16, dtype=t.int32, device=self.device)
        if boards.numel():
            _lib.check(_lib.use_device(self.index).g2048_unpack(boards.data_ptr(), out.data_ptr(), boards.numel(), self._stream()))
        return out.to(t.float32)

    def _expand(self, boards, actions, game, call, draw0):
        """One launch over all items -> (next_boards[n,8], reward[n,8], done[n,8], count[n], draws[n])."""
        t = self.torch
        n = boards.numel()
        z = dict(device=self.device)
        nb = t.zeros(n, 8, dtype=t.int64, **z); rw = t.zeros(n, 8, dtype=t.float64, **z)
        dn = t.zeros(n, 8, dtype=t.uint8, **z); cnt = t.zeros(n, dtype=t.int32, **z); used = t.zeros(n, dtype=t.int32, **z)
        _lib.check(_lib.use_device(self.index).g2048_hybrid_expand_items(
            boards.data_ptr(), actions.data_ptr(), game.data_ptr(), call.data_ptr(), draw0.data_ptr(), nb.data_ptr(),
            rw.data_ptr(), dn.data_ptr(), cnt.data_ptr(), used.data_ptr(), n, self.seed, self.game0, self._stream()))
        return nb, rw, dn, cnt, used

    def get_actions(self, boards, call=0, game0=0):
        """boards: packed int64[G] (device).  call: int or int32[G] (index of the agent's call, addresses the
        sampling stream of game game0 + g).  Returns (actions int64[G], scores float64[G,4] -- the action_scores
        of hybrid.py:881-890, NaN where an action is absent or the Q-network path was taken)."""
        t = self.torch
        G = boards.numel()
        dev = self.device
        self.game0 = int(game0)
        lib = _lib.use_device(self.index)
        call_t = call.to(t.int32) if t.is_tensor(call) else t.full((G,), int(call), dtype=t.int32, device=dev)
        values = self._values(boards.contiguous())
        actions = t.zeros(G, dtype=t.int64, device=dev)
        scores = t.full((G, 4), float("nan"), dtype=t.float64, device=dev)
        # hybrid.py:821-834: simple boards go straight to the Q-network, invalid moves masked
        simple = (values.max(dim=1).values < self.threshold) | ((values > 0).sum(dim=1) < 8)
        if bool(simple.any()):
            idx = simple.nonzero().squeeze(1)
            with t.no_grad():
                q = self.model(values[idx]).to(t.float32).clone()
            legal = t.empty(G, dtype=t.uint8, device=dev)
            _lib.check(lib.g2048_legal_masks(boards.data_ptr(), legal.data_ptr(), 0, G, self._stream()))
            bits = (legal[idx].to(t.int64).unsqueeze(1) >> t.arange(4, device=dev)) & 1
            q[bits == 0] = -1e9
            actions[idx] = q.argmax(dim=1)
        todo = (~simple).nonzero().squeeze(1)
        if todo.numel() == 0:
            return actions, scores
        g_of = todo.to(t.int32)                                   # item -> game (for the sampling stream)
        n = todo.numel()
        W = self.beam_width
        beam = boards[todo].clone().unsqueeze(1)                  # [n, B]
        first = t.full((n, 1), -1, dtype=t.int64, device=dev)
        cum = t.zeros(n, 1, dtype=t.float64, device=dev)
        prob = t.ones(n, 1, dtype=t.float64, device=dev)
        alive = t.ones(n, 1, dtype=t.bool, device=dev)
        draw = t.zeros(n, dtype=t.int32, device=dev)              # draws consumed so far in this call, per game
        for step in range(self.search_depth):
            B = beam.shape[1]
            # items in the reference's order: (beam entry, action) within a game
            ib = beam.unsqueeze(2).expand(n, B, 4).reshape(-1).contiguous()
            ia = t.arange(4, dtype=t.uint8, device=dev).repeat(n * B)
            ig = g_of.unsqueeze(1).expand(n, B * 4).reshape(-1).contiguous()
            ic = call_t[todo].unsqueeze(1).expand(n, B * 4).reshape(-1).contiguous()
            live = alive.unsqueeze(2).expand(n, B, 4).reshape(n, B * 4)
            zero = t.zeros(n * B * 4, dtype=t.int32, device=dev)
            _, _, _, _, used = self._expand(ib, ia, ig, ic, zero)     # pass 1: how many draws each item takes
            used = used.reshape(n, B * 4) * live
            offs = draw.unsqueeze(1) + used.cumsum(dim=1, dtype=t.int32) - used
            nb, rw, dn, cnt, _ = self._expand(ib, ia, ig, ic, offs.reshape(-1).to(t.int32).contiguous())
            draw = draw + used.sum(dim=1, dtype=t.int32)
            K = B * 4 * 8                                          # candidate slots per game, generation order
            cnt2 = cnt.reshape(n, B * 4)
            slot_ok = (t.arange(8, device=dev).view(1, 1, 8) < cnt2.unsqueeze(2)) & live.unsqueeze(2)
            nb = nb.reshape(n, K); rw = rw.reshape(n, K); dn = dn.reshape(n, K).to(t.bool); ok = slot_ok.reshape(n, K)
            parent_cum = cum.unsqueeze(2).expand(n, B, 32).reshape(n, K)
            parent_prob = prob.unsqueeze(2).expand(n, B, 32).reshape(n, K)
            total = parent_cum + rw
            need_value = ok & (dn | (step == self.search_depth - 1))
            if bool(need_value.any()):                             # one Q-network call per level (hybrid.py:851-855)
                sel = need_value.nonzero()
                with t.no_grad():
                    v = self.model(self._values(nb[sel[:, 0], sel[:, 1]].contiguous())).max(dim=1).values.to(t.float64)
                gv = self.gamma * v
                gv = gv * (1 - dn[sel[:, 0], sel[:, 1]].to(t.float64))
                total[sel[:, 0], sel[:, 1]] = (parent_cum + rw)[sel[:, 0], sel[:, 1]] + gv
            new_prob = parent_prob / cnt2.clamp(min=1).to(t.float64).unsqueeze(2).expand(n, B * 4, 8).reshape(n, K)
            act = t.arange(4, device=dev).view(1, 1, 4, 1).expand(n, B, 4, 8).reshape(n, K)
            parent_first = first.unsqueeze(2).expand(n, B, 32).reshape(n, K)
            new_first = t.where(parent_first >= 0, parent_first, act)
            key = t.where(ok, total * new_prob, t.full_like(total, float("-inf")))
            order = t.sort(key, dim=1, descending=True, stable=True).indices[:, :W]   # hybrid.py:867-868
            take = lambda x: x.gather(1, order)                    # noqa: E731
            beam, first, cum, prob, alive = take(nb), take(new_first), take(total), take(new_prob), take(ok)
            if self.early_exit:                                    # hybrid.py:871
                break
        # hybrid.py:881-893: scores per first action summed in beam order, first maximum in order of appearance
        B = beam.shape[1]
        contrib = t.where(alive, cum * prob, t.zeros_like(cum))
        sc = t.zeros(n, 4, dtype=t.float64, device=dev)
        seen_at = t.full((n, 4), B, dtype=t.int64, device=dev)
        rows = t.arange(n, device=dev)
        for b in range(B):
            a_b = first[:, b].clamp(min=0)
            live_b = alive[:, b]
            sc[rows, a_b] = t.where(live_b, sc[rows, a_b] + contrib[:, b], sc[rows, a_b])
            seen_at[rows, a_b] = t.where(live_b & (seen_at[rows, a_b] == B), t.full_like(a_b, b), seen_at[rows, a_b])
        present = seen_at < B
        masked = t.where(present, sc, t.full_like(sc, float("-inf")))
        best = masked.max(dim=1, keepdim=True).values
        tie_rank = t.where(masked == best, seen_at, t.full_like(seen_at, B + 1))
        actions[todo] = tie_rank.argmin(dim=1)
        scores[todo] = t.where(present, sc, t.full_like(sc, float("nan")))
        return actions, scores
