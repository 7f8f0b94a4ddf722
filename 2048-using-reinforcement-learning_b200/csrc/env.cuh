// env.cuh -- one Game2048Env.step (environment/game_2048.py:170-210) on a packed board.
#pragma once
#include "board.cuh"

namespace g2048 {

struct EnvState {
    Board board;
    int32_t score;        // cumulative merge score (np.int32 in the reference)
    uint32_t highest;     // log2(env.highest_tile), 0 when highest_tile == 0
    uint32_t spawn_ctr;   // spawns drawn so far from the env stream
};

struct StepResult {
    double reward;
    uint32_t score_delta;
    bool valid;
    bool done;
};

// Game2048Env.reset (env:29-48): two spawns on an empty board.
__device__ __forceinline__ void env_reset(EnvState &s, uint32_t k0, uint32_t k1, uint32_t game)
{
    s.board = Board(0u, 0u);
    s.score = 0;
    SpawnWords a = spawn_words(k0, k1, game, 0u, DOM_ENV, s.spawn_ctr);
    place_tile(s.board, a.pos, a.val);
    SpawnWords b = spawn_words(k0, k1, game, 0u, DOM_ENV, s.spawn_ctr + 1u);
    place_tile(s.board, b.pos, b.val);
    s.spawn_ctr += 2u;
    s.highest = max_exponent(s.board);
}

// One step.  kReward: compute the float64 shaped reward (env:212-277).
// inject: nullptr or the two raw words to use instead of the env stream.
template <bool kRowShared, bool kCodeShared, bool kReward>
__device__ __forceinline__ StepResult env_step(EnvState &s, uint32_t action, const uint16_t *row,
                                               const uint8_t *code, uint32_t k0, uint32_t k1, uint32_t game,
                                               const uint32_t *inject, unsigned long long *overflow)
{
    StepResult r;
    const Board prev = s.board;
    const Board line = to_line(prev, action);
    Board next = from_line(move_left<kRowShared>(line, row), action);          // env:185
    uint32_t codes = merge_codes<kCodeShared>(line, code);
    const bool in_range = action < 4u;                                      // env:99-114 has no else branch
    next = select(in_range, next, prev);
    codes = in_range ? codes : 0u;
    uint32_t sat;
    r.score_delta = decode_score(codes, &sat);
    if (sat & 0x10000u) atomicAdd(overflow, 1ull);                          // 32768+32768: nibble saturated
    s.score += (int32_t)r.score_delta;
    r.valid = next != prev;                                                 // env:188
    const int empty_before = kReward ? count_empty(prev) : 0;
    int empty_after = count_empty(next);
    if (r.valid) {                                                          // env:191-192, always >= 1 empty here
        SpawnWords w;
        if (inject) { w.pos = inject[0]; w.val = inject[1]; }
        else { w = spawn_words(k0, k1, game, 0u, DOM_ENV, s.spawn_ctr); s.spawn_ctr += 1u; }
        place_tile(next, w.pos, w.val);
        empty_after -= 1;
    }
    r.reward = 0.0;
    if (kReward)                                                            // env:195, uses the OLD highest_tile
        r.reward = shaped_reward(r.valid, empty_before, next, empty_after, r.score_delta, s.highest,
                                 max_exponent(prev));
    // env:198: a board with both tiles and empty cells always has a legal move; only a full
    // (or a tile-less) board needs the neighbour test
    r.done = (empty_after == 0 || empty_after == 16) && env_game_over(next);
    s.highest = max(s.highest, max_exponent(next));                         // env:200-203
    s.board = next;
    return r;
}

}  // namespace g2048
