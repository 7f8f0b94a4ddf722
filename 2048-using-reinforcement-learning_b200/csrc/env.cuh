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
__device__ __forceinline__ void env_reset(EnvState &s, const PhiloxKey &K, uint32_t game)
{
    s.board = Board(0u, 0u);
    s.score = 0;
    SpawnWords a = spawn_words(K, game, 0u, DOM_ENV, s.spawn_ctr);
    place_tile(s.board, a.pos, a.val);
    SpawnWords b = spawn_words(K, game, 0u, DOM_ENV, s.spawn_ctr + 1u);
    place_tile(s.board, b.pos, b.val);
    s.spawn_ctr += 2u;
    s.highest = max_exponent(s.board);
}

// One step.  kReward: compute the float64 shaped reward (env:212-277).
// inject: nullptr or the two raw words to use instead of the env stream.
template <bool kRowShared, bool kCodeShared, bool kReward>
__device__ __forceinline__ StepResult env_step(EnvState &s, uint32_t action, const uint16_t *row,
                                               const uint8_t *code, const PhiloxKey &K, uint32_t game,
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
        else { w = spawn_words(K, game, 0u, DOM_ENV, s.spawn_ctr); s.spawn_ctr += 1u; }
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

// The same transition for the per-step kernel (env_step_fused_kernel): merge score, tile sums and
// edge sum come from the 512-entry pair table in shared memory (board.cuh), the move itself either
// from the row tables through L1/L2 (kSwarMove = false) or table-free (move_left_half).  Bit-exact
// with env_step(): `shaped_reward_tracked` adds the reference's exact terms as integers.
// Also returns the legal mask of the new board (get_valid_moves, env:69-95); done = no legal move.
struct StepResult2 {
    double reward;
    uint32_t score_delta, legal;
    bool valid, done;
};
// `tables_ready()` is called once, by every thread, right before the first use of `pairs`: the kernel publishes its
// shared-memory copy of the tables there (a block-wide barrier), so the table-free move above it overlaps the copy.
template <bool kSwarMove, bool kReward, class TablesReady>
__device__ __forceinline__ StepResult2 env_step_pairs(EnvState &s, uint32_t action, const uint16_t *row, const uint8_t *code,
                                                      const uint32_t *pairs, const PhiloxKey &K, uint32_t game,
                                                      const uint32_t *inject, unsigned long long *overflow,
                                                      TablesReady tables_ready)
{
    StepResult2 r;
    const Board prev = s.board;
    const uint32_t prev_max = max_exponent(prev);
    const Board line = to_line(prev, action);
    const bool in_range = action < 4u;                                       // env:99-114 has no else branch
    Board moved;
    uint32_t gained;
    int empty_before;
    bool max_merged = false;                 // two tiles of the board's largest exponent merged
    if (kSwarMove) {
        const HalfMove lo = move_left_half(line.lo), hi = move_left_half(line.hi);
        moved = Board(lo.rows, hi.rows);
        tables_ready();
        gained = code_score_pairs(lo.codes, hi.codes, pairs);
        // by-products of the move: the line view is a permutation of the board, so its occupancy counts the board's
        // empty cells; a code nibble is the exponent of the two tiles that merged, and only a merge of two tiles of
        // the largest exponent raises the board's maximum (a zero nibble of codes ^ max..max; "any zero nibble" by the
        // borrow trick is exact as an any-test).  A tile-less board has no merges.
        empty_before = 16 - __popc(lo.occupied) - __popc(hi.occupied);
        const uint32_t rep = prev_max * LSB4, x = lo.codes ^ rep, y = hi.codes ^ rep;
        max_merged = in_range && prev_max != 0u && ((((x - LSB4) & ~x) | ((y - LSB4) & ~y)) & MSB4) != 0u;
    } else {
        moved = move_left<false>(line, row);
        tables_ready();
        gained = merge_score_pairs<false>(line, code, pairs);
        empty_before = count_empty(prev);
    }
    Board next = select(in_range, from_line(moved, action), prev);           // env:185
    gained = in_range ? gained : 0u;
    r.score_delta = gained & (kPairSaturated - 1u);
    if ((gained >> 28) && overflow) atomicAdd(overflow, 1ull);               // 32768+32768: nibble saturated (null: not counted)
    s.score += (int32_t)r.score_delta;
    r.valid = next != prev;                                                  // env:188
    uint32_t nzl = nz_flags(next.lo), nzh = nz_flags(next.hi);
    const uint32_t zl = nzl ^ LSB4, zh = nzh ^ LSB4;
    const int cl = __popc(zl);
    int empty_after = cl + __popc(zh);
    SpawnWords w;
    if (inject) { w.pos = inject[0]; w.val = inject[1]; }
    else w = spawn_words(K, game, 0u, DOM_ENV, s.spawn_ctr);                  // computed for every lane, used by the valid ones
    const SpawnPick sp = pick_spawn(zl, zh, cl, empty_after, w.pos, w.val);
    uint32_t spawned = 0u;
    if (r.valid) {                                                           // env:191-192, always >= 1 empty here
        next.lo |= sp.flag_lo * sp.exponent;
        next.hi |= sp.flag_hi * sp.exponent;
        nzl |= sp.flag_lo;
        nzh |= sp.flag_hi;
        if (!inject) s.spawn_ctr += 1u;
        empty_after -= 1;
        spawned = sp.exponent;
    }
    r.reward = 0.0;
    if (kReward)                                                             // env:195, uses the OLD highest_tile
        r.reward = shaped_reward_tracked(r.valid, empty_before, next, empty_after, nzl, nzh, r.score_delta, s.highest,
                                         prev_max, tile_total_pairs(next, pairs), pairs);
    r.legal = env_legal_mask_flags(next, nzl, nzh);
    r.done = r.legal == 0u;                                                  // env:198
    // env:200-203.  The new board's largest exponent: the old one, one more when two such tiles merged (a merge of
    // two 32768s saturates at 15, row_tables.h), or the spawned tile on a board that had none above it
    const uint32_t next_max = kSwarMove ? max(min(prev_max + (max_merged ? 1u : 0u), 15u), spawned) : max_exponent(next);
    s.highest = max(s.highest, next_max);
    s.board = next;
    return r;
}

template <bool kSwarMove, bool kReward>
__device__ __forceinline__ StepResult2 env_step_pairs(EnvState &s, uint32_t action, const uint16_t *row, const uint8_t *code,
                                                      const uint32_t *pairs, const PhiloxKey &K, uint32_t game,
                                                      const uint32_t *inject, unsigned long long *overflow)
{
    return env_step_pairs<kSwarMove, kReward>(s, action, row, code, pairs, K, game, inject, overflow, [] {});
}

// ---- fused-rollout fast path ----------------------------------------------------------------
// Same transition as env_step(), for callers that keep an env in registers across steps and
// therefore can carry what a step already knows into the next one: the empty count, the sum
// of the tile values (a move conserves it, a spawn adds 2 or 4) and the board's max exponent
// (it grows by one exactly when two max tiles merge).  Bit-exact with env_step(); the GPU
// tests hold both to the same reference results.
struct TrackedEnv {
    EnvState s;
    int n_empty;          // empty cells of s.board
    uint32_t total;       // sum of the tile values of s.board
    uint32_t bmax;        // largest exponent on s.board
};

__device__ __forceinline__ void track(TrackedEnv &t)
{
    t.n_empty = count_empty(t.s.board);
    t.total = (uint32_t)pow2_sum(t.s.board.hi, LSB4, pow2_sum(t.s.board.lo, LSB4, 0.0f));
    t.bmax = max_exponent(t.s.board);
}

// env_reset() + track() for the fused rollout.  Same two spawns; a fresh board is known without
// looking at it (14 empty cells, two tiles of exponent 1 or 2) and the k-th empty cell of an empty
// board is cell k.  The reset draws spawns i and i+1; it also hands back the words of spawn i+2
// (`next0`) and, when i is even, of spawn i+3 (`next1`), which sit in the second block it computes:
// the caller's spawn-word schedule (rollout_steps) continues from them.
__device__ __forceinline__ void reset_tracked(TrackedEnv &t, const PhiloxKey &K, uint32_t game, SpawnWords &next0,
                                              SpawnWords &next1)
{
    EnvState &s = t.s;
    const uint32_t i = s.spawn_ctr;
    const Philox4 p = philox4x32_10(i >> 1, 0u, game, DOM_ENV, K);
    const Philox4 q = philox4x32_10((i >> 1) + 1u, 0u, game, DOM_ENV, K);
    const bool odd = (i & 1u) != 0u;
    const uint32_t pos0 = odd ? p.w[2] : p.w[0], val0 = odd ? p.w[3] : p.w[1];
    const uint32_t pos1 = odd ? q.w[0] : p.w[2], val1 = odd ? q.w[1] : p.w[3];
    next0.pos = odd ? q.w[2] : q.w[0]; next0.val = odd ? q.w[3] : q.w[1];
    next1.pos = q.w[2]; next1.val = q.w[3];
    const uint32_t e0 = val0 < 3865470567u ? 1u : 2u, e1 = val1 < 3865470567u ? 1u : 2u;
    const uint32_t c0 = pos0 >> 28;                                  // (pos * 16) >> 32
    const uint32_t k1 = __umulhi(pos1, 15u);
    const uint32_t c1 = k1 + (k1 >= c0 ? 1u : 0u);                   // k1-th of the 15 cells left
    const uint32_t t0 = e0 << ((4u * c0) & 31u), t1 = e1 << ((4u * c1) & 31u);
    s.board = Board((c0 < 8u ? t0 : 0u) | (c1 < 8u ? t1 : 0u), (c0 < 8u ? 0u : t0) | (c1 < 8u ? 0u : t1));
    s.score = 0;
    s.spawn_ctr = i + 2u;
    s.highest = max(e0, e1);
    t.n_empty = 14;
    t.total = 2u * (e0 + e1);                                        // 2^1 = 2*1, 2^2 = 2*2
    t.bmax = s.highest;
}

// The step is split in two so that a caller can software-pipeline it: `step_move` produces
// the next board (move, spawn, game-over test) and everything the reward needs; `step_reward`
// turns that into the float64 reward.  The reward of step t does not feed step t+1, so a loop
// that issues step_move(t+1) and step_reward(t) back to back gives the scheduler two
// independent dependency chains (the kernel has only ~3.5 warps per scheduler to hide latency).
//
// kTrackMax = false: the caller guarantees highest == board max on entry, which the transition
// preserves (highest only ever follows the board's max), so the dead "new highest tile" branch
//  cannot fire and neither value is maintained per step; the caller sets
// highest = max_exponent(board) when it is done.
struct PendingReward {
    Board cur;                 // board after move + spawn
    uint32_t nzl, nzh;         // its occupancy flags
    uint32_t score_delta, total, highest_before, prev_max;
    int empty_before, empty_after;
    bool valid;
};

// `w`: the words of spawn number s.spawn_ctr of the env stream (used iff the move is valid).
template <bool kTrackMax>
__device__ __forceinline__ PendingReward step_move(TrackedEnv &t, uint32_t action, const uint16_t *row,
                                                   const uint8_t *code, const uint32_t *pairs, const SpawnWords &w,
                                                   uint32_t &saturated, bool &full)
{
    PendingReward p;
    EnvState &s = t.s;
    const Board prev = s.board;
    const Board line = to_line(prev, action);
    Board next = from_line(move_left<true>(line, row), action);
    const uint32_t gained = merge_score_pairs<true>(line, code, pairs);      // action is always 0..3 here
    p.score_delta = gained & (kPairSaturated - 1u);
    saturated |= gained;                                                    // bits 28.. : see rollout_saturated()
    s.score += (int32_t)p.score_delta;
    p.valid = next != prev;
    uint32_t nzl = nz_flags(next.lo), nzh = nz_flags(next.hi);             // occupancy before the spawn
    const uint32_t zl = nzl ^ LSB4, zh = nzh ^ LSB4;
    const int cl = __popc(zl);
    int empty_after = cl + __popc(zh);
    // spawn (env:59-67) computed unconditionally and masked by `valid` (env:191-192; a valid move
    // always leaves an empty cell)
    const SpawnPick sp = pick_spawn(zl, zh, cl, empty_after, w.pos, w.val);
    const uint32_t exponent = sp.exponent, flag_lo = sp.flag_lo, flag_hi = sp.flag_hi;
    uint32_t spawn_value = 0u, spawn_exp = 0u;
    if (p.valid) {
        next.lo |= flag_lo * exponent;
        next.hi |= flag_hi * exponent;
        nzl |= flag_lo;
        nzh |= flag_hi;
        s.spawn_ctr += 1u;
        empty_after -= 1;
        spawn_exp = exponent;
        spawn_value = 2u * exponent;                                       // 2^1 = 2*1, 2^2 = 2*2
    }
    p.total = t.total + spawn_value;
    p.cur = next; p.nzl = nzl; p.nzh = nzh;
    p.empty_before = t.n_empty; p.empty_after = empty_after;
    p.highest_before = kTrackMax ? s.highest : 0u;
    p.prev_max = kTrackMax ? t.bmax : 0u;
    if (kTrackMax) {
        // two tiles of the current maximum merged <=> some merge code equals bmax
        const uint32_t codes = merge_codes<true>(line, code);
        uint32_t bmax = t.bmax + (zero_flags(codes ^ (t.bmax * LSB4)) != 0u ? 1u : 0u);
        bmax = max(bmax, spawn_exp);
        s.highest = max(s.highest, bmax);
        t.bmax = bmax;
    }
    full = empty_after == 0;              // only a full board can be game over (env:279-288)
    s.board = next;
    t.n_empty = empty_after; t.total = p.total;
    return p;
}

__device__ __forceinline__ double step_reward(const PendingReward &p, const uint32_t *pairs)
{
    return shaped_reward_tracked(p.valid, p.empty_before, p.cur, p.empty_after, p.nzl, p.nzh, p.score_delta,
                                 p.highest_before, p.prev_max, p.total, pairs);
}
// `saturated` of step_move() accumulates raw table sums; some merge produced 2^16 iff a flag bit is set
__device__ __forceinline__ bool rollout_saturated(uint32_t saturated) { return (saturated & ~(kPairSaturated - 1u)) != 0u; }

// A full board is over when no two neighbours are equal.  (x ^ shifted) | guard has a zero
// nibble exactly where a real pair is equal; "any zero nibble" via the borrow trick is exact.
__device__ __forceinline__ bool full_board_game_over(Board b)
{
    auto zero_nibbles = [](uint32_t v) { return (v - LSB4) & ~v; };        // bit 3 of the lowest zero nibble (and maybe above it)
    uint32_t z = zero_nibbles((b.lo ^ (b.lo >> 4)) | 0xF000F000u) | zero_nibbles((b.hi ^ (b.hi >> 4)) | 0xF000F000u) |
                 zero_nibbles(b.lo ^ __funnelshift_r(b.lo, b.hi, 16)) | zero_nibbles((b.hi ^ (b.hi >> 16)) | 0xFFFF0000u);
    return (z & MSB4) == 0u;                                                // no equal neighbours anywhere
}

// The fused rollout's step loop for one env (env_rollout_kernel; the host emulation runs the same
// code).  A finished game is reset right after the step that ended it.  Returns whether some merge
// saturated a nibble.
//  * Software-pipelined: step_move(t) and step_reward(t-1) share a basic block.
//  * Actions: one Philox block of the action stream holds 64 two-bit actions (16 per word), so the
//    loop nest is block -> word -> step and a step pays one AND and one shift for its action.
//  * Spawn words: a Philox block of the env stream serves two spawns, and a step uses at most one,
//    so blocks are fetched on a fixed schedule that is the same for every lane of a warp: one block
//    per PAIR of steps (t even, t+1).  With c = spawn counter at the pair's start the block is
//    (c+1) >> 1: for even c it holds spawns c, c+1; for odd c it holds c+1, c+2 and spawn c is the
//    word pair `cache` left over from the block before.  Whatever the two steps consume (0, 1 or 2
//    spawns), the counter they leave is covered again: if it is odd, `cache` is set to its spawn.
template <bool kTrackMax>
__device__ __forceinline__ bool rollout_steps(TrackedEnv &e, int32_t steps, uint32_t t0, const PhiloxKey &K,
                                              uint32_t game, const uint16_t *row, const uint8_t *code,
                                              const uint32_t *pairs, double &rsum, int32_t &episodes)
{
    if (steps <= 0) return false;
    const uint32_t end = t0 + (uint32_t)steps;
    uint32_t saturated = 0u;
    bool full;
    // the cache of an odd counter at entry: second word pair of block c >> 1
    const Philox4 entry = philox4x32_10(e.s.spawn_ctr >> 1, 0u, game, DOM_ENV, K);
    SpawnWords cache = {entry.w[2], entry.w[3]}, q0, q1, q2;
    // q0, q1, q2 := the spawns c, c+1, c+2 as far as block (c+1) >> 1 and the cache give them
    auto fetch = [&]() {
        const uint32_t c = e.s.spawn_ctr;
        const Philox4 p = philox4x32_10((c + 1u) >> 1, 0u, game, DOM_ENV, K);
        const bool odd = (c & 1u) != 0u;
        q0.pos = odd ? cache.pos : p.w[0]; q0.val = odd ? cache.val : p.w[1];
        q1.pos = odd ? p.w[0] : p.w[2];    q1.val = odd ? p.w[1] : p.w[3];
        q2.pos = p.w[2];                   q2.val = p.w[3];
    };
    // game over -> reset; the stream continues with the words the reset hands back.  They are filed
    // as "the spawn after one more" (q1) and "after two more" (q2) and the step is counted as having
    // consumed one (`a`), which is where a step that follows in the same pair and the cache rule look
    // for them.  (A reset normally follows a valid move; a caller may also hand in a dead board.)
    bool a;
    auto reset_if_over = [&]() {
        if (full && full_board_game_over(e.s.board)) { ++episodes; reset_tracked(e, K, game, q1, q2); a = true; }
    };
    // first step of the launch: no pending reward to overlap with yet
    fetch();
    PendingReward pend = step_move<kTrackMax>(e, random_action(K, game, t0), row, code, pairs, q0, saturated, full);
    a = pend.valid;
    reset_if_over();
    cache = a ? q1 : q0;
    uint32_t t = t0 + 1u;
    while (t < end) {
        const Philox4 act = philox4x32_10(t >> 6, 0u, game, DOM_ACTION, K);
        const uint32_t block_end = min(end, (t | 63u) + 1u);
        while (t < block_end) {
            const uint32_t sel = (t >> 4) & 3u;
            uint32_t word = sel == 0 ? act.w[0] : sel == 1 ? act.w[1] : sel == 2 ? act.w[2] : act.w[3];
            word >>= 2u * (t & 15u);
            const uint32_t word_end = min(block_end, (t | 15u) + 1u);
            while (t < word_end) {
                fetch();
                PendingReward cur = step_move<kTrackMax>(e, word & 3u, row, code, pairs, q0, saturated, full);
                rsum = __dadd_rn(rsum, step_reward(pend, pairs));      // float64 sum stays in step order
                pend = cur;
                a = cur.valid;
                reset_if_over();
                if ((t & 1u) || t + 1u == word_end) {                  // lone step: before an even t, or the last one
                    cache = a ? q1 : q0;
                    word >>= 2; ++t;
                    continue;
                }
                const SpawnWords wb = a ? q1 : q0;                      // second step of the pair
                cur = step_move<kTrackMax>(e, (word >> 2) & 3u, row, code, pairs, wb, saturated, full);
                rsum = __dadd_rn(rsum, step_reward(pend, pairs));
                pend = cur;
                const bool b = cur.valid;
                cache = a ? (b ? q2 : q1) : (b ? q1 : q0);             // spawn number (c + consumed)
                if (full && full_board_game_over(e.s.board)) {         // reset after the second step: its
                    SpawnWords unused;                                 // next spawn is the new cache
                    ++episodes;
                    reset_tracked(e, K, game, cache, unused);
                }
                word >>= 4; t += 2;
            }
        }
    }
    rsum = __dadd_rn(rsum, step_reward(pend, pairs));
    if (!kTrackMax) e.s.highest = max_exponent(e.s.board);
    return rollout_saturated(saturated);
}

}  // namespace g2048
