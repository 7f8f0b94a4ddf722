// Host emulation of the packed-board device primitives (csrc/board.cuh, csrc/env.cuh).
// TEST INFRASTRUCTURE ONLY: lets the CPU test-suite check the SWAR arithmetic of the CUDA
// kernels against the oracle without a GPU.  It is never loaded by the product.
#include <stdint.h>
#include <algorithm>
#include <cmath>

#define G2048_HOST_EMUL 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__

using std::max;
using std::min;

static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
    return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (sh & 31u)) >> 32);
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    uint64_t v = (((uint64_t)b) << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b)
{
    uint32_t lo = std::max(a & 0xFFFFu, b & 0xFFFFu), hi = std::max(a >> 16, b >> 16);
    return lo | (hi << 16);
}
static inline uint32_t __vsadu4(uint32_t a, uint32_t b)
{
    uint32_t s = 0;
    for (int i = 0; i < 4; ++i) {
        int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF;
        s += (uint32_t)(x > y ? x - y : y - x);
    }
    return s;
}
static inline float __uint_as_float(uint32_t u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }

#include "../../2048-using-reinforcement-learning_b200/csrc/env.cuh"
#include "../../2048-using-reinforcement-learning_b200/csrc/row_tables.h"

using namespace g2048;

// the product's own table builder (csrc/row_tables.h), so the tables are under test too
static uint16_t g_row[65536];
static uint8_t g_code[65536];
static uint32_t g_pairs[kPairEntries];
static uint32_t g_corners[256];
static unsigned long long g_overflow;

extern "C" {

void emul_init(void)
{
    build_row_tables(g_row, g_code);
    for (int i = 0; i < kPairEntries; ++i) g_pairs[i] = pair_table_entry(i);
    for (int i = 0; i < 256; ++i) g_corners[i] = corner_table_entry(i);
}
uint32_t emul_row(uint32_t r) { return g_row[r & 0xFFFF]; }
uint32_t emul_code(uint32_t r) { return g_code[r & 0xFFFF]; }
uint64_t emul_transpose(uint64_t b) { return transpose(Board(b)).u64(); }
uint64_t emul_flip_rows(uint64_t b) { return flip_rows(Board(b)).u64(); }
uint64_t emul_flip_row_order(uint64_t b) { return flip_row_order(Board(b)).u64(); }
uint64_t emul_rot180(uint64_t b) { return rot180(Board(b)).u64(); }
uint64_t emul_env_move(uint64_t b, uint32_t action) { return env_move<false>(Board(b), action, g_row).u64(); }
uint32_t emul_move_score(uint64_t b, uint32_t action)
{
    uint32_t sat;
    return action < 4 ? decode_score(merge_codes<false>(to_line(Board(b), action), g_code), &sat) : 0;
}
uint32_t emul_env_legal(uint64_t b) { return env_legal_mask(Board(b)); }
uint32_t emul_agent_legal(uint64_t b)
{
    Board x(b);
    uint32_t m = env_legal_mask(x);
    return (m & 7u) | ((rot180(env_move<false>(x, 3u, g_row)) != x) ? 8u : 0u);
}
uint64_t emul_agent_child(uint64_t b, uint32_t action)
{
    Board x(b);
    if (action == 3) {
        Board t = transpose(x);
        return transpose(flip_row_order(move_left<false>(flip_rows(t), g_row))).u64();
    }
    return env_move<false>(x, action, g_row).u64();
}
// table-free LEFT move of a half board (two rows): result rows in the low word, code word in the high one
uint64_t emul_move_left_half(uint32_t x)
{
    const HalfMove m = move_left_half(x);
    return ((uint64_t)m.codes << 32) | m.rows;
}
uint32_t emul_code_score(uint32_t codes_lo, uint32_t codes_hi) { return code_score_pairs(codes_lo, codes_hi, g_pairs); }
uint32_t emul_tile_total(uint64_t b) { return tile_total_pairs(Board(b), g_pairs); }
int emul_count_empty(uint64_t b) { return count_empty(Board(b)); }
uint32_t emul_max_exponent(uint64_t b) { return max_exponent(Board(b)); }
uint64_t emul_place_tile(uint64_t b, uint32_t pw, uint32_t vw) { Board x(b); place_tile(x, pw, vw); return x.u64(); }
int emul_fast_eval(uint64_t b) { Board x(b); return fast_eval(x, count_empty(x), max_exponent(x)); }
double emul_full_eval(uint64_t b, int phase) { Board x(b); return full_eval(x, count_empty(x), max_exponent(x), phase); }
// the fused rollout's table-driven pieces against their arithmetic twins
uint32_t emul_edge_sum(uint64_t b, int use_pairs)
{
    Board x(b);
    const uint32_t total = tile_sum_half(x.lo, LSB4) + tile_sum_half(x.hi, LSB4);
    if (use_pairs) return edge_sum_pairs(x, total, g_pairs);
    return total - (tile_sum_half(x.lo, 0x01100000u) + tile_sum_half(x.hi, 0x00000110u))
                 + (tile_sum_half(x.lo, 0x00001001u) + tile_sum_half(x.hi, 0x10010000u));
}
uint32_t emul_move_score_pairs(uint64_t b, uint32_t action)     // bits 28.. = rows with a saturated merge
{
    return merge_score_pairs<false>(to_line(Board(b), action), g_code, g_pairs);
}
void emul_ordered_lines(uint64_t b, int use_flags, int *line)
{
    Board x(b);
    if (use_flags) ordered_pairs_flags(x, nz_flags(x.lo), nz_flags(x.hi), line);
    else ordered_pairs(x, line);
}
// the beam kernels' variants: corner term from the 256-entry table, flags handed in
int emul_fast_eval_lut(uint64_t b)
{
    Board x(b);
    return fast_eval_flags(x, nz_flags(x.lo), nz_flags(x.hi), count_empty(x), max_exponent(x), g_corners);
}
double emul_full_eval_lut(uint64_t b, int phase)
{
    Board x(b);
    return full_eval_flags(x, nz_flags(x.lo), nz_flags(x.hi), count_empty(x), max_exponent(x), phase, g_corners);
}
double emul_ppo_heuristic(uint64_t b) { return ppo_heuristic(Board(b)); }
double emul_ppo_top4(uint64_t b) { return ppo_top4_bonus(Board(b)); }
void emul_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out)
{
    PhiloxKey K; { uint32_t a = k0, b = k1; for (int r = 0; r < 10; ++r) { K.k0[r] = a; K.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; } }
    Philox4 p = philox4x32_10(c0, c1, c2, c3, K);
    for (int i = 0; i < 4; ++i) out[i] = p.w[i];
}
uint32_t emul_random_action(uint64_t seed, uint32_t game, uint32_t t)
{
    return random_action(make_philox_key(seed), game, t);
}

struct EmulEnv { uint64_t board; int32_t score; uint32_t highest; uint32_t spawn_ctr; };
struct EmulStep { double reward; uint32_t score_delta; int32_t valid; int32_t done; };

void emul_env_reset(EmulEnv *e, uint64_t seed, uint32_t game)
{
    EnvState s; s.spawn_ctr = e->spawn_ctr;
    env_reset(s, make_philox_key(seed), game);
    e->board = s.board.u64(); e->score = s.score; e->highest = s.highest; e->spawn_ctr = s.spawn_ctr;
}
void emul_env_step(EmulEnv *e, uint32_t action, const uint32_t *inject, uint64_t seed, uint32_t game, EmulStep *o)
{
    EnvState s; s.board = Board(e->board); s.score = e->score; s.highest = e->highest; s.spawn_ctr = e->spawn_ctr;
    StepResult r = env_step<false, false, true>(s, action, g_row, g_code, make_philox_key(seed), game,
                                                inject, &g_overflow);
    e->board = s.board.u64(); e->score = s.score; e->highest = s.highest; e->spawn_ctr = s.spawn_ctr;
    o->reward = r.reward; o->score_delta = r.score_delta; o->valid = r.valid; o->done = r.done;
}
// the per-step kernel's transition (env_step_fused_kernel): pair-table scores, table-free or table move
void emul_env_step_pairs(EmulEnv *e, uint32_t action, const uint32_t *inject, uint64_t seed, uint32_t game, int swar,
                         EmulStep *o, uint32_t *legal)
{
    EnvState s; s.board = Board(e->board); s.score = e->score; s.highest = e->highest; s.spawn_ctr = e->spawn_ctr;
    const PhiloxKey K = make_philox_key(seed);
    StepResult2 r = swar ? env_step_pairs<true, true>(s, action, g_row, g_code, g_pairs, K, game, inject, &g_overflow)
                         : env_step_pairs<false, true>(s, action, g_row, g_code, g_pairs, K, game, inject, &g_overflow);
    e->board = s.board.u64(); e->score = s.score; e->highest = s.highest; e->spawn_ctr = s.spawn_ctr;
    o->reward = r.reward; o->score_delta = r.score_delta; o->valid = r.valid; o->done = r.done;
    *legal = r.legal;
}
// `steps` tracked steps with the random-policy action stream and auto-reset, like env_rollout_kernel
}  // extern "C"
template <bool kTrackMax>
static void rollout_tracked(EmulEnv *e, int steps, uint32_t t0, uint64_t seed, uint32_t game, double *reward_sum, int *episodes)
{
    TrackedEnv t; t.s.board = Board(e->board); t.s.score = e->score; t.s.highest = e->highest; t.s.spawn_ctr = e->spawn_ctr;
    track(t);
    const PhiloxKey K = make_philox_key(seed);
    // the kernel's own step loop (env.cuh): action words, software pipeline, reset points
    if (rollout_steps<kTrackMax>(t, steps, t0, K, game, g_row, g_code, g_pairs, *reward_sum, *episodes)) g_overflow += 1;
    e->board = t.s.board.u64(); e->score = t.s.score; e->highest = t.s.highest; e->spawn_ctr = t.s.spawn_ctr;
}
// the same transitions one step at a time with the action recomputed per step and both game-over
// tests compared: an independent walk through what rollout_steps() does in its loop nest
template <bool kTrackMax>
static void rollout_stepwise(EmulEnv *e, int steps, uint32_t t0, uint64_t seed, uint32_t game, double *reward_sum, int *episodes)
{
    TrackedEnv t; t.s.board = Board(e->board); t.s.score = e->score; t.s.highest = e->highest; t.s.spawn_ctr = e->spawn_ctr;
    track(t);
    const PhiloxKey K = make_philox_key(seed);
    uint32_t saturated = 0;
    for (int i = 0; i < steps; ++i) {
        bool full;
        const SpawnWords w = spawn_words(K, game, 0u, DOM_ENV, t.s.spawn_ctr);       // one block per step, no schedule
        PendingReward p = step_move<kTrackMax>(t, random_action(K, game, t0 + i), g_row, g_code, g_pairs, w, saturated, full);
        *reward_sum += step_reward(p, g_pairs);
        bool done = full && full_board_game_over(t.s.board);
        if (full && done != env_game_over(t.s.board)) __builtin_trap();   // the two game-over tests must agree
        if (done) { SpawnWords n0, n1; ++*episodes; reset_tracked(t, K, game, n0, n1); }
    }
    g_overflow += rollout_saturated(saturated) ? 1u : 0u;
    if (!kTrackMax) t.s.highest = max_exponent(t.s.board);
    e->board = t.s.board.u64(); e->score = t.s.score; e->highest = t.s.highest; e->spawn_ctr = t.s.spawn_ctr;
}
extern "C" {
void emul_rollout_tracked(EmulEnv *e, int steps, uint32_t t0, uint64_t seed, uint32_t game, double *reward_sum, int *episodes, int track_max)
{
    if (track_max) rollout_tracked<true>(e, steps, t0, seed, game, reward_sum, episodes);
    else rollout_tracked<false>(e, steps, t0, seed, game, reward_sum, episodes);
}
void emul_rollout_stepwise(EmulEnv *e, int steps, uint32_t t0, uint64_t seed, uint32_t game, double *reward_sum, int *episodes, int track_max)
{
    if (track_max) rollout_stepwise<true>(e, steps, t0, seed, game, reward_sum, episodes);
    else rollout_stepwise<false>(e, steps, t0, seed, game, reward_sum, episodes);
}
unsigned long long emul_overflow(void) { return g_overflow; }

}  // extern "C"
