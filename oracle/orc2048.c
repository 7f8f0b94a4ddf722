/*
 * orc2048.c -- CPU oracle (plain C) for the batched 2048 engine.
 * TEST INFRASTRUCTURE ONLY: see orc2048.h for who may load it.
 *
 * Every function cites the reference lines it restates
 *   env   = /root/reference/environment/game_2048.py
 *   agent = /root/reference/agents/beam_search_agent.py
 * Boards are int32[16] row-major tile values (0 = empty), as in the reference.
 * Build with -ffp-contract=off: the shaped reward and the full evaluation are
 * float64 with a fixed operation order and must not be FMA-contracted.
 */
#include "orc2048.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator   */
/* both sides draw spawns from.  Verified against the Random123 KATs   */
/* in tests/test_philox.py.                                            */
/* ------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout (DESIGN.md "Random streams"): counter = (block, call, game, domain),
 * key = (seed lo, seed hi).  Sequential draw j of a stream is word j&3 of block j>>2;
 * a spawn consumes two draws (position, then value), so spawn i is words
 * 2(i&1), 2(i&1)+1 of block i>>1. */
static void stream_block(uint64_t seed, uint32_t game, uint32_t call, uint32_t domain,
                         uint32_t block, uint32_t out[4])
{
    uint32_t ctr[4] = { block, call, game, domain };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    orc_philox4x32_10(ctr, key, out);
}

void orc_spawn_words(uint64_t seed, uint32_t game, uint32_t call, uint32_t domain,
                     uint32_t i, uint32_t *pos_word, uint32_t *val_word)
{
    uint32_t w[4];
    stream_block(seed, game, call, domain, i >> 1, w);
    *pos_word = w[2 * (i & 1)];
    *val_word = w[2 * (i & 1) + 1];
}

int orc_random_action(uint64_t seed, uint32_t game, uint32_t t)
{
    uint32_t w[4];
    stream_block(seed, game, 0, ORC_DOM_ACTION, t >> 6, w);
    return (int)((w[(t >> 4) & 3] >> (2 * (t & 15))) & 3u);
}

void orc_synthetic_board(uint64_t seed, uint32_t game, int32_t board[16])
{
    for (uint32_t blk = 0; blk < 4; ++blk) {
        uint32_t w[4];
        stream_block(seed, game, 0, ORC_DOM_BOARD, blk, w);
        for (int j = 0; j < 4; ++j) {
            uint32_t x = w[j];
            int e = 1 + (int)(((x >> 16) * 11u) >> 16);
            board[4 * blk + j] = ((x & 0xFFFFu) < 19661u) ? 0 : (1 << e);
        }
    }
}

/* index = floor(u * n / 2^32): the integer-exact form of "uniform choice among n". */
static inline int pick_index(uint32_t word, int n) { return (int)(((uint64_t)word * (uint64_t)n) >> 32); }
/* random.random() < 0.9  with random() = word / 2^32  <=>  word < ceil(0.9 * 2^32) */
static inline int32_t pick_tile(uint32_t word) { return word < 3865470567u ? 2 : 4; }

/* ------------------------------------------------------------------ */
/* Row kernel shared by env and agent.                                  */
/* env:116-168 and agent:213-242 have identical row semantics: drop the */
/* zeros, scan left to right merging an equal neighbour once, pad.     */
/* `line` lists the 4 cell indices in the order tiles travel towards.  */
/* ------------------------------------------------------------------ */
static int64_t slide_line(int32_t *cells, const int line[4])
{
    int32_t packed[4]; int m = 0;
    for (int j = 0; j < 4; ++j) if (cells[line[j]] != 0) packed[m++] = cells[line[j]];
    int32_t res[4] = {0, 0, 0, 0}; int k = 0; int64_t gained = 0;
    for (int j = 0; j < m; ) {
        if (j + 1 < m && packed[j] == packed[j + 1]) {
            int32_t v = packed[j] * 2;
            res[k++] = v; gained += v; j += 2;
        } else {
            res[k++] = packed[j]; j += 1;
        }
    }
    for (int j = 0; j < 4; ++j) cells[line[j]] = res[j];
    return gained;
}

/* env:97-114: 0 LEFT, 1 UP (transpose), 2 RIGHT (fliplr), 3 DOWN (T, fliplr, ..., fliplr, T).
 * Any other action is a no-op there (no else branch). */
int64_t orc_env_move(int32_t board[16], int action)
{
    int64_t gained = 0;
    if (action < 0 || action > 3) return 0;
    for (int i = 0; i < 4; ++i) {
        int line[4];
        for (int j = 0; j < 4; ++j) {
            switch (action) {
            case 0: line[j] = 4 * i + j;        break;   /* row i, towards column 0 */
            case 1: line[j] = 4 * j + i;        break;   /* column i, towards row 0 */
            case 2: line[j] = 4 * i + (3 - j);  break;   /* row i, towards column 3 */
            default: line[j] = 4 * (3 - j) + i; break;   /* column i, towards row 3 */
            }
        }
        gained += slide_line(board, line);
    }
    return gained;
}

int orc_env_legal_mask(const int32_t board[16])          /* env:69-95 */
{
    int mask = 0;
    for (int a = 0; a < 4; ++a) {
        int32_t tmp[16];
        memcpy(tmp, board, sizeof tmp);
        orc_env_move(tmp, a);
        if (memcmp(tmp, board, sizeof tmp) != 0) mask |= 1 << a;
    }
    return mask;
}

static int count_empty(const int32_t b[16]) { int n = 0; for (int i = 0; i < 16; ++i) n += (b[i] == 0); return n; }
static int32_t max_tile(const int32_t b[16]) { int32_t m = 0; for (int i = 0; i < 16; ++i) if (b[i] > m) m = b[i]; return m; }

/* env:59-67 / agent:260-269: k-th empty cell in row-major order gets 2 (90 %) or 4.
 * No draw and no change when the board is full.  Returns 1 if a tile was placed. */
static int place_tile(int32_t b[16], uint32_t pos_word, uint32_t val_word)
{
    int n = count_empty(b);
    if (n == 0) return 0;
    int k = pick_index(pos_word, n);
    for (int i = 0; i < 16; ++i) {
        if (b[i] == 0) { if (k == 0) { b[i] = pick_tile(val_word); return 1; } --k; }
    }
    return 0;
}

static void env_spawn(orc_env *env)
{
    if (count_empty(env->board) == 0) return;          /* env:61: no draw on a full board */
    uint32_t pw, vw;
    orc_spawn_words(env->seed, env->game, 0, ORC_DOM_ENV, env->spawn_ctr, &pw, &vw);
    env->spawn_ctr += 1;
    place_tile(env->board, pw, vw);
}

void orc_env_reset(orc_env *env)                         /* env:29-48 */
{
    memset(env->board, 0, sizeof env->board);
    env->score = 0;
    env->game_over = 0;
    env->highest_tile = 0;
    env_spawn(env);
    env_spawn(env);
    env->highest_tile = max_tile(env->board);
}

/* env:212-277.  float64, operation order exactly as written there.
 * highest_tile_before is env.highest_tile at the time _calculate_reward runs,
 * i.e. BEFORE the update at env:200-203. */
double orc_env_reward(int valid, const int32_t prev[16], const int32_t cur[16],
                      int64_t score_delta, int32_t highest_tile_before)
{
    double reward = (double)score_delta / 4.0;                       /* :225-226 */
    if (highest_tile_before > max_tile(prev)) {                      /* :229 */
        reward += 2.0 * log2((double)highest_tile_before);           /* :231 */
        if (highest_tile_before >= 256)  reward += 50;
        if (highest_tile_before >= 512)  reward += 100;
        if (highest_tile_before >= 1024) reward += 200;
        if (highest_tile_before >= 2048) reward += 500;
    }
    if (!valid) reward -= 2.0;                                       /* :244-245 */
    int empty_before = count_empty(prev), empty_after = count_empty(cur);
    reward += (double)(empty_after - empty_before) * 0.5;            /* :249-251 */
    int64_t edge = 0, total = 0;
    for (int j = 0; j < 4; ++j) edge += cur[j];                      /* row 0   */
    for (int j = 0; j < 4; ++j) edge += cur[12 + j];                 /* row 3   */
    for (int j = 0; j < 4; ++j) edge += cur[4 * j];                  /* col 0   */
    for (int j = 0; j < 4; ++j) edge += cur[4 * j + 3];              /* col 3   */
    for (int i = 0; i < 16; ++i) total += cur[i];
    reward += ((double)edge / (double)total) * 1.0;                  /* :254-259 */
    if (empty_after <= 2) reward -= 2.0;                             /* :262-263 */
    for (int i = 0; i < 4; ++i) {                                    /* :267-275 */
        int row_ordered = 0, col_ordered = 0;
        for (int j = 1; j < 4; ++j) {
            int32_t a = cur[4 * i + j - 1], b = cur[4 * i + j];
            if (a > 0 && b > 0 && b >= a) ++row_ordered;
            int32_t c = cur[4 * (j - 1) + i], d = cur[4 * j + i];
            if (c > 0 && d > 0 && d >= c) ++col_ordered;
        }
        reward += (double)(row_ordered + col_ordered) * 0.1;
    }
    return reward;
}

void orc_env_step(orc_env *env, int action, const uint32_t *inject, orc_step_out *out)   /* env:170-210 */
{
    int32_t prev[16];
    memcpy(prev, env->board, sizeof prev);
    int64_t gained = orc_env_move(env->board, action);               /* :185 */
    env->score += gained;
    int valid = memcmp(prev, env->board, sizeof prev) != 0;          /* :188 */
    if (valid) {                                                     /* :191-192 */
        if (inject) place_tile(env->board, inject[0], inject[1]);
        else env_spawn(env);
    }
    double reward = orc_env_reward(valid, prev, env->board, gained, env->highest_tile);  /* :195 */
    env->game_over = (orc_env_legal_mask(env->board) == 0);          /* :198 */
    int32_t hi = max_tile(env->board);                               /* :200-203 */
    if (hi > env->highest_tile) env->highest_tile = hi;
    out->reward = reward; out->valid = valid; out->done = env->game_over; out->score_delta = gained;
}

/* env:341-387 simulate_move(state, action): every empty cell x {2, 4} after the move.
 * Restated WITH the reference's quirk: the loop writes `self.board = new_state.copy()` (env:378)
 * for the game-over test and never restores it inside the loop, so `new_state = self.board.copy()`
 * (env:371) starts from the PREVIOUS outcome (tiles accumulate: earlier cells stay filled with
 * their last value, 4) and the reward (env:375) is computed on that previous board.
 * Returns the number of outcomes (<= 30). */
int orc_env_simulate_move(const int32_t state[16], int action, int32_t highest_tile,
                          int32_t out_boards[30][16], double out_reward[30], int32_t out_done[30])
{
    int32_t cur[16];
    memcpy(cur, state, sizeof cur);
    int64_t gained = orc_env_move(cur, action);                       /* env:362 (score stays bumped) */
    if (memcmp(cur, state, sizeof cur) == 0) return 0;                /* env:363-366 */
    int empties[16], n = 0;
    for (int i = 0; i < 16; ++i) if (cur[i] == 0) empties[n++] = i;   /* env:367, taken once */
    int k = 0;
    for (int e = 0; e < n; ++e) {
        for (int tile = 2; tile <= 4; tile += 2) {
            int32_t ns[16];
            memcpy(ns, cur, sizeof ns);                               /* env:371: self.board, not the moved board */
            ns[empties[e]] = tile;
            out_reward[k] = orc_env_reward(1, state, cur, gained, highest_tile);   /* env:375 */
            memcpy(cur, ns, sizeof cur);                              /* env:378 */
            out_done[k] = orc_env_legal_mask(cur) == 0;               /* env:379 */
            memcpy(out_boards[k], ns, sizeof ns);
            ++k;
        }
    }
    return k;
}

/* agents/hybrid.py:578-635: the hybrid agent's own Game2048Env.simulate_move (monkey-patched onto a
 * private copy of the env class, hybrid.py:695-697) with _calculate_simulation_reward (:676-692).
 * An invalid move yields the single outcome (board, -1.0, False).  Otherwise min(3, #empty) cells
 * are drawn with random.sample -- served by the shim as a partial Fisher-Yates over the row-major
 * empty list, one draw per pick, j = i + floor(u * (n - i) / 2^32) -- and each yields a 2-tile
 * outcome (reward * 0.9) and a 4-tile outcome (reward * 0.1).  *draw is the sequential draw index
 * in stream (seed, game, call, ORC_DOM_HYBRID), advanced by the picks made. */
int orc_hybrid_simulate_move(const int32_t board[16], int action, uint64_t seed, uint32_t game, uint32_t call,
                             uint32_t *draw, int32_t out_boards[6][16], double out_reward[6], int32_t out_done[6])
{
    int32_t moved[16];
    memcpy(moved, board, sizeof moved);
    orc_env_move(moved, action);                                       /* hybrid.py:588-602 (true directions) */
    if (memcmp(moved, board, sizeof moved) == 0) {                     /* :605-609 */
        memcpy(out_boards[0], moved, sizeof moved); out_reward[0] = -1.0; out_done[0] = 0;
        return 1;
    }
    int cells[16], n = 0;
    for (int i = 0; i < 16; ++i) if (moved[i] == 0) cells[n++] = i;    /* :612 */
    if (n == 0) { memcpy(out_boards[0], moved, sizeof moved); out_reward[0] = 0.0; out_done[0] = 1; return 1; }   /* :613-615 */
    int k = n < 3 ? n : 3, count = 0;                                  /* :621-622 */
    int64_t old_sum = 0; int32_t old_max = max_tile(board);
    for (int i = 0; i < 16; ++i) old_sum += board[i];
    for (int i = 0; i < k; ++i) {
        uint32_t w[4];
        stream_block(seed, game, call, ORC_DOM_HYBRID, *draw >> 2, w);
        int j = i + pick_index(w[*draw & 3], n - i);
        *draw += 1;
        int t = cells[i]; cells[i] = cells[j]; cells[j] = t;
        for (int tile = 2; tile <= 4; tile += 2) {                     /* :624-634 */
            int32_t nb[16];
            memcpy(nb, moved, sizeof nb);
            nb[cells[i]] = tile;
            int64_t new_sum = 0; int32_t new_max = max_tile(nb);
            for (int c = 0; c < 16; ++c) new_sum += nb[c];
            int64_t merge_reward = new_sum - old_sum;                  /* :679-681 */
            int64_t bonus = new_max > old_max ? new_max : 0;           /* :684-688 */
            double empty_bonus = (double)count_empty(nb) * 0.1;        /* :691 */
            double reward = (double)(merge_reward + bonus) + empty_bonus;
            memcpy(out_boards[count], nb, sizeof nb);
            out_reward[count] = reward * (tile == 2 ? 0.9 : 0.1);
            out_done[count] = 0;
            ++count;
        }
    }
    return count;
}

/* env:313-339 _evaluate_pattern (never called by the reference): max of the snake-weighted and the
 * corner-weighted sum of tile VALUES, each / 100. */
double orc_env_pattern(const int32_t b[16])
{
    static const double SNAKE[16] = {16, 15, 14, 13, 9, 10, 11, 12, 8, 7, 6, 5, 1, 2, 3, 4};
    static const double CORNER[16] = {16, 8, 4, 2, 8, 4, 2, 1, 4, 2, 1, 0.5, 2, 1, 0.5, 0.25};
    double s = 0.0, c = 0.0;                 /* all partial sums are exact dyadic numbers */
    for (int i = 0; i < 16; ++i) { s += (double)b[i] * SNAKE[i]; c += (double)b[i] * CORNER[i]; }
    s /= 100.0; c /= 100.0;
    return s > c ? s : c;
}

/* ------------------------------------------------------------------ */
/* Beam-search agent                                                    */
/* ------------------------------------------------------------------ */

/* agent:194-258.  LEFT/UP/RIGHT agree with the env.  For DOWN the post-rotation
 * is applied in the wrong order (agent:251-253 vs the pre-rotation at :210), so
 * the returned board is the true DOWN result rotated by 180 degrees, and
 * `valid` compares THAT with the input. */
int orc_agent_move(const int32_t board[16], int action, int32_t out[16], int64_t *merge_score)
{
    int32_t tmp[16];
    memcpy(tmp, board, sizeof tmp);
    int64_t gained = orc_env_move(tmp, action);
    if (action == 3) {
        for (int i = 0; i < 16; ++i) out[i] = tmp[15 - i];
    } else {
        memcpy(out, tmp, sizeof tmp);
    }
    if (merge_score) *merge_score = gained;
    return memcmp(out, board, sizeof tmp) != 0;
}

int orc_agent_legal_mask(const int32_t board[16])        /* agent:183-192 */
{
    int mask = 0; int32_t tmp[16];
    for (int a = 0; a < 4; ++a) if (orc_agent_move(board, a, tmp, NULL)) mask |= 1 << a;
    return mask;
}

static int ilog2(int32_t v) { int e = 0; while (v > 1) { v >>= 1; ++e; } return e; }

double orc_fast_eval(const int32_t b[16])                /* agent:280-314 */
{
    double empty_score = (double)count_empty(b) * 10.0;
    int32_t mx = max_tile(b);
    double max_score = mx > 0 ? (double)ilog2(mx) * 2.0 : 0.0;
    static const int corner[4] = {0, 3, 12, 15};
    int32_t corner_score = 0;
    for (int c = 0; c < 4; ++c) { int32_t s = b[corner[c]] * 2; if (s > 0 && s > corner_score) corner_score = s; }
    int merges = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) if (b[4 * i + j] == b[4 * i + j + 1] && b[4 * i + j] > 0) ++merges;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) if (b[4 * i + j] == b[4 * i + 4 + j] && b[4 * i + j] > 0) ++merges;
    return ((empty_score + max_score) + (double)corner_score) + (double)(merges * 2);
}

int orc_phase(int32_t mx, int32_t early_thr, int32_t mid_thr)   /* agent:271-278 */
{
    if (mx < early_thr) return 0;
    if (mx < mid_thr) return 1;
    return 2;
}

double orc_full_eval(const int32_t b[16], int phase)      /* agent:316-373 (+375-403) */
{
    static const double W_EMPTY[3]  = {15.0, 10.0, 8.0};
    static const double W_MAX[3]    = {1.0, 1.5, 2.0};
    static const double W_CORNER[3] = {2.0, 2.5, 3.0};
    static const double W_MERGE[3]  = {2.0, 1.5, 1.0};
    static const int SNAKE[16] = {15, 14, 13, 12,  8, 9, 10, 11,  7, 6, 5, 4,  0, 1, 2, 3};  /* :37-42 */

    int n0 = count_empty(b);
    double empty_score = (double)n0 * W_EMPTY[phase];                 /* :338-339 */
    if (n0 <= 2) empty_score -= 10.0;                                 /* :342-343 */
    int32_t mx = max_tile(b);
    double max_score = mx > 0 ? (double)ilog2(mx) * W_MAX[phase] : 0.0;   /* :346-347 */
    if (mx >= 512)  max_score *= 1.2;                                 /* :350-355 */
    if (mx >= 1024) max_score *= 1.5;
    if (mx >= 2048) max_score *= 2.0;
    int32_t cmax = b[0];                                              /* :375-385 */
    if (b[3] > cmax) cmax = b[3];
    if (b[12] > cmax) cmax = b[12];
    if (b[15] > cmax) cmax = b[15];
    double corner_bonus = (cmax > 0 ? (double)ilog2(cmax) * 2.0 : 0.0) * W_CORNER[phase];   /* :358 */
    double pot = 0.0;                                                 /* :387-403 */
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j)
        if (b[4 * i + j] > 0 && b[4 * i + j] == b[4 * i + j + 1]) pot += (double)ilog2(b[4 * i + j]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j)
        if (b[4 * i + j] > 0 && b[4 * i + j] == b[4 * i + 4 + j]) pot += (double)ilog2(b[4 * i + j]);
    double merge_potential = pot * W_MERGE[phase];                    /* :361 */
    double snake = 0.0;                                               /* :364-370 */
    for (int i = 0; i < 16; ++i) if (b[i] > 0) snake += (double)ilog2(b[i]) * (double)SNAKE[i];
    snake /= 100.0;
    return (((empty_score + max_score) + corner_bonus) + merge_potential) + snake;   /* :373 */
}

/* ------------------------------------------------------------------ */
/* PPO-side features: agents/ppo_agent.py            */
/* ------------------------------------------------------------------ */
/* ppo_agent.py:184-195 normalize_state: log2(tile)/15, 0 for empty (float32) */
void orc_ppo_observe(const int32_t b[16], float obs[16])
{
    for (int i = 0; i < 16; ++i) obs[i] = b[i] > 0 ? (float)ilog2(b[i]) / 15.0f : 0.0f;
}

/* ppo_agent.py:271-333 evaluate_heuristic: 2 * best-direction monotonicity / 24
 * + 1 if the largest corner is the largest tile - 0.1 * #(tiles >= 8) */
double orc_ppo_heuristic(const int32_t b[16])
{
    int h_le = 0, h_ge = 0, v_le = 0, v_ge = 0;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 3; ++c) {
        int32_t x = b[4 * r + c], y = b[4 * r + c + 1];
        if (x > 0 && y > 0) { h_le += x <= y; h_ge += x >= y; }
    }
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) {
        int32_t x = b[4 * r + c], y = b[4 * r + 4 + c];
        if (x > 0 && y > 0) { v_le += x <= y; v_ge += x >= y; }
    }
    /* max over the four (row_dir, col_dir) combinations of (rows + cols) / 24.0 */
    int best = (h_le > h_ge ? h_le : h_ge) + (v_le > v_ge ? v_le : v_ge);
    double score = 2.0 * ((double)best / 24.0);
    int32_t cmax = b[0];
    if (b[3] > cmax) cmax = b[3];
    if (b[12] > cmax) cmax = b[12];
    if (b[15] > cmax) cmax = b[15];
    if (cmax == max_tile(b)) score += 1.0;
    int high = 0;
    for (int i = 0; i < 16; ++i) high += b[i] >= 8;
    if (high > 0) score += -0.1 * (double)high;
    return score;
}

/* ppo_agent.py:251-254: 0.1 * sum(log2 of the four largest tiles that are > 0) */
double orc_ppo_top4_bonus(const int32_t b[16])
{
    int32_t t[16];
    memcpy(t, b, sizeof t);
    for (int i = 0; i < 4; ++i)                 /* partial selection sort, descending */
        for (int j = i + 1; j < 16; ++j) if (t[j] > t[i]) { int32_t x = t[i]; t[i] = t[j]; t[j] = x; }
    double sum = 0.0;
    for (int i = 3; i >= 0; --i) if (t[i] > 0) sum += (double)ilog2(t[i]);     /* ascending, as np.sort()[-4:] */
    return 0.1 * sum;
}

/* ppo_agent.py:234-269 PPOAgent.remember: the reward it stores.  `highest_tile_seen` is the agent's
 * running maximum (in/out, starts at 2, :171); `novel` says whether hash(next_state.tobytes()) was absent
 * from agent.seen_states (:257-260; the caller keeps the set).  float64, the reference's order. */
double orc_ppo_shape_reward(const int32_t state[16], const int32_t next_state[16], double reward,
                            int32_t *highest_tile_seen, int novel)
{
    int32_t cur = 0, nxt = 0;
    for (int i = 0; i < 16; ++i) { if (state[i] > cur) cur = state[i]; if (next_state[i] > nxt) nxt = next_state[i]; }
    if (nxt > *highest_tile_seen) {                                        /* :241-246 */
        double tile_bonus = 5.0 * ((double)ilog2(nxt) - (double)ilog2(*highest_tile_seen));
        *highest_tile_seen = nxt;
        reward += tile_bonus;
    }
    if (nxt < cur)                                                         /* :249-251 (nxt > 0 on this path) */
        reward += -2.0 * ((double)ilog2(cur) - (double)ilog2(nxt));
    reward += orc_ppo_top4_bonus(next_state);                              /* :254-256 */
    if (novel) reward += 0.2;                                              /* :259-262, novelty_factor :175 */
    reward += 0.3 * orc_ppo_heuristic(next_state);                         /* :265-266, heuristic_weight :181 */
    return reward;
}

typedef struct { int32_t board[16]; int first; double score; } cand_t;

/* sorted(..., key=score, reverse=True)[:k]  -- Python's sort is stable, so equal
 * scores keep generation order (agent:131-132, :174-175). */
static int keep_top(cand_t *c, int n, int k)
{
    for (int i = 1; i < n; ++i) {            /* stable insertion sort, descending */
        cand_t x = c[i]; int j = i - 1;
        while (j >= 0 && c[j].score < x.score) { c[j + 1] = c[j]; --j; }
        c[j + 1] = x;
    }
    return n < k ? n : k;
}

void orc_beam_get_action(const int32_t board[16], int legal_mask,
                         int beam_width, int search_depth,
                         int32_t early_thr, int32_t mid_thr,
                         uint64_t seed, uint32_t game, uint32_t call,
                         orc_beam_out *out)
{
    memset(out, 0, sizeof *out);
    int vm = legal_mask < 0 ? orc_agent_legal_mask(board) : (legal_mask & 15);   /* :82-84 */
    if (vm == 0) { out->action = 0; out->prob = 0.5f; return; }                  /* :86-88 */
    if ((vm & (vm - 1)) == 0) {                                                  /* :91-93 */
        int a = 0; while (!((vm >> a) & 1)) ++a;
        out->action = a; out->prob = 1.0f; return;
    }
    int phase = orc_phase(max_tile(board), early_thr, mid_thr);                  /* :96-97 */
    int n0 = count_empty(board), depth;                                          /* :100-106 */
    if (n0 <= 4)       depth = search_depth + 5 < 25 ? search_depth + 5 : 25;
    else if (n0 >= 10) depth = search_depth - 5 < 10 ? search_depth - 5 : 10;
    else               depth = search_depth;
    out->depth_used = depth;

    int cap = 4 * (beam_width > 1 ? beam_width : 1);
    cand_t *cur = (cand_t *)malloc(sizeof(cand_t) * (size_t)(cap + 4));
    cand_t *nxt = (cand_t *)malloc(sizeof(cand_t) * (size_t)(cap + 4));
    int ncur = 0; uint32_t spawn_i = 0; int nodes = 0;

    for (int a = 0; a < 4; ++a) {                                                /* :112-123 */
        if (!((vm >> a) & 1)) continue;
        cand_t c;
        if (!orc_agent_move(board, a, c.board, NULL)) continue;
        if (count_empty(c.board) > 0) {
            uint32_t pw, vw;
            orc_spawn_words(seed, game, call, ORC_DOM_BEAM, spawn_i++, &pw, &vw);
            place_tile(c.board, pw, vw);
        }
        c.first = a; c.score = orc_fast_eval(c.board); ++nodes;
        cur[ncur++] = c;
    }
    if (ncur == 0) {                                                             /* :126-128 */
        int list[4], m = 0;
        for (int a = 0; a < 4; ++a) if ((vm >> a) & 1) list[m++] = a;
        uint32_t w[4];
        uint32_t ctr[4] = {0, call, game, ORC_DOM_BEAM};
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        orc_philox4x32_10(ctr, key, w);
        out->action = list[pick_index(w[0], m)]; out->prob = 0.5f;
        out->nodes = 0; free(cur); free(nxt); return;
    }
    ncur = keep_top(cur, ncur, beam_width);                                      /* :131-132 */

    for (int d = 1; d < depth; ++d) {                                            /* :135-175 */
        int nn = 0;
        for (int r = 0; r < ncur; ++r) {
            int lv = orc_agent_legal_mask(cur[r].board);                         /* :147 */
            for (int a = 0; a < 4; ++a) {
                if (!((lv >> a) & 1)) continue;
                cand_t c;
                if (!orc_agent_move(cur[r].board, a, c.board, NULL)) continue;   /* :152-153 */
                if (count_empty(c.board) > 0) {                                  /* :155, :262-263 */
                    uint32_t pw, vw;
                    orc_spawn_words(seed, game, call, ORC_DOM_BEAM, spawn_i++, &pw, &vw);
                    place_tile(c.board, pw, vw);
                }
                c.first = cur[r].first;
                c.score = d > 3 ? orc_fast_eval(c.board) : orc_full_eval(c.board, phase);   /* :139, :158-161 */
                ++nodes;
                nxt[nn++] = c;
            }
        }
        if (nn == 0) break;                                                      /* :170-171 */
        ncur = keep_top(nxt, nn, beam_width);
        cand_t *t = cur; cur = nxt; nxt = t;
    }
    out->action = cur[0].first; out->prob = 1.0f;                                /* :178-181 */
    out->nodes = nodes; out->best_score = cur[0].score; out->spawns = (int32_t)spawn_i;
    free(cur); free(nxt);
}

void orc_play_game(uint64_t seed, uint32_t game, int beam_width, int search_depth,
                   int32_t early_thr, int32_t mid_thr, int max_moves, orc_game_out *out)
{
    /* evaluate_beam_search.py:16-98: env = Game2048Env(); state = env.reset(); loop
     * agent.get_action(state) (no valid_moves) -> env.step -> stats. */
    orc_env env; memset(&env, 0, sizeof env);
    env.seed = seed; env.game = game; env.spawn_ctr = 0;
    orc_env_reset(&env);          /* ctor reset  (env:27)  */
    orc_env_reset(&env);          /* explicit reset (evaluate_beam_search.py:30) */
    memset(out, 0, sizeof *out);
    for (int m = 0; m < 8; ++m) out->milestone_move[m] = -1;
    int moves = 0, done = 0;
    while (!done && moves < max_moves) {
        orc_beam_out b;
        orc_beam_get_action(env.board, -1, beam_width, search_depth, early_thr, mid_thr,
                            seed, game, (uint32_t)moves, &b);
        out->nodes += b.nodes;
        orc_step_out s;
        orc_env_step(&env, b.action, NULL, &s);
        done = s.done;
        for (int m = 0; m < 8; ++m)           /* milestones 64..8192, evaluate_beam_search.py:42-43,61-64: */
            if (out->milestone_move[m] < 0 && max_tile(env.board) >= (64 << m)) out->milestone_move[m] = moves;   /* index before moves += 1 (:86) */
        if (s.valid) out->valid_moves++; else out->invalid_moves++;
        ++moves;
    }
    out->score = env.score; out->highest_tile = env.highest_tile; out->moves = moves;
}

/* ------------------------------------------------------------------ */
/* Batched helpers (pthreads; one contiguous or strided slice per thread) */
/* ------------------------------------------------------------------ */
int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef void (*item_fn)(int64_t i, void *arg);
typedef struct { item_fn fn; void *arg; int64_t n; int tid, nthreads; } slice_t;

static void *slice_main(void *p)
{
    slice_t *s = (slice_t *)p;
    for (int64_t i = s->tid; i < s->n; i += s->nthreads) s->fn(i, s->arg);   /* strided: balances uneven items */
    return NULL;
}

static void parallel_for(int64_t n, int threads, item_fn fn, void *arg)
{
    if (threads <= 0) threads = orc_max_threads();
    if (threads > n) threads = (int)(n > 0 ? n : 1);
    if (threads <= 1) { for (int64_t i = 0; i < n; ++i) fn(i, arg); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    slice_t *sl = (slice_t *)malloc(sizeof(slice_t) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        sl[t].fn = fn; sl[t].arg = arg; sl[t].n = n; sl[t].tid = t; sl[t].nthreads = threads;
        pthread_create(&th[t], NULL, slice_main, &sl[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(sl);
}

typedef struct {
    int32_t *boards; int64_t *score; int32_t *highest; uint32_t *spawn_ctr;
    double *reward_sum; int32_t *episodes; int steps; uint32_t t0; uint64_t seed; uint32_t game0;
} rollout_args;

static void rollout_item(int64_t i, void *p)
{
    rollout_args *a = (rollout_args *)p;
    orc_env env;
    memcpy(env.board, a->boards + 16 * i, sizeof env.board);
    env.score = a->score[i]; env.highest_tile = a->highest[i]; env.game_over = 0;
    env.spawn_ctr = a->spawn_ctr[i]; env.game = a->game0 + (uint32_t)i; env.seed = a->seed;
    double rs = a->reward_sum[i]; int32_t ep = a->episodes[i];
    for (int s = 0; s < a->steps; ++s) {
        orc_step_out o;
        orc_env_step(&env, orc_random_action(a->seed, env.game, a->t0 + (uint32_t)s), NULL, &o);
        rs += o.reward;                                  /* fp64, summed in step order */
        if (o.done) { ++ep; orc_env_reset(&env); }       /* harness-side auto-reset (SURVEY 8d cfg 2) */
    }
    memcpy(a->boards + 16 * i, env.board, sizeof env.board);
    a->score[i] = env.score; a->highest[i] = env.highest_tile; a->spawn_ctr[i] = env.spawn_ctr;
    a->reward_sum[i] = rs; a->episodes[i] = ep;
}

void orc_rollout(int32_t *boards, int64_t *score, int32_t *highest, uint32_t *spawn_ctr,
                 double *reward_sum, int32_t *episodes,
                 int64_t n, int steps, uint32_t t0, uint64_t seed, uint32_t game0, int threads)
{
    rollout_args a = { boards, score, highest, spawn_ctr, reward_sum, episodes, steps, t0, seed, game0 };
    parallel_for(n, threads, rollout_item, &a);
}

typedef struct {
    const int32_t *boards; int beam_width, search_depth; uint64_t seed; uint32_t game0, call;
    int32_t *action; float *prob; int32_t *nodes; double *best_score;
} beam_args;

static void beam_item(int64_t i, void *p)
{
    beam_args *a = (beam_args *)p;
    orc_beam_out b;
    orc_beam_get_action(a->boards + 16 * i, -1, a->beam_width, a->search_depth, 512, 1024,
                        a->seed, a->game0 + (uint32_t)i, a->call, &b);
    a->action[i] = b.action; a->prob[i] = b.prob; a->nodes[i] = b.nodes; a->best_score[i] = b.best_score;
}

void orc_beam_batch(const int32_t *boards, int64_t n, int beam_width, int search_depth,
                    uint64_t seed, uint32_t game0, uint32_t call,
                    int32_t *action, float *prob, int32_t *nodes, double *best_score, int threads)
{
    beam_args a = { boards, beam_width, search_depth, seed, game0, call, action, prob, nodes, best_score };
    parallel_for(n, threads, beam_item, &a);
}

typedef struct {
    uint64_t seed; uint32_t game0; int beam_width, search_depth, max_moves; orc_game_out *out;
} games_args;

static void games_item(int64_t i, void *p)
{
    games_args *a = (games_args *)p;
    orc_play_game(a->seed, a->game0 + (uint32_t)i, a->beam_width, a->search_depth, 512, 1024,
                  a->max_moves, &a->out[i]);
}

void orc_play_games(uint64_t seed, uint32_t game0, int64_t n, int beam_width, int search_depth,
                    int max_moves, orc_game_out *out, int threads)
{
    games_args a = { seed, game0, beam_width, search_depth, max_moves, out };
    parallel_for(n, threads, games_item, &a);
}
