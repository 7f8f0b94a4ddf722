/*
 * orc2048.h -- CPU oracle for the batched 2048 engine.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's hot path
 *   /root/reference/environment/game_2048.py      (env transition)
 *   /root/reference/agents/beam_search_agent.py   (beam search + heuristics)
 * working on unpacked int32 tile VALUES (like the reference, unlike the CUDA
 * product which works on packed nibble exponents), so that it is an
 * independent check of the kernels.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product never does.
 *
 * Parity pin: the reference has no tests or golden vectors, so
 * this file is pinned against the live reference itself: tests/test_oracle_vs_reference.py
 * drives both with the same injected Philox spawn stream (oracle/philox.py
 * shim), and tests/golden/ holds vectors generated from the live reference by
 * oracle/make_golden.py.
 */
#ifndef ORC2048_H
#define ORC2048_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Random stream domains (word c3 of the Philox counter). */
enum { ORC_DOM_ENV = 0, ORC_DOM_BEAM = 1, ORC_DOM_ACTION = 2, ORC_DOM_BOARD = 3, ORC_DOM_HYBRID = 4 };

/* Philox4x32-10.  ctr/key/out are plain uint32 words. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* The i-th (position word, value word) pair of stream (seed, game, call, domain). */
void orc_spawn_words(uint64_t seed, uint32_t game, uint32_t call, uint32_t domain,
                     uint32_t i, uint32_t *pos_word, uint32_t *val_word);
/* Uniform action in 0..3 for step t of the random-policy stream. */
int orc_random_action(uint64_t seed, uint32_t game, uint32_t t);
/* Synthetic mid-game board: each cell empty w.p. ~0.3, else 2^U{1..11}. */
void orc_synthetic_board(uint64_t seed, uint32_t game, int32_t board[16]);

/* ---- environment (game_2048.py) ---- */
typedef struct {
    int32_t  board[16];     /* row-major tile values, 0 = empty          */
    int64_t  score;         /* cumulative merge score                    */
    int32_t  highest_tile;  /* env.highest_tile                          */
    int32_t  game_over;     /* env.game_over                             */
    uint32_t spawn_ctr;     /* spawns drawn so far from the env stream   */
    uint32_t game;          /* global game id (Philox counter word c2)   */
    uint64_t seed;
} orc_env;

/* game_2048.py:116-168 on one direction; returns gained score, writes board in place */
int64_t orc_env_move(int32_t board[16], int action);
/* game_2048.py:69-95 -> bit a set when action a changes the board */
int orc_env_legal_mask(const int32_t board[16]);
/* game_2048.py:29-48; continues the env stream at env->spawn_ctr */
void orc_env_reset(orc_env *env);
/* game_2048.py:170-210.  inject: NULL or two raw words {pos_word, val_word}
 * used instead of the env stream (the stream counter is not advanced then). */
typedef struct { double reward; int32_t valid; int32_t done; int64_t score_delta; } orc_step_out;
void orc_env_step(orc_env *env, int action, const uint32_t *inject, orc_step_out *out);
/* game_2048.py:212-277 as a pure function */
double orc_env_reward(int valid, const int32_t prev_board[16], const int32_t new_board[16],
                      int64_t score_delta, int32_t highest_tile_before);

/* game_2048.py:341-387 (with its accumulating-board quirk) and :313-339 */
int orc_env_simulate_move(const int32_t state[16], int action, int32_t highest_tile,
                          int32_t out_boards[30][16], double out_reward[30], int32_t out_done[30]);
double orc_env_pattern(const int32_t board[16]);
/* agents/hybrid.py:578-692: the hybrid agent's sampled one-move expansion */
int orc_hybrid_simulate_move(const int32_t board[16], int action, uint64_t seed, uint32_t game, uint32_t call,
                             uint32_t *draw, int32_t out_boards[6][16], double out_reward[6], int32_t out_done[6]);

/* ---- beam-search agent (beam_search_agent.py) ---- */
/* :194-258 incl. the DOWN quirk.  Returns valid flag. */
int orc_agent_move(const int32_t board[16], int action, int32_t out[16], int64_t *merge_score);
int orc_agent_legal_mask(const int32_t board[16]);           /* :183-192 */
double orc_fast_eval(const int32_t board[16]);               /* :280-314 */
/* :316-403; phase 0 early, 1 mid, 2 late */
double orc_full_eval(const int32_t board[16], int phase);
int orc_phase(int32_t max_tile, int32_t early_thr, int32_t mid_thr);  /* :271-278 */

/* ---- PPO-side features (agents/ppo_agent.py), SURVEY 8f row 1 ---- */
void orc_ppo_observe(const int32_t board[16], float obs[16]);      /* :184-195 */
double orc_ppo_heuristic(const int32_t board[16]);                 /* :271-333 */
double orc_ppo_top4_bonus(const int32_t board[16]);                /* :251-254 */
/* :234-269 the reward PPOAgent.remember stores; highest_tile_seen in/out, novel = state not yet in seen_states */
double orc_ppo_shape_reward(const int32_t state[16], const int32_t next_state[16], double reward,
                            int32_t *highest_tile_seen, int novel);

typedef struct {
    int32_t action;        /* chosen action                                */
    float   prob;          /* 0.5 or 1.0 as the reference returns           */
    int32_t nodes;         /* evaluated children (one eval call each)       */
    int32_t depth_used;    /* adaptive depth (actual_depth), 0 on fast exit */
    double  best_score;    /* score of candidates[0] at the last level      */
    int32_t spawns;        /* (pos,val) pairs consumed from the beam stream */
} orc_beam_out;

/* :71-181.  legal_mask < 0 => valid_moves=None (agent's own legality);
 * otherwise bit a = valid_moves[a]. */
void orc_beam_get_action(const int32_t board[16], int legal_mask,
                         int beam_width, int search_depth,
                         int32_t early_thr, int32_t mid_thr,
                         uint64_t seed, uint32_t game, uint32_t call,
                         orc_beam_out *out);

/* evaluate_beam_search.py:16-98 (run_game) restated: get_action(state) with
 * no valid_moves, env.step, until done or max_moves. */
typedef struct {
    int64_t score; int32_t highest_tile; int32_t moves; int32_t valid_moves;
    int32_t invalid_moves; int32_t milestone_move[8]; /* 64..8192, -1 = never */
    int64_t nodes;
} orc_game_out;
void orc_play_game(uint64_t seed, uint32_t game, int beam_width, int search_depth,
                   int32_t early_thr, int32_t mid_thr, int max_moves, orc_game_out *out);

/* ---- batched helpers (pthreads) used as the CPU baseline and by big parity tests ---- */
/* T steps of the random-policy rollout on n envs (auto-reset when done), see DESIGN.md. */
void orc_rollout(int32_t *boards /*[n][16] in/out*/, int64_t *score, int32_t *highest,
                 uint32_t *spawn_ctr, double *reward_sum, int32_t *episodes,
                 int64_t n, int steps, uint32_t t0, uint64_t seed, uint32_t game0, int threads);
void orc_beam_batch(const int32_t *boards /*[n][16]*/, int64_t n, int beam_width, int search_depth,
                    uint64_t seed, uint32_t game0, uint32_t call,
                    int32_t *action, float *prob, int32_t *nodes, double *best_score, int threads);
void orc_play_games(uint64_t seed, uint32_t game0, int64_t n, int beam_width, int search_depth,
                    int max_moves, orc_game_out *out, int threads);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
