/*
 * g2048.h -- C ABI of libg2048 (B200 / sm_100a batched 2048 engine).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (vivek-tiwari-vt/2048-Using-Reinforcement-Learning) is pure Python and has no
 * FFI of its own; each entry point below names the reference method(s) it
 * replaces (env = environment/game_2048.py, agent = agents/beam_search_agent.py)
 * and INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *  - Plain pointers and sizes only.  `g2048_*` take DEVICE pointers and a CUDA
 *    stream handle (`void*` = cudaStream_t, NULL = default stream); they enqueue
 *    and return without synchronising.  `g2048_host_*` take HOST pointers, do
 *    their own H2D/D2H copies and return after the results are in host memory.
 *  - The caller owns every buffer.  The library allocates only its lookup
 *    tables (g2048_init) and, for g2048_host_*, internal staging buffers.
 *  - Return value: 0 on success, a negative G2048_E* code otherwise;
 *    g2048_last_error() gives a message.  Nothing throws across the boundary.
 *  - Board: uint64, cell (r,c) = nibble 4r+c = log2(tile), 0 = empty
 *    (row r = bits [16r,16r+16), LEFT moves towards nibble 0 of a row).
 *  - Actions 0 LEFT, 1 UP, 2 RIGHT, 3 DOWN (env:11-16); any other value is a
 *    no-op / invalid move, as in env:97-114.
 *  - Random spawns come from Philox4x32-10, counter = (block, call, game, domain),
 *    key = seed; env i of a batch is game `game0 + i` (DESIGN.md "Random streams").
 *  - Optional outputs may be NULL.
 */
#ifndef G2048_H
#define G2048_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2048_ABI_VERSION 1

enum {
    G2048_OK        = 0,
    G2048_EINVAL    = -1,   /* bad argument (NULL required pointer, n < 0, width out of range ...) */
    G2048_ENOTINIT  = -2,   /* g2048_init not called for the current device */
    G2048_ECUDA     = -3,   /* a CUDA runtime call failed, see g2048_last_error() */
    G2048_ENODEVICE = -4    /* no CUDA device: there is no CPU fallback */
};

#define G2048_MAX_BEAM_WIDTH 32          /* widths up to this take the warp-shuffle fast path */
#define G2048_MAX_WIDE_BEAM_WIDTH 128    /* widths up to this are accepted (33 and up: shared-memory path) */

/* ---- lifecycle ---------------------------------------------------------- */
int         g2048_abi_version(void);
/* Builds the row tables and uploads them to `device`.  Idempotent per device. */
int         g2048_init(int device);
/* Makes `device` (already initialised) the one the calling thread's next calls run on. */
int         g2048_set_device(int device);
const char *g2048_last_error(void);
/* Sticky count of merges that would have produced a 65536 tile (nibble saturation;
 * the reference's int32 board has no ceiling, env:149).  Synchronises `stream`. */
int         g2048_overflow_count(uint64_t *count, void *stream);

/* ---- board format ------------------------------------------------------- */
/* int32[n][16] tile values (env.get_state(), env:50-57)  <->  packed boards */
int g2048_pack(const int32_t *values, uint64_t *boards, int64_t n, void *stream);
int g2048_unpack(const uint64_t *boards, int32_t *values, int64_t n, void *stream);
/* float32[n][16] observation log2(tile)/15, 0 for empty (agents/ppo_agent.py:184-195) */
int g2048_observe(const uint64_t *boards, float *obs, int64_t n, void *stream);
/* PPO-side features of agents/ppo_agent.py in one pass: obs = normalize_state
 * (:184-195, float32[n][16]); heuristic = evaluate_heuristic (:271-333, float64, bit-exact);
 * top4_bonus = 0.1 * sum(log2 of the four largest tiles) (:251-254).  Outputs optional. */
int g2048_ppo_features(const uint64_t *boards, float *obs, double *heuristic, double *top4_bonus,
                       int64_t n, void *stream);
/* PPOAgent.remember's reward shaping (agents/ppo_agent.py:234-269) for n transitions (state, next_state,
 * reward) in one call, float64 in the reference's order:
 *   + 5 * (log2 max(next) - log2 highest_tile_seen) when next_state holds a new highest tile (:241-246;
 *     highest_seen_exp[i] = log2 of the running maximum of env i, in/out, start it at 1 = tile 2, :171),
 *   - 2 * (log2 max(state) - log2 max(next)) when the maximum regressed (:249-251),
 *   + 0.1 * sum(log2 of the four largest tiles) (:254-256),
 *   + 0.2 when next_state was not in the agent's seen_states (:259-262), then it is added,
 *   + 0.3 * evaluate_heuristic(next_state) (:265-266).
 * seen_states is ONE set for the whole batch, an open-addressing table the caller owns (set_keys /
 * set_claims, uint64[set_capacity], capacity a power of two, prepared by g2048_novelty_set_init; pass NULL
 * to leave the novelty term out).  The result equals calling remember() for env 0, 1, ..., n-1 in order:
 * of several envs reaching the same new board in one call the lowest env gets the bonus.  `step` must grow
 * with every call on the same table.  novel[i] (optional) = the bonus was given; *set_dropped (optional,
 * device counter) counts states that found no slot within 256 probes (size the table for < 70 % load). */
int g2048_novelty_set_init(uint64_t *set_keys, uint64_t *set_claims, int64_t set_capacity, void *stream);
int g2048_ppo_shape_rewards(const uint64_t *state_boards, const uint64_t *next_boards, const double *reward_in,
                            uint8_t *highest_seen_exp, uint64_t *set_keys, uint64_t *set_claims, int64_t set_capacity,
                            uint32_t step, double *reward_out, uint8_t *novel, uint64_t *set_dropped,
                            int64_t n, void *stream);
/* Synthetic mid-game boards: cell empty w.p. ~0.3 else 2^U{1..11} (bench workloads). */
int g2048_synthetic_boards(uint64_t *boards, int64_t n, uint64_t seed, uint32_t game0, void *stream);

/* ---- environment (environment/game_2048.py) ------------------------------ */
/* Game2048Env.reset (env:29-48): zero board, score 0, two spawns continuing the env
 * stream at spawn_ctr[i], highest = max exponent.  spawn_ctr is in/out. */
int g2048_env_reset(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                    int64_t n, uint64_t seed, uint32_t game0, void *stream);

/* Game2048Env.reset for exactly the envs with done[i] != 0 (what a rollout harness does after a
 * step, `if done: state = env.reset()` -- train.py:49,107), without leaving the device.
 * episodes[i] += 1 for those envs (optional). */
int g2048_env_reset_done(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                         const uint8_t *done, int32_t *episodes, int64_t n, uint64_t seed, uint32_t game0,
                         void *stream);

/* Game2048Env.step (env:170-210) for n independent envs.
 *   in/out : boards, score (cumulative merge score), highest_exp (log2 highest_tile), spawn_ctr
 *   in     : actions[n]; spawn_inject NULL or uint32[n][2] raw (position word, value word)
 *            used instead of the env stream (spawn_ctr is then left untouched)
 *   out    : reward (float64, bit-exact env:212-277), reward32, score_delta,
 *            valid (info["valid_move"]), legal (get_valid_moves of the NEW board, bit a),
 *            done (is_game_over, env:279-288).  All optional. */
int g2048_env_step(uint64_t *boards, const uint8_t *actions, const uint32_t *spawn_inject,
                   int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                   double *reward, float *reward32, int32_t *score_delta,
                   uint8_t *valid, uint8_t *legal, uint8_t *done,
                   int64_t n, uint64_t seed, uint32_t game0, void *stream);

/* Game2048Env.simulate_move (env:341-387): every (empty cell of the moved board, tile 2 then 4)
 * outcome of `actions[i]` on `boards[i]`, in the reference's order and with its accumulating-board
 * quirk (env:371,378: each outcome starts from the previous one; the reward of env:375 is taken on
 * that previous board).  Outcome k of board i is at index 32*i + k, count[i] <= 30 (0 = invalid move).
 * highest_exp: log2(env.highest_tile) per board or NULL (0). */
int g2048_simulate_move(const uint64_t *boards, const uint8_t *actions, const uint8_t *highest_exp,
                        uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count,
                        int64_t n, void *stream);
/* The hybrid agent's sampled one-move expansion, agents/hybrid.py:578-692 (its own simulate_move and
 * _calculate_simulation_reward): outcome k of (boards[i], actions[i]) at index 8*i + k, count[i] <= 6.
 * Invalid move: one outcome (board, -1.0, not done).  Else min(3, #empty) cells, picked by a partial
 * Fisher-Yates over the row-major empty list with draws draw0[i].. of stream (seed, game0+i,
 * call[i] or call0, domain 4), each as a 2-tile (reward*0.9) and a 4-tile (reward*0.1) outcome.
 * draws[i] = picks made.  Optional: call, draw0, every output. */
int g2048_hybrid_expand(const uint64_t *boards, const uint8_t *actions, const uint32_t *call, uint32_t call0,
                        const uint32_t *draw0, uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count,
                        uint32_t *draws, int64_t n, uint64_t seed, uint32_t game0, void *stream);
/* The same expansion for the items of a beam (DQNAgent.beam_search, agents/hybrid.py:814-907): item i belongs to
 * game `game0 + game[i]`, so that all (beam entry, action) items of one game draw from ONE stream at the
 * offsets draw0[i] the caller assigns in the reference's order (hybrid.py:840-844). */
int g2048_hybrid_expand_items(const uint64_t *boards, const uint8_t *actions, const uint32_t *game, const uint32_t *call,
                              const uint32_t *draw0, uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count,
                              uint32_t *draws, int64_t n, uint64_t seed, uint32_t game0, void *stream);
/* Game2048Env._evaluate_pattern (env:313-339): max(snake, corner weighted tile sums) / 100. */
int g2048_evaluate_pattern(const uint64_t *boards, double *pattern, int64_t n, void *stream);

/* g2048_env_step followed, for the envs whose game just ended, by Game2048Env.reset -- the
 * `if done: state = env.reset()` of every training loop (train.py:49,107), in the same launch.
 * done[i] still reports the end of the old game; boards/score/highest/legal describe the fresh one;
 * episodes[i] += 1.  spawn_ctr and episodes are required. */
int g2048_env_step_autoreset(uint64_t *boards, const uint8_t *actions,
                             int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                             double *reward, float *reward32, int32_t *score_delta,
                             uint8_t *valid, uint8_t *legal, uint8_t *done, int32_t *episodes,
                             int64_t n, uint64_t seed, uint32_t game0, void *stream);

/* The per-step call of a training loop in ONE launch: g2048_env_step, then
 *   stepped_boards / final_score / final_highest_exp = the state right after the step (the `next_state`,
 *                    info["score"], info["highest_tile"] a loop reads at `done`, train.py:74-133),
 *   the reset of the envs whose game ended (only when `episodes` is given; episodes[i] += 1),
 *   legal          = get_valid_moves of the board the caller acts on next (after a reset: the fresh one),
 *   obs            = float32[n][16] PPOAgent.normalize_state of that board (agents/ppo_agent.py:184-195).
 * Every output is optional; boards, actions are required. */
int g2048_env_step_fused(uint64_t *boards, const uint8_t *actions,
                         int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                         double *reward, float *reward32, int32_t *score_delta,
                         uint8_t *valid, uint8_t *legal, uint8_t *done, int32_t *episodes,
                         float *obs, uint64_t *stepped_boards, int32_t *final_score, uint8_t *final_highest_exp,
                         int64_t n, uint64_t seed, uint32_t game0, void *stream);

/* Game2048Env.get_valid_moves (env:69-95) -> env_legal; BeamSearchAgent._check_valid_moves
 * (agent:183-192, DOWN quirk included) -> agent_legal.  Bit a = action a.  Either may be NULL. */
int g2048_legal_masks(const uint64_t *boards, uint8_t *env_legal, uint8_t *agent_legal,
                      int64_t n, void *stream);

/* Fused rollout: `steps` consecutive env.step calls per env with boards held in registers.
 * Action of env i at step t0+s comes from the Philox action stream (uniform in 0..3);
 * when an env reports done it is reset (env.reset, same env stream) and episodes[i] += 1.
 * reward_sum[i] += the float64 rewards in step order.  All state arrays are in/out. */
int g2048_env_rollout(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                      double *reward_sum, int32_t *episodes,
                      int64_t n, int32_t steps, uint32_t t0, uint64_t seed, uint32_t game0, void *stream);

/* ---- beam search (agents/beam_search_agent.py) --------------------------- */
/* _fast_evaluate (agent:280-314) -> fast[n] (exact integer);
 * _evaluate_state (agent:316-403) -> full[n][3] float64 for phases early/mid/late, bit-exact. */
int g2048_evaluate(const uint64_t *boards, int32_t *fast, double *full, int64_t n, void *stream);

/* BeamSearchAgent.get_action (agent:71-181) for n independent root boards, one warp per root.
 *   legal    : NULL (valid_moves=None -> the agent's own legality) or uint8[n] bit masks
 *              (valid_moves passed by the caller, e.g. env.get_valid_moves())
 *   call     : NULL or uint32[n] per-root call index for the beam stream; else `call0` for all
 *   out      : action[n], prob[n] (0.5 / 1.0 as the reference returns), best_score[n]
 *              (score of candidates[0] at the last level), nodes[n] (evaluated children)
 *   beam_width 1..G2048_MAX_WIDE_BEAM_WIDTH (33 and up take a slower shared-memory path), search_depth >= 1,
 *   early_thr / mid_thr = agent.early_game_threshold / mid_game_threshold (512 / 1024). */
int g2048_beam_search(const uint64_t *roots, const uint8_t *legal, const uint32_t *call, uint32_t call0,
                      uint8_t *action, float *prob, double *best_score, int32_t *nodes,
                      int64_t n, int32_t beam_width, int32_t search_depth,
                      int32_t early_thr, int32_t mid_thr,
                      uint64_t seed, uint32_t game0, void *stream);

/* Whole games, evaluate_beam_search.run_game (evaluate_beam_search.py:16-98): Game2048Env(),
 * reset(), then get_action(state) / env.step(action) until done or max_moves.
 * Per-game outputs (all optional): score, highest_exp, moves, valid, invalid,
 * milestone[n][8] = 0-based index of the first move after which max tile >= 64,128,...,8192
 * (evaluate_beam_search.py:59-64; -1 = never),
 * nodes (evaluated children), final_board. */
int g2048_play_games(int64_t n, int32_t beam_width, int32_t search_depth,
                     int32_t early_thr, int32_t mid_thr, int32_t max_moves,
                     uint64_t seed, uint32_t game0,
                     int32_t *score, uint8_t *highest_exp, int32_t *moves, int32_t *valid, int32_t *invalid,
                     int32_t *milestone, int64_t *nodes, uint64_t *final_board, void *stream);

/* Statistics of a batch of finished games (evaluate_beam_search.py:127-135,201-213):
 * stats[0..17]  histogram of highest_exp 0..17, [18] sum score, [19] sum moves, [20] sum valid,
 * [21] sum invalid, [22] games, [23] max score, [24..31] games that reached 64..8192,
 * [32] sum nodes.  int64[G2048_STATS_LEN], ADDED to (max'ed into [23]) the existing
 * contents so that per-rank vectors can be all-reduced (sum; [23] with max). */
#define G2048_STATS_LEN 40
#define G2048_STATS_MAXSCORE 23
int g2048_stats_reduce(const int32_t *score, const uint8_t *highest_exp, const int32_t *moves,
                       const int32_t *valid, const int32_t *invalid, const int32_t *milestone,
                       const int64_t *nodes, int64_t n, int64_t *stats, void *stream);

/* ---- host-buffer entry points (H2D + kernels + D2H inside the call) ------ */
/* What a reference-side binding calls with numpy buffers; see INTEGRATION.md. */
int g2048_host_env_step(uint64_t *boards, const uint8_t *actions, const uint32_t *spawn_inject,
                        int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                        double *reward, int32_t *score_delta, uint8_t *valid, uint8_t *legal, uint8_t *done,
                        int64_t n, uint64_t seed, uint32_t game0);
int g2048_host_env_reset(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                         int64_t n, uint64_t seed, uint32_t game0);
int g2048_host_env_rollout(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                           double *reward_sum, int32_t *episodes,
                           int64_t n, int32_t steps, uint32_t t0, uint64_t seed, uint32_t game0);
int g2048_host_legal_masks(const uint64_t *boards, uint8_t *env_legal, uint8_t *agent_legal, int64_t n);
int g2048_host_evaluate(const uint64_t *boards, int32_t *fast, double *full, int64_t n);
int g2048_host_hybrid_expand(const uint64_t *boards, const uint8_t *actions, const uint32_t *call, uint32_t call0,
                             const uint32_t *draw0, uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count,
                             uint32_t *draws, int64_t n, uint64_t seed, uint32_t game0);
int g2048_host_simulate_move(const uint64_t *boards, const uint8_t *actions, const uint8_t *highest_exp,
                             uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count, double *pattern, int64_t n);
int g2048_host_ppo_features(const uint64_t *boards, float *obs, double *heuristic, double *top4_bonus, int64_t n);
int g2048_host_beam_search(const uint64_t *roots, const uint8_t *legal, const uint32_t *call, uint32_t call0,
                           uint8_t *action, float *prob, double *best_score, int32_t *nodes,
                           int64_t n, int32_t beam_width, int32_t search_depth,
                           int32_t early_thr, int32_t mid_thr, uint64_t seed, uint32_t game0);
int g2048_host_play_games(int64_t n, int32_t beam_width, int32_t search_depth,
                          int32_t early_thr, int32_t mid_thr, int32_t max_moves,
                          uint64_t seed, uint32_t game0,
                          int32_t *score, uint8_t *highest_exp, int32_t *moves, int32_t *valid, int32_t *invalid,
                          int32_t *milestone, int64_t *nodes, uint64_t *final_board, int64_t *stats);

/* Scheduling knobs of the beam-search launchers (results never depend on them; tests and
 * profiles/ use them to force a path).  Process-wide.
 *   G2048_TUNE_SEARCH_MODE      g2048_beam_search: 0 = auto (one team of four warps per root while the
 *                               roots fit the GPU's team slots, else one warp per root), 1 = always one
 *                               warp per root, 2 = always teams
 *   G2048_TUNE_TEAM_DIRECT_MAX  g2048_play_games: up to this many games are played by teams from the
 *                               first move (-1 = default: the team slots of the device, six per SM)
 *   G2048_TUNE_TAIL_THRESHOLD   g2048_play_games with more games: one warp per game until this many are
 *                               left alive, then teams (-1 = default: the team slots; 0 = never)
 *   G2048_TUNE_STEP_TABLES      g2048_env_step*: 0 = table-free SWAR row move, 1 = row tables read through L1/L2,
 *                               -1 = default: table-free below 262,144 envs per launch, tables from there on
 *   G2048_TUNE_SPLIT_STALLS     g2048_play_games: 1 = the calls of a stall (the agent keeps choosing an invalid move)
 *                               are cut into ranges that every free warp of the GPU searches (default); 0 = a stalled
 *                               game is parked for the stall breaker (one-warp kernel) or played through move by
 *                               move (team kernel): a test path, slow on games that stall up to the move cap
 *   G2048_TUNE_STEP_BLOCK_WARPS g2048_env_step*: warps per block (-1 = default: sized to the batch, see env_step_launch)
 *   G2048_TUNE_PDL              g2048_env_step*: 1 = launched with programmatic dependent launch, 0 = plain,
 *                               -1 = default: with it (and blocks of one / seven warps) up to 65,536 and from 196,608
 *                               envs per launch on, plain in between (see env_step_launch)
 *   G2048_TUNE_PENDING_CAP      g2048_play_games: slots of the ring through which resumed / migrating games reach a
 *                               free team (-1 = default: games + 64, which can never fill up; tests shrink it to
 *                               exercise the wrap-around; a ring smaller than the
 *                               number of games can block for good when every consumer waits to push -- experiments only)
 *   G2048_TUNE_STEP_OUTPUTS     g2048_env_step*: -1 = default: the host looks at the optional pointers once and launches
 *                               the kernel variant that takes the set it found for granted (2 = all of them, 1 = the
 *                               core set, see env_step_fused_kernel); 0 / 1 = use at most that variant (0 tests every
 *                               optional array per thread).  A test path: every variant must give the same results */
enum { G2048_TUNE_SEARCH_MODE = 0, G2048_TUNE_TEAM_DIRECT_MAX = 1, G2048_TUNE_TAIL_THRESHOLD = 2,
       G2048_TUNE_STEP_TABLES = 3, G2048_TUNE_SPLIT_STALLS = 4, G2048_TUNE_STEP_BLOCK_WARPS = 5, G2048_TUNE_PDL = 6,
       G2048_TUNE_PENDING_CAP = 7, G2048_TUNE_STEP_OUTPUTS = 8, G2048_TUNE_COUNT = 9 };
int g2048_set_tuning(int key, int value);

/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
uint64_t g2048_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* G2048_H */
