// g2048_api.cu -- C ABI of libg2048 (include/g2048.h): table construction, environment
// kernels, board-format kernels, statistics, and the host-buffer entry points.
// There is no CPU fallback: every entry point needs a CUDA device.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "env.cuh"
#include "row_tables.h"
#include "stage.cuh"

namespace g2048 {

// ------------------------------------------------------------------------------------------
// host state
// ------------------------------------------------------------------------------------------
static DeviceState g_dev[kMaxDevices];
static std::mutex g_mutex;
static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return code;
}
int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return G2048_OK;
    return set_error(G2048_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

DeviceState *current_device_state()
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        set_error(G2048_ENODEVICE, "no CUDA device (libg2048 has no CPU fallback)");
        return nullptr;
    }
    if (!g_dev[dev].ready) {
        set_error(G2048_ENOTINIT, "g2048_init(%d) has not been called", dev);
        return nullptr;
    }
    return &g_dev[dev];
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
#ifndef G2048_ENV_THREADS
#define G2048_ENV_THREADS 256
#endif
constexpr int kEnvThreads = G2048_ENV_THREADS;

// Both row tables into dynamic shared memory (192 KiB) by TMA bulk copy (stage.cuh).
__device__ __forceinline__ void stage_tables(uint8_t *smem, const uint16_t *row, const uint8_t *code, bool with_code)
{
    stage_bulk(smem, row, (uint32_t)kRowTableBytes, smem + kRowTableBytes, code, with_code ? (uint32_t)kCodeTableBytes : 0u);
}

struct StepArgs {
    uint64_t *boards; const uint8_t *actions; const uint32_t *inject;
    int32_t *score; uint8_t *highest; uint32_t *spawn_ctr;
    double *reward; float *reward32; int32_t *score_delta; uint8_t *valid; uint8_t *legal; uint8_t *done;
    int64_t n; PhiloxKey K; uint32_t game0;
    const uint16_t *row; const uint8_t *code; unsigned long long *overflow;
    const uint32_t *tables;        // DeviceState::step_tables
    int32_t *episodes;             // non-null: reset an env right after the step that ended its game
    // fused extras (all optional)
    float *obs;                    // float32[n][16] observation of the board the caller acts on next
    uint64_t *stepped;             // board after the step, BEFORE an auto-reset (the `next_state` of the transition)
    int32_t *final_score;          // score / highest exponent after the step, before an auto-reset
    uint8_t *final_highest;
    int32_t tables_early;          // the block publishes its copy of the tables before it waits for the previous launch (PDL)
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization
// attribute may start while its predecessor in the stream drains; `pdl_wait` blocks until that
// predecessor has completed and its writes are visible.  Everything before it (building the
// block's lookup tables) overlaps the predecessor's tail; without the attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// A global load that stays where it is written: ptxas sinks plain loads (and non-volatile asm loads) to their first use.
__device__ __forceinline__ uint64_t load_now(const uint64_t *p)
{
    uint64_t v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint32_t load_now(const uint32_t *p)
{
    uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint32_t load_now(const uint8_t *p)
{
    uint32_t v; asm volatile("ld.volatile.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// The reset of a finished game.  (-DG2048_STEP_RESET_NOINLINE moves it out of line: few envs end in a given
// step; measured slower -- the call's register convention costs every thread more than the skipped fetch saves.)
#ifdef G2048_STEP_RESET_NOINLINE
__device__ __noinline__ void env_reset_cold(EnvState &s, const PhiloxKey &K, uint32_t game) { env_reset(s, K, game); }
#else
__device__ __forceinline__ void env_reset_cold(EnvState &s, const PhiloxKey &K, uint32_t game) { env_reset(s, K, game); }
#endif

// Game2048Env.step for n envs, one env per thread, one launch per step, with everything a
// training loop wants next to it in the same launch: legal mask, done, the reset of a finished
// game, the policy's observation (agents/ppo_agent.py:184-195) and the pre-reset state.
// No 192 KiB tables: the move is table-free SWAR (kSwarMove) or reads the tables through L1/L2;
// scores and tile sums come from a 2 KiB pair table built per block.
constexpr int kStepMaxThreads = 1024;
// kOutputs: what the host already knows about the optional arrays (env_step_launch looks at the pointers once, so the
// kernel does not test them per thread: a test is a 64-bit compare and a branch around every load / store, ~90 of the
// step's instructions when everything is asked for, and the instruction count is the launch time here):
//   0  any combination;  1  score, highest, spawn_ctr, reward, reward32, score_delta, valid, legal, done present and no
//   injected spawn words (every step of BatchedGame2048Env);  2  also episodes, obs, stepped, final_score, final_highest
//   (BatchedGame2048Env.step_fused of a PPO rollout).
template <int kOutputs, int kLevel, class T>
__device__ __forceinline__ bool present(T *p) { return kOutputs >= kLevel || p != nullptr; }
template <bool kSwarMove, int kOutputs>
__global__ void __launch_bounds__(kStepMaxThreads) env_step_fused_kernel(StepArgs a)
{
    __shared__ __align__(16) uint32_t tables[kStepTableWords];
    // 1,024 words = 256 16-byte pieces (constant tables, written once by g2048_init: reading them does not have to
    // wait for the launch before this one).  A thread fetches piece tid into registers here and stores it only where
    // the step first needs a table (`tables_ready`, after the table-free move): the fetch, the state loads and the
    // move then overlap instead of queueing behind one another.  Blocks of fewer than 8 warps copy the rest in a
    // plain loop, one piece in flight per thread.  (Measured and dropped: reading the tables in place through L1 --
    // 2.7 -> 3.6 us per step at 4,096 envs -- and, for the one-warp blocks of small batches, all seven loads of a
    // thread in flight together -- 3.0 -> 4.7 us at 32,768 envs: those blocks belong to the NEXT launch and run their
    // prologue while the current one computes; the slow loop keeps them out of its way.)
    static_assert(kStepTableWords == 1024, "copied as 256 16-byte pieces, one per thread of the first eight warps");
    uint4 piece = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < 256u) piece = __ldg(reinterpret_cast<const uint4 *>(a.tables) + threadIdx.x);
#pragma unroll 1
    for (unsigned int extra = threadIdx.x + blockDim.x; extra < 256u; extra += blockDim.x)
        reinterpret_cast<uint4 *>(tables)[extra] = __ldg(reinterpret_cast<const uint4 *>(a.tables) + extra);
    const uint32_t *pairs = tables;
    const char *obs_pairs = reinterpret_cast<const char *>(tables + kPairEntries);      // float2[256]
    // With programmatic dependent launch the block is here while the launch before it still runs: it publishes the
    // tables at once (that wait is free) instead of inside the step.
    const bool early = a.tables_early != 0;
    auto publish = [&]() {
        if (threadIdx.x < 256u) reinterpret_cast<uint4 *>(tables)[threadIdx.x] = piece;
        __syncthreads();
    };
    auto tables_ready = [&]() { if (!early) publish(); };
    if (early) publish();
    pdl_launch_dependents();
    pdl_wait();
    // one env per thread; the launch covers the batch (env_step_launch), so there is no loop.  The threads past the end
    // of the batch run the step of the last env (every thread has to reach the barrier inside it) and store nothing.
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < (uint32_t)a.n;
    const uint32_t j = live ? i : (uint32_t)a.n - 1u;
    // (all state loads are issued here, together: left to itself the compiler sinks the ones the move does not need
    // to their first use, behind the barrier, and the spawn's Philox then waits a second memory round trip)
    EnvState s;
    s.board = Board(load_now(a.boards + j));
    const uint32_t action = load_now(a.actions + j);
    s.spawn_ctr = present<kOutputs, 1>(a.spawn_ctr) ? load_now(a.spawn_ctr + j) : 0u;
    s.score = present<kOutputs, 1>(a.score) ? (int32_t)load_now(reinterpret_cast<const uint32_t *>(a.score) + j) : 0;
    s.highest = present<kOutputs, 1>(a.highest) ? load_now(a.highest + j) : 0u;
    uint32_t inj[2];
    const bool injected = kOutputs == 0 && a.inject != nullptr;
    if (injected) { inj[0] = a.inject[2 * (size_t)j]; inj[1] = a.inject[2 * (size_t)j + 1]; }
    const uint32_t game = a.game0 + j;
    // one code path: a launch runs every instruction once per warp, and a warp runs them nearly one dependent
    // instruction at a time, so the instruction count is the launch time; a second, reward-free copy of the step
    // would only add instruction-cache misses
    StepResult2 r = env_step_pairs<kSwarMove, true>(s, action, a.row, a.code, pairs, a.K, game, injected ? inj : nullptr,
                                                    live ? a.overflow : nullptr, tables_ready);
    if (!live) return;
    if (present<kOutputs, 2>(a.stepped)) a.stepped[i] = s.board.u64();
    if (present<kOutputs, 2>(a.final_score)) a.final_score[i] = s.score;
    if (present<kOutputs, 2>(a.final_highest)) a.final_highest[i] = (uint8_t)s.highest;
    if (present<kOutputs, 2>(a.episodes) && r.done) {   // `if done: state = env.reset()` of the caller's loop (train.py:49,107)
        env_reset_cold(s, a.K, game);
        a.episodes[i] += 1;
        r.legal = env_legal_mask(s.board);
    }
    a.boards[i] = s.board.u64();
    if (present<kOutputs, 1>(a.score)) a.score[i] = s.score;
    if (present<kOutputs, 1>(a.highest)) a.highest[i] = (uint8_t)s.highest;
    if (present<kOutputs, 1>(a.spawn_ctr)) a.spawn_ctr[i] = s.spawn_ctr;
    if (present<kOutputs, 1>(a.reward)) a.reward[i] = r.reward;
    if (present<kOutputs, 1>(a.reward32)) a.reward32[i] = (float)r.reward;
    if (present<kOutputs, 1>(a.score_delta)) a.score_delta[i] = (int32_t)r.score_delta;
    if (present<kOutputs, 1>(a.valid)) a.valid[i] = r.valid;
    if (present<kOutputs, 1>(a.legal)) a.legal[i] = (uint8_t)r.legal;
    if (present<kOutputs, 1>(a.done)) a.done[i] = r.done;
    if (present<kOutputs, 2>(a.obs)) {
        // float32[16] = log2(tile) / 15 of the board the policy acts on next (after a reset, if there was one), from
        // byte-indexed float2 lookups.  Lanes 2j and 2j+1 write their two envs together: every 16-byte store of the
        // even lane and the one of the odd lane next to it fill one 32-byte sector (a lane writing its own 64 bytes
        // alone would half-fill two sectors per store, and the launch is then bound by store transactions).
        auto quad = [&](uint32_t cells16) {                 // four cells -> float4
            const float2 p = *reinterpret_cast<const float2 *>(obs_pairs + ((cells16 << 3) & 0x7F8u));
            const float2 q = *reinterpret_cast<const float2 *>(obs_pairs + ((cells16 >> 5) & 0x7F8u));
            return make_float4(p.x, p.y, q.x, q.y);
        };
        const uint32_t lo = s.board.lo, hi = s.board.hi;
        if ((i | 31u) < (uint32_t)a.n) {                    // the whole warp is here
            const uint32_t plo = __shfl_xor_sync(0xFFFFFFFFu, lo, 1), phi = __shfl_xor_sync(0xFFFFFFFFu, hi, 1);
            const bool odd = (i & 1u) != 0u;
            const uint32_t elo = odd ? plo : lo, ehi = odd ? phi : hi;       // board of the pair's even env
            const uint32_t olo = odd ? lo : plo, ohi = odd ? hi : phi;       // ... and of its odd env
            const uint32_t sh = odd ? 16u : 0u;
            float4 *even_env = reinterpret_cast<float4 *>(a.obs + 16 * (size_t)(i & ~1u)) + (odd ? 1 : 0);
            even_env[0] = quad(elo >> sh);
            even_env[4] = quad(olo >> sh);
            even_env[2] = quad(ehi >> sh);
            even_env[6] = quad(ohi >> sh);
        } else {
            float4 *out = reinterpret_cast<float4 *>(a.obs + 16 * (size_t)i);
            out[0] = quad(lo); out[1] = quad(lo >> 16); out[2] = quad(hi); out[3] = quad(hi >> 16);
        }
    }
}

__global__ void __launch_bounds__(kEnvThreads) env_reset_kernel(uint64_t *boards, int32_t *score, uint8_t *highest,
                                                                 uint32_t *spawn_ctr, int64_t n, PhiloxKey K,
                                                                 uint32_t game0)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        EnvState s;
        s.spawn_ctr = spawn_ctr ? spawn_ctr[i] : 0u;
        env_reset(s, K, game0 + (uint32_t)i);
        boards[i] = s.board.u64();
        if (score) score[i] = 0;
        if (highest) highest[i] = (uint8_t)s.highest;
        if (spawn_ctr) spawn_ctr[i] = s.spawn_ctr;
    }
}

// env.reset() for the envs whose done flag is set (the harness-side "reset when done" of a rollout
// loop, without a device->host round trip); the others are left untouched.
__global__ void __launch_bounds__(kEnvThreads) env_reset_done_kernel(uint64_t *boards, int32_t *score, uint8_t *highest,
                                                                      uint32_t *spawn_ctr, const uint8_t *done,
                                                                      int32_t *episodes, int64_t n, PhiloxKey K,
                                                                      uint32_t game0)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!done[i]) continue;
        EnvState s;
        s.spawn_ctr = spawn_ctr ? spawn_ctr[i] : 0u;
        env_reset(s, K, game0 + (uint32_t)i);
        boards[i] = s.board.u64();
        if (score) score[i] = 0;
        if (highest) highest[i] = (uint8_t)s.highest;
        if (spawn_ctr) spawn_ctr[i] = s.spawn_ctr;
        if (episodes) episodes[i] += 1;
    }
}

struct RolloutArgs {
    uint64_t *boards; int32_t *score; uint8_t *highest; uint32_t *spawn_ctr;
    double *reward_sum; int32_t *episodes;
    int64_t n; int32_t steps; uint32_t t0; PhiloxKey K; uint32_t game0;
    const uint16_t *row; const uint8_t *code; unsigned long long *overflow;
};

// `steps` env.step calls per env with the board in registers; random-policy actions from the
// Philox action stream (one block = 64 actions), auto-reset on game over.
#ifndef G2048_ROLLOUT_THREADS
#define G2048_ROLLOUT_THREADS 512
#endif
constexpr int kRolloutThreads = G2048_ROLLOUT_THREADS;

__global__ void __launch_bounds__(kRolloutThreads, 1) env_rollout_kernel(RolloutArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t pairs[kPairEntries];
    for (int i = threadIdx.x; i < kPairEntries; i += blockDim.x) pairs[i] = pair_table_entry(i);
    stage_tables(smem, a.row, a.code, true);                                  // its block-wide barrier also publishes `pairs`
    const uint16_t *row = reinterpret_cast<const uint16_t *>(smem);
    const uint8_t *code = smem + kRowTableBytes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t game = a.game0 + (uint32_t)i;
        TrackedEnv e;
        e.s.board = Board(a.boards[i]);
        e.s.score = a.score[i];
        e.s.highest = a.highest[i];
        e.s.spawn_ctr = a.spawn_ctr[i];
        double rsum = a.reward_sum ? a.reward_sum[i] : 0.0;
        int32_t episodes = a.episodes ? a.episodes[i] : 0;
        int32_t steps = a.steps;
        uint32_t t0 = a.t0;
        // A tile-less board (only a caller can hand one in: moves keep tiles, resets place two) has no
        // legal move: env.step reports an invalid move and game over (env:188,198), the harness resets.
        // The step loop below only tests FULL boards for game over, so this one step is taken here.
        if (steps > 0 && (e.s.board.lo | e.s.board.hi) == 0u) {
            rsum = __dadd_rn(rsum, shaped_reward(false, 16, e.s.board, 16, 0u, e.s.highest, 0u));
            env_reset(e.s, a.K, game);
            ++episodes; --steps; ++t0;
        }
        track(e);
        // highest > board max only if the caller poked it (env:229): such warps take the
        // variant that maintains both per step; everybody else skips that bookkeeping.
        const bool poked = __any_sync(__activemask(), e.s.highest > e.bmax);
        bool saturated;
        if (poked) saturated = rollout_steps<true>(e, steps, t0, a.K, game, row, code, pairs, rsum, episodes);
        else       saturated = rollout_steps<false>(e, steps, t0, a.K, game, row, code, pairs, rsum, episodes);
        if (saturated) atomicAdd(a.overflow, 1ull);
        a.boards[i] = e.s.board.u64();
        a.score[i] = e.s.score;
        a.highest[i] = (uint8_t)e.s.highest;
        a.spawn_ctr[i] = e.s.spawn_ctr;
        if (a.reward_sum) a.reward_sum[i] = rsum;
        if (a.episodes) a.episodes[i] = episodes;
    }
}

__global__ void legal_masks_kernel(const uint64_t *boards, uint8_t *env_legal, uint8_t *agent_legal, int64_t n,
                                   const uint16_t *row)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Board b(boards[i]);
        uint32_t m = env_legal_mask(b);
        if (env_legal) env_legal[i] = (uint8_t)m;
        if (agent_legal) {
            // agent:209-210 vs :251-253: the DOWN result comes back rotated by 180 degrees and is
            // compared with the input in that state
            Board down = rot180(env_move<false>(b, 3u, row));
            agent_legal[i] = (uint8_t)((m & 7u) | (down != b ? 8u : 0u));
        }
    }
}

__global__ void pack_kernel(const int32_t *values, uint64_t *boards, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 *v = reinterpret_cast<const int4 *>(values + 16 * i);
        uint64_t b = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int4 x = v[q];
            int t[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t e = t[j] > 0 ? 31u - (uint32_t)__clz(t[j]) : 0u;     // log2 of a power of two
                b |= (uint64_t)(e & 15u) << (4 * (4 * q + j));
            }
        }
        boards[i] = b;
    }
}

__global__ void unpack_kernel(const uint64_t *boards, int32_t *values, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t b = boards[i];
        int4 *v = reinterpret_cast<int4 *>(values + 16 * i);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t e = (uint32_t)(b >> (4 * (4 * q + j))) & 15u;
                t[j] = e ? (1 << e) : 0;
            }
            v[q] = make_int4(t[0], t[1], t[2], t[3]);
        }
    }
}

// ppo_agent.py:184-195: log2(tile)/15 for tile > 0, else 0
__global__ void observe_kernel(const uint64_t *boards, float *obs, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t b = boards[i];
        float4 *o = reinterpret_cast<float4 *>(obs + 16 * i);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = (float)((uint32_t)(b >> (4 * (4 * q + j))) & 15u) / 15.0f;
            o[q] = make_float4(t[0], t[1], t[2], t[3]);
        }
    }
}

// Game2048Env.simulate_move (env:341-387): all (empty cell, tile in {2,4}) outcomes of a move, one
// (board, action) per thread.  The reference's loop never restores self.board (env:378), so every
// outcome starts from the PREVIOUS outcome and the reward (env:375) is computed on that previous
// board; reproduced as is.  Outcome k of board i lands at index 32*i + k; count[i] <= 30.
constexpr int kSimStride = 32;
__global__ void simulate_move_kernel(const uint64_t *boards, const uint8_t *actions, const uint8_t *highest,
                                     uint64_t *out_boards, double *out_reward, uint8_t *out_done, int32_t *count,
                                     int64_t n, const uint16_t *row, const uint8_t *code)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Board state(boards[i]);
        const uint32_t action = actions[i];
        const uint32_t hi_exp = highest ? highest[i] : 0u;
        const Board line = to_line(state, action);
        Board cur = from_line(move_left<false>(line, row), action);
        uint32_t sat;
        uint32_t gained = decode_score(merge_codes<false>(line, code), &sat);
        if (action >= 4u) { cur = state; gained = 0u; }
        int k = 0;
        if (cur != state) {
            const uint64_t moved = cur.u64();
            const int empty_before = count_empty(state);
            const uint32_t prev_max = max_exponent(state);
            for (int cell = 0; cell < 16; ++cell) {
                if ((moved >> (4 * cell)) & 15ull) continue;                 // env:367: empties of the moved board
                for (uint32_t e = 1; e <= 2; ++e) {
                    uint64_t ns = (cur.u64() & ~(15ull << (4 * cell))) | ((uint64_t)e << (4 * cell));
                    double r = shaped_reward(true, empty_before, cur, count_empty(cur), gained, hi_exp, prev_max);
                    cur = Board(ns);
                    G2048_ASSERT(k < 30);
                    if (out_boards) out_boards[kSimStride * i + k] = ns;
                    if (out_reward) out_reward[kSimStride * i + k] = r;
                    if (out_done) out_done[kSimStride * i + k] = env_game_over(cur);
                    ++k;
                }
            }
        }
        if (count) count[i] = k;
    }
}

// The hybrid agent's sampled expansion (agents/hybrid.py:578-692): one (board,
// action) per thread -> up to 6 outcomes at index 8*i + k.  An invalid move gives the single
// outcome (board, -1.0, not done).  Otherwise min(3, #empty) cells are picked by a partial
// Fisher-Yates over the row-major empty list (one draw per pick, the shim's `random.sample`) and
// each yields a 2-tile outcome (reward * 0.9) and a 4-tile outcome (reward * 0.1), with
// reward = (sum(new) - sum(old)) + (new max if it grew) + 0.1 * #empty(new)  (hybrid.py:676-692).
constexpr int kHybridStride = 8;
__global__ void hybrid_expand_kernel(const uint64_t *boards, const uint8_t *actions, const uint32_t *call,
                                     const uint32_t *draw0, uint64_t *out_boards, double *out_reward, uint8_t *out_done,
                                     int32_t *count, uint32_t *draws, int64_t n, PhiloxKey K, uint32_t game0, uint32_t call0,
                                     const uint16_t *row, const uint32_t *game_of)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Board state(boards[i]);
        const uint32_t action = actions[i];
        Board moved = env_move<false>(state, action, row);
        int k = 0;
        uint32_t draw = draw0 ? draw0[i] : 0u, used = 0u;
        if (moved == state) {
            if (out_boards) out_boards[kHybridStride * i] = state.u64();
            if (out_reward) out_reward[kHybridStride * i] = -1.0;
            if (out_done) out_done[kHybridStride * i] = 0;
            k = 1;
        } else {
            const uint64_t m64 = moved.u64();
            int cells[16], ne = 0;
            for (int c = 0; c < 16; ++c) if (((m64 >> (4 * c)) & 15ull) == 0ull) cells[ne++] = c;
            const uint32_t old_max = max_exponent(state);
            const int picks = ne < 3 ? ne : 3;
            for (int p = 0; p < picks; ++p) {
                Philox4 blk = philox4x32_10(draw >> 2, call ? call[i] : call0, game0 + (game_of ? game_of[i] : (uint32_t)i), DOM_HYBRID, K);
                const uint32_t word = (draw & 3u) == 0 ? blk.w[0] : (draw & 3u) == 1 ? blk.w[1] : (draw & 3u) == 2 ? blk.w[2] : blk.w[3];
                ++draw; ++used;
                const int j = p + (int)__umulhi(word, (uint32_t)(ne - p));
                const int t = cells[p]; cells[p] = cells[j]; cells[j] = t;
                for (uint32_t e = 1; e <= 2; ++e) {
                    const uint64_t nb = m64 | ((uint64_t)e << (4 * cells[p]));
                    // moves conserve the tile sum: sum(new) - sum(old) is the spawned tile
                    const long long merge_reward = 1ll << e;
                    const uint32_t new_max = max(max_exponent(moved), e);
                    const long long bonus = new_max > old_max ? (1ll << new_max) : 0ll;
                    const double empty_bonus = __dmul_rn((double)(ne - 1), 0.1);
                    const double reward = __dadd_rn((double)(merge_reward + bonus), empty_bonus);
                    G2048_ASSERT(k < 6 && cells[p] >= 0 && cells[p] < 16);
                    if (out_boards) out_boards[kHybridStride * i + k] = nb;
                    if (out_reward) out_reward[kHybridStride * i + k] = __dmul_rn(reward, e == 1 ? 0.9 : 0.1);
                    if (out_done) out_done[kHybridStride * i + k] = 0;
                    ++k;
                }
            }
        }
        if (count) count[i] = k;
        if (draws) draws[i] = used;
    }
}

// Game2048Env._evaluate_pattern (env:313-339): max(snake-weighted, corner-weighted tile sum) / 100
__global__ void pattern_kernel(const uint64_t *boards, double *out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t b = boards[i];
        double snake = 0.0, corner = 0.0;          // exact: dyadic partial sums far below 2^53
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int r = c >> 2, q = c & 3;
            const double ws = (r & 1) ? (double)(16 - 4 * r - 3 + q) : (double)(16 - 4 * r - q);   // 16 15 14 13 / 9 10 11 12 / 8 7 6 5 / 1 2 3 4
            const double wc = 16.0 / (double)(1 << (r + q));                                         // 16 8 4 2 / 8 4 2 1 / ...
            uint32_t e = (uint32_t)(b >> (4 * c)) & 15u;
            double v = e ? (double)(1u << e) : 0.0;
            snake = __dadd_rn(snake, __dmul_rn(v, ws));
            corner = __dadd_rn(corner, __dmul_rn(v, wc));
        }
        snake = __ddiv_rn(snake, 100.0);
        corner = __ddiv_rn(corner, 100.0);
        out[i] = snake > corner ? snake : corner;
    }
}

// normalize_state + the per-state shaping terms of PPOAgent.remember (ppo_agent.py:184-195,
// 251-254, 271-333) in one pass over the boards; every output is optional.
__global__ void ppo_features_kernel(const uint64_t *boards, float *obs, double *heuristic, double *top4, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t raw = boards[i];
        Board b(raw);
        if (obs) {
            float4 *o = reinterpret_cast<float4 *>(obs + 16 * i);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] = (float)((uint32_t)(raw >> (4 * (4 * q + j))) & 15u) / 15.0f;
                o[q] = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        if (heuristic) heuristic[i] = ppo_heuristic(b);
        if (top4) top4[i] = ppo_top4_bonus(b);
    }
}

// ---- PPOAgent.remember's reward shaping (agents/ppo_agent.py:234-269), SURVEY 8f row 1 -----------------
// The agent's `seen_states` set as an open-addressing table of packed boards.  One agent remembers the
// transitions of all n envs in env order, every step; to reproduce that order in parallel a slot carries a
// CLAIM = (step << 32 | env): claims only shrink (atomicMin) and steps grow, so after the insert pass
// a slot holds the first step the board was ever seen in and the lowest env that saw it then -- exactly the
// env whose `state_hash not in self.seen_states` test succeeds in the sequential loop.
constexpr int kNoveltyProbes = 256;
__device__ __forceinline__ uint64_t novelty_key(uint64_t board) { return board ? board : ~0ull; }   // 0 marks a free slot
__device__ __forceinline__ uint64_t novelty_hash(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
__global__ void novelty_insert_kernel(const uint64_t *next_boards, unsigned long long *keys, unsigned long long *claims,
                                      uint64_t mask, uint32_t step, int64_t n, unsigned long long *dropped)
{
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = novelty_key(next_boards[i]);
        uint64_t slot = novelty_hash(key) & mask;
        bool placed = false;
        for (int probe = 0; probe < kNoveltyProbes; ++probe, slot = (slot + 1) & mask) {
            const unsigned long long old = atomicCAS(&keys[slot], 0ull, key);
            if (old == 0ull || old == key) {
                atomicMin(&claims[slot], ((unsigned long long)step << 32) | (unsigned long long)(uint32_t)i);
                placed = true;
                break;
            }
        }
        if (!placed && dropped) atomicAdd(dropped, 1ull);                  // table (nearly) full: the state is not recorded
    }
}
__global__ void ppo_shape_kernel(const uint64_t *state_boards, const uint64_t *next_boards, const double *reward_in,
                                 uint8_t *highest_seen, const unsigned long long *keys, const unsigned long long *claims,
                                 uint64_t mask, uint32_t step, double *reward_out, uint8_t *novel_out, int64_t n)
{
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Board cur(state_boards[i]), nxt(next_boards[i]);
        const uint32_t cur_max = max_exponent(cur), nxt_max = max_exponent(nxt);
        double reward = reward_in[i];
        uint32_t seen = highest_seen[i];
        if (nxt_max > seen) {                                              // ppo_agent.py:241-246
            reward = __dadd_rn(reward, __dmul_rn(5.0, (double)(int)(nxt_max - seen)));
            highest_seen[i] = (uint8_t)nxt_max;
        }
        if (nxt_max < cur_max)                                             // :249-251
            reward = __dadd_rn(reward, __dmul_rn(-2.0, (double)(int)(cur_max - nxt_max)));
        reward = __dadd_rn(reward, ppo_top4_bonus(nxt));                   // :254-256
        bool novel = false;
        if (keys) {                                                        // :259-262
            const unsigned long long key = novelty_key(next_boards[i]);
            uint64_t slot = novelty_hash(key) & mask;
            for (int probe = 0; probe < kNoveltyProbes; ++probe, slot = (slot + 1) & mask) {
                const unsigned long long k = keys[slot];
                if (k == key) { novel = claims[slot] == (((unsigned long long)step << 32) | (unsigned long long)(uint32_t)i); break; }
                if (k == 0ull) break;
            }
        }
        if (novel) reward = __dadd_rn(reward, 0.2);
        reward = __dadd_rn(reward, __dmul_rn(0.3, ppo_heuristic(nxt)));    // :265-266
        reward_out[i] = reward;
        if (novel_out) novel_out[i] = novel;
    }
}

__global__ void synthetic_kernel(uint64_t *boards, int64_t n, PhiloxKey K, uint32_t game0)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t b = 0;
#pragma unroll
        for (uint32_t blk = 0; blk < 4; ++blk) {
            Philox4 p = philox4x32_10(blk, 0u, game0 + (uint32_t)i, DOM_BOARD, K);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t x = p.w[j];
                uint32_t e = 1u + (((x >> 16) * 11u) >> 16);
                e = (x & 0xFFFFu) < 19661u ? 0u : e;
                b |= (uint64_t)e << (4 * (4 * blk + j));
            }
        }
        boards[i] = b;
    }
}

__global__ void evaluate_kernel(const uint64_t *boards, int32_t *fast, double *full, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Board b(boards[i]);
        int n0 = count_empty(b);
        uint32_t emax = max_exponent(b);
        if (fast) fast[i] = fast_eval(b, n0, emax);
        if (full) {
            full[3 * i + 0] = full_eval(b, n0, emax, 0);
            full[3 * i + 1] = full_eval(b, n0, emax, 1);
            full[3 * i + 2] = full_eval(b, n0, emax, 2);
        }
    }
}

// Per-rank game statistics; added into `stats` so that ranks can all-reduce the vector.
__global__ void stats_kernel(const int32_t *score, const uint8_t *highest, const int32_t *moves, const int32_t *valid,
                             const int32_t *invalid, const int32_t *milestone, const int64_t *nodes, int64_t n,
                             unsigned long long *stats)
{
    __shared__ unsigned long long acc[G2048_STATS_LEN];
    for (int j = threadIdx.x; j < G2048_STATS_LEN; j += blockDim.x) acc[j] = 0ull;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = highest ? highest[i] : 0u;
        atomicAdd(&acc[h < 18u ? h : 17u], 1ull);
        if (score) {
            atomicAdd(&acc[18], (unsigned long long)(long long)score[i]);
            atomicMax(reinterpret_cast<long long *>(&acc[G2048_STATS_MAXSCORE]), (long long)score[i]);
        }
        if (moves) atomicAdd(&acc[19], (unsigned long long)moves[i]);
        if (valid) atomicAdd(&acc[20], (unsigned long long)valid[i]);
        if (invalid) atomicAdd(&acc[21], (unsigned long long)invalid[i]);
        atomicAdd(&acc[22], 1ull);
        if (milestone)
            for (int m = 0; m < 8; ++m)
                if (milestone[8 * i + m] >= 0) atomicAdd(&acc[24 + m], 1ull);
        if (nodes) atomicAdd(&acc[32], (unsigned long long)nodes[i]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < G2048_STATS_LEN; j += blockDim.x) {
        if (j == G2048_STATS_MAXSCORE) atomicMax(reinterpret_cast<long long *>(&stats[j]), (long long)acc[j]);
        else if (acc[j]) atomicAdd(&stats[j], acc[j]);
    }
}

static int grid_for(int64_t n, int threads, int sm_count, int blocks_per_sm)
{
    int64_t want = (n + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

int step_tuning(int key);      // beam.cu: g2048_set_tuning values

// Launch with the programmatic-serialization attribute (see pdl_wait above).
template <typename... Params, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kernel)(Params...), int grid, int threads, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

}  // namespace g2048

using namespace g2048;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int g2048_abi_version(void) { return G2048_ABI_VERSION; }
const char *g2048_last_error(void) { return g_error; }
uint64_t g2048_launch_count(void) { return g_launches.load(); }
int g2048_set_tuning(int key, int value) { return set_tuning(key, value); }

int g2048_init(int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return set_error(G2048_ENODEVICE, "no CUDA device (libg2048 has no CPU fallback)");
    if (device < 0 || device >= count || device >= kMaxDevices) return set_error(G2048_EINVAL, "bad device %d", device);
    std::lock_guard<std::mutex> lock(g_mutex);
    G2048_CUDA(cudaSetDevice(device));
    DeviceState &st = g_dev[device];
    if (st.ready) return G2048_OK;
    std::vector<uint16_t> row(kRowEntries);
    std::vector<uint8_t> code(kRowEntries);
    build_row_tables(row.data(), code.data());
    G2048_CUDA(cudaMalloc(&st.row, kRowTableBytes));
    G2048_CUDA(cudaMalloc(&st.code, kCodeTableBytes));
    G2048_CUDA(cudaMalloc(&st.overflow, sizeof(unsigned long long)));
    G2048_CUDA(cudaMemcpy(st.row, row.data(), kRowTableBytes, cudaMemcpyHostToDevice));
    G2048_CUDA(cudaMemcpy(st.code, code.data(), kCodeTableBytes, cudaMemcpyHostToDevice));
    G2048_CUDA(cudaMemset(st.overflow, 0, sizeof(unsigned long long)));
    G2048_CUDA(cudaDeviceGetAttribute(&st.sm_count, cudaDevAttrMultiProcessorCount, device));
    std::vector<uint32_t> step_tables(kStepTableWords);
    for (int i = 0; i < kPairEntries; ++i) step_tables[i] = pair_table_entry((uint32_t)i);
    for (int b = 0; b < 256; ++b) {
        const float v[2] = {(float)(b & 15) / 15.0f, (float)(b >> 4) / 15.0f};  // log2(tile) / 15.0 in float32
        memcpy(&step_tables[kPairEntries + 2 * b], v, sizeof v);
    }
    G2048_CUDA(cudaMalloc(&st.step_tables, kStepTableWords * sizeof(uint32_t)));
    G2048_CUDA(cudaMemcpy(st.step_tables, step_tables.data(), kStepTableWords * sizeof(uint32_t), cudaMemcpyHostToDevice));
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    G2048_CUDA(cudaMemPoolCreate(&st.pool, &props));
    uint64_t keep = UINT64_MAX;
    G2048_CUDA(cudaMemPoolSetAttribute(st.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    const int smem = (int)(kRowTableBytes + kCodeTableBytes);
    G2048_CUDA(cudaFuncSetAttribute(env_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    st.ready = true;
    return G2048_OK;
}

int g2048_set_device(int device)
{
    if (device < 0 || device >= kMaxDevices || !g_dev[device].ready)
        return set_error(G2048_ENOTINIT, "g2048_init(%d) has not been called", device);
    G2048_CUDA(cudaSetDevice(device));
    return G2048_OK;
}

int g2048_overflow_count(uint64_t *count, void *stream)
{
    DeviceState *st = current_device_state();
    if (!st) return G2048_ENOTINIT;
    if (!count) return set_error(G2048_EINVAL, "count is NULL");
    unsigned long long v = 0;
    G2048_CUDA(cudaMemcpyAsync(&v, st->overflow, sizeof v, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    G2048_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *count = v;
    return G2048_OK;
}

#define G2048_ENTER(required_ok)                                                   \
    DeviceState *st = current_device_state();                                      \
    if (!st) return G2048_ENOTINIT;                                                \
    if (n < 0 || !(required_ok)) return set_error(G2048_EINVAL, "%s: bad argument", __func__); \
    if (n == 0) return G2048_OK;                                                   \
    cudaStream_t s = (cudaStream_t)stream

#define G2048_LAUNCHED()                                      \
    count_launch();                                           \
    return check_cuda(cudaGetLastError(), __func__)

int g2048_pack(const int32_t *values, uint64_t *boards, int64_t n, void *stream)
{
    G2048_ENTER(values && boards);
    pack_kernel<<<grid_for(n, 256, st->sm_count, 16), 256, 0, s>>>(values, boards, n);
    G2048_LAUNCHED();
}

int g2048_unpack(const uint64_t *boards, int32_t *values, int64_t n, void *stream)
{
    G2048_ENTER(values && boards);
    unpack_kernel<<<grid_for(n, 256, st->sm_count, 16), 256, 0, s>>>(boards, values, n);
    G2048_LAUNCHED();
}

int g2048_observe(const uint64_t *boards, float *obs, int64_t n, void *stream)
{
    G2048_ENTER(boards && obs);
    observe_kernel<<<grid_for(n, 256, st->sm_count, 16), 256, 0, s>>>(boards, obs, n);
    G2048_LAUNCHED();
}

int g2048_simulate_move(const uint64_t *boards, const uint8_t *actions, const uint8_t *highest_exp,
                        uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count, int64_t n, void *stream)
{
    G2048_ENTER(boards && actions);
    simulate_move_kernel<<<grid_for(n, 128, st->sm_count, 16), 128, 0, s>>>(boards, actions, highest_exp, next_boards, reward,
                                                                          done, count, n, st->row, st->code);
    G2048_LAUNCHED();
}

int g2048_hybrid_expand(const uint64_t *boards, const uint8_t *actions, const uint32_t *call, uint32_t call0,
                        const uint32_t *draw0, uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count,
                        uint32_t *draws, int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    G2048_ENTER(boards && actions);
    hybrid_expand_kernel<<<grid_for(n, 128, st->sm_count, 16), 128, 0, s>>>(boards, actions, call, draw0, next_boards, reward,
                                                                          done, count, draws, n, make_philox_key(seed), game0,
                                                                          call0, st->row, nullptr);
    G2048_LAUNCHED();
}

int g2048_hybrid_expand_items(const uint64_t *boards, const uint8_t *actions, const uint32_t *game, const uint32_t *call,
                              const uint32_t *draw0, uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count,
                              uint32_t *draws, int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    G2048_ENTER(boards && actions && game);
    hybrid_expand_kernel<<<grid_for(n, 128, st->sm_count, 16), 128, 0, s>>>(boards, actions, call, draw0, next_boards, reward,
                                                                          done, count, draws, n, make_philox_key(seed), game0,
                                                                          0u, st->row, game);
    G2048_LAUNCHED();
}

int g2048_evaluate_pattern(const uint64_t *boards, double *pattern, int64_t n, void *stream)
{
    G2048_ENTER(boards && pattern);
    pattern_kernel<<<grid_for(n, 256, st->sm_count, 8), 256, 0, s>>>(boards, pattern, n);
    G2048_LAUNCHED();
}

int g2048_ppo_features(const uint64_t *boards, float *obs, double *heuristic, double *top4_bonus, int64_t n, void *stream)
{
    G2048_ENTER(boards);
    ppo_features_kernel<<<grid_for(n, 256, st->sm_count, 8), 256, 0, s>>>(boards, obs, heuristic, top4_bonus, n);
    G2048_LAUNCHED();
}

int g2048_novelty_set_init(uint64_t *set_keys, uint64_t *set_claims, int64_t capacity, void *stream)
{
    const int64_t n = capacity;
    G2048_ENTER(set_keys && set_claims && (capacity & (capacity - 1)) == 0);
    G2048_CUDA(cudaMemsetAsync(set_keys, 0, (size_t)capacity * sizeof(uint64_t), s));
    return check_cuda(cudaMemsetAsync(set_claims, 0xFF, (size_t)capacity * sizeof(uint64_t), s), __func__);
}

int g2048_ppo_shape_rewards(const uint64_t *state_boards, const uint64_t *next_boards, const double *reward_in,
                            uint8_t *highest_seen_exp, uint64_t *set_keys, uint64_t *set_claims, int64_t set_capacity,
                            uint32_t step, double *reward_out, uint8_t *novel, uint64_t *set_dropped,
                            int64_t n, void *stream)
{
    G2048_ENTER(state_boards && next_boards && reward_in && highest_seen_exp && reward_out &&
                (!set_keys || (set_claims && set_capacity > 0 && (set_capacity & (set_capacity - 1)) == 0)));
    const int grid = grid_for(n, 256, st->sm_count, 8);
    const uint64_t mask = set_keys ? (uint64_t)set_capacity - 1 : 0;
    if (set_keys) {
        cudaError_t e = launch_pdl(true, novelty_insert_kernel, grid, 256, s, next_boards,
                                   reinterpret_cast<unsigned long long *>(set_keys),
                                   reinterpret_cast<unsigned long long *>(set_claims), mask, step, n,
                                   reinterpret_cast<unsigned long long *>(set_dropped));
        count_launch();
        G2048_CUDA(e);
    }
    cudaError_t e = launch_pdl(true, ppo_shape_kernel, grid, 256, s, state_boards, next_boards, reward_in, highest_seen_exp,
                               reinterpret_cast<const unsigned long long *>(set_keys),
                               reinterpret_cast<const unsigned long long *>(set_claims), mask, step, reward_out, novel, n);
    count_launch();
    return check_cuda(e, __func__);
}

int g2048_synthetic_boards(uint64_t *boards, int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    G2048_ENTER(boards);
    synthetic_kernel<<<grid_for(n, 256, st->sm_count, 16), 256, 0, s>>>(boards, n, make_philox_key(seed), game0);
    G2048_LAUNCHED();
}

int g2048_env_reset(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                    int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    G2048_ENTER(boards);
    env_reset_kernel<<<grid_for(n, kEnvThreads, st->sm_count, 8), kEnvThreads, 0, s>>>(
        boards, score, highest_exp, spawn_ctr, n, make_philox_key(seed), game0);
    G2048_LAUNCHED();
}

int g2048_env_reset_done(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                         const uint8_t *done, int32_t *episodes, int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    G2048_ENTER(boards && done);
    env_reset_done_kernel<<<grid_for(n, kEnvThreads, st->sm_count, 8), kEnvThreads, 0, s>>>(
        boards, score, highest_exp, spawn_ctr, done, episodes, n, make_philox_key(seed), game0);
    G2048_LAUNCHED();
}

static int env_step_launch(StepArgs a, void *stream)
{
    const int64_t n = a.n;
    G2048_ENTER(a.boards && a.actions);
    a.row = st->row; a.code = st->code; a.overflow = st->overflow; a.tables = st->step_tables; a.tables_early = 0;
    // Launch geometry (profiles/step_geometry.py; us per step as a CUDA graph, with / without the observation).
    //  * Programmatic dependent launch lets the next step's blocks become resident and copy their tables while this
    //    step still computes, which hides most of the ~1.1 us a link in a chain of dependent launches costs.  It pays
    //    when the blocks are small enough to enter SMs that are still busy: ONE-WARP blocks up to 65,536 envs
    //    (32,768 envs: 2.9 us against 4.2 us as 147 blocks of 7 warps without it; 16,384: 2.7 against 3.0) and blocks
    //    of 7 warps from 196,608 envs on (1,048,576: 41 / 30 us against 50 / 34).
    //  * In between, the waiting blocks of the next launch crowd the running one out (131,072 envs: 7.8 us with it,
    //    6.9 without), so those launches run plain, as one block per SM sized to the batch (<= 28 warps, else 14).
    const int64_t warps = (n + 31) / 32;
    int64_t per_block = (warps + st->sm_count - 1) / st->sm_count;
    if (per_block > 28) per_block = 14;
    const int pdl_knob = step_tuning(G2048_TUNE_PDL);
    const bool pdl = pdl_knob < 0 ? (n <= 65536 || n >= 196608) : pdl_knob != 0;
    if (pdl && pdl_knob < 0) per_block = n <= 65536 ? 1 : 7;
    const int forced = step_tuning(G2048_TUNE_STEP_BLOCK_WARPS);
    if (forced > 0 && forced <= kStepMaxThreads / 32) per_block = forced;
    const int threads = 32 * (int)per_block;
    const int grid = (int)((warps + per_block - 1) / per_block);
    a.tables_early = pdl ? 1 : 0;
    cudaError_t e;
    // table-free move: 15 % faster up to 65,536 envs (no dependent table round trip on a latency-bound launch);
    // at a million envs the launch is ALU-bound and the table reads are hidden: row tables win by 8 %
    const int tables = step_tuning(G2048_TUNE_STEP_TABLES);
    const bool swar = !(tables == 1 || (tables < 0 && n >= (1 << 18)));
    // what the kernel may take for granted about the optional arrays (kOutputs of env_step_fused_kernel)
    const bool core = a.score && a.highest && a.spawn_ctr && a.reward && a.reward32 && a.score_delta && a.valid && a.legal &&
                      a.done && !a.inject;
    const bool all = core && a.episodes && a.obs && a.stepped && a.final_score && a.final_highest;
    const int known = all ? 2 : core ? 1 : 0, cap = step_tuning(G2048_TUNE_STEP_OUTPUTS);
    const int outputs = cap >= 0 && cap < known ? cap : known;
    auto go = [&](auto kernel) { return launch_pdl(pdl, kernel, grid, threads, s, a); };
    if (swar) e = outputs == 2 ? go(env_step_fused_kernel<true, 2>) : outputs == 1 ? go(env_step_fused_kernel<true, 1>)
                                                                                  : go(env_step_fused_kernel<true, 0>);
    else      e = outputs == 2 ? go(env_step_fused_kernel<false, 2>) : outputs == 1 ? go(env_step_fused_kernel<false, 1>)
                                                                                   : go(env_step_fused_kernel<false, 0>);
    count_launch();
    return check_cuda(e, __func__);
}

int g2048_env_step(uint64_t *boards, const uint8_t *actions, const uint32_t *spawn_inject,
                   int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                   double *reward, float *reward32, int32_t *score_delta,
                   uint8_t *valid, uint8_t *legal, uint8_t *done,
                   int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    StepArgs a{boards, actions, spawn_inject, score, highest_exp, spawn_ctr, reward, reward32, score_delta,
               valid, legal, done, n, make_philox_key(seed), game0, nullptr, nullptr, nullptr, nullptr, nullptr,
               nullptr, nullptr, nullptr, nullptr};
    return env_step_launch(a, stream);
}

int g2048_env_step_autoreset(uint64_t *boards, const uint8_t *actions,
                             int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                             double *reward, float *reward32, int32_t *score_delta,
                             uint8_t *valid, uint8_t *legal, uint8_t *done, int32_t *episodes,
                             int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    if (!episodes || !spawn_ctr) return set_error(G2048_EINVAL, "g2048_env_step_autoreset: episodes and spawn_ctr are required");
    StepArgs a{boards, actions, nullptr, score, highest_exp, spawn_ctr, reward, reward32, score_delta,
               valid, legal, done, n, make_philox_key(seed), game0, nullptr, nullptr, nullptr, nullptr, episodes,
               nullptr, nullptr, nullptr, nullptr};
    return env_step_launch(a, stream);
}

int g2048_env_step_fused(uint64_t *boards, const uint8_t *actions,
                         int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                         double *reward, float *reward32, int32_t *score_delta,
                         uint8_t *valid, uint8_t *legal, uint8_t *done, int32_t *episodes,
                         float *obs, uint64_t *stepped_boards, int32_t *final_score, uint8_t *final_highest_exp,
                         int64_t n, uint64_t seed, uint32_t game0, void *stream)
{
    if (episodes && !spawn_ctr) return set_error(G2048_EINVAL, "g2048_env_step_fused: auto-reset (episodes) needs spawn_ctr");
    StepArgs a{boards, actions, nullptr, score, highest_exp, spawn_ctr, reward, reward32, score_delta,
               valid, legal, done, n, make_philox_key(seed), game0, nullptr, nullptr, nullptr, nullptr, episodes,
               obs, stepped_boards, final_score, final_highest_exp};
    return env_step_launch(a, stream);
}

int g2048_legal_masks(const uint64_t *boards, uint8_t *env_legal, uint8_t *agent_legal, int64_t n, void *stream)
{
    G2048_ENTER(boards);
    legal_masks_kernel<<<grid_for(n, 256, st->sm_count, 8), 256, 0, s>>>(boards, env_legal, agent_legal, n, st->row);
    G2048_LAUNCHED();
}

int g2048_env_rollout(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                      double *reward_sum, int32_t *episodes,
                      int64_t n, int32_t steps, uint32_t t0, uint64_t seed, uint32_t game0, void *stream)
{
    G2048_ENTER(boards && score && highest_exp && spawn_ctr && steps >= 0);
    RolloutArgs a{boards, score, highest_exp, spawn_ctr, reward_sum, episodes, n, steps, t0,
                  make_philox_key(seed), game0, st->row, st->code, st->overflow};
    // one 192 KiB block per SM; spread the envs over ALL SMs (65,536 envs -> 148 blocks x 448 threads)
    int grid = (int)(n < st->sm_count ? n : st->sm_count);
    int64_t per_block = (n + grid - 1) / grid;
    int threads = (int)(per_block >= kRolloutThreads ? kRolloutThreads : ((per_block + 31) / 32) * 32);
    env_rollout_kernel<<<grid, threads, kRowTableBytes + kCodeTableBytes, s>>>(a);
    G2048_LAUNCHED();
}

int g2048_evaluate(const uint64_t *boards, int32_t *fast, double *full, int64_t n, void *stream)
{
    G2048_ENTER(boards);
    evaluate_kernel<<<grid_for(n, 256, st->sm_count, 8), 256, 0, s>>>(boards, fast, full, n);
    G2048_LAUNCHED();
}

int g2048_beam_search(const uint64_t *roots, const uint8_t *legal, const uint32_t *call, uint32_t call0,
                      uint8_t *action, float *prob, double *best_score, int32_t *nodes,
                      int64_t n, int32_t beam_width, int32_t search_depth,
                      int32_t early_thr, int32_t mid_thr, uint64_t seed, uint32_t game0, void *stream)
{
    G2048_ENTER(roots && action && beam_width >= 1 && beam_width <= G2048_MAX_WIDE_BEAM_WIDTH && search_depth >= 1);
    return launch_beam_search(st, roots, legal, call, call0, action, prob, best_score, nodes, n, beam_width,
                              search_depth, early_thr, mid_thr, seed, game0, s);
}

int g2048_play_games(int64_t n, int32_t beam_width, int32_t search_depth,
                     int32_t early_thr, int32_t mid_thr, int32_t max_moves, uint64_t seed, uint32_t game0,
                     int32_t *score, uint8_t *highest_exp, int32_t *moves, int32_t *valid, int32_t *invalid,
                     int32_t *milestone, int64_t *nodes, uint64_t *final_board, void *stream)
{
    G2048_ENTER(beam_width >= 1 && beam_width <= G2048_MAX_WIDE_BEAM_WIDTH && search_depth >= 1 && max_moves >= 0);
    return launch_play_games(st, n, beam_width, search_depth, early_thr, mid_thr, max_moves, seed, game0, score,
                             highest_exp, moves, valid, invalid, milestone, nodes, final_board, s);
}

int g2048_stats_reduce(const int32_t *score, const uint8_t *highest_exp, const int32_t *moves,
                       const int32_t *valid, const int32_t *invalid, const int32_t *milestone,
                       const int64_t *nodes, int64_t n, int64_t *stats, void *stream)
{
    G2048_ENTER(stats);
    stats_kernel<<<grid_for(n, 256, st->sm_count, 2), 256, 0, s>>>(score, highest_exp, moves, valid, invalid, milestone,
                                                                  nodes, n, reinterpret_cast<unsigned long long *>(stats));
    G2048_LAUNCHED();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// host-buffer entry points: H2D, kernels, D2H on the library's own stream
// ------------------------------------------------------------------------------------------
namespace g2048 {

// A grow-only device arena per device; the host_* calls carve their temporaries out of it.
struct Arena {
    uint8_t *base = nullptr;
    size_t cap = 0, used = 0;
    cudaStream_t stream = nullptr;
};
static Arena g_arena[kMaxDevices];
// one staging arena, stream and lock per DEVICE: host calls for different GPUs do not serialise behind each other
static std::mutex g_host_mutex[kMaxDevices];
static std::mutex &host_mutex()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return g_host_mutex[dev];
}

static int arena_begin(Arena **out, size_t need)
{
    int dev = 0;
    G2048_CUDA(cudaGetDevice(&dev));
    Arena &a = g_arena[dev];
    if (!a.stream) G2048_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
    if (need > a.cap) {
        if (a.base) { G2048_CUDA(cudaStreamSynchronize(a.stream)); G2048_CUDA(cudaFree(a.base)); a.base = nullptr; }
        size_t cap = need + need / 4 + (1u << 16);
        G2048_CUDA(cudaMalloc(&a.base, cap));
        a.cap = cap;
    }
    a.used = 0;
    *out = &a;
    return G2048_OK;
}
template <typename T>
static T *arena_take(Arena *a, int64_t count)
{
    size_t bytes = ((size_t)count * sizeof(T) + 255u) & ~(size_t)255u;
    T *p = reinterpret_cast<T *>(a->base + a->used);
    a->used += bytes;
    return p;
}
template <typename T>
static size_t arena_bytes(int64_t count) { return ((size_t)count * sizeof(T) + 255u) & ~(size_t)255u; }

template <typename T>
static int to_device(Arena *a, T **dptr, const T *host, int64_t count, bool copy)
{
    *dptr = arena_take<T>(a, count);
    if (copy && host) G2048_CUDA(cudaMemcpyAsync(*dptr, host, (size_t)count * sizeof(T), cudaMemcpyHostToDevice, a->stream));
    return G2048_OK;
}
template <typename T>
static int to_host(Arena *a, T *host, const T *dptr, int64_t count)
{
    if (host) G2048_CUDA(cudaMemcpyAsync(host, dptr, (size_t)count * sizeof(T), cudaMemcpyDeviceToHost, a->stream));
    return G2048_OK;
}

}  // namespace g2048

#define G2048_TRY(expr)                        \
    do {                                       \
        int _rc = (expr);                      \
        if (_rc != G2048_OK) return _rc;       \
    } while (0)

extern "C" {

int g2048_host_env_reset(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                         int64_t n, uint64_t seed, uint32_t game0)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards) return set_error(G2048_EINVAL, "g2048_host_env_reset: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + arena_bytes<int32_t>(n) + arena_bytes<uint8_t>(n) + arena_bytes<uint32_t>(n)));
    uint64_t *d_b; int32_t *d_s; uint8_t *d_h; uint32_t *d_c;
    G2048_TRY(to_device(a, &d_b, boards, n, false));
    G2048_TRY(to_device(a, &d_s, score, n, false));
    G2048_TRY(to_device(a, &d_h, highest_exp, n, false));
    G2048_TRY(to_device(a, &d_c, spawn_ctr, n, true));
    if (!spawn_ctr) G2048_CUDA(cudaMemsetAsync(d_c, 0, (size_t)n * sizeof(uint32_t), a->stream));
    G2048_TRY(g2048_env_reset(d_b, d_s, d_h, d_c, n, seed, game0, a->stream));
    G2048_TRY(to_host(a, boards, d_b, n));
    G2048_TRY(to_host(a, score, d_s, n));
    G2048_TRY(to_host(a, highest_exp, d_h, n));
    G2048_TRY(to_host(a, spawn_ctr, d_c, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_env_step(uint64_t *boards, const uint8_t *actions, const uint32_t *spawn_inject,
                        int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                        double *reward, int32_t *score_delta, uint8_t *valid, uint8_t *legal, uint8_t *done,
                        int64_t n, uint64_t seed, uint32_t game0)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards || !actions) return set_error(G2048_EINVAL, "g2048_host_env_step: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + arena_bytes<uint32_t>(2 * n) + 2 * arena_bytes<int32_t>(n) +
                                  arena_bytes<uint32_t>(n) + arena_bytes<double>(n) + 5 * arena_bytes<uint8_t>(n)));
    uint64_t *d_b; uint8_t *d_a; uint32_t *d_inj = nullptr; int32_t *d_s; uint8_t *d_h; uint32_t *d_c;
    double *d_r = nullptr; int32_t *d_sd = nullptr; uint8_t *d_v = nullptr, *d_l = nullptr, *d_d = nullptr;
    G2048_TRY(to_device(a, &d_b, (const uint64_t *)boards, n, true));
    G2048_TRY(to_device(a, &d_a, actions, n, true));
    if (spawn_inject) G2048_TRY(to_device(a, &d_inj, spawn_inject, 2 * n, true));
    G2048_TRY(to_device(a, &d_s, (const int32_t *)score, n, true));
    if (!score) G2048_CUDA(cudaMemsetAsync(d_s, 0, (size_t)n * sizeof(int32_t), a->stream));
    G2048_TRY(to_device(a, &d_h, (const uint8_t *)highest_exp, n, true));
    if (!highest_exp) G2048_CUDA(cudaMemsetAsync(d_h, 0, (size_t)n, a->stream));
    G2048_TRY(to_device(a, &d_c, (const uint32_t *)spawn_ctr, n, true));
    if (!spawn_ctr) G2048_CUDA(cudaMemsetAsync(d_c, 0, (size_t)n * sizeof(uint32_t), a->stream));
    if (reward) d_r = arena_take<double>(a, n);
    if (score_delta) d_sd = arena_take<int32_t>(a, n);
    if (valid) d_v = arena_take<uint8_t>(a, n);
    if (legal) d_l = arena_take<uint8_t>(a, n);
    if (done) d_d = arena_take<uint8_t>(a, n);
    G2048_TRY(g2048_env_step(d_b, d_a, d_inj, d_s, d_h, d_c, d_r, nullptr, d_sd, d_v, d_l, d_d, n, seed, game0, a->stream));
    G2048_TRY(to_host(a, boards, d_b, n));
    G2048_TRY(to_host(a, score, d_s, n));
    G2048_TRY(to_host(a, highest_exp, d_h, n));
    G2048_TRY(to_host(a, spawn_ctr, d_c, n));
    G2048_TRY(to_host(a, reward, d_r, n));
    G2048_TRY(to_host(a, score_delta, d_sd, n));
    G2048_TRY(to_host(a, valid, d_v, n));
    G2048_TRY(to_host(a, legal, d_l, n));
    G2048_TRY(to_host(a, done, d_d, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_env_rollout(uint64_t *boards, int32_t *score, uint8_t *highest_exp, uint32_t *spawn_ctr,
                           double *reward_sum, int32_t *episodes,
                           int64_t n, int32_t steps, uint32_t t0, uint64_t seed, uint32_t game0)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards || !score || !highest_exp || !spawn_ctr)
        return set_error(G2048_EINVAL, "g2048_host_env_rollout: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + 2 * arena_bytes<int32_t>(n) + arena_bytes<uint8_t>(n) +
                                  arena_bytes<uint32_t>(n) + arena_bytes<double>(n)));
    uint64_t *d_b; int32_t *d_s; uint8_t *d_h; uint32_t *d_c; double *d_r; int32_t *d_e;
    G2048_TRY(to_device(a, &d_b, (const uint64_t *)boards, n, true));
    G2048_TRY(to_device(a, &d_s, (const int32_t *)score, n, true));
    G2048_TRY(to_device(a, &d_h, (const uint8_t *)highest_exp, n, true));
    G2048_TRY(to_device(a, &d_c, (const uint32_t *)spawn_ctr, n, true));
    G2048_TRY(to_device(a, &d_r, (const double *)reward_sum, n, true));
    if (!reward_sum) G2048_CUDA(cudaMemsetAsync(d_r, 0, (size_t)n * sizeof(double), a->stream));
    G2048_TRY(to_device(a, &d_e, (const int32_t *)episodes, n, true));
    if (!episodes) G2048_CUDA(cudaMemsetAsync(d_e, 0, (size_t)n * sizeof(int32_t), a->stream));
    G2048_TRY(g2048_env_rollout(d_b, d_s, d_h, d_c, d_r, d_e, n, steps, t0, seed, game0, a->stream));
    G2048_TRY(to_host(a, boards, d_b, n));
    G2048_TRY(to_host(a, score, d_s, n));
    G2048_TRY(to_host(a, highest_exp, d_h, n));
    G2048_TRY(to_host(a, spawn_ctr, d_c, n));
    G2048_TRY(to_host(a, reward_sum, d_r, n));
    G2048_TRY(to_host(a, episodes, d_e, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_legal_masks(const uint64_t *boards, uint8_t *env_legal, uint8_t *agent_legal, int64_t n)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards) return set_error(G2048_EINVAL, "g2048_host_legal_masks: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + 2 * arena_bytes<uint8_t>(n)));
    uint64_t *d_b; uint8_t *d_e = nullptr, *d_g = nullptr;
    G2048_TRY(to_device(a, &d_b, boards, n, true));
    if (env_legal) d_e = arena_take<uint8_t>(a, n);
    if (agent_legal) d_g = arena_take<uint8_t>(a, n);
    G2048_TRY(g2048_legal_masks(d_b, d_e, d_g, n, a->stream));
    G2048_TRY(to_host(a, env_legal, d_e, n));
    G2048_TRY(to_host(a, agent_legal, d_g, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_ppo_features(const uint64_t *boards, float *obs, double *heuristic, double *top4_bonus, int64_t n)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards) return set_error(G2048_EINVAL, "g2048_host_ppo_features: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + arena_bytes<float>(16 * n) + 2 * arena_bytes<double>(n)));
    uint64_t *d_b; float *d_o = nullptr; double *d_h = nullptr, *d_t = nullptr;
    G2048_TRY(to_device(a, &d_b, boards, n, true));
    if (obs) d_o = arena_take<float>(a, 16 * n);
    if (heuristic) d_h = arena_take<double>(a, n);
    if (top4_bonus) d_t = arena_take<double>(a, n);
    G2048_TRY(g2048_ppo_features(d_b, d_o, d_h, d_t, n, a->stream));
    G2048_TRY(to_host(a, obs, d_o, 16 * n));
    G2048_TRY(to_host(a, heuristic, d_h, n));
    G2048_TRY(to_host(a, top4_bonus, d_t, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_simulate_move(const uint64_t *boards, const uint8_t *actions, const uint8_t *highest_exp,
                             uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count, double *pattern, int64_t n)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards || !actions) return set_error(G2048_EINVAL, "g2048_host_simulate_move: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + 2 * arena_bytes<uint8_t>(n) + arena_bytes<uint64_t>(32 * n) +
                                  arena_bytes<double>(32 * n) + arena_bytes<uint8_t>(32 * n) + arena_bytes<int32_t>(n) +
                                  arena_bytes<double>(n)));
    uint64_t *d_b; uint8_t *d_a; uint8_t *d_h = nullptr;
    G2048_TRY(to_device(a, &d_b, boards, n, true));
    G2048_TRY(to_device(a, &d_a, actions, n, true));
    if (highest_exp) G2048_TRY(to_device(a, &d_h, highest_exp, n, true));
    uint64_t *d_nb = arena_take<uint64_t>(a, 32 * n);
    double *d_r = arena_take<double>(a, 32 * n);
    uint8_t *d_d = arena_take<uint8_t>(a, 32 * n);
    int32_t *d_c = arena_take<int32_t>(a, n);
    double *d_p = arena_take<double>(a, n);
    G2048_TRY(g2048_simulate_move(d_b, d_a, d_h, d_nb, d_r, d_d, d_c, n, a->stream));
    if (pattern) G2048_TRY(g2048_evaluate_pattern(d_b, d_p, n, a->stream));
    G2048_TRY(to_host(a, next_boards, d_nb, 32 * n));
    G2048_TRY(to_host(a, reward, d_r, 32 * n));
    G2048_TRY(to_host(a, done, d_d, 32 * n));
    G2048_TRY(to_host(a, count, d_c, n));
    G2048_TRY(to_host(a, pattern, d_p, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_hybrid_expand(const uint64_t *boards, const uint8_t *actions, const uint32_t *call, uint32_t call0,
                             const uint32_t *draw0, uint64_t *next_boards, double *reward, uint8_t *done, int32_t *count,
                             uint32_t *draws, int64_t n, uint64_t seed, uint32_t game0)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards || !actions) return set_error(G2048_EINVAL, "g2048_host_hybrid_expand: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + arena_bytes<uint8_t>(n) + 3 * arena_bytes<uint32_t>(n) +
                                  arena_bytes<uint64_t>(8 * n) + arena_bytes<double>(8 * n) + arena_bytes<uint8_t>(8 * n) +
                                  arena_bytes<int32_t>(n)));
    uint64_t *d_b; uint8_t *d_a; uint32_t *d_call = nullptr, *d_d0 = nullptr;
    G2048_TRY(to_device(a, &d_b, boards, n, true));
    G2048_TRY(to_device(a, &d_a, actions, n, true));
    if (call) G2048_TRY(to_device(a, &d_call, call, n, true));
    if (draw0) G2048_TRY(to_device(a, &d_d0, draw0, n, true));
    uint64_t *d_nb = arena_take<uint64_t>(a, 8 * n);
    double *d_r = arena_take<double>(a, 8 * n);
    uint8_t *d_dn = arena_take<uint8_t>(a, 8 * n);
    int32_t *d_c = arena_take<int32_t>(a, n);
    uint32_t *d_u = arena_take<uint32_t>(a, n);
    G2048_TRY(g2048_hybrid_expand(d_b, d_a, d_call, call0, d_d0, d_nb, d_r, d_dn, d_c, d_u, n, seed, game0, a->stream));
    G2048_TRY(to_host(a, next_boards, d_nb, 8 * n));
    G2048_TRY(to_host(a, reward, d_r, 8 * n));
    G2048_TRY(to_host(a, done, d_dn, 8 * n));
    G2048_TRY(to_host(a, count, d_c, n));
    G2048_TRY(to_host(a, draws, d_u, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_evaluate(const uint64_t *boards, int32_t *fast, double *full, int64_t n)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !boards) return set_error(G2048_EINVAL, "g2048_host_evaluate: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + arena_bytes<int32_t>(n) + arena_bytes<double>(3 * n)));
    uint64_t *d_b; int32_t *d_f = nullptr; double *d_u = nullptr;
    G2048_TRY(to_device(a, &d_b, boards, n, true));
    if (fast) d_f = arena_take<int32_t>(a, n);
    if (full) d_u = arena_take<double>(a, 3 * n);
    G2048_TRY(g2048_evaluate(d_b, d_f, d_u, n, a->stream));
    G2048_TRY(to_host(a, fast, d_f, n));
    G2048_TRY(to_host(a, full, d_u, 3 * n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_beam_search(const uint64_t *roots, const uint8_t *legal, const uint32_t *call, uint32_t call0,
                           uint8_t *action, float *prob, double *best_score, int32_t *nodes,
                           int64_t n, int32_t beam_width, int32_t search_depth,
                           int32_t early_thr, int32_t mid_thr, uint64_t seed, uint32_t game0)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0 || !roots || !action) return set_error(G2048_EINVAL, "g2048_host_beam_search: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, arena_bytes<uint64_t>(n) + 2 * arena_bytes<uint8_t>(n) + arena_bytes<uint32_t>(n) +
                                  arena_bytes<float>(n) + arena_bytes<double>(n) + arena_bytes<int32_t>(n)));
    uint64_t *d_b; uint8_t *d_l = nullptr; uint32_t *d_c = nullptr; uint8_t *d_a; float *d_p; double *d_s; int32_t *d_n;
    G2048_TRY(to_device(a, &d_b, roots, n, true));
    if (legal) G2048_TRY(to_device(a, &d_l, legal, n, true));
    if (call) G2048_TRY(to_device(a, &d_c, call, n, true));
    d_a = arena_take<uint8_t>(a, n);
    d_p = arena_take<float>(a, n);
    d_s = arena_take<double>(a, n);
    d_n = arena_take<int32_t>(a, n);
    G2048_TRY(g2048_beam_search(d_b, d_l, d_c, call0, d_a, d_p, d_s, d_n, n, beam_width, search_depth, early_thr, mid_thr,
                                seed, game0, a->stream));
    G2048_TRY(to_host(a, action, d_a, n));
    G2048_TRY(to_host(a, prob, d_p, n));
    G2048_TRY(to_host(a, best_score, d_s, n));
    G2048_TRY(to_host(a, nodes, d_n, n));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

int g2048_host_play_games(int64_t n, int32_t beam_width, int32_t search_depth,
                          int32_t early_thr, int32_t mid_thr, int32_t max_moves, uint64_t seed, uint32_t game0,
                          int32_t *score, uint8_t *highest_exp, int32_t *moves, int32_t *valid, int32_t *invalid,
                          int32_t *milestone, int64_t *nodes, uint64_t *final_board, int64_t *stats)
{
    if (!current_device_state()) return G2048_ENOTINIT;
    if (n < 0) return set_error(G2048_EINVAL, "g2048_host_play_games: bad argument");
    if (n == 0) return G2048_OK;
    std::lock_guard<std::mutex> lock(host_mutex());
    Arena *a;
    G2048_TRY(arena_begin(&a, 4 * arena_bytes<int32_t>(n) + arena_bytes<uint8_t>(n) + arena_bytes<int32_t>(8 * n) +
                                  arena_bytes<int64_t>(n) + arena_bytes<uint64_t>(n) + arena_bytes<int64_t>(G2048_STATS_LEN)));
    int32_t *d_s = arena_take<int32_t>(a, n);
    uint8_t *d_h = arena_take<uint8_t>(a, n);
    int32_t *d_m = arena_take<int32_t>(a, n);
    int32_t *d_v = arena_take<int32_t>(a, n);
    int32_t *d_i = arena_take<int32_t>(a, n);
    int32_t *d_ms = arena_take<int32_t>(a, 8 * n);
    int64_t *d_n = arena_take<int64_t>(a, n);
    uint64_t *d_f = arena_take<uint64_t>(a, n);
    int64_t *d_st = arena_take<int64_t>(a, G2048_STATS_LEN);
    G2048_TRY(g2048_play_games(n, beam_width, search_depth, early_thr, mid_thr, max_moves, seed, game0,
                               d_s, d_h, d_m, d_v, d_i, d_ms, d_n, d_f, a->stream));
    if (stats) {
        G2048_CUDA(cudaMemsetAsync(d_st, 0, sizeof(int64_t) * G2048_STATS_LEN, a->stream));
        G2048_TRY(g2048_stats_reduce(d_s, d_h, d_m, d_v, d_i, d_ms, d_n, n, d_st, a->stream));
    }
    G2048_TRY(to_host(a, score, d_s, n));
    G2048_TRY(to_host(a, highest_exp, d_h, n));
    G2048_TRY(to_host(a, moves, d_m, n));
    G2048_TRY(to_host(a, valid, d_v, n));
    G2048_TRY(to_host(a, invalid, d_i, n));
    G2048_TRY(to_host(a, milestone, d_ms, 8 * n));
    G2048_TRY(to_host(a, nodes, d_n, n));
    G2048_TRY(to_host(a, final_board, d_f, n));
    G2048_TRY(to_host(a, stats, d_st, G2048_STATS_LEN));
    G2048_CUDA(cudaStreamSynchronize(a->stream));
    return G2048_OK;
}

}  // extern "C"
