// beam.cu -- BeamSearchAgent.get_action (agents/beam_search_agent.py:71-181) as a warp-per-game
// search, and the whole-game driver (evaluate_beam_search.py:16-98) built on it.
//
// One warp owns one root board.  Per level:
//   A  parent-per-lane : lane p holds beam entry p in registers, makes its four children with
//                        the row table in shared memory (agent's DOWN quirk included) and
//                        writes the valid ones, compacted in (rank, action) order, to scratch.
//   B  child-per-lane  : each valid child draws its spawn (ordinal = ballot prefix over
//                        "has an empty cell", exactly the reference's sequential draw order),
//                        is evaluated and gets a sort key (score, -generation index, first action).
//   C  top-k           : warp-shuffle bitonic sort of each 32-key row + bitonic top-32 merges;
//                        lane i < k picks up rank i.  Ties keep generation order (Python's
//                        stable sort, agent:131,174).
#include "common.cuh"
#include "env.cuh"
#include "stage.cuh"

#ifndef G2048_BEAM_WARPS
#define G2048_BEAM_WARPS 24
#endif

namespace g2048 {

constexpr int kBeamWarps = G2048_BEAM_WARPS;                             // roots in flight per block
constexpr int kBeamThreads = kBeamWarps * 32;
constexpr int kMaxCand = 4 * G2048_MAX_BEAM_WIDTH;          // 128 children per level at most
constexpr uint32_t FULL = 0xFFFFFFFFu;

constexpr uint32_t kSpawnRing = 128;                        // spawn ordinals held per warp (power of two)
struct __align__(16) WarpScratch {
    uint64_t cand[kMaxCand];     // children of this level, generation order
    double score[kMaxCand];      // float64 scores of the full-evaluation levels
    uint4 rng[kSpawnRing / 2];   // spawn words of ordinals [ring_end - 128, ring_end): one Philox block per entry
    uint8_t first[kMaxCand];     // bits 0-1: first action of the child's path; bits 2-5: its largest exponent
};
constexpr size_t kBeamSmemBytes = kRowTableBytes + kBeamWarps * sizeof(WarpScratch);

struct BeamParams {
    int width, depth;
    int early_thr, mid_thr;
    PhiloxKey K;
};
struct BeamResult {
    uint32_t action;
    float prob;
    double best;
    int nodes;
};

// The block's corner table (board.cuh: corner_table_entry); stage_row_table() fills it.
__device__ __forceinline__ uint32_t *corner_table()
{
    __shared__ uint32_t table[256];
    return table;
}

// ---- warp-wide sorting network on unique uint32 keys, descending (lane 0 = largest) --------
// Bitonic sort in its "flip" form: every compare-exchange keeps the larger key in the lower
// lane, so the direction of a stage is one lane-id bit test (five loop-invariant predicates)
// instead of a per-stage direction computation.
__device__ __forceinline__ uint32_t exchange(uint32_t v, int xor_mask, bool lower)
{
    uint32_t o = __shfl_xor_sync(FULL, v, xor_mask);
    uint32_t hi = max(v, o), lo = min(v, o);
    return lower ? hi : lo;
}
__device__ __forceinline__ uint32_t sort_desc32(uint32_t v, uint32_t lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        v = exchange(v, k - 1, (lane & (k >> 1)) == 0);       // mirror inside the block of k
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) v = exchange(v, j, (lane & j) == 0);
    }
    return v;
}
// a, b sorted descending -> the 32 largest of their union, sorted descending
__device__ __forceinline__ uint32_t merge_top32(uint32_t a, uint32_t b, uint32_t lane)
{
    uint32_t v = max(a, __shfl_sync(FULL, b, 31 - lane));   // bitonic
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) v = exchange(v, j, (lane & j) == 0);
    return v;
}

// BeamSearchAgent._make_move (agent:194-258) for one direction.  DOWN returns the true DOWN
// result rotated by 180 degrees (agent:251-253 undoes agent:210 in the wrong order).
__device__ __forceinline__ Board agent_child_any(Board b, uint32_t action, const uint16_t *row)
{
    Board r = from_line(move_left<true>(to_line(b, action), row), action);
    return select(action == 3u, rot180(r), r);
}
__device__ __forceinline__ void agent_children(Board b, const uint16_t *row, Board c[4])
{
    Board t = transpose(b);
    c[0] = move_left<true>(b, row);
    c[2] = flip_rows(move_left<true>(flip_rows(b), row));
    c[1] = transpose(move_left<true>(t, row));
    // rot180(T(flip(L(flip(T b))))) == T(flipud(L(flip(T b))))
    c[3] = transpose(flip_row_order(move_left<true>(flip_rows(t), row)));
}

__device__ __forceinline__ int tile_value(uint32_t e) { return e ? (1 << e) : 0; }

// Phase B for candidate c of this level: spawn (agent:155), evaluate (agent:158-161), sort key.
// Branch-free on purpose (the Philox block is computed for every lane and masked by `draws`) so that
// two rounds issued back to back form one basic block the scheduler can interleave: a warp
// working alone at the tail of a whole-game run is bound by dependent-issue latency.
struct ChildEval {
    bool active;
    int fast;          // _fast_evaluate (valid when !kFull)
    double full;       // _evaluate_state (valid when kFull; also stored to ws.score[c])
    uint32_t first;    // first action of the child's path
};

// The beam stream hands out spawn words in ordinal order, two ordinals per Philox block.  One pass
// of this loop has lane j compute block ring_end / 2 + j, i.e. 64 ordinals for the instructions a
// per-batch draw would spend on 32 (or on far fewer: the last batch of a level is mostly idle
// lanes), and what a level leaves over serves the next one.  Callers ask for at most 64 ordinals
// past spawn_base at a time, so everything they still need stays inside the 128-ordinal ring.
__device__ __forceinline__ void fill_spawn_ring(WarpScratch &ws, uint32_t &ring_end, uint32_t need_end,
                                                const BeamParams &P, uint32_t game, uint32_t call, uint32_t lane)
{
    if (ring_end >= need_end) return;                           // warp-uniform
    __syncwarp();
    do {
        const Philox4 p = philox4x32_10((ring_end >> 1) + lane, call, game, DOM_BEAM, P.K);
        ws.rng[((ring_end & (kSpawnRing - 1u)) >> 1) + lane] = make_uint4(p.w[0], p.w[1], p.w[2], p.w[3]);
        ring_end += 64u;
    } while (ring_end < need_end);
    __syncwarp();
}
template <typename Scratch> struct HasSpawnRing { static constexpr bool value = false; };
template <> struct HasSpawnRing<WarpScratch> { static constexpr bool value = true; };
template <typename Scratch>
__device__ __forceinline__ SpawnWords ring_words(const Scratch &, uint32_t) { return SpawnWords{0u, 0u}; }
template <>
__device__ __forceinline__ SpawnWords ring_words<WarpScratch>(const WarpScratch &ws, uint32_t ordinal)
{
    const uint2 w = reinterpret_cast<const uint2 *>(ws.rng)[ordinal & (kSpawnRing - 1u)];
    return SpawnWords{w.x, w.y};
}

template <bool kFull, typename Scratch>
__device__ __forceinline__ ChildEval spawn_and_eval(int c, int n_valid, uint32_t lane_lt, uint32_t &spawn_base,
                                                    const BeamParams &P, uint32_t game, uint32_t call, int phase,
                                                    Scratch &ws)
{
    ChildEval ev;
    ev.active = c < n_valid;
    Board b = ev.active ? Board(ws.cand[c]) : Board(0u, 0u);
    const uint32_t fe = ev.active ? ws.first[c] : 0u;
    uint32_t nzl = nz_flags(b.lo), nzh = nz_flags(b.hi);
    const uint32_t zl = nzl ^ LSB4, zh = nzh ^ LSB4;
    const int cl = __popc(zl);
    int n_empty = cl + __popc(zh);
    const bool draws = ev.active && n_empty > 0;                // agent:262-263: no draw on a full board
    const uint32_t bal = __ballot_sync(FULL, draws);
    const uint32_t ordinal = spawn_base + (uint32_t)__popc(bal & lane_lt);
    G2048_ASSERT(!ev.active || c < (int)(sizeof(ws.cand) / sizeof(ws.cand[0])));
    // fast path: the caller has filled the ring (fill_spawn_ring); wide path: one block per child
    const SpawnWords w = HasSpawnRing<Scratch>::value ? ring_words(ws, ordinal)
                                                      : spawn_words(P.K, game, call, DOM_BEAM, ordinal);
    spawn_base += (uint32_t)__popc(bal);
    const SpawnPick sp = pick_spawn(zl, zh, cl, n_empty, w.pos, w.val);
    if (draws) {
        b.lo |= sp.flag_lo * sp.exponent;
        b.hi |= sp.flag_hi * sp.exponent;
        nzl |= sp.flag_lo;
        nzh |= sp.flag_hi;
        n_empty -= 1;
    }
    const uint32_t pmax = fe >> 2;                              // parent's largest exponent
    const uint32_t emax = pmax + ((pmax < 15u && has_exponent(b, pmax + 1u)) ? 1u : 0u);
    ev.first = fe & 3u;
    ev.fast = 0;
    ev.full = 0.0;
    if (kFull) ev.full = full_eval_flags(b, nzl, nzh, n_empty, emax, phase, corner_table());
    else       ev.fast = fast_eval_flags(b, nzl, nzh, n_empty, emax, corner_table());
    if (ev.active) {
        ws.cand[c] = b.u64();
        ws.first[c] = (uint8_t)(ev.first | (emax << 2));
        if (kFull) ws.score[c] = ev.full;
    }
    return ev;
}

// fast path (width <= 32): unique 32-bit sort key (score, 127 - generation index, first action)
template <bool kFull>
__device__ __forceinline__ uint32_t spawn_and_score(int c, int n_valid, uint32_t lane_lt, uint32_t &spawn_base,
                                                    const BeamParams &P, uint32_t game, uint32_t call, int phase,
                                                    WarpScratch &ws)
{
    const ChildEval ev = spawn_and_eval<kFull>(c, n_valid, lane_lt, spawn_base, P, game, call, phase, ws);
    const uint32_t tail = ((uint32_t)(127 - c) << 2) | ev.first;
    const uint32_t key = kFull ? tail : (((uint32_t)ev.fast << 9) | tail);   // kFull: rank field filled in by the caller
    return ev.active ? key : 0u;
}

// kRows independent 32-key sorts in lockstep: same network, kRows exchanges per stage that do not
// depend on one another, so a lone warp hides the shuffle latency of one behind the others.
template <int kRows>
__device__ __forceinline__ void sort_desc32_rows(uint32_t (&k)[4], uint32_t lane)
{
#pragma unroll
    for (int blk = 2; blk <= 32; blk <<= 1) {
        bool lower = (lane & (blk >> 1)) == 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) k[r] = exchange(k[r], blk - 1, lower);
#pragma unroll
        for (int j = blk >> 2; j > 0; j >>= 1) {
            lower = (lane & j) == 0;
#pragma unroll
            for (int r = 0; r < kRows; ++r) k[r] = exchange(k[r], j, lower);
        }
    }
}
// top-32 of (a0 U b0) and of (a1 U b1) in lockstep
__device__ __forceinline__ void merge_top32_x2(uint32_t &a0, uint32_t b0, uint32_t &a1, uint32_t b1, uint32_t lane)
{
    a0 = max(a0, __shfl_sync(FULL, b0, 31 - lane));
    a1 = max(a1, __shfl_sync(FULL, b1, 31 - lane));
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const bool lower = (lane & j) == 0;
        a0 = exchange(a0, j, lower);
        a1 = exchange(a1, j, lower);
    }
}

// Full-evaluation levels: rank = #candidates with a strictly larger float64 score, written into the
// score field of the keys of the kRows rows in use (candidate c sits in row c / 32, lane c % 32).
template <int kRows>
__device__ __forceinline__ void rank_by_count(const WarpScratch &ws, int n_valid, uint32_t lane, uint32_t (&key)[4])
{
    int rank[kRows];
    double mine_s[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int c = r * 32 + (int)lane;
        rank[r] = 0;
        mine_s[r] = c < n_valid ? ws.score[c] : 0.0;
    }
    for (int j = 0; j < n_valid; ++j) {
        const double s = ws.score[j];
#pragma unroll
        for (int r = 0; r < kRows; ++r) rank[r] += (s > mine_s[r]) ? 1 : 0;
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int c = r * 32 + (int)lane;
        if (c < n_valid) key[r] |= (uint32_t)(n_valid - rank[r]) << 9;
    }
}

// All 32 lanes call this with the same root / parameters.
__device__ __forceinline__ BeamResult beam_search_warp(Board root, int legal_given, const BeamParams &P, uint32_t game,
                                       uint32_t call, const uint16_t *row, WarpScratch &ws)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    BeamResult res;
    res.nodes = 0;
    res.best = 0.0;

    // agent:82-93 -- legality and the two fast exits
    Board root_child = agent_child_any(root, lane & 3u, row);
    bool root_child_valid = root_child != root;
    uint32_t agent_mask = __ballot_sync(FULL, root_child_valid) & 15u;
    uint32_t vm = legal_given >= 0 ? ((uint32_t)legal_given & 15u) : agent_mask;
    if (vm == 0u) { res.action = 0u; res.prob = 0.5f; return res; }
    if ((vm & (vm - 1u)) == 0u) { res.action = (uint32_t)__ffs((int)vm) - 1u; res.prob = 1.0f; return res; }

    // agent:96-106 -- phase from the root's max tile, adaptive depth from its empty count
    const int root_max = tile_value(max_exponent(root));
    const int phase = root_max < P.early_thr ? 0 : root_max < P.mid_thr ? 1 : 2;
    const int n0 = count_empty(root);
    // the root expansion (agent:112-132) always runs; `for depth in range(1, actual_depth)` adds levels
    const int depth = max(1, n0 <= 4 ? min(P.depth + 5, 25) : n0 >= 10 ? min(P.depth - 5, 10) : P.depth);

    Board mine(0u, 0u);          // beam entry of rank `lane`
    uint32_t my_first = 0u;      // first action of its path
    uint32_t my_emax = 0u;       // its largest exponent (a child's is the parent's or one more)
    const uint32_t root_emax = max_exponent(root);
    int nb = 0;                  // beam entries alive
    uint32_t spawn_base = 0u;    // spawns drawn so far in this call
    uint32_t ring_end = 0u;      // ws.rng holds the spawn words of ordinals [ring_end - 128, ring_end)

    for (int d = 0; d < depth; ++d) {
        // ---- A: expand ----------------------------------------------------------------------
        int n_valid;
        if (d == 0) {                                                   // agent:112-123
            bool is_cand = lane < 4u && ((vm >> lane) & 1u) && root_child_valid;
            uint32_t bal = __ballot_sync(FULL, is_cand);
            if (is_cand) {
                int pos = __popc(bal & lt_mask);
                ws.cand[pos] = root_child.u64();
                ws.first[pos] = (uint8_t)(lane | (root_emax << 2));
            }
            n_valid = __popc(bal);
        } else {                                                        // agent:142-167
            Board c[4];
            uint32_t v = 0u;
            if ((int)lane < nb) {
                agent_children(mine, row, c);
#pragma unroll
                for (int a = 0; a < 4; ++a) v |= (c[a] != mine) ? (1u << a) : 0u;
            }
            int incl = __popc(v);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(FULL, incl, o);
                if ((int)lane >= o) incl += t;
            }
            n_valid = __shfl_sync(FULL, incl, 31);
            int pos = incl - __popc(v);
            G2048_ASSERT(incl <= kMaxCand);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if ((v >> a) & 1u) {
                    ws.cand[pos] = c[a].u64();
                    ws.first[pos] = (uint8_t)(my_first | (my_emax << 2));
                    ++pos;
                }
            }
        }
        __syncwarp();
        if (n_valid == 0) {
            if (d > 0) break;                                           // agent:170-171 keeps the old beam
            // agent:126-128: random.choice among the caller's valid moves, one draw
            Philox4 p = philox4x32_10(0u, call, game, DOM_BEAM, P.K);
            int pick = (int)__umulhi(p.w[0], (uint32_t)__popc(vm));
            uint32_t m = vm;
            for (int i = 0; i < pick; ++i) m &= m - 1u;
            res.action = (uint32_t)__ffs((int)m) - 1u;
            res.prob = 0.5f;
            return res;
        }
        res.nodes += n_valid;

        // ---- B: spawn + evaluate ---------------------------------------------------------------
        const bool full_level = d >= 1 && d <= 3;                       // agent:139,158-161
        uint32_t key[4] = {0u, 0u, 0u, 0u};
        const int L = (int)lane;
        fill_spawn_ring(ws, ring_end, spawn_base + (uint32_t)min(n_valid, 64), P, game, call, lane);
        if (full_level) {
            key[0] = spawn_and_score<true>(L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
            if (n_valid > 32) key[1] = spawn_and_score<true>(32 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
            if (n_valid > 64) {
                fill_spawn_ring(ws, ring_end, spawn_base + (uint32_t)(n_valid - 64), P, game, call, lane);
                key[2] = spawn_and_score<true>(64 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
            }
            if (n_valid > 96) key[3] = spawn_and_score<true>(96 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
        } else if (n_valid > 32) {                                      // the common case: two rounds, one basic block
            key[0] = spawn_and_score<false>(L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
            key[1] = spawn_and_score<false>(32 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
            if (n_valid > 64) {
                fill_spawn_ring(ws, ring_end, spawn_base + (uint32_t)(n_valid - 64), P, game, call, lane);
                key[2] = spawn_and_score<false>(64 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
            }
            if (n_valid > 96) key[3] = spawn_and_score<false>(96 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
        } else {
            key[0] = spawn_and_score<false>(L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
        }
        __syncwarp();
        if (full_level) {
            // float64 scores: rank = #candidates with a strictly larger score; equal scores are
            // separated by the generation index already in the key (stable sort).
            const int rows = (n_valid + 31) >> 5;
            if (rows <= 2)      rank_by_count<2>(ws, n_valid, lane, key);
            else if (rows == 3) rank_by_count<3>(ws, n_valid, lane, key);
            else                rank_by_count<4>(ws, n_valid, lane, key);
        }

        // ---- C: stable top-k ------------------------------------------------------------------------
        nb = min(P.width, n_valid);                                     // agent:132,175
        uint32_t top = 0u;
        if (n_valid > 96) {
            sort_desc32_rows<4>(key, lane);
            merge_top32_x2(key[0], key[1], key[2], key[3], lane);
            top = merge_top32(key[0], key[2], lane);
        } else if (n_valid > 64) {
            sort_desc32_rows<3>(key, lane);
            top = merge_top32(merge_top32(key[0], key[1], lane), key[2], lane);
        } else if (n_valid > 32) {
            sort_desc32_rows<2>(key, lane);
            top = merge_top32(key[0], key[1], lane);
        } else {
            top = sort_desc32(key[0], lane);
        }
        const int pick = 127 - (int)((top >> 2) & 127u);
        G2048_ASSERT((int)lane >= nb || (pick >= 0 && pick < n_valid && n_valid <= kMaxCand));
        G2048_ASSERT(spawn_base <= ring_end && ring_end - spawn_base <= kSpawnRing);     // every word read was inside the ring
        if ((int)lane < nb) {
            mine = Board(ws.cand[pick]);
            my_first = top & 3u;
            my_emax = ws.first[pick] >> 2;
        }
        const int pick0 = __shfl_sync(FULL, pick, 0);
        const uint32_t top0 = __shfl_sync(FULL, top, 0);
        res.best = full_level ? ws.score[pick0] : (double)(top0 >> 9);
        __syncwarp();                                                   // scratch is rewritten next level
    }
    res.action = __shfl_sync(FULL, my_first, 0);                        // agent:178
    res.prob = 1.0f;
    return res;
}

__device__ __forceinline__ void stage_row_table(uint8_t *smem, const uint16_t *row)
{
    uint32_t *corners = corner_table();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) corners[i] = corner_table_entry(i);   // published by the barrier in stage_bulk
    stage_bulk(smem, row, (uint32_t)kRowTableBytes, nullptr, nullptr, 0u);       // TMA bulk copy, stage.cuh
}

struct BeamArgs {
    const uint64_t *roots; const uint8_t *legal; const uint32_t *call; uint32_t call0;
    uint8_t *action; float *prob; double *best; int32_t *nodes;
    int64_t n; BeamParams P; uint32_t game0; const uint16_t *row; unsigned int *work;
};

__global__ void __launch_bounds__(kBeamThreads, 1) beam_search_kernel(BeamArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    stage_row_table(smem, a.row);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(smem);
    const int warp = threadIdx.x >> 5;
    WarpScratch &ws = reinterpret_cast<WarpScratch *>(smem + kRowTableBytes)[warp];
    const uint32_t lane = threadIdx.x & 31u;
    // roots differ a lot in cost (fast exits, adaptive depth 10 / d / 25): warps pull the next
    // root from a global queue instead of striding over a fixed share
    for (;;) {
        unsigned int i = 0;
        if (lane == 0) i = atomicAdd(a.work, 1u);
        i = __shfl_sync(FULL, i, 0);
        if ((int64_t)i >= a.n) break;
        Board root(a.roots[i]);
        int legal = a.legal ? (int)a.legal[i] : -1;
        uint32_t call = a.call ? a.call[i] : a.call0;
        BeamResult r = beam_search_warp(root, legal, a.P, a.game0 + (uint32_t)i, call, row, ws);
        if (lane == 0) {
            a.action[i] = (uint8_t)r.action;
            if (a.prob) a.prob[i] = r.prob;
            if (a.best) a.best[i] = r.best;
            if (a.nodes) a.nodes[i] = r.nodes;
        }
        __syncwarp();
    }
}

}  // namespace g2048
#include "team.cuh"
namespace g2048 {

// Scratch of team `quad` of the block: it aliases the WarpScratch slots of the team's four warps.
__device__ __forceinline__ TeamScratch &team_scratch(uint8_t *smem, int quad)
{
    return *reinterpret_cast<TeamScratch *>(smem + kRowTableBytes + (size_t)quad * kTeamWarps * sizeof(WarpScratch));
}

// Small batches (fewer roots than the GPU has warp slots for): one TEAM of four warps per root, so a
// single get_action takes ~1/4 of the one-warp latency.  blockDim.x is a multiple of 128.
__global__ void __launch_bounds__(kBeamThreads, 1) beam_search_team_kernel(BeamArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    stage_row_table(smem, a.row);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(smem);
    const int quad = threadIdx.x >> 7;
    TeamScratch &ts = team_scratch(smem, quad);
    const int bar = 1 + quad;
    const bool leader = (threadIdx.x & (kTeamThreads - 1)) == 0;
    for (;;) {
        if (leader) ts.next_item = atomicAdd(a.work, 1u);
        team_barrier(bar);
        const unsigned int i = ts.next_item;
        team_barrier(bar);                                 // everyone has read it before the next round rewrites it
        if ((int64_t)i >= a.n) break;
        Board root(a.roots[i]);
        int legal = a.legal ? (int)a.legal[i] : -1;
        uint32_t call = a.call ? a.call[i] : a.call0;
        BeamResult r = beam_search_team(root, legal, a.P, a.game0 + (uint32_t)i, call, row, ts, bar);
        if (leader) {
            a.action[i] = (uint8_t)r.action;
            if (a.prob) a.prob[i] = r.prob;
            if (a.best) a.best[i] = r.best;
            if (a.nodes) a.nodes[i] = r.nodes;
        }
    }
}

// ---- wide beams (33 <= beam_width <= 128) ---------------------------------------------------------
// Same algorithm with the beam and up to 512 candidates per level in shared memory and the stable
// top-k done by counting (rank = #candidates ahead; score as float64 on every level, ties by
// generation index).  O(n^2 / 32) per lane and level: a compatibility path for wide beams, not a
// tuned one -- the reference's own configurations use widths 10-20.
constexpr int kWideMaxWidth = 128;
constexpr int kWideMaxCand = 4 * kWideMaxWidth;
constexpr int kWideWarps = 8;
struct __align__(16) WideScratch {
    uint64_t beam[kWideMaxWidth];
    uint64_t cand[kWideMaxCand];
    double score[kWideMaxCand];
    uint8_t beam_first[kWideMaxWidth];     // first action | largest exponent << 2, as in WarpScratch::first
    uint8_t first[kWideMaxCand];
    uint16_t slot[kWideMaxWidth];          // slot[rank] = candidate index
};
constexpr size_t kWideSmemBytes = kRowTableBytes + kWideWarps * sizeof(WideScratch);

__device__ BeamResult beam_search_wide_warp(Board root, int legal_given, const BeamParams &P, uint32_t game,
                                            uint32_t call, const uint16_t *row, WideScratch &ws)
{
    const uint32_t lane = threadIdx.x & 31u;
    const int L = (int)lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
    BeamResult res;
    res.nodes = 0;
    res.best = 0.0;
    Board root_child = agent_child_any(root, lane & 3u, row);
    bool root_child_valid = root_child != root;
    uint32_t agent_mask = __ballot_sync(FULL, root_child_valid) & 15u;
    uint32_t vm = legal_given >= 0 ? ((uint32_t)legal_given & 15u) : agent_mask;
    if (vm == 0u) { res.action = 0u; res.prob = 0.5f; return res; }
    if ((vm & (vm - 1u)) == 0u) { res.action = (uint32_t)__ffs((int)vm) - 1u; res.prob = 1.0f; return res; }
    const uint32_t root_emax = max_exponent(root);
    const int root_max = tile_value(root_emax);
    const int phase = root_max < P.early_thr ? 0 : root_max < P.mid_thr ? 1 : 2;
    const int n0 = count_empty(root);
    const int depth = max(1, n0 <= 4 ? min(P.depth + 5, 25) : n0 >= 10 ? min(P.depth - 5, 10) : P.depth);
    int nb = 0;
    uint32_t spawn_base = 0u;
    for (int d = 0; d < depth; ++d) {
        int n_valid = 0;
        if (d == 0) {
            bool is_cand = lane < 4u && ((vm >> lane) & 1u) && root_child_valid;
            uint32_t bal = __ballot_sync(FULL, is_cand);
            if (is_cand) {
                int pos = __popc(bal & lt_mask);
                ws.cand[pos] = root_child.u64();
                ws.first[pos] = (uint8_t)(lane | (root_emax << 2));
            }
            n_valid = __popc(bal);
        } else {
            for (int p0 = 0; p0 < nb; p0 += 32) {                     // parents in rank order, 32 at a time
                const int p = p0 + L;
                Board c[4];
                uint32_t v = 0u;
                Board mine(0u, 0u);
                if (p < nb) {
                    mine = Board(ws.beam[p]);
                    agent_children(mine, row, c);
#pragma unroll
                    for (int a = 0; a < 4; ++a) v |= (c[a] != mine) ? (1u << a) : 0u;
                }
                int incl = __popc(v);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(FULL, incl, o);
                    if (L >= o) incl += t;
                }
                int pos = n_valid + incl - __popc(v);
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if ((v >> a) & 1u) { ws.cand[pos] = c[a].u64(); ws.first[pos] = ws.beam_first[p]; ++pos; }
                n_valid += __shfl_sync(FULL, incl, 31);
            }
        }
        __syncwarp();
        if (n_valid == 0) {
            if (d > 0) break;
            Philox4 pw = philox4x32_10(0u, call, game, DOM_BEAM, P.K);
            int pick = (int)__umulhi(pw.w[0], (uint32_t)__popc(vm));
            uint32_t msk = vm;
            for (int i = 0; i < pick; ++i) msk &= msk - 1u;
            res.action = (uint32_t)__ffs((int)msk) - 1u;
            res.prob = 0.5f;
            return res;
        }
        res.nodes += n_valid;
        const bool full_level = d >= 1 && d <= 3;
        for (int c0 = 0; c0 < n_valid; c0 += 32) {
            if (full_level) {
                spawn_and_eval<true>(c0 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
            } else {
                ChildEval ev = spawn_and_eval<false>(c0 + L, n_valid, lt_mask, spawn_base, P, game, call, phase, ws);
                if (ev.active) ws.score[c0 + L] = (double)ev.fast;
            }
        }
        __syncwarp();
        nb = min(P.width, n_valid);
        for (int c = L; c < n_valid; c += 32) {                       // stable rank by counting
            const double s = ws.score[c];
            int rank = 0;
            for (int j = 0; j < n_valid; ++j) {
                const double sj = ws.score[j];
                rank += (sj > s || (sj == s && j < c)) ? 1 : 0;
            }
            if (rank < nb) ws.slot[rank] = (uint16_t)c;
        }
        __syncwarp();
        for (int r = L; r < nb; r += 32) {
            const int c = ws.slot[r];
            ws.beam[r] = ws.cand[c];
            ws.beam_first[r] = ws.first[c];
        }
        res.best = ws.score[ws.slot[0]];
        __syncwarp();
    }
    res.action = ws.beam_first[0] & 3u;
    res.prob = 1.0f;
    return res;
}

__global__ void __launch_bounds__(kWideWarps * 32, 1) beam_search_wide_kernel(BeamArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    stage_row_table(smem, a.row);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(smem);
    const int warp = threadIdx.x >> 5;
    WideScratch &ws = reinterpret_cast<WideScratch *>(smem + kRowTableBytes)[warp];
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        unsigned int i = 0;
        if (lane == 0) i = atomicAdd(a.work, 1u);
        i = __shfl_sync(FULL, i, 0);
        if ((int64_t)i >= a.n) break;
        Board root(a.roots[i]);
        int legal = a.legal ? (int)a.legal[i] : -1;
        uint32_t call = a.call ? a.call[i] : a.call0;
        BeamResult r = beam_search_wide_warp(root, legal, a.P, a.game0 + (uint32_t)i, call, row, ws);
        if (lane == 0) {
            a.action[i] = (uint8_t)r.action;
            if (a.prob) a.prob[i] = r.prob;
            if (a.best) a.best[i] = r.best;
            if (a.nodes) a.nodes[i] = r.nodes;
        }
        __syncwarp();
    }
}

// A game that can be resumed exactly: every random draw is addressed by (game, call / spawn
// counter), so the state below is all there is.
struct GameState {
    uint64_t board;
    long long nodes;
    int32_t score, moves, valid, invalid;
    uint32_t highest, spawn_ctr, index;      // index = position in the caller's output arrays
    int32_t streak;                          // consecutive invalid moves so far
    int32_t reserved;                        // PendingKind of a pending entry (a migrated game's pusher reserved an idle group)
    int32_t ms[8];
};

// A long stall, split: the calls [gs.moves, max_moves) of a game whose agent keeps choosing an invalid move
// are cut into `segs` ranges of `seg_len` calls that different SMs search at once (every call sees the same
// board: an invalid move changes nothing).  The lowest range that holds a valid call decides where the stall
// ends; ranges above it are cancelled.  The group that finishes the last range puts the game together again.
constexpr int kMaxSegs = 32;
constexpr int kWarpSegCalls = 8;           // calls per range at least (ranges searched by one warp, play_games_kernel)
constexpr int kWarpSplitFirst = 64;        // calls the first split of a stall covers; every dry split doubles the stretch
constexpr int kWarpSplitMax = 1024;        // ... up to this many (the one-warp kernel ends when its last record does)
constexpr int kBulkStreak = 256;           // a stall this long is cut up to the move cap at once, into ranges of
constexpr int kBulkSegCalls = 32;          // ... this many calls (at most kBulkMaxSegs ranges), searched only by
constexpr int kBulkMaxSegs = 1023;         // warps / teams whose SM plays no game: bulk work must not slow the games
constexpr unsigned int kVoidTask = 0xFFFFFFFFu;    // a queue slot that was handed out but holds no range
struct SegResult { int32_t first_valid; uint32_t action; long long nodes; };
struct StallRecord {
    GameState gs;                          // the game at the split; gs.moves = first call of range 0
    int32_t seg_len, segs;
    int32_t calls;                         // calls the record covers: [gs.moves, gs.moves + calls)
    int32_t done;                          // ranges finished
    int32_t first_valid_seg;               // lowest range that found a valid call (kNoValidSeg: none so far)
    SegResult *big;                        // results of a bulk record's ranges (more than kMaxSegs), else nullptr
    SegResult res[kMaxSegs];
};
constexpr int32_t kNoValidSeg = 1 << 20;
__device__ __forceinline__ SegResult *results_of(StallRecord *rec)
{
    SegResult *big = reinterpret_cast<SegResult *>(__ldcg(reinterpret_cast<const unsigned long long *>(&rec->big)));
    return big ? big : rec->res;
}
enum PendingKind { kEntryStalled = 0, kEntryMigrated = 1, kEntryResumed = 3 };   // GameState::reserved of a pending entry

// Device-side counters of one g2048_play_games call (per-launch scratch, zeroed before the kernels).
struct GameCounters {
    unsigned int work;            // queue head of the one-warp kernel
    unsigned int finished;        // games written or handed to the stall breaker
    unsigned int tail_count;      // games handed to the team kernel
    unsigned int team_work;       // queue head of the team kernel
    unsigned int pending_count;   // stalled games handed to the stall breaker
    unsigned int finish_work;     // queue head of the stall breaker
    unsigned int written;         // games whose results are final (the stall breaker leaves when this reaches n)
    int idle_groups;              // stall-breaker groups waiting for work that no migrating game has reserved yet
    unsigned int record_count;    // StallRecords handed out
    unsigned int unused0;
    unsigned int seg_tail;        // ranges pushed to segq (one-warp kernel)
    unsigned int seg_head;        // ranges claimed from segq
    unsigned int in_stall;        // games of the one-warp kernel that are inside a split stall (not counted as alive
                                  // for the hand-over to the team kernel: they come back when their stall ends)
    unsigned int handover;        // the one-warp kernel has started handing its games to the team kernel (sticky)
    unsigned int bulk_tail;       // ranges pushed to / claimed from bulkq (long stalls: searched where no game is played)
    unsigned int bulk_head;
    unsigned int big_count;       // SegResult slots handed out of big_results
    unsigned int pad[3];
};

struct GamesArgs {
    int64_t n; BeamParams P; int max_moves; uint32_t game0;
    int32_t *score; uint8_t *highest; int32_t *moves; int32_t *valid; int32_t *invalid;
    int32_t *milestone; int64_t *nodes; uint64_t *final_board;
    const uint16_t *row; const uint8_t *code; unsigned long long *overflow;
    GameCounters *ctr;
    GameState *pending;            // stalled games handed to finish_games_kernel (nullptr: play them in place)
    unsigned int *pending_ready;   // pending[i] is complete (the stall breaker runs beside its producers)
    unsigned int pending_cap;      // entries pending[] can hold (games + ranges of split stalls)
    StallRecord *records;          // split stalls (nullptr: never split)
    unsigned int record_cap;
    unsigned int *segq;            // queued ranges of split stalls: (record << 10 | range) + 1, 0 = not written yet
    unsigned int segq_cap;
    unsigned int *bulkq;           // ... of stalls that have lasted kBulkStreak calls already (most likely up to the move cap)
    unsigned int bulkq_cap;
    SegResult *big_results;        // result slots of bulk records
    unsigned int big_cap;
    GameState *tail;               // live games handed to team_games_kernel once few are left (nullptr: never)
    unsigned int tail_threshold;   // ... i.e. once n - finished <= tail_threshold
};

__device__ __forceinline__ void write_game(const GamesArgs &a, const GameState &g)
{
    const uint32_t i = g.index;
    if (a.score) a.score[i] = g.score;
    if (a.highest) a.highest[i] = (uint8_t)g.highest;
    if (a.moves) a.moves[i] = g.moves;
    if (a.valid) a.valid[i] = g.valid;
    if (a.invalid) a.invalid[i] = g.invalid;
    if (a.milestone)
#pragma unroll
        for (int m = 0; m < 8; ++m) a.milestone[8 * i + m] = g.ms[m];
    if (a.nodes) a.nodes[i] = g.nodes;
    if (a.final_board) a.final_board[i] = g.board;
    GAME_EVT(0, i);
}

// Game2048Env() -> __init__ calls reset (env:27); state = env.reset() (evaluate_beam_search.py:30)
__device__ __forceinline__ void start_game(GameState &gs, EnvState &s, const PhiloxKey &K, uint32_t game, uint32_t index)
{
    s.spawn_ctr = 0u;
    env_reset(s, K, game);
    env_reset(s, K, game);
    gs.index = index; gs.nodes = 0; gs.moves = 0; gs.valid = 0; gs.invalid = 0; gs.streak = 0; gs.reserved = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) gs.ms[m] = -1;
}
__device__ __forceinline__ void load_env(const GameState &gs, EnvState &s)
{
    s.board = Board(gs.board); s.score = gs.score; s.highest = gs.highest; s.spawn_ctr = gs.spawn_ctr;
}
__device__ __forceinline__ void store_env(GameState &gs, const EnvState &s)
{
    gs.board = s.board.u64(); gs.score = s.score; gs.highest = s.highest; gs.spawn_ctr = s.spawn_ctr;
}
// env.step(action) and the bookkeeping of evaluate_beam_search.py:56-86; returns done.
__device__ __forceinline__ bool play_move(GameState &gs, EnvState &s, const BeamResult &r, const GamesArgs &a,
                                          const uint16_t *row, uint32_t game)
{
    gs.nodes += r.nodes;
    const StepResult st = env_step<true, false, false>(s, r.action, row, a.code, a.P.K, game, nullptr, a.overflow);
    // evaluate_beam_search.py:59-64 records the move index BEFORE `moves += 1` (:86)
#pragma unroll
    for (int m = 0; m < 8; ++m)
        if (gs.ms[m] < 0 && s.highest >= (uint32_t)(6 + m)) gs.ms[m] = gs.moves;
    if (st.valid) { ++gs.valid; gs.streak = 0; } else { ++gs.invalid; ++gs.streak; }
    ++gs.moves;
    return st.done;
}

// Consecutive invalid moves after which a game is handed to the stall breaker.  The reference's
// agent keeps choosing its fake-valid DOWN on some boards for thousands of moves
// (report.md: games reaching the move limit); one warp would grind through them one call at a time.
#ifndef G2048_STALL_STREAK
#define G2048_STALL_STREAK 32
#endif
constexpr int kStallStreak = G2048_STALL_STREAK;

// A game leaves a play loop finished (done / move cap), stalled (-> stall breaker) or, in the
// one-warp kernel, because few games are left (-> team kernel).  One thread calls this.
// Read a record another block wrote (after its flag / counter was seen): through L2, not this SM's L1
template <typename T>
__device__ __forceinline__ T load_shared_record(const T *p)
{
    static_assert(sizeof(T) % 8 == 0, "copied in 8-byte pieces");
    T out;
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(p);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(&out);
#pragma unroll
    for (size_t i = 0; i < sizeof(T) / 8; ++i) dst[i] = __ldcg(src + i);
    return out;
}
// pending[] is a ring: ticket t lives in slot t % cap, and its flag holds the ticket's generation (t / cap + 1)
// while the entry waits, 0 once it has been read.  A game is in one place at a time, so fewer than `cap` entries
// are ever outstanding and a writer practically never waits for its slot.
__device__ __forceinline__ void push_pending(const GamesArgs &a, const GameState &entry)
{
    const unsigned int t = atomicAdd(&a.ctr->pending_count, 1u);
    const unsigned int slot = t % a.pending_cap, gen = t / a.pending_cap + 1u;
    volatile unsigned int *flag = &a.pending_ready[slot];
    while (*flag != 0u) __nanosleep(100);                  // the slot's previous entry has been read
    a.pending[slot] = entry;
    __threadfence();                                       // the entry before its flag
    *flag = gen;
}
// Claims the oldest waiting entry, if any.  One thread.
__device__ __forceinline__ bool pop_pending(const GamesArgs &a, GameState &out)
{
    volatile GameCounters *c = a.ctr;
    for (;;) {
        const unsigned int h = c->finish_work;
        if (h >= c->pending_count) return false;
        if (atomicCAS(&a.ctr->finish_work, h, h + 1u) != h) continue;
        const unsigned int slot = h % a.pending_cap, gen = h / a.pending_cap + 1u;
        volatile unsigned int *flag = &a.pending_ready[slot];
        while (*flag != gen) __nanosleep(200);
        __threadfence();
        out = load_shared_record(&a.pending[slot]);
        __threadfence();
        *flag = 0u;
        return true;
    }
}
enum RetireTo { kRetirePending = 0, kRetireTail = 1, kRetireMigrate = 2 };
__device__ __forceinline__ void retire_game(const GamesArgs &a, GameState &gs, bool done, int to)
{
    if (done || gs.moves >= a.max_moves) {
        write_game(a, gs);
        atomicAdd(&a.ctr->finished, 1u);
        atomicAdd(&a.ctr->written, 1u);
    } else if (to == kRetireTail) {
        const unsigned int slot = atomicAdd(&a.ctr->tail_count, 1u);
        G2048_ASSERT(slot < (unsigned int)a.n);
        a.tail[slot] = gs;
        GAME_EVT(2, gs.index);
    } else {
        GAME_EVT(to == kRetireMigrate ? 3 : 1, gs.index);
        gs.reserved = to == kRetireMigrate ? kEntryMigrated : kEntryStalled;
        push_pending(a, gs);
        if (to == kRetirePending) atomicAdd(&a.ctr->finished, 1u);
    }
}

// The hand-over to the team kernel starts when the games the one-warp kernel is still PLAYING (not finished, not
// parked, not inside a split stall, not handed over yet) fall to the threshold; from then on every game leaves
// after its current move, and a game that comes back from a stall is handed over at once (the flag is sticky).
__device__ __forceinline__ bool handover_started(const GamesArgs &a)
{
    volatile GameCounters *c = a.ctr;
    if (c->handover) return true;
    if ((int)a.n - (int)c->finished - (int)c->in_stall - (int)c->tail_count > (int)a.tail_threshold) return false;
    c->handover = 1u;
    return true;
}

// Whole games (evaluate_beam_search.py:16-98), throughput form: one warp plays one game, fetching
// the next game id from a global queue so long games do not strand SMs.  Once the games that are
// still alive would all fit the team kernel (one team of four warps each), every warp hands its
// game over after its current move and leaves: from there on the chain of moves of the longest
// game, not the ALU pipe, bounds the run.
//
// Stalls are broken HERE, as they appear (a parked game would wait for the whole phase and then be the
// last one to finish): the warp whose game has made kStallStreak invalid moves in a row cuts the next
// calls -- 64 at first, twice the stretch after every dry split -- into ranges of >= 8 calls
// (StallRecord), takes range 0 itself and puts the others into `segq`.  Every warp that needs work
// takes a range before it takes a new game; warps without a game (few games per SM) wait for ranges.
// A range is searched call by call (every call sees the same board) until one chooses a valid move
// or a lower range has found one; the warp that finishes the LAST range of a record puts the game
// together again (same arithmetic as the stall breaker's run_segment) and plays on with it.
// The loop makes ONE search per iteration whatever the warp is doing, so the search is inlined once.
// ---- stalls cut into ranges searched by single warps (both games kernels) --------------------------------------
// Cuts the next calls of the stalled game `gs` (its env already stored in it) into ranges and queues ranges
// first_pushed .. segs-1 (the caller searches the ones below itself).  One thread.  Returns the record index,
// or a.record_cap when no record is left.  A stall that has lasted kBulkStreak calls is cut up to the move cap
// and goes to the bulk queue (all its ranges: first_pushed is ignored, *bulk = true).
__device__ __forceinline__ unsigned int split_stall(const GamesArgs &a, const GameState &gs, int first_pushed, bool *bulk)
{
    const int rem = a.max_moves - gs.moves;
    *bulk = false;
    int calls = min(rem, min(kWarpSplitMax, max(kWarpSplitFirst, gs.streak)));
    int seg_len = max(kWarpSegCalls, (calls + kMaxSegs - 1) / kMaxSegs);
    int segs = (calls + seg_len - 1) / seg_len;                  // no range is empty
    SegResult *big = nullptr;
    unsigned int qslot = 0u;
    if (a.bulkq && gs.streak >= kBulkStreak && rem > kWarpSplitMax) {
        const int blen = max(kBulkSegCalls, (rem + kBulkMaxSegs - 1) / kBulkMaxSegs);
        const int bsegs = (rem + blen - 1) / blen;
        const unsigned int off = atomicAdd(&a.ctr->big_count, (unsigned int)bsegs);
        if (off + (unsigned int)bsegs <= a.big_cap) {
            qslot = atomicAdd(&a.ctr->bulk_tail, (unsigned int)bsegs);
            if (qslot + (unsigned int)bsegs <= a.bulkq_cap) {
                big = a.big_results + off;
                calls = rem; seg_len = blen; segs = bsegs; first_pushed = 0;
                *bulk = true;
            } else {                                             // queue full: mark the slots it handed out as void
                for (unsigned int k = qslot; k < a.bulkq_cap; ++k)
                    *reinterpret_cast<volatile unsigned int *>(&a.bulkq[k]) = kVoidTask;
            }
        }
    }
    const unsigned int idx = atomicAdd(&a.ctr->record_count, 1u);
    if (idx >= a.record_cap) {
        G2048_ASSERT(!big);                                      // sized so that records outlast the bulk slots
        return a.record_cap;
    }
    StallRecord *r = &a.records[idx];
    r->gs = gs;
    r->segs = segs;
    r->seg_len = seg_len;
    r->calls = calls;
    r->done = 0;
    r->first_valid_seg = kNoValidSeg;
    r->big = big;
    GAME_EVT(5, gs.index);
    __threadfence();
    if (big) {
        for (int k = 0; k < segs; ++k)
            *reinterpret_cast<volatile unsigned int *>(&a.bulkq[qslot + (unsigned int)k]) = ((idx << 10) | (unsigned int)k) + 1u;
    } else if (segs > first_pushed) {
        const unsigned int count = (unsigned int)(segs - first_pushed);
        const unsigned int slot = atomicAdd(&a.ctr->seg_tail, count);
        G2048_ASSERT(slot + count <= a.segq_cap);
        for (int k = first_pushed; k < segs; ++k)
            *reinterpret_cast<volatile unsigned int *>(&a.segq[slot + (unsigned int)(k - first_pushed)]) = ((idx << 10) | (unsigned int)k) + 1u;
    }
    return idx;
}
// Claims a queued range: (record << 10 | range) + 1, or 0 when there is none.  One thread.
__device__ __forceinline__ unsigned int pop_queue(unsigned int *head_ptr, const unsigned int *tail_ptr, const unsigned int *q, unsigned int cap)
{
    volatile unsigned int *head = head_ptr;
    const volatile unsigned int *tail = tail_ptr;
    for (;;) {
        const unsigned int h = *head;
        if (h >= *tail || h >= cap) return 0u;
        if (atomicCAS(head_ptr, h, h + 1u) != h) continue;
        unsigned int task;
        while ((task = *reinterpret_cast<const volatile unsigned int *>(&q[h])) == 0u) __nanosleep(100);
        if (task == kVoidTask) continue;
        __threadfence();
        return task;
    }
}
// urgent ranges first; bulk ranges only for a caller whose SM plays no game
__device__ __forceinline__ unsigned int pop_range(const GamesArgs &a, bool bulk_ok)
{
    unsigned int task = pop_queue(&a.ctr->seg_head, &a.ctr->seg_tail, a.segq, a.segq_cap);
    if (!task && bulk_ok && a.bulkq) task = pop_queue(&a.ctr->bulk_head, &a.ctr->bulk_tail, a.bulkq, a.bulkq_cap);
    return task;
}
__device__ __forceinline__ bool ranges_waiting(const GamesArgs &a, bool bulk_ok)
{
    const volatile GameCounters *c = a.ctr;
    if (c->seg_head < c->seg_tail) return true;
    return bulk_ok && a.bulkq && c->bulk_head < c->bulk_tail && c->bulk_head < a.bulkq_cap;
}

// What a warp of play_games_kernel knows about its game / range, in shared memory: it is touched once per move
// (by lane 0), while the search between two moves wants every register.
struct WarpGame {
    GameState gs;
    StallRecord *rec;              // the range being searched: calls [m, hi) of `rec`, `lo` = its first call
    int seg, lo, hi, m;
    long long seg_nodes;
};

__global__ void __launch_bounds__(kBeamThreads, 1) play_games_kernel(GamesArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ WarpGame games[kBeamWarps];
    stage_row_table(smem, a.row);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(smem);
    const int warp = threadIdx.x >> 5;
    WarpScratch &ws = reinterpret_cast<WarpScratch *>(smem + kRowTableBytes)[warp];
    WarpGame &wg = games[warp];
    GameState &gs = wg.gs;
    const uint32_t lane = threadIdx.x & 31u;
    const bool can_split = a.pending && a.records && a.segq;
    const bool bulk_ok = a.tail == nullptr;                // bulk ranges wait for the team kernel, if one follows
#ifdef G2048_TEAM_PROFILE
    if (threadIdx.x == 0) atomicMin(&g_game_ts[6][0], prof_now());
#endif
    enum { kNeedWork = 0, kPlaying = 1, kInRange = 2 };
    int state = kNeedWork;
    unsigned int nap = 2000u;                              // lane 0: sleep between two looks at the queues while idle
    bool queue_dry = false, done = false;
    EnvState s;                    // board, score, highest tile, spawn counter: registers (the search starts from the board)
    uint32_t game = 0u, legal = 0u;

    // env.step(action) and the bookkeeping of evaluate_beam_search.py:56-86 (play_move) on the shared-memory record
    auto make_move = [&](uint32_t action, int nodes) {
        const StepResult st = env_step<true, false, false>(s, action, row, a.code, a.P.K, game, nullptr, a.overflow);
        if (lane == 0) {
            gs.nodes += nodes;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (gs.ms[k] < 0 && s.highest >= (uint32_t)(6 + k)) gs.ms[k] = gs.moves;
            if (st.valid) { ++gs.valid; gs.streak = 0; } else { ++gs.invalid; ++gs.streak; }
            ++gs.moves;
        }
        __syncwarp();
        done = st.done;
    };
    auto begin_range = [&](StallRecord *r, int k) {
        if (lane == 0) {
            gs = load_shared_record(&r->gs);
            const int seg_len = __ldcg(&r->seg_len);
            wg.rec = r; wg.seg = k;
            wg.lo = gs.moves + k * seg_len;
            wg.hi = min(wg.lo + seg_len, gs.moves + __ldcg(&r->calls));
            wg.m = wg.lo;
            wg.seg_nodes = 0;
        }
        __syncwarp();
        load_env(gs, s);
        game = a.game0 + gs.index;
        legal = env_legal_mask(s.board);
        state = kInRange;
    };
    // Ends the range (found = offset of its valid call or -1).  The warp that finishes the record's last range
    // re-assembles the game: then state = kPlaying with the game after the stall's valid move, or still inside
    // the stall when the record covered no valid call (the next split is longer).
    auto end_range = [&](int found, uint32_t action) {
        unsigned int last = 0u;
        StallRecord *rec = wg.rec;
        const int segs = __ldcg(&rec->segs);
        if (lane == 0) {
            const int seg = wg.seg;
            SegResult *out = results_of(rec) + seg;
            out->first_valid = found; out->action = action; out->nodes = wg.seg_nodes;
            if (found >= 0) atomicMin(&rec->first_valid_seg, seg);
            __threadfence();
            last = atomicAdd(&rec->done, 1) + 1 == segs;
        }
        last = __shfl_sync(FULL, last, 0);
        state = kNeedWork;
        if (!last) return;
        __threadfence();
        const int f = *reinterpret_cast<volatile int32_t *>(&rec->first_valid_seg);
        const volatile SegResult *res = results_of(rec);
        const int seg_len = __ldcg(&rec->seg_len), calls = __ldcg(&rec->calls);
        state = kPlaying;
        done = false;
        if (f >= segs) {                                   // no valid call in this stretch: all of it was invalid moves
            if (lane == 0) {
                for (int k = 0; k < segs; ++k) gs.nodes += res[k].nodes;
                gs.invalid += calls; gs.moves += calls; gs.streak += calls;
            }
            __syncwarp();
            return;
        }
        const int before = f * seg_len + res[f].first_valid;      // invalid moves: nothing else changes (env:188-192)
        if (lane == 0) {
            for (int k = 0; k <= f; ++k) gs.nodes += res[k].nodes;
            gs.invalid += before;
            gs.moves += before;
            gs.reserved = 0;
            atomicSub(&a.ctr->in_stall, 1u);
        }
        __syncwarp();
        make_move(res[f].action, 0);
    };

    for (;;) {
        if (state == kNeedWork) {
            if (can_split) {                               // a range of somebody's stall comes first
                unsigned int task = 0u;
                if (lane == 0) task = pop_range(a, bulk_ok);
                task = __shfl_sync(FULL, task, 0);
                if (task) { nap = 2000u; begin_range(&a.records[(task - 1u) >> 10], (int)((task - 1u) & 1023u)); }
            }
            if (state == kNeedWork && !queue_dry) {
                unsigned int g = 0;
                if (lane == 0) g = atomicAdd(&a.ctr->work, 1u);
                g = __shfl_sync(FULL, g, 0);
                if ((int64_t)g >= a.n) queue_dry = true;
                else {
                    game = a.game0 + g;
                    GameState fresh;
                    start_game(fresh, s, a.P.K, game, g);
                    if (lane == 0) gs = fresh;
                    __syncwarp();
                    done = false;
                    state = kPlaying;
                }
            }
            if (state == kNeedWork) {                      // no game, no range
                if (!can_split) break;
                unsigned int over = 0u;
                if (lane == 0) {
                    const volatile GameCounters *c = a.ctr;
                    over = (a.tail ? handover_started(a) : c->finished >= (unsigned int)a.n) && !ranges_waiting(a, bulk_ok);
                    if (!over) { __nanosleep(nap); nap = min(2u * nap, 32000u); }   // idle warps back off
                }
                if (__shfl_sync(FULL, over, 0)) break;
                continue;
            }
        }
        if (state == kPlaying) {
            const int moves = gs.moves, streak = gs.streak;
            bool leave = done || moves >= a.max_moves;
            int to = kRetirePending;
            if (!leave && a.tail) {                        // warp-uniform: lane 0 reads, everyone follows
                unsigned int go = 0u;
                if (lane == 0) go = handover_started(a);
                if (__shfl_sync(FULL, go, 0)) { leave = true; to = kRetireTail; }
            }
            if (!leave && a.pending && streak >= kStallStreak) {
                // split the next stretch of calls into ranges
                unsigned int idx = a.record_cap, bulk = 0u;
                if (can_split && lane == 0) {
                    store_env(gs, s);
                    const bool counted = gs.reserved != 0;
                    bool b = false;
                    gs.reserved = 1;
                    idx = split_stall(a, gs, 1, &b);
                    bulk = b;
                    if (idx >= a.record_cap) gs.reserved = counted ? 1 : 0;
                    else if (!counted) atomicAdd(&a.ctr->in_stall, 1u);        // counted until the stall ends
                }
                idx = __shfl_sync(FULL, idx, 0);
                bulk = __shfl_sync(FULL, bulk, 0);
                if (idx >= a.record_cap) leave = true;     // no record left: park the game for the stall breaker
                else if (bulk) { state = kNeedWork; continue; }                // all of it is queued: the game comes back later
                else begin_range(&a.records[idx], 0);
            }
            if (leave) {
                if (lane == 0) {
                    store_env(gs, s);
                    if (gs.reserved) { atomicSub(&a.ctr->in_stall, 1u); gs.reserved = 0; }   // move cap / handed over inside a stall
                    retire_game(a, gs, done, to);
                }
                __syncwarp();
                state = kNeedWork;
                continue;
            }
        }
        // ---- one get_action call: the game's next move, or call m of the range -----------------------------
        const uint32_t call = state == kInRange ? (uint32_t)wg.m : (uint32_t)gs.moves;
        const BeamResult r = beam_search_warp(s.board, -1, a.P, game, call, row, ws);
        if (state == kPlaying) {
            make_move(r.action, r.nodes);
        } else {
            const int m = wg.m;
            __syncwarp();
            if (lane == 0) { wg.seg_nodes += r.nodes; wg.m = m + 1; }
            __syncwarp();
            if ((legal >> r.action) & 1u) end_range(m - wg.lo, r.action);
            else {
                unsigned int cancelled = 0u;               // a lower range ended the stall
                if (lane == 0) cancelled = *reinterpret_cast<volatile int32_t *>(&wg.rec->first_valid_seg) < wg.seg;
                cancelled = __shfl_sync(FULL, cancelled, 0);
                if (m + 1 >= wg.hi || cancelled) end_range(-1, 0u);
            }
        }
    }
}

// Whole games with a wide beam (33..128): the same loop on the shared-memory search path, one warp
// per game, no stall breaker (a compatibility path like beam_search_wide_kernel).
__global__ void __launch_bounds__(kWideWarps * 32, 1) play_games_wide_kernel(GamesArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    stage_row_table(smem, a.row);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(smem);
    const int warp = threadIdx.x >> 5;
    WideScratch &ws = reinterpret_cast<WideScratch *>(smem + kRowTableBytes)[warp];
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        unsigned int g = 0;
        if (lane == 0) g = atomicAdd(&a.ctr->work, 1u);
        g = __shfl_sync(FULL, g, 0);
        if ((int64_t)g >= a.n) break;
        const uint32_t game = a.game0 + g;
        EnvState s;
        GameState gs;
        start_game(gs, s, a.P.K, game, g);
        bool done = false;
        while (!done && gs.moves < a.max_moves) {
            const BeamResult r = beam_search_wide_warp(s.board, -1, a.P, game, (uint32_t)gs.moves, row, ws);
            done = play_move(gs, s, r, a, row, game);
        }
        store_env(gs, s);
        if (lane == 0) write_game(a, gs);
        __syncwarp();
    }
}

// Stall breaker: kSpecWarps warps per game search the next kSpecWarps get_action calls of the SAME
// board at once.  An invalid move changes nothing in the env (no spawn, no counter), so call m+1
// sees the board of call m whenever move m is invalid: the first call whose action is valid is the
// one the sequential loop would have reached, the calls before it are its invalid moves, the calls
// after it are discarded.  Exact, and up to kSpecWarps times faster through a stall.
// kSpecWarps = 8 (three games per block) when many games may be stalled, = 24 (one game per
// block) when the launch is small enough that a whole SM per stalled game is available.
// After a valid move the next call most likely is valid too: that one call is searched by the
// group's first four warps as a team (latency form); the width doubles again with every round
// that finds no valid move.
struct SpecSlot { uint32_t action; int32_t nodes; };

// The whole block calls this.  It returns when every game of the g2048_play_games call is final, so
// it may run BESIDE the producers of stalled games: entries are claimed as they are published.
template <int kSpecWarps>
__device__ __noinline__ void break_stalls(const GamesArgs &a, uint8_t *smem, const uint16_t *row)
{
    constexpr int kSpecGroups = kBeamWarps / kSpecWarps;
    static_assert(kSpecGroups * kSpecWarps == kBeamWarps && kSpecWarps % kTeamWarps == 0, "groups must tile the block");
    __shared__ SpecSlot slots[kSpecGroups][2][kSpecWarps];
    __shared__ unsigned int next_game[kSpecGroups];
    __shared__ GameState taken[kSpecGroups];               // the entry the leader claimed from pending[]
    const int warp = threadIdx.x >> 5;
    WarpScratch &ws = reinterpret_cast<WarpScratch *>(smem + kRowTableBytes)[warp];
    const uint32_t lane = threadIdx.x & 31u;
    const int group = warp / kSpecWarps, w = warp % kSpecWarps;
    const bool leader = w == 0 && lane == 0;
    const int quad = warp >> 2;                            // the group's team = its first four warps
    TeamScratch &ts = team_scratch(smem, quad);
    auto group_barrier = [&]() { asm volatile("barrier.sync %0, %1;" ::"r"(8 + group), "r"(kSpecWarps * 32) : "memory"); };
    constexpr unsigned int kNone = 0xFFFFFFFFu;
    int buf = 0;

    for (;;) {
        if (leader) {
            volatile unsigned int *written = &a.ctr->written;
            unsigned int got = kNone;
            bool counted_idle = false;                     // this group is in ctr->idle_groups
            for (;;) {
                if (pop_pending(a, taken[group])) {
                    // a migrated game's pusher took one idle group off the count for it; keep the count
                    // right whoever ends up with the entry
                    const bool reserved = taken[group].reserved == kEntryMigrated;
                    if (counted_idle && !reserved) atomicSub(&a.ctr->idle_groups, 1);
                    if (!counted_idle && reserved) atomicAdd(&a.ctr->idle_groups, 1);
                    got = 0u;
                    break;
                }
                if (*written >= (unsigned int)a.n) break;
                if (!counted_idle) { atomicAdd(&a.ctr->idle_groups, 1); counted_idle = true; }
                __nanosleep(1000);
            }
            next_game[group] = got;
        }
        group_barrier();
        const unsigned int p = next_game[group];
        group_barrier();                                   // everyone has read it before the next round rewrites it
        if (p == kNone) break;
        GameState gs = taken[group];                       // every warp of the group keeps an identical copy
        group_barrier();                                   // ... before the leader claims the next entry
        EnvState s;
        bool done = false;
        if (leader) { GAMES_PROF_ADD(6, 1); GAME_EVT(4, gs.index); }
        load_env(gs, s);
        const uint32_t game = a.game0 + gs.index;
        int width = gs.streak >= kStallStreak && gs.reserved == kEntryStalled ? kSpecWarps : 1;   // inside a stall, or normal play
        while (!done && gs.moves < a.max_moves) {
            const int allowed = min(width, a.max_moves - gs.moves);
            if (allowed == 1) {
                if (w < kTeamWarps) {
                    const BeamResult r = beam_search_team(s.board, -1, a.P, game, (uint32_t)gs.moves, row, ts, 1 + quad);
                    if (leader) { slots[group][buf][0].action = r.action; slots[group][buf][0].nodes = r.nodes; }
                }
            } else if (w < allowed) {
                const BeamResult r = beam_search_warp(s.board, -1, a.P, game, (uint32_t)(gs.moves + w), row, ws);
                if (lane == 0) { slots[group][buf][w].action = r.action; slots[group][buf][w].nodes = r.nodes; }
            }
            group_barrier();
            if (leader) GAMES_PROF_ADD(7, 1);
            const uint32_t legal = env_legal_mask(s.board);
            int first_valid = -1;
            for (int j = 0; j < allowed; ++j) {
                gs.nodes += slots[group][buf][j].nodes;
                if ((legal >> slots[group][buf][j].action) & 1u) { first_valid = j; break; }
            }
            const int n_invalid = first_valid >= 0 ? first_valid : allowed;
            gs.invalid += n_invalid;
            gs.moves += n_invalid;                          // invalid moves: nothing else changes (env:188-192)
            if (first_valid >= 0) {
                BeamResult r;
                r.action = slots[group][buf][first_valid].action;
                r.nodes = 0;
                done = play_move(gs, s, r, a, row, game);
                width = 1;
            } else {
                width = min(kSpecWarps, 2 * width);
            }
            buf ^= 1;
        }
        store_env(gs, s);
        if (leader) { write_game(a, gs); __threadfence(); atomicAdd(&a.ctr->written, 1u); }
    }
}

// The stall breaker on its own: after the one-warp kernel when no team kernel follows it.
template <int kSpecWarps>
__global__ void __launch_bounds__(kBeamThreads, 1) finish_games_kernel(GamesArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    stage_row_table(smem, a.row);
    break_stalls<kSpecWarps>(a, smem, reinterpret_cast<const uint16_t *>(smem));
}

// Whole games, latency form: one TEAM of four warps plays one game (beam_search_team).  Games come from
// `in` (hand-overs of play_games_kernel, ctr->tail_count of them) or, with in == nullptr, are the fresh
// games 0..n-1, and later from pending[] (games that resume after a stall, games that migrate).
// Always launched with kBeamThreads threads on every SM; every team stays until every game of the call
// is final, because a team without a game is what breaks stalls:
//   * a team whose game has made kStallStreak invalid moves in a row cuts the game's next calls into ranges
//     (split_stall) and queues them; it is then free for other work;
//   * a free team first looks for a game in pending[] (a game coming back from a stall is most likely on
//     the critical path), then for a fresh game, then each of its four warps takes ONE queued range and
//     searches it with the one-warp search -- so whatever part of the GPU is not playing searches ranges,
//     from the first stall on, and a resumed game waits for the end of one range at most;
//   * the warp that finishes the last range of a record puts the game together: it goes to pending[] when
//     the stall ended in a valid move, is cut again (twice the stretch) when the stretch held none.
// Once the queue of fresh games is dry, SMs that still hold several games run them slower (six teams share
// four schedulers; three run at full speed) than an SM holding few: a team that is one of more than three
// playing on its SM checks every fourth move for teams that wait on an emptier SM and sends its game there.
struct TeamJob {
    GameState gs;                  // the game a team is about to play (leader -> team)
    int32_t kind;
};
enum { kJobExit = 0, kJobGame = 1, kJobRanges = 2 };

__global__ void __launch_bounds__(kBeamThreads, 1) team_games_kernel(GamesArgs a, const GameState *in)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kTeams = kBeamWarps / kTeamWarps;
    __shared__ TeamJob jobs[kTeams];
    __shared__ WarpGame ranges[kBeamWarps];
    __shared__ int active_teams;                           // teams of this block that are playing a game
    stage_row_table(smem, a.row);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(smem);
    const int warp = threadIdx.x >> 5, quad = warp >> 2;
    const uint32_t lane = threadIdx.x & 31u;
    TeamScratch &ts = team_scratch(smem, quad);            // overlays the four WarpScratch the range searches use
    WarpScratch &ws = reinterpret_cast<WarpScratch *>(smem + kRowTableBytes)[warp];
    TeamJob &job = jobs[quad];
    WarpGame &wg = ranges[warp];
    const int bar = 1 + quad;
    const bool leader = (threadIdx.x & (kTeamThreads - 1)) == 0;
    const bool can_split = a.pending && a.records && a.segq;
    const unsigned int total = in ? a.ctr->tail_count : (unsigned int)a.n;
    if (threadIdx.x == 0) active_teams = 0;
    __syncthreads();
    if (threadIdx.x == 0) GAMES_PROF_MIN(0);
    // Team q of a block asks a little later than team q - 1: when there are fewer games than teams, they
    // spread over all SMs instead of filling a few.
    if (leader) __nanosleep(4000u * (unsigned int)quad);
    // ctr->idle_groups = room for full-speed games: the sum over all blocks of max(0, 3 - games played on the block),
    // kept exact by the transitions of every block's `active_teams` (a migrating game reserves its place beforehand).
    if (threadIdx.x == 0) atomicAdd(&a.ctr->idle_groups, 3);
    // A team takes a game that waits in pending[] at once when its block plays fewer than three (room for a full-speed
    // game), otherwise only after a while; bulk ranges are searched by the teams of blocks that play no game.
    // (A/B knobs, measured within noise of each other on cfg 5: which teams of a block take games from pending[],
    // and which search bulk ranges while the block plays no game)
#ifndef G2048_RECEIVERS
#define G2048_RECEIVERS 6
#endif
#ifndef G2048_BULK_TEAMS_FROM
#define G2048_BULK_TEAMS_FROM 0
#endif
    const bool receiver = quad < G2048_RECEIVERS;
    const bool bulk_team = quad >= G2048_BULK_TEAMS_FROM;
    bool queue_dry = false;                                // leader only

    for (;;) {
        // ---- the leader finds the team's next job -------------------------------------------------------------
        if (leader) {
            volatile GameCounters *c = a.ctr;
            int kind = -1;
            bool deferred = false, reserved = false;
            unsigned int nap = 1000u;                      // idle leaders back off: five of them share an SM with a lone game
            while (kind < 0) {
                if (a.pending && c->finish_work < c->pending_count) {           // 1. a game that waits in pending[]
                    // ... goes to a receiver on a block with room; anybody else waits a while first (then nobody with room is free)
                    if (!(receiver && *reinterpret_cast<volatile int *>(&active_teams) < 3) && !deferred) {
                        deferred = true;
                        __nanosleep(50000);
                        continue;
                    }
                    if (pop_pending(a, job.gs)) {
                        reserved = job.gs.reserved == kEntryMigrated;          // its pusher took the room it needs off the count
                        GAME_EVT(4, job.gs.index);
                        kind = kJobGame;
                        break;
                    }
                }
                if (!queue_dry) {                                              // 2. a fresh (or handed-over) game
                    const unsigned int p = atomicAdd(&a.ctr->team_work, 1u);
                    if (p < total) {
                        G2048_ASSERT(!in || p < (unsigned int)a.n);
                        if (in) job.gs = in[p];
                        else {
                            EnvState s0;
                            start_game(job.gs, s0, a.P.K, a.game0 + p, p);
                            store_env(job.gs, s0);
                        }
                        kind = kJobGame;
                        break;
                    }
                    queue_dry = true;
                    GAMES_PROF_MIN(1);
                }
                if (can_split && ranges_waiting(a, bulk_team && *reinterpret_cast<volatile int *>(&active_teams) == 0)) {   // 3. ranges of split stalls
                    kind = kJobRanges;
                    break;
                }
                if (c->written >= (unsigned int)a.n) { kind = kJobExit; break; }               // 4. every game is final
                __nanosleep(nap);
                nap = min(2u * nap, 32000u);
            }
            if (kind == kJobGame) {
                const bool uses_room = atomicAdd(&active_teams, 1) < 3;
                if (reserved && !uses_room) atomicAdd(&a.ctr->idle_groups, 1);                 // the reservation goes back
                if (!reserved && uses_room) atomicSub(&a.ctr->idle_groups, 1);
            }
            job.kind = kind;
        }
        team_barrier(bar);
        const int kind = job.kind;
        if (kind == kJobExit) break;

        if (kind == kJobGame) {
            // The game's record stays in shared memory (job.gs, kept by the leader): across a search the threads only
            // hold the env (five registers), the move count and the streak -- the search wants every register.
            GameState &gs = job.gs;
            EnvState s;
            load_env(gs, s);
            const uint32_t game = a.game0 + gs.index;
            int moves = gs.moves, streak = gs.streak;
            bool done = false, migrate = false, split = false, may_split = can_split;
            while (!done && moves < a.max_moves) {
                if (a.pending && streak >= kStallStreak && may_split) {
                    // the agent keeps choosing an invalid move: cut the next calls into ranges for every free warp
                    if (leader) {
                        store_env(gs, s);
                        gs.moves = moves; gs.streak = streak; gs.reserved = 0;
                        bool bulk;
                        ts.next_item = split_stall(a, gs, 0, &bulk);
                    }
                    team_barrier(bar);
                    split = ts.next_item < a.record_cap;
                    team_barrier(bar);
                    if (split) break;
                    may_split = false;                     // no record left: play through the stall
                }
                // every fourth move: does the game share a crowded SM while a team waits on an emptier one?
#ifndef G2048_NO_MIGRATE
                if (a.pending && (moves & 3) == 3) {
                    if (leader) {
                        unsigned int go = 0u;
                        if (*reinterpret_cast<volatile int *>(&active_teams) > 3 &&
                            *reinterpret_cast<volatile unsigned int *>(&a.ctr->team_work) >= total &&
                            *reinterpret_cast<volatile int *>(&a.ctr->idle_groups) > 0) {
                            if (atomicSub(&a.ctr->idle_groups, 1) > 0) go = 1u;            // reserved room on an emptier block
                            else atomicAdd(&a.ctr->idle_groups, 1);
                        }
                        ts.next_item = go;
                    }
                    team_barrier(bar);
                    migrate = ts.next_item != 0u;
                    team_barrier(bar);
                    if (migrate) break;
                }
#endif
                const BeamResult r = beam_search_team(s.board, -1, a.P, game, (uint32_t)moves, row, ts, bar);
                // env.step(action) and the bookkeeping of evaluate_beam_search.py:56-86 (play_move), the latter by the leader
                const StepResult st = env_step<true, false, false>(s, r.action, row, a.code, a.P.K, game, nullptr, a.overflow);
                if (leader) {
                    gs.nodes += r.nodes;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (gs.ms[k] < 0 && s.highest >= (uint32_t)(6 + k)) gs.ms[k] = moves;
                    if (st.valid) ++gs.valid; else ++gs.invalid;
                }
                streak = st.valid ? 0 : streak + 1;
                ++moves;
                done = st.done;
            }
            if (leader) {
                if (atomicSub(&active_teams, 1) <= 3) atomicAdd(&a.ctr->idle_groups, 1);       // room for a full-speed game again
                if (!split) {                              // finished (written) or migrating (to pending[])
                    store_env(gs, s);
                    gs.moves = moves; gs.streak = streak;
                    retire_game(a, gs, done, kRetireMigrate);
                    GAMES_PROF_MAX(2);
                }
            }
            team_barrier(bar);
            continue;
        }

        // ---- kJobRanges: every warp of the team takes one queued range ------------------------------------------
        unsigned int task = 0u;
        if (lane == 0) task = pop_range(a, bulk_team && *reinterpret_cast<volatile int *>(&active_teams) == 0);
        task = __shfl_sync(FULL, task, 0);
        if (task) {
            StallRecord *rec = &a.records[(task - 1u) >> 10];
            const int seg = (int)((task - 1u) & 1023u);
            GameState &gs = wg.gs;
            if (lane == 0) gs = load_shared_record(&rec->gs);
            __syncwarp();
            EnvState s;
            load_env(gs, s);
            const uint32_t game = a.game0 + gs.index;
            const uint32_t legal = env_legal_mask(s.board);
            const int seg_len = __ldcg(&rec->seg_len), calls = __ldcg(&rec->calls), segs = __ldcg(&rec->segs);
            const int first = gs.moves;
            const int lo = first + seg * seg_len, hi = min(lo + seg_len, first + calls);
            long long nodes = 0;
            int found = -1;
            uint32_t action = 0u;
            for (int m = lo; m < hi; ++m) {
                const BeamResult r = beam_search_warp(s.board, -1, a.P, game, (uint32_t)m, row, ws);
                nodes += r.nodes;
                if ((legal >> r.action) & 1u) { found = m - lo; action = r.action; break; }
                unsigned int cancelled = 0u;               // a lower range ended the stall
                if (lane == 0) cancelled = *reinterpret_cast<volatile int32_t *>(&rec->first_valid_seg) < seg;
                if (__shfl_sync(FULL, cancelled, 0)) break;
            }
            unsigned int last = 0u;
            if (lane == 0) {
                SegResult *out = results_of(rec) + seg;
            out->first_valid = found; out->action = action; out->nodes = nodes;
                if (found >= 0) atomicMin(&rec->first_valid_seg, seg);
                __threadfence();
                last = atomicAdd(&rec->done, 1) + 1 == segs;
            }
            last = __shfl_sync(FULL, last, 0);
            if (last) {                                    // this warp puts the game together again
                __threadfence();
                const int f = *reinterpret_cast<volatile int32_t *>(&rec->first_valid_seg);
                const volatile SegResult *res = results_of(rec);
                bool done = false;
                if (f >= segs) {                           // no valid call in this stretch: all of it was invalid moves
                    if (lane == 0) {
                        for (int k = 0; k < segs; ++k) gs.nodes += res[k].nodes;
                        gs.invalid += calls; gs.moves += calls; gs.streak += calls;
                    }
                } else {
                    const int before = f * seg_len + res[f].first_valid;      // invalid moves: nothing else changes (env:188-192)
                    if (lane == 0) {
                        for (int k = 0; k <= f; ++k) gs.nodes += res[k].nodes;
                        gs.invalid += before;
                        gs.moves += before;
                    }
                    __syncwarp();
                    GameState g2 = gs;
                    BeamResult r;
                    r.action = res[f].action;
                    r.nodes = 0;
                    done = play_move(g2, s, r, a, row, game);
                    store_env(g2, s);
                    if (lane == 0) gs = g2;
                }
                __syncwarp();
                if (lane == 0) {
                    bool bulk_unused;
                    gs.reserved = 0;
                    if (done || gs.moves >= a.max_moves) retire_game(a, gs, done, kRetirePending);       // written
                    else if (f >= segs && split_stall(a, gs, 0, &bulk_unused) < a.record_cap) { }       // still stalled: cut again
                    else { gs.reserved = kEntryResumed; push_pending(a, gs); GAME_EVT(3, gs.index); }   // plays on with a team
                }
                __syncwarp();
            }
        }
        team_barrier(bar);
    }
    if (threadIdx.x == 0) GAMES_PROF_MAX(5);
}

#ifdef G2048_TEAM_PROFILE
}  // namespace g2048
// [0..2] cycles in phases A/B/C of beam_search_team, [3] levels, [4] searches, [5] cycles in searches; resets them
extern "C" int g2048_debug_team_profile(unsigned long long *out8)
{
    unsigned long long zero[8] = {0};
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out8, g2048::g_team_prof, sizeof zero) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(g2048::g_team_prof, zero, sizeof zero) == cudaSuccess ? 0 : -1;
}
// timeline of the last team_games_kernel launches (see g_games_prof); resets it (min slots to ~0)
extern "C" int g2048_debug_games_profile(unsigned long long *out8)
{
    unsigned long long init[8] = {~0ull, ~0ull, 0, ~0ull, 0, 0, 0, 0};
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out8, g2048::g_games_prof, sizeof init) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(g2048::g_games_prof, init, sizeof init) == cudaSuccess ? 0 : -1;
}
// per-game event times of the last g2048_play_games call (see g_game_ts): out = uint64[7][16384]; resets them
extern "C" int g2048_debug_game_times(unsigned long long *out)
{
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out, g2048::g_game_ts, sizeof g2048::g_game_ts) != cudaSuccess) return -1;
    void *p = nullptr;
    if (cudaGetSymbolAddress(&p, g2048::g_game_ts) != cudaSuccess) return -1;
    if (cudaMemset(p, 0, sizeof g2048::g_game_ts) != cudaSuccess) return -1;
    const unsigned long long big = ~0ull;
    return cudaMemcpy(reinterpret_cast<unsigned long long *>(p) + 6 * g2048::kProfGames, &big, sizeof big, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
}
namespace g2048 {
#endif

static int g_attr_done[kMaxDevices];
static int g_tuning[G2048_TUNE_COUNT] = {0, -1, -1, -1, 1, -1, -1, -1, -1};
int step_tuning(int key) { return g_tuning[key]; }

int set_tuning(int key, int value)
{
    if (key < 0 || key >= G2048_TUNE_COUNT) return set_error(G2048_EINVAL, "g2048_set_tuning: unknown key %d", key);
    g_tuning[key] = value;
    return G2048_OK;
}

static int ensure_attrs()
{
    int dev = 0;
    G2048_CUDA(cudaGetDevice(&dev));
    if (!g_attr_done[dev]) {
        G2048_CUDA(cudaFuncSetAttribute(beam_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBeamSmemBytes));
        G2048_CUDA(cudaFuncSetAttribute(beam_search_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBeamSmemBytes));
        G2048_CUDA(cudaFuncSetAttribute(beam_search_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWideSmemBytes));
        G2048_CUDA(cudaFuncSetAttribute(play_games_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWideSmemBytes));
        G2048_CUDA(cudaFuncSetAttribute(play_games_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBeamSmemBytes));
        G2048_CUDA(cudaFuncSetAttribute(team_games_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBeamSmemBytes));
        G2048_CUDA(cudaFuncSetAttribute(finish_games_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBeamSmemBytes));
        G2048_CUDA(cudaFuncSetAttribute(finish_games_kernel<kBeamWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBeamSmemBytes));
        g_attr_done[dev] = 1;
    }
    return G2048_OK;
}

// Per-launch device scratch from the stream-ordered allocator: launches on different streams (or
// baked into different CUDA graphs) never share a queue head, and the memory goes back to the
// pool when the launch's kernels are done -- on every return path.
struct LaunchScratch {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    int alloc(DeviceState *st, size_t bytes, size_t zero_bytes, cudaStream_t s)
    {
        stream = s;
        G2048_CUDA(cudaMallocFromPoolAsync(&ptr, bytes, st->pool, s));
        G2048_CUDA(cudaMemsetAsync(ptr, 0, zero_bytes, s));
        return G2048_OK;
    }
    ~LaunchScratch() { if (ptr) cudaFreeAsync(ptr, stream); }
};

int launch_beam_search(DeviceState *st, const uint64_t *roots, const uint8_t *legal, const uint32_t *call,
                       uint32_t call0, uint8_t *action, float *prob, double *best_score, int32_t *nodes,
                       int64_t n, int beam_width, int search_depth, int early_thr, int mid_thr,
                       uint64_t seed, uint32_t game0, cudaStream_t stream)
{
    int rc = ensure_attrs();
    if (rc != G2048_OK) return rc;
    LaunchScratch scratch;
    rc = scratch.alloc(st, sizeof(unsigned int), sizeof(unsigned int), stream);
    if (rc != G2048_OK) return rc;
    BeamArgs a{roots, legal, call, call0, action, prob, best_score, nodes, n,
               BeamParams{beam_width, search_depth, early_thr, mid_thr, make_philox_key(seed)},
               game0, st->row, static_cast<unsigned int *>(scratch.ptr)};
    if (beam_width > 32) {                                          // wide beams: shared-memory beam, counting top-k
        int wgrid = (int)(n < st->sm_count ? n : st->sm_count);
        int64_t wwarps = (n + wgrid - 1) / wgrid;
        int wthreads = 32 * (int)(wwarps < kWideWarps ? wwarps : kWideWarps);
        beam_search_wide_kernel<<<wgrid, wthreads, kWideSmemBytes, stream>>>(a);
        count_launch();
        return check_cuda(cudaGetLastError(), "beam_search_wide_kernel");
    }
    const int teams_per_block = kBeamWarps / kTeamWarps;
    const int mode = g_tuning[G2048_TUNE_SEARCH_MODE];
    const bool team = mode == 2 || (mode == 0 && n <= (int64_t)st->sm_count * teams_per_block);
    int grid = (int)(n < st->sm_count ? n : st->sm_count);          // one block per SM; a small batch is spread over all SMs
    if (team) {                                                     // fewer roots than team slots: latency form
        int64_t per_block = (n + grid - 1) / grid;
        int threads = kTeamThreads * (int)(per_block < teams_per_block ? per_block : teams_per_block);
        beam_search_team_kernel<<<grid, threads, kBeamSmemBytes, stream>>>(a);
        count_launch();
        return check_cuda(cudaGetLastError(), "beam_search_team_kernel");
    }
    int64_t warps = (n + grid - 1) / grid;
    int threads = 32 * (int)(warps < kBeamWarps ? warps : kBeamWarps);
    beam_search_kernel<<<grid, threads, kBeamSmemBytes, stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "beam_search_kernel");
}

int launch_play_games(DeviceState *st, int64_t n, int beam_width, int search_depth, int early_thr, int mid_thr,
                      int max_moves, uint64_t seed, uint32_t game0, int32_t *score, uint8_t *highest_exp,
                      int32_t *moves, int32_t *valid, int32_t *invalid, int32_t *milestone, int64_t *nodes,
                      uint64_t *final_board, cudaStream_t stream)
{
    int rc = ensure_attrs();
    if (rc != G2048_OK) return rc;
    const bool wide = beam_width > 32;
    const int teams_per_block = kBeamWarps / kTeamWarps;
    const int64_t team_slots = (int64_t)st->sm_count * teams_per_block;
    // Games that fit the team slots (six per SM): teams from the first move.  More: one warp per game until the
    // live ones fit -- while every SM holds more than six games the run is bound by instruction throughput, and
    // a team spends ~1.8x the instructions of a lone warp on a move (measured: 1,250 games at 20/40 take 0.199 s
    // with teams from the start, 0.175 s with the hand-over at 888 live games).
    const int64_t direct_max = g_tuning[G2048_TUNE_TEAM_DIRECT_MAX] >= 0 ? g_tuning[G2048_TUNE_TEAM_DIRECT_MAX] : team_slots;
    const int64_t tail_thr = g_tuning[G2048_TUNE_TAIL_THRESHOLD] >= 0 ? g_tuning[G2048_TUNE_TAIL_THRESHOLD] : team_slots;
    const bool direct = !wide && n <= direct_max;
    // (games that are inside a split stall at the hand-over follow later: the tail array can hold every game)
    const int64_t tail_cap = wide || direct || tail_thr <= 0 ? 0 : n;
    const int64_t tail_threshold = tail_thr < n ? tail_thr : n;
    // scratch: counters | ready flags of the pending entries | range queue of the one-warp kernel | pending entries
    // (stalled / migrated games, ranges of stalls the stall breaker split) | split-stall records | games for the
    // team kernel.  Everything before the pending entries is zeroed.
    const size_t record_cap = wide ? 0 : 2 * (size_t)n + 64;
    const size_t pending_cap = wide ? 0 : g_tuning[G2048_TUNE_PENDING_CAP] > 0 ? (size_t)g_tuning[G2048_TUNE_PENDING_CAP]
                                                                                : (size_t)n + 64;   // a ring: a game is in one place at a time
    const size_t segq_cap = wide ? 0 : kMaxSegs * (2 * (size_t)n + 64);
    // bulk records (stalls that run to the move cap, ~5 % of the games): n / 8 + 16 of them, cut into ranges of 32 calls
    const size_t bulk_segs = (size_t)((max_moves + kBulkSegCalls - 1) / kBulkSegCalls < kBulkMaxSegs
                                          ? (max_moves + kBulkSegCalls - 1) / kBulkSegCalls : kBulkMaxSegs);
    const size_t bulkq_cap = wide ? 0 : ((size_t)n / 8 + 16) * (bulk_segs > 0 ? bulk_segs : 1);
    auto round256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t flags_off = 256, segq_off = flags_off + round256(pending_cap * sizeof(unsigned int));
    const size_t bulkq_off = segq_off + round256(segq_cap * sizeof(unsigned int));
    const size_t pending_off = bulkq_off + round256(bulkq_cap * sizeof(unsigned int));
    const size_t records_off = pending_off + round256(pending_cap * sizeof(GameState));
    const size_t big_off = records_off + round256(record_cap * sizeof(StallRecord));
    const size_t tail_off = big_off + round256(bulkq_cap * sizeof(SegResult));
    LaunchScratch scratch;
    rc = scratch.alloc(st, tail_off + (size_t)tail_cap * sizeof(GameState), pending_off, stream);
    if (rc != G2048_OK) return rc;
    uint8_t *base = static_cast<uint8_t *>(scratch.ptr);
    const bool split = !wide && g_tuning[G2048_TUNE_SPLIT_STALLS] != 0;
    GamesArgs a{n, BeamParams{beam_width, search_depth, early_thr, mid_thr, make_philox_key(seed)},
                max_moves, game0, score, highest_exp, moves, valid, invalid, milestone, nodes, final_board,
                st->row, st->code, st->overflow, reinterpret_cast<GameCounters *>(base),
                wide ? nullptr : reinterpret_cast<GameState *>(base + pending_off),
                wide ? nullptr : reinterpret_cast<unsigned int *>(base + flags_off), (unsigned int)pending_cap,
                split ? reinterpret_cast<StallRecord *>(base + records_off) : nullptr,
                (unsigned int)record_cap,
                split && segq_cap ? reinterpret_cast<unsigned int *>(base + segq_off) : nullptr, (unsigned int)segq_cap,
                split && bulkq_cap ? reinterpret_cast<unsigned int *>(base + bulkq_off) : nullptr, (unsigned int)bulkq_cap,
                reinterpret_cast<SegResult *>(base + big_off), (unsigned int)bulkq_cap,
                tail_cap ? reinterpret_cast<GameState *>(base + tail_off) : nullptr, (unsigned int)tail_threshold};
    const int grid = (int)(n < st->sm_count ? n : st->sm_count);    // spread small runs over all SMs (see beam search)
    if (wide) {                                                     // wide beams: compatibility path, no stall breaker
        int64_t wwarps = (n + grid - 1) / grid;
        int wthreads = 32 * (int)(wwarps < kWideWarps ? wwarps : kWideWarps);
        play_games_wide_kernel<<<grid, wthreads, kWideSmemBytes, stream>>>(a);
        count_launch();
        return check_cuda(cudaGetLastError(), "play_games_wide_kernel");
    }
    // ~5 % of games stall.  Up to 2048 games that is at most ~100 of them: a whole SM (24 speculative calls
    // per round) each; beyond, three games per SM (8 calls per round each).
    const bool whole_sm = n <= 2048;
    if (direct) {
        team_games_kernel<<<st->sm_count, kBeamThreads, kBeamSmemBytes, stream>>>(a, nullptr);
        count_launch();
        return check_cuda(cudaGetLastError(), "team_games_kernel");
    }
    play_games_kernel<<<st->sm_count, kBeamThreads, kBeamSmemBytes, stream>>>(a);
    count_launch();
    G2048_CUDA(cudaGetLastError());
    if (tail_cap) {                                                 // how many were handed over is only known on the device
        team_games_kernel<<<st->sm_count, kBeamThreads, kBeamSmemBytes, stream>>>(a, a.tail);
        count_launch();
        return check_cuda(cudaGetLastError(), "team_games_kernel");
    }
    if (whole_sm) {
        finish_games_kernel<kBeamWarps><<<grid, kBeamThreads, kBeamSmemBytes, stream>>>(a);
    } else {
        int64_t groups = (n + 2) / 3;
        int grid2 = (int)(groups < st->sm_count ? groups : st->sm_count);
        finish_games_kernel<8><<<grid2, kBeamThreads, kBeamSmemBytes, stream>>>(a);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "finish_games_kernel");
}

}  // namespace g2048
