// team.cuh -- BeamSearchAgent.get_action (agents/beam_search_agent.py:71-181) by a TEAM of four
// warps, one child per thread.  Included by beam.cu after the one-warp search.
//
// The one-warp search (beam_search_warp) is the throughput form: 24 independent roots per SM keep
// the ALU pipe busy, but a single warp runs a level's ~1,100 instructions as one dependent chain
// (IPC ~0.3, ~60 us per move at width 20 / depth 40).  Whole games are sequential chains of moves,
// so once fewer games are alive than warps fit on the GPU the chain, not the pipe, sets the time.
// Here warp a of the team owns ACTION a and lane r owns the beam entry of RANK r: every thread
// makes one child (a warp-uniform direction: no selects), spawns and evaluates it, and the stable
// top-k is a counting rank over the level's <= 128 keys in shared memory (broadcast 16-byte
// loads, no shuffle network).  Three named barriers per level replace ~45 dependent shuffles.
//   generation order (agent:142-167) = (rank, action)  ->  position and spawn ordinal of a child
//   come from the four validity / draw ballots: popc over lower ranks + the lower actions of its
//   own parent.  Results are bit-identical to the one-warp search (same tests).
#pragma once

namespace g2048 {

constexpr int kTeamWarps = 4;
constexpr int kTeamThreads = kTeamWarps * 32;
constexpr uint32_t kTeamRing = 512;                  // spawn ordinals held per team (power of two)

struct __align__(16) TeamScratch {
    uint4 rng[kTeamRing / 2];       // spawn words of ordinals [ring_end - 512, ring_end): one Philox block per entry
    double score[kTeamThreads + 8]; // float64 scores of a full-evaluation level, generation order (+ padding of a chunk)
    uint32_t key[kTeamThreads];     // sort keys of the level's children in generation order, padded with 0 to a multiple of 16
    uint64_t beam[32];              // parents of the level, rank order
    uint32_t ballots[8];            // [a]: valid children of action a (bit = parent rank); [4 + a]: those that draw
    uint8_t meta[32];               // first action | largest exponent << 2 of beam[i]
    double out_best;                // score of the best candidate of the last level
    uint32_t out_first;             // its first action
    uint32_t next_item;             // work-queue hand-off of the kernels built on the team
};
static_assert(sizeof(TeamScratch) <= kTeamWarps * sizeof(WarpScratch), "a team reuses the scratch of its four warps");

// -DG2048_TEAM_PROFILE: cycles per phase of the team search, summed by each team's first thread
// (profiles/tail.py reads them through g2048_debug_team_profile; never part of the product build)
#ifdef G2048_TEAM_PROFILE
__device__ unsigned long long g_team_prof[8];
// timeline of a team_games_kernel launch (globaltimer, ns): [0] first block started, [1] queue first found
// empty, [2] last game retired by a team, [3] first / [4] last block turning stall breaker, [5] kernel end,
// [6] stalled games taken by the stall breaker, [7] speculative rounds
__device__ unsigned long long g_games_prof[8];
__device__ __forceinline__ unsigned long long prof_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// per-game event times (globaltimer, ns): [0] final results written, [1] parked as stalled, [2] handed to the team
// kernel, [3] migrated, [4] claimed by a stall-breaker group, [5] its stall was split; [6][0] = start of the one-warp kernel
constexpr int kProfGames = 16384;
__device__ unsigned long long g_game_ts[7][kProfGames];
#define GAME_EVT(kind, index) (g_game_ts[kind][(index) & (kProfGames - 1)] = prof_now())
#define GAMES_PROF_MIN(i) atomicMin(&g_games_prof[i], prof_now())
#define GAMES_PROF_MAX(i) atomicMax(&g_games_prof[i], prof_now())
#define GAMES_PROF_ADD(i, v) atomicAdd(&g_games_prof[i], (unsigned long long)(v))
#define TEAM_PROF_DECL long long _tp = clock64(), _tp0 = _tp; unsigned long long _pa = 0, _pb = 0, _pc = 0, _pl = 0
#define TEAM_PROF_MARK(acc) do { const long long _n = clock64(); acc += (unsigned long long)(_n - _tp); _tp = _n; } while (0)
#define TEAM_PROF_FLUSH() do { if (tid == 0u) { atomicAdd(&g_team_prof[0], _pa); atomicAdd(&g_team_prof[1], _pb); \
    atomicAdd(&g_team_prof[2], _pc); atomicAdd(&g_team_prof[3], _pl); atomicAdd(&g_team_prof[4], 1ull); \
    atomicAdd(&g_team_prof[5], (unsigned long long)(clock64() - _tp0)); } } while (0)
#else
#define TEAM_PROF_DECL
#define TEAM_PROF_MARK(acc)
#define TEAM_PROF_FLUSH()
#define GAMES_PROF_MIN(i)
#define GAMES_PROF_MAX(i)
#define GAMES_PROF_ADD(i, v)
#define GAME_EVT(kind, index)
#endif

__device__ __forceinline__ void team_barrier(int id)
{
    asm volatile("barrier.sync %0, %1;" ::"r"(id), "n"(kTeamThreads) : "memory");
}

// BeamSearchAgent._make_move (agent:194-258) for a warp-uniform direction (DOWN with its rotation
// bug): same results as agent_children(), one direction per call.
__device__ __forceinline__ Board agent_child_uniform(Board b, uint32_t action, const uint16_t *row)
{
    switch (action) {
    case 0: return move_left<true>(b, row);
    case 1: return transpose(move_left<true>(transpose(b), row));
    case 2: return flip_rows(move_left<true>(flip_rows(b), row));
    default: return transpose(flip_row_order(move_left<true>(flip_rows(transpose(b)), row)));
    }
}

// #keys among keys[0 .. n) larger than `key`; keys[n .. n + 15] must hold 0 (chunks of 16).  The keys sit in
// generation order without gaps (a level of width 20 has 80 slots by thread but ~55 valid children: only those are
// compared).  Four independent accumulators: the loads are broadcasts, the compares independent, so a lone warp
// issues them back to back instead of walking one dependent chain.
__device__ __forceinline__ uint32_t count_larger_keys(const uint32_t *keys, int n, uint32_t key)
{
    uint32_t r0 = 0u, r1 = 0u, r2 = 0u, r3 = 0u;
#pragma unroll 1
    for (int j = 0; j < n; j += 16) {
        const uint4 *k4 = reinterpret_cast<const uint4 *>(keys + j);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 k = k4[g];
            // keys are below 2^31, so the sign of key - k is "k is larger": a subtract and a shift-add
            // (LEA.HI) per key instead of compare / increment / select
            r0 += (key - k.x) >> 31; r1 += (key - k.y) >> 31; r2 += (key - k.z) >> 31; r3 += (key - k.w) >> 31;
        }
    }
    return (r0 + r1) + (r2 + r3);
}
// #scores among score[0 .. n) larger than `mine`; score[n .. n + 7] must hold -inf (chunks of 8)
__device__ __forceinline__ int count_larger_scores(const double *score, int n, double mine)
{
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    for (int j = 0; j < n; j += 8) {
        const double2 a = *reinterpret_cast<const double2 *>(score + j), b = *reinterpret_cast<const double2 *>(score + j + 2);
        const double2 c = *reinterpret_cast<const double2 *>(score + j + 4), d = *reinterpret_cast<const double2 *>(score + j + 6);
        r0 += (a.x > mine ? 1 : 0) + (c.x > mine ? 1 : 0); r1 += (a.y > mine ? 1 : 0) + (c.y > mine ? 1 : 0);
        r2 += (b.x > mine ? 1 : 0) + (d.x > mine ? 1 : 0); r3 += (b.y > mine ? 1 : 0) + (d.y > mine ? 1 : 0);
    }
    return (r0 + r1) + (r2 + r3);
}

// All 128 threads of the team call this with the same root / parameters; `bar` is the team's
// named barrier.  Returns the same BeamResult on every thread.
__device__ __forceinline__ BeamResult beam_search_team(Board root, int legal_given, const BeamParams &P, uint32_t game,
                                                       uint32_t call, const uint16_t *row, TeamScratch &ts, int bar)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tw = (threadIdx.x >> 5) & 3u;                        // warp in the team = action it expands
    const uint32_t tid = tw * 32u + lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
    BeamResult res;
    res.nodes = 0;
    res.best = 0.0;

    // agent:82-93 -- legality and the two fast exits (every warp computes them, no exchange needed)
    const Board root_child = agent_child_any(root, lane & 3u, row);
    const uint32_t agent_mask = __ballot_sync(FULL, root_child != root) & 15u;
    const uint32_t vm = legal_given >= 0 ? ((uint32_t)legal_given & 15u) : agent_mask;
    if (vm == 0u) { res.action = 0u; res.prob = 0.5f; return res; }
    if ((vm & (vm - 1u)) == 0u) { res.action = (uint32_t)__ffs((int)vm) - 1u; res.prob = 1.0f; return res; }

    // agent:96-106
    const uint32_t root_emax = max_exponent(root);
    const int root_max = tile_value(root_emax);
    const int phase = root_max < P.early_thr ? 0 : root_max < P.mid_thr ? 1 : 2;
    const int n0 = count_empty(root);
    const int depth = max(1, n0 <= 4 ? min(P.depth + 5, 25) : n0 >= 10 ? min(P.depth - 5, 10) : P.depth);

    int nb = 1;                  // parents of the level (level 0: the root)
    uint32_t spawn_base = 0u;    // spawns drawn so far in this call
    uint32_t ring_end = 0u;      // ts.rng holds the spawn words of ordinals [ring_end - 512, ring_end)
    const uint32_t *corners = corner_table();

    // A level draws at most 128 spawns; one pass (a Philox block per thread) adds 256 ordinals and overwrites
    // ordinals below ring_end - 256, all consumed (spawn_base > ring_end - 128).  The first pass is made here, the
    // later ones inside phase C of the level before the one that needs them: the block's ten rounds of multiplies
    // then interleave with the key loads of the counting rank instead of standing alone at the head of a level.
    {
        const Philox4 p = philox4x32_10(tid, call, game, DOM_BEAM, P.K);
        ts.rng[tid] = make_uint4(p.w[0], p.w[1], p.w[2], p.w[3]);
        ring_end = 256u;
    }
    TEAM_PROF_DECL;
    for (int d = 0; d < depth; ++d) {
        // ---- A: my child (agent:112-123 for the root, agent:142-167 below it) ------------------------
        Board parent = root;
        uint32_t pmeta = tw | (root_emax << 2);
        bool has_parent = lane == 0u && ((vm >> tw) & 1u);
        if (d > 0) {
            has_parent = (int)lane < nb;
            parent = Board(ts.beam[lane]);
            pmeta = ts.meta[lane];
        }
        Board b = agent_child_uniform(parent, tw, row);
        const bool valid = has_parent && b != parent;
        uint32_t nzl = nz_flags(b.lo), nzh = nz_flags(b.hi);
        const uint32_t zl = nzl ^ LSB4, zh = nzh ^ LSB4;
        const int cl = __popc(zl);
        int n_empty = cl + __popc(zh);
        const bool draws = valid && n_empty > 0;                        // agent:262-263: no draw on a full board
        const uint32_t vb = __ballot_sync(FULL, valid), db = __ballot_sync(FULL, draws);
        if (lane == 0u) { ts.ballots[tw] = vb; ts.ballots[4u + tw] = db; }
        team_barrier(bar);                                              // (1) ballots (and the ring) are visible
        TEAM_PROF_MARK(_pa);
        const uint4 V = *reinterpret_cast<const uint4 *>(&ts.ballots[0]);
        const uint4 D = *reinterpret_cast<const uint4 *>(&ts.ballots[4]);
        const int n_valid = (__popc(V.x) + __popc(V.y)) + (__popc(V.z) + __popc(V.w));
        if (n_valid == 0) {
            team_barrier(bar);                                          // everyone has read the ballots before a next call rewrites them
            if (d > 0) break;                                           // agent:170-171 keeps the old beam
            // agent:126-128: random.choice among the caller's valid moves, one draw
            const Philox4 p = philox4x32_10(0u, call, game, DOM_BEAM, P.K);
            const int pick = (int)__umulhi(p.w[0], (uint32_t)__popc(vm));
            uint32_t m = vm;
            for (int i = 0; i < pick; ++i) m &= m - 1u;
            res.action = (uint32_t)__ffs((int)m) - 1u;
            res.prob = 0.5f;
            return res;
        }
        res.nodes += n_valid;
        // children of lower ranks come first, then the lower actions of my own parent
        const uint32_t own = (tw > 0u ? (V.x >> lane) & 1u : 0u) + (tw > 1u ? (V.y >> lane) & 1u : 0u) +
                             (tw > 2u ? (V.z >> lane) & 1u : 0u);
        const uint32_t pos = (uint32_t)((__popc(V.x & lt_mask) + __popc(V.y & lt_mask)) +
                                        (__popc(V.z & lt_mask) + __popc(V.w & lt_mask))) + own;
        const uint32_t down = (tw > 0u ? (D.x >> lane) & 1u : 0u) + (tw > 1u ? (D.y >> lane) & 1u : 0u) +
                              (tw > 2u ? (D.z >> lane) & 1u : 0u);
        const uint32_t ordinal = spawn_base + (uint32_t)((__popc(D.x & lt_mask) + __popc(D.y & lt_mask)) +
                                                         (__popc(D.z & lt_mask) + __popc(D.w & lt_mask))) + down;
        spawn_base += (uint32_t)((__popc(D.x) + __popc(D.y)) + (__popc(D.z) + __popc(D.w)));

        // ---- B: spawn + evaluate (agent:155-161) ---------------------------------------------------------
        G2048_ASSERT(!valid || (pos < (uint32_t)n_valid && n_valid <= kTeamThreads));
        G2048_ASSERT(!draws || (ordinal < ring_end && ring_end - ordinal <= kTeamRing));      // the word is in the ring
        const uint2 w = reinterpret_cast<const uint2 *>(ts.rng)[ordinal & (kTeamRing - 1u)];
        const SpawnPick sp = pick_spawn(zl, zh, cl, n_empty, w.x, w.y);
        if (draws) {
            b.lo |= sp.flag_lo * sp.exponent;
            b.hi |= sp.flag_hi * sp.exponent;
            nzl |= sp.flag_lo;
            nzh |= sp.flag_hi;
            n_empty -= 1;
        }
        const uint32_t pmax = pmeta >> 2;                               // parent's largest exponent
        const uint32_t emax = pmax + ((pmax < 15u && has_exponent(b, pmax + 1u)) ? 1u : 0u);
        const uint32_t first = pmeta & 3u;
        const uint32_t tail = ((127u - pos) << 2) | first;
        const bool full_level = d >= 1 && d <= 3;                       // agent:139,158-161
        uint32_t key = 0u;
        double full = 0.0;
        if (full_level) {
            // float64 scores: rank = #candidates with a strictly larger score; equal scores are
            // separated by the generation index in the key (Python's stable sort)
            full = full_eval_flags(b, nzl, nzh, n_empty, emax, phase, corners);
            if (valid) ts.score[pos] = full;
            if (tid < 8u) ts.score[n_valid + (int)tid] = -INFINITY;       // padding of the last chunk of 8
            team_barrier(bar);
            const int larger = count_larger_scores(ts.score, n_valid, full);
            if (valid) key = ((uint32_t)(n_valid - larger) << 9) | tail;
        } else {
            const int fast = fast_eval_flags(b, nzl, nzh, n_empty, emax, corners);
            if (valid) key = ((uint32_t)fast << 9) | tail;
        }
        if (valid) ts.key[pos] = key;                                   // generation order, no gaps
        if (tid >= (uint32_t)n_valid && tid < (((uint32_t)n_valid + 15u) & ~15u)) ts.key[tid] = 0u;   // the last chunk's padding
        team_barrier(bar);                                              // (2) all keys of the level are visible
        TEAM_PROF_MARK(_pb);
        if (ring_end - spawn_base < 128u) {                             // the next level's spawn words (see above)
            const Philox4 p = philox4x32_10((ring_end >> 1) + tid, call, game, DOM_BEAM, P.K);
            ts.rng[((ring_end & (kTeamRing - 1u)) >> 1) + tid] = make_uint4(p.w[0], p.w[1], p.w[2], p.w[3]);
            ring_end += 256u;
        }

        // ---- C: stable top-k by counting (agent:131-132,174-175) -----------------------------------------
        const uint32_t rank = count_larger_keys(ts.key, n_valid, key);
        nb = min(P.width, n_valid);
        G2048_ASSERT(nb >= 1 && nb <= 32 && (!valid || rank < (uint32_t)n_valid));
        if (valid && (int)rank < nb) {
            ts.beam[rank] = b.u64();
            ts.meta[rank] = (uint8_t)(first | (emax << 2));
            if (rank == 0u) {
                ts.out_first = first;
                ts.out_best = full_level ? full : (double)(key >> 9);
            }
        }
        team_barrier(bar);                                              // (3) the next level's parents are in place
        TEAM_PROF_MARK(_pc);
#ifdef G2048_TEAM_PROFILE
        ++_pl;
#endif
    }
    TEAM_PROF_FLUSH();
    res.action = ts.out_first;                                          // agent:178
    res.best = ts.out_best;
    res.prob = 1.0f;
    return res;
}

}  // namespace g2048
