// board.cuh -- device primitives on packed 2048 boards (sm_100a).
//
// A board is 16 nibble exponents in a uint64 (cell (r,c) = nibble 4r+c, 0 = empty),
// held in registers as two 32-bit halves: lo = rows 0,1, hi = rows 2,3.  Everything
// here is branch-free integer SWAR sized for the ALU pipe (LOP3 / SHF / PRMT / IADD3 /
// POPC); the only memory operation is the row-table lookup of move_left().
#pragma once
#include <stdint.h>
#ifndef G2048_HOST_EMUL   // tests/host_emul compiles this header with g++ and intrinsic stand-ins
#include <cuda_runtime.h>
#endif

// -DG2048_DEBUG (make ../libg2048_debug.so): device-side asserts on every scratch / queue index.  The product
// build compiles them away; the GPU suite is run against the debug library once per round.
#if defined(G2048_DEBUG) && !defined(G2048_HOST_EMUL)
#include <cassert>
#define G2048_ASSERT(cond) assert(cond)
#else
#define G2048_ASSERT(cond) ((void)0)
#endif

namespace g2048 {

struct Board {
    uint32_t lo, hi;
    __device__ __forceinline__ Board() {}
    __device__ __forceinline__ Board(uint32_t l, uint32_t h) : lo(l), hi(h) {}
    __device__ __forceinline__ explicit Board(uint64_t b) : lo((uint32_t)b), hi((uint32_t)(b >> 32)) {}
    __device__ __forceinline__ uint64_t u64() const { return ((uint64_t)hi << 32) | lo; }
};
__device__ __forceinline__ bool operator==(const Board &a, const Board &b) { return ((a.lo ^ b.lo) | (a.hi ^ b.hi)) == 0; }
__device__ __forceinline__ bool operator!=(const Board &a, const Board &b) { return ((a.lo ^ b.lo) | (a.hi ^ b.hi)) != 0; }

constexpr uint32_t LSB4 = 0x11111111u;   // bit 0 of every nibble
constexpr uint32_t MSB4 = 0x88888888u;   // bit 3 of every nibble

template <int k> __device__ __forceinline__ uint32_t shr(uint32_t x) { return x >> k; }

// bit 4i set  <=>  nibble i of x is non-zero
__device__ __forceinline__ uint32_t nz_flags(uint32_t x)
{
    uint32_t t = x | shr<1>(x);
    t |= shr<2>(t);
    return t & LSB4;
}
// bit 4i set  <=>  nibble i of x is zero
__device__ __forceinline__ uint32_t zero_flags(uint32_t x) { return ~nz_flags(x) & LSB4; }

// (a & mask) | (b & ~mask) as ONE LOP3 (truth table 0xE2 with the mask as the middle, immediate
// operand).  Written in C the compiler re-associates nested selects into and/or chains with one
// distinct mask per term, which costs three LOP3 where two (transpose) or one (nibble swap) do.
__device__ __forceinline__ uint32_t bitselect(uint32_t mask, uint32_t a, uint32_t b)
{
#ifdef G2048_HOST_EMUL
    return (a & mask) | (b & ~mask);
#else
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE2;" : "=r"(d) : "r"(a), "r"(mask), "r"(b));
    return d;
#endif
}

// ---- geometry -------------------------------------------------------------
// 4x4 nibble transpose: swap inside 2x2 blocks (per half), then swap the two
// off-diagonal 2x2 blocks (a byte permutation across the halves).
__device__ __forceinline__ Board transpose(Board b)
{
    uint32_t l = bitselect(0x0000F0F0u, shr<12>(b.lo), bitselect(0x0F0F0000u, b.lo << 12, b.lo));
    uint32_t h = bitselect(0x0000F0F0u, shr<12>(b.hi), bitselect(0x0F0F0000u, b.hi << 12, b.hi));
    return Board(__byte_perm(l, h, 0x6240), __byte_perm(l, h, 0x7351));
}
__device__ __forceinline__ uint32_t swap_nibbles_in_bytes(uint32_t x)
{
    return bitselect(0xF0F0F0F0u, x << 4, shr<4>(x));
}
// np.fliplr: reverse the 4 nibbles of every row
__device__ __forceinline__ Board flip_rows(Board b)
{
    return Board(swap_nibbles_in_bytes(__byte_perm(b.lo, 0, 0x2301)),
                 swap_nibbles_in_bytes(__byte_perm(b.hi, 0, 0x2301)));
}
// np.flipud: reverse the order of the rows
__device__ __forceinline__ Board flip_row_order(Board b)
{
    return Board(__byte_perm(b.hi, 0, 0x1032), __byte_perm(b.lo, 0, 0x1032));
}
// np.rot90(b, 2): reverse all 16 nibbles
__device__ __forceinline__ Board rot180(Board b)
{
    return Board(swap_nibbles_in_bytes(__byte_perm(b.hi, 0, 0x0123)),
                 swap_nibbles_in_bytes(__byte_perm(b.lo, 0, 0x0123)));
}

// ---- row tables -------------------------------------------------------------
// row[r]  : the row r (4 nibbles) after a LEFT move (env:116-168 / agent:213-242)
// code[r] : the (at most two) merges of that move, one nibble each: 0 = none, else
//           merged exponent - 1, i.e. the merge scored 2 << nibble.
struct RowTables {
    const uint16_t *row;
    const uint8_t *code;
};

template <bool kShared>
__device__ __forceinline__ uint32_t lut16(const uint16_t *t, uint32_t i)
{
    if (kShared) return t[i];
    return __ldg(t + i);
}
template <bool kShared>
__device__ __forceinline__ uint32_t lut8(const uint8_t *t, uint32_t i)
{
    if (kShared) return t[i];
    return __ldg(t + i);
}

// LEFT move of all four rows.
template <bool kShared>
__device__ __forceinline__ Board move_left(Board b, const uint16_t *row)
{
    uint32_t r0 = lut16<kShared>(row, b.lo & 0xFFFFu);
    uint32_t r1 = lut16<kShared>(row, b.lo >> 16);
    uint32_t r2 = lut16<kShared>(row, b.hi & 0xFFFFu);
    uint32_t r3 = lut16<kShared>(row, b.hi >> 16);
    return Board(__byte_perm(r0, r1, 0x5410), __byte_perm(r2, r3, 0x5410));
}
// merge codes of the same move, one byte per row
template <bool kShared>
__device__ __forceinline__ uint32_t merge_codes(Board b, const uint8_t *code)
{
    uint32_t c0 = lut8<kShared>(code, b.lo & 0xFFFFu);
    uint32_t c1 = lut8<kShared>(code, b.lo >> 16);
    uint32_t c2 = lut8<kShared>(code, b.hi & 0xFFFFu);
    uint32_t c3 = lut8<kShared>(code, b.hi >> 16);
    return __byte_perm(__byte_perm(c0, c1, 0x0040), __byte_perm(c2, c3, 0x0040), 0x5410);
}
// Sum of the merged tile values encoded by 8 code nibbles; *sat gets bit 16 set when a
// merge produced 2^16 (nibble saturation, counted by the caller).
__device__ __forceinline__ uint32_t decode_score(uint32_t codes, uint32_t *sat)
{
    uint32_t total = 0, any = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t v = (2u << ((codes >> (4 * i)) & 15u)) & ~3u;   // code 0 -> 2 -> masked to 0
        total += v;
        any |= v;
    }
    *sat = any;
    return total;
}

// 2^c for a 4-bit field c >= 1 and 0 for c == 0, summed over selected nibbles on the FMA pipe:
// the field is shifted into the float exponent (c << 23 is the normal number 2^(c-127)) and
// scaled by 2^127.  All terms and partial sums are small powers of two / integers < 2^24, so
// the float arithmetic is exact.  `cells` has bit 4i set for every nibble i to include.
__device__ __forceinline__ float pow2_field(uint32_t x, int i)
{
    const uint32_t field = 15u << 23;
    uint32_t bits = (4 * i <= 23 ? (x << (23 - 4 * i)) : (x >> (4 * i - 23))) & field;
    return __uint_as_float(bits);
}
__device__ __forceinline__ float pow2_sum(uint32_t x, uint32_t cells, float acc)
{
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if ((cells >> (4 * i)) & 1u) acc = fmaf(pow2_field(x, i), 0x1p127f, acc);
    return acc;
}
// ---- pair tables of the fused rollout (512 x u32, shared memory) ------------------------------
// Entry b < 256: sum of the two tile values whose exponents are the nibbles of byte b (0 = empty).
// Entry 256 + c: score of the (up to two) merges a code byte c describes (2 << nibble each), plus
// kPairSaturated when a nibble is 15, i.e. a merge produced 2^16.  One lookup replaces two
// shift/mask/FMA extractions on the ALU pipe, which is the pipe the rollout is bound by.
constexpr int kPairEntries = 512;
constexpr uint32_t kPairSaturated = 1u << 28;          // above any sum of four row scores (< 2^20)
__host__ __device__ __forceinline__ uint32_t pair_table_entry(uint32_t i)
{
    const uint32_t lo = i & 15u, hi = (i >> 4) & 15u;
    if (i < 256u) return (lo ? 1u << lo : 0u) + (hi ? 1u << hi : 0u);
    return (lo ? 2u << lo : 0u) + (hi ? 2u << hi : 0u) + ((lo == 15u || hi == 15u) ? kPairSaturated : 0u);
}
// pairs[offset4 / 4] for a byte offset that is already a multiple of 4 (saves the index scaling)
__device__ __forceinline__ uint32_t pair_at(const uint32_t *pairs, uint32_t offset4)
{
    return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(pairs) + offset4);
}
// Score of a LEFT move of all four rows; bits 28.. count the rows with a saturated merge.
template <bool kShared>
__device__ __forceinline__ uint32_t merge_score_pairs(Board b, const uint8_t *code, const uint32_t *pairs)
{
    uint32_t c0 = lut8<kShared>(code, b.lo & 0xFFFFu);
    uint32_t c1 = lut8<kShared>(code, b.lo >> 16);
    uint32_t c2 = lut8<kShared>(code, b.hi & 0xFFFFu);
    uint32_t c3 = lut8<kShared>(code, b.hi >> 16);
    const uint32_t *m = pairs + 256;
    return (m[c0] + m[c1]) + (m[c2] + m[c3]);
}
// edge_sum of env:254-259 = total - inner 2x2 + corners (a corner is counted by a row and a column)
__device__ __forceinline__ uint32_t edge_sum_pairs(Board b, uint32_t total, const uint32_t *pairs)
{
    uint32_t inner = pair_at(pairs, (b.lo >> 18) & 0x3FCu) + pair_at(pairs, (b.hi >> 2) & 0x3FCu);   // cells 5,6 and 9,10
    // cells 15 and 0 are neighbours in the rotated 64-bit word; cells 3 and 12 after a byte pick
    uint32_t c015 = __funnelshift_l(b.hi, b.lo, 6) & 0x3FCu;
    uint32_t c312 = (__byte_perm(b.lo, b.hi, 0x0061) >> 2) & 0x3FCu;
    uint32_t corners = pair_at(pairs, c015) + pair_at(pairs, c312);
    return total - inner + corners;
}

// ---- LEFT move without tables ------------------------------------------------------------------
// The per-step kernel makes ONE move per launch: staging 192 KiB of tables per block cannot pay and
// reading them through L1/L2 puts two dependent memory round trips (state, then table) on a path
// that is latency-bound end to end.  This is the same row semantics (env:116-168) as pure SWAR on a
// half board (two rows), both rows at once:
//   1. compress: nibble i moves left by s_i = #empty cells left of it, as two masked shifts
//      (s_i & 1 -> one nibble, s_i & 2 -> two; adjacent tiles share s_i, so nothing collides);
//   2. merge: e_i = (y_i == y_i+1 != 0) for i = 0..2; the scan takes a pair only if the pair before
//      it was not taken: m0 = e0, m1 = e1 & ~m0, m2 = e2 & ~m1; the left cell gains one (a merged
//      exponent of 16 saturates to 15, as in row_tables.h), the right cell empties;
//   3. close the holes: cells right of a hole move one nibble left.
// `codes` gets, on the cell that received a merge, the exponent before it (= merged exponent - 1,
// the code of row_tables.h): one pair-table lookup per byte then gives the score and the saturation flag.
struct HalfMove { uint32_t rows, codes, occupied; };    // occupied: nz_flags of the INPUT rows (a by-product)
__device__ __forceinline__ HalfMove move_left_half(uint32_t x)
{
    const uint32_t nz = nz_flags(x), z = nz ^ LSB4;
    const uint32_t c = (z << 4) & 0xFFF0FFF0u;                                // nibble i: cell i-1 is empty
    const uint32_t s = c + ((c << 4) & 0xFFF0FFF0u) + ((c << 8) & 0xFF00FF00u);   // empty cells left of i (0..3)
    const uint32_t m1 = (s & nz) * 15u, m2 = ((s >> 1) & nz) * 15u;           // full-nibble masks of the movers
    uint32_t y = (x & ~m1) | ((x & m1) >> 4);
    const uint32_t m2a = (m2 & ~m1) | ((m2 & m1) >> 4);                       // the second mask travels with its cells
    y = (y & ~m2a) | ((y & m2a) >> 8);
    // merge scan over the packed row
    const uint32_t d = y ^ (y >> 4);
    uint32_t e = zero_flags(d) & nz_flags(y) & 0x01110111u;                    // e_i on nibble i, i = 0..2
    e &= ~((e & 0x00010001u) << 4);                                           // m1 = e1 & ~m0
    e &= ~((e & 0x00100010u) << 4);                                           // m2 = e2 & ~m1
    const uint32_t em = e * 15u;
    HalfMove r;
    r.occupied = nz;
    r.codes = y & em;
    const uint32_t sat = r.codes & (r.codes >> 1) & (r.codes >> 2) & (r.codes >> 3) & LSB4;   // merging two 32768s
    y = (y + (e ^ sat)) & ~(em << 4);
    const uint32_t mv = (((e & 0x00110011u) << 8) | ((e & 0x00010001u) << 12)) * 15u;
    r.rows = (y & ~mv) | ((y & mv) >> 4);
    return r;
}
// Sum over the 8 bytes of the two code words of pairs[256 + byte]: merge score + saturation flags
__device__ __forceinline__ uint32_t code_score_pairs(uint32_t codes_lo, uint32_t codes_hi, const uint32_t *pairs)
{
    const uint32_t *m = pairs + 256;
    return ((m[codes_lo & 0xFFu] + m[(codes_lo >> 8) & 0xFFu]) + (m[(codes_lo >> 16) & 0xFFu] + m[codes_lo >> 24])) +
           ((m[codes_hi & 0xFFu] + m[(codes_hi >> 8) & 0xFFu]) + (m[(codes_hi >> 16) & 0xFFu] + m[codes_hi >> 24]));
}
// Sum of the tile values of a board from the pair table (8 lookups)
__device__ __forceinline__ uint32_t tile_total_pairs(Board b, const uint32_t *pairs)
{
    return ((pair_at(pairs, (b.lo << 2) & 0x3FCu) + pair_at(pairs, (b.lo >> 6) & 0x3FCu)) +
            (pair_at(pairs, (b.lo >> 14) & 0x3FCu) + pair_at(pairs, (b.lo >> 22) & 0x3FCu))) +
           ((pair_at(pairs, (b.hi << 2) & 0x3FCu) + pair_at(pairs, (b.hi >> 6) & 0x3FCu)) +
            (pair_at(pairs, (b.hi >> 14) & 0x3FCu) + pair_at(pairs, (b.hi >> 22) & 0x3FCu)));
}

// Direction wrappers.  `to_line` brings the rows the tiles travel along into LEFT-move
// position, `from_line` undoes it.  Lane-varying actions use selects, not branches.
__device__ __forceinline__ Board select(bool p, Board a, Board b) { return Board(p ? a.lo : b.lo, p ? a.hi : b.hi); }

__device__ __forceinline__ Board to_line(Board b, uint32_t action)
{
    Board t = select(action & 1u, transpose(b), b);        // UP/DOWN travel along columns
    return select(action & 2u, flip_rows(t), t);           // RIGHT/DOWN travel towards index 3
}
__device__ __forceinline__ Board from_line(Board m, uint32_t action)
{
    Board t = select(action & 2u, flip_rows(m), m);
    return select(action & 1u, transpose(t), t);
}

// env._execute_move (env:97-114).  action outside 0..3 leaves the board alone.
template <bool kShared>
__device__ __forceinline__ Board env_move(Board b, uint32_t action, const uint16_t *row)
{
    Board m = from_line(move_left<kShared>(to_line(b, action), row), action);
    return select(action < 4u, m, b);
}

// ---- legality without tables ------------------------------------------------
// Bit a set <=> env action a changes the board (env:69-95): some tile can slide into an
// empty neighbour in that direction or two equal neighbours along it can merge.
__device__ __forceinline__ uint32_t eq_flags(uint32_t a, uint32_t b) { return zero_flags(a ^ b); }

// nl / nh: occupancy flags (bit 0 of every non-empty nibble) of b.lo / b.hi
__device__ __forceinline__ uint32_t env_legal_mask_flags(Board b, uint32_t nl, uint32_t nh)
{
    // horizontal pairs (c, c+1): flag sits on nibble c, c = 0..2 of each row
    const uint32_t HP = 0x01110111u;
    uint32_t nl1 = nl >> 4, nh1 = nh >> 4;                       // occupancy of the right neighbour
    uint32_t mh = ((eq_flags(b.lo, b.lo >> 4) & nl) | (eq_flags(b.hi, b.hi >> 4) & nh)) & HP;
    uint32_t sl = ((~nl & nl1) | (~nh & nh1)) & HP;              // empty cell, tile to its right  -> LEFT slides
    uint32_t sr = ((nl & ~nl1) | (nh & ~nh1)) & HP;              // tile, empty cell to its right  -> RIGHT slides
    // vertical pairs (r, r+1): flag sits on row r, r = 0..2
    uint32_t below_lo = __funnelshift_r(b.lo, b.hi, 16);         // rows 1,2 aligned under rows 0,1
    uint32_t below_hi = b.hi >> 16;                              // row 3 aligned under row 2
    uint32_t nbl = __funnelshift_r(nl, nh, 16), nbh = nh >> 16;
    const uint32_t VH = 0x00001111u;                             // only row 2 has a row below it in `hi`
    uint32_t mv = (eq_flags(b.lo, below_lo) & nl) | (eq_flags(b.hi, below_hi) & nh & VH);
    uint32_t su = (~nl & nbl) | (~nh & nbh & VH);                // empty cell, tile below -> UP slides
    uint32_t sd = (nl & ~nbl & LSB4) | (nh & ~nbh & VH);         // tile, empty cell below -> DOWN slides
    uint32_t mask = 0;
    mask |= ((mh | sl) != 0u) ? 1u : 0u;
    mask |= ((mv | su) != 0u) ? 2u : 0u;
    mask |= ((mh | sr) != 0u) ? 4u : 0u;
    mask |= ((mv | sd) != 0u) ? 8u : 0u;
    return mask;
}
__device__ __forceinline__ uint32_t env_legal_mask(Board b) { return env_legal_mask_flags(b, nz_flags(b.lo), nz_flags(b.hi)); }
// is_game_over (env:279-288): no empty cell and no equal neighbours
__device__ __forceinline__ bool env_game_over(Board b) { return env_legal_mask(b) == 0u; }
// Out-of-line copy for hot loops: a full board is rare, and a call cannot be if-converted

// ---- counting ----------------------------------------------------------------
__device__ __forceinline__ int count_empty(Board b) { return __popc(zero_flags(b.lo)) + __popc(zero_flags(b.hi)); }

// largest nibble of the board
__device__ __forceinline__ uint32_t max_exponent(Board b)
{
    const uint32_t M = 0x000F000Fu;
    uint32_t m0 = __vmaxu2(b.lo & M, (b.lo >> 4) & M);
    uint32_t m1 = __vmaxu2((b.lo >> 8) & M, (b.lo >> 12) & M);
    uint32_t m2 = __vmaxu2(b.hi & M, (b.hi >> 4) & M);
    uint32_t m3 = __vmaxu2((b.hi >> 8) & M, (b.hi >> 12) & M);
    uint32_t m = __vmaxu2(__vmaxu2(m0, m1), __vmaxu2(m2, m3));
    return max(m & 0xFFFFu, m >> 16);
}
// does any nibble of the board equal v (1..15)?
__device__ __forceinline__ bool has_exponent(Board b, uint32_t v)
{
    uint32_t rep = v * LSB4, x = b.lo ^ rep, y = b.hi ^ rep;
    // "some nibble is zero" by the borrow trick (exact as an any-test)
    return ((((x - LSB4) & ~x) | ((y - LSB4) & ~y)) & MSB4) != 0u;
}

// ---- Philox4x32-10 ---------------------------------------------------------------
struct Philox4 { uint32_t w[4]; };

// The ten round keys of a seed (key schedule k + r * Weyl constant).  Built once on the host and
// passed by value in the kernel arguments, so every round reads its keys as constant-bank
// operands instead of spending two integer adds per round per thread.
struct PhiloxKey {
    uint32_t k0[10], k1[10];
};
inline PhiloxKey make_philox_key(uint64_t seed)
{
    PhiloxKey K;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { K.k0[r] = a; K.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    return K;
}

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const PhiloxKey &K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

enum : uint32_t { DOM_ENV = 0, DOM_BEAM = 1, DOM_ACTION = 2, DOM_BOARD = 3, DOM_HYBRID = 4 };

struct SpawnWords { uint32_t pos, val; };
// spawn i of stream (seed, game, call, domain): words 2(i&1), 2(i&1)+1 of block i>>1
__device__ __forceinline__ SpawnWords spawn_words(const PhiloxKey &K, uint32_t game, uint32_t call,
                                                  uint32_t domain, uint32_t i)
{
    Philox4 p = philox4x32_10(i >> 1, call, game, domain, K);
    SpawnWords s;
    s.pos = (i & 1u) ? p.w[2] : p.w[0];
    s.val = (i & 1u) ? p.w[3] : p.w[1];
    return s;
}
__device__ __forceinline__ uint32_t random_action(const PhiloxKey &K, uint32_t game, uint32_t t)
{
    Philox4 p = philox4x32_10(t >> 6, 0u, game, DOM_ACTION, K);
    uint32_t s = (t >> 4) & 3u;
    uint32_t w = s == 0 ? p.w[0] : s == 1 ? p.w[1] : s == 2 ? p.w[2] : p.w[3];
    return (w >> (2u * (t & 15u))) & 3u;
}

// ---- spawn (env:59-67, agent:260-269) ----------------------------------------------
// Puts exponent 1 (word < 0.9*2^32) or 2 into the k-th empty cell in row-major order,
// k = (pos_word * n_empty) >> 32.  No-op on a full board.  Returns n_empty before.
__device__ __forceinline__ int place_tile(Board &b, uint32_t pos_word, uint32_t val_word)
{
    uint32_t zl = zero_flags(b.lo), zh = zero_flags(b.hi);
    int cl = __popc(zl), n = cl + __popc(zh);
    uint32_t k = __umulhi(pos_word, (uint32_t)n);
    bool in_hi = k >= (uint32_t)cl;
    uint32_t kk = in_hi ? k - (uint32_t)cl : k;
    uint32_t z = in_hi ? zh : zl;
    uint32_t prefix = z * LSB4;                       // nibble j = #empty among nibbles 0..j (<= 8)
    uint32_t s = prefix + (7u - kk) * LSB4;           // bit 3 of nibble j set <=> prefix_j > kk
    uint32_t bit = __ffs((int)(s & MSB4)) - 1;        // 4j+3 of the first such nibble
    uint32_t tile = (val_word < 3865470567u ? 1u : 2u) << ((bit - 3u) & 31u);
    tile = n > 0 ? tile : 0u;
    b.lo |= in_hi ? 0u : tile;
    b.hi |= in_hi ? tile : 0u;
    return n;
}

// The same spawn for callers that keep OCCUPANCY flags (bit 0 of every non-empty nibble) and
// want them updated: the lowest set bit of the comparison word is isolated with v & -v instead of
// ffs, which directly gives the new tile's occupancy flag; the tile is flag * exponent (FMA pipe).
// zl/zh = zero flags, cl = popc(zl), n = number of empty cells (a full board gives flag 0).
struct SpawnPick { uint32_t flag_lo, flag_hi, exponent; };
__device__ __forceinline__ SpawnPick pick_spawn(uint32_t zl, uint32_t zh, int cl, int n, uint32_t pos_word,
                                                uint32_t val_word)
{
    const uint32_t k = __umulhi(pos_word, (uint32_t)max(n, 1));
    const bool in_hi = k >= (uint32_t)cl;
    const uint32_t kk = in_hi ? k - (uint32_t)cl : k;
    const uint32_t z = in_hi ? zh : zl;
    const uint32_t v = (z * LSB4 + (7u - kk) * LSB4) & MSB4;              // bit 3 of nibble j set <=> #empty(0..j) > kk
    const uint32_t flag = (v & (0u - v)) >> 3;                              // bit 0 of the first such nibble
    SpawnPick sp;
    sp.flag_lo = in_hi ? 0u : flag;
    sp.flag_hi = in_hi ? flag : 0u;
    sp.exponent = val_word < 3865470567u ? 1u : 2u;
    return sp;
}

// ---- heuristics -----------------------------------------------------------------------
// Sum over a set of cells of 2^e (0 for an empty cell); `cells` has bit 4i set for cell i.
__device__ __forceinline__ uint32_t tile_sum_half(uint32_t x, uint32_t cells)
{
    uint32_t total = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if ((cells >> (4 * i)) & 1u) total += (1u << ((x >> (4 * i)) & 15u)) & ~1u;
    return total;
}

// #equal non-empty neighbour pairs (agent:302-312) and the sum of their exponents (agent:387-403)
__device__ __forceinline__ void merge_pairs_flags(Board b, uint32_t nl, uint32_t nh, int *pairs, int *exponent_sum);
__device__ __forceinline__ void merge_pairs(Board b, int *pairs, int *exponent_sum)
{
    merge_pairs_flags(b, nz_flags(b.lo), nz_flags(b.hi), pairs, exponent_sum);
}
// nl / nh: occupancy flags (bit 0 of every non-empty nibble) of b.lo / b.hi
__device__ __forceinline__ void merge_pairs_flags(Board b, uint32_t nl, uint32_t nh, int *pairs, int *exponent_sum)
{
    // a == b per nibble, exactly, on bit 3: adding 7 to the low three bits of a ^ b carries into bit 3
    // iff they are non-zero (one LOP3 less per word than smearing, and the add can go to the FMA pipe)
    auto eq8 = [](uint32_t a, uint32_t c) { uint32_t x = a ^ c; return ~(((x & 0x77777777u) + 0x77777777u) | x) & MSB4; };
    const uint32_t nl8 = nl * 8u, nh8 = nh * 8u;                            // occupancy on bit 3
    uint32_t hl = eq8(b.lo, b.lo >> 4) & nl8 & 0x08880888u;
    uint32_t hh = eq8(b.hi, b.hi >> 4) & nh8 & 0x08880888u;
    uint32_t vl = eq8(b.lo, __funnelshift_r(b.lo, b.hi, 16)) & nl8;
    uint32_t vh = eq8(b.hi, b.hi >> 16) & nh8 & 0x00008888u;
    *pairs = (__popc(hl) + __popc(hh)) + (__popc(vl) + __popc(vh));
    if (exponent_sum) {
        // keep the exponent of the first cell of every pair, then add all nibbles up
        hl >>= 3; hh >>= 3; vl >>= 3; vh >>= 3;
        uint32_t a = b.lo & (hl * 15u), c = b.hi & (hh * 15u), d = b.lo & (vl * 15u), e = b.hi & (vh * 15u);
        // nibble sums: split even/odd nibbles into bytes, bytes never overflow (<= 4 * 15)
        uint32_t ev = (a & 0x0F0F0F0Fu) + (c & 0x0F0F0F0Fu) + (d & 0x0F0F0F0Fu) + (e & 0x0F0F0F0Fu);
        uint32_t od = ((a >> 4) & 0x0F0F0F0Fu) + ((c >> 4) & 0x0F0F0F0Fu) + ((d >> 4) & 0x0F0F0F0Fu) + ((e >> 4) & 0x0F0F0F0Fu);
        *exponent_sum = (int)__vsadu4(ev, 0u) + (int)__vsadu4(od, 0u);
    }
}

// Corner table of the beam kernels (256 x u32, shared memory), indexed by a byte holding two corner
// exponents: bits 24.. = the larger exponent, bits 0..19 = 2 * its tile value (0 for two empty cells).
// Both fields grow with the exponent, so the larger entry of the two corner pairs describes the best corner.
__host__ __device__ __forceinline__ uint32_t corner_table_entry(uint32_t i)
{
    const uint32_t lo = i & 15u, hi = (i >> 4) & 15u, m = lo > hi ? lo : hi;
    return (m << 24) | (m ? 2u << m : 0u);
}
__device__ __forceinline__ uint32_t best_corner(Board b, const uint32_t *corners)
{
    // cells 15 and 0 are neighbours in the rotated 64-bit word; cells 3 and 12 after a byte pick
    const uint32_t c015 = __funnelshift_l(b.hi, b.lo, 6) & 0x3FCu;
    const uint32_t c312 = (__byte_perm(b.lo, b.hi, 0x0061) >> 2) & 0x3FCu;
    return max(pair_at(corners, c015), pair_at(corners, c312));
}

// BeamSearchAgent._fast_evaluate (agent:280-314): always an exact integer.
__device__ __forceinline__ int fast_eval_flags(Board b, uint32_t nl, uint32_t nh, int n_empty, uint32_t emax,
                                               const uint32_t *corners = nullptr)
{
    int corner_score;                                             // 2 * tile value of the best corner
    if (corners) {
        corner_score = (int)(best_corner(b, corners) & 0xFFFFFu);
    } else {
        uint32_t corner = max(max(b.lo & 15u, (b.lo >> 12) & 15u), max((b.hi >> 16) & 15u, b.hi >> 28));
        corner_score = corner ? (int)(2u << corner) : 0;
    }
    int pairs;
    merge_pairs_flags(b, nl, nh, &pairs, nullptr);
    return n_empty * 10 + (int)emax * 2 + corner_score + pairs * 2;
}
__device__ __forceinline__ int fast_eval(Board b, int n_empty, uint32_t emax)
{
    return fast_eval_flags(b, nz_flags(b.lo), nz_flags(b.hi), n_empty, emax);
}

// BeamSearchAgent._evaluate_state (agent:316-373) in float64 with the reference's operation
// order and no FMA contraction: bit-exact.  phase 0 early, 1 mid, 2 late.
__device__ __forceinline__ double full_eval_flags(Board b, uint32_t nl, uint32_t nh, int n_empty, uint32_t emax, int phase,
                                                  const uint32_t *corners = nullptr);
__device__ __forceinline__ double full_eval(Board b, int n_empty, uint32_t emax, int phase)
{
    return full_eval_flags(b, nz_flags(b.lo), nz_flags(b.hi), n_empty, emax, phase);
}
__device__ __forceinline__ double full_eval_flags(Board b, uint32_t nl, uint32_t nh, int n_empty, uint32_t emax, int phase,
                                                  const uint32_t *corners)
{
    const double w_empty  = phase == 0 ? 15.0 : phase == 1 ? 10.0 : 8.0;
    const double w_max    = phase == 0 ? 1.0  : phase == 1 ? 1.5  : 2.0;
    const double w_corner = phase == 0 ? 2.0  : phase == 1 ? 2.5  : 3.0;
    const double w_merge  = phase == 0 ? 2.0  : phase == 1 ? 1.5  : 1.0;
    double empty_score = __dmul_rn((double)n_empty, w_empty);
    if (n_empty <= 2) empty_score = __dadd_rn(empty_score, -10.0);
    double max_score = __dmul_rn((double)emax, w_max);
    if (emax >= 9)  max_score = __dmul_rn(max_score, 1.2);
    if (emax >= 10) max_score = __dmul_rn(max_score, 1.5);
    if (emax >= 11) max_score = __dmul_rn(max_score, 2.0);
    uint32_t corner = corners ? best_corner(b, corners) >> 24
                              : max(max(b.lo & 15u, (b.lo >> 12) & 15u), max((b.hi >> 16) & 15u, b.hi >> 28));
    double corner_bonus = __dmul_rn(__dmul_rn((double)corner, 2.0), w_corner);
    int pairs, esum;
    merge_pairs_flags(b, nl, nh, &pairs, &esum);
    double merge_potential = __dmul_rn((double)esum, w_merge);
    // snake weights 15 14 13 12 / 8 9 10 11 / 7 6 5 4 / 0 1 2 3 (agent:37-42), exact in int
    int snake = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int wl = i < 4 ? 15 - i : 4 + i;          // rows 0,1
        const int wh = i < 4 ? 7 - i : i - 4;           // rows 2,3
        snake += (int)((b.lo >> (4 * i)) & 15u) * wl + (int)((b.hi >> (4 * i)) & 15u) * wh;
    }
    double snake_score = __ddiv_rn((double)snake, 100.0);
    return __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(empty_score, max_score), corner_bonus), merge_potential), snake_score);
}

// ---- shaped reward (env:212-277), float64, fixed operation order, no FMA ------------------
// #(both non-empty and next >= previous) per row / per column, packed one nibble per line.
__device__ __forceinline__ uint32_t ge_flags(uint32_t next, uint32_t prev)
{
    // per-nibble unsigned next >= prev, result on bit 3 of every nibble
    uint32_t t = (next | MSB4) - (prev & ~MSB4);
    return ((next & ~prev) | (~(next ^ prev) & t)) & MSB4;
}

__device__ __forceinline__ void ordered_pairs(Board b, int line[4])
{
    // line[i] = row_ordered_i + col_ordered_i of env:267-275; a pair's flag sits on bit 3 of
    // the nibble of its FIRST cell (left cell of a row pair, upper cell of a column pair).
    const uint32_t nl = nz_flags(b.lo) << 3, nh = nz_flags(b.hi) << 3;      // occupancy on bit 3
    uint32_t hl = ge_flags(b.lo >> 4, b.lo) & nl & (nl >> 4) & 0x08880888u;  // rows 0,1
    uint32_t hh = ge_flags(b.hi >> 4, b.hi) & nh & (nh >> 4) & 0x08880888u;  // rows 2,3
    uint32_t below_lo = __funnelshift_r(b.lo, b.hi, 16), nbl = __funnelshift_r(nl, nh, 16);
    uint32_t vl = ge_flags(below_lo, b.lo) & nl & nbl;                       // row pairs (0,1), (1,2)
    uint32_t vh = ge_flags(b.hi >> 16, b.hi) & nh & (nh >> 16) & 0x00008888u; // row pair (2,3)
    uint32_t v = vl | (vh << 1);              // (2,3) flags parked on bit 0 of the next nibble
    line[0] = __popc(hl & 0x0000FFFFu) + __popc(v & 0x00080018u);
    line[1] = __popc(hl & 0xFFFF0000u) + __popc(v & 0x00800180u);
    line[2] = __popc(hh & 0x0000FFFFu) + __popc(v & 0x08001800u);
    line[3] = __popc(hh & 0xFFFF0000u) + __popc(v & 0x80018000u);
}

// ordered_pairs() with the occupancy flags (bit 0 of every non-empty nibble) already known
__device__ __forceinline__ void ordered_pairs_flags(Board b, uint32_t nzl, uint32_t nzh, int line[4])
{
    // "both cells non-empty and next >= previous" == "previous non-empty and next >= previous"
    // (next >= previous >= 1 already makes next non-empty), so only the first cell's occupancy masks
    const uint32_t nl = nzl << 3, nh = nzh << 3;
    uint32_t hl = ge_flags(b.lo >> 4, b.lo) & nl & 0x08880888u;
    uint32_t hh = ge_flags(b.hi >> 4, b.hi) & nh & 0x08880888u;
    uint32_t below_lo = __funnelshift_r(b.lo, b.hi, 16);
    uint32_t vl = ge_flags(below_lo, b.lo) & nl;
    uint32_t vh = ge_flags(b.hi >> 16, b.hi) & nh & 0x00008888u;
    // all 24 flags in one word: rows 0,1 on bit 3 of their nibbles, rows 2,3 on bit 2, column pairs
    // (0,1),(1,2) on bit 1 and (2,3) on bit 0; one popc per line instead of two
    uint32_t x = hl | (hh >> 1) | (vl >> 2) | (vh >> 3);
    line[0] = __popc(x & 0x0002088Bu);
    line[1] = __popc(x & 0x08A80030u);
    line[2] = __popc(x & 0x02000744u);
    line[3] = __popc(x & 0x24443000u);
}

// shaped_reward() for callers that track the tile total (moves conserve it, a spawn adds its
// value) and the occupancy flags; identical float64 operation sequence.
__device__ __forceinline__ double shaped_reward_tracked(bool valid, int empty_before, Board cur, int empty_after,
                                                        uint32_t nzl, uint32_t nzh, uint32_t score_delta,
                                                        uint32_t highest_exp_before, uint32_t prev_max_exp,
                                                        uint32_t total, const uint32_t *pairs)
{
    // env:225-251: score/4, the "new highest tile" bonuses, -2 for an invalid move and
    // 0.5 per empty cell gained are all multiples of 0.25 far below 2^53, so every partial sum of the
    // reference's float64 sequence is exact and the four terms can be added as integers (in quarters)
    int quarters = (int)score_delta + 2 * (empty_after - empty_before) - (valid ? 0 : 8);
    if (highest_exp_before > prev_max_exp) {
        quarters += 8 * (int)highest_exp_before;
        if (highest_exp_before >= 8)  quarters += 4 * 50;
        if (highest_exp_before >= 9)  quarters += 4 * 100;
        if (highest_exp_before >= 10) quarters += 4 * 200;
        if (highest_exp_before >= 11) quarters += 4 * 500;
    }
    double reward = __dmul_rn((double)quarters, 0.25);
    uint32_t edge = edge_sum_pairs(cur, total, pairs);
    reward = __dadd_rn(reward, __ddiv_rn((double)edge, (double)total));      // `* 1.0` (env:259) is the identity in IEEE-754
    if (empty_after <= 2) reward = __dadd_rn(reward, -2.0);
    int line[4];
    ordered_pairs_flags(cur, nzl, nzh, line);
#pragma unroll
    for (int i = 0; i < 4; ++i) reward = __dadd_rn(reward, __dmul_rn((double)line[i], 0.1));
    return reward;
}

__device__ __forceinline__ double shaped_reward(bool valid, int empty_before, Board cur, int empty_after,
                                                uint32_t score_delta, uint32_t highest_exp_before,
                                                uint32_t prev_max_exp)
{
    double reward = __dmul_rn((double)score_delta, 0.25);                  // / 4.0 (exact either way)
    if (highest_exp_before > prev_max_exp) {                               // env:229-241 (dead inside step)
        reward = __dadd_rn(reward, __dmul_rn(2.0, (double)highest_exp_before));
        if (highest_exp_before >= 8)  reward = __dadd_rn(reward, 50.0);
        if (highest_exp_before >= 9)  reward = __dadd_rn(reward, 100.0);
        if (highest_exp_before >= 10) reward = __dadd_rn(reward, 200.0);
        if (highest_exp_before >= 11) reward = __dadd_rn(reward, 500.0);
    }
    if (!valid) reward = __dadd_rn(reward, -2.0);
    reward = __dadd_rn(reward, __dmul_rn((double)(empty_after - empty_before), 0.5));
    // edge_sum = rows 0,3 + columns 0,3 (corners twice) = total - inner 2x2 + corners
    uint32_t total = tile_sum_half(cur.lo, LSB4) + tile_sum_half(cur.hi, LSB4);
    uint32_t inner = tile_sum_half(cur.lo, 0x01100000u) + tile_sum_half(cur.hi, 0x00000110u);
    uint32_t corners = tile_sum_half(cur.lo, 0x00001001u) + tile_sum_half(cur.hi, 0x10010000u);
    uint32_t edge = total - inner + corners;
    reward = __dadd_rn(reward, __ddiv_rn((double)edge, (double)total));      // `* 1.0` (env:259) is the identity in IEEE-754
    if (empty_after <= 2) reward = __dadd_rn(reward, -2.0);
    int line[4];
    ordered_pairs(cur, line);
#pragma unroll
    for (int i = 0; i < 4; ++i) reward = __dadd_rn(reward, __dmul_rn((double)line[i], 0.1));
    return reward;
}

// ---- PPO-side features (agents/ppo_agent.py), SURVEY 8f row 1 -----------------------------------
// evaluate_heuristic (ppo_agent.py:271-333): 2 * best-direction monotonicity / 24
// + 1 if the largest corner holds the largest tile - 0.1 * #(tiles >= 8); float64, same order.
__device__ __forceinline__ double ppo_heuristic(Board b)
{
    const uint32_t nl = nz_flags(b.lo) << 3, nh = nz_flags(b.hi) << 3;       // occupancy on bit 3
    const uint32_t HP = 0x08880888u;
    // horizontal pairs (left, right), flag on the left cell
    uint32_t hml = nl & (nl >> 4) & HP, hmh = nh & (nh >> 4) & HP;
    int h_le = __popc(ge_flags(b.lo >> 4, b.lo) & hml) + __popc(ge_flags(b.hi >> 4, b.hi) & hmh);
    int h_ge = __popc(ge_flags(b.lo, b.lo >> 4) & hml) + __popc(ge_flags(b.hi, b.hi >> 4) & hmh);
    // vertical pairs (upper, lower), flag on the upper cell
    uint32_t below_lo = __funnelshift_r(b.lo, b.hi, 16), below_hi = b.hi >> 16;
    uint32_t vml = nl & __funnelshift_r(nl, nh, 16), vmh = nh & (nh >> 16) & 0x00008888u;
    int v_le = __popc(ge_flags(below_lo, b.lo) & vml) + __popc(ge_flags(below_hi, b.hi) & vmh);
    int v_ge = __popc(ge_flags(b.lo, below_lo) & vml) + __popc(ge_flags(b.hi, below_hi) & vmh);
    int best = max(h_le, h_ge) + max(v_le, v_ge);         // max over the four (row_dir, col_dir) pairs
    double score = __dmul_rn(2.0, __ddiv_rn((double)best, 24.0));
    uint32_t corner = max(max(b.lo & 15u, (b.lo >> 12) & 15u), max((b.hi >> 16) & 15u, b.hi >> 28));
    if (corner == max_exponent(b)) score = __dadd_rn(score, 1.0);
    // tiles >= 8  <=>  exponent >= 3
    int high = __popc(ge_flags(b.lo, 0x33333333u)) + __popc(ge_flags(b.hi, 0x33333333u));
    if (high > 0) score = __dadd_rn(score, __dmul_rn(-0.1, (double)high));
    return score;
}

// 0.1 * sum(log2 of the four largest tiles) (ppo_agent.py:251-254): exponents are summed exactly
__device__ __forceinline__ double ppo_top4_bonus(Board b)
{
    // counting sort over the 16 possible exponents: cnt[e] packed 5 bits each would overflow a
    // word, so walk the exponents from the top and take what is still needed
    int need = 4, sum = 0;
#pragma unroll
    for (int e = 15; e >= 1; --e) {
        uint32_t rep = (uint32_t)e * LSB4;
        int c = __popc(zero_flags(b.lo ^ rep)) + __popc(zero_flags(b.hi ^ rep));
        int take = min(c, need);
        sum += take * e;
        need -= take;
    }
    return __dmul_rn(0.1, (double)sum);
}

}  // namespace g2048
