// ecg_core.cuh -- one match-3 board, register resident, stepped by one thread.
//
// Restates (bit-exactly, see tests/) the reference hot path
//   match3tile/boardFunctions.py : legal_actions :26-112, swap :115, get_matches :121-156,
//                                  get_match_spawn_mask :159-169, shuffle :16-23
//   match3tile/boardv2.py        : apply_action :43-207
// on a bit-sliced board: 4 code bit-planes p[0..3] over the padded row-major bitboard of
// ecg_bits.cuh.  Cell codes: 0 empty, 1..11 plain token of that type, 12 h_line, 13 v_line,
// 14 bomb, 15 mega (specials are typeless in the reference: value & type_mask == 0).
//
// __host__ __device__ throughout: nvcc builds the sm_100a kernels from it, g++ builds the
// test-only host simulator (tests/hostsim) that is fuzzed against the CPU oracle.
#pragma once
#include "ecg_bits.cuh"

namespace ecg {

enum : uint32_t {
    ST_TERMINAL = 1,        // moves_left < 1: board returned unchanged (boardv2.py:44)
    ST_STREAM_OVERFLOW = 2, // replay stream exhausted
    ST_SHUFFLE_CAP = 4,     // shuffle loop capped (the reference would spin forever)
    ST_BAD_ACTION = 8,      // action outside [0, action_space) (reference: KeyError, boardv2.py:48)
    ST_NO_LEGAL = 16,       // random pick on an empty legal set (reference: ValueError)
    ST_BAD_CELL = 32,       // pack: cell value outside the engine's closed code set
    ST_CASCADE_CAP = 64,    // cascade loop capped (the reference has no cap; tiny type counts never settle)
};
constexpr int SHUFFLE_CAP = 64;
constexpr int CASCADE_CAP = 1024;

enum : int { K_HLINE = 0, K_VLINE = 1, K_BOMB = 2, K_MEGA = 3 };

// host experiments only (-DECG_COUNT_RARE): why the common-case build hands boards off
#if defined(ECG_COUNT_RARE) && !defined(__CUDA_ARCH__)
extern long long g_rare_reason[8];
#define ECG_RARE(i) (g_rare_reason[i]++)
#else
#define ECG_RARE(i) ((void)0)
#endif

// ------------------------------------------------------------------ RNG sources

// Philox4x32-10 (Salmon et al. SC'11).  One substream per (board, step):
// u32 #k of the substream = philox(ctr = (k>>2, step, board_lo, board_hi), key)[k & 3].
ECG_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0;
        c1 = l1;
        c2 = n2;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Out-of-line copy for the rare block fetches inside loops (a refill of more than 12 tiles, shuffles): values in,
// values out, so the callers' state stays in registers and the loops carry no inlined Philox rounds.
struct PhiloxBlock {
    uint32_t w[4];
};
template <int UNUSED = 0> // a template only so that the header can define it
ECG_HD_NOINLINE PhiloxBlock philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    PhiloxBlock b;
    philox4x32_10(c0, c1, c2, c3, k0, k1, b.w);
    return b;
}

// Sixteen refill tiles at once: the four base-n digits of each of the four words (most significant first), each
// stored as (digit + 1) in a 4-bit field, first tile in the low nibble of lo.  Fields are never 0, so "lo == 0"
// means "all used".  Out of line: only a refill of more than 12 tiles in one cascade iteration gets here.
struct TileNibbles {
    uint32_t lo, hi;
};
ECG_HD void pack_tiles(uint32_t x, uint32_t n, uint32_t &out, int shift) { // four digits of x -> out[shift .. shift+15]
#pragma unroll
    for (int t = 0; t < 4; t++) {
        out += (mulhi32(x, n) + 1u) << (shift + 4 * t);
        x *= n;
    }
}
template <int UNUSED = 0>
ECG_HD_NOINLINE TileNibbles philox_block_tiles(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t n) {
    uint32_t w[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, w);
    TileNibbles t = {0u, 0u};
    pack_tiles(w[0], n, t.lo, 0);
    pack_tiles(w[1], n, t.lo, 16);
    pack_tiles(w[2], n, t.hi, 0);
    pack_tiles(w[3], n, t.hi, 16);
    return t;
}

// Throughput mode: bounded ints by multiply-high of one u32 (no rejection).
// Draw addressing inside the (board, step) substream (engine-defined; mirrored by the oracle's Philox mode):
//   word 0: the random action pick of this step (philox_pick);
//   cascade iteration j (0-based): refill tiles come from words j*2048 + 1, + 2, ...: four base-n digits per
//   word, handed to the holes in ROW-MAJOR order (tiles are i.i.d., so the order is free; row-major = bit
//   order makes the loop convergent);
//   shuffle draws of that iteration start at word j*2048 + 1024.
struct PhiloxRng {
    static constexpr bool ROW_MAJOR = true;
    uint32_t k0, k1, b0, b1, step;
    uint32_t pos;
    uint32_t blk[4];
    uint32_t blk_idx;
    uint32_t dig_x, dig_left; // digit(): up to 4 base-n digits are taken from one word
    bool overflow;
    ECG_HD void init(uint64_t key, uint64_t board, uint32_t step_) {
        k0 = (uint32_t)key;
        k1 = (uint32_t)(key >> 32);
        b0 = (uint32_t)board;
        b1 = (uint32_t)(board >> 32);
        step = step_;
        pos = 0;
        blk_idx = 0xFFFFFFFFu;
        blk[0] = blk[1] = blk[2] = blk[3] = 0u;
        dig_x = dig_left = 0u;
        overflow = false;
    }
    ECG_HD void reseed() {} // counter-based: one substream per (board, step), never restarted
    ECG_HD void seek(uint32_t p) {
        pos = p;
        dig_left = 0u;
    }
    // hand over a block computed earlier (the step kernel computes one block per trip for all lanes)
    ECG_HD void preset_block(uint32_t index, const uint32_t w[4]) {
        blk[0] = w[0];
        blk[1] = w[1];
        blk[2] = w[2];
        blk[3] = w[3];
        blk_idx = index;
    }
    ECG_HD uint32_t u32() {
        const uint32_t k = pos++;
        const uint32_t b = k >> 2;
        if (b != blk_idx) {
            const PhiloxBlock nb = philox_block(b, step, b0, b1, k0, k1);
            blk[0] = nb.w[0];
            blk[1] = nb.w[1];
            blk[2] = nb.w[2];
            blk[3] = nb.w[3];
            blk_idx = b;
        }
        const uint32_t j = k & 3u;
        return j == 0 ? blk[0] : j == 1 ? blk[1] : j == 2 ? blk[2] : blk[3];
    }
    ECG_HD uint32_t below(uint32_t n) { return n <= 1u ? 0u : mulhi32(u32(), n); }
    // The refill's view of the same digit stream (see digit() below): after seek(p) with p = 1 mod 4 and the block
    // of p preset, first_tiles() packs the 12 tiles of words p, p+1, p+2 as nibbles (TileNibbles layout) and
    // more_tiles() the 16 of every following block.  One multiply per tile on the FMA pipe, done by all lanes of
    // the warp together, instead of digit bookkeeping inside the divergent per-hole loop.
    ECG_HD void first_tiles(uint32_t n, uint32_t &lo, uint32_t &hi) {
        if ((pos >> 2) != blk_idx) { // not preset by the caller
            const PhiloxBlock nb = philox_block(pos >> 2, step, b0, b1, k0, k1);
            blk[0] = nb.w[0];
            blk[1] = nb.w[1];
            blk[2] = nb.w[2];
            blk[3] = nb.w[3];
            blk_idx = pos >> 2;
        }
        lo = hi = 0u;
        pack_tiles(blk[1], n, lo, 0);
        pack_tiles(blk[2], n, lo, 16);
        pack_tiles(blk[3], n, hi, 0);
        pos = (pos | 3u) + 1u; // first word of the next block
    }
    ECG_HD void more_tiles(uint32_t n, uint32_t &lo, uint32_t &hi) {
        const TileNibbles t = philox_block_tiles(pos >> 2, step, b0, b1, k0, k1, n);
        lo = t.lo;
        hi = t.hi;
        pos += 4u;
    }
    // Refill tiles: successive base-n digits of the fraction word / 2^32 (digit = hi32(x * n), x = lo32(x * n)),
    // 4 per word.  Each digit is uniform up to n^4 / 2^32 (3e-7 for 6 types).
    ECG_HD uint32_t digit(uint32_t n) {
        if (dig_left == 0u) {
            dig_x = u32();
            dig_left = 4u;
        }
        dig_left--;
        const uint32_t v = mulhi32(dig_x, n);
        dig_x *= n;
        return v;
    }
};

// The action pick of a Philox lockstep step: idx = mulhi(word 0 of the (board, step) substream, n).
// blk receives block 0, whose words 1..3 are the first refill words of the step (cached by the step kernel).
ECG_HD uint32_t philox_pick(uint64_t key, uint64_t board, uint32_t step, uint32_t n, uint32_t blk[4]) {
    philox4x32_10(0u, step, (uint32_t)board, (uint32_t)(board >> 32), (uint32_t)key, (uint32_t)(key >> 32), blk);
    return n <= 1u ? 0u : mulhi32(blk[0], n);
}
ECG_HD uint32_t philox_pick(uint64_t key, uint64_t board, uint32_t step, uint32_t n) {
    uint32_t blk[4];
    return philox_pick(key, board, step, n, blk);
}

// Parity mode: replays the raw u32 output of numpy's legacy MT19937 (np.random.seed(cfg.seed)
// restarts it at the top of every apply_action, boardv2.py:46, and in shuffle, boardFunctions.py:17)
// with numpy's masked-rejection bounded integers (RandomState.randint / random_interval).
struct ReplayRng {
    static constexpr bool ROW_MAJOR = false; // the reference's order: columns left to right, top cell first
    const uint32_t *raw;
    uint32_t len, pos;
    bool overflow;
    ECG_HD void init(const uint32_t *raw_, uint32_t len_, uint32_t pos_) {
        raw = raw_;
        len = len_;
        pos = pos_;
        overflow = false;
    }
    ECG_HD void reseed() { pos = 0; }
    ECG_HD void seek(uint32_t) {} // strictly sequential like the reference
    ECG_HD uint32_t u32() {
        const uint32_t k = pos++;
        if (k >= len) {
            overflow = true;
            return 0u;
        }
        return raw[k];
    }
    ECG_HD uint32_t below(uint32_t n) {
        if (n <= 1u) return 0u;
        const uint32_t rng = n - 1u;
        uint32_t mask = rng;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        for (;;) {
            const uint32_t v = u32() & mask;
            if (v <= rng || overflow) return v <= rng ? v : 0u;
        }
    }
    ECG_HD uint32_t digit(uint32_t n) { return below(n); } // np.random.randint(1, types + 1) per tile
};

// Replay mode, common-case kernel.  Within one apply_action the refill tiles are a PREFIX of the stream's
// accepted-tile sequence: np.random.randint(1, types + 1) is masked rejection with one fixed mask over the raw words,
// restarted by np.random.seed(cfg.seed) at the top of the step (boardv2.py:46, :172) -- until a shuffle draws with
// another bound (boardFunctions.py:22), which the common-case build hands to the exact build anyway.  So the tiles
// are precomputed once per stream: tile j as a nibble (1..types, 8 per word, first tile in the low nibble) and
// wpos[j] = raw words consumed once j tiles are taken (REPLAY_TILES_END past the end of the stream); the refill then
// reads 16 tiles per window like the Philox path instead of one raw word per rejection round.
constexpr uint32_t REPLAY_TILES_END = 0xFFFFu;
ECG_HD_CONSTEXPR int replay_tile_words(int stream_len) { return (stream_len + 7) / 8 + 2; } // + 2: window over-read
ECG_HD void build_replay_tiles(const uint32_t *raw, int len, uint32_t types, uint32_t *tiles, uint16_t *wpos) {
    const int tw = replay_tile_words(len);
    for (int i = 0; i < tw; i++) tiles[i] = 0u;
    const uint32_t rng = types - 1u;
    uint32_t mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    int j = 0;
    wpos[0] = 0;
    if (rng == 0u) { // randint(1, 2): no word is drawn
        for (j = 1; j <= len; j++) {
            tiles[(j - 1) >> 3] |= 1u << (4 * ((j - 1) & 7));
            wpos[j] = 0;
        }
        return;
    }
    for (int k = 0; k < len; k++) {
        const uint32_t v = raw[k] & mask;
        if (v <= rng) {
            tiles[j >> 3] |= (v + 1u) << (4 * (j & 7));
            wpos[++j] = (uint16_t)(k + 1);
        }
    }
    for (int t = j + 1; t <= len; t++) wpos[t] = (uint16_t)REPLAY_TILES_END;
}

struct ReplayTileRng {
    static constexpr bool ROW_MAJOR = false; // the reference's order: columns left to right, top cell first
    static constexpr bool TILE_WINDOW = true;
    const uint32_t *tiles;
    const uint16_t *wpos;
    uint32_t tpos, len; // tiles taken since the last reseed; words (= table entries) of the stream
    bool overflow;
    ECG_HD void init(const uint32_t *tiles_, const uint16_t *wpos_, uint32_t len_, uint32_t tpos_) {
        tiles = tiles_;
        wpos = wpos_;
        len = len_;
        tpos = tpos_;
        overflow = false;
    }
    // after a refill: were all the tiles taken so far really there?  (a window may end inside the stream's last tiles)
    ECG_HD void check_end() {
        if (tpos > len || wpos[tpos > len ? len : tpos] == REPLAY_TILES_END) overflow = true;
    }
    ECG_HD void reseed() { tpos = 0; }
    ECG_HD void seek(uint32_t) {}
    // the 16 tiles from tile t on, one per nibble, first in the low nibble of lo (zero nibbles past the stream's
    // last tile: check_end() after the refill tells)
    ECG_HD void window(uint32_t t, uint32_t &lo, uint32_t &hi) {
        if (t >= len) { // the table holds at most len tiles
            overflow = true;
            lo = hi = 0x11111111u;
            return;
        }
        const uint32_t *p = tiles + (t >> 3);
        const int sh = (int)(t & 7u) * 4;
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
        lo = funnel_r(w0, w1, sh);
        hi = funnel_r(w1, w2, sh);
    }
    ECG_HD uint32_t words() const { return wpos[tpos]; } // raw words behind the tiles taken (np.random's position)
};
template <class RNG>
struct UsesTileWindow {
    static constexpr bool value = false;
};
template <>
struct UsesTileWindow<ReplayTileRng> {
    static constexpr bool value = true;
};

// ------------------------------------------------------------------ board

template <class G>
struct Board {
    BB<G::W> p[4];
};

// Static per-shape parameters.  TPL = token planes that can differ between plain tokens
// (3 when types <= 7, else 4); CLIP = np.clip(next_state, 0, 32) (boardv2.py:163) bites,
// i.e. type_mask == 15: spawned bombs (48) and megas (64) become 32 == v_line.
template <int R_, int C_, int TPL_, bool CLIP_>
struct Shape {
    using G = Geo<R_, C_>;
    static constexpr int TPL = TPL_;
    static constexpr bool CLIP = CLIP_;
};

template <class G>
struct Derived { // views of a board used by match / legal logic
    BB<G::W> spec; // special cells
    BB<G::W> nz;   // cells holding a plain token
    BB<G::W> q[4]; // token planes (zero on specials)
};

template <class SH>
ECG_PHASE Derived<typename SH::G> derive(const Board<typename SH::G> &b) {
    using G = typename SH::G;
    Derived<G> d;
    d.spec = b.p[3] & b.p[2];
    d.nz = andn(b.p[0] | b.p[1] | b.p[2] | b.p[3], d.spec);
#pragma unroll
    for (int k = 0; k < 4; k++) d.q[k] = andn(b.p[k], d.spec);
    return d;
}

// bit b: token(b) == token(b + K) and both are plain tokens
template <class SH, int K>
ECG_HD BB<SH::G::W> eq_at(const Derived<typename SH::G> &d) {
    constexpr int W = SH::G::W;
    BB<W> diff = d.q[0] ^ shr<K>(d.q[0]);
#pragma unroll
    for (int k = 1; k < SH::TPL; k++) diff |= d.q[k] ^ shr<K>(d.q[k]);
    return andn(d.nz, diff);
}

template <class G>
struct Matches {
    BB<G::W> mask;       // cells cleared by get_matches
    BB<G::W> sp;         // spawn cells (get_match_spawn_mask != 0)
    BB<G::W> sk0, sk1;   // spawn kind bits
    BB<G::W> A, D;       // eq-right / eq-down of the analysed board (reused by legal_mask)
    bool found;
    bool rare;           // find_matches<SH, true> only: a rare case was met and nothing was computed (see there)
};

template <class SH>
ECG_HD void add_spawn(Matches<typename SH::G> &m, const BB<SH::G::W> &cells, int kind) {
    if (SH::CLIP && kind >= K_BOMB) kind = K_VLINE; // np.clip(.., 0, 32): 48 and 64 -> 32 == v_line
    // a later group overwrites an earlier one at the same centre (boardFunctions.py:164-168)
    m.sk0 = andn(m.sk0, cells);
    m.sk1 = andn(m.sk1, cells);
    m.sp |= cells;
    if (kind & 1) m.sk0 |= cells;
    if (kind & 2) m.sk1 |= cells;
}

// spawns of groups whose centres cannot coincide (disjoint groups): plain OR
template <class SH>
ECG_HD void add_spawn_disjoint(Matches<typename SH::G> &m, const BB<SH::G::W> &cells, int kind) {
    if (SH::CLIP && kind >= K_BOMB) kind = K_VLINE;
    m.sp |= cells;
    if (kind & 1) m.sk0 |= cells;
    if (kind & 2) m.sk1 |= cells;
}

template <class G>
ECG_HD int run_right(const BB<G::W> &A, int b) { // cells in the A-chain starting at b (>= 1)
    return 1 + ctz32(~extract32(A, b));
}
template <class G>
ECG_HD int run_down(const BB<G::W> &D, int b) {
    int n = 1;
    while (testbit(D, b)) {
        n++;
        b += G::S;
    }
    return n;
}

// Straight-run spawns for runs of length >= 5 (mega tokens), rare.
template <class SH>
ECG_HD void mega_spawns(Matches<typename SH::G> &m, BB<SH::G::W> L5h, BB<SH::G::W> L5v) {
    using G = typename SH::G;
    while (any(L5h)) {
        const int b = pop_lowest(L5h);
        const int n = run_right<G>(m.A, b);
        add_spawn_disjoint<SH>(m, onehot<G::W>(b + n / 2), K_MEGA); // sorted(group)[len // 2], boardFunctions.py:13
    }
    while (any(L5v)) {
        const int b = pop_lowest(L5v);
        const int n = run_down<G>(m.D, b);
        add_spawn_disjoint<SH>(m, onehot<G::W>(b + (n / 2) * G::S), K_MEGA);
    }
}

// Origins that fire both a horizontal and a vertical run (L / T corner at the origin):
// group = H cells + V cells with the origin listed twice -> never a line -> bomb.
template <class SH>
ECG_HD void corner_spawns(Matches<typename SH::G> &m, BB<SH::G::W> K) {
    using G = typename SH::G;
    while (any(K)) {
        const int b = pop_lowest(K);
        const int hl = run_right<G>(m.A, b), vl = run_down<G>(m.D, b);
        const int idx = (hl + vl) / 2; // >= 3
        // sorted multiset: (r,c),(r,c),(r,c+1)..(r,c+hl-1),(r+1,c)..(r+vl-1,c)
        const int centre = idx <= hl ? b + idx - 1 : b + (idx - hl) * G::S;
        add_spawn_disjoint<SH>(m, onehot<G::W>(centre), K_BOMB);
    }
}

// General group bookkeeping when a horizontal run crosses a vertical run that started in
// an earlier row (boardFunctions.py:126-131 merges the later match into the FIRST earlier
// group sharing a cell, keeping duplicates).  Rare; kept out of line.
template <class SH>
ECG_HD void merged_spawns(Matches<typename SH::G> &m, const BB<SH::G::W> &HO, const BB<SH::G::W> &VO,
                                   const BB<SH::G::W> &VC) {
    using G = typename SH::G;
    constexpr int W = G::W, MAXM = G::R * G::C / 3 + 1;
    // record: origin bit (9) | hlen (5) << 9 | vlen (5) << 14 | gid (8) << 19
    uint32_t rec[MAXM];
    int nm = 0;
    BB<W> O = HO | VO;
    while (any(O)) {
        const int b = pop_lowest(O);
        const int r = b / G::S, c = b - r * G::S;
        const int hl = testbit(HO, b) ? run_right<G>(m.A, b) : 0;
        const int vl = testbit(VO, b) ? run_down<G>(m.D, b) : 0;
        int gid = nm;
        uint32_t cross = hl ? (extract32(VC, b) & ((1u << hl) - 1u)) : 0u; // H cells covered by an earlier vertical run
        while (cross) {
            const int x = ctz32(cross);
            cross &= cross - 1u;
            for (int j = 0; j < nm; j++) {
                const int ob = rec[j] & 511, ovl = (rec[j] >> 14) & 31, og = (int)(rec[j] >> 19);
                const int orow = ob / G::S, ocol = ob - orow * G::S;
                if (ovl && ocol == c + x && orow < r && r <= orow + ovl - 1) gid = og < gid ? og : gid;
            }
        }
        rec[nm++] = (uint32_t)b | ((uint32_t)hl << 9) | ((uint32_t)vl << 14) | ((uint32_t)gid << 19);
    }
    for (int g = 0; g < nm; g++) {
        if ((int)(rec[g] >> 19) != g) continue; // not a group creator
        BB<W> once = bb_zero<W>(), twice = bb_zero<W>(), thrice = bb_zero<W>();
        int n = 0, members = 0, hl0 = 0, vl0 = 0;
        for (int j = g; j < nm; j++) {
            if ((int)(rec[j] >> 19) != g) continue;
            const int b = rec[j] & 511, hl = (rec[j] >> 9) & 31, vl = (rec[j] >> 14) & 31;
            if (members == 0) {
                hl0 = hl;
                vl0 = vl;
            }
            members++;
            n += hl + vl;
            if (hl) { // the horizontal run is a contiguous bit range
                const BB<W> run = bitrange<W>(b, b + hl);
                thrice |= twice & run;
                twice |= once & run;
                once |= run;
            }
            if (vl) { // the vertical run: one column, rows r .. r+vl-1
                const int r = b / G::S, c = b - r * G::S;
                const BB<W> run = shl_rt(G::col0(), c) & bitrange<W>(r * G::S, (r + vl) * G::S);
                thrice |= twice & run;
                twice |= once & run;
                once |= run;
            }
        }
        if (n <= 3) continue; // boardFunctions.py:161
        int kind;
        if (members == 1 && vl0 == 0) kind = n > 4 ? K_MEGA : K_VLINE;      // one row   (:163-164)
        else if (members == 1 && hl0 == 0) kind = n > 4 ? K_MEGA : K_HLINE; // one column (:165-166)
        else kind = K_BOMB;                                                  // (:168)
        // centre = sorted(group)[n // 2] with duplicates (:8-13): skip whole words by popcount, then walk
        int idx = n / 2, centre = 0;
        bool found = false;
#pragma unroll
        for (int wi = 0; wi < W; wi++) {
            if (found) continue;
            const int cw = popc32(once.w[wi]) + popc32(twice.w[wi]) + popc32(thrice.w[wi]);
            if (idx >= cw) {
                idx -= cw;
                continue;
            }
            uint32_t o = once.w[wi];
            while (o) {
                const uint32_t bit = o & (0u - o);
                o ^= bit;
                const int mult = 1 + ((twice.w[wi] & bit) ? 1 : 0) + ((thrice.w[wi] & bit) ? 1 : 0);
                if (idx < mult) {
                    centre = 32 * wi + ctz32(bit);
                    found = true;
                    break;
                }
                idx -= mult;
            }
        }
        add_spawn<SH>(m, onehot<W>(centre), kind);
    }
}

// Exact scan-order semantics of get_matches when horizontal and vertical runs intersect
// (SURVEY.md 8a row A4): rows top to bottom, carrying vcov = cells of the row covered by
// vertical runs started above.  Per maximal equal segment [s,e] (len >= 3) the first column
// p in [s, e-2] not in vcov fires and marks [p, e]; vertical origins are the uncovered cells
// not in (p, e] with two equal cells below.  Carry-propagation finds p and [p, e] for all
// segments of a row at once.
template <class SH>
ECG_HD void scan_order_matches(Matches<typename SH::G> &m, const BB<SH::G::W> &hs, const BB<SH::G::W> &vs) {
    using G = typename SH::G;
    constexpr int W = G::W;
    constexpr uint32_t RM = (1u << G::C) - 1u;
    BB<W> HF = bb_zero<W>(), HO = bb_zero<W>(), VO = bb_zero<W>(), VC = bb_zero<W>();
    uint32_t vcov = 0;
    for (int r = 0; r < G::R; r++) { // rolled on purpose: rare path, keep the instruction footprint small
        const int b0 = r * G::S;
        const uint32_t hs_r = extract32(hs, b0) & RM, A_r = extract32(m.A, b0) & RM;
        const uint32_t vs_r = extract32(vs, b0) & RM, D_r = extract32(m.D, b0) & RM;
        const uint32_t cand = hs_r & ~vcov;
        const uint32_t first = cand & ~(A_r + cand);
        const uint32_t fill = (A_r + first) ^ A_r;
        const uint32_t vorig = vs_r & ~vcov & ~(fill & ~first);
        // deposit the row words (values < 2^16, shift < 32 within a 2-word window)
        const int wi = b0 >> 5, s = b0 & 31;
#pragma unroll
        for (int i = 0; i < W; i++) {
            if (i == wi) {
                HF.w[i] |= fill << s;
                HO.w[i] |= first << s;
                VO.w[i] |= vorig << s;
                VC.w[i] |= vcov << s;
            }
            if (i == wi + 1 && s) {
                HF.w[i] |= fill >> (32 - s);
                HO.w[i] |= first >> (32 - s);
                VO.w[i] |= vorig >> (32 - s);
                VC.w[i] |= vcov >> (32 - s);
            }
        }
        vcov = (vcov | vorig) & D_r;
    }
    m.mask = HF | VO | VC;
    if (any(HF & VC)) { // a fired horizontal run crosses an earlier vertical run: groups merge
        merged_spawns<SH>(m, HO, VO, VC);
        return;
    }
    // every match is its own group
    const BB<W> K = HO & VO;
    const BB<W> HOs = andn(HO, K), VOs = andn(VO, K);
    const BB<W> L4h = HOs & shr<2>(m.A), L4v = VOs & shr<2 * G::S>(m.D);
    const BB<W> L5h = L4h & shr<3>(m.A), L5v = L4v & shr<3 * G::S>(m.D);
    add_spawn_disjoint<SH>(m, shl<2>(andn(L4h, L5h)), K_VLINE);        // horizontal 4-run -> v_line (:164)
    add_spawn_disjoint<SH>(m, shl<2 * G::S>(andn(L4v, L5v)), K_HLINE); // vertical 4-run -> h_line (:166)
    if (any(L5h | L5v)) mega_spawns<SH>(m, L5h, L5v);
    if (any(K)) corner_spawns<SH>(m, K);
}

// One straight run as its own group (boardFunctions.py:161-166): start bit, length, step (1 = horizontal)
template <class SH>
ECG_HD void straight_spawn(Matches<typename SH::G> &m, int start, int n, int step) {
    if (n <= 3) return;
    const int kind = n > 4 ? K_MEGA : (step == 1 ? K_VLINE : K_HLINE);
    add_spawn_disjoint<SH>(m, onehot<SH::G::W>(start + (n / 2) * step), kind);
}

// Exactly ONE cell x belongs to both a horizontal and a vertical run (L, T and + shapes): the outcome of
// the reference's scan has a closed form.  H = [s, e] in row r, V = [t, b] in
// column c, x = (r, c); every other run of the board is disjoint from both and forms its own group.
//   t < r, c > s : V fires first from (t, c) and covers x; H fires from s and crosses it -> ONE merged group,
//                  x listed twice -> bomb at the median of the sorted multiset.
//   t < r, c == s: x is covered, so H can only fire from s + 1: its own group [s+1, e] if it still has 3 cells,
//                  else the two cells right of x stay on the board.
//   t == r, c == s: the origin fires both runs (corner): one group, origin listed twice -> bomb.
//   t == r, c > s : H fires from s and consumes x, so V can only start one row lower: its own group
//                  [r+1, b] if it still has 3 cells, else the two cells below x stay.
template <class SH>
ECG_HD void single_cross_matches(Matches<typename SH::G> &m, const BB<SH::G::W> &hs, const BB<SH::G::W> &vs,
                                          const BB<SH::G::W> &HV3, int x) {
    using G = typename SH::G;
    constexpr int W = G::W, S = G::S;
    const int r = x / S, c = x - r * S;
    // H = [s, e] in row r and V = [t, b] in column c from the dense row word of A and column word of D
    // (A is zero in the last column and the pad column, D in the last row: the chains end by themselves)
    const uint32_t Ar = extract32(m.A, r * S);
    const uint32_t Dc = column_bits<G::R, S>(m.D, c);
    const int e = c + ctz32(~(Ar >> c));
    const int s = bitlen32(~Ar & ((1u << c) - 1u)); // one past the highest A == 0 left of c
    const int b = r + ctz32(~(Dc >> r));
    const int t = bitlen32(~Dc & ((1u << r) - 1u));
    const int hl = e - s + 1, vl = b - t + 1;
    m.mask = HV3;
    // every run other than H and V is disjoint from both and forms its own group: the formulas of the
    // intersection-free case, with the run starts of H and V taken out; skipped when H and V are all there is
    if (popcount(HV3) != hl + vl - 1) {
        const BB<W> own = onehot<W>(r * S + s) | onehot<W>(t * S + c);
        const BB<W> L4h = andn(andn(hs, shl<1>(m.A)) & shr<2>(m.A), own);
        const BB<W> L4v = andn(andn(vs, shl<S>(m.D)) & shr<2 * S>(m.D), own);
        if (any(L4h | L4v)) {
            const BB<W> L5h = L4h & shr<3>(m.A), L5v = L4v & shr<3 * S>(m.D);
            add_spawn_disjoint<SH>(m, shl<2>(andn(L4h, L5h)), K_VLINE);
            add_spawn_disjoint<SH>(m, shl<2 * S>(andn(L4v, L5v)), K_HLINE);
            if (any(L5h | L5v)) mega_spawns<SH>(m, L5h, L5v);
        }
    }
    // groups are created in scan order; the pair's groups may be created between the others', but spawn
    // centres of disjoint groups never coincide, so the order of the add_spawn calls is irrelevant here.
    // Each case ends in at most one spawn (start cell, run length, step) or one bomb cell, and at most one
    // pair of cells (first, first + gap) that stays on the board.
    int bomb = -1, st0 = 0, stn = 0, stp = 1, keep = -1, gap = 1;
    if (t < r) {
        if (c != s) { // merged T / + shape
            const int a = r - t; // vertical cells above row r
            int idx = (hl + vl) / 2;
            if (idx < a) {
                bomb = (t + idx) * S + c;
            } else if (idx - a < hl + 1) { // row r holds s..c, c, c+1..e
                idx -= a;
                bomb = r * S + (idx <= c - s ? s + idx : s + idx - 1);
            } else {
                bomb = (r + 1 + (idx - a - (hl + 1))) * S + c;
            }
        } else { // V is its own group; H can only fire from s + 1
            st0 = t * S + c;
            stn = vl;
            stp = S;
            if (hl >= 4) { // two straight groups: the second one spawns here, the first one below
                straight_spawn<SH>(m, r * S + s + 1, hl - 1, 1);
            } else {
                keep = x + 1;
            }
        }
    } else {
        if (c == s) { // corner at the origin: (r,c),(r,c),(r,c+1)..(r,e),(r+1,c)..(b,c)
            const int idx = (hl + vl) / 2;
            bomb = idx <= hl ? x + idx - 1 : x + (idx - hl) * S;
        } else { // H is its own group; V can only start one row lower
            st0 = r * S + s;
            stn = hl;
            if (vl >= 4) {
                straight_spawn<SH>(m, (r + 1) * S + c, vl - 1, S);
            } else {
                keep = x + S;
                gap = S;
            }
        }
    }
    if (bomb >= 0) add_spawn_disjoint<SH>(m, onehot<W>(bomb), K_BOMB);
    straight_spawn<SH>(m, st0, stn, stp);
    if (keep >= 0) m.mask = andn(m.mask, onehot<W>(keep) | onehot<W>(keep + gap));
}

// Out-of-line entry points of the rarely taken match paths.  They take and return VALUES only: nothing of
// the caller's register-resident state (board, Matches) ever has its address taken, so the hot loop keeps it
// in registers and pays no local-memory traffic for the existence of these paths.
template <class G>
struct MatchOut {
    BB<G::W> mask, sp, sk0, sk1;
};
template <class SH>
ECG_HD_NOINLINE MatchOut<typename SH::G> crossing_matches(BB<SH::G::W> A, BB<SH::G::W> D, BB<SH::G::W> hs,
                                                          BB<SH::G::W> vs, BB<SH::G::W> HV3, BB<SH::G::W> X) {
    using G = typename SH::G;
    Matches<G> m;
    m.A = A;
    m.D = D;
    m.sp = bb_zero<G::W>();
    m.sk0 = bb_zero<G::W>();
    m.sk1 = bb_zero<G::W>();
    if (popcount(X) == 1) single_cross_matches<SH>(m, hs, vs, HV3, pop_lowest(X));
    else scan_order_matches<SH>(m, hs, vs);
    MatchOut<G> o;
    o.mask = m.mask;
    o.sp = m.sp;
    o.sk0 = m.sk0;
    o.sk1 = m.sk1;
    return o;
}
template <class SH>
ECG_HD_NOINLINE MatchOut<typename SH::G> long_run_spawns(BB<SH::G::W> A, BB<SH::G::W> D, BB<SH::G::W> L5h,
                                                         BB<SH::G::W> L5v) {
    using G = typename SH::G;
    Matches<G> m;
    m.A = A;
    m.D = D;
    m.sp = bb_zero<G::W>();
    m.sk0 = bb_zero<G::W>();
    m.sk1 = bb_zero<G::W>();
    mega_spawns<SH>(m, L5h, L5v);
    MatchOut<G> o;
    o.mask = bb_zero<G::W>();
    o.sp = m.sp;
    o.sk0 = m.sk0;
    o.sk1 = m.sk1;
    return o;
}

// get_matches + get_match_spawn_mask of the token board (boardFunctions.py:121-169).
// FAST = true is the common-case build used by the first of the two step kernels (ecg_shape_kernels.cu): the two
// rare cases -- a horizontal and a vertical run share a cell (scan-order dependent result, 1.4 % of the calls), or
// a run of 6 or more -- are not computed; m.rare is set instead (m.found is true) and the caller hands the board
// over to the exact kernel.  Their code (2 000 instructions) then never enters the hot kernel's instruction stream.
template <class SH, bool FAST = false>
ECG_PHASE void find_matches(const Derived<typename SH::G> &d, Matches<typename SH::G> &m) {
    using G = typename SH::G;
    constexpr int W = G::W, S = G::S;
    m.A = eq_at<SH, 1>(d);
    m.D = eq_at<SH, S>(d);
    m.sp = bb_zero<W>();
    m.sk0 = bb_zero<W>();
    m.sk1 = bb_zero<W>();
    m.rare = false;
    const BB<W> hs = m.A & shr<1>(m.A); // b, b+1, b+2 equal
    const BB<W> vs = m.D & shr<S>(m.D);
    m.found = any(hs | vs);
    if (!m.found) {
        m.mask = bb_zero<W>();
        return;
    }
    const BB<W> H3 = hs | shl<1>(hs) | shl<2>(hs);
    const BB<W> V3 = vs | shl<S>(vs) | shl<2 * S>(vs);
    const BB<W> X = H3 & V3;
    const BB<W> L4h = andn(hs, shl<1>(m.A)) & shr<2>(m.A); // run starts with >= 4 cells
    const BB<W> L4v = andn(vs, shl<S>(m.D)) & shr<2 * S>(m.D);
    if (any(X)) { // intersecting runs: the reference's result depends on scan order
        if constexpr (FAST) {
            // Two thirds of these calls are ONE shared cell x on a board whose runs are all plain triples:
            // single_cross_matches with hl == vl == 3 and no other spawn collapses to four cases.
            //   x top of V and left end of H (corner)   : everything cleared, bomb on the right end of H
            //   x top of V, not left end (H fires first): the two cells below x stay
            //   x left end, not top (V fires first)     : the two cells right of x stay
            //   neither (merged T / + shape)            : everything cleared, bomb on x -- or on the cell left
            //                                             of x when x is the bottom of V and the right end of H
            // ONE shared cell with a run of four or more on the board is three quarters of the hand-offs.  On small
            // boards, which hand off ten times as often (6x6x4: 4.8 % of the steps against 0.55 % at 9x9x6), it is
            // resolved here by the exact build's out-of-line closed form: 4.85e9 -> 5.10e9 env-steps/s at 6x6x4;
            // at 9x9x6 the call costs the common-case kernel more than the exact kernel saves (7.19e9 -> 7.00e9, r04d).
#if defined(ECG_FAST_CROSS_CALL)
            constexpr bool CROSS_CALL = true;
#else
            constexpr bool CROSS_CALL = G::R <= 6;
#endif
            if (CROSS_CALL && popcount(X) == 1 && any(L4h | L4v)) {
                const MatchOut<G> o = crossing_matches<SH>(m.A, m.D, hs, vs, H3 | V3, X);
                m.mask = o.mask;
                m.sp = o.sp;
                m.sk0 = o.sk0;
                m.sk1 = o.sk1;
                return;
            }
            if (popcount(X) != 1 || any(L4h | L4v)) {
                ECG_RARE(popcount(X) != 1 ? 0 : 1);
#if defined(ECG_COUNT_RARE) && !defined(__CUDA_ARCH__)
                if (popcount(X) == 1) {
                    const BB<W> h4 = hs & shr<1>(hs), v4 = vs & shr<S>(vs);
                    const BB<W> c4 = h4 | shl<1>(h4) | shl<2>(h4) | shl<3>(h4) | v4 | shl<S>(v4) | shl<2 * S>(v4) | shl<3 * S>(v4);
                    ECG_RARE(any(X & c4) ? 5 : 6);
                }
#endif
                m.rare = true;
                m.mask = bb_zero<W>();
                return;
            }
            const bool top = any(andn(X, shl<S>(m.D))), left = any(andn(X, shl<1>(m.A)));
            BB<W> keep = bb_zero<W>(), bomb = bb_zero<W>();
            if (top && !left) keep = shl<S>(X) | shl<2 * S>(X);
            if (!top && left) keep = shl<1>(X) | shl<2>(X);
            if (top && left) bomb = shl<2>(X);
            if (!top && !left) bomb = any(X & (m.A | m.D)) ? X : shr<1>(X);
            m.mask = andn(H3 | V3, keep);
            add_spawn_disjoint<SH>(m, bomb, K_BOMB);
            return;
        }
        const MatchOut<G> o = crossing_matches<SH>(m.A, m.D, hs, vs, H3 | V3, X);
        m.mask = o.mask;
        m.sp = o.sp;
        m.sk0 = o.sk0;
        m.sk1 = o.sk1;
        return;
    }
    // disjoint straight runs: every maximal run is one group
    m.mask = H3 | V3;
    if (any(L4h | L4v)) {
        const BB<W> L5h = L4h & shr<3>(m.A), L5v = L4v & shr<3 * S>(m.D);
        add_spawn_disjoint<SH>(m, shl<2>(andn(L4h, L5h)), K_VLINE);
        add_spawn_disjoint<SH>(m, shl<2 * S>(andn(L4v, L5v)), K_HLINE);
        if (any(L5h | L5v)) { // runs of exactly 5: mega token on the third cell; longer ones out of line
            const BB<W> L6h = L5h & shr<4>(m.A), L6v = L5v & shr<4 * S>(m.D);
            add_spawn_disjoint<SH>(m, shl<2>(andn(L5h, L6h)) | shl<2 * S>(andn(L5v, L6v)), K_MEGA);
            if (any(L6h | L6v)) {
                if constexpr (FAST) {
                    ECG_RARE(2);
                    m.rare = true;
                    return;
                }
                const MatchOut<G> o = long_run_spawns<SH>(m.A, m.D, L6h, L6v);
                m.sp |= o.sp;
                m.sk0 |= o.sk0;
                m.sk1 |= o.sk1;
            }
        }
    }
}

// ------------------------------------------------------------------ legal mask

// legal_actions (boardFunctions.py:26-112) as two swap bitboards: HL bit x = swap (x, x+1),
// VL bit x = swap (x, x+S).  A swap is legal iff either token is 0 (special / empty, :100), or
// the tokens differ (:103) and a moved token completes a run of three with two equal tokens
// that are not its swap partner (:41-61, :74-94).  A/D must be eq-right/eq-down of this board.
template <class SH>
ECG_PHASE void legal_swaps(const Derived<typename SH::G> &d, const BB<SH::G::W> &A, const BB<SH::G::W> &D,
                        BB<SH::G::W> &HL, BB<SH::G::W> &VL) {
    using G = typename SH::G;
    constexpr int W = G::W, S = G::S;
    const BB<W> A2 = eq_at<SH, 2>(d) & G::notlastcol(); // tok(b) == tok(b+2); col C-1 would wrap into the next row
    const BB<W> D2 = eq_at<SH, 2 * S>(d);
    const BB<W> F = eq_at<SH, S + 1>(d); // down-right diagonal
    const BB<W> Gd = eq_at<SH, S - 1>(d); // down-left diagonal
    const BB<W> Gup = shl<S - 1>(Gd);    // tok(x) == tok(x - S + 1)
    const BB<W> Fup = shl<S>(F);         // tok(x + 1) == tok(x - S)
    const BB<W> G1 = shr<1>(Gd);         // tok(x + 1) == tok(x + S)
    const BB<W> F1 = shl<1>(F);          // tok(x - 1) == tok(x + S)
    const BB<W> Aup2 = shl<2>(A);
    const BB<W> Dup2 = shl<2 * S>(D);
    // horizontal swap (x, x+1)
    BB<W> h = (A2 & shr<2>(A))                 // x -> x+1 joins (x+2, x+3)
              | (Gup & F)                       // x -> x+1 between (x+1-S) and (x+1+S)
              | (Gup & shl<2 * S - 1>(D))       // x -> x+1 under (x+1-2S, x+1-S)
              | (F & shr<S + 1>(D))             // x -> x+1 above (x+1+S, x+1+2S)
              | (shl<1>(A2) & Aup2)             // x+1 -> x joins (x-2, x-1)
              | (Fup & G1)                      // x+1 -> x between (x-S) and (x+S)
              | (Fup & Dup2)                    // x+1 -> x under (x-2S, x-S)
              | (G1 & shr<S>(D));               // x+1 -> x above (x+S, x+2S)
    // vertical swap (x, x+S)
    BB<W> v = (D2 & shr<2 * S>(D))             // x -> x+S joins (x+2S, x+3S)
              | (Gd & F)                        // x -> x+S between (x+S-1) and (x+S+1)
              | (Gd & shr<S - 2>(A))            // x -> x+S right of (x+S-2, x+S-1)
              | (F & shr<S + 1>(A))             // x -> x+S left of (x+S+1, x+S+2)
              | (shl<S>(D2) & Dup2)             // x+S -> x joins (x-2S, x-S)
              | (F1 & G1)                       // x+S -> x between (x-1) and (x+1)
              | (F1 & Aup2)                     // x+S -> x right of (x-2, x-1)
              | (G1 & shr<1>(A));               // x+S -> x left of (x+1, x+2)
    const BB<W> zt = andn(G::valid(), d.nz);   // token == 0
    HL = (andn(h, A) | zt | shr<1>(zt)) & G::hvalid();
    VL = (andn(v, D) | zt | shr<S>(zt)) & G::vvalid();
}

// swap bitboards -> action-ordered mask, action = r*(2C-1) + c (horizontal) | + (C-1) + c (vertical)
// (boardConfig.py:45-69)
template <class G>
ECG_PHASE void swaps_to_actions(const BB<G::W> &HL, const BB<G::W> &VL, uint32_t out[G::AW]) {
#pragma unroll
    for (int i = 0; i < G::AW; i++) out[i] = 0u;
#pragma unroll
    for (int r = 0; r < G::R; r++) {
        constexpr uint32_t HM = (1u << (G::C - 1)) - 1u, VM = (1u << G::C) - 1u;
        const uint64_t row = (uint64_t)(extract32(HL, r * G::S) & HM) |
                             ((uint64_t)(extract32(VL, r * G::S) & VM) << (G::C - 1));
        const int pos = r * G::ROWA, wi = pos >> 5, s = pos & 31;
        const uint64_t sh = row << s; // <= 31 + 31 bits
#pragma unroll
        for (int i = 0; i < G::AW; i++) {
            if (i == wi) out[i] |= (uint32_t)sh;
            if (i == wi + 1) out[i] |= (uint32_t)(sh >> 32);
        }
    }
}

// legal_actions[k] (ascending ACTION order, what np.random.choice indexes) straight from the swap bitboards: board
// row r owns actions r*ROWA ..: its C-1 horizontal swaps, then its C vertical swaps (boardConfig.py:45-59).  One pass
// over the rows instead of building the action-ordered mask first.
template <class G>
ECG_HD int swaps_select_action(const BB<G::W> &HL, const BB<G::W> &VL, int k) {
    constexpr uint32_t HM = (1u << (G::C - 1)) - 1u, VM = (1u << G::C) - 1u;
    uint32_t w = 0;
    int base = 0;
    bool done = false;
#pragma unroll
    for (int r = 0; r < G::R; r++) {
        const uint32_t row = (extract32(HL, r * G::S) & HM) | ((extract32(VL, r * G::S) & VM) << (G::C - 1));
        const int c = popc32(row);
        if (!done) {
            if (k < c) {
                w = row;
                base = r * G::ROWA;
                done = true;
            } else {
                k -= c;
            }
        }
    }
    int pos = 0, c;
    c = popc32(w & 0xFFFFu);
    if (k >= c) { k -= c; pos += 16; w >>= 16; }
    c = popc32(w & 0xFFu);
    if (k >= c) { k -= c; pos += 8; w >>= 8; }
    c = popc32(w & 0xFu);
    if (k >= c) { k -= c; pos += 4; w >>= 4; }
    c = popc32(w & 0x3u);
    if (k >= c) { k -= c; pos += 2; w >>= 2; }
    if (k >= (int)(w & 1u)) pos += 1;
    return base + pos;
}

// The legal set in "swap-bitboard order": all horizontal swaps by (row, col), then all vertical swaps by
// (row, col).  This is the order of the packed legal mask in HBM (HL words, then VL words) and of the
// Philox-mode random pick; the reference's ascending-action order is produced by swaps_to_actions.
template <class G>
ECG_HD int swaps_count(const BB<G::W> &HL, const BB<G::W> &VL) {
    return popcount(HL) + popcount(VL);
}
// k-th legal swap (swap-bitboard order) -> bit of its source cell; vertical tells the direction
template <class G>
ECG_HD int swaps_select_bit(const BB<G::W> &HL, const BB<G::W> &VL, int k, bool &vertical) {
    uint32_t w = 0;
    int base = 0;
    bool done = false;
#pragma unroll
    for (int i = 0; i < 2 * G::W; i++) {
        const uint32_t x = i < G::W ? HL.w[i < G::W ? i : 0] : VL.w[i < G::W ? 0 : i - G::W];
        const int c = popc32(x);
        if (!done) {
            if (k < c) {
                w = x;
                base = 32 * i;
                done = true;
            } else {
                k -= c;
            }
        }
    }
    int pos = 0, c;
    c = popc32(w & 0xFFFFu);
    if (k >= c) { k -= c; pos += 16; w >>= 16; }
    c = popc32(w & 0xFFu);
    if (k >= c) { k -= c; pos += 8; w >>= 8; }
    c = popc32(w & 0xFu);
    if (k >= c) { k -= c; pos += 4; w >>= 4; }
    c = popc32(w & 0x3u);
    if (k >= c) { k -= c; pos += 2; w >>= 2; }
    if (k >= (int)(w & 1u)) pos += 1;
    int bit = base + pos;
    vertical = bit >= 32 * G::W;
    if (vertical) bit -= 32 * G::W;
    return bit;
}
template <class G>
ECG_HD int action_of_swap(int bit, bool vertical) { // boardConfig.py:61-69 (encode)
    const int r = bit / G::S, col = bit - r * G::S;
    return r * G::ROWA + col + (vertical ? G::C - 1 : 0);
}
template <class G>
ECG_HD int swaps_select(const BB<G::W> &HL, const BB<G::W> &VL, int k) { // k-th legal swap -> action id
    bool vertical;
    const int bit = swaps_select_bit<G>(HL, VL, k, vertical);
    return action_of_swap<G>(bit, vertical);
}

// ------------------------------------------------------------------ step

template <class G>
ECG_HD void decode_action(int a, int &b1, int &b2) { // boardConfig.py:45-59 (columns >= 4)
    const int r = a / G::ROWA, k = a - r * G::ROWA;
    if (k < G::C - 1) {
        b1 = r * G::S + k;
        b2 = b1 + 1;
    } else {
        b1 = r * G::S + (k - (G::C - 1));
        b2 = b1 + G::S;
    }
}

template <class G>
ECG_HD int cell_code(const Board<G> &b, int bit) {
    return (testbit(b.p[0], bit) ? 1 : 0) | (testbit(b.p[1], bit) ? 2 : 0) | (testbit(b.p[2], bit) ? 4 : 0) |
           (testbit(b.p[3], bit) ? 8 : 0);
}
template <class G>
ECG_HD void set_code(Board<G> &b, int bit, int code) { // cell must be empty
    const int wi = bit >> 5;
    const uint32_t m = 1u << (bit & 31);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t mk = ((code >> k) & 1) ? m : 0u;
#pragma unroll
        for (int i = 0; i < G::W; i++) b.p[k].w[i] |= (i == wi) ? mk : 0u;
    }
}

// effects of every special token on the board (boardv2.py:141-154): all of them fire in every
// cascade iteration because a typeless special always has token_board == 0.  Specials are on some
// board of the warp in almost every trip, so lines are expanded bit-parallel, without a per-token loop.
template <class G>
ECG_HD BB<G::W> fill_rows_with_any(const BB<G::W> &x) { // every row holding a bit of x, completely
    // per row field (C cells + pad bit): x + (2^C - 1) carries into the pad bit iff the row is non-empty
    constexpr int W = G::W;
    const BB<W> v = G::valid(), pad = shl<G::C>(G::col0());
    BB<W> sum;
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const uint64_t t = (uint64_t)x.w[i] + v.w[i] + carry;
        sum.w[i] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
    }
    const BB<W> flag = shr<G::C>(sum & pad); // bit r*S set for non-empty rows
    // flag * (2^C - 1) = (flag << C) - flag: no borrow leaves a row field
    const BB<W> hi = shl<G::C>(flag);
    BB<W> r;
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const uint64_t t = (uint64_t)hi.w[i] - flag.w[i] - borrow;
        r.w[i] = (uint32_t)t;
        borrow = (uint32_t)(t >> 63);
    }
    return r;
}
template <class G>
ECG_HD BB<G::W> fill_cols_with_any(const BB<G::W> &x) { // every column holding a bit of x, completely
    constexpr int S = G::S;
    BB<G::W> u = x | shr<S>(x); // fold all rows onto row 0 ...
    u |= shr<2 * S>(u);
    if (G::R > 4) u |= shr<4 * S>(u);
    if (G::R > 8) u |= shr<8 * S>(u);
    u = u & G::template rect<0, 1, 0, G::C>();
    u |= shl<S>(u); // ... and broadcast row 0 back down
    u |= shl<2 * S>(u);
    if (G::R > 4) u |= shl<4 * S>(u);
    if (G::R > 8) u |= shl<8 * S>(u);
    return u & G::valid();
}
template <class SH>
ECG_HD BB<SH::G::W> trigger_specials(const Board<typename SH::G> &bd, const BB<SH::G::W> &spec) {
    using G = typename SH::G;
    constexpr int W = G::W;
    const BB<W> k0 = bd.p[0] & spec, k1 = bd.p[1] & spec;
    const BB<W> hl = andn(andn(spec, k0), k1);
    BB<W> z = bb_zero<W>();
    if (any(hl)) z = fill_rows_with_any<G>(hl); // h_line: its row (:147-148)
    const BB<W> vl = andn(k0, k1);
    if (any(vl)) z |= fill_cols_with_any<G>(vl);               // v_line: its column (:149-150)
    BB<W> bombs = andn(k1, k0);
    while (any(bombs)) { // bomb at (i, j): token_board[j-1:j+1, i-1:i+1] -- transposed 2x2; a negative start
        const int b = pop_lowest(bombs); // (i == 0 or j == 0) makes the Python slice empty (:151-154)
        const int i = b / G::S, j = b - i * G::S;
        if (i > 0 && j > 0) {
            const int t = (j - 1) * G::S + (i - 1); // cells t, t+1, t+S, t+S+1
            const BB<W> two = onehot<W>(t) | onehot<W>(t + 1);
            z |= (two | shl<G::S>(two)) & G::valid();
        }
    }
    return z;
}

// gravity (boardv2.py:166-173): every column keeps its non-empty cells, in order, at the bottom
#if !defined(ECG_GRAVITY_COUNT)
// cells with a hole anywhere below (log-step smear) move down one row; repeat until stable
template <class G>
ECG_PHASE void gravity(Board<G> &b) {
    constexpr int W = G::W, S = G::S;
#if defined(ECG_SIM_HOOKS) && !defined(__CUDA_ARCH__) // host experiments only (scripts/experiments): found by ADL
    ecg_sim_gravity_hook(b);
#endif
    for (;;) {
        const BB<W> occ = b.p[0] | b.p[1] | b.p[2] | b.p[3];
        const BB<W> holes = andn(G::valid(), occ);
        BB<W> u = shr<S>(holes); // a hole 1 row below ...
        u |= shr<S>(u);          // ... 1..2 rows below
        u |= shr<2 * S>(u);      // ... 1..4
        if (G::R > 5) u |= shr<4 * S>(u); // ... 1..8 (the bottom row is R - 1 rows below the top one)
        if (G::R > 9) u |= shr<8 * S>(u); // ... 1..16
        const BB<W> f = occ & u;
        if (!any(f)) return;
#pragma unroll
        for (int k = 0; k < 4; k++) b.p[k] = andn(b.p[k], f) | shl<S>(b.p[k] & f);
    }
}
#else
// -DECG_GRAVITY_COUNT, a measured alternative (7 % SLOWER at 9x9x6: 5.99e9 vs 6.47e9 env-steps/s, r03a; 250 instead
// of 83 static instructions, and the row-by-row loop's iterations are cheaper than they look because most lanes of a
// warp need the same 1..3 of them): every cell falls by the number of holes below it in its column.  That count is built as a
// bit-sliced number (one bitboard per binary digit) by doubling the window -- 1, 2, 4, 8 rows below -- with ripple
// adders of bitboards; then cells whose count has bit k set move down 2^k rows, lowest bit first (the order of
// Hacker's Delight's compress: two cells of a column never meet, because their remaining distances stay ordered).
// The counter digits still to be used travel with their cells.  No loop whose trip count differs between the
// boards of a warp (the row-by-row form ran to the deepest column of the 32 boards).
template <int K, int W>
ECG_HD void count_window(BB<W> *cnt, int have, bool grow) { // cnt (have digits) += cnt moved up by K bits
    BB<W> carry = bb_zero<W>();
#pragma unroll
    for (int k = 0; k < have; k++) {
        const BB<W> e = shr<K>(cnt[k]);
        const BB<W> x = cnt[k] ^ e;
        const BB<W> nc = (cnt[k] & e) | (carry & x);
        cnt[k] = x ^ carry;
        carry = nc;
    }
    if (grow) cnt[have] = carry; // else the sum is known to fit
}
template <int K, class G>
ECG_HD void fall_by(Board<G> &b, BB<G::W> *cnt, int k, int nb) { // cells with digit k set move down K bits
    const BB<G::W> m = cnt[k];
    if (!any(m)) return;
#pragma unroll
    for (int j = 0; j < 4; j++) b.p[j] = andn(b.p[j], m) | shl<K>(b.p[j] & m);
#pragma unroll
    for (int j = k + 1; j < nb; j++) cnt[j] = andn(cnt[j], m) | shl<K>(cnt[j] & m);
}
template <class G>
ECG_PHASE void gravity(Board<G> &b) {
    constexpr int W = G::W, S = G::S, R = G::R;
    constexpr int NB = R > 8 ? 4 : R > 4 ? 3 : R > 2 ? 2 : 1; // binary digits of R - 1 (rows below the top cell)
    const BB<W> occ = b.p[0] | b.p[1] | b.p[2] | b.p[3];
    const BB<W> holes = andn(G::valid(), occ);
    BB<W> cnt[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) cnt[k] = bb_zero<W>();
    cnt[0] = shr<S>(holes);                               // holes in the 1 row below
    if (R > 2) count_window<S>(cnt, 1, NB > 1);           // 2 rows below
    if (R > 3) count_window<2 * S>(cnt, 2, NB > 2);       // 4 rows
    if (R > 5) count_window<4 * S>(cnt, 3, NB > 3);       // 8 rows
    if (R > 9) count_window<8 * S>(cnt, 4, false);        // 16 rows
#pragma unroll
    for (int k = 0; k < NB; k++) cnt[k] &= occ;
    fall_by<S>(b, cnt, 0, NB);
    if (NB > 1) fall_by<2 * S>(b, cnt, 1, NB);
    if (NB > 2) fall_by<4 * S>(b, cnt, 2, NB);
    if (NB > 3) fall_by<8 * S>(b, cnt, 3, NB);
}
#endif

// plane |= bit when (v & M): one predicate-setting LOP3 and one predicated LOP3 (the compiler's own branch-free
// form is shift, arithmetic shift, and, add)
template <uint32_t M>
ECG_HD void deposit_bit(uint32_t &plane, uint32_t v, uint32_t bit) {
#if defined(__CUDA_ARCH__) && !defined(ECG_NO_PRED_DEPOSIT)
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 q, t, 0;\n\t@q or.b32 %0, %0, %3;\n\t}"
        : "+r"(plane)
        : "r"(v), "n"(M), "r"(bit));
#else
    if (v & M) plane |= bit;
#endif
}

// refill (boardv2.py:172-173).  Replay: the reference's order (columns left to right, first draw =
// topmost hole).  Philox: holes in bit (row-major) order, draws addressed per cascade iteration.
template <class SH, class RNG>
ECG_PHASE void refill(Board<typename SH::G> &b, RNG &rng, uint32_t types, int iter) {
    using G = typename SH::G;
    const BB<G::W> holes = andn(G::valid(), b.p[0] | b.p[1] | b.p[2] | b.p[3]);
    if constexpr (RNG::ROW_MAJOR) {
        rng.seek((uint32_t)iter * 2048u + 1u);
#if !defined(ECG_REFILL_DIGIT_LOOP)
        uint32_t lo, hi; // the next tiles, one per nibble
        rng.first_tiles(types, lo, hi);
#pragma unroll
        for (int w = 0; w < G::W; w++) {
            uint32_t h = holes.w[w];
            for (;;) {
                while (h != 0u && lo != 0u) { // one compare-and-branch per tile; running out of tiles is rare
                    const uint32_t bit = h & (0u - h);
                    h ^= bit;
                    deposit_bit<1>(b.p[0].w[w], lo, bit);
                    deposit_bit<2>(b.p[1].w[w], lo, bit);
                    deposit_bit<4>(b.p[2].w[w], lo, bit);
                    if (SH::TPL > 3) deposit_bit<8>(b.p[3].w[w], lo, bit);
                    lo = funnel_r(lo, hi, 4);
                    hi >>= 4;
                }
                if (h == 0u) break;
                rng.more_tiles(types, lo, hi);
            }
        }
#else
#pragma unroll
        for (int w = 0; w < G::W; w++) {
            uint32_t h = holes.w[w];
            while (h) {
                const uint32_t bit = h & (0u - h);
                h ^= bit;
                const uint32_t v = 1u + rng.digit(types);
                deposit_bit<1>(b.p[0].w[w], v, bit);
                deposit_bit<2>(b.p[1].w[w], v, bit);
                deposit_bit<4>(b.p[2].w[w], v, bit);
                if (SH::TPL > 3) deposit_bit<8>(b.p[3].w[w], v, bit);
            }
        }
#endif
    } else if constexpr (UsesTileWindow<RNG>::value) {
        // the reference's order with precomputed tiles (ReplayTileRng): 16 tiles per window, one nibble per hole.
        // ONE flat loop, one hole per trip, the next hole chosen without a branch (down the column, else the top of
        // the next column with a hole): as nested column / hole loops the lanes of a warp waited for each other at
        // every column change (24 trips per warp iteration at 5 of 32 lanes, ncu r04c / r04d).
        uint32_t cols = holes.w[0] & ((1u << G::C) - 1u); // holes are top-aligned after gravity
        if (cols) {
            uint32_t t = rng.tpos;
            int bit = ctz32(cols);
            bool more = true;
            do { // one window of 16 tiles per pass: a second pass is rare
                uint32_t lo, hi;
                rng.window(t, lo, hi);
                t += 16u;
                int left = 16;
                do {
                    const int wi = bit >> 5;
                    const uint32_t m = 1u << (bit & 31);
#pragma unroll
                    for (int i = 0; i < G::W; i++) {
                        const uint32_t mi = (i == wi) ? m : 0u;
                        deposit_bit<1>(b.p[0].w[i], lo, mi);
                        deposit_bit<2>(b.p[1].w[i], lo, mi);
                        deposit_bit<4>(b.p[2].w[i], lo, mi);
                        if (SH::TPL > 3) deposit_bit<8>(b.p[3].w[i], lo, mi);
                    }
                    lo = funnel_r(lo, hi, 4);
                    hi >>= 4;
                    const int nb = bit + G::S;
                    const bool down = testbit(holes, nb); // (bits past the board are never set)
                    const uint32_t rest = cols & (cols - 1u);
                    more = down || rest != 0u;
                    cols = down ? cols : rest;
                    bit = down ? nb : ctz32(rest | 0x80000000u);
                } while (more && --left);
            } while (more);
            rng.tpos += (uint32_t)popcount(holes);
            rng.check_end();
        }
    } else {
        uint32_t cols = holes.w[0] & ((1u << G::C) - 1u); // holes are top-aligned after gravity
        while (cols) {
            int bit = ctz32(cols);
            cols &= cols - 1u;
            do {
                set_code(b, bit, 1 + (int)rng.digit(types));
                bit += G::S;
            } while (bit < G::NB && testbit(holes, bit));
        }
    }
}

// boardFunctions.shuffle (:16-23): reseed, permute ROWS (numpy legacy Fisher-Yates), then cells
// that held a special before the shuffle get their old value back.
template <class SH, class RNG>
ECG_HD void shuffle_rows_impl(Board<typename SH::G> &b, RNG &rng) {
    using G = typename SH::G;
    constexpr int W = G::W;
    constexpr uint32_t RM = (1u << G::C) - 1u;
    rng.reseed();
    const Board<G> old = b;
    const BB<W> old_spec = old.p[3] & old.p[2];
    uint32_t rows[4][G::R];
    for (int k = 0; k < 4; k++)
        for (int r = 0; r < G::R; r++) rows[k][r] = extract32(b.p[k], r * G::S) & RM;
    for (int i = G::R - 1; i >= 1; i--) {
        const int j = (int)rng.below((uint32_t)i + 1u);
        for (int k = 0; k < 4; k++) {
            const uint32_t t = rows[k][j];
            rows[k][j] = rows[k][i];
            rows[k][i] = t;
        }
    }
    for (int k = 0; k < 4; k++) {
        BB<W> p = bb_zero<W>();
        for (int r = 0; r < G::R; r++) {
            const int b0 = r * G::S, wi = b0 >> 5, s = b0 & 31;
            for (int i = 0; i < W; i++) {
                if (i == wi) p.w[i] |= rows[k][r] << s;
                if (i == wi + 1 && s) p.w[i] |= rows[k][r] >> (32 - s);
            }
        }
        b.p[k] = andn(p, old_spec) | (old.p[k] & old_spec);
    }
}

template <class SH, class RNG>
struct ShuffleOut {
    Board<typename SH::G> b;
    RNG rng;
};
template <class SH, class RNG>
ECG_HD_NOINLINE ShuffleOut<SH, RNG> shuffle_rows(Board<typename SH::G> b, RNG rng) {
    ShuffleOut<SH, RNG> o;
    o.b = b;
    o.rng = rng;
    shuffle_rows_impl<SH>(o.b, o.rng);
    return o;
}

struct StepOut {
    int reward;       // points of this step (boardv2.py:157-158 summed over the cascade)
    int cascades;     // iterations of the cascade loop (:138), >= 1
    uint32_t status;  // ST_* bits
};

// Swapping two special tokens (boardv2.py:81-132).  s1/s2: 0 none, 1 h_line, 2 v_line, 3 bomb, 4 mega of
// the source / target cell after the swap; `target` is the bit of the second (lower/right) cell, all regions
// are relative to it.  Returns true when the normal get_matches path applies (:134-136).
template <class G>
struct PairOut {
    BB<G::W> cleared;
    bool matched;
};
template <class G>
ECG_HD bool special_pair_impl(int s1, int s2, int target, BB<G::W> &cleared) {
    const int tr = target / G::S, tc = target - tr * G::S;
    const int lo = s1 < s2 ? s1 : s2, hi = s1 < s2 ? s2 : s1;
    if (hi == 4) { // a mega token is involved
        if (lo == 4) cleared = G::valid(); // :81-82
        // mega + bomb / line / plain (:84-103): token = max(token1, token2) is the mega value itself, never
        // present in token_board -> no cell changes and get_matches is skipped
        return false;
    }
    if (hi == 3 && lo >= 1) { // bomb + bomb: 4x4 block (:112-116); bomb + line: 4 columns and 4 rows (:123-125)
        const BB<G::W> rr = G::rows(tr - 2, tr + 2), cc = G::cols(tc - 2, tc + 2);
        cleared = lo == 3 ? (rr & cc) : (rr | cc);
        return false;
    }
    if (lo == 1 && hi == 2) { // h_line + v_line (:130-132): ROW slices [:tc] and [tr:]
        cleared = andn(G::valid(), G::rows(tc, tr));
        return false;
    }
    return true; // one special + plain, or two equal lines
}
template <class G>
ECG_HD_NOINLINE PairOut<G> special_pair(int s1, int s2, int target) {
    PairOut<G> o;
    o.cleared = bb_zero<G::W>();
    o.matched = special_pair_impl<G>(s1, s2, target, o.cleared);
    return o;
}

// One in-flight BoardV2.apply_action (boardv2.py:43-207), cut at the cascade-loop boundary so a kernel
// can interleave boards: step_begin = swap + special-pair branch / first get_matches (:46-136),
// step_iter = ONE iteration of the cascade loop (:138-202).  The GPU step kernel keeps one Lane per
// thread and lets threads whose cascade ended fetch the next board while others keep cascading.
template <class SH>
struct Lane {
    Board<typename SH::G> bd;
    BB<SH::G::W> cleared;       // plain cells whose token_board entry was zeroed (match mask / special region)
    BB<SH::G::W> sp, sk0, sk1;  // pending spawns (get_match_spawn_mask) of the coming iteration
    int reward, cascades;
    uint32_t status;
};

// b1 = bit of the source (upper / left) cell, d = 1 (horizontal swap) or S (vertical): target = b1 + d.
// FAST = true (see find_matches): returns true, with L half-updated, when the step needs the exact build -- a
// swap of two special tokens, or a rare match case; the caller drops L and steps the board again from its
// unchanged input with FAST = false.
template <class SH, bool FAST = false>
ECG_HD bool step_begin_at(Lane<SH> &L, int b1, int d) {
    using G = typename SH::G;
    constexpr int W = G::W;
    Board<G> &bd = L.bd;
    L.reward = 0;
    L.cascades = 0;
    L.status = 0;
    const int b2 = b1 + d;
    // swap (:51).  Per plane the two cells' bits are gathered into one word (they sit at different bit positions
    // modulo 32, because d is 1 or S < 32): the plane changes iff exactly one of them is set.
    const BB<W> m1 = onehot<W>(b1), m2 = onehot<W>(b2);
    const BB<W> both = m1 | m2;
    uint32_t pair = 0;
#pragma unroll
    for (int i = 0; i < W; i++) pair |= both.w[i];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < W; i++) x |= bd.p[k].w[i] & both.w[i];
        const uint32_t flip = (x != 0u && x != pair) ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int i = 0; i < W; i++) bd.p[k].w[i] ^= both.w[i] & flip;
    }
    bool matched = true;
    L.sp = bb_zero<W>();
    L.sk0 = bb_zero<W>();
    L.sk1 = bb_zero<W>();
    L.cleared = bb_zero<W>();
    const BB<W> spec2 = bd.p[3] & bd.p[2] & both;
    if (any(spec2)) { // a special token was moved
        if constexpr (FAST) {
            // one special + a plain token (1.9 % of random legal steps: a typeless special makes every swap with
            // it legal) is special_pair_impl's last line: get_matches decides -- except mega + plain, which clears
            // nothing (:84-103).  Two specials go to the exact build.
            if (any(spec2 & m1) && any(spec2 & m2)) {
                ECG_RARE(3);
                return true;
            }
            matched = !any(bd.p[0] & bd.p[1] & spec2); // code 15 = mega
        } else {
            const int c1 = cell_code<G>(bd, b1), c2 = cell_code<G>(bd, b2); // source / target AFTER the swap
            const int s1 = c1 >= 12 ? c1 - 11 : 0, s2 = c2 >= 12 ? c2 - 11 : 0; // 0 none, 1 h, 2 v, 3 bomb, 4 mega
            const PairOut<G> o = special_pair<G>(s1, s2, b2);
            L.cleared = o.cleared;
            matched = o.matched;
        }
    }
    if (matched) {
        const Derived<G> d_ = derive<SH>(bd);
        Matches<G> m;
        find_matches<SH, FAST>(d_, m);
        if (FAST && m.rare) return true;
        L.cleared = m.mask;
        L.sp = m.sp;
        L.sk0 = m.sk0;
        L.sk1 = m.sk1;
    }
    return false;
}
template <class SH>
ECG_HD void step_begin(Lane<SH> &L, int action) {
    int b1, b2;
    decode_action<typename SH::G>(action, b1, b2); // source, target (:48)
    step_begin_at<SH>(L, b1, b2 - b1);
}

// One cascade iteration.  Returns true when the step is over; then HL/VL are the legal swaps of the
// final board (the reference computes legal_actions there too, :188, to decide about shuffling).
// FAST = true (see find_matches): `rare` is set (and false returned) when the iteration needs the exact build:
// a rare match case, or a final board without a legal swap (the shuffle loop).
// DEFER_LEGAL = true (the pooled step kernel): when the cascade ends, HL / VL are left alone; the caller computes the
// legal swaps of the final board later, when a full warp of finished boards has gathered (and hands the board over
// itself if none is legal).
template <class SH, class RNG, bool FAST, bool DEFER_LEGAL = false>
ECG_HD bool step_iter(Lane<SH> &L, RNG &rng, uint32_t types, BB<SH::G::W> &HL, BB<SH::G::W> &VL, bool &rare) {
    using G = typename SH::G;
    constexpr int W = G::W;
    Board<G> &bd = L.bd;
    rare = false;
    L.cascades++;
    { // :141-163 trigger pass, points, clear, spawn, clip
        const BB<W> spec = bd.p[3] & bd.p[2];
        const BB<W> occ = bd.p[0] | bd.p[1] | bd.p[2] | bd.p[3];
        BB<W> z = L.cleared | spec | andn(G::valid(), occ); // token_board == 0
        if (any(spec)) z |= trigger_specials<SH>(bd, spec);
        const BB<W> k1 = bd.p[1] & spec, k0 = bd.p[0] & spec; // points (:58-65, :157-158)
        L.reward += 2 * popcount(andn(z, spec)) + 25 * popcount(andn(spec, k1)) + 50 * popcount(andn(k1, k0)) +
                    250 * popcount(k1 & k0);
#pragma unroll
        for (int k = 0; k < 4; k++) bd.p[k] = andn(bd.p[k], z);
        bd.p[3] |= L.sp; // spawns always land on cleared cells
        bd.p[2] |= L.sp;
        bd.p[0] |= L.sk0;
        bd.p[1] |= L.sk1;
    }
    gravity<G>(bd);                             // :166-170
    refill<SH>(bd, rng, types, L.cascades - 1); // :172-173
    if (FAST && rng.overflow) { // replay stream exhausted: the exact build reports it
        rare = true;
        return false;
    }
    Derived<G> d = derive<SH>(bd);
    Matches<G> m;
    find_matches<SH, FAST>(d, m); // :181
    if (FAST && m.rare) {
        rare = true;
        return false;
    }
    bool done = false;
    if constexpr (DEFER_LEGAL) {
        if (!m.found) return true;
        L.cleared = m.mask;
        L.sp = m.sp;
        L.sk0 = m.sk0;
        L.sk1 = m.sk1;
        if (rng.overflow || L.cascades >= CASCADE_CAP) { // the pending match stays unapplied, like below
            L.status |= rng.overflow ? ST_STREAM_OVERFLOW : ST_CASCADE_CAP;
            return true;
        }
        return false;
    }
    if (!m.found) {
        legal_swaps<SH>(d, m.A, m.D, HL, VL);
        if (FAST && !any(HL | VL)) {
            ECG_RARE(4);
            rare = true;
            return false;
        }
        if constexpr (!FAST) {
            int shuffles = 0;
            while (!m.found && !any(HL | VL)) { // :188-194
                if (shuffles++ >= SHUFFLE_CAP) {
                    L.status |= ST_SHUFFLE_CAP;
                    break;
                }
                if (shuffles == 1) rng.seek((uint32_t)(L.cascades - 1) * 2048u + 1024u);
                {
                    const ShuffleOut<SH, RNG> so = shuffle_rows<SH, RNG>(bd, rng);
                    bd = so.b;
                    rng = so.rng;
                }
                d = derive<SH>(bd);
                find_matches<SH>(d, m);
                if (!m.found) legal_swaps<SH>(d, m.A, m.D, HL, VL);
            }
        }
        done = !m.found; // :195
    }
    if (!done) {
        L.cleared = m.mask; // :199
        L.sp = m.sp;        // :202
        L.sk0 = m.sk0;
        L.sk1 = m.sk1;
        if (rng.overflow || L.cascades >= CASCADE_CAP) {
            if (!rng.overflow) L.status |= ST_CASCADE_CAP;
            legal_swaps<SH>(d, m.A, m.D, HL, VL);
            done = true;
        }
    }
    if (done && rng.overflow) L.status |= ST_STREAM_OVERFLOW;
    return done;
}

// What the coming cascade iteration of a lane will cost, from its pending match (the pooled step kernel runs lanes
// of one class together, so the gravity / refill loops of a warp end together): 0 = one horizontal run of three on a
// board without special tokens (three holes in one row: one gravity move, three tiles), 1 = one vertical run of three
// (up to three moves, three tiles), 2 = everything else.
constexpr int CASCADE_CLASSES = 3;
template <class SH>
ECG_HD int cascade_class(const Lane<SH> &L) {
    using G = typename SH::G;
    const bool plain3 = popcount(L.cleared) == 3 && !any((L.bd.p[3] & L.bd.p[2]) | L.sp);
    if (!plain3) return 2;
    return any(L.cleared & shl<G::S>(L.cleared)) ? 1 : 0;
}

template <class SH, class RNG>
ECG_HD bool step_iter(Lane<SH> &L, RNG &rng, uint32_t types, BB<SH::G::W> &HL, BB<SH::G::W> &VL) {
    bool rare;
    return step_iter<SH, RNG, false>(L, rng, types, HL, VL, rare);
}

// BoardV2.apply_action (boardv2.py:43-207) minus the terminal test, run to completion.
template <class SH, class RNG>
ECG_HD void step_board(Board<typename SH::G> &bd, int action, uint32_t types, RNG &rng, StepOut &out,
                       BB<SH::G::W> &HL, BB<SH::G::W> &VL) {
    Lane<SH> L;
    L.bd = bd;
    rng.reseed(); // :46
    step_begin<SH>(L, action);
    while (!step_iter<SH>(L, rng, types, HL, VL)) {
    }
    bd = L.bd;
    out.reward = L.reward;
    out.cascades = L.cascades;
    out.status = L.status;
}

// The two-kernel step on one board: the common-case build first, the exact build from the unchanged input when it
// hands the board off.  Host tests use it to fuzz the FAST logic against the oracle; returns true on a hand-off.
// frng drives the common-case pass, rng the exact pass (replay: ReplayTileRng / ReplayRng over the same stream); when
// no board is handed off, rng is left untouched and frng holds the position.
template <class SH, class FRNG, class RNG>
ECG_HD bool step_board_two_pass(Board<typename SH::G> &bd, int action, uint32_t types, FRNG &frng, RNG &rng,
                                StepOut &out, BB<SH::G::W> &HL, BB<SH::G::W> &VL) {
    Lane<SH> L;
    L.bd = bd;
    frng.reseed();
    int b1, b2;
    decode_action<typename SH::G>(action, b1, b2);
    bool rare = step_begin_at<SH, true>(L, b1, b2 - b1);
    while (!rare && !step_iter<SH, FRNG, true>(L, frng, types, HL, VL, rare)) {
    }
    if (rare) {
        step_board<SH>(bd, action, types, rng, out, HL, VL);
        return true;
    }
    bd = L.bd;
    out.reward = L.reward;
    out.cascades = L.cascades;
    out.status = L.status;
    return false;
}
template <class SH, class RNG>
ECG_HD bool step_board_two_pass(Board<typename SH::G> &bd, int action, uint32_t types, RNG &rng, StepOut &out,
                                BB<SH::G::W> &HL, BB<SH::G::W> &VL) {
    RNG frng = rng;
    const bool handed_off = step_board_two_pass<SH, RNG, RNG>(bd, action, types, frng, rng, out, HL, VL);
    if (!handed_off) rng = frng;
    return handed_off;
}

// BoardV2.__init__ (boardv2.py:20-27): draw a board, redraw matched cells until clean.
template <class SH, class RNG>
ECG_HD void init_board(Board<typename SH::G> &bd, uint32_t types, RNG &rng) {
    using G = typename SH::G;
    rng.reseed();
#pragma unroll
    for (int k = 0; k < 4; k++) bd.p[k] = bb_zero<G::W>();
    for (int r = 0; r < G::R; r++)
        for (int c = 0; c < G::C; c++) set_code(bd, r * G::S + c, 1 + (int)rng.below(types));
    Matches<G> m;
    for (int guard = 0; guard < 100000; guard++) {
        Derived<G> d = derive<SH>(bd);
        find_matches<SH>(d, m);
        if (!m.found || rng.overflow) break;
        // a full board of fresh draws is consumed every round; only masked cells take theirs (:25-26)
        Board<G> fresh;
#pragma unroll
        for (int k = 0; k < 4; k++) fresh.p[k] = bb_zero<G::W>();
        for (int r = 0; r < G::R; r++)
            for (int c = 0; c < G::C; c++) set_code(fresh, r * G::S + c, 1 + (int)rng.below(types));
#pragma unroll
        for (int k = 0; k < 4; k++) bd.p[k] = andn(bd.p[k], m.mask) | (fresh.p[k] & m.mask);
    }
}

// ------------------------------------------------------------------ dataset augmentation (dataset.py:86-112)

// np.fliplr(observation) (dataset.py:94): reverse the C cells of every row, plane by plane
template <class G>
ECG_HD Board<G> mirror_board(const Board<G> &b) {
    constexpr uint32_t RM = (1u << G::C) - 1u;
    Board<G> o;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o.p[k] = bb_zero<G::W>();
#pragma unroll
        for (int r = 0; r < G::R; r++) {
            const uint32_t row = extract32(b.p[k], r * G::S) & RM;
            uint32_t rev = 0;
#pragma unroll
            for (int c = 0; c < G::C; c++) rev |= ((row >> c) & 1u) << (G::C - 1 - c);
            const int b0 = r * G::S, wi = b0 >> 5, s = b0 & 31;
#pragma unroll
            for (int i = 0; i < G::W; i++) {
                if (i == wi) o.p[k].w[i] |= rev << s;
                if (i == wi + 1 && s) o.p[k].w[i] |= rev >> (32 - s);
            }
        }
    }
    return o;
}

// every cell code c becomes lut[c] (type permutation: plain codes permuted, 0 and the specials kept)
template <class G>
ECG_HD Board<G> remap_codes(const Board<G> &b, const uint8_t lut[16]) {
    Board<G> o;
#pragma unroll
    for (int k = 0; k < 4; k++) o.p[k] = bb_zero<G::W>();
    for (int code = 1; code < 16; code++) { // rolled: an augmentation pass, not the step path
        BB<G::W> m = G::valid();
#pragma unroll
        for (int k = 0; k < 4; k++) m = ((code >> k) & 1) ? (m & b.p[k]) : andn(m, b.p[k]);
        const int t = lut[code];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((t >> k) & 1) o.p[k] |= m;
    }
    return o;
}

// ------------------------------------------------------------------ cell codec

struct CellCodec { // boardConfig.py:29-43
    int type_mask, h_line, v_line, bomb, mega;
};
ECG_HD CellCodec make_codec(int types) {
    int bits = 0;
    while ((1 << bits) < types + 1) bits++;
    CellCodec c;
    c.type_mask = (1 << bits) - 1;
    c.h_line = c.type_mask + 1;
    c.v_line = 2 * c.h_line;
    c.bomb = (1 << (bits + 1)) + 1 + c.type_mask;
    c.mega = c.type_mask + c.bomb + 1;
    return c;
}
ECG_HD int encode_cell(const CellCodec &c, long long v) { // -> code or -1
    if (v == 0) return 0;
    if (v > 0 && v <= c.type_mask && v <= 11) return (int)v;
    if (v == c.h_line) return 12;
    if (v == c.v_line) return 13;
    if (v == c.bomb) return 14;
    if (v == c.mega) return 15;
    return -1;
}
ECG_HD int decode_cell(const CellCodec &c, int code) {
    return code < 12 ? code : code == 12 ? c.h_line : code == 13 ? c.v_line : code == 14 ? c.bomb : c.mega;
}

} // namespace ecg
