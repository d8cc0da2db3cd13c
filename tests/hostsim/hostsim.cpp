// hostsim.cpp -- TEST-ONLY host build of the kernel core (element-crush-gym_b200/csrc/ecg_core.cuh).
// g++ compiles the exact __host__ __device__ board logic that nvcc puts into the sm_100a kernels, so
// the bitboard algorithm can be fuzzed against the CPU oracle without a GPU.  Not part of the product:
// nothing under element-crush-gym_b200/ loads this library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../element-crush-gym_b200/csrc/ecg_core.cuh"

using namespace ecg;

template <class SH>
static uint32_t pack_board(const CellCodec &cc, const int64_t *cells, Board<typename SH::G> &b) {
    using G = typename SH::G;
    uint32_t st = 0;
    for (int k = 0; k < 4; k++) b.p[k] = bb_zero<G::W>();
    for (int r = 0; r < G::R; r++)
        for (int c = 0; c < G::C; c++) {
            int code = encode_cell(cc, cells[r * G::C + c]);
            if (code < 0) {
                st |= ST_BAD_CELL;
                code = 0;
            }
            set_code(b, r * G::S + c, code);
        }
    return st;
}

template <class SH>
static void unpack_board(const CellCodec &cc, const Board<typename SH::G> &b, int64_t *cells) {
    using G = typename SH::G;
    for (int r = 0; r < G::R; r++)
        for (int c = 0; c < G::C; c++) cells[r * G::C + c] = decode_cell(cc, cell_code<G>(b, r * G::S + c));
}

template <class SH>
static void legal_bytes(const BB<SH::G::W> &HL, const BB<SH::G::W> &VL, uint8_t *out) {
    using G = typename SH::G;
    uint32_t m[G::AW];
    swaps_to_actions<G>(HL, VL, m);
    for (int a = 0; a < G::A; a++) out[a] = (m[a >> 5] >> (a & 31)) & 1u;
}

struct StepArgs {
    int types, mode;
    const uint32_t *raw;
    int64_t raw_stride, raw_len;
    uint64_t key, board0;
    uint32_t step_ctr;
    const int64_t *in;
    const int32_t *actions, *moves_left;
    int64_t *out, *reward;
    int32_t *cascades;
    uint8_t *status, *legal;
    int64_t n;
    int64_t *handoffs; // modes 3 / 4: incremented per board the common-case build handed off
    uint32_t *words;   // optional, replay: raw words consumed by the step
    int64_t *class_errors; // mode 6: lanes whose pending match contradicts their cascade_class
};

template <class SH>
static void step_t(const StepArgs &a) {
    using G = typename SH::G;
    const CellCodec cc = make_codec(a.types);
    const int N = G::R * G::C;
    for (int64_t i = 0; i < a.n; i++) {
        Board<G> b;
        uint32_t st = pack_board<SH>(cc, a.in + i * N, b);
        StepOut so = {0, 0, 0};
        BB<G::W> HL, VL;
        const bool terminal = a.moves_left && a.moves_left[i] < 1;
        const bool bad = a.actions[i] < 0 || a.actions[i] >= G::A;
        if (terminal || bad) {
            so.status = terminal ? ST_TERMINAL : ST_BAD_ACTION;
            Derived<G> d = derive<SH>(b);
            legal_swaps<SH>(d, eq_at<SH, 1>(d), eq_at<SH, G::S>(d), HL, VL);
        } else if (a.mode == 6) {
            // the pooled step kernel's decomposition of the common-case pass: BEGIN, then one step_iter<DEFER_LEGAL>
            // per pass on a lane whose gravity / refill work its cascade_class predicted, FINISH = the legal swaps of
            // the final board (a board without any goes to the exact build, like any rare case)
            PhiloxRng rng;
            rng.init(a.key, a.board0 + (uint64_t)i, a.step_ctr);
            Lane<SH> L;
            L.bd = b;
            int b1, b2;
            decode_action<G>(a.actions[i], b1, b2);
            bool rare = step_begin_at<SH, true>(L, b1, b2 - b1);
            bool fin = false;
            while (!rare && !fin) {
                const int cls = cascade_class<SH>(L);
                const BB<G::W> spec = L.bd.p[3] & L.bd.p[2];
                const BB<G::W> gone = L.cleared | spec; // cleared cells (a class 0 / 1 board holds no special)
                if (cls < 2 && (popcount(gone) != 3 || any(spec) || any(L.sp))) ++*a.class_errors;
                if (cls == 0 && any(gone & shl<G::S>(gone))) ++*a.class_errors; // three holes in ONE row
                if (cls == 1 && !any(gone & shl<2 * G::S>(gone))) ++*a.class_errors; // three in one column
                PhiloxRng r2 = rng;
                fin = step_iter<SH, PhiloxRng, true, true>(L, r2, (uint32_t)a.types, HL, VL, rare);
            }
            if (!rare) {
                Derived<G> d = derive<SH>(L.bd);
                legal_swaps<SH>(d, eq_at<SH, 1>(d), eq_at<SH, G::S>(d), HL, VL);
                rare = !any(HL | VL);
            }
            if (rare) {
                if (a.handoffs) ++*a.handoffs;
                step_board<SH>(b, a.actions[i], (uint32_t)a.types, rng, so, HL, VL);
            } else {
                b = L.bd;
                so.reward = L.reward;
                so.cascades = L.cascades;
                so.status = L.status;
            }
        } else if (a.mode == 2 || a.mode == 4) { // mode + 2: common-case build first, exact build on a hand-off
            PhiloxRng rng;
            rng.init(a.key, a.board0 + (uint64_t)i, a.step_ctr);
            if (a.mode == 4) {
                if (step_board_two_pass<SH>(b, a.actions[i], (uint32_t)a.types, rng, so, HL, VL) && a.handoffs)
                    ++*a.handoffs;
            } else {
                step_board<SH>(b, a.actions[i], (uint32_t)a.types, rng, so, HL, VL);
            }
        } else {
            ReplayRng rng;
            rng.init(a.raw + i * a.raw_stride, (uint32_t)a.raw_len, 0);
            if (a.mode == 3) { // the replay two-kernel step: precomputed tiles in the common-case pass
                std::vector<uint32_t> tiles(replay_tile_words((int)a.raw_len));
                std::vector<uint16_t> wpos(a.raw_len + 1);
                build_replay_tiles(a.raw + i * a.raw_stride, (int)a.raw_len, (uint32_t)a.types, tiles.data(), wpos.data());
                ReplayTileRng frng;
                frng.init(tiles.data(), wpos.data(), (uint32_t)a.raw_len, 0);
                const bool handed_off = step_board_two_pass<SH>(b, a.actions[i], (uint32_t)a.types, frng, rng, so, HL, VL);
                if (handed_off && a.handoffs) ++*a.handoffs;
                // np.random's position behind the step: the tile table's word count must equal the exact build's
                if (a.words) a.words[i] = handed_off ? rng.pos : frng.words();
            } else {
                step_board<SH>(b, a.actions[i], (uint32_t)a.types, rng, so, HL, VL);
                if (a.words) a.words[i] = rng.pos;
            }
        }
        unpack_board<SH>(cc, b, a.out + i * N);
        if (a.reward) a.reward[i] = so.reward;
        if (a.cascades) a.cascades[i] = so.cascades;
        if (a.status) a.status[i] = (uint8_t)(so.status | st);
        if (a.legal) legal_bytes<SH>(HL, VL, a.legal + i * G::A);
    }
}

template <class SH>
static void legal_t(int types, const int64_t *boards, uint8_t *legal, int64_t n) {
    using G = typename SH::G;
    const CellCodec cc = make_codec(types);
    for (int64_t i = 0; i < n; i++) {
        Board<G> b;
        pack_board<SH>(cc, boards + i * G::R * G::C, b);
        Derived<G> d = derive<SH>(b);
        BB<G::W> HL, VL;
        legal_swaps<SH>(d, eq_at<SH, 1>(d), eq_at<SH, G::S>(d), HL, VL);
        legal_bytes<SH>(HL, VL, legal + i * G::A);
    }
}

// get_matches + get_match_spawn_mask on the token board of `boards` (plain tokens only matter)
template <class SH>
static void matches_t(int types, const int64_t *boards, uint8_t *mask, int32_t *spawn, int64_t n) {
    using G = typename SH::G;
    const CellCodec cc = make_codec(types);
    const int kinds[4] = {cc.h_line, cc.v_line, cc.bomb, cc.mega};
    for (int64_t i = 0; i < n; i++) {
        Board<G> b;
        pack_board<SH>(cc, boards + i * G::R * G::C, b);
        Derived<G> d = derive<SH>(b);
        Matches<G> m;
        find_matches<SH>(d, m);
        for (int r = 0; r < G::R; r++)
            for (int c = 0; c < G::C; c++) {
                const int bit = r * G::S + c;
                mask[i * G::R * G::C + r * G::C + c] = m.found && testbit(m.mask, bit);
                int v = 0;
                if (m.found && testbit(m.sp, bit)) v = kinds[(testbit(m.sk0, bit) ? 1 : 0) | (testbit(m.sk1, bit) ? 2 : 0)];
                spawn[i * G::R * G::C + r * G::C + c] = v;
            }
    }
}

template <class SH>
static void init_t(int types, const uint32_t *raw, int64_t raw_stride, int64_t raw_len, uint64_t key,
                   uint64_t board0, int mode, int64_t *out, uint8_t *status, int64_t n) {
    using G = typename SH::G;
    const CellCodec cc = make_codec(types);
    for (int64_t i = 0; i < n; i++) {
        Board<G> b;
        bool ovf;
        if (mode == 2) {
            PhiloxRng rng;
            rng.init(key, board0 + (uint64_t)i, 0xFFFFFFFFu);
            init_board<SH>(b, (uint32_t)types, rng);
            ovf = false;
        } else {
            ReplayRng rng;
            rng.init(raw + i * raw_stride, (uint32_t)raw_len, 0);
            init_board<SH>(b, (uint32_t)types, rng);
            ovf = rng.overflow;
        }
        unpack_board<SH>(cc, b, out + i * G::R * G::C);
        if (status) status[i] = ovf ? ST_STREAM_OVERFLOW : 0;
    }
}

// One library per board size (-DHS_SIZE=N, tests/hostsim/hostsim.py builds them in parallel): the core is a large
// header and thirteen sizes x two type widths in one translation unit took g++ nine minutes.
#ifndef HS_SIZE
#error "compile with -DHS_SIZE=<board size>"
#endif
#define DISPATCH(rows, types, CALL)                              \
    do {                                                         \
        if ((rows) != HS_SIZE) return -1;                        \
        if ((types) >= 8) {                                      \
            using SH = Shape<HS_SIZE, HS_SIZE, 4, true>;         \
            CALL;                                                \
        } else {                                                 \
            using SH = Shape<HS_SIZE, HS_SIZE, 3, false>;        \
            CALL;                                                \
        }                                                        \
        return 0;                                                \
    } while (0)

static int64_t g_handoffs = 0, g_class_errors = 0;
static uint32_t *g_words = nullptr; // hs_set_words: where the next replay hs_step reports the words each step drew

extern "C" {

int hs_step(int rows, int cols, int types, int mode, const uint32_t *raw, int64_t raw_stride, int64_t raw_len,
            uint64_t key, uint64_t board0, uint32_t step_ctr, const int64_t *in, const int32_t *actions,
            const int32_t *moves_left, int64_t *out, int64_t *reward, int32_t *cascades, uint8_t *status,
            uint8_t *legal, int64_t n) {
    if (rows != cols || types < 1 || types > 11) return -1;
    g_handoffs = 0;
    g_class_errors = 0;
    StepArgs a = {types, mode, raw, raw_stride, raw_len, key, board0, step_ctr, in, actions, moves_left,
                  out, reward, cascades, status, legal, n, &g_handoffs, g_words, &g_class_errors};
    g_words = nullptr;
    DISPATCH(rows, types, step_t<SH>(a));
}

void hs_set_words(uint32_t *words) { g_words = words; }

// boards the common-case build handed off in the last hs_step call (modes 3 / 4)
int64_t hs_handoffs(void) { return g_handoffs; }
int64_t hs_class_errors(void) { return g_class_errors; }

int hs_legal(int rows, int cols, int types, const int64_t *boards, uint8_t *legal, int64_t n) {
    if (rows != cols || types < 1 || types > 11) return -1;
    DISPATCH(rows, types, legal_t<SH>(types, boards, legal, n));
}

int hs_matches(int rows, int cols, int types, const int64_t *boards, uint8_t *mask, int32_t *spawn, int64_t n) {
    if (rows != cols || types < 1 || types > 11) return -1;
    DISPATCH(rows, types, matches_t<SH>(types, boards, mask, spawn, n));
}

int hs_init(int rows, int cols, int types, int mode, const uint32_t *raw, int64_t raw_stride, int64_t raw_len,
            uint64_t key, uint64_t board0, int64_t *out, uint8_t *status, int64_t n) {
    if (rows != cols || types < 1 || types > 11) return -1;
    DISPATCH(rows, types, init_t<SH>(types, raw, raw_stride, raw_len, key, board0, mode, out, status, n));
}
}
