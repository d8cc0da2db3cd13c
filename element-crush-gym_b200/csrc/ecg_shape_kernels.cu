// ecg_shape_kernels.cu -- sm_100a kernels of libecg.so for ONE board size (compiled once per
// -DECG_SIZE=N so the sizes build in parallel); ecg_api.cu holds the C-ABI (include/ecg.h).
//
// One thread owns one board for the whole step: the board is 4 bit-planes x W words
// (12 registers for 9x9), every operation is a 32-bit integer op on those registers
// (LOP3 / SHF / IADD3 / POPC), no shared memory, no shuffles, no tensor cores.
//
// HBM layout ("tile-interleaved"): boards are grouped in tiles of 32; inside a tile the
// 16-byte chunk k of board `lane` sits at uint4 index (tile*CH + k)*32 + lane, so one warp
// moves a whole tile with CH fully coalesced 512-byte LDG.128/STG.128 transactions.
// Legal masks use the same tiling with 4-byte words: u32 index (tile*MW + w)*32 + lane.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/ecg.h"
#include "ecg_core.cuh"
#include "ecg_ops.h"

#ifndef ECG_SIZE
#error "compile with -DECG_SIZE=<board size>"
#endif

using namespace ecg;

static_assert(ECG_ST_TERMINAL == ST_TERMINAL && ECG_ST_STREAM_OVERFLOW == ST_STREAM_OVERFLOW &&
                  ECG_ST_SHUFFLE_CAP == ST_SHUFFLE_CAP && ECG_ST_BAD_ACTION == ST_BAD_ACTION &&
                  ECG_ST_NO_LEGAL == ST_NO_LEGAL && ECG_ST_BAD_CELL == ST_BAD_CELL &&
                  ECG_ST_CASCADE_CAP == ST_CASCADE_CAP,
              "status bits of include/ecg.h and ecg_core.cuh must agree");

namespace {

// Board / mask traffic uses plain (write-back) accesses: lanes of a warp finish their boards in different trips,
// so a 32-byte sector is written 16 bytes at a time; with evict-first hints (-DECG_STREAMING_LDST: __ldcs/__stcs)
// the halves reached HBM separately (ncu: 4.23 GB of DRAM traffic per step against 2.85 GB of buffers), with
// write-back the L2 merges them (2.99 GB).
#if defined(ECG_STREAMING_LDST)
#define ECG_LD(p) __ldcs(p)
#define ECG_ST(p, v) __stcs(p, v)
#else
#define ECG_LD(p) (*(p))
#define ECG_ST(p, v) (*(p) = (v))
#endif

#ifndef ECG_BLOCK
#define ECG_BLOCK 128
#endif
constexpr int BLOCK = ECG_BLOCK;


// ------------------------------------------------------------------ packed I/O

template <class G>
__device__ __forceinline__ void load_board(const void *boards, long long i, Board<G> &b) {
    constexpr int BW = 4 * G::W, CH = BW / 4;
    const uint4 *base = reinterpret_cast<const uint4 *>(boards) + (i >> 5) * (CH * 32) + (i & 31);
    uint32_t w[BW];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        const uint4 v = ECG_LD(base + k * 32);
        w[4 * k] = v.x;
        w[4 * k + 1] = v.y;
        w[4 * k + 2] = v.z;
        w[4 * k + 3] = v.w;
    }
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < G::W; j++) b.p[p].w[j] = w[p * G::W + j];
}

template <class G>
__device__ __forceinline__ void store_board(void *boards, long long i, const Board<G> &b) {
    constexpr int BW = 4 * G::W, CH = BW / 4;
    uint4 *base = reinterpret_cast<uint4 *>(boards) + (i >> 5) * (CH * 32) + (i & 31);
    uint32_t w[BW];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < G::W; j++) w[p * G::W + j] = b.p[p].w[j];
#pragma unroll
    for (int k = 0; k < CH; k++) ECG_ST(base + k * 32, make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]));
}

// packed legal mask = the two swap bitboards, HL words then VL words, tiled like the boards with 4-byte words
template <class G>
__device__ __forceinline__ void load_mask(const uint32_t *mask, long long i, BB<G::W> &HL, BB<G::W> &VL) {
    const uint32_t *base = mask + (i >> 5) * (2 * G::W * 32) + (i & 31);
#pragma unroll
    for (int k = 0; k < G::W; k++) HL.w[k] = ECG_LD(base + k * 32);
#pragma unroll
    for (int k = 0; k < G::W; k++) VL.w[k] = ECG_LD(base + (G::W + k) * 32);
}
template <class G>
__device__ __forceinline__ void store_mask(uint32_t *mask, long long i, const BB<G::W> &HL, const BB<G::W> &VL) {
    uint32_t *base = mask + (i >> 5) * (2 * G::W * 32) + (i & 31);
#pragma unroll
    for (int k = 0; k < G::W; k++) ECG_ST(base + k * 32, HL.w[k]);
#pragma unroll
    for (int k = 0; k < G::W; k++) ECG_ST(base + (G::W + k) * 32, VL.w[k]);
}

// replay: the raw MT19937 words of board i (ecg_refill.stream_index redirects boards to shared streams)
__device__ __forceinline__ const uint32_t *stream_of(const RefillDev &rf, long long i) {
    return rf.stream + (rf.stream_index ? (long long)rf.stream_index[i] : i) * rf.stream_stride;
}

__device__ __forceinline__ long long stream_id(const RefillDev &rf, long long i) {
    return rf.stream_stride ? (rf.stream_index ? (long long)rf.stream_index[i] : i) : 0;
}

template <class SH>
__device__ __forceinline__ void legal_of(const Board<typename SH::G> &b, BB<SH::G::W> &HL, BB<SH::G::W> &VL) {
    using G = typename SH::G;
    const Derived<G> d = derive<SH>(b);
    legal_swaps<SH>(d, eq_at<SH, 1>(d), eq_at<SH, G::S>(d), HL, VL);
}

// ------------------------------------------------------------------ kernels

template <class SH, typename T>
__global__ void __launch_bounds__(BLOCK) pack_kernel(const T *__restrict__ cells, void *boards, uint8_t *status,
                                                     int types, long long n) {
    using G = typename SH::G;
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const CellCodec cc = make_codec(types);
    const T *src = cells + i * (G::R * G::C);
    Board<G> b;
#pragma unroll
    for (int k = 0; k < 4; k++) b.p[k] = bb_zero<G::W>();
    uint32_t st = 0;
    for (int r = 0; r < G::R; r++)
        for (int c = 0; c < G::C; c++) {
            int code = encode_cell(cc, (long long)src[r * G::C + c]);
            if (code < 0) {
                st = ST_BAD_CELL;
                code = 0;
            }
            set_code(b, r * G::S + c, code);
        }
    store_board<G>(boards, i, b);
    if (status) status[i] = (uint8_t)st;
}

// board.array / the observation: one warp per tile of 32 boards.  The tile (32 x 4W words) is staged in shared
// memory with coalesced 16-byte loads; the lanes then walk the tile's 32*R*C output cells in order, so every
// store instruction of the warp writes one contiguous run (128 B for uint8 cells, 4 per lane; 256 B for int64).
// `cells` must be 4-byte aligned for uint8 output (op_unpack checks and otherwise uses unpack_kernel_simple).
template <class SH, typename T>
__global__ void __launch_bounds__(BLOCK) unpack_kernel(const void *boards, T *__restrict__ cells, int types,
                                                       long long n) {
    using G = typename SH::G;
    constexpr int BW = 4 * G::W, RC = G::R * G::C, VEC = sizeof(T) == 1 ? 4 : 1;
    static_assert((32 * RC) % VEC == 0, "tile size must be a multiple of the store width");
    __shared__ uint32_t tile_words[BLOCK / 32][32 * BW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long tile = (long long)blockIdx.x * (BLOCK / 32) + wib;
    if (tile * 32 >= n) return;
    uint32_t *tw = tile_words[wib];
    { // chunk k of board `lane` is uint4 (tile*CH + k)*32 + lane; keep that order in shared memory
        const uint4 *src = reinterpret_cast<const uint4 *>(boards) + tile * (BW / 4 * 32);
#pragma unroll
        for (int k = 0; k < BW / 4; k++) reinterpret_cast<uint4 *>(tw)[k * 32 + lane] = src[k * 32 + lane];
    }
    __syncwarp();
    const CellCodec cc = make_codec(types);
    const long long rest = n - tile * 32;
    const int total = (int)(rest < 32 ? rest : 32) * RC; // cells of this tile
    T *dst = cells + tile * 32 * RC;
    for (int e0 = lane * VEC; e0 < total; e0 += 32 * VEC) {
        uint32_t packed = 0;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int e = e0 + v;
            const int b = e / RC, idx = e - b * RC;
            const int r = idx / G::C, c = idx - r * G::C, bit = r * G::S + c;
            int code = 0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int j = p * G::W + (bit >> 5); // word j of the board = word (j & 3) of chunk (j >> 2)
                code |= (int)((tw[((j >> 2) * 32 + b) * 4 + (j & 3)] >> (bit & 31)) & 1u) << p;
            }
            const int value = decode_cell(cc, code);
            if (VEC == 1) dst[e] = (T)value;
            else packed |= (uint32_t)value << (8 * v);
        }
        if (VEC == 4) {
            uint8_t *d8 = reinterpret_cast<uint8_t *>(dst) + e0;
            if (e0 + VEC <= total) {
                *reinterpret_cast<uint32_t *>(d8) = packed;
            } else { // the last word of a partial tile: only the cells that exist (total need not be a multiple of 4)
                for (int v = 0; e0 + v < total; v++) d8[v] = (uint8_t)(packed >> (8 * v));
            }
        }
    }
}

// The compact observation: 4-bit cell CODES (0 empty, 1..11 plain token, 12 h_line, 13 v_line, 14 bomb, 15 mega),
// row-major, two cells per byte (cell 2k in the low nibble of byte k), ceil(R*C/2) bytes per board, boards back to
// back.  Same staging as unpack_kernel; a lane's 32-bit store carries 8 cells, and a word may span two boards.
template <class SH>
__global__ void __launch_bounds__(BLOCK) unpack_nibbles_kernel(const void *boards, uint8_t *__restrict__ out,
                                                               long long n) {
    using G = typename SH::G;
    constexpr int BW = 4 * G::W, RC = G::R * G::C, NBY = (RC + 1) / 2;
    static_assert((32 * NBY) % 4 == 0, "a tile's output is a whole number of words");
    __shared__ uint32_t tile_words[BLOCK / 32][32 * BW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long tile = (long long)blockIdx.x * (BLOCK / 32) + wib;
    if (tile * 32 >= n) return;
    uint32_t *tw = tile_words[wib];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(boards) + tile * (BW / 4 * 32);
#pragma unroll
        for (int k = 0; k < BW / 4; k++) reinterpret_cast<uint4 *>(tw)[k * 32 + lane] = src[k * 32 + lane];
    }
    __syncwarp();
    const long long rest = n - tile * 32;
    const int total = (int)(rest < 32 ? rest : 32) * NBY; // bytes of this tile
    uint8_t *dst = out + tile * 32 * NBY;
    for (int y0 = lane * 4; y0 < total; y0 += 32 * 4) {
        uint32_t packed = 0;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const int y = y0 + v;
            const int b = y / NBY, k = y - b * NBY;
            uint32_t byte = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int idx = 2 * k + h;
                if (idx < RC) {
                    const int r = idx / G::C, c = idx - r * G::C, bit = r * G::S + c;
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const int j = p * G::W + (bit >> 5);
                        byte |= ((tw[((j >> 2) * 32 + b) * 4 + (j & 3)] >> (bit & 31)) & 1u) << (4 * h + p);
                    }
                }
            }
            packed |= byte << (8 * v);
        }
        if (y0 + 4 <= total) {
            *reinterpret_cast<uint32_t *>(dst + y0) = packed;
        } else {
            for (int v = 0; y0 + v < total; v++) dst[y0 + v] = (uint8_t)(packed >> (8 * v));
        }
    }
}

template <class SH, typename T>
__global__ void __launch_bounds__(BLOCK) unpack_kernel_simple(const void *boards, T *__restrict__ cells, int types,
                                                              long long n) {
    using G = typename SH::G;
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const CellCodec cc = make_codec(types);
    Board<G> b;
    load_board<G>(boards, i, b);
    T *dst = cells + i * (G::R * G::C);
    for (int r = 0; r < G::R; r++)
        for (int c = 0; c < G::C; c++) dst[r * G::C + c] = (T)decode_cell(cc, cell_code<G>(b, r * G::S + c));
}

template <class SH>
__global__ void __launch_bounds__(BLOCK) unpack_mask_kernel(const uint32_t *mask, uint8_t *__restrict__ out,
                                                            long long n) {
    using G = typename SH::G;
    // one thread per (board, action): coalesced byte stores
    const long long t = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (t >= n * G::A) return;
    const long long i = t / G::A;
    const int a = (int)(t - i * G::A);
    const int r = a / G::ROWA, k = a - r * G::ROWA; // boardConfig.py:45-59
    const bool vertical = k >= G::C - 1;
    const int bit = r * G::S + (vertical ? k - (G::C - 1) : k);
    const uint32_t w = mask[(i >> 5) * (2 * G::W * 32) + ((vertical ? G::W : 0) + (bit >> 5)) * 32 + (i & 31)];
    out[t] = (w >> (bit & 31)) & 1u;
}

// nnx.one_hot(board.array, channels) (elementCrush.py:92): one thread per (board, cell) writes `channels` values
template <class SH, typename T>
__global__ void __launch_bounds__(BLOCK) onehot_kernel(const void *boards, T *__restrict__ out, int channels, T one,
                                                       int types, long long n) {
    using G = typename SH::G;
    const long long t = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (t >= n * (G::R * G::C)) return;
    const long long i = t / (G::R * G::C);
    const int cell = (int)(t - i * (G::R * G::C));
    const int r = cell / G::C, c = cell - r * G::C, bit = r * G::S + c;
    constexpr int CH = G::W; // chunks of 4 words: word index plane*W + (bit >> 5)
    const uint32_t *w = reinterpret_cast<const uint32_t *>(boards) + (i >> 5) * (CH * 32 * 4);
    int code = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int j = p * G::W + (bit >> 5);
        const uint32_t x = w[((j >> 2) * 32 + (int)(i & 31)) * 4 + (j & 3)];
        code |= (int)((x >> (bit & 31)) & 1u) << p;
    }
    const int value = decode_cell(make_codec(types), code);
    T *dst = out + t * channels;
    for (int k = 0; k < channels; k++) dst[k] = k == value ? one : (T)0;
}

// dataset augmentation (dataset.py:86-112, 114-176): mirrored and / or type-permuted copies of packed boards
template <class SH>
__global__ void __launch_bounds__(BLOCK) augment_kernel(const void *boards_in, void *boards_out, bool mirror, bool remap,
                                                        CodeLut lut, long long n) {
    using G = typename SH::G;
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    Board<G> b;
    load_board<G>(boards_in, i, b);
    if (mirror) b = mirror_board<G>(b);
    if (remap) b = remap_codes<G>(b, lut.v);
    store_board<G>(boards_out, i, b);
}

template <class SH>
__global__ void __launch_bounds__(BLOCK) legal_kernel(const void *boards, uint32_t *mask, long long n) {
    using G = typename SH::G;
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    Board<G> b;
    load_board<G>(boards, i, b);
    BB<G::W> HL, VL;
    legal_of<SH>(b, HL, VL);
    store_mask<G>(mask, i, HL, VL);
}


template <bool PHILOX, bool FAST = false>
struct RngOf {
    using type = PhiloxRng;
};
template <>
struct RngOf<false, false> {
    using type = ReplayRng;
};
template <>
struct RngOf<false, true> { // the common-case kernel of a replay step reads the precomputed tile tables
    using type = ReplayTileRng;
};

template <class SH, bool PHILOX>
__global__ void __launch_bounds__(BLOCK) random_action_kernel(RefillDev rf, const uint32_t *mask, int32_t *actions,
                                                              uint8_t *status, long long n) {
    using G = typename SH::G;
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    BB<G::W> HL, VL;
    load_mask<G>(mask, i, HL, VL);
    const int cnt = swaps_count<G>(HL, VL);
    int a = -1;
    uint32_t st = 0;
    if (cnt == 0) {
        st = ST_NO_LEGAL;
    } else if (PHILOX) { // k-th legal swap in swap-bitboard order
        a = swaps_select<G>(HL, VL, (int)philox_pick(rf.key, rf.board0 + (unsigned long long)i, rf.step_ctr, (uint32_t)cnt));
    } else { // np.random.choice(legal_actions): k-th legal action in ascending action order
        ReplayRng rng;
        rng.init(stream_of(rf, i), (uint32_t)rf.stream_len, rf.stream_pos ? rf.stream_pos[i] : 0u);
        a = swaps_select_action<G>(HL, VL, (int)rng.below((uint32_t)cnt));
        if (rf.stream_pos) rf.stream_pos[i] = rng.pos;
        if (rng.overflow) st = ST_STREAM_OVERFLOW;
    }
    actions[i] = a;
    if (status) status[i] = (uint8_t)st;
}

template <class SH, bool PHILOX>
__global__ void __launch_bounds__(BLOCK) init_kernel(RefillDev rf, void *boards, uint8_t *status, int types,
                                                     long long n) {
    using G = typename SH::G;
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    Board<G> b;
    typename RngOf<PHILOX>::type rng;
    if constexpr (PHILOX) rng.init(rf.key, rf.board0 + (unsigned long long)i, 0xFFFFFFFFu);
    else rng.init(stream_of(rf, i), (uint32_t)rf.stream_len, 0u);
    init_board<SH>(b, (uint32_t)types, rng);
    store_board<G>(boards, i, b);
    if (status) status[i] = rng.overflow ? (uint8_t)ST_STREAM_OVERFLOW : (uint8_t)0;
}


// ---- the lockstep step: a persistent warp loop over work chunks ------------------------------------
//
// Cascade lengths differ per board (mean 1.5 iterations, long tail), so "one thread = one board for the
// whole launch" leaves most lanes idle while the slowest board of the warp finishes.  Instead every lane
// runs a small state machine: FETCH a board (load, choose the action, swap, first match pass), then one
// cascade ITERATION per trip of the warp loop, then FINISH (legal mask, stores) and fetch the next board
// in the same trip.  Every trip therefore runs one cascade iteration on (almost) all 32 lanes.
// Boards are handed out in chunks of CHUNK_BOARDS consecutive boards; warp w owns chunks w, w + nwarps, ...
// (static, no atomics); lanes that ask in the same trip receive consecutive boards, so their 16-byte
// chunk loads/stores stay contiguous inside the 32-board tile.
#ifndef ECG_CHUNK
#define ECG_CHUNK 256
#endif
constexpr int CHUNK_BOARDS = ECG_CHUNK;
#ifndef ECG_MIN_CHUNKS
#define ECG_MIN_CHUNKS 8 // chunks every warp should get before the chunk size stops shrinking (16 / 32 / 64: no gain, r08)
#endif
// The lane kernel runs ONE block of LANE_BLOCK threads per SM whose warps walk the trip loop together (one
// __syncthreads_or per trip).  The loop body is ~35 KB of SASS, more than the SM's 32 KB instruction cache; 16
// free-running warps each streamed it on their own (stall_no_inst 56 % of stall samples, GPC instruction-fetch path
// at 96 % of peak).  In lockstep a cache line fetched by one warp is reused by the other 15: +11 % env-steps/s
// (4.39e9 -> 4.87e9 at 9x9x6; 128 x 4 in lockstep: 4.40e9, 256 x 2: 4.84e9, 512 x 1: 4.87e9).
// Boards of 16x16 (9 words per plane) need ~250 registers: 512-thread blocks (128 registers) spilled 650 B per thread
// (r04b: 1.42e9 env-steps/s; 256 threads x 255 registers: 1.98e9; 384 x 168: 1.67e9).  256 also wins at 15x15 (2.04e9 ->
// 2.54e9) and 14x14 (2.09e9 -> 2.28e9); 13x13 (2.89e9 vs 2.83e9) and 12x12 (3.85e9 vs 3.45e9) are best at 512 (r04c).
#if ECG_SIZE >= 14 && !defined(ECG_LANE_BLOCK) && !defined(ECG_FAST_BLOCK)
#define ECG_LANE_BLOCK 256
#define ECG_FAST_BLOCK 256
#endif
#ifndef ECG_LANE_BLOCK
#define ECG_LANE_BLOCK 512
#endif
#ifndef ECG_LANE_MINB
#define ECG_LANE_MINB 1
#endif
constexpr int LANE_BLOCK = ECG_LANE_BLOCK;
// The common-case kernel fits 96 registers per thread, i.e. 640-thread blocks (20 warps per SM instead of 16), but
// measures the same with them (6.47e9 vs 6.46e9 env-steps/s): it is issue bound, not latency bound.  576 threads
// (18 warps, uneven over the 4 sub-partitions) are 9 % slower.
#ifndef ECG_FAST_BLOCK
#define ECG_FAST_BLOCK 512
#endif
__host__ __device__ constexpr int lane_block(bool fast) { return fast ? ECG_FAST_BLOCK : LANE_BLOCK; }
#if !defined(ECG_LANE_FREE_RUNNING)
#define ECG_TRIP_ANY(p) __syncthreads_or(p)
#else
#define ECG_TRIP_ANY(p) __any_sync(0xffffffffu, p)
#endif

template <class SH, bool PHILOX>
__device__ __forceinline__ void finish_board(const RefillDev &rf, const StepDev &io, long long i, long long src,
                                             const Board<typename SH::G> &b, bool stepped, int action, int moves,
                                             int reward, int cascades, uint32_t status, const BB<SH::G::W> &HL,
                                             const BB<SH::G::W> &VL, uint32_t rpos, int score_in = 0,
                                             bool have_score = false) {
    using G = typename SH::G;
    if (stepped || io.boards_out != io.boards_in) store_board<G>(io.boards_out, i, b);
    if (io.mask_out) store_mask<G>(io.mask_out, i, HL, VL);
    if (io.actions_out) io.actions_out[i] = action;
    if (io.moves_left) io.moves_left[i] = moves;
    if (io.reward) io.reward[i] = reward;
    int score = reward;
    if (io.score) {
        score += have_score ? score_in : io.score[src];
        io.score[i] = score;
    }
    if (io.cascades) io.cascades[i] = cascades;
    if (io.flags) { // env.py:54-55
        const bool won = score >= io.env_goal;
        io.flags[i] = (uint8_t)((won || moves == 0 ? ECG_FLAG_DONE : 0) | (won ? ECG_FLAG_WON : 0));
    }
    if (io.status) io.status[i] = (uint8_t)status;
    if constexpr (!PHILOX)
        if (rf.stream_pos) rf.stream_pos[i] = rpos;
}

struct RolloutDev {
    void *boards;
    const int32_t *moves_left;
    long long *total_reward;
    int32_t *steps_done;
    uint8_t *status;
};

// One kernel for the lockstep step (ROLLOUT = false: every board takes ONE action, all per-step outputs are
// written) and for whole-episode rollouts (ROLLOUT = true: a board keeps stepping with random legal actions
// until moves_left reaches 0; only the final board and the collected reward go back to HBM).
// Lane states: IDLE (needs a board) -> READY (board + legal swaps in registers, action not chosen yet)
//              -> ACTIVE (inside the cascade loop) -> IDLE | READY (rollout: next action of the same board).
//
// FAST = true is the first kernel of a two-kernel step (StepDev::handoff): the common-case build of the board logic
// (find_matches<SH, true>: no intersecting runs, no run of 6+, no swap of two specials, no shuffle).  A lane that
// meets one of those drops its board -- nothing of it has been written -- and appends the job to the hand-off list;
// the exact kernel (FAST = false, launched right after on the same stream) steps those boards from their unchanged
// inputs.  Without the rare paths the trip body fits the SM's instruction cache, so the FAST kernel's warps run
// free (no trip barrier); the exact kernel keeps the one-barrier-per-trip lockstep described above.
template <class SH, bool PHILOX, bool ROLLOUT, bool FAST>
__global__ void __launch_bounds__(lane_block(FAST), ECG_LANE_MINB) lane_kernel(RefillDev rf, StepDev io, RolloutDev ro, int types,
                                                                    int n) {
    using G = typename SH::G;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int IDLE = 0, READY = 1, ACTIVE = 2;
    static_assert(!(FAST && ROLLOUT && !PHILOX), "two-kernel rollouts are Philox-mode only");
    if constexpr (!FAST)
        if (io.n_jobs) n = *io.n_jobs;
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); // (blockDim.x <= lane_block(FAST))
    // The hand-off list of a two-kernel step is short (0.6 % of the boards: 94 k jobs for 2 368 warps): every warp takes
    // ONE contiguous range of (almost) the same length.  Chunks of 32 handed round-robin left some warps with 64 jobs and
    // others with 32 (6.98e9 env-steps/s at 9x9x6; chunks of 16: 7.11e9; equal ranges: 7.18e9).
    const int nwarps_all = (int)((gridDim.x * blockDim.x) >> 5);
    // Chunks of CHUNK_BOARDS boards -- smaller (down to one tile) when the batch is small, so that every warp still gets
    // eight of them: with fixed 256-board chunks a batch of 2^17 boards filled 32 of the 148 SMs and one of 2^21 left
    // a quarter of the warps idle for the last chunk.
    int chunk = ::CHUNK_BOARDS;
    while (chunk > 32 && (long long)n < (long long)nwarps_all * chunk * ECG_MIN_CHUNKS) chunk >>= 1;
    const int CHUNK = (!FAST && io.n_jobs) ? (n + nwarps_all - 1) / nwarps_all + (n == 0) : chunk;
    const int stride = nwarps_all * CHUNK; // host guarantees n + stride + CHUNK < 2^31
    // warp-uniform cursor over this warp's chunks
    int chunk0 = warp * CHUNK;
    int next = chunk0 < n ? chunk0 : n;
    int end = chunk0 + CHUNK < n ? chunk0 + CHUNK : n;

    Lane<SH> L;
    L.cascades = 0;
    int state = IDLE;
    int idx = 0, moves = 0, action = -1;
    int src = 0; // board whose state, refill stream and Philox id this job uses (== idx unless io.src_index)
    uint32_t step = rf.step_ctr; // philox step counter of this lane's board (advances inside a rollout)
    uint32_t rpos = 0;           // replay: words consumed since the last reseed
    long long total = 0;         // rollout: points collected so far
    int steps_done = 0;
    uint32_t st_acc = 0;
    int score_in = 0; // FAST: the board's score, loaded with the board
    BB<G::W> HL, VL;

    for (;;) {
#if defined(ECG_BARRIER_TOP)
        __syncthreads();
#endif
        // ---- cursor: idle lanes are handed the next boards
        const unsigned need = __ballot_sync(FULL, state == IDLE);
        int cand = n;
        if (need) {
            const int rank = __popc(need & ((1u << lane) - 1u));
            const int cnt = __popc(need);
            const int avail = end - next;
            if (rank < avail) {
                cand = next + rank;
            } else if (chunk0 + stride < n) { // the request spills into this warp's next chunk
                const int c2 = chunk0 + stride;
                const int e2 = c2 + CHUNK < n ? c2 + CHUNK : n;
                cand = c2 + (rank - avail);
                if (cand >= e2) cand = n;
            }
            if (cnt >= avail) {
                if (chunk0 + stride < n) {
                    chunk0 += stride;
                    end = chunk0 + CHUNK < n ? chunk0 + CHUNK : n;
                    next = chunk0 + (cnt - avail);
                    if (next > end) next = end;
                } else {
                    chunk0 = next = end = n;
                }
            } else {
                next += cnt;
            }
        }
        if constexpr (FAST) {
            // free-running warps: pull the lines of the boards this warp hands out next into L1 one trip ahead
            const int pf = next + lane;
            if (need && pf < end && !io.src_index) {
                constexpr int CH = G::W;
                const void *pfb = ROLLOUT ? (const void *)ro.boards : io.boards_in;
                const uint4 *bp = reinterpret_cast<const uint4 *>(pfb) + (long long)(pf >> 5) * (CH * 32) + (pf & 31);
#pragma unroll
                for (int k = 0; k < CH; k++) asm volatile("prefetch.global.L1 [%0];" ::"l"(bp + k * 32));
                // (prefetching the mask / action / moves / score lines as well measured 0.5 % slower, r03c)
            }
        }
        // ---- LOAD: IDLE -> READY
        if (state == IDLE && cand < n) {
            idx = cand;
            if constexpr (!FAST)
                if (io.jobs) idx = io.jobs[cand];
            src = idx;
            if constexpr (!ROLLOUT)
                if (io.src_index) src = io.src_index[idx];
            step = rf.step_ctr;
            rpos = 0;
            if constexpr (!PHILOX) // expanding (board, action) pairs: stream_pos is output only, indexed by job
                rpos = (rf.stream_pos && (ROLLOUT || !(io.actions && io.src_index))) ? rf.stream_pos[src] : 0u;
            if constexpr (ROLLOUT) {
                load_board<G>(ro.boards, idx, L.bd);
                moves = ro.moves_left[idx];
                total = 0;
                steps_done = 0;
                st_acc = 0;
                if constexpr (!FAST)
                    if (io.jobs) { // an episode the common-case rollout kernel handed over: go on where it stopped
                        steps_done = ro.steps_done[idx];
                        total = ro.total_reward[idx];
                        st_acc = ro.status ? ro.status[idx] : 0u;
                        moves -= steps_done;
                        step += (uint32_t)steps_done;
                    }
                legal_of<SH>(L.bd, HL, VL);
            } else {
                load_board<G>(io.boards_in, src, L.bd);
                moves = io.moves_left ? io.moves_left[src] : 1;
                if (!io.actions) load_mask<G>(io.mask_in, src, HL, VL);
                if constexpr (FAST) score_in = io.score ? io.score[src] : 0; // registers to spare: no load at FINISH
            }
            state = READY;
        }
        // ---- one Philox block per trip for every lane that needs random words in this trip: block 0 of the
        // (board, step) substream for a READY lane (word 0 = action pick, words 1..3 = first refill words),
        // block 512*j for a lane entering cascade iteration j
        uint32_t blk[4];
        uint32_t blk_index = 0;
        if constexpr (PHILOX) {
            // lanes that loaded a board above and lanes that did not must meet again here: without the explicit
            // reconvergence the compiler ran the ten Philox rounds twice per trip, once for each group (ncu r03:
            // 3.15 executions per 32 boards at 16.0 active threads, against 1.59 trips)
            __syncwarp();
            if (state != IDLE) {
                blk_index = state == READY ? 0u : (uint32_t)L.cascades * 512u;
                const unsigned long long board = rf.board0 + (unsigned long long)src;
                philox4x32_10(blk_index, step, (uint32_t)board, (uint32_t)(board >> 32), (uint32_t)rf.key,
                              (uint32_t)(rf.key >> 32), blk);
            }
        }
        // ---- BEGIN: READY -> ACTIVE: choose the action, swap, first match pass
        bool handoff = false; // FAST: this lane's board goes to the exact kernel
        if (state == READY) {
            uint32_t st = 0;
            action = -1;
            int b1 = -1, d = 0; // Philox pick: the swap as (source bit, 1 | S), no action decode needed
            if (moves < 1) { // boardv2.py:44
                st = ST_TERMINAL;
            } else if (!ROLLOUT && io.actions) {
                action = io.actions[idx];
                if (action < 0 || action >= G::A) {
                    st = ST_BAD_ACTION;
                    action = -1;
                }
            } else { // board.random_action(): uniform over the legal set of the current board
                const int c = swaps_count<G>(HL, VL);
                if (c == 0) {
                    st = ST_NO_LEGAL;
                } else if constexpr (PHILOX) {
                    bool vertical;
                    b1 = swaps_select_bit<G>(HL, VL, (int)mulhi32(blk[0], (uint32_t)c), vertical);
                    d = vertical ? G::S : 1;
                    action = action_of_swap<G>(b1, vertical);
                } else { // np.random.choice(legal_actions): ascending action order, numpy's masked rejection
                    ReplayRng rng;
                    rng.init(stream_of(rf, src), (uint32_t)rf.stream_len, rpos);
                    action = swaps_select_action<G>(HL, VL, (int)rng.below((uint32_t)c));
                    rpos = rng.pos;
                    if (rng.overflow) st = ST_STREAM_OVERFLOW;
                }
            }
            if (action >= 0) {
                rpos = 0; // np.random.seed(cfg.seed) at the top of apply_action (boardv2.py:46)
                if (b1 < 0) { // action given by the caller / replay pick: boardConfig.decode
                    int b2;
                    decode_action<G>(action, b1, b2);
                    d = b2 - b1;
                }
                // a rollout that may hand this step over keeps the board of the step's start in HBM (the first step's
                // is still there): the exact kernel resumes the episode from it
                if constexpr (ROLLOUT && FAST)
                    if (steps_done > 0) store_board<G>(ro.boards, idx, L.bd);
                handoff = step_begin_at<SH, FAST>(L, b1, d);
                L.status |= st;
                moves -= 1;
                state = handoff ? IDLE : ACTIVE;
            } else if constexpr (ROLLOUT) { // the episode is over (terminal, or nothing legal)
                if (st != ST_TERMINAL) st_acc |= st;
                store_board<G>(ro.boards, idx, L.bd);
                ro.total_reward[idx] = total;
                if (ro.steps_done) ro.steps_done[idx] = steps_done;
                if (ro.status) ro.status[idx] = (uint8_t)st_acc;
                if constexpr (!PHILOX)
                    if (rf.stream_pos) rf.stream_pos[idx] = rpos;
                state = IDLE;
            } else { // no-op boards are finished on the spot
                if (io.mask_out) legal_of<SH>(L.bd, HL, VL);
                finish_board<SH, PHILOX>(rf, io, idx, src, L.bd, false, -1, moves, 0, 0, st, HL, VL, rpos, score_in, FAST);
                state = IDLE;
            }
        }
        if constexpr (FAST) { // a board dropped in this trip is still to be appended to the hand-off list below
#if defined(ECG_FAST_LOCKSTEP) // measured alternative for the largest boards, whose unrolled body outgrows the i-cache
            if (!__syncthreads_or(state != IDLE || handoff)) break;
#else
            if (!__any_sync(FULL, state != IDLE || handoff)) break;
#endif
        } else {
            if (!ROLLOUT && io.n_jobs) { // the short job list of a two-kernel step: too few trips to pay for barriers
                // (parked rollout episodes run ten more steps each, several per lane: they keep the trip barrier --
                // 2^22 episodes in 13.35 instead of 13.77 ms, r06)
                if (!__any_sync(FULL, state != IDLE)) break;
            } else if (!ECG_TRIP_ANY(state != IDLE)) break;
        }
        // ---- ITERATE: one cascade iteration on every active lane; FINISH the steps whose cascade ended
        if (state == ACTIVE) {
            typename RngOf<PHILOX, FAST>::type rng;
            if constexpr (PHILOX) {
                rng.init(rf.key, rf.board0 + (unsigned long long)src, step);
                rng.preset_block(blk_index, blk);
            } else if constexpr (FAST) { // rpos counts TILES while the step is in flight (converted at the end)
                const long long sid = stream_id(rf, src);
                rng.init(rf.tiles + sid * replay_tile_words(rf.stream_len), rf.tile_wpos + sid * (rf.stream_len + 1),
                         (uint32_t)rf.stream_len, rpos);
            } else {
                rng.init(stream_of(rf, src), (uint32_t)rf.stream_len, rpos);
            }
            const bool fin = step_iter<SH, typename RngOf<PHILOX, FAST>::type, FAST>(L, rng, (uint32_t)types, HL, VL, handoff);
            if constexpr (!PHILOX) {
                if constexpr (FAST) rpos = fin ? rng.words() : rng.tpos; // np.random's position once the step is over
                else rpos = rng.pos;
            }
            if (FAST && handoff) state = IDLE;
            if (fin) {
                if constexpr (ROLLOUT) {
                    total += L.reward;
                    st_acc |= L.status;
                    steps_done++;
                    step++;
                    if (st_acc & ST_STREAM_OVERFLOW) moves = 0; // invalid from here on: end the episode
                    state = READY; // next action of this board (or the end of the episode)
                } else {
                    finish_board<SH, PHILOX>(rf, io, idx, src, L.bd, true, action, moves, L.reward, L.cascades, L.status,
                                             HL, VL, rpos, score_in, FAST);
                    state = IDLE;
                }
            }
        }
        if constexpr (FAST) { // hand the dropped boards over (idx is still this trip's board)
            const unsigned hm = __ballot_sync(FULL, handoff);
            if (hm) {
                int base = 0;
                if (lane == __ffs((int)hm) - 1) base = atomicAdd(io.handoff, __popc(hm));
                base = __shfl_sync(FULL, base, __ffs((int)hm) - 1);
                if (handoff) {
                    io.handoff[1 + base + __popc(hm & ((1u << lane) - 1u))] = idx;
                    if constexpr (ROLLOUT) { // the episode so far (its board of this step's start is already in HBM)
                        ro.total_reward[idx] = total;
                        ro.steps_done[idx] = steps_done;
                        if (ro.status) ro.status[idx] = (uint8_t)st_acc;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ pooled common-case step kernel
//
// The lane kernel above keeps one board per lane for its whole step; lanes of a warp are in different phases (65 % of
// them finish their cascade in a trip), so FINISH / FETCH / BEGIN ran at 20 of 32 lanes and the gravity / refill loops
// to the warp's worst board (ncu r09: 18 of 32 threads per instruction).  Here a BLOCK owns a pool of boards in shared
// memory (PoolGeo::SLOTS slots, word w of slot s at pool[w * SLOTS + s]) and runs ONE phase per pass on POOL_BLOCK of
// them, all threads together:
//   BEGIN   POOL_BLOCK boards from HBM: action pick, swap, first match pass -> slot, listed by cascade_class
//   ITERATE POOL_BLOCK slots of ONE class: one cascade iteration in place -> listed by its new class, or as finished
//   FINISH  POOL_BLOCK finished slots: legal swaps of the final board, outputs to HBM, slot freed
// Lists of slot numbers per class / finished / free, appended to with one shared-memory atomic per warp and list; two
// barriers per pass (lists merged and read | slots processed and listed).  All warps of a block run the same phase, so
// the instruction cache holds one phase at a time (warp-private pools were measured first, profiles/r10_warp_pool_*:
// 16 warps in three different phases stream 54 KB of SASS, i-cache hit 81 %).  Philox lockstep steps without
// src_index only; batches too small to keep every block's pool busy stay with the lane kernel.
//
// MEASURED AND REJECTED (round 2, profiles/r10_block_pool_*): bit-exact, but 2.64 ms per 2^24-board launch against the
// lane kernel's 2.15 ms.  The phases shrink as planned (BEGIN 830 -> 650, FINISH 475 -> 374 warp-instructions per 32
// boards at 32 of 32 lanes, gravity loop 390 -> 290, refill loops 290 -> 145), but moving the 28-word lane states
// through shared memory every pass, the lists, the per-pass decision and the class test cost ~630, and the issue rate
// falls from 60 % to 41-50 % (the loads of a BEGIN pass are no longer hidden behind other lanes' cascades).  Compiled
// only with -DECG_POOL=1 (scripts/build_variant.sh); tests/test_gpu_pool.py covers it when it is.
#ifndef ECG_POOL
#define ECG_POOL 0
#endif
#if ECG_POOL
#ifndef ECG_POOL_BLOCK
#define ECG_POOL_BLOCK 64
#endif
#ifndef ECG_POOL_PER_SM
#define ECG_POOL_PER_SM (512 / ECG_POOL_BLOCK)
#endif
constexpr int POOL_BLOCK = ECG_POOL_BLOCK, POOL_PER_SM = ECG_POOL_PER_SM;
constexpr int POOL_LISTS = CASCADE_CLASSES + 2; // classes, finished, free
constexpr int POOL_HDR = 64;                    // two sets of list-growth counters
template <class SH>
struct PoolGeo {
    static constexpr int W = SH::G::W;
    static constexpr int SLOT_WORDS = 8 * W + 4; // board 4W, cleared / sp / sk0 / sk1 W each, idx, action, reward, cascades | status
    // 228 KB of shared memory per SM, 1 KB per block reserved, at most 227 KB per block
    static constexpr int per_block = (233472 / POOL_PER_SM - 1024) & ~15;
    static constexpr int BLOCK_BYTES = per_block > 232448 ? 232448 : per_block;
    static constexpr int SLOTS = ((BLOCK_BYTES - POOL_HDR) / (4 * SLOT_WORDS + 2 * POOL_LISTS)) & ~3;
    static constexpr bool OK = SLOTS >= 3 * POOL_BLOCK && SLOTS < 65536;
};

template <class SH>
__global__ void __launch_bounds__(POOL_BLOCK, POOL_PER_SM) pool_step_kernel(RefillDev rf, StepDev io, int types, int n) {
    using G = typename SH::G;
    using PG = PoolGeo<SH>;
    constexpr int W = G::W, NS = PG::SLOTS, KB = POOL_BLOCK;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int OTHER = CASCADE_CLASSES - 1, FIN = CASCADE_CLASSES, FREE = CASCADE_CLASSES + 1;
    extern __shared__ __align__(16) unsigned char pool_smem[];
    int *const delta = reinterpret_cast<int *>(pool_smem); // [2][8]
    uint32_t *const pool = reinterpret_cast<uint32_t *>(pool_smem + POOL_HDR);
    uint16_t *const lists = reinterpret_cast<uint16_t *>(pool + NS * PG::SLOT_WORDS);
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    // slot word offsets
    constexpr int O_BD = 0, O_CL = 4 * W, O_SP = 5 * W, O_K0 = 6 * W, O_K1 = 7 * W, O_IDX = 8 * W, O_ACT = O_IDX + 1,
                  O_REW = O_IDX + 2, O_CS = O_IDX + 3;

    int group = blockIdx.x; // boards [group * KB, (group + 1) * KB), then group += gridDim.x
    const int groups = (n + KB - 1) / KB;

    int cnt[POOL_LISTS]; // list lengths, the same in every thread of the block
#pragma unroll
    for (int c = 0; c < POOL_LISTS; c++) cnt[c] = 0;
    for (int s = tid; s < NS; s += KB) lists[FREE * NS + s] = (uint16_t)s;
    cnt[FREE] = NS;
    if (tid < 16) delta[tid] = 0;
    int cur = 0; // the counter set this pass appends with

    for (;;) {
        __syncthreads(); // the previous pass is complete: its slots and list entries are written
        {
            const int *dl = delta + (cur ^ 1) * 8;
#pragma unroll
            for (int c = 0; c < POOL_LISTS; c++) cnt[c] += dl[c];
        }
        // ---- which phase, on how many slots (the same decision in every thread)
        int phase = -1, k = KB; // phase: 0 .. CLASSES-1 ITERATE that class, FIN, FREE = BEGIN
        if (cnt[FIN] >= KB) {
            phase = FIN;
        } else {
#pragma unroll
            for (int c = 0; c < CASCADE_CLASSES; c++)
                if (phase < 0 && cnt[c] >= KB) phase = c;
            if (phase < 0) {
                if (group < groups && cnt[FREE] >= KB) {
                    phase = FREE;
                } else {
                    int best = 0;
#pragma unroll
                    for (int c = 0; c < CASCADE_CLASSES; c++)
                        if (cnt[c] > best) {
                            best = cnt[c];
                            phase = c;
                        }
                    if (phase < 0) {
                        if (cnt[FIN] == 0) break;
                        phase = FIN;
                        best = cnt[FIN];
                    }
                    k = best; // < KB
                }
            }
        }
        int idx = 0;
        if (phase == FREE) {
            const int base = group * KB;
            k = n - base < KB ? n - base : KB;
            idx = base + tid;
            group += gridDim.x;
            if (group < groups) { // the lines of this block's next boards
                constexpr int CH = G::W;
                const int pf = group * KB + tid;
                const uint4 *bp = reinterpret_cast<const uint4 *>(io.boards_in) + (long long)(pf >> 5) * (CH * 32) + (pf & 31);
                if (pf < n) {
#pragma unroll
                    for (int q = 0; q < CH; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(bp + q * 32));
                }
            }
        }
        // ---- take k entries off the phase's list
        int have = 0;
#pragma unroll
        for (int c = 0; c < POOL_LISTS; c++)
            if (phase == c) {
                have = cnt[c];
                cnt[c] -= k;
            }
        const bool on = tid < k;
        const int id = on ? (int)lists[phase * NS + have - k + tid] : 0;
        uint32_t *const sl = pool + id;
        __syncthreads(); // every thread has read the lists: this pass may append to them
        if (tid < 8) delta[(cur ^ 1) * 8 + tid] = 0;

        Lane<SH> L;
        BB<W> HL, VL;
        int dest = -1, action = -1, moves = 0;
        bool handoff = false;
        uint32_t st = 0;
        if (on) {
            if (phase == FREE) { // FETCH
                load_board<G>(io.boards_in, idx, L.bd);
                moves = io.moves_left ? io.moves_left[idx] : 1;
                if (!io.actions) load_mask<G>(io.mask_in, idx, HL, VL);
                L.reward = 0;
                L.cascades = 0;
                L.status = 0;
            } else {
#pragma unroll
                for (int p = 0; p < 4; p++)
#pragma unroll
                    for (int j = 0; j < W; j++) L.bd.p[p].w[j] = sl[(O_BD + p * W + j) * NS];
                idx = (int)sl[O_IDX * NS];
                L.reward = (int)sl[O_REW * NS];
                const uint32_t cs = sl[O_CS * NS];
                L.cascades = (int)(cs & 0xffffu);
                L.status = cs >> 16;
            }
        }
        if (phase != FIN) {
            // the Philox block of this cascade iteration (block 0 holds the action pick and the first refill words)
            uint32_t blk[4];
            const uint32_t blk_index = phase == FREE ? 0u : (uint32_t)L.cascades * 512u;
            const unsigned long long board = rf.board0 + (unsigned long long)idx;
            philox4x32_10(blk_index, rf.step_ctr, (uint32_t)board, (uint32_t)(board >> 32), (uint32_t)rf.key,
                          (uint32_t)(rf.key >> 32), blk);
            if (phase == FREE) {
                // ---- BEGIN
                if (on) {
                    int b1 = -1, d = 0;
                    if (moves < 1) { // boardv2.py:44
                        st = ST_TERMINAL;
                    } else if (io.actions) {
                        action = io.actions[idx];
                        if (action < 0 || action >= G::A) {
                            st = ST_BAD_ACTION;
                            action = -1;
                        }
                    } else { // board.random_action(): uniform over the legal set of the current board
                        const int c = swaps_count<G>(HL, VL);
                        if (c == 0) {
                            st = ST_NO_LEGAL;
                        } else {
                            bool vertical;
                            b1 = swaps_select_bit<G>(HL, VL, (int)mulhi32(blk[0], (uint32_t)c), vertical);
                            d = vertical ? G::S : 1;
                            action = action_of_swap<G>(b1, vertical);
                        }
                    }
                    if (action >= 0) {
                        if (b1 < 0) { // action given by the caller: boardConfig.decode
                            int b2;
                            decode_action<G>(action, b1, b2);
                            d = b2 - b1;
                        }
                        handoff = step_begin_at<SH, true>(L, b1, d);
                        dest = handoff ? FREE : cascade_class<SH>(L);
                    } else { // a no-op board: FINISH writes it (action -1, nothing stepped)
                        L.status = st;
                        dest = FIN;
                    }
                }
            } else {
                // ---- ITERATE: one cascade iteration on slots of class `phase`
                if (on) {
#pragma unroll
                    for (int j = 0; j < W; j++) L.cleared.w[j] = sl[(O_CL + j) * NS];
                    if (phase == OTHER) {
#pragma unroll
                        for (int j = 0; j < W; j++) {
                            L.sp.w[j] = sl[(O_SP + j) * NS];
                            L.sk0.w[j] = sl[(O_K0 + j) * NS];
                            L.sk1.w[j] = sl[(O_K1 + j) * NS];
                        }
                    } else {
                        L.sp = bb_zero<W>();
                        L.sk0 = bb_zero<W>();
                        L.sk1 = bb_zero<W>();
                    }
                    PhiloxRng rng;
                    rng.init(rf.key, rf.board0 + (unsigned long long)idx, rf.step_ctr);
                    rng.preset_block(blk_index, blk);
                    const bool fin = step_iter<SH, PhiloxRng, true, true>(L, rng, (uint32_t)types, HL, VL, handoff);
                    dest = handoff ? FREE : fin ? FIN : cascade_class<SH>(L);
                }
            }
            // ---- keep the lane in its slot
            if (dest >= 0 && dest <= FIN) {
#pragma unroll
                for (int p = 0; p < 4; p++)
#pragma unroll
                    for (int j = 0; j < W; j++) sl[(O_BD + p * W + j) * NS] = L.bd.p[p].w[j];
                sl[O_REW * NS] = (uint32_t)L.reward;
                sl[O_CS * NS] = (uint32_t)L.cascades | (L.status << 16);
                if (phase == FREE) {
                    sl[O_IDX * NS] = (uint32_t)idx;
                    sl[O_ACT * NS] = (uint32_t)action;
                }
                if (dest < FIN) {
#pragma unroll
                    for (int j = 0; j < W; j++) sl[(O_CL + j) * NS] = L.cleared.w[j];
                    if (dest == OTHER) {
#pragma unroll
                        for (int j = 0; j < W; j++) {
                            sl[(O_SP + j) * NS] = L.sp.w[j];
                            sl[(O_K0 + j) * NS] = L.sk0.w[j];
                            sl[(O_K1 + j) * NS] = L.sk1.w[j];
                        }
                    }
                }
            }
        } else if (on) {
            // ---- FINISH: legal swaps of the final board, outputs
            action = (int)sl[O_ACT * NS];
            legal_of<SH>(L.bd, HL, VL);
            dest = FREE;
            if (!any(HL | VL)) { // the shuffle loop (or a no-op board without a legal swap): exact kernel
                handoff = true;
            } else {
                const bool stepped = action >= 0;
                moves = io.moves_left ? io.moves_left[idx] - (stepped ? 1 : 0) : (stepped ? 0 : 1);
                finish_board<SH, true>(rf, io, idx, idx, L.bd, stepped, action, moves, L.reward, L.cascades, L.status, HL,
                                       VL, 0u);
            }
        }
        // ---- list the slot, hand boards over
        {
            // one shared-memory atomic per thread: the LSU pipe idles in this kernel, the ALU pipe is what it runs on
            // (warp-aggregated appends cost ~40 ALU instructions per list)
            if (dest >= 0) {
                int at = 0;
#pragma unroll
                for (int c = 0; c < POOL_LISTS; c++)
                    if (dest == c) at = c * NS + cnt[c];
                lists[at + atomicAdd(delta + cur * 8 + dest, 1)] = (uint16_t)id;
            }
            const unsigned hm = __ballot_sync(FULL, handoff);
            if (hm) {
                int base = 0;
                if (lane == __ffs((int)hm) - 1) base = atomicAdd(io.handoff, __popc(hm));
                base = __shfl_sync(FULL, base, __ffs((int)hm) - 1);
                if (handoff) io.handoff[1 + base + __popc(hm & lt)] = idx;
            }
        }
        cur ^= 1;
    }
}

#endif // ECG_POOL

// ------------------------------------------------------------------ launchers

inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

using SHN = Shape<ECG_SIZE, ECG_SIZE, 3, false>; // types <= 7
using SHW = Shape<ECG_SIZE, ECG_SIZE, 4, true>;  // types 8..11

void op_pack(bool wide, const void *cells, int eb, void *boards, uint8_t *status, int types, long long n,
             cudaStream_t s) {
    const unsigned g = grid_for(n, BLOCK);
    if (eb == 8) {
        if (wide) pack_kernel<SHW, long long><<<g, BLOCK, 0, s>>>((const long long *)cells, boards, status, types, n);
        else pack_kernel<SHN, long long><<<g, BLOCK, 0, s>>>((const long long *)cells, boards, status, types, n);
    } else {
        if (wide) pack_kernel<SHW, uint8_t><<<g, BLOCK, 0, s>>>((const uint8_t *)cells, boards, status, types, n);
        else pack_kernel<SHN, uint8_t><<<g, BLOCK, 0, s>>>((const uint8_t *)cells, boards, status, types, n);
    }
}
void op_unpack(bool wide, const void *boards, void *cells, int eb, int types, long long n, cudaStream_t s) {
    const unsigned g = grid_for((n + 31) / 32, BLOCK / 32); // one warp per tile
    if (eb == 8) {
        if (wide) unpack_kernel<SHW, long long><<<g, BLOCK, 0, s>>>(boards, (long long *)cells, types, n);
        else unpack_kernel<SHN, long long><<<g, BLOCK, 0, s>>>(boards, (long long *)cells, types, n);
    } else if ((reinterpret_cast<uintptr_t>(cells) & 3u) == 0) {
        if (wide) unpack_kernel<SHW, uint8_t><<<g, BLOCK, 0, s>>>(boards, (uint8_t *)cells, types, n);
        else unpack_kernel<SHN, uint8_t><<<g, BLOCK, 0, s>>>(boards, (uint8_t *)cells, types, n);
    } else { // unaligned byte output: one thread per board
        const unsigned g1 = grid_for(n, BLOCK);
        if (wide) unpack_kernel_simple<SHW, uint8_t><<<g1, BLOCK, 0, s>>>(boards, (uint8_t *)cells, types, n);
        else unpack_kernel_simple<SHN, uint8_t><<<g1, BLOCK, 0, s>>>(boards, (uint8_t *)cells, types, n);
    }
}
void op_unpack_nibbles(const void *boards, uint8_t *out, long long n, cudaStream_t s) {
    unpack_nibbles_kernel<SHN><<<grid_for((n + 31) / 32, BLOCK / 32), BLOCK, 0, s>>>(boards, out, n);
}
void op_unpack_mask(const uint32_t *mask, uint8_t *out, long long n, cudaStream_t s) {
    unpack_mask_kernel<SHN><<<grid_for(n * SHN::G::A, BLOCK), BLOCK, 0, s>>>(mask, out, n);
}
void op_init(bool wide, bool philox, RefillDev rf, void *boards, uint8_t *status, int types, long long n,
             cudaStream_t s) {
    const unsigned g = grid_for(n, BLOCK);
    if (wide) {
        if (philox) init_kernel<SHW, true><<<g, BLOCK, 0, s>>>(rf, boards, status, types, n);
        else init_kernel<SHW, false><<<g, BLOCK, 0, s>>>(rf, boards, status, types, n);
    } else {
        if (philox) init_kernel<SHN, true><<<g, BLOCK, 0, s>>>(rf, boards, status, types, n);
        else init_kernel<SHN, false><<<g, BLOCK, 0, s>>>(rf, boards, status, types, n);
    }
}
void op_legal(bool wide, const void *boards, uint32_t *mask, long long n, cudaStream_t s) {
    const unsigned g = grid_for(n, BLOCK);
    if (wide) legal_kernel<SHW><<<g, BLOCK, 0, s>>>(boards, mask, n);
    else legal_kernel<SHN><<<g, BLOCK, 0, s>>>(boards, mask, n);
}
void op_random_action(bool philox, RefillDev rf, const uint32_t *mask, int32_t *actions, uint8_t *status, long long n,
                      cudaStream_t s) {
    const unsigned g = grid_for(n, BLOCK);
    if (philox) random_action_kernel<SHN, true><<<g, BLOCK, 0, s>>>(rf, mask, actions, status, n);
    else random_action_kernel<SHN, false><<<g, BLOCK, 0, s>>>(rf, mask, actions, status, n);
}
// persistent grid: all resident warps, but no more than there are chunks
template <class K>
unsigned persistent_grid(K kernel, long long n, int block) {
    static int resident = 0; // per kernel instantiation
    if (resident == 0) {
        int dev = 0, sms = 148, per_sm = 4;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
        resident = sms * (per_sm > 0 ? per_sm : 1);
    }
    const long long chunks = (n + 31) / 32; // the kernel shrinks its chunks down to one tile for small batches
    const long long blocks = (chunks + (block / 32) - 1) / (block / 32);
    return (unsigned)(blocks < resident ? blocks : resident);
}

template <class SH, bool PHILOX, bool ROLLOUT, bool FAST = false>
void launch_lanes(RefillDev rf, StepDev io, RolloutDev ro, int types, long long n, cudaStream_t s) {
    constexpr int B = lane_block(FAST);
    lane_kernel<SH, PHILOX, ROLLOUT, FAST>
        <<<persistent_grid(lane_kernel<SH, PHILOX, ROLLOUT, FAST>, n, B), B, 0, s>>>(rf, io, ro, types, (int)n);
}

// The exact kernel over the hand-off list of a two-kernel STEP.  The list is short (0.6 % of the boards) and its boards
// cascade long (4 iterations on average against 1.5): with all 16 warps per SM every warp got 40 jobs for its 32 lanes,
// ran to its slowest lane (7.6 trips at 17 active lanes) and shared the issue slots with 15 others.  Fewer, smaller
// blocks give every warp more jobs to refill its lanes with and faster trips.
#ifndef ECG_JOBS_BLOCK
#define ECG_JOBS_BLOCK lane_block(false)
#endif
#ifndef ECG_JOBS_PER_SM
#define ECG_JOBS_PER_SM 1
#endif
template <class SH, bool PHILOX>
void launch_jobs(RefillDev rf, StepDev io, int types, long long n, cudaStream_t s) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const RolloutDev ro = {};
    lane_kernel<SH, PHILOX, false, false><<<sms * ECG_JOBS_PER_SM, ECG_JOBS_BLOCK, 0, s>>>(rf, io, ro, types, (int)n);
}

#if ECG_POOL
// The pooled common-case kernel: POOL_PER_SM blocks per SM; a block needs a few dozen passes to amortise filling and
// draining its pool, so small batches stay with the lane kernel
#ifndef ECG_POOL_MIN_GROUPS
#define ECG_POOL_MIN_GROUPS 16
#endif
template <class SH>
bool launch_pool_step(RefillDev rf, StepDev io, long long n, cudaStream_t s) {
    using PG = PoolGeo<SH>;
    constexpr int SMEM = PG::BLOCK_BYTES;
    static int blocks = 0;
    if (blocks == 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(pool_step_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        blocks = sms * POOL_PER_SM;
    }
    long long min_boards = (long long)blocks * POOL_BLOCK * ECG_POOL_MIN_GROUPS;
    if (const char *e = getenv("ECG_POOL_MIN_BOARDS")) min_boards = atoll(e); // tests: 0 runs every batch through the pool
    if (n < min_boards) return false;
    pool_step_kernel<SH><<<blocks, POOL_BLOCK, SMEM, s>>>(rf, io, io.types, (int)n);
    return true;
}
#endif // ECG_POOL

// The two-kernel step (Philox mode, or replay mode with tile tables): the common-case kernel over all boards, then the
// exact kernel over the jobs it handed off (their number is read on the device: no host synchronisation in between)
template <class SH, bool PHILOX>
void launch_two_kernel_step(RefillDev rf, StepDev io, long long n, cudaStream_t s) {
    const RolloutDev ro = {};
    cudaMemsetAsync(io.handoff, 0, sizeof(int32_t), s);
    bool pooled = false;
#if ECG_POOL
    if constexpr (PHILOX && PoolGeo<SH>::OK) {
        if (!io.src_index) pooled = launch_pool_step<SH>(rf, io, n, s);
    }
#endif
    if (!pooled) launch_lanes<SH, PHILOX, false, true>(rf, io, ro, io.types, n, s);
    if (io.mid_event) cudaEventRecord((cudaEvent_t)io.mid_event, s);
    StepDev io2 = io;
    io2.jobs = io.handoff + 1;
    io2.n_jobs = io.handoff;
    io2.handoff = nullptr;
    launch_jobs<SH, PHILOX>(rf, io2, io.types, n, s);
}

int op_step(bool wide, bool philox, RefillDev rf, StepDev io, long long n, cudaStream_t s) {
    const RolloutDev ro = {};
    // (the job cursor's 32-bit arithmetic: 2 n + warps < 2^31)
    if (io.handoff && n <= (1ll << 29) && (philox || (rf.tiles && rf.tile_wpos))) {
        if (philox) {
            if (wide) launch_two_kernel_step<SHW, true>(rf, io, n, s);
            else launch_two_kernel_step<SHN, true>(rf, io, n, s);
        } else {
            if (wide) launch_two_kernel_step<SHW, false>(rf, io, n, s);
            else launch_two_kernel_step<SHN, false>(rf, io, n, s);
        }
        return 2;
    }
    io.handoff = nullptr;
    if (wide) {
        if (philox) launch_lanes<SHW, true, false>(rf, io, ro, io.types, n, s);
        else launch_lanes<SHW, false, false>(rf, io, ro, io.types, n, s);
    } else {
        if (philox) launch_lanes<SHN, true, false>(rf, io, ro, io.types, n, s);
        else launch_lanes<SHN, false, false>(rf, io, ro, io.types, n, s);
    }
    return 1;
}
// Philox rollouts with a work list: the common-case kernel plays the episodes; an episode that meets a rare case is
// parked (board of the step's start, reward and step count so far) and finished by the exact kernel
template <class SH>
void launch_two_kernel_rollout(RefillDev rf, RolloutDev ro, int32_t *handoff, int types, long long n, cudaStream_t s) {
    cudaMemsetAsync(handoff, 0, sizeof(int32_t), s);
    StepDev io = {};
    io.handoff = handoff;
    launch_lanes<SH, true, true, true>(rf, io, ro, types, n, s);
    StepDev io2 = {};
    io2.jobs = handoff + 1;
    io2.n_jobs = handoff;
    launch_lanes<SH, true, true, false>(rf, io2, ro, types, n, s);
}

int op_rollout(bool wide, bool philox, RefillDev rf, void *boards, const int32_t *moves_left, long long *total_reward,
               int32_t *steps_done, uint8_t *status, int32_t *scratch, int types, long long n, cudaStream_t s) {
    const StepDev io = {};
    const RolloutDev ro = {boards, moves_left, total_reward, steps_done, status};
    if (philox && scratch && steps_done && n <= (1ll << 29)) {
        if (wide) launch_two_kernel_rollout<SHW>(rf, ro, scratch, types, n, s);
        else launch_two_kernel_rollout<SHN>(rf, ro, scratch, types, n, s);
        return 2;
    }
    if (wide) {
        if (philox) launch_lanes<SHW, true, true>(rf, io, ro, types, n, s);
        else launch_lanes<SHW, false, true>(rf, io, ro, types, n, s);
    } else {
        if (philox) launch_lanes<SHN, true, true>(rf, io, ro, types, n, s);
        else launch_lanes<SHN, false, true>(rf, io, ro, types, n, s);
    }
    return 1;
}

void op_onehot(const void *boards, void *out, int channels, int elem_kind, int types, long long n, cudaStream_t s) {
    const unsigned g = grid_for(n * SHN::G::R * SHN::G::C, BLOCK);
    if (elem_kind == 0) onehot_kernel<SHN, uint8_t><<<g, BLOCK, 0, s>>>(boards, (uint8_t *)out, channels, (uint8_t)1, types, n);
    else if (elem_kind == 1) onehot_kernel<SHN, uint32_t><<<g, BLOCK, 0, s>>>(boards, (uint32_t *)out, channels, 0x3F800000u, types, n);
    else onehot_kernel<SHN, uint16_t><<<g, BLOCK, 0, s>>>(boards, (uint16_t *)out, channels,
                                                          (uint16_t)(elem_kind == 2 ? 0x3F80 : 0x3C00), types, n);
}

void op_augment(const void *boards_in, void *boards_out, bool mirror, bool remap, CodeLut lut, long long n,
                cudaStream_t s) {
    augment_kernel<SHN><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(boards_in, boards_out, mirror, remap, lut, n);
}

const ShapeOps k_ops = {op_pack, op_unpack, op_unpack_nibbles, op_unpack_mask, op_init, op_legal, op_random_action, op_step, op_rollout,
                        op_onehot, op_augment};

} // namespace

#define ECG_CAT2(a, b) a##b
#define ECG_CAT(a, b) ECG_CAT2(a, b)
namespace ecg {
const ShapeOps *ECG_CAT(shape_ops_, ECG_SIZE)() { return &k_ops; }
} // namespace ecg
