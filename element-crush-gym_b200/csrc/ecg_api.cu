// ecg_api.cu -- the C-ABI of libecg.so (include/ecg.h): argument checks, dispatch to the per-size
// kernel objects (ecg_shape_kernels.cu), and the two size-independent kernels.
#include <cuda_runtime.h>

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ecg.h"
#include "ecg_core.cuh"
#include "ecg_ops.h"

using namespace ecg;

static_assert(ECG_ST_TERMINAL == ST_TERMINAL && ECG_ST_STREAM_OVERFLOW == ST_STREAM_OVERFLOW &&
                  ECG_ST_SHUFFLE_CAP == ST_SHUFFLE_CAP && ECG_ST_BAD_ACTION == ST_BAD_ACTION &&
                  ECG_ST_NO_LEGAL == ST_NO_LEGAL && ECG_ST_BAD_CELL == ST_BAD_CELL &&
                  ECG_ST_CASCADE_CAP == ST_CASCADE_CAP,
              "status bits of include/ecg.h and ecg_core.cuh must agree");

namespace {

thread_local char g_err[256] = "";
std::atomic<long long> g_launches{0};
thread_local void *g_mid_event = nullptr; // ecg_step_mark_event

int fail(const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return -1;
}
int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// numpy legacy MT19937: init_genrand(seed) then genrand_int32; thread-local state
__global__ void __launch_bounds__(64) mt19937_kernel(const uint32_t *__restrict__ seeds, uint32_t *__restrict__ out,
                                                     int len, long long n) {
    const long long i = (long long)blockIdx.x * 64 + threadIdx.x;
    if (i >= n) return;
    uint32_t mt[624];
    mt[0] = seeds[i];
    for (int k = 1; k < 624; k++) mt[k] = 1812433253u * (mt[k - 1] ^ (mt[k - 1] >> 30)) + (uint32_t)k;
    int mti = 624;
    uint32_t *dst = out + i * len;
    for (int j = 0; j < len; j++) {
        if (mti >= 624) {
            for (int k = 0; k < 624; k++) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        dst[j] = y;
    }
}

// per-stream tile tables of the replay two-kernel step (ecg_core.cuh build_replay_tiles): one thread per stream
__global__ void __launch_bounds__(64) replay_tiles_kernel(const uint32_t *__restrict__ raw, int len, int types,
                                                          uint32_t *__restrict__ tiles, uint16_t *__restrict__ wpos,
                                                          long long n) {
    const long long i = (long long)blockIdx.x * 64 + threadIdx.x;
    if (i >= n) return;
    build_replay_tiles(raw + i * len, len, (uint32_t)types, tiles + i * replay_tile_words(len), wpos + i * (len + 1));
}

__global__ void __launch_bounds__(256) stats_kernel(const int32_t *__restrict__ score, const uint8_t *__restrict__ flags,
                                                    long long *out, long long n) {
    long long sum = 0, sq = 0, cnt = 0, wins = 0, mn = 0x7fffffffffffffffLL, mx = -0x7fffffffffffffffLL - 1;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long s = score[i];
        sum += s;
        sq += s * s;
        cnt++;
        mn = s < mn ? s : mn;
        mx = s > mx ? s : mx;
        if (flags && (flags[i] & ECG_FLAG_WON)) wins++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        sq += __shfl_down_sync(0xffffffffu, sq, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        wins += __shfl_down_sync(0xffffffffu, wins, o);
        const long long a = __shfl_down_sync(0xffffffffu, mn, o), b = __shfl_down_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *)&out[0], (unsigned long long)sum);
        atomicAdd((unsigned long long *)&out[1], (unsigned long long)cnt);
        atomicMin(&out[2], mn);
        atomicMax(&out[3], mx);
        atomicAdd((unsigned long long *)&out[4], (unsigned long long)wins);
        atomicAdd((unsigned long long *)&out[5], (unsigned long long)sq);
    }
}


RefillDev to_dev(const ecg_refill *rf) {
    RefillDev d;
    d.stream = rf->stream;
    d.stream_stride = rf->stream_stride;
    d.stream_pos = rf->stream_pos;
    d.key = rf->philox_key;
    d.board0 = rf->board0;
    d.step_ctr = rf->step_ctr;
    d.stream_len = rf->stream_len;
    d.stream_index = rf->stream_index;
    d.tiles = rf->tiles;
    d.tile_wpos = rf->tile_wpos;
    return d;
}

// -DECG_ONLY_SIZE=N: a one-size experiment build (scripts/build_variant.sh)
const ShapeOps *ops_for(int rows) {
    switch (rows) {
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 4
    case 4: return shape_ops_4();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 5
    case 5: return shape_ops_5();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 6
    case 6: return shape_ops_6();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 7
    case 7: return shape_ops_7();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 8
    case 8: return shape_ops_8();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 9
    case 9: return shape_ops_9();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 10
    case 10: return shape_ops_10();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 11
    case 11: return shape_ops_11();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 12
    case 12: return shape_ops_12();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 13
    case 13: return shape_ops_13();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 14
    case 14: return shape_ops_14();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 15
    case 15: return shape_ops_15();
#endif
#if !defined(ECG_ONLY_SIZE) || ECG_ONLY_SIZE == 16
    case 16: return shape_ops_16();
#endif
    default: return nullptr;
    }
}

int check_cfg(const ecg_config *cfg) {
    if (!cfg) return fail("cfg is NULL");
    ecg_config ref;
    if (ecg_config_init(&ref, cfg->rows, cfg->cols, cfg->types) != 0) return -1;
    if (memcmp(&ref, cfg, sizeof(ref)) != 0) return fail("ecg_config was not produced by ecg_config_init");
    return 0;
}
int check_refill(const ecg_refill *rf) {
    if (!rf) return fail("refill is NULL");
    if (rf->mode == ECG_REFILL_PHILOX) return 0;
    if (rf->mode == ECG_REFILL_REPLAY) {
        if (!rf->stream || rf->stream_len <= 0 || rf->stream_stride < 0)
            return fail("replay refill needs stream, stream_len > 0, stream_stride >= 0");
        if ((rf->tiles != nullptr) != (rf->tile_wpos != nullptr) || (rf->tiles && rf->stream_len >= (int32_t)REPLAY_TILES_END))
            return fail("replay refill: tiles and tile_wpos come together (ecg_replay_tiles, stream_len < 65535)");
        if (rf->tiles && rf->stream_stride != 0 && rf->stream_stride != rf->stream_len)
            return fail("replay refill: tile tables need stream_stride == 0 or == stream_len");
        return 0;
    }
    return fail("refill mode must be ECG_REFILL_REPLAY or ECG_REFILL_PHILOX");
}

} // namespace

extern "C" {

int ecg_version(void) { return ECG_VERSION; }
int ecg_sizeof(int which) {
    switch (which) {
    case ECG_SIZEOF_CONFIG: return (int)sizeof(ecg_config);
    case ECG_SIZEOF_REFILL: return (int)sizeof(ecg_refill);
    case ECG_SIZEOF_STEP_IO: return (int)sizeof(ecg_step_io);
    default: return -1;
    }
}
const char *ecg_last_error(void) { return g_err; }
int64_t ecg_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int ecg_config_init(ecg_config *cfg, int rows, int cols, int types) {
    if (!cfg) return fail("cfg is NULL");
    if (rows != cols)
        return fail("only square boards: the reference's action space rows*(cols-1)*2 (boardConfig.py:27) is wrong otherwise "
                    "and its own front end only builds (height, height) boards (main.py:168)");
    if (!ops_for(rows)) return fail("board size must be in 4..16 (boardConfig.decode needs columns >= 4; 16 is the engine's limit)");
    if (types < 1 || types > 11) return fail("types must be in 1..11 (4-bit cell codes)");
    memset(cfg, 0, sizeof(*cfg));
    cfg->rows = rows;
    cfg->cols = cols;
    cfg->types = types;
    int bits = 0;
    while ((1 << bits) < types + 1) bits++;
    cfg->bits = bits;
    cfg->type_mask = (1 << bits) - 1;
    cfg->special_type_mask = (1 << (bits + 1)) + 1 + cfg->type_mask;
    cfg->h_line = cfg->type_mask + 1;
    cfg->v_line = 2 * cfg->h_line;
    cfg->bomb = cfg->special_type_mask;
    cfg->mega_token = cfg->type_mask + cfg->special_type_mask + 1;
    cfg->action_space = rows * (cols - 1) * 2;
    cfg->board_words = 4 * ((rows * (cols + 1) + 31) / 32);
    cfg->mask_words = cfg->board_words / 2; /* two swap bitboards (horizontal, vertical) */
    return 0;
}

int64_t ecg_boards_bytes(const ecg_config *cfg, int64_t n) {
    return ((n + ECG_TILE - 1) / ECG_TILE) * ECG_TILE * (int64_t)cfg->board_words * 4;
}
int64_t ecg_masks_bytes(const ecg_config *cfg, int64_t n) {
    return ((n + ECG_TILE - 1) / ECG_TILE) * ECG_TILE * (int64_t)cfg->mask_words * 4;
}

int ecg_pack(const ecg_config *cfg, const void *cells, int elem_bytes, void *boards, uint8_t *status, int64_t n,
             void *stream) {
    if (check_cfg(cfg)) return -1;
    if (!cells || !boards) return fail("ecg_pack: NULL buffer");
    if (elem_bytes != 8 && elem_bytes != 1) return fail("ecg_pack: elem_bytes must be 8 (int64) or 1 (uint8)");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->pack(cfg->types >= 8, cells, elem_bytes, boards, status, cfg->types, n, (cudaStream_t)stream);
    return check_launch("ecg_pack");
}

int ecg_unpack(const ecg_config *cfg, const void *boards, void *cells, int elem_bytes, int64_t n, void *stream) {
    if (check_cfg(cfg)) return -1;
    if (!cells || !boards) return fail("ecg_unpack: NULL buffer");
    if (elem_bytes != 8 && elem_bytes != 1) return fail("ecg_unpack: elem_bytes must be 8 (int64) or 1 (uint8)");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->unpack(cfg->types >= 8, boards, cells, elem_bytes, cfg->types, n, (cudaStream_t)stream);
    return check_launch("ecg_unpack");
}

int ecg_unpack_nibbles(const ecg_config *cfg, const void *boards, uint8_t *out, int64_t n, void *stream) {
    if (check_cfg(cfg)) return -1;
    if (!out || !boards) return fail("ecg_unpack_nibbles: NULL buffer");
    if (reinterpret_cast<uintptr_t>(out) & 3u) return fail("ecg_unpack_nibbles: out must be 4-byte aligned");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->unpack_nibbles(boards, out, n, (cudaStream_t)stream);
    return check_launch("ecg_unpack_nibbles");
}

int ecg_unpack_mask(const ecg_config *cfg, const uint32_t *mask, uint8_t *out, int64_t n, void *stream) {
    if (check_cfg(cfg)) return -1;
    if (!mask || !out) return fail("ecg_unpack_mask: NULL buffer");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->unpack_mask(mask, out, n, (cudaStream_t)stream);
    return check_launch("ecg_unpack_mask");
}

int ecg_mt19937_stream(const uint32_t *seeds, uint32_t *out, int32_t len, int64_t n, void *stream) {
    if (!seeds || !out || len <= 0) return fail("ecg_mt19937_stream: bad argument");
    if (n <= 0) return 0;
    mt19937_kernel<<<grid_for(n, 64), 64, 0, (cudaStream_t)stream>>>(seeds, out, len, n);
    return check_launch("ecg_mt19937_stream");
}

int64_t ecg_replay_tiles_words(int32_t stream_len) { return stream_len > 0 ? replay_tile_words(stream_len) : 0; }

int ecg_replay_tiles(const uint32_t *streams, int32_t stream_len, int types, uint32_t *tiles, uint16_t *tile_wpos,
                     int64_t n_streams, void *stream) {
    if (!streams || !tiles || !tile_wpos) return fail("ecg_replay_tiles: NULL buffer");
    if (stream_len <= 0 || stream_len >= (int32_t)REPLAY_TILES_END)
        return fail("ecg_replay_tiles: stream_len must be in 1..65534 (16-bit word positions)");
    if (types < 1 || types > 11) return fail("ecg_replay_tiles: types must be in 1..11");
    if (n_streams <= 0) return 0;
    replay_tiles_kernel<<<grid_for(n_streams, 64), 64, 0, (cudaStream_t)stream>>>(streams, stream_len, types, tiles,
                                                                                  tile_wpos, n_streams);
    return check_launch("ecg_replay_tiles");
}

int ecg_init_boards(const ecg_config *cfg, const ecg_refill *rf, void *boards, uint8_t *status, int64_t n,
                    void *stream) {
    if (check_cfg(cfg) || check_refill(rf)) return -1;
    if (!boards) return fail("ecg_init_boards: NULL buffer");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->init(cfg->types >= 8, rf->mode == ECG_REFILL_PHILOX, to_dev(rf), boards, status, cfg->types, n,
                             (cudaStream_t)stream);
    return check_launch("ecg_init_boards");
}

int ecg_legal_mask(const ecg_config *cfg, const void *boards, uint32_t *mask, int64_t n, void *stream) {
    if (check_cfg(cfg)) return -1;
    if (!boards || !mask) return fail("ecg_legal_mask: NULL buffer");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->legal(cfg->types >= 8, boards, mask, n, (cudaStream_t)stream);
    return check_launch("ecg_legal_mask");
}

int ecg_random_action(const ecg_config *cfg, const ecg_refill *rf, const uint32_t *mask, int32_t *actions,
                      uint8_t *status, int64_t n, void *stream) {
    if (check_cfg(cfg) || check_refill(rf)) return -1;
    if (!mask || !actions) return fail("ecg_random_action: NULL buffer");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->random_action(rf->mode == ECG_REFILL_PHILOX, to_dev(rf), mask, actions, status, n,
                                      (cudaStream_t)stream);
    return check_launch("ecg_random_action");
}

int ecg_step_mark_event(void *event) {
    g_mid_event = event;
    return 0;
}

int ecg_step(const ecg_config *cfg, const ecg_refill *rf, const ecg_step_io *io, int64_t n, void *stream) {
    if (check_cfg(cfg) || check_refill(rf)) return -1;
    if (!io || !io->boards_in || !io->boards_out) return fail("ecg_step: boards_in/boards_out are required");
    if (!io->actions && !io->mask_in) return fail("ecg_step: actions == NULL needs mask_in (random legal action)");
    if (io->flags && !(io->score && io->moves_left)) return fail("ecg_step: flags need score and moves_left");
    if (n <= 0) return 0;
    if (n > (1ll << 30)) return fail("ecg_step: at most 2^30 boards per call");
    if (io->src_index) {
        // job i reads board src_index[i] and writes every output at i: an in/out array would be read at one index
        // while another job writes it
        if (io->boards_out == io->boards_in) return fail("ecg_step: src_index needs boards_out != boards_in");
        if (io->moves_left || io->score)
            return fail("ecg_step: src_index cannot be combined with the in/out arrays moves_left / score");
        if (rf->mode == ECG_REFILL_REPLAY && rf->stream_pos && !io->actions)
            return fail("ecg_step: src_index with random picks (actions == NULL) cannot take an in/out stream_pos");
    }
    StepDev sd;
    sd.boards_in = io->boards_in;
    sd.boards_out = io->boards_out;
    sd.actions = io->actions;
    sd.mask_in = io->mask_in;
    sd.actions_out = io->actions_out;
    sd.moves_left = io->moves_left;
    sd.reward = io->reward;
    sd.score = io->score;
    sd.cascades = io->cascades;
    sd.mask_out = io->mask_out;
    sd.flags = io->flags;
    sd.status = io->status;
    sd.env_goal = io->env_goal;
    sd.types = cfg->types;
    sd.src_index = io->src_index;
    sd.handoff = io->scratch;
    sd.jobs = nullptr;
    sd.n_jobs = nullptr;
    sd.mid_event = g_mid_event;
    g_mid_event = nullptr;
    const int launched =
        ops_for(cfg->rows)->step(cfg->types >= 8, rf->mode == ECG_REFILL_PHILOX, to_dev(rf), sd, n, (cudaStream_t)stream);
    g_launches.fetch_add(launched - 1, std::memory_order_relaxed);
    return check_launch("ecg_step");
}

int ecg_rollout_scratch(const ecg_config *cfg, const ecg_refill *rf, void *boards, const int32_t *moves_left,
                        int64_t *total_reward, int32_t *steps_done, uint8_t *status, int32_t *scratch, int64_t n,
                        void *stream) {
    if (check_cfg(cfg) || check_refill(rf)) return -1;
    if (!boards || !moves_left || !total_reward) return fail("ecg_rollout: boards, moves_left and total_reward are required");
    if (n <= 0) return 0;
    if (n > (1ll << 30)) return fail("ecg_rollout: at most 2^30 boards per call");
    const int launched = ops_for(cfg->rows)->rollout(cfg->types >= 8, rf->mode == ECG_REFILL_PHILOX, to_dev(rf), boards,
                                                     moves_left, (long long *)total_reward, steps_done, status, scratch,
                                                     cfg->types, n, (cudaStream_t)stream);
    g_launches.fetch_add(launched - 1, std::memory_order_relaxed);
    return check_launch("ecg_rollout");
}

int ecg_rollout(const ecg_config *cfg, const ecg_refill *rf, void *boards, const int32_t *moves_left,
                int64_t *total_reward, int32_t *steps_done, uint8_t *status, int64_t n, void *stream) {
    return ecg_rollout_scratch(cfg, rf, boards, moves_left, total_reward, steps_done, status, nullptr, n, stream);
}

int ecg_observe_onehot(const ecg_config *cfg, const void *boards, void *out, int channels, int elem_kind, int64_t n,
                       void *stream) {
    if (check_cfg(cfg)) return -1;
    if (!boards || !out) return fail("ecg_observe_onehot: NULL buffer");
    if (channels < 1 || channels > 256) return fail("ecg_observe_onehot: channels must be in 1..256");
    if (elem_kind < 0 || elem_kind > 3) return fail("ecg_observe_onehot: elem_kind 0=u8 1=f32 2=bf16 3=f16");
    if (n <= 0) return 0;
    ops_for(cfg->rows)->onehot(boards, out, channels, elem_kind, cfg->types, n, (cudaStream_t)stream);
    return check_launch("ecg_observe_onehot");
}

int ecg_augment(const ecg_config *cfg, const void *boards_in, void *boards_out, int mirror, const uint8_t *type_perm,
                int64_t n, void *stream) {
    if (check_cfg(cfg)) return -1;
    if (!boards_in || !boards_out) return fail("ecg_augment: NULL buffer");
    CodeLut lut;
    for (int c = 0; c < 16; c++) lut.v[c] = (uint8_t)c;
    if (type_perm) { // type t (1..types) becomes type_perm[t - 1]; must be a permutation of 1..types
        const int top = cfg->type_mask < 11 ? cfg->type_mask : 11; // plain codes the engine can hold
        unsigned seen = 0;
        for (int t = 1; t <= cfg->types; t++) {
            const int v = type_perm[t - 1];
            if (v < 1 || v > cfg->types || v > top || (seen >> v) & 1u)
                return fail("ecg_augment: type_perm must be a permutation of 1..types");
            seen |= 1u << v;
            lut.v[t] = (uint8_t)v;
        }
    }
    if (n <= 0) return 0;
    ops_for(cfg->rows)->augment(boards_in, boards_out, mirror != 0, type_perm != nullptr, lut, n, (cudaStream_t)stream);
    return check_launch("ecg_augment");
}

int ecg_episode_stats(const int32_t *score, const uint8_t *flags, int64_t *out, int64_t n, void *stream) {
    if (!score || !out) return fail("ecg_episode_stats: NULL buffer");
    if (n <= 0) return 0;
    unsigned g = grid_for(n, 256);
    if (g > 148 * 8) g = 148 * 8;
    stats_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(score, flags, (long long *)out, n);
    return check_launch("ecg_episode_stats");
}
}

// ------------------------------------------------------------------ host side of the compact observation

namespace {

struct ExpandGeo {
    int rc, nb;        // cells and nibble bytes per board
    uint8_t lut[16];   // code -> cell value (boardConfig.py:29-43)
    uint16_t tab[256]; // nibble byte -> two cell values (low nibble first)
};

ExpandGeo expand_geo(const ecg_config *cfg) {
    ExpandGeo g;
    g.rc = cfg->rows * cfg->cols;
    g.nb = (g.rc + 1) / 2;
    for (int c = 0; c < 12; c++) g.lut[c] = (uint8_t)c;
    g.lut[12] = (uint8_t)cfg->h_line;
    g.lut[13] = (uint8_t)cfg->v_line;
    g.lut[14] = (uint8_t)cfg->bomb;
    g.lut[15] = (uint8_t)cfg->mega_token;
    for (int b = 0; b < 256; b++) g.tab[b] = (uint16_t)(g.lut[b & 15] | (g.lut[b >> 4] << 8));
    return g;
}

void expand_scalar(const ExpandGeo &g, const uint8_t *src, uint8_t *dst, int from_byte) {
    for (int k = from_byte; k < g.nb; k++) {
        const uint16_t v = g.tab[src[k]];
        dst[2 * k] = (uint8_t)v;
        if (2 * k + 1 < g.rc) dst[2 * k + 1] = (uint8_t)(v >> 8);
    }
}

// 16 nibble bytes -> 32 cell bytes: two table lookups (pshufb) and an interleave
__attribute__((target("ssse3"))) inline void expand16(const __m128i lut, const uint8_t *src, uint8_t *dst) {
    const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
    const __m128i m = _mm_set1_epi8(0x0F);
    const __m128i lo = _mm_shuffle_epi8(lut, _mm_and_si128(x, m));
    const __m128i hi = _mm_shuffle_epi8(lut, _mm_and_si128(_mm_srli_epi16(x, 4), m));
    _mm_storeu_si128(reinterpret_cast<__m128i *>(dst), _mm_unpacklo_epi8(lo, hi));
    _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + 16), _mm_unpackhi_epi8(lo, hi));
}

// Boards of 16+ nibble bytes: whole 16-byte blocks, then one block that ends with the board's last byte (it overlaps
// the previous one).  With an odd cell count that block writes one byte past the board -- the first byte of the next
// board, which is written afterwards -- so the last board of a range takes the scalar tail instead.
__attribute__((target("ssse3"))) void expand_range_ssse3(const ExpandGeo &g, const uint8_t *nib, uint8_t *cells,
                                                         int64_t n) {
    const __m128i lut = _mm_loadu_si128(reinterpret_cast<const __m128i *>(g.lut));
    const int full = g.nb / 16 * 16, tail = g.nb - 16;
    const bool spill = (g.rc & 1) != 0;
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *src = nib + i * g.nb;
        uint8_t *dst = cells + i * g.rc;
        for (int off = 0; off < full; off += 16) expand16(lut, src + off, dst + 2 * off);
        if (full < g.nb) {
            if (spill && i == n - 1) expand_scalar(g, src, dst, full);
            else expand16(lut, src + tail, dst + 2 * tail);
        }
    }
}

// The same through a stack buffer of 16 boards (16 * rc bytes: a whole number of 16-byte blocks) that leaves with
// non-temporal stores: the output is written once and not read again by these threads, and a cached store would first
// read every line it is about to overwrite (1.36 GB of extra DRAM reads per 2^24 9x9 boards).  cells must be 16-byte
// aligned; n a multiple of 16.
__attribute__((target("ssse3"))) void expand_range_ssse3_nt(const ExpandGeo &g, const uint8_t *nib, uint8_t *cells,
                                                            int64_t n) {
    const __m128i lut = _mm_loadu_si128(reinterpret_cast<const __m128i *>(g.lut));
    const int full = g.nb / 16 * 16, tail = g.nb - 16;
    alignas(64) uint8_t buf[16 * 256 + 32];
    for (int64_t i0 = 0; i0 < n; i0 += 16) {
        for (int k = 0; k < 16; k++) {
            const uint8_t *src = nib + (i0 + k) * g.nb;
            uint8_t *dst = buf + k * g.rc;
            for (int off = 0; off < full; off += 16) expand16(lut, src + off, dst + 2 * off);
            if (full < g.nb) expand16(lut, src + tail, dst + 2 * tail);
        }
        __m128i *out = reinterpret_cast<__m128i *>(cells + i0 * g.rc);
        const __m128i *in = reinterpret_cast<const __m128i *>(buf);
        for (int q = 0; q < g.rc; q++) _mm_stream_si128(out + q, _mm_load_si128(in + q));
    }
    _mm_sfence();
}

void expand_range(const ExpandGeo &g, const uint8_t *nib, uint8_t *cells, int64_t n) {
    static const bool ssse3 = __builtin_cpu_supports("ssse3");
    static const bool nt = !(getenv("ECG_EXPAND_NT") && getenv("ECG_EXPAND_NT")[0] == '0');
    if (ssse3 && g.nb >= 16) {
        const int64_t n16 = (nt && (reinterpret_cast<uintptr_t>(cells) & 15u) == 0) ? (n & ~(int64_t)15) : 0;
        if (n16) expand_range_ssse3_nt(g, nib, cells, n16);
        if (n16 < n) expand_range_ssse3(g, nib + n16 * g.nb, cells + n16 * g.rc, n - n16);
        return;
    }
    for (int64_t i = 0; i < n; i++) expand_scalar(g, nib + i * g.nb, cells + i * g.rc, 0);
}

struct Expander {
    struct Job {
        ExpandGeo g;
        const uint8_t *nib;
        uint8_t *cells;
        int64_t n;
        cudaEvent_t event;
    };
    std::vector<std::thread> workers;
    std::deque<Job> queue;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    int64_t pending = 0;
    int device = 0;
    bool stop = false, failed = false;
    char err[200] = "";

    explicit Expander(int threads) {
        cudaGetDevice(&device);
        for (int t = 0; t < threads; t++) workers.emplace_back([this] { run(); });
    }
    ~Expander() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &w : workers) w.join();
    }
    void run() {
        cudaSetDevice(device); // the events belong to the creator's device
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [this] { return stop || !queue.empty(); });
                if (queue.empty()) return;
                j = queue.front();
                queue.pop_front();
            }
            cudaError_t e = j.event ? cudaEventSynchronize(j.event) : cudaSuccess;
            if (e == cudaSuccess) expand_range(j.g, j.nib, j.cells, j.n);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (e != cudaSuccess && !failed) {
                    failed = true;
                    snprintf(err, sizeof(err), "ecg_host_expander: %s", cudaGetErrorString(e));
                }
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
};

} // namespace

extern "C" {

int ecg_host_expand_nibbles(const ecg_config *cfg, const uint8_t *nibbles, uint8_t *cells, int64_t n) {
    if (check_cfg(cfg)) return -1;
    if (!nibbles || !cells) return fail("ecg_host_expand_nibbles: NULL buffer");
    if (n <= 0) return 0;
    expand_range(expand_geo(cfg), nibbles, cells, n);
    return 0;
}

void *ecg_host_expander_create(int threads) {
    if (threads < 1 || threads > 1024) {
        fail("ecg_host_expander_create: threads must be in 1..1024");
        return nullptr;
    }
    return new Expander(threads);
}

int ecg_host_expander_submit(void *expander, const ecg_config *cfg, const uint8_t *nibbles, uint8_t *cells, int64_t n,
                             void *event, int split) {
    if (!expander) return fail("ecg_host_expander_submit: NULL expander");
    if (check_cfg(cfg)) return -1;
    if (!nibbles || !cells) return fail("ecg_host_expander_submit: NULL buffer");
    if (n <= 0) return 0;
    Expander *x = static_cast<Expander *>(expander);
    if (split < 1) split = 1;
    if (split > n) split = (int)n;
    const ExpandGeo g = expand_geo(cfg);
    {
        std::lock_guard<std::mutex> lk(x->mu);
        for (int s = 0; s < split; s++) { // pieces start at multiples of 16 boards (16-byte aligned output)
            const int64_t lo = (n * s / split) & ~(int64_t)15, hi = s + 1 == split ? n : (n * (s + 1) / split) & ~(int64_t)15;
            if (hi <= lo) continue;
            x->queue.push_back({g, nibbles + lo * g.nb, cells + lo * g.rc, hi - lo, (cudaEvent_t)event});
            x->pending++;
        }
    }
    x->cv_work.notify_all();
    return 0;
}

int ecg_host_expander_wait(void *expander) {
    if (!expander) return fail("ecg_host_expander_wait: NULL expander");
    Expander *x = static_cast<Expander *>(expander);
    std::unique_lock<std::mutex> lk(x->mu);
    x->cv_done.wait(lk, [x] { return x->pending == 0; });
    if (x->failed) {
        x->failed = false;
        return fail(x->err);
    }
    return 0;
}

void ecg_host_expander_destroy(void *expander) { delete static_cast<Expander *>(expander); }

} // extern "C"
