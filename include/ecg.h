/*
 * ecg.h -- C-ABI of libecg.so: the B200 (sm_100a) batched stepping engine for
 * Element-Crush-Gym's match-3 environment.
 *
 * The reference has no FFI: its seam is the Python `State` ABC
 * (mctslib/abc/mcts.py:8-30) implemented by `BoardV2` (match3tile/boardv2.py:11-226),
 * the free functions of match3tile/boardFunctions.py and `Match3Env`
 * (match3tile/env.py:8-66).  Each entry point below names the reference code it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (e.g. torch CUDA tensors);
 *    the library never allocates, never synchronises, and launches on `stream`
 *    (a cudaStream_t passed as void*; NULL = the legacy default stream);
 *  - every function returns 0 on success, <0 on a bad argument or CUDA error
 *    (message via ecg_last_error()); data-dependent conditions are reported per
 *    board in `status` (ECG_ST_* bits), never as errors;
 *  - boards live in HBM in the engine's packed format (4-bit cells, bit-sliced,
 *    32-board interleaved tiles); ecg_pack/ecg_unpack convert from
 *    and to the reference's row-major cell arrays (BoardV2.array, int64 or uint8).
 */
#ifndef ECG_H
#define ECG_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECG_VERSION 106

/* per-board status bits */
#define ECG_ST_TERMINAL 1        /* n_actions < 1: board returned unchanged (boardv2.py:44-45) */
#define ECG_ST_STREAM_OVERFLOW 2 /* replay stream exhausted; results of this board are invalid */
#define ECG_ST_SHUFFLE_CAP 4     /* boardv2.py:188-194 shuffle loop stopped after 64 rounds (reference: may never end) */
#define ECG_ST_BAD_ACTION 8      /* action outside [0, action_space): no-op (reference: KeyError, boardv2.py:48) */
#define ECG_ST_NO_LEGAL 16       /* random pick on an empty legal set: no-op (reference: ValueError in np.random.choice) */
#define ECG_ST_BAD_CELL 32       /* ecg_pack: value outside {0, 1..min(type_mask,11), h_line, v_line, bomb, mega_token} */
#define ECG_ST_CASCADE_CAP 64    /* cascade loop (boardv2.py:138) stopped after 1024 iterations */

/* env flags written by ecg_step (env.py:54-55) */
#define ECG_FLAG_DONE 1
#define ECG_FLAG_WON 2

#define ECG_REFILL_REPLAY 1 /* raw u32 output of numpy's legacy MT19937, restarted every step (boardv2.py:46) */
#define ECG_REFILL_PHILOX 2 /* counter-based Philox4x32-10, keyed by (key, global board index, step) */

#define ECG_TILE 32 /* boards per interleave tile */

/* BoardConfig (match3tile/boardConfig.py:5-43) plus the packed-layout sizes. */
typedef struct ecg_config {
    int32_t rows, cols, types;
    int32_t bits, type_mask, special_type_mask, h_line, v_line, bomb, mega_token; /* boardConfig.py:29-43 */
    int32_t action_space; /* rows * (cols - 1) * 2, boardConfig.py:27 */
    int32_t board_words;  /* u32 words per packed board = 4 planes x ceil(rows*(cols+1)/32) */
    int32_t mask_words;   /* u32 words per packed legal mask = 2 swap bitboards (horizontal, vertical) = board_words/2 */
    int32_t reserved;
} ecg_config;

/* Where refill tiles (boardv2.py:172), shuffle permutations (boardFunctions.py:16-23) and random action
 * picks (np.random.choice, samplerTasks.py:13 / mctslib/standard/mcts.py:17) come from. */
typedef struct ecg_refill {
    int32_t mode;            /* ECG_REFILL_* */
    int32_t stream_len;      /* replay: u32 words available per board */
    const uint32_t *stream;  /* replay: raw MT19937 words (ecg_mt19937_stream) */
    int64_t stream_stride;   /* replay: words between consecutive boards' streams; 0 = one shared stream */
    uint32_t *stream_pos;    /* replay, optional in/out [n]: words consumed since the last reseed */
    uint64_t philox_key;     /* philox: run key */
    uint64_t board0;         /* philox: global index of board 0 of this call (shard offset) */
    uint32_t step_ctr;       /* philox: step counter of this call (a rollout uses step_ctr + t) */
    uint32_t reserved;
    const int32_t *stream_index; /* replay, optional [n]: board i replays the stream at stream + stream_index[i] *
                                    stream_stride instead of i * stream_stride (children of an expansion keep their
                                    parent's stream, boardv2.py:46: every BoardV2 reseeds with the same cfg.seed) */
    const uint32_t *tiles;       /* replay, optional: per-stream tile tables made by ecg_replay_tiles from `stream` for this */
    const uint16_t *tile_wpos;   /* board config's `types`.  With them (and ecg_step_io.scratch) the replay step runs as two
                                    kernels like the Philox step; results are identical with and without.  Streams must be
                                    packed (stream_stride == stream_len, or 0 for one shared stream). */
} ecg_refill;

/* Buffers of one lockstep step.  n-element arrays are plain board-indexed; boards/masks are packed. */
typedef struct ecg_step_io {
    const void *boards_in;   /* packed boards */
    void *boards_out;        /* packed boards; may alias boards_in */
    const int32_t *actions;  /* [n] or NULL: NULL = pick uniformly from mask_in (board.random_action()) */
    const uint32_t *mask_in; /* packed legal mask of boards_in; required when actions == NULL */
    int32_t *actions_out;    /* optional [n]: the action applied (-1 = none) */
    int32_t *moves_left;     /* optional in/out [n]: BoardV2.n_actions; < 1 = terminal no-op; decremented */
    int32_t *reward;         /* optional out [n]: points of this step (BoardV2 reward delta = env move_score) */
    int32_t *score;          /* optional in/out [n]: cumulative reward (BoardV2._reward, env.score) */
    int32_t *cascades;       /* optional out [n]: cascade-loop iterations of this step ("combo count") */
    uint32_t *mask_out;      /* optional out: packed legal mask of boards_out (legal_actions of the new state) */
    uint8_t *flags;          /* optional out [n]: ECG_FLAG_DONE | ECG_FLAG_WON (env.py:54-55); needs score+moves_left */
    uint8_t *status;         /* optional out [n]: ECG_ST_* */
    int32_t env_goal;        /* env.py:17 env_goal */
    int32_t reserved;
    const int32_t *src_index; /* optional [n]: job i steps board src_index[i] of boards_in (and uses that board's
                                 refill stream / Philox id, moves_left, score, mask_in) and writes every output at i:
                                 "expand these (board, action) pairs" = Node.expand / greedy_action (boardv2.py:209-218,
                                 mctslib/standard/mcts.py:31-42).  Needs boards_out != boards_in, and moves_left ==
                                 score == NULL (in/out arrays would be read at src_index[i] while job src_index[i]
                                 writes them); replay mode with actions == NULL also needs stream_pos == NULL.  With
                                 explicit actions stream_pos is output only: [i] = words this step consumed. */
    int32_t *scratch;        /* optional work list, [n + 1] int32, contents irrelevant on entry and exit.  Philox mode, or
                                replay mode with tile tables (ecg_refill.tiles): when given, the step runs as TWO kernels -- the common-case kernel over all n
                                boards, which hands the boards that need a rare path (intersecting runs, runs of 6+,
                                a swap of two special tokens, the shuffle loop; about 5 % of the 9x9x6 steps) over to
                                the exact kernel through this list.  Results are identical with and without it. */
} ecg_step_io;

int ecg_version(void);
/* sizeof the library's own structs, for binders to check their stubs against (a short ecg_step_io would make
 * ecg_step read past the caller's struct): which = ECG_SIZEOF_* */
#define ECG_SIZEOF_CONFIG 0
#define ECG_SIZEOF_REFILL 1
#define ECG_SIZEOF_STEP_IO 2
int ecg_sizeof(int which);
const char *ecg_last_error(void);
/* kernels launched by this process so far (bench.py's gpu_launches) */
int64_t ecg_launch_count(void);

/* BoardConfig.__post_init__ (boardConfig.py:26-43).  rows == cols in 4..16, types in 1..11;
 * every size 4..16 is built into the library (one kernel object per size). */
int ecg_config_init(ecg_config *cfg, int rows, int cols, int types);
/* bytes of a packed board / mask buffer for n boards (n is rounded up to a whole tile) */
int64_t ecg_boards_bytes(const ecg_config *cfg, int64_t n);
int64_t ecg_masks_bytes(const ecg_config *cfg, int64_t n);

/* BoardV2.array -> packed.  cells: [n, rows, cols] row-major, elem_bytes 8 (int64) or 1 (uint8).
 * status (optional, [n]) gets ECG_ST_BAD_CELL for boards with an unrepresentable value. */
int ecg_pack(const ecg_config *cfg, const void *cells, int elem_bytes, void *boards, uint8_t *status, int64_t n,
             void *stream);
/* packed -> BoardV2.array (the env's observation, env.py:56) */
int ecg_unpack(const ecg_config *cfg, const void *boards, void *cells, int elem_bytes, int64_t n, void *stream);
/* packed -> the compact observation: 4-bit cell CODES (0 empty, 1..11 plain token of that type, 12 h_line, 13 v_line,
 * 14 bomb, 15 mega_token), row-major, two cells per byte (cell 2k in the low nibble of byte k), ceil(rows*cols/2)
 * bytes per board, boards back to back: BoardV2.array at half the bytes of the uint8 form (env.py:56 observation for
 * callers that cross PCIe).  out must be 4-byte aligned. */
int ecg_unpack_nibbles(const ecg_config *cfg, const void *boards, uint8_t *out, int64_t n, void *stream);
/* packed legal mask -> [n, action_space] bytes (1 = legal): membership form of BoardV2.legal_actions */
int ecg_unpack_mask(const ecg_config *cfg, const uint32_t *mask, uint8_t *out, int64_t n, void *stream);

/* raw u32 output of np.random.seed(seed) (numpy legacy MT19937 init_genrand): out[i*len + k] */
int ecg_mt19937_stream(const uint32_t *seeds, uint32_t *out, int32_t len, int64_t n, void *stream);

/* Tile tables of the replay streams: within one apply_action the refill tiles np.random.randint(1, types + 1)
 * (boardv2.py:172) are a prefix of the stream's accepted values (masked rejection restarted by np.random.seed,
 * boardv2.py:46).  tiles: ecg_replay_tiles_words(stream_len) u32 per stream (tile j = nibble j), tile_wpos:
 * stream_len + 1 u16 per stream (raw words consumed once j tiles are taken).  streams: [n_streams, stream_len]. */
int64_t ecg_replay_tiles_words(int32_t stream_len);
int ecg_replay_tiles(const uint32_t *streams, int32_t stream_len, int types, uint32_t *tiles, uint16_t *tile_wpos,
                     int64_t n_streams, void *stream);

/* BoardV2.__init__ (boardv2.py:20-27): draw boards, redraw matched cells until no match remains.
 * replay: consumes the board's stream from position 0; philox: substream (board, step 0xFFFFFFFF). */
int ecg_init_boards(const ecg_config *cfg, const ecg_refill *rf, void *boards, uint8_t *status, int64_t n,
                    void *stream);

/* boardFunctions.legal_actions (:26-112) for every board, as a packed mask */
int ecg_legal_mask(const ecg_config *cfg, const void *boards, uint32_t *mask, int64_t n, void *stream);

/* np.random.choice(state.legal_actions) (samplerTasks.py:13): the idx-th legal action, ascending.
 * replay: idx-th legal action in ascending action order, idx by numpy's masked rejection at stream_pos;
 * philox: idx = mulhi(word 0 of the (board, step) substream, n), counted in swap-bitboard order (all horizontal
 * swaps by (row, col), then all vertical swaps) */
int ecg_random_action(const ecg_config *cfg, const ecg_refill *rf, const uint32_t *mask, int32_t *actions,
                      uint8_t *status, int64_t n, void *stream);

/* BoardV2.apply_action (boardv2.py:43-207) + Match3Env.step bookkeeping (env.py:48-56) for n boards */
int ecg_step(const ecg_config *cfg, const ecg_refill *rf, const ecg_step_io *io, int64_t n, void *stream);
/* Measurement hook: the next two-kernel ecg_step of this thread records `event` (a cudaEvent_t) on its stream
 * between the common-case kernel and the exact kernel, so a caller can time the two separately (bench.py). */
int ecg_step_mark_event(void *event);

/* MCTS.rollout / random_task (mctslib/standard/mcts.py:14-19, samplerTasks.py:9-14): play random legal
 * actions until moves_left reaches 0.  boards are updated in place; total_reward[n] receives the points
 * collected; steps_done (optional, [n]) the number of actions applied. */
int ecg_rollout(const ecg_config *cfg, const ecg_refill *rf, void *boards, const int32_t *moves_left,
                int64_t *total_reward, int32_t *steps_done, uint8_t *status, int64_t n, void *stream);
/* The same with a work list (scratch: [n + 1] int32, contents irrelevant on entry and exit; needs steps_done).  Philox
 * mode: the episodes are played by the common-case kernel; an episode that meets a rare case (see ecg_step_io.scratch)
 * is parked -- the board of that step's start, the reward and step count so far -- and finished by the exact kernel.
 * Results are identical with and without the list. */
int ecg_rollout_scratch(const ecg_config *cfg, const ecg_refill *rf, void *boards, const int32_t *moves_left,
                        int64_t *total_reward, int32_t *steps_done, uint8_t *status, int32_t *scratch, int64_t n,
                        void *stream);

/* Observation for the policy/value net: nnx.one_hot(board.array, channels) (elementCrush.py:66,92), layout
 * [n, rows, cols, channels]; a cell value >= channels (e.g. the mega token, 32 of 32 channels) is all zeros, as in
 * jax.nn.one_hot.  elem_kind: 0 = uint8, 1 = float32, 2 = bfloat16, 3 = float16. */
int ecg_observe_onehot(const ecg_config *cfg, const void *boards, void *out, int channels, int elem_kind, int64_t n,
                       void *stream);

/* Dataset augmentation (dataset.py:86-112 mirror, :114-176 type_switch) on packed boards: boards_out[i] =
 * fliplr(boards_in[i]) when mirror != 0, and every plain token t (1..types) renamed to type_perm[t-1] when
 * type_perm (HOST pointer, a permutation of 1..types) is not NULL; empty cells and special tokens keep their
 * value.  boards_out may alias boards_in.  The matching policy re-indexing (action -> mirrored action,
 * dataset.py:99-107) is a fixed permutation of the action ids and lives in the host layer. */
int ecg_augment(const ecg_config *cfg, const void *boards_in, void *boards_out, int mirror, const uint8_t *type_perm,
                int64_t n, void *stream);

/* Host side of the compact observation (no CUDA kernel): 4-bit cell codes in HOST memory (the ecg_unpack_nibbles
 * layout, e.g. the pinned buffer a D2H copy filled) -> BoardV2.array as uint8 cell VALUES [n, rows, cols] (env.py:56),
 * byte-identical to ecg_unpack(..., elem_bytes = 1) of the same boards.  PCIe is the whole end-to-end cost of a step
 * (the observation is 81 of 91 bytes per 9x9 board): crossing it with 41 bytes and widening on the host cores halves
 * the transfer.  Single-threaded; the expander below runs it on a pool of host threads. */
int ecg_host_expand_nibbles(const ecg_config *cfg, const uint8_t *nibbles, uint8_t *cells, int64_t n);
/* A pool of `threads` host threads (bound to the calling thread's CUDA device).  submit() queues the expansion of n
 * boards, cut into `split` >= 1 pieces; a piece starts once `event` (a cudaEvent_t recorded after the D2H copy of the
 * nibbles, or NULL) has completed, so chunks are widened while later chunks are still crossing PCIe.  wait() blocks
 * until everything submitted is done and returns 0, or -1 with ecg_last_error() if a piece failed. */
void *ecg_host_expander_create(int threads);
int ecg_host_expander_submit(void *expander, const ecg_config *cfg, const uint8_t *nibbles, uint8_t *cells, int64_t n,
                             void *event, int split);
int ecg_host_expander_wait(void *expander);
void ecg_host_expander_destroy(void *expander);

/* episode statistics (main.py:240-267 sample()): out[0]=sum(score) out[1]=n out[2]=min out[3]=max
 * out[4]=#flags&WON out[5]=sum(score^2); out must be zero-initialised except out[2]=INT64_MAX, out[3]=INT64_MIN */
int ecg_episode_stats(const int32_t *score, const uint8_t *flags, int64_t *out, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif
