// ecg_ops.h -- internal seam between the per-size kernel objects (ecg_shape_kernels.cu, one
// object per -DECG_SIZE) and the C-ABI (ecg_api.cu).  Not installed; not part of include/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ecg {

struct RefillDev {
    const uint32_t *stream;
    long long stream_stride;
    uint32_t *stream_pos;
    unsigned long long key, board0;
    uint32_t step_ctr;
    int stream_len;
    const int32_t *stream_index; // replay, optional: board i replays stream stream_index[i]
    const uint32_t *tiles;       // replay, optional: per-stream tile tables (ecg_replay_tiles) -> two-kernel replay step
    const uint16_t *tile_wpos;
};

struct StepDev {
    const void *boards_in;
    void *boards_out;
    const int32_t *actions;
    const uint32_t *mask_in;
    int32_t *actions_out, *moves_left, *reward, *score, *cascades;
    uint32_t *mask_out;
    uint8_t *flags, *status;
    int env_goal, types;
    const int32_t *src_index;
    // two-kernel step (ecg_step_io.scratch): the common-case kernel appends the jobs it hands off to
    // handoff[1 + handoff[0]++]; the exact kernel then runs over jobs[0 .. *n_jobs)
    int32_t *handoff;
    const int32_t *jobs, *n_jobs;
    void *mid_event; // host side only: cudaEvent_t recorded between the two kernels (ecg_step_mark_event)
};

struct CodeLut {
    uint8_t v[16];
};

// `wide` selects the types >= 8 instantiation (4 token planes + np.clip(.., 0, 32) quirk)
struct ShapeOps {
    void (*pack)(bool wide, const void *cells, int elem_bytes, void *boards, uint8_t *status, int types, long long n,
                 cudaStream_t s);
    void (*unpack)(bool wide, const void *boards, void *cells, int elem_bytes, int types, long long n, cudaStream_t s);
    void (*unpack_nibbles)(const void *boards, uint8_t *out, long long n, cudaStream_t s);
    void (*unpack_mask)(const uint32_t *mask, uint8_t *out, long long n, cudaStream_t s);
    void (*init)(bool wide, bool philox, RefillDev rf, void *boards, uint8_t *status, int types, long long n,
                 cudaStream_t s);
    void (*legal)(bool wide, const void *boards, uint32_t *mask, long long n, cudaStream_t s);
    void (*random_action)(bool philox, RefillDev rf, const uint32_t *mask, int32_t *actions, uint8_t *status,
                          long long n, cudaStream_t s);
    int (*step)(bool wide, bool philox, RefillDev rf, StepDev io, long long n, cudaStream_t s); // kernels launched
    int (*rollout)(bool wide, bool philox, RefillDev rf, void *boards, const int32_t *moves_left,
                   long long *total_reward, int32_t *steps_done, uint8_t *status, int32_t *scratch, int types,
                   long long n, cudaStream_t s); // kernels launched
    void (*onehot)(const void *boards, void *out, int channels, int elem_kind, int types, long long n, cudaStream_t s);
    void (*augment)(const void *boards_in, void *boards_out, bool mirror, bool remap, CodeLut lut, long long n,
                    cudaStream_t s);
};

const ShapeOps *shape_ops_4();
const ShapeOps *shape_ops_5();
const ShapeOps *shape_ops_6();
const ShapeOps *shape_ops_7();
const ShapeOps *shape_ops_8();
const ShapeOps *shape_ops_9();
const ShapeOps *shape_ops_10();
const ShapeOps *shape_ops_11();
const ShapeOps *shape_ops_12();
const ShapeOps *shape_ops_13();
const ShapeOps *shape_ops_14();
const ShapeOps *shape_ops_15();
const ShapeOps *shape_ops_16();

} // namespace ecg
