// ecg_bits.cuh -- multi-word bitboards for the match-3 stepping engine.
//
// A board of R x C cells is laid out row-major with one zero pad column:
// cell (r, c) <-> bit r*S + c, S = C + 1.  "right" = +1, "down" = +S.  The pad
// column (and every bit >= R*S) is zero in all board planes, which is what makes
// single-bit horizontal shifts wrap-free.  9x9 -> 90 bits -> 3 x u32.
//
// Everything here is __host__ __device__ so the identical logic is compiled by
// g++ into the test-only host simulator (tests/hostsim) and by nvcc into the
// sm_100a kernels.  Shift amounts are compile-time constants: one funnel shift
// (SHF) per word.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ECG_HD __host__ __device__ __forceinline__
#define ECG_HD_NOINLINE __host__ __device__ __noinline__
#define ECG_HD_CONSTEXPR __host__ __device__ constexpr
#else
#define ECG_HD_CONSTEXPR constexpr
#define ECG_HD inline __attribute__((always_inline))
#define ECG_HD_NOINLINE __attribute__((noinline))
#endif

// -DECG_PROFILE_PHASES keeps the major phases out of line so ncu attributes instructions per phase
#if defined(ECG_PROFILE_PHASES)
#define ECG_PHASE ECG_HD_NOINLINE
#else
#define ECG_PHASE ECG_HD
#endif

namespace ecg {

ECG_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// index of lowest set bit; x != 0
ECG_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// 1 + index of the highest set bit; 0 for x == 0
ECG_HD int bitlen32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 32 - __clz((int)x);
#else
    return x ? 32 - __builtin_clz(x) : 0;
#endif
}
ECG_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// low 32 bits of (hi:lo) >> s, 0 <= s < 32
ECG_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
// high 32 bits of (hi:lo) << s, 0 <= s < 32
ECG_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, int s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

template <int W>
struct BB {
    uint32_t w[W];
};

template <int W>
ECG_HD BB<W> bb_zero() {
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = 0u;
    return r;
}
template <int W>
ECG_HD BB<W> operator&(const BB<W> &a, const BB<W> &b) {
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = a.w[i] & b.w[i];
    return r;
}
template <int W>
ECG_HD BB<W> operator|(const BB<W> &a, const BB<W> &b) {
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = a.w[i] | b.w[i];
    return r;
}
template <int W>
ECG_HD BB<W> operator^(const BB<W> &a, const BB<W> &b) {
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = a.w[i] ^ b.w[i];
    return r;
}
// a & ~b
template <int W>
ECG_HD BB<W> andn(const BB<W> &a, const BB<W> &b) {
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = a.w[i] & ~b.w[i];
    return r;
}
template <int W>
ECG_HD BB<W> &operator|=(BB<W> &a, const BB<W> &b) {
#pragma unroll
    for (int i = 0; i < W; i++) a.w[i] |= b.w[i];
    return a;
}
template <int W>
ECG_HD BB<W> &operator&=(BB<W> &a, const BB<W> &b) {
#pragma unroll
    for (int i = 0; i < W; i++) a.w[i] &= b.w[i];
    return a;
}
template <int W>
ECG_HD bool any(const BB<W> &a) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < W; i++) x |= a.w[i];
    return x != 0;
}
template <int W>
ECG_HD int popcount(const BB<W> &a) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < W; i++) n += popc32(a.w[i]);
    return n;
}

// bit b of result = a(b + K): moves cells toward lower indices (left / up)
template <int K, int W>
ECG_HD BB<W> shr(const BB<W> &a) {
    constexpr int q = K / 32, s = K % 32;
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) {
        uint32_t lo = (i + q < W) ? a.w[(i + q < W) ? i + q : 0] : 0u;
        uint32_t hi = (i + q + 1 < W) ? a.w[(i + q + 1 < W) ? i + q + 1 : 0] : 0u;
        r.w[i] = (s == 0) ? lo : funnel_r(lo, hi, s);
    }
    return r;
}
// bit b of result = a(b - K): moves cells toward higher indices (right / down)
template <int K, int W>
ECG_HD BB<W> shl(const BB<W> &a) {
    constexpr int q = K / 32, s = K % 32;
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) {
        uint32_t hi = (i - q >= 0) ? a.w[(i - q >= 0) ? i - q : 0] : 0u;
        uint32_t lo = (i - q - 1 >= 0) ? a.w[(i - q - 1 >= 0) ? i - q - 1 : 0] : 0u;
        r.w[i] = (s == 0) ? hi : funnel_l(lo, hi, s);
    }
    return r;
}

// ---- run-time indexed helpers (rare paths: swap, special tokens, refill) ----

template <int W>
ECG_HD bool testbit(const BB<W> &a, int b) {
    uint32_t x = 0;
    const int wi = b >> 5;
#pragma unroll
    for (int i = 0; i < W; i++) x = (i == wi) ? a.w[i] : x;
    return (x >> (b & 31)) & 1u;
}
template <int W>
ECG_HD BB<W> onehot(int b) {
    BB<W> r;
    const int wi = b >> 5;
    const uint32_t m = 1u << (b & 31);
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = (i == wi) ? m : 0u;
    return r;
}
// bits [lo, hi) set; 0 <= lo <= hi <= 32*W
template <int W>
ECG_HD BB<W> bitrange(int lo, int hi) {
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) {
        int l = lo - 32 * i, h = hi - 32 * i;
        l = l < 0 ? 0 : l;
        h = h > 32 ? 32 : h;
        uint32_t m = 0u;
        if (h > l) m = ((h >= 32) ? 0xFFFFFFFFu : ((1u << h) - 1u)) & ~((l >= 32) ? 0xFFFFFFFFu : ((1u << l) - 1u));
        r.w[i] = m;
    }
    return r;
}
// 32 bits of a starting at bit b (bits beyond the array read as zero)
template <int W>
ECG_HD uint32_t extract32(const BB<W> &a, int b) {
    const int wi = b >> 5, s = b & 31;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < W; i++) {
        lo = (i == wi) ? a.w[i] : lo;
        hi = (i == wi + 1) ? a.w[i] : hi;
    }
    return funnel_r(lo, hi, s);
}
// bit i of the result = a(i*S + c), i < R: column c of a board-shaped bitboard as a dense word
template <int R, int S, int W>
ECG_HD uint32_t column_bits(const BB<W> &a, int c) {
    uint32_t sh[W]; // a >> c (c < 32)
#pragma unroll
    for (int i = 0; i < W; i++) sh[i] = funnel_r(a.w[i], i + 1 < W ? a.w[i + 1 < W ? i + 1 : 0] : 0u, c);
    uint32_t col = 0;
#pragma unroll
    for (int i = 0; i < R; i++) {
        constexpr int dummy = 0;
        (void)dummy;
        const int bp = (i * S) & 31, wi = (i * S) >> 5; // compile-time after unrolling
        col |= (bp >= i ? sh[wi] >> (bp - i) : sh[wi] << (i - bp)) & (1u << i);
    }
    return col;
}
// a << j for a run-time 0 <= j < 32
template <int W>
ECG_HD BB<W> shl_rt(const BB<W> &a, int j) {
    BB<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = funnel_l(i > 0 ? a.w[i > 0 ? i - 1 : 0] : 0u, a.w[i], j);
    return r;
}
// index of the lowest set bit of a (a != 0) and clear it
template <int W>
ECG_HD int pop_lowest(BB<W> &a) {
    int b = 0;
    bool done = false;
#pragma unroll
    for (int i = 0; i < W; i++) {
        if (!done && a.w[i]) {
            b = 32 * i + ctz32(a.w[i]);
            a.w[i] &= a.w[i] - 1u;
            done = true;
        }
    }
    return b;
}

// ---- compile-time geometry ----

template <int R_, int C_>
struct Geo {
    static constexpr int R = R_, C = C_, S = C_ + 1, NB = R_ * (C_ + 1), W = (R_ * (C_ + 1) + 31) / 32;
    static constexpr int A = R_ * (C_ - 1) + C_ * (R_ - 1); // == rows*(cols-1)*2 (boardConfig.py:27) for square boards
    static constexpr int AW = (A + 31) / 32;
    static constexpr int ROWA = 2 * C_ - 1; // actions per board row (boardConfig.py:46)
    static_assert(R_ == C_, "the reference's action space (boardConfig.py:27) is only consistent for square boards");
    static_assert(C_ >= 4 && C_ <= 16, "boardConfig.decode (:50) needs columns >= 4; 16 is the engine's limit");

    // word i of the set { bit r*S + c : r0 <= r < r1, c0 <= c < c1 }
    static ECG_HD_CONSTEXPR uint32_t rect_word(int i, int r0, int r1, int c0, int c1) {
        uint32_t m = 0;
        for (int r = r0; r < r1; r++)
            for (int c = c0; c < c1; c++) {
                int b = r * S + c;
                if ((b >> 5) == i) m |= 1u << (b & 31);
            }
        return m;
    }
    template <int R0, int R1, int C0, int C1>
    static ECG_HD BB<W> rect() {
        BB<W> r;
#pragma unroll
        for (int i = 0; i < W; i++) r.w[i] = rect_word_c<R0, R1, C0, C1>(i);
        return r;
    }
    template <int R0, int R1, int C0, int C1>
    static ECG_HD uint32_t rect_word_c(int i) {
        // constant-folded per unrolled i
        constexpr uint32_t t0 = rect_word(0, R0, R1, C0, C1), t1 = rect_word(1, R0, R1, C0, C1),
                           t2 = rect_word(2, R0, R1, C0, C1), t3 = rect_word(3, R0, R1, C0, C1),
                           t4 = rect_word(4, R0, R1, C0, C1), t5 = rect_word(5, R0, R1, C0, C1),
                           t6 = rect_word(6, R0, R1, C0, C1), t7 = rect_word(7, R0, R1, C0, C1),
                           t8 = rect_word(8, R0, R1, C0, C1);
        return i == 0 ? t0 : i == 1 ? t1 : i == 2 ? t2 : i == 3 ? t3 : i == 4 ? t4 : i == 5 ? t5 : i == 6 ? t6
               : i == 7 ? t7 : t8;
    }
    static ECG_HD BB<W> valid() { return rect<0, R, 0, C>(); }           // all real cells
    static ECG_HD BB<W> hvalid() { return rect<0, R, 0, C - 1>(); }      // left cell of a horizontal swap
    static ECG_HD BB<W> vvalid() { return rect<0, R - 1, 0, C>(); }      // upper cell of a vertical swap
    static ECG_HD BB<W> col0() { return rect<0, R, 0, 1>(); }
    static ECG_HD BB<W> notlastcol() { return rect<0, R, 0, C - 1>(); }

    // run-time row / column sets (special-token paths)
    static ECG_HD BB<W> rows(int r0, int r1) { // rows [r0, r1), all real columns
        r0 = r0 < 0 ? 0 : r0;
        r1 = r1 > R ? R : r1;
        if (r1 <= r0) return bb_zero<W>();
        return bitrange<W>(r0 * S, r1 * S) & valid();
    }
    static ECG_HD BB<W> cols(int c0, int c1) { // columns [c0, c1), all rows
        c0 = c0 < 0 ? 0 : c0;
        c1 = c1 > C ? C : c1;
        BB<W> r = bb_zero<W>();
        for (int c = c0; c < c1; c++) r |= shl_rt(col0(), c);
        return r;
    }
};

} // namespace ecg
