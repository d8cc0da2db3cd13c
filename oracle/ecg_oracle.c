/*
 * ecg_oracle.c -- CPU ORACLE (test infrastructure, NOT the product). See ecg_oracle.h.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root).  Boards are int64[rows][cols] row-major exactly like the
 * reference's NumPy arrays, so arbitrary cell values (typed specials, zeros)
 * behave as they do there.
 */
#include "ecg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------ config */

/* boardConfig.py:26-43 */
void ecgo_cfg_init(ecgo_cfg *cfg, int rows, int cols, int types) {
    cfg->rows = rows;
    cfg->cols = cols;
    cfg->types = types;
    int bits = 0; /* int(np.ceil(np.log2(types + 1))) */
    while ((1 << bits) < types + 1) bits++;
    cfg->bits = bits;
    cfg->type_mask = (1 << bits) - 1;
    cfg->special_type_mask = (1 << (bits + 1)) + 1 + cfg->type_mask;
    cfg->h_line = cfg->type_mask + 1;
    cfg->v_line = 2 * cfg->h_line;
    cfg->bomb = cfg->special_type_mask;
    cfg->mega_token = cfg->type_mask + cfg->special_type_mask + 1;
    cfg->action_space = rows * (cols - 1) * 2;
}

/* Python int(x) on a float quotient: truncation toward zero */
static int py_int_div(int num, int den) { return (int)trunc((double)num / (double)den); }
/* Python % with a positive modulus */
static int py_mod(int a, int m) {
    int r = a % m;
    return r < 0 ? r + m : r;
}

/* boardConfig.py:45-59 (including the literal "- 3" of :50) */
void ecgo_decode(const ecgo_cfg *cfg, int action, int out[4]) {
    int a = 2 * cfg->cols - 1;
    int b = cfg->cols - 1;
    int row1, col1, row2, col2;
    if (action - a * py_int_div(action, a) >= b) {
        col1 = py_mod(action, a) - b;
        row1 = py_int_div(action - 3 - col1, a);
        col2 = col1;
        row2 = row1 + 1;
    } else {
        col1 = py_mod(action, a);
        row1 = py_int_div(action - col1, a);
        col2 = col1 + 1;
        row2 = row1;
    }
    out[0] = row1;
    out[1] = col1;
    out[2] = row2;
    out[3] = col2;
}

/* boardConfig.py:61-69 */
int ecgo_encode(const ecgo_cfg *cfg, int r1, int c1, int r2, int c2) {
    int a = 2 * cfg->cols - 1;
    int b = (c1 == c2) ? cfg->cols - 1 : 0;
    int rmin = r1 < r2 ? r1 : r2, cmin = c1 < c2 ? c1 : c2;
    return rmin * a + b + cmin;
}

/* --------------------------------------------------------------------- RNG */

/* numpy/random/src/mt19937/mt19937.c: mt19937_seed == init_genrand */
static void mt_seed(ecgo_rng *r, uint32_t seed) {
    r->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->mti = 624;
}

static uint32_t mt_next(ecgo_rng *r) {
    if (r->mti >= 624) {
        uint32_t *mt = r->mt;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->mti = 0;
    }
    uint32_t y = r->mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* Philox4x32-10 (Salmon et al., SC'11), the engine's counter-based refill mode */
void ecgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

void ecgo_rng_init_mt(ecgo_rng *r, uint32_t seed) {
    memset(r, 0, sizeof(*r));
    r->mode = 0;
    r->seed = seed;
    mt_seed(r, seed);
}

void ecgo_rng_init_replay(ecgo_rng *r, const uint32_t *raw, int64_t len) {
    memset(r, 0, sizeof(*r));
    r->mode = 1;
    r->raw = raw;
    r->raw_len = len;
}

void ecgo_rng_init_philox(ecgo_rng *r, uint64_t key, uint64_t board, uint32_t step) {
    memset(r, 0, sizeof(*r));
    r->mode = 2;
    r->key[0] = (uint32_t)key;
    r->key[1] = (uint32_t)(key >> 32);
    r->board[0] = (uint32_t)board;
    r->board[1] = (uint32_t)(board >> 32);
    r->step = step;
    r->blk_idx = -1;
}

/* np.random.seed(cfg.seed) (boardv2.py:20,46; boardFunctions.py:17): the stream
 * restarts.  Philox mode has no reseed: one substream per (board, step). */
void ecgo_rng_reseed(ecgo_rng *r) {
    if (r->mode == 0) {
        mt_seed(r, r->seed);
        r->pos = 0;
    } else if (r->mode == 1) {
        r->pos = 0;
    }
}

uint32_t ecgo_rng_u32(ecgo_rng *r) {
    int64_t k = r->pos++;
    if (r->mode == 0) return mt_next(r);
    if (r->mode == 1) {
        if (k >= r->raw_len) {
            r->overflow = 1;
            return 0;
        }
        return r->raw[k];
    }
    int64_t b = k >> 2;
    if (b != r->blk_idx) {
        uint32_t ctr[4] = {(uint32_t)b, r->step, r->board[0], r->board[1]};
        ecgo_philox4x32_10(ctr, r->key, r->blk);
        r->blk_idx = b;
    }
    return r->blk[k & 3];
}

/* Uniform integer in [0, n).  Modes 0/1: numpy legacy RandomState.randint /
 * random_interval = masked rejection over 32-bit draws
 * (numpy/random/src/distributions/distributions.c, rng = n-1, mask = next 2^k-1;
 * rng == 0 consumes nothing).  Mode 2: multiply-high of one u32. */
uint32_t ecgo_rng_below(ecgo_rng *r, uint32_t n) {
    if (n <= 1) return 0;
    if (r->mode == 2) return (uint32_t)(((uint64_t)ecgo_rng_u32(r) * n) >> 32);
    uint32_t rng = n - 1, mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    for (;;) {
        uint32_t v = ecgo_rng_u32(r) & mask;
        if (v <= rng) return v;
        if (r->overflow) return 0;
    }
}

/* Philox mode only: jump inside the (board, step) substream (see PhiloxRng in the engine's ecg_core.cuh) */
void ecgo_rng_seek(ecgo_rng *r, int64_t pos) {
    if (r->mode == 2) {
        r->pos = pos;
        r->dig_left = 0;
    }
}

/* One refill tile minus 1.  Modes 0/1: np.random.randint(1, types+1) == 1 + below(types).  Mode 2 (engine
 * defined): successive base-n digits of word / 2^32, four per word (uniform up to n^4 / 2^32). */
uint32_t ecgo_rng_digit(ecgo_rng *r, uint32_t n) {
    if (r->mode != 2) return ecgo_rng_below(r, n);
    if (r->dig_left == 0) {
        r->dig_x = ecgo_rng_u32(r);
        r->dig_left = 4;
    }
    r->dig_left--;
    uint64_t p = (uint64_t)r->dig_x * n;
    r->dig_x = (uint32_t)p;
    return (uint32_t)(p >> 32);
}

void ecgo_mt_raw(uint32_t seed, uint32_t *out, int64_t n) {
    ecgo_rng r;
    ecgo_rng_init_mt(&r, seed);
    for (int64_t i = 0; i < n; i++) out[i] = mt_next(&r);
}

/* ----------------------------------------------------------- legal_actions */

/* boardFunctions.py:26-112.  tb = array & type_mask. */
static int horizontal_check(const int64_t *tb, int H, int W, int64_t left_token, int64_t right_token,
                            int l_r, int l_c, int r_r, int r_c) {
#define TB(r, c) tb[(r) * W + (c)]
    /* :41-45 */
    if (l_c - 2 >= 0 && TB(l_r, l_c - 2) == TB(l_r, l_c - 1) && TB(l_r, l_c - 1) == left_token) return 1;
    if (r_c + 2 < W && TB(r_r, r_c + 1) == TB(r_r, r_c + 2) && TB(r_r, r_c + 2) == right_token) return 1;
    /* :47-61 check_above_and_below for (l, left_token) then (r, right_token) */
    for (int side = 0; side < 2; side++) {
        int r = side ? r_r : l_r, c = side ? r_c : l_c;
        int64_t token = side ? right_token : left_token;
        int above = r - 1 >= 0 && TB(r - 1, c) == token;
        int below = r + 1 < H && TB(r + 1, c) == token;
        int res;
        if (!(above || below)) res = 0;
        else if (above && below) res = 1;
        else if (above) res = r - 2 >= 0 && TB(r - 2, c) == token;
        else res = r + 2 < H && TB(r + 2, c) == token;
        if (res) return 1;
    }
    return 0;
}

static int vertical_check(const int64_t *tb, int H, int W, int64_t above_token, int64_t below_token,
                          int a_r, int a_c, int b_r, int b_c) {
    /* :74-78 */
    if (b_r + 2 < H && TB(b_r + 1, b_c) == TB(b_r + 2, b_c) && TB(b_r + 2, b_c) == below_token) return 1;
    if (a_r - 2 >= 0 && TB(a_r - 2, a_c) == TB(a_r - 1, a_c) && TB(a_r - 1, a_c) == above_token) return 1;
    /* :80-94 check_left_and_right for (below, below_token) then (above, above_token) */
    for (int side = 0; side < 2; side++) {
        int r = side ? a_r : b_r, c = side ? a_c : b_c;
        int64_t token = side ? above_token : below_token;
        int left = c - 1 >= 0 && TB(r, c - 1) == token;
        int right = c + 1 < W && TB(r, c + 1) == token;
        int res;
        if (!(left || right)) res = 0;
        else if (left && right) res = 1;
        else if (left) res = c - 2 >= 0 && TB(r, c - 2) == token;
        else res = c + 2 < W && TB(r, c + 2) == token;
        if (res) return 1;
    }
    return 0;
#undef TB
}

int ecgo_legal_actions(const ecgo_cfg *cfg, const int64_t *arr, int *out) {
    int H = cfg->rows, W = cfg->cols, n = 0;
    int64_t tb[ECGO_MAX_CELLS];
    for (int i = 0; i < H * W; i++) tb[i] = arr[i] & cfg->type_mask; /* :96 */
    for (int action = 0; action < cfg->action_space; action++) {     /* :97 cfg.actions.items() */
        int d[4];
        ecgo_decode(cfg, action, d);
        int r1 = d[0], c1 = d[1], r2 = d[2], c2 = d[3];
        int64_t t1 = tb[r1 * W + c1], t2 = tb[r2 * W + c2];
        /* :100 special tokens */
        if (t1 == 0 || t2 == 0 || (arr[r1 * W + c1] > cfg->type_mask && arr[r2 * W + c2] > cfg->type_mask)) {
            out[n++] = action;
            continue;
        }
        if (t1 == t2) continue; /* :103 */
        if (c1 == c2) {         /* :105 vertical: vertical_check(token2, token1, cell1, cell2) */
            /* signature (above_token, below_token, above, below); inside, :74 binds b_* to `below`
             * (= cell2, the lower cell) and a_* to `above` (= cell1): the token of the upper cell
             * lands on the lower cell and vice versa. */
            if (vertical_check(tb, H, W, t2, t1, r1, c1, r2, c2)) out[n++] = action;
        } else { /* :109 horizontal_check(token2, token1, cell1, cell2) */
            if (horizontal_check(tb, H, W, t2, t1, r1, c1, r2, c2)) out[n++] = action;
        }
    }
    return n;
}

/* ----------------------------------------------------- swap / get_matches */

/* boardFunctions.py:115-118 */
void ecgo_swap(const ecgo_cfg *cfg, const int64_t *in, int r1, int c1, int r2, int c2, int64_t *out) {
    int W = cfg->cols;
    if (out != in) memcpy(out, in, sizeof(int64_t) * cfg->rows * W);
    int64_t a = in[r1 * W + c1], b = in[r2 * W + c2];
    out[r1 * W + c1] = b;
    out[r2 * W + c2] = a;
}

typedef struct {
    int n_groups;
    int start[ECGO_MAX_CELLS + 1]; /* not contiguous: each group owns cap cells in pool */
    int len[ECGO_MAX_CELLS];
    int cap[ECGO_MAX_CELLS];
    int16_t *cells; /* pool: (r<<8|c) */
    int pool_used, pool_cap;
} groups_t;

static void groups_init(groups_t *g) {
    g->n_groups = 0;
    g->pool_cap = 4096;
    g->pool_used = 0;
    g->cells = (int16_t *)malloc(sizeof(int16_t) * g->pool_cap);
}
static void groups_free(groups_t *g) { free(g->cells); }

static void group_reserve(groups_t *g, int idx, int extra) {
    if (g->len[idx] + extra <= g->cap[idx]) return;
    int ncap = (g->len[idx] + extra) * 2 + 8;
    if (g->pool_used + ncap > g->pool_cap) {
        while (g->pool_used + ncap > g->pool_cap) g->pool_cap *= 2;
        g->cells = (int16_t *)realloc(g->cells, sizeof(int16_t) * g->pool_cap);
    }
    memcpy(g->cells + g->pool_used, g->cells + g->start[idx], sizeof(int16_t) * g->len[idx]);
    g->start[idx] = g->pool_used;
    g->cap[idx] = ncap;
    g->pool_used += ncap;
}

static int group_contains(const groups_t *g, int idx, int16_t cell) {
    const int16_t *p = g->cells + g->start[idx];
    for (int i = 0; i < g->len[idx]; i++)
        if (p[i] == cell) return 1;
    return 0;
}

/* boardFunctions.py:126-131 add_to_matches: merge into the first group sharing a cell;
 * `item not in matches` (:129) compares a tuple with lists -> always true -> every cell appended. */
static void add_to_matches(groups_t *g, const int16_t *match, int n) {
    for (int idx = 0; idx < g->n_groups; idx++) {
        int any = 0;
        for (int i = 0; i < n && !any; i++) any = group_contains(g, idx, match[i]);
        if (any) {
            group_reserve(g, idx, n);
            memcpy(g->cells + g->start[idx] + g->len[idx], match, sizeof(int16_t) * n);
            g->len[idx] += n;
            return;
        }
    }
    int idx = g->n_groups++;
    g->start[idx] = g->pool_used;
    g->len[idx] = 0;
    g->cap[idx] = 0;
    group_reserve(g, idx, n);
    memcpy(g->cells + g->start[idx], match, sizeof(int16_t) * n);
    g->len[idx] = n;
}

/* boardFunctions.py:121-156 */
static void get_matches(int rows, int cols, const int64_t *arr, uint8_t *mask, groups_t *g) {
    memset(mask, 0, (size_t)rows * cols);
    int16_t match[2 * ECGO_MAX_DIM + 2];
    for (int row = 0; row < rows; row++) {
        for (int col = 0; col < cols; col++) {
            int64_t value = arr[row * cols + col];
            if (value == 0) continue; /* :136 */
            int16_t me = (int16_t)((row << 8) | col);
            int seen = 0;
            for (int idx = 0; idx < g->n_groups && !seen; idx++) seen = group_contains(g, idx, me);
            if (seen) continue;
            int n = 0;
            /* :140 horizontal */
            if (col <= cols - 3 && arr[row * cols + col + 1] == value && arr[row * cols + col + 2] == value) {
                int k = col;
                while (k < cols && arr[row * cols + k] == value) {
                    match[n++] = (int16_t)((row << 8) | k);
                    mask[row * cols + k] = 1;
                    k++;
                }
            }
            /* :148 vertical */
            if (row <= rows - 3 && arr[(row + 1) * cols + col] == value && arr[(row + 2) * cols + col] == value) {
                int k = row;
                while (k < rows && arr[k * cols + col] == value) {
                    match[n++] = (int16_t)((k << 8) | col);
                    mask[k * cols + col] = 1;
                    k++;
                }
            }
            if (n > 2) add_to_matches(g, match, n); /* :154 */
        }
    }
}

static int cmp_i16(const void *a, const void *b) { return (int)*(const int16_t *)a - (int)*(const int16_t *)b; }

/* boardFunctions.py:159-169 get_match_spawn_mask, :8-13 get_center */
static void get_match_spawn_mask(const ecgo_cfg *cfg, groups_t *g, int32_t *spawn) {
    memset(spawn, 0, sizeof(int32_t) * cfg->rows * cfg->cols);
    for (int idx = 0; idx < g->n_groups; idx++) {
        int n = g->len[idx];
        if (n <= 3) continue; /* :161 len(match) > 3 */
        int16_t *p = g->cells + g->start[idx];
        qsort(p, n, sizeof(int16_t), cmp_i16); /* sort by (row, col) == by (row<<8|col) */
        int16_t center = p[n / 2];
        int cr = center >> 8, cc = center & 0xff;
        int same_row = 1, same_col = 1;
        for (int i = 0; i < n; i++) {
            if ((p[i] >> 8) != (p[0] >> 8)) same_row = 0;
            if ((p[i] & 0xff) != (p[0] & 0xff)) same_col = 0;
        }
        int32_t v;
        if (same_row) v = (int32_t)(n > 4 ? cfg->mega_token : cfg->v_line);      /* :163-164 */
        else if (same_col) v = (int32_t)(n > 4 ? cfg->mega_token : cfg->h_line); /* :165-166 */
        else v = (int32_t)cfg->bomb;                                              /* :168 */
        spawn[cr * cfg->cols + cc] = v;
    }
}

int ecgo_get_matches(int rows, int cols, const int64_t *arr, uint8_t *mask, int *ncells_out) {
    groups_t g;
    groups_init(&g);
    get_matches(rows, cols, arr, mask, &g);
    int n = g.n_groups;
    if (ncells_out)
        for (int i = 0; i < n; i++) ncells_out[i] = g.len[i];
    groups_free(&g);
    return n;
}

int ecgo_matches_and_spawn(const ecgo_cfg *cfg, const int64_t *arr, uint8_t *mask, int32_t *spawn) {
    groups_t g;
    groups_init(&g);
    get_matches(cfg->rows, cfg->cols, arr, mask, &g);
    get_match_spawn_mask(cfg, &g, spawn);
    int n = g.n_groups;
    groups_free(&g);
    return n;
}

/* ---------------------------------------------------------------- shuffle */

/* boardFunctions.py:16-23: reseed; np.random.shuffle(array) permutes ROWS with numpy's
 * legacy Fisher-Yates (i = n-1..1, j = random_interval(i)); cells that held specials
 * before the shuffle get their old special value back. */
void ecgo_shuffle(const ecgo_cfg *cfg, ecgo_rng *rng, int64_t *arr) {
    int H = cfg->rows, W = cfg->cols;
    ecgo_rng_reseed(rng);
    int64_t special[ECGO_MAX_CELLS];
    uint8_t smask[ECGO_MAX_CELLS];
    for (int i = 0; i < H * W; i++) {
        smask[i] = arr[i] > cfg->type_mask;
        special[i] = smask[i] ? (int64_t)(int32_t)arr[i] : 0;
    }
    int64_t buf[ECGO_MAX_DIM];
    for (int i = H - 1; i >= 1; i--) {
        int j = (int)ecgo_rng_below(rng, (uint32_t)i + 1);
        memcpy(buf, arr + j * W, sizeof(int64_t) * W);
        memcpy(arr + j * W, arr + i * W, sizeof(int64_t) * W);
        memcpy(arr + i * W, buf, sizeof(int64_t) * W);
    }
    for (int i = 0; i < H * W; i++)
        if (smask[i]) arr[i] = special[i];
}

/* --------------------------------------------------------------- init board */

/* boardv2.py:20-27 */
void ecgo_init_board(const ecgo_cfg *cfg, ecgo_rng *rng, int64_t *out) {
    int n = cfg->rows * cfg->cols;
    ecgo_rng_reseed(rng);
    for (int i = 0; i < n; i++) out[i] = 1 + ecgo_rng_below(rng, (uint32_t)cfg->types);
    uint8_t mask[ECGO_MAX_CELLS];
    int64_t fresh[ECGO_MAX_CELLS];
    for (int guard = 0; guard < 100000; guard++) {
        int ng = ecgo_get_matches(cfg->rows, cfg->cols, out, mask, NULL);
        if (ng == 0 || rng->overflow) break;
        for (int i = 0; i < n; i++) fresh[i] = 1 + ecgo_rng_below(rng, (uint32_t)cfg->types);
        for (int i = 0; i < n; i++)
            if (mask[i]) out[i] = fresh[i];
    }
}

/* ------------------------------------------------------------ apply_action */

/* Python slice [start:stop] on an axis of length n -> [lo, hi) */
static void py_slice(int start, int stop, int n, int *lo, int *hi) {
    if (start < 0) {
        start += n;
        if (start < 0) start = 0;
    }
    if (stop < 0) {
        stop += n;
        if (stop < 0) stop = 0;
    }
    if (start > n) start = n;
    if (stop > n) stop = n;
    *lo = start;
    *hi = stop > start ? stop : start;
}

static int lower_clamp(int v) { return v < 0 ? 0 : v; }               /* util/quickMath.py:1-2 */
static int upper_clamp(int v, int m) { return v > m ? m : v; }        /* util/quickMath.py:5-6 */

/* boardv2.py:58-65 */
static int64_t point_of(const ecgo_cfg *cfg, int64_t x) {
    if (x <= cfg->type_mask) return 2;
    if (x == cfg->mega_token) return 250;
    if (x < cfg->special_type_mask) return 25;
    return 50;
}

/* boardv2.py:43-207.  `in` is not modified; `out` receives the next board.
 * Returns status bits.  The caller handles is_terminal (:44). */
int ecgo_apply_action(const ecgo_cfg *cfg, ecgo_rng *rng, const int64_t *in, int action, int64_t *out,
                      int64_t *reward_out, int *cascades_out, int *draws_out) {
    const int H = cfg->rows, W = cfg->cols, N = H * W;
    const int64_t type_mask = cfg->type_mask, stm = cfg->special_type_mask;
    const int64_t h_line = cfg->h_line, v_line = cfg->v_line, bomb = cfg->bomb, mega = cfg->mega_token;
    int status = 0, cascades = 0, draws = 0;
    int64_t reward = 0;

    if (action < 0 || action >= cfg->action_space) { /* :48 KeyError */
        if (out != in) memcpy(out, in, sizeof(int64_t) * N);
        *reward_out = 0;
        if (cascades_out) *cascades_out = 0;
        if (draws_out) *draws_out = 0;
        return ECGO_ST_BAD_ACTION;
    }
    ecgo_rng_reseed(rng); /* :46 */
    int d[4];
    ecgo_decode(cfg, action, d);
    const int sr = d[0], sc = d[1], tr = d[2], tc = d[3]; /* source, target */

    int64_t next_state[ECGO_MAX_CELLS], points_board[ECGO_MAX_CELLS], special_tokens[ECGO_MAX_CELLS],
        token_board[ECGO_MAX_CELLS];
    int32_t token_spawn[ECGO_MAX_CELLS];
    uint8_t zeros_mask[ECGO_MAX_CELLS];
    ecgo_swap(cfg, in, sr, sc, tr, tc, next_state); /* :51 */

#define REBUILD_SUB_BOARDS()                                                   \
    for (int i = 0; i < N; i++) {                                              \
        points_board[i] = point_of(cfg, next_state[i]);                        \
        special_tokens[i] = next_state[i] > type_mask ? next_state[i] : 0;     \
        token_board[i] = next_state[i] & type_mask;                            \
    }
    REBUILD_SUB_BOARDS(); /* :68-70 */
    memset(token_spawn, 0, sizeof(int32_t) * N); /* :71 */

    const int64_t token1 = in[sr * W + sc], token2 = in[tr * W + tc];                        /* :73 */
    const int64_t t1t = special_tokens[sr * W + sc], t2t = special_tokens[tr * W + tc];      /* :74 */
#define ARE(a, b) ((t1t == (a) && t2t == (b)) || (t2t == (a) && t1t == (b)))               /* :76-77 */

    groups_t g;
    groups_init(&g);
    if (ARE(mega, mega)) { /* :81 */
        for (int i = 0; i < N; i++) token_board[i] = 0;
    } else if (ARE(mega, bomb)) { /* :84-89 */
        int64_t token = token1 > token2 ? token1 : token2;
        for (int i = 0; i < N; i++)
            if (token_board[i] == token && special_tokens[i] == 0) special_tokens[i] = token + bomb;
    } else if (ARE(mega, h_line) || ARE(mega, v_line)) { /* :91-99 */
        int64_t token = token1 > token2 ? token1 : token2;
        uint8_t m[ECGO_MAX_CELLS];
        for (int i = 0; i < N; i++) m[i] = (token_board[i] == token && special_tokens[i] == 0);
        for (int i = 0; i < N; i++)
            if (m[i]) token_board[i] = 0;
        int n = 0;
        for (int i = 0; i < N; i++) {
            if (!m[i]) continue;
            if (special_tokens[i] == 0) special_tokens[i] = (n % 2 == 0) ? v_line : h_line;
            n++;
        }
    } else if (ARE(mega, 0)) { /* :101-103 */
        int64_t token = token1 > token2 ? token1 : token2;
        for (int i = 0; i < N; i++)
            if (token_board[i] == token) token_board[i] = 0;
    } else if (ARE(bomb, bomb)) { /* :112-116 */
        int r0 = lower_clamp(tr - 2), r1 = upper_clamp(tr + 2, H);
        int c0 = lower_clamp(tc - 2), c1 = upper_clamp(tc + 2, W);
        for (int r = r0; r < r1; r++)
            for (int c = c0; c < c1; c++) token_board[r * W + c] = 0;
    } else if (ARE(bomb, h_line) || ARE(bomb, v_line)) { /* :123-125 */
        int c0 = lower_clamp(tc - 2), c1 = upper_clamp(tc + 2, W);
        for (int r = 0; r < H; r++)
            for (int c = c0; c < c1; c++) token_board[r * W + c] = 0;
        int r0 = lower_clamp(tr - 2), r1 = upper_clamp(tr + 2, H);
        for (int r = r0; r < r1; r++)
            for (int c = 0; c < W; c++) token_board[r * W + c] = 0;
    } else if (ARE(h_line, v_line)) { /* :130-132: token_board[:target[1]] and token_board[target[0]:] are ROW slices */
        int lo, hi;
        py_slice(0, tc, H, &lo, &hi);
        for (int r = lo; r < hi; r++)
            for (int c = 0; c < W; c++) token_board[r * W + c] = 0;
        py_slice(tr, H, H, &lo, &hi);
        for (int r = lo; r < hi; r++)
            for (int c = 0; c < W; c++) token_board[r * W + c] = 0;
    } else { /* :134-136 */
        get_matches(H, W, token_board, zeros_mask, &g);
        for (int i = 0; i < N; i++)
            if (zeros_mask[i]) token_board[i] = 0;
        get_match_spawn_mask(cfg, &g, token_spawn);
    }

    for (;;) { /* :138 */
        cascades++;
        /* :141-154 trigger pass; the list is fixed before any effect is applied */
        for (int i = 0; i < N; i++)
            if (token_board[i] != 0) special_tokens[i] = 0;
        for (int i = 0; i < H; i++) {
            for (int j = 0; j < W; j++) {
                if (special_tokens[i * W + j] == 0) continue;
                int64_t st = special_tokens[i * W + j] & stm; /* :144 */
                if (st == h_line) {
                    for (int c = 0; c < W; c++) token_board[i * W + c] = 0;
                } else if (st == v_line) {
                    for (int r = 0; r < H; r++) token_board[r * W + j] = 0;
                } else if (st == bomb) { /* :151-154: [start_col:end_col, start_row:end_row] (transposed) */
                    int r0, r1, c0, c1;
                    py_slice(j - 1, j + 1, H, &r0, &r1);
                    py_slice(i - 1, i + 1, W, &c0, &c1);
                    for (int r = r0; r < r1; r++)
                        for (int c = c0; c < c1; c++) token_board[r * W + c] = 0;
                }
            }
        }
        /* :157-158 */
        for (int i = 0; i < N; i++)
            if (token_board[i] == 0) reward += points_board[i];
        /* :161-163 */
        for (int i = 0; i < N; i++)
            if (token_board[i] == 0) next_state[i] = 0;
        for (int i = 0; i < N; i++)
            if (token_spawn[i] != 0) next_state[i] += token_spawn[i];
        for (int i = 0; i < N; i++) next_state[i] = next_state[i] < 0 ? 0 : (next_state[i] > 32 ? 32 : next_state[i]);
        /* :166-173 gravity + refill, columns left to right, first draw = topmost cell */
        if (rng->mode != 2) {
            for (int col = 0; col < W; col++) {
                int64_t tokens[ECGO_MAX_DIM];
                int nt = 0;
                for (int r = 0; r < H; r++)
                    if (next_state[r * W + col] > 0) tokens[nt++] = next_state[r * W + col];
                if (nt == H) continue;
                int k = H - nt;
                for (int r = 0; r < k; r++) {
                    next_state[r * W + col] = 1 + ecgo_rng_digit(rng, (uint32_t)cfg->types);
                    draws++;
                }
                for (int r = 0; r < nt; r++) next_state[(k + r) * W + col] = tokens[r];
            }
        } else {
            /* Philox mode (engine-defined addressing): same gravity; the i.i.d. tiles are assigned to the
             * holes in row-major order from words (iteration * 2048 + d) of the (board, step) substream */
            for (int col = 0; col < W; col++) {
                int64_t tokens[ECGO_MAX_DIM];
                int nt = 0;
                for (int r = 0; r < H; r++)
                    if (next_state[r * W + col] > 0) tokens[nt++] = next_state[r * W + col];
                int k = H - nt;
                for (int r = 0; r < k; r++) next_state[r * W + col] = 0;
                for (int r = 0; r < nt; r++) next_state[(k + r) * W + col] = tokens[r];
            }
            ecgo_rng_seek(rng, (int64_t)(cascades - 1) * 2048 + 1);
            for (int i = 0; i < N; i++)
                if (next_state[i] == 0) {
                    next_state[i] = 1 + ecgo_rng_digit(rng, (uint32_t)cfg->types);
                    draws++;
                }
        }
        REBUILD_SUB_BOARDS(); /* :176-178 */
        g.n_groups = 0;
        g.pool_used = 0;
        get_matches(H, W, token_board, zeros_mask, &g); /* :181 */
        /* :188-194 shuffle while no matches and no legal action */
        int shuffles = 0;
        while (g.n_groups == 0) {
            int legal[ECGO_MAX_ACTIONS];
            if (ecgo_legal_actions(cfg, next_state, legal) != 0) break;
            if (shuffles++ >= ECGO_SHUFFLE_CAP) { /* the reference has no cap and can spin forever */
                status |= ECGO_ST_SHUFFLE_CAP;
                break;
            }
            if (shuffles == 1) ecgo_rng_seek(rng, (int64_t)(cascades - 1) * 2048 + 1024);
            ecgo_shuffle(cfg, rng, next_state);
            REBUILD_SUB_BOARDS();
            g.n_groups = 0;
            g.pool_used = 0;
            get_matches(H, W, token_board, zeros_mask, &g);
        }
        if (g.n_groups == 0) break; /* :195 */
        if (cascades >= ECGO_CASCADE_CAP) { /* the reference has no cap (tiny type counts cascade ~forever) */
            status |= ECGO_ST_CASCADE_CAP;
            break;
        }
        for (int i = 0; i < N; i++)
            if (zeros_mask[i]) token_board[i] = 0; /* :199 */
        get_match_spawn_mask(cfg, &g, token_spawn); /* :202 */
        if (rng->overflow) break;
    }
    groups_free(&g);
    if (rng->overflow) status |= ECGO_ST_STREAM_OVERFLOW;
    memcpy(out, next_state, sizeof(int64_t) * N); /* :205 */
    *reward_out = reward;
    if (cascades_out) *cascades_out = cascades;
    if (draws_out) *draws_out = draws;
    return status;
#undef ARE
#undef REBUILD_SUB_BOARDS
}

/* ------------------------------------------------------------ batch helpers */

static int g_threads = 0; /* 0 = all online cores */

void ecgo_set_threads(int n) { g_threads = n; }

int ecgo_max_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

typedef void (*range_fn)(void *ctx, int64_t lo, int64_t hi);
typedef struct {
    range_fn fn;
    void *ctx;
    int64_t n, chunk;
    volatile int64_t *next;
} pf_arg;

static void *pf_worker(void *p) {
    pf_arg *a = (pf_arg *)p;
    for (;;) {
        int64_t lo = __atomic_fetch_add(a->next, a->chunk, __ATOMIC_RELAXED);
        if (lo >= a->n) break;
        int64_t hi = lo + a->chunk < a->n ? lo + a->chunk : a->n;
        a->fn(a->ctx, lo, hi);
    }
    return NULL;
}

/* dynamic-chunk parallel for over [0, n) on ecgo_max_threads() pthreads */
static void parallel_for(int64_t n, int64_t chunk, range_fn fn, void *ctx) {
    int nt = ecgo_max_threads();
    if (nt > 256) nt = 256;
    if (n <= chunk || nt <= 1) {
        fn(ctx, 0, n);
        return;
    }
    volatile int64_t next = 0;
    pf_arg a = {fn, ctx, n, chunk, &next};
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nt - 1; t++)
        if (pthread_create(&th[started], NULL, pf_worker, &a) == 0) started++;
    pf_worker(&a);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

typedef struct {
    const ecgo_cfg *cfg;
    const int64_t *boards;
    uint8_t *mask;
} lm_ctx;

static void lm_range(void *p, int64_t lo, int64_t hi) {
    lm_ctx *c = (lm_ctx *)p;
    const int N = c->cfg->rows * c->cfg->cols, A = c->cfg->action_space;
    for (int64_t i = lo; i < hi; i++) {
        int legal[ECGO_MAX_ACTIONS];
        int k = ecgo_legal_actions(c->cfg, c->boards + i * N, legal);
        memset(c->mask + i * A, 0, (size_t)A);
        for (int j = 0; j < k; j++) c->mask[i * A + legal[j]] = 1;
    }
}

/* mask[n][action_space] bytes */
void ecgo_legal_mask_batch(const ecgo_cfg *cfg, const int64_t *boards, uint8_t *mask, int64_t n) {
    lm_ctx c = {cfg, boards, mask};
    parallel_for(n, 256, lm_range, &c);
}

typedef struct {
    const ecgo_cfg *cfg;
    int rng_mode;
    const uint32_t *raw;
    int64_t raw_stride, raw_len;
    uint64_t philox_key, board0;
    uint32_t step_ctr;
    const int64_t *in;
    const int32_t *actions, *moves_left;
    int64_t *out, *reward;
    int32_t *cascades;
    uint8_t *status, *legal_out;
} sb_ctx;

static void sb_range(void *p, int64_t lo, int64_t hi) {
    sb_ctx *s = (sb_ctx *)p;
    const ecgo_cfg *cfg = s->cfg;
    const int N = cfg->rows * cfg->cols, A = cfg->action_space;
    for (int64_t i = lo; i < hi; i++) {
        ecgo_rng rng;
        if (s->rng_mode == 0) ecgo_rng_init_mt(&rng, s->raw[i * s->raw_stride]);
        else if (s->rng_mode == 1) ecgo_rng_init_replay(&rng, s->raw + i * s->raw_stride, s->raw_len);
        else ecgo_rng_init_philox(&rng, s->philox_key, s->board0 + (uint64_t)i, s->step_ctr);
        int64_t r = 0;
        int c = 0, st;
        if (s->moves_left && s->moves_left[i] < 1) { /* boardv2.py:44 */
            if (s->out != s->in) memcpy(s->out + i * N, s->in + i * N, sizeof(int64_t) * N);
            st = ECGO_ST_TERMINAL;
        } else {
            int64_t tmp[ECGO_MAX_CELLS];
            st = ecgo_apply_action(cfg, &rng, s->in + i * N, s->actions[i], tmp, &r, &c, NULL);
            memcpy(s->out + i * N, tmp, sizeof(int64_t) * N);
        }
        if (s->reward) s->reward[i] = r;
        if (s->cascades) s->cascades[i] = c;
        if (s->status) s->status[i] = (uint8_t)st;
        if (s->legal_out) {
            int legal[ECGO_MAX_ACTIONS];
            int k = ecgo_legal_actions(cfg, s->out + i * N, legal);
            memset(s->legal_out + i * A, 0, (size_t)A);
            for (int j = 0; j < k; j++) s->legal_out[i * A + legal[j]] = 1;
        }
    }
}

/* One lockstep step of n boards.  moves_left may be NULL (never terminal).
 * rng_mode 0: live MT19937 seeded per board with raw[i*raw_stride] taken as the seed
 * (raw_len ignored); 1: replay of the raw u32 stream raw + i*raw_stride (stride 0 = one
 * shared stream); 2: philox(key, board0+i, step_ctr). */
void ecgo_step_batch(const ecgo_cfg *cfg, int rng_mode, const uint32_t *raw, int64_t raw_stride, int64_t raw_len,
                     uint64_t philox_key, uint64_t board0, uint32_t step_ctr, const int64_t *in,
                     const int32_t *actions, const int32_t *moves_left, int64_t *out, int64_t *reward,
                     int32_t *cascades, uint8_t *status, uint8_t *legal_out, int64_t n) {
    sb_ctx s = {cfg, rng_mode, raw, raw_stride, raw_len, philox_key, board0, step_ctr, in,
                actions, moves_left, out, reward, cascades, status, legal_out};
    parallel_for(n, 64, sb_range, &s);
}

typedef struct {
    const ecgo_cfg *cfg;
    uint64_t key, board0;
    uint32_t step0;
    int n_moves;
    int64_t *boards, *reward, *steps;
} pe_ctx;

static void pe_range(void *p, int64_t lo, int64_t hi) {
    pe_ctx *c = (pe_ctx *)p;
    const int N = c->cfg->rows * c->cfg->cols;
    for (int64_t i = lo; i < hi; i++) {
        int64_t st = 0;
        int64_t r = ecgo_philox_episode(c->cfg, c->key, c->board0 + (uint64_t)i, c->step0, c->n_moves, c->boards + i * N, &st);
        if (c->reward) c->reward[i] = r;
        if (c->steps) c->steps[i] = st;
    }
}

/* n independent Philox lockstep episodes on all host threads (bench cpu_baseline / --impl reference) */
void ecgo_philox_episode_batch(const ecgo_cfg *cfg, uint64_t key, uint64_t board0, uint32_t step0, int n_moves,
                               int64_t *boards, int64_t *reward, int64_t *steps, int64_t n) {
    pe_ctx c = {cfg, key, board0, step0, n_moves, boards, reward, steps};
    parallel_for(n, 16, pe_range, &c);
}

typedef struct {
    const ecgo_cfg *cfg;
    const uint32_t *seeds;
    int n_moves;
    int64_t *reward, *steps;
} re_ctx;

static void re_range(void *p, int64_t lo, int64_t hi) {
    re_ctx *c = (re_ctx *)p;
    for (int64_t i = lo; i < hi; i++) {
        int64_t st = 0;
        int64_t r = ecgo_random_episode(c->cfg, c->seeds[i], c->n_moves, &st, NULL);
        if (c->reward) c->reward[i] = r;
        if (c->steps) c->steps[i] = st;
    }
}

/* n random_task episodes (samplerTasks.py:9-14) = the reference's own CPU workload
 * (util/multiprocessingAutoBatcher.py:37-56 fans these over processes; here: threads) */
void ecgo_random_episode_batch(const ecgo_cfg *cfg, const uint32_t *seeds, int n_moves, int64_t *reward,
                               int64_t *steps, int64_t n) {
    re_ctx c = {cfg, seeds, n_moves, reward, steps};
    parallel_for(n, 8, re_range, &c);
}

/* samplerTasks.py:9-14 random_task with an explicit seed:
 *   state = BoardV2(n_moves, BoardConfig(seed)); np.random.seed(seed)
 *   while not terminal: state = state.apply_action(np.random.choice(state.legal_actions))
 * np.random.choice(list) == list[randint(0, len)] on the same legacy stream. */
int64_t ecgo_random_episode(const ecgo_cfg *cfg, uint32_t seed, int n_moves, int64_t *steps, int64_t *final_board) {
    ecgo_rng rng;
    ecgo_rng_init_mt(&rng, seed);
    int64_t board[ECGO_MAX_CELLS], next[ECGO_MAX_CELLS], total = 0;
    ecgo_init_board(cfg, &rng, board);
    ecgo_rng_reseed(&rng); /* samplerTasks.py:11 */
    int64_t nsteps = 0;
    for (int m = n_moves; m >= 1; m--) {
        int legal[ECGO_MAX_ACTIONS];
        int k = ecgo_legal_actions(cfg, board, legal);
        if (k == 0) break; /* reference: ValueError */
        int a = legal[ecgo_rng_below(&rng, (uint32_t)k)];
        int64_t r;
        ecgo_apply_action(cfg, &rng, board, a, next, &r, NULL, NULL);
        memcpy(board, next, sizeof(int64_t) * cfg->rows * cfg->cols);
        total += r;
        nsteps++;
    }
    if (steps) *steps = nsteps;
    if (final_board) memcpy(final_board, board, sizeof(int64_t) * cfg->rows * cfg->cols);
    return total;
}

/* The engine's Philox lockstep episode: at step t the action is the
 * idx-th legal action (ascending), idx = mulhi(word 0 of the (board, t) substream, n_legal);
 * refills come from the same substream (words j*2048 + 1.. in cascade iteration j). */
int64_t ecgo_philox_episode(const ecgo_cfg *cfg, uint64_t key, uint64_t board_index, uint32_t step0, int n_moves,
                            int64_t *board_io, int64_t *steps) {
    int64_t next[ECGO_MAX_CELLS], total = 0, nsteps = 0;
    const uint32_t k2[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    for (uint32_t t = step0; t < step0 + (uint32_t)n_moves; t++) {
        int legal[ECGO_MAX_ACTIONS];
        int k = ecgo_legal_actions(cfg, board_io, legal);
        if (k == 0) break;
        uint32_t ctr[4] = {0u, t, (uint32_t)board_index, (uint32_t)(board_index >> 32)}, o[4];
        ecgo_philox4x32_10(ctr, k2, o);
        /* the engine counts the pick in swap-bitboard order: all horizontal swaps by (row, col), then all
         * vertical swaps by (row, col); within each class ascending action id is already (row, col) order */
        int ordered[ECGO_MAX_ACTIONS], no = 0;
        const int per_row = 2 * cfg->cols - 1;
        for (int j = 0; j < k; j++)
            if (legal[j] % per_row < cfg->cols - 1) ordered[no++] = legal[j];
        for (int j = 0; j < k; j++)
            if (legal[j] % per_row >= cfg->cols - 1) ordered[no++] = legal[j];
        int a = ordered[(uint32_t)(((uint64_t)o[0] * (uint32_t)k) >> 32)];
        ecgo_rng rng;
        ecgo_rng_init_philox(&rng, key, board_index, t);
        int64_t r;
        ecgo_apply_action(cfg, &rng, board_io, a, next, &r, NULL, NULL);
        memcpy(board_io, next, sizeof(int64_t) * cfg->rows * cfg->cols);
        total += r;
        nsteps++;
    }
    if (steps) *steps = nsteps;
    return total;
}
