// Host-only experiment (no GPU): how many gravity-loop moves and refill deposits ONE cascade iteration of the
// common-case kernel needs per board, over natural Philox play at 9x9x6, and what a warp pays for them -- the
// loops run to the MAX over the lanes that share a warp.  Variants: a pre-pass that lets every cell above three
// (two) vertically contiguous holes fall by three (two) rows at once; 8 / 16 / 32 boards per warp (the sub-warp
// mappings of VERDICT r1 item 3 iv).  Build and run:
//   g++ -O2 -std=c++17 -o /tmp/grm scripts/experiments/gravity_refill_model.cpp && /tmp/grm 30000
// Result of the committed run: profiles/r09_gravity_refill_model.txt
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#define ECG_SIM_HOOKS 1
#include "../../element-crush-gym_b200/csrc/ecg_core.cuh"
using namespace ecg;
using SH = Shape<9, 9, 3, false>;
using G = SH::G;
constexpr int W = G::W, S = G::S;

static int grav_iters(Board<G> b) {  // current loop: number of moves
    int it = 0;
    for (;;) {
        const BB<W> occ = b.p[0] | b.p[1] | b.p[2] | b.p[3];
        const BB<W> holes = andn(G::valid(), occ);
        BB<W> u = shr<S>(holes);
        u |= shr<S>(u); u |= shr<2 * S>(u); u |= shr<4 * S>(u);
        const BB<W> f = occ & u;
        if (!any(f)) return it;
        for (int k = 0; k < 4; k++) b.p[k] = andn(b.p[k], f) | shl<S>(b.p[k] & f);
        it++;
    }
}
template <int K>
static bool prepass(Board<G> &b) { // cells above K vertically contiguous holes fall by K
    const BB<W> occ = b.p[0] | b.p[1] | b.p[2] | b.p[3];
    const BB<W> holes = andn(G::valid(), occ);
    BB<W> v = holes;
    if (K >= 2) v = v & shr<S>(holes);
    if (K >= 3) v = v & shr<2 * S>(holes);
    BB<W> up = shr<S>(v);
    up |= shr<S>(up); up |= shr<2 * S>(up); up |= shr<4 * S>(up);
    const BB<W> f = occ & up;
    if (!any(f)) return false;
    for (int k = 0; k < 4; k++) b.p[k] = andn(b.p[k], f) | shl<K * S>(b.p[k] & f);
    return true;
}
struct Rec { int g, g3, g32, holes, deep, cls; };
static int g_cls = 0; // 1: the step started with a special token on the board
static std::vector<Rec> recs;
namespace ecg { void ecg_sim_gravity_hook(const Board<G> &b) {
    Rec r;
    r.g = grav_iters(b);
    Board<G> c = b; prepass<3>(c); r.g3 = grav_iters(c);
    Board<G> d = b; prepass<3>(d); prepass<2>(d); r.g32 = grav_iters(d);
    const BB<W> occ = b.p[0] | b.p[1] | b.p[2] | b.p[3];
    const BB<W> holes = andn(G::valid(), occ);
    r.holes = popcount(holes);
    r.deep = 0;
    r.cls = g_cls;
    recs.push_back(r);
} }

int main(int argc, char **argv) {
    const int NB = argc > 1 ? atoi(argv[1]) : 20000, STEPS = 20;
    const uint64_t key = 12345;
    long handoffs = 0, steps = 0, n_cls[2] = {0, 0};
    for (int i = 0; i < NB; i++) {
        Board<G> bd;
        PhiloxRng r0; r0.init(key ^ 0x9999, i, 0xFFFFFFFFu);
        init_board<SH>(bd, 6, r0);
        recs.clear();
    }
    recs.clear();
    std::vector<Board<G>> boards(NB);
    for (int i = 0; i < NB; i++) { PhiloxRng r0; r0.init(key ^ 0x9999, i, 0xFFFFFFFFu); init_board<SH>(boards[i], 6, r0); }
    recs.clear();
    for (int s = 0; s < STEPS; s++)
        for (int i = 0; i < NB; i++) {
            Board<G> &bd = boards[i];
            Derived<G> d = derive<SH>(bd);
            BB<W> HL, VL;
            legal_swaps<SH>(d, eq_at<SH, 1>(d), eq_at<SH, S>(d), HL, VL);
            const int c = swaps_count<G>(HL, VL);
            if (!c) continue;
            uint32_t blk[4];
            const uint32_t k = philox_pick(key, i, s, c, blk);
            const int a = swaps_select<G>(HL, VL, (int)k);
            g_cls = any(bd.p[3] & bd.p[2]) ? 1 : 0;
            n_cls[g_cls]++;
            PhiloxRng rng; rng.init(key, i, s);
            StepOut so;
            size_t before = recs.size();
            bool h = step_board_two_pass<SH>(bd, a, 6, rng, so, HL, VL);
            if (h) { handoffs++; recs.resize(before); } // keep only common-case iterations
            steps++;
        }
    printf("steps %ld handoffs %ld cascade-iterations %zu (%.3f per step)\n", steps, handoffs, recs.size(), recs.size() / (double)steps);
    auto stat = [&](const char *name, auto f) {
        double mean = 0; int hist[12] = {0};
        for (auto &r : recs) { int v = f(r); mean += v; hist[std::min(v, 11)]++; }
        mean /= recs.size();
        printf("%-12s mean/lane %.3f  E[max of 8/16/32]", name, mean);
        for (int lanes = 8; lanes <= 32; lanes *= 2) {
            std::mt19937 g(1); double wm = 0; const int T = 200000;
            for (int t = 0; t < T; t++) { int m = 0; for (int l = 0; l < lanes; l++) m = std::max(m, f(recs[g() % recs.size()])); wm += m; }
            printf(" %.3f", wm / T);
        }
        printf("  hist 0..11+:");
        for (int i = 0; i < 12; i++) printf(" %.3f", hist[i] / (double)recs.size());
        printf("\n");
    };
    printf("steps starting with a special on the board: %.4f\n", n_cls[1] / (double)steps);
    std::vector<Rec> all = recs;
    for (int c = -1; c < 2; c++) {
    if (c >= 0) { recs.clear(); for (auto &r : all) if (r.cls == c) recs.push_back(r); printf("-- class %d (%s): %.4f of the cascade iterations\n", c, c ? "special at the start of the step" : "no special at the start of the step", recs.size() / (double)all.size()); }
    stat("gravity", [](const Rec &r) { return r.g; });
    stat("after v3", [](const Rec &r) { return r.g3; });
    stat("after v3+v2", [](const Rec &r) { return r.g32; });
    stat("holes", [](const Rec &r) { return r.holes; });
    }
    return 0;
}
