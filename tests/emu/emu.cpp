// Host build of sq_recovery_b200/csrc/sq_core.cuh -- TEST TOOL, never part of the product.
//
// The per-point core is written as host+device functions so that the forward/backward algebra and the
// finalize Jacobians can be checked against the oracle on a machine without a GPU.  This file replays the
// kernels' loop structure (one "thread" per column, fp32 per-column sums, fp64 across columns) serially.
// MUFU approximations are replaced by libm (ex2f/log2f/1/x), so this checks the maths, not the MUFU error.
#include "sq_core.cuh"
#include <vector>
#include <cstring>

using namespace sq;

static void acc_to_double(const Acc& a, double* d) {
    for (int i = 0; i < 3; ++i) d[i] += a.gs[i];
    for (int i = 0; i < 9; ++i) d[3 + i] += a.gm[i];
    for (int i = 0; i < 3; ++i) d[12 + i] += a.wa[i];
    d[15] += a.ge[0]; d[16] += a.ge[1]; d[17] += a.loss;
}

static int g_emu_refine = 1, g_emu_queue = 1, g_emu_pool = kBwdPool;

extern "C" {

void emu_set_refine(int on) { g_emu_refine = on; }     // fp64 refinement of the pooled points near the surface
void emu_set_queue(int on) { g_emu_queue = on; }       // 0: every gradient point on the spot (two-moment path)
void emu_set_pool(int slots) { g_emu_pool = slots > 0 && slots <= kBwdPool ? slots : kBwdPool; }   // small: exercises the overflow path

// the pool entries' tags (sq_core.cuh BwdQueue): plane index | lane | link packed into one float and taken apart again.
// Returns 0 when plane, lane and link all survive the round trip (cf0 = the sample's "index" of plane 0, below 1).
int emu_tag_roundtrip(float cf, float cf0, int lane, int link) {
    Sample S{}; S.cf0 = cf0;
    const float tagged = bits_f32((f32_bits(cf) & ~kTagMask) | (unsigned)lane | ((unsigned)link << 5));
    return (entry_cf(S, tagged) != cf) | ((entry_lane(tagged) != lane) << 1) | ((entry_link(tagged) != link) << 2);
}

// accuracy probes of the fp64 primitives
double emu_exp2_acc(double y) { return exp2_acc(y, default_tabs()); }
double emu_log2_acc(double m) { return log2_acc(m, default_tabs()); }

// target: [B, n, n] in image orientation (row, col); depth_out optional [B, n, n] image orientation
int emu_implicit(const double* pred, int B, int n, double step, double z0, const float* target, float tau, float k,
                 double* loss_out, double* grad /*[B,12] or null*/, float* depth_out /*or null*/) {
    Grid g = make_grid(n, step, z0);
    ImplicitParams P{k * kLog2e, tau * kLog2e, implicit_cull_bound(k * kLog2e), implicit_active_bits(k * kLog2e, n, B)};
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        double p[12]; for (int i = 0; i < 12; ++i) p[i] = pred[12 * b + i];
        SampleFull S; prep_sample(p, true, g, S);
        double accd[kAccN] = {0};
        // The kernels' compacted backward, one group of 32 columns (raster order) at a time: the columns walk one after the
        // other here, all appending to the group's pool of gradient-carrying points (front: the entries near the surface,
        // refined in fp64; back: the others), then refinement, corrected suffix weights and the backward per pooled entry.
        for (int g0 = 0; g0 < n * n; g0 += 32) {
            float qcf[kBwdPool], qpre[kBwdPool], qx[kBwdPool], qd[kBwdPool];
            BwdQueue q{qcf, qpre, qx, qd, (n <= kPoolMaxPlanes && S.cf0 < 1.0f) ? g_emu_pool : 0, 0, 0u};
            int nr = 0, top = q.cap - 1;
            float bh[32][3], bl[32][3], U[32], w[32], dx[32], dy[32];
            int head[32];
            Acc a[32];
            const int cols = n * n - g0 < 32 ? n * n - g0 : 32;
            for (int L = 0; L < cols; ++L) {
                const int ib = (g0 + L) / n, ia = (g0 + L) - ib * n;
                float cg[11];
                column_base(S, g, ia, ib, bh[L], bl[L]);
                int c_lo, c_hi;
                column_range(S, g, P.bound, bh[L], c_lo, c_hi);
                warp_range(n, c_lo, c_hi);
                U[L] = 0.f; head[L] = kNoLink; bool spilled = false;
                q.lane = L;
                const int own_lo = c_hi >= c_lo ? c_lo : n;
                const float depth = !grad ? implicit_column<false>(S, g, P, bh[L], bl[L], c_lo, c_hi, own_lo, cg)
                                  : g_emu_queue ? implicit_column<true, true, true>(S, g, P, bh[L], bl[L], c_lo, c_hi, own_lo, cg, &q, &U[L], &nr, &top, &spilled, &head[L])
                                                : implicit_column<true>(S, g, P, bh[L], bl[L], c_lo, c_hi, own_lo, cg);
                const int row = n - 1 - ib, col = ia;
                if (depth_out) depth_out[(size_t)b * n * n + row * n + col] = depth;
                const float tgt = target ? target[(size_t)b * n * n + row * n + col] : 0.f;
                const float diff = depth - tgt;
                acc_zero(a[L]);
                a[L].loss = fabsf(diff);
                w[L] = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                dx[L] = (float)(grid_coord(g, ia) - S.t[0]); dy[L] = (float)(grid_coord(g, ib) - S.t[1]);
                if (grad && (!g_emu_queue || spilled)) implicit_fold(a[L], cg, w[L], dx[L], dy[L]);
            }
            if (grad && g_emu_queue) {
                if (!g_emu_refine) {                               // front entries keep x: turn it into the weight, no correction
                    for (int e = 0; e < nr; ++e) queue_store_refined(q, e, qx[e], qx[e]);
                } else {
                    for (int e = 0; e < nr; ++e) {
                        const int L = entry_lane(qcf[e]);
                        queue_refine_entry(S, g.step, P.kl, q, e, (double)bh[L][0] + (double)bl[L][0], (double)bh[L][1] + (double)bl[L][1],
                                           (double)bh[L][2] + (double)bl[L][2], default_tabs());
                    }
                }
                for (int L = 0; L < cols; ++L) a[L].loss += w[L] * tau * g.inv_n * queue_depth_shift(q, head[L], U[L]);
                const int np = q.cap - 1 - top;
                for (int j = 0; j < nr + np; ++j) {
                    const int at = j < nr ? j : q.cap - 1 - (j - nr), L = entry_lane(qcf[at]);
                    const float cf = entry_cf(S, qcf[at]);
                    Bwd bq;
                    queue_entry_backward<true>(S, bh[L], bl[L], cf, qx[at], queue_suffix_weight(q, at, head[L], U[L], tau), w[L], true, bq);
                    acc_add_point(a[L], bq, cf, dx[L], dy[L]);
                }
            }
            for (int L = 0; L < cols; ++L) acc_to_double(a[L], accd);
        }
        total += accd[17] / ((double)n * n);
        if (grad) {
            const double scale = -(double)k * tau / ((double)n * n * n * B);
            finalize_sample(S, g, accd, scale, true, grad + 12 * b);
        }
    }
    *loss_out = total / B;
    return 0;
}

int emu_explicit(const double* tru, const double* pred, int B, int n, double step, double z0, float k, float mult,
                 double* loss_out, double* grad) {
    Grid g = make_grid(n, step, z0);
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        double pt[12], pp[12];
        for (int i = 0; i < 12; ++i) { pt[i] = tru[12 * b + i]; pp[i] = pred[12 * b + i]; }
        SampleFull St, Sp; prep_sample(pt, true, g, St); prep_sample(pp, true, g, Sp);
        double accd[kAccN] = {0};
        for (int ib = 0; ib < n; ++ib) for (int ia = 0; ia < n; ++ia) {
            float bht[3], blt[3], bhp[3], blp[3];
            column_base(St, g, ia, ib, bht, blt);
            column_base(Sp, g, ia, ib, bhp, blp);
            Acc a; acc_zero(a);
            const float dx = (float)(grid_coord(g, ia) - Sp.t[0]), dy = (float)(grid_coord(g, ib) - Sp.t[1]);
            const float bound = cull_bound_bits(k * kLog2e, 24.0f);
            Range rt, rp;
            column_range(St, g, bound, bht, rt.lo, rt.hi); warp_range(n, rt.lo, rt.hi);
            column_range(Sp, g, bound, bhp, rp.lo, rp.hi); warp_range(n, rp.lo, rp.hi);
            a.loss = grad ? explicit_column<true>(St, Sp, g, k * kLog2e, bht, blt, bhp, blp, rt, rp, dx, dy, a)
                          : explicit_column<false>(St, Sp, g, k * kLog2e, bht, blt, bhp, blp, rt, rp, dx, dy, a);
            acc_to_double(a, accd);
        }
        const double n3 = (double)n * n * n;
        total += mult * accd[17] / n3;
        if (grad) finalize_sample(Sp, g, accd, 2.0 * k * mult / (n3 * B), true, grad + 12 * b);
    }
    *loss_out = total / B;
    return 0;
}

int emu_iou(const double* tru, const double* pred, int B, int n, double step, long long* inter, long long* uni) {
    Grid g = make_grid(n, step, 0.0);
    for (int b = 0; b < B; ++b) {
        double pt[12], pp[12];
        for (int i = 0; i < 12; ++i) { pt[i] = tru[12 * b + i]; pp[i] = pred[12 * b + i]; }
        SampleFull St, Sp; prep_sample(pt, false, g, St); prep_sample(pp, false, g, Sp);
        long long I = 0, U = 0;
        for (int ib = 0; ib < n; ++ib) for (int ia = 0; ia < n; ++ia) {
            float bht[3], blt[3], bhp[3], blp[3];
            column_base(St, g, ia, ib, bht, blt);
            column_base(Sp, g, ia, ib, bhp, blp);
            unsigned i = 0, u = 0;
            Range rt, rp;
            column_range(St, g, kIoUBound, bht, rt.lo, rt.hi); warp_range(n, rt.lo, rt.hi);
            column_range(Sp, g, kIoUBound, bhp, rp.lo, rp.hi); warp_range(n, rp.lo, rp.hi);
            iou_column(St, Sp, &St, &Sp, g, ia, ib, bht, blt, bhp, blp, rt, rp, i, u);
            I += i; U += u;
        }
        inter[b] = I; uni[b] = U;
    }
    return 0;
}

// points: [sum m, 3] (x, y, z) with offsets[B+1]
int emu_lsq(const double* pred, int B, const float* points, const int* offsets, double* loss_out, double* grad) {
    Grid g = make_grid(2, 1.0, 0.0);
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        double p[12]; for (int i = 0; i < 12; ++i) p[i] = pred[12 * b + i];
        SampleFull S; prep_sample(p, true, g, S);
        double accd[kAccN] = {0};
        for (int j = offsets[b]; j < offsets[b + 1]; ++j) {
            Acc a; acc_zero(a);
            a.loss = grad ? lsq_point<true>(S, points[3 * j], points[3 * j + 1], points[3 * j + 2], a)
                          : lsq_point<false>(S, points[3 * j], points[3 * j + 1], points[3 * j + 2], a);
            acc_to_double(a, accd);
        }
        const double vol = S.a[0] * S.a[1] * S.a[2];
        total += vol * accd[17];
        if (grad) {
            finalize_sample(S, g, accd, 2.0 * vol / B, false, grad + 12 * b);
            for (int i = 0; i < 3; ++i) grad[12 * b + i] += S.mask[i] * (vol / S.a[i]) * accd[17] / B;
        }
    }
    *loss_out = total / B;
    return 0;
}

}  // extern "C"

extern "C" {

// Culling claims of the column kernels, checked against the fp64 inside-outside function (TEST TOOL):
//  * for every column, the planes OUTSIDE column_range() (computed from the fp32 base, as the kernels do) have F >= Fmin;
//  * for every 8 x 4 patch the plan kernel would call proven empty (footprint_planes == 0), ALL planes of all its columns do.
// Returns the number of violations; *patches_empty / *patches_total report how much the proof catches.
long long emu_check_culling(const double* params, int B, int n, double step, double z0, int clamp, float bound,
                            double Fmin, long long* patches_empty, long long* patches_total) {
    Grid g = make_grid(n, step, z0);
    long long bad = 0, pe = 0, ptot = 0;
    for (int b = 0; b < B; ++b) {
        double p[12]; for (int i = 0; i < 12; ++i) p[i] = params[12 * b + i];
        SampleFull S; prep_sample(p, clamp != 0, g, S);
        std::vector<char> inside((size_t)n * n * n);
        for (int ia = 0; ia < n; ++ia) for (int ib = 0; ib < n; ++ib) for (int ic = 0; ic < n; ++ic) {
            const double gx = grid_coord(g, ia), gy = grid_coord(g, ib), gz = grid_coord(g, ic);
            double s[3];
            for (int i = 0; i < 3; ++i)
                s[i] = (S.M[3 * i] * (gx - S.t[0]) + S.M[3 * i + 1] * (gy - S.t[1]) + S.M[3 * i + 2] * (gz - S.t[2])) / S.a[i];
            const double A = pow(s[0] * s[0], 1.0 / S.e[1]), Bv = pow(s[1] * s[1], 1.0 / S.e[1]), C = pow(s[2] * s[2], 1.0 / S.e[0]);
            const double F = pow(pow(A + Bv, S.e[1] / S.e[0]) + C, S.e[0]);
            inside[((size_t)ia * n + ib) * n + ic] = F < Fmin;
        }
        for (int ia = 0; ia < n; ++ia) for (int ib = 0; ib < n; ++ib) {
            float b32[3]; column_base_f32(S, g, ia, ib, b32);
            int c_lo, c_hi; column_range(S, g, bound, b32, c_lo, c_hi);
            for (int ic = 0; ic < n; ++ic)
                if (inside[((size_t)ia * n + ib) * n + ic] && (ic < c_lo || ic > c_hi)) ++bad;
        }
        if (n % 8 != 0)      // x-fastest layout: 32-slot groups that wrap over several rows when n < 32
            for (int group = 0; group * 32 < n * n; ++group) {
                ++ptot;
                float cx, cy, hx, hy;
                xfast_group_footprint(n, group, cx, cy, hx, hy);
                if (footprint_planes(S, g, bound, cx, cy, hx, hy) != 0) continue;
                ++pe;
                for (int slot = group * 32; slot < group * 32 + 32 && slot < n * n; ++slot) {
                    const int ib = slot / n, ia = slot - ib * n;
                    for (int ic = 0; ic < n; ++ic) if (inside[((size_t)ia * n + ib) * n + ic]) ++bad;
                }
            }
        if (n % 8 == 0)
            for (int pa = 0; pa < n / 8; ++pa) for (int pb = 0; pb < n / 4; ++pb) {
                ++ptot;
                if (footprint_planes(S, g, bound, 8 * pa + 3.5f, 4 * pb + 1.5f, 3.5f, 1.5f) != 0) continue;
                ++pe;
                for (int ia = 8 * pa; ia < 8 * pa + 8; ++ia) for (int ib = 4 * pb; ib < 4 * pb + 4; ++ib)
                    for (int ic = 0; ic < n; ++ic) if (inside[((size_t)ia * n + ib) * n + ic]) ++bad;
            }
    }
    *patches_empty = pe; *patches_total = ptot;
    return bad;
}

// The plan kernel's walk estimate (footprint_walk) only orders the work, but an estimate of 0 planes means "proven empty" and
// the item is never processed: it must be 0 exactly where the plane count is, and never exceed it.  Returns the violations;
// *cut_groups counts the groups whose estimate the early exit shortened.
long long emu_check_walk_estimate(const double* params, int B, int n, double step, double z0, float bound, float die,
                                  long long* cut_groups, long long* live_groups) {
    Grid g = make_grid(n, step, z0);
    long long bad = 0, cut = 0, live = 0;
    for (int b = 0; b < B; ++b) {
        double p[12]; for (int i = 0; i < 12; ++i) p[i] = params[12 * b + i];
        SampleFull S; prep_sample(p, true, g, S);
        const int groups = n % 8 == 0 ? (n / 8) * (n / 4) : (n * n + 31) / 32;
        for (int group = 0; group < groups; ++group) {
            float cx, cy, hx, hy;
            if (n % 8 == 0) { const int pw = n >> 3, pb = group / pw, pa = group - pb * pw; cx = 8 * pa + 3.5f; cy = 4 * pb + 1.5f; hx = 3.5f; hy = 1.5f; }
            else xfast_group_footprint(n, group, cx, cy, hx, hy);
            const int planes = footprint_planes(S, g, bound, cx, cy, hx, hy);
            const int walk = footprint_walk(S, g, bound, die, cx, cy, hx, hy);
            if ((planes == 0) != (walk == 0) || walk > planes || walk < 0) ++bad;
            if (planes > 0) { ++live; if (walk < planes) ++cut; }
        }
    }
    *cut_groups = cut; *live_groups = live;
    return bad;
}

}  // extern "C"
