// Slider end-point recompute of the diffusion `denoised_fn` on the device.
//   reference: DiffisionPipeline.sample_part.denoised_fn (diffusion_pipeline.py:203-222) -> SliderPath(curve_type, control_points)
//   .get_distance() / .position_at(length / max_length)  (osuT5/osuT5/inference/slider_path.py:26-230) over the flattening routines of
//   osuT5/osuT5/inference/path_approximator.py (adaptive bezier subdivision :12-81,173-222; circular arc :100-161; Catmull :84-97).
// Without this every slider-bearing map (i.e. every real map) had to leave the fused 100-step loop once per step for a host round trip.
//
// One thread per slider: a slider has 2-10 control points and flattens to a few hundred vertices — latency, not throughput, and a few
// hundred sliders per chunk run in parallel.  The polyline is never stored: the path is generated twice through the same
// deterministic walker, pass 1 sums the length (get_distance), pass 2 stops at distance min(length, max_length) and interpolates on
// the segment that contains it (binary search + interpolate_vertices collapse into "first vertex whose cumulative length >= d").
// Arithmetic: bezier / Catmull in float64 (the reference's work buffers are float64 np.empty arrays); the circular arc in float32
// without FMA contraction, operation by operation as numpy evaluates it on float32 control points — near-collinear anchors give radii
// of 1e4+ px where float32 and float64 differ by pixels, and parity is against what the reference computes.
#include "common.cuh"
#include "kernels.h"

namespace mb200 {
namespace {

constexpr int SL_MAX_CP = 64;        // control points per slider
constexpr int SL_MAX_SPAN = 32;      // control points per bezier sub-path (its degree + 1)
constexpr int SL_STACK = 24;         // subdivision stack depth (one pending right half per level)
constexpr double SL_BEZIER_TOL = 0.25;
constexpr int SL_CATMULL_DETAIL = 50;
constexpr float SL_ARC_TOL = 0.1f;

struct P2 { double x, y; };

struct Walker {                      // consumes the vertices of the calculated path in order
    bool locate;                     // false: measure; true: find the point at distance `target`
    double target;
    double cum;
    bool have_prev, found;
    P2 prev, result;
    __device__ void emit(P2 p) {
        if (have_prev && p.x == prev.x && p.y == prev.y) return;                      // slider_path.py:133-139
        if (have_prev) {
            const double dx = p.x - prev.x, dy = p.y - prev.y;
            const double seg = sqrt(dx * dx + dy * dy);
            if (locate && !found && cum + seg >= target) {
                const double d0 = cum, d1 = cum + seg;
                if (fabs(d1 - d0) <= 1e-8 + 1e-5 * fabs(d1)) result = prev;            // np.isclose(d0, d1)
                else { const double w = (target - d0) / (d1 - d0); result = P2{prev.x + dx * w, prev.y + dy * w}; }
                found = true;
            }
            cum += seg;
        } else if (locate && target <= 0.0) {
            result = p; found = true;
        }
        prev = p; have_prev = true;
    }
};

__device__ void bezier_subdivide(const P2* cp, int count, P2* left, P2* right, P2* mid) {
    for (int i = 0; i < count; ++i) mid[i] = cp[i];
    for (int i = 0; i < count; ++i) {
        left[i] = mid[0];
        right[count - i - 1] = mid[count - i - 1];
        for (int j = 0; j < count - i - 1; ++j) mid[j] = P2{(mid[j].x + mid[j + 1].x) / 2, (mid[j].y + mid[j + 1].y) / 2};
    }
}

__device__ bool bezier_flat_enough(const P2* cp, int count) {
    for (int i = 1; i < count - 1; ++i) {
        const double px = cp[i - 1].x - 2 * cp[i].x + cp[i + 1].x, py = cp[i - 1].y - 2 * cp[i].y + cp[i + 1].y;
        if (px * px + py * py > SL_BEZIER_TOL * SL_BEZIER_TOL * 4) return false;
    }
    return true;
}

// approximate_bezier (= approximate_b_spline with p = 0 -> full degree), depth-first so vertices come out in curve order
__device__ bool flatten_bezier(const P2* cp, int count, Walker& w, P2* stack, P2* cur, P2* left, P2* right, P2* mid) {
    if (count <= 0) return true;
    int sp = 0;
    for (int i = 0; i < count; ++i) stack[i] = cp[i];
    sp = 1;
    while (sp > 0) {
        --sp;
        for (int i = 0; i < count; ++i) cur[i] = stack[sp * SL_MAX_SPAN + i];
        if (bezier_flat_enough(cur, count)) {
            // bezier_approximate (:200-222): subdivide once more, emit cp[0] and the smoothed interior points
            bezier_subdivide(cur, count, left, right, mid);
            w.emit(cur[0]);
            for (int i = 1; i < count - 1; ++i) {
                const int k = 2 * i;                               // index into left ++ right[1:]
                auto both = [&](int q) -> P2 { return q < count ? left[q] : right[q - count + 1]; };
                const P2 a = both(k - 1), b = both(k), c = both(k + 1);
                w.emit(P2{0.25 * (a.x + 2 * b.x + c.x), 0.25 * (a.y + 2 * b.y + c.y)});
            }
            continue;
        }
        if (sp + 2 > SL_STACK) return false;
        bezier_subdivide(cur, count, left, right, mid);
        for (int i = 0; i < count; ++i) { stack[sp * SL_MAX_SPAN + i] = right[i]; stack[(sp + 1) * SL_MAX_SPAN + i] = left[i]; }
        sp += 2;
    }
    w.emit(cp[count - 1]);
    return true;
}

__device__ void flatten_catmull(const P2* cp, int count, Walker& w) {
    for (int i = 0; i < count - 1; ++i) {
        const P2 v1 = i > 0 ? cp[i - 1] : cp[i];
        const P2 v2 = cp[i];
        const P2 v3 = i < count - 1 ? cp[i + 1] : P2{v2.x + v2.x - v1.x, v2.y + v2.y - v1.y};
        const P2 v4 = i < count - 2 ? cp[i + 2] : P2{v3.x + v3.x - v2.x, v3.y + v3.y - v2.y};
        for (int c = 0; c < SL_CATMULL_DETAIL; ++c) {
            for (int e = 0; e < 2; ++e) {
                const double t = (double)(c + e) / SL_CATMULL_DETAIL, t2 = t * t, t3 = t * t2;
                auto f = [&](double a, double b, double cc, double d) {
                    return 0.5 * (2 * b + (-a + cc) * t + (2 * a - 5 * b + 4 * cc - d) * t2 + (-a + 3 * b - 3 * cc + d) * t3);
                };
                w.emit(P2{f(v1.x, v2.x, v3.x, v4.x), f(v1.y, v2.y, v3.y, v4.y)});
            }
        }
    }
}

// approximate_circular_arc in float32, no contraction (each numpy operation rounds once); false = degenerate -> caller uses bezier
__device__ bool flatten_arc(const P2* cp, Walker& w) {
    const float ax = (float)cp[0].x, ay = (float)cp[0].y, bx = (float)cp[1].x, by = (float)cp[1].y, cx = (float)cp[2].x, cy = (float)cp[2].y;
    auto dot2 = [](float x0, float y0, float x1, float y1) { return __fadd_rn(__fmul_rn(x0, x1), __fmul_rn(y0, y1)); };
    const float a_sq = dot2(bx - cx, by - cy, bx - cx, by - cy), b_sq = dot2(ax - cx, ay - cy, ax - cx, ay - cy), c_sq = dot2(ax - bx, ay - by, ax - bx, ay - by);
    if (fabsf(a_sq) <= 1e-8f || fabsf(b_sq) <= 1e-8f || fabsf(c_sq) <= 1e-8f) return false;
    const float s = __fmul_rn(a_sq, __fsub_rn(__fadd_rn(b_sq, c_sq), a_sq));
    const float t = __fmul_rn(b_sq, __fsub_rn(__fadd_rn(a_sq, c_sq), b_sq));
    const float u = __fmul_rn(c_sq, __fsub_rn(__fadd_rn(a_sq, b_sq), c_sq));
    const float sum = __fadd_rn(__fadd_rn(s, t), u);
    if (fabsf(sum) <= 1e-8f) return false;
    const float ox = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(s, ax), __fmul_rn(t, bx)), __fmul_rn(u, cx)), sum);
    const float oy = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(s, ay), __fmul_rn(t, by)), __fmul_rn(u, cy)), sum);
    const float dax = ax - ox, day = ay - oy, dcx = cx - ox, dcy = cy - oy;
    const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(dax, dax), __fmul_rn(day, day)));
    const float theta_start = atan2f(day, dax);
    float theta_end = atan2f(dcy, dcx);
    const float two_pi = 6.283185307179586f;
    while (theta_end < theta_start) theta_end = __fadd_rn(theta_end, two_pi);
    float direction = 1.f;
    float theta_range = __fsub_rn(theta_end, theta_start);
    const float ortx = cy - ay, orty = -(cx - ax);
    if (dot2(ortx, orty, bx - ax, by - ay) < 0.f) { direction = -1.f; theta_range = __fsub_rn(two_pi, theta_range); }
    int n;
    if (__fmul_rn(2.f, r) <= SL_ARC_TOL) n = 2;
    else n = max(2, (int)ceilf(__fdiv_rn(theta_range, __fmul_rn(2.f, acosf(__fsub_rn(1.f, __fdiv_rn(SL_ARC_TOL, r)))))));
    for (int i = 0; i < n; ++i) {
        const float fract = (float)(direction * ((double)i / (double)(n - 1)));
        const float theta = __fadd_rn(theta_start, __fmul_rn(fract, theta_range));
        w.emit(P2{(double)__fadd_rn(ox, __fmul_rn(cosf(theta), r)), (double)__fadd_rn(oy, __fmul_rn(sinf(theta), r))});
    }
    return true;
}

// SliderPath.calculate_path: sub-paths end where two consecutive control points coincide (or at the last point)
__device__ bool walk_path(int type, const P2* cps, int n, Walker& w, P2* stack, P2* cur, P2* left, P2* right, P2* mid) {
    int start = 0;
    for (int i = 0; i < n; ++i) {
        if (i == n - 1 || (cps[i].x == cps[i + 1].x && cps[i].y == cps[i + 1].y)) {
            const P2* span = cps + start;
            const int count = i + 1 - start;
            if (count > SL_MAX_SPAN) return false;
            if (type == 3) {
                for (int k = 0; k < count; ++k) w.emit(span[k]);
            } else if (type == 2) {
                flatten_catmull(span, count, w);
            } else {
                bool done = false;
                if (type == 1 && n == 3 && count == 3) done = flatten_arc(span, w);
                if (!done && !flatten_bezier(span, count, w, stack, cur, left, right, mid)) return false;
            }
            start = i + 1;
        }
    }
    return true;
}

__global__ void slider_end_kernel(SliderSet sl, float* __restrict__ pix, int T, int* __restrict__ error_flag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= sl.n) return;
    P2 cps[SL_MAX_CP];
    P2 stack[SL_STACK * SL_MAX_SPAN], cur[SL_MAX_SPAN], left[SL_MAX_SPAN], right[SL_MAX_SPAN], mid[SL_MAX_SPAN];
    const int c0 = sl.cp_offsets[k], n = sl.cp_offsets[k + 1] - c0;
    if (n > SL_MAX_CP || n <= 0) { if (n > SL_MAX_CP) atomicExch(error_flag, 1); return; }
    for (int i = 0; i < n; ++i) {
        const int t = sl.cp_index[c0 + i];
        cps[i] = P2{(double)pix[t], (double)pix[T + t]};
    }
    Walker w{};
    w.locate = false;
    if (!walk_path(sl.type[k], cps, n, w, stack, cur, left, right, mid)) { atomicExch(error_flag, 2); return; }
    const double max_length = w.cum;
    if (!w.have_prev || max_length == 0.0) return;                                        // diffusion_pipeline.py:215-216
    const double progress = fmin(fmax((double)sl.length[k] / max_length, 0.0), 1.0);
    Walker w2{};
    w2.locate = true; w2.target = progress * max_length;
    walk_path(sl.type[k], cps, n, w2, stack, cur, left, right, mid);
    const P2 e = w2.found ? w2.result : w2.prev;
    pix[sl.end_index[k]] = (float)e.x;
    pix[T + sl.end_index[k]] = (float)e.y;
}

// to_positions of the conditional half (diffusion_pipeline.py:172-177): pix[ch][t] = ((x[0][ch][t] + 1) / 2) * size_ch
__global__ void to_pixels_kernel(const float* __restrict__ x, int T, float* __restrict__ pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * T) return;
    const int ch = i / T;
    pix[i] = __fmul_rn(__fdiv_rn(__fadd_rn(x[i], 1.f), 2.f), ch == 0 ? 512.f : 384.f);
}

// x[:, :, :] = pix / (512, 384) * 2 - 1 for BOTH halves (diffusion_pipeline.py:220 broadcasts the conditional positions)
__global__ void from_pixels_kernel(const float* __restrict__ pix, int N, int T, float* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 2 * T) return;
    const int rem = i % (2 * T), ch = rem / T;
    x[i] = __fsub_rn(__fmul_rn(__fdiv_rn(pix[rem], ch == 0 ? 512.f : 384.f), 2.f), 1.f);
}

}  // namespace

int launch_slider_recompute(const SliderSet& sl, float* x, int N, int T, float* pix_scratch, int* error_flag, cudaStream_t st) {
    if (sl.n <= 0) return 0;
    to_pixels_kernel<<<(2 * T + 255) / 256, 256, 0, st>>>(x, T, pix_scratch);
    MB_LAUNCH_CHECK();
    slider_end_kernel<<<(sl.n + 63) / 64, 64, 0, st>>>(sl, pix_scratch, T, error_flag);
    MB_LAUNCH_CHECK();
    from_pixels_kernel<<<(N * 2 * T + 255) / 256, 256, 0, st>>>(pix_scratch, N, T, x);
    MB_LAUNCH_CHECK();
    g_launch_count += 3;
    return 0;
}

}  // namespace mb200
