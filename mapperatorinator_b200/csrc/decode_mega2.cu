// Dataflow token-loop megakernel: the same 98 dependent micro-phases per token as decode_mega.cu, the same device arithmetic
// (gemv_dot, the attention unit's score / softmax / PV order, the fused logits chain), but NO grid barrier between them.
//
// Why: ncu + the clock64 timeline of round 1 showed the barrier kernel at 348 us / token is latency-bound, and 27 % of a token is the
// barrier itself (98 x ~0.95 us = store drain + release atomic + acquire poll) with another ~0.3 us per phase for the activation load
// that can only start after it.  Here every value that crosses CTAs is an 8-byte pair {fp32 bits | tag << 32} written by ONE 64-bit
// store and polled by whoever needs it (NCCL's "LL" protocol): the consumer's load IS the wait, the producer never drains or
// releases, and a CTA that has nothing to consume in a phase simply runs ahead.  tag = (step + 1) * 128 + phase + 1 is unique per
// (token, phase), buffers are zeroed before every launch.
//
// Which CTA reads what, and why a buffer can be overwritten without a second copy: a version of a buffer is only overwritten by a
// phase whose inputs transitively require EVERY reader of that version to have produced its own output first (e.g. x after out_proj
// is read by the 128 CTAs that own rows of the cross-q projection; the next writer of x, the cross out_proj, needs the merged cross
// attention, which needs every row of cross-q).  The full table is in DESIGN.md §4.1.
//
// The K/V rows appended to the cache are the one thing readers pick up long after the fact (next token onwards) through plain
// ld.cg: their writer fences right after the plain stores (hidden: that warp then waits for the attention phase anyway) and the
// NEWEST row travels as tagged pairs (`kvnew`) like every other same-token value.
//
// Every wait is bounded; on a timeout the error flag is raised, every other wait sees it and the launch drains.
#include "common.cuh"
#include "kernels.h"
#include "decode_device.cuh"

namespace mb200 {

namespace {

constexpr int M2_THREADS = SAMPLE_THREADS;              // 512
constexpr int M2_WARPS = M2_THREADS / 32;
constexpr int M2_NB_MAX = 2;
constexpr int M2_XS_FLOATS = M2_NB_MAX * 3072;
constexpr int M2_XRAW_FLOATS = M2_NB_MAX * 1024;
constexpr int M2_PART = 66;                             // pairs per split partial: o[64], m, l
constexpr int M3_SLOTS = 16;                            // K-split GEMV: output rows per thread (row slots), two passes of 8
constexpr int M3_ROWS = 32;                             // K-split GEMV: output rows per CTA and phase

struct __align__(16) M2Smem {
    float wbuf[2][MEGA_WBUF_FLOATS];
    union __align__(16) {
        float xs[M2_XS_FLOATS];
        SampleSmem sample;
        struct { __align__(16) float sc[128]; float red[4][64]; __align__(16) float qs[64]; float kns[64]; float vns[64]; float stat[2]; } attn;
    } u;
    __align__(16) float xraw[M2_XRAW_FLOATS];   // raw residual stream as of this CTA's last LayerNorm staging (residual source of its rows)
    Mega2Phase phase[3];                        // descriptor of phase i lives in slot i % 3 (the K-split mode has no end-of-phase barrier)
    SampleParams sample_params;
    int ctrl[8];                                // cur_len, all_finished, error, prompt_len, encoder slots
    float ln_red[32];
    // K-split GEMV mode (MODE 1): warp partial sums [parity][decoder row][row slot][warp] and the second LayerNorm statistic
    __align__(16) float red[2][M2_NB_MAX][M3_ROWS][M2_WARPS];     // [phase parity][decoder row][output row of this CTA][warp of the row's group]
    float ln_red2[32];
    unsigned long long mbar[2];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    unsigned long long pol;      // weights stream through L2 once per token: evict-first keeps the small hot set (exchange buffers, K/V) resident
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void cta_rows(int N, int cta, int rpc, int& r0, int& r1) {
    r0 = min(N, cta * rpc);
    r1 = min(N, r0 + rpc);
}
__device__ __forceinline__ void prefetch_weights(const float* W, long long ldw, int N, int K, float* dst, unsigned long long* bar, int cta, int rpc) {
    int r0, r1;
    cta_rows(N, cta, rpc, r0, r1);
    const unsigned bytes = (unsigned)(r1 - r0) * (unsigned)K * 4u;
    if (bytes == 0 || (c_ll_debug & 1)) { mbar_arrive(bar); return; }
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(dst, W + (long long)r0 * ldw, bytes, bar);
}
__device__ __forceinline__ bool wait_weights(unsigned long long* bar, unsigned parity, int* error_flag) {
    for (long long spin = 0; spin < (1ll << 24); ++spin)
        if (mbar_try_wait(bar, parity)) return true;
    atomicCAS(error_flag, 0, 2);
    return false;
}

__device__ __forceinline__ unsigned ll_tag(int step, int pi) { return (unsigned)((step + 1) * 128 + pi + 1); }

__device__ __forceinline__ ll_t* ll_buf(const MegaLL& ll, int sel) {
    switch (sel) {
        case LL_X: return ll.x;
        case LL_Q: return ll.q;
        case LL_K: return ll.kvnew;
        case LL_V: return ll.kvnew;
        case LL_ATT: return ll.att;
        case LL_H: return ll.h;
        case LL_LOGITS: return ll.logits;
        default: return nullptr;
    }
}

// ---- activation staging ---------------------------------------------------------------------------------------------------------
// LayerNorm prologue over the tagged residual stream: the reduction structure of gemv_stage_x (chunk of 32 float4 per warp, 8 chunk
// partials per row reduced by a fixed tree), operands polled instead of loaded; the raw values are kept in `xraw` (this CTA adds them
// back as the residual of the rows it owns two phases later, so the residual needs no second trip through L2).
template <int NB>
__device__ __forceinline__ void m2_stage_ln(const GemvParams& p, const ll_t* ll_in, unsigned tag, float* xs, float* xraw, float* red, int tid, int* err) {
    const int lane = tid & 31, warp = tid >> 5;
    const int K = p.K, K4 = K >> 2;
    constexpr int RG = M2_WARPS / 8;                   // rows staged per round (16 warps: 2)
    const float inv = 1.0f / (float)K;
    for (int g0 = 0; g0 < NB; g0 += RG) {
        const int rl = warp >> 3, bb = g0 + rl, c0 = warp & 7;
        const bool row_ok = bb < NB && bb < p.B;
        const int idx = c0 * 32 + lane;
        const bool ok = row_ok && idx < K4;
        float4 v = make_float4(0, 0, 0, 0), lw = v, lb = v;
        if (ok) {
            lw = __ldg(reinterpret_cast<const float4*>(p.ln_w) + idx);
            lb = __ldg(reinterpret_cast<const float4*>(p.ln_b) + idx);
            v = ll_wait4(ll_in + (long long)bb * K + idx * 4, tag, err);
            reinterpret_cast<float4*>(xraw + bb * K)[idx] = v;
        }
        const float sc = warp_sum((v.x + v.y) + (v.z + v.w));
        if (lane == 0) red[rl * 8 + c0] = sc;
        __syncthreads();
        const float* r = red + rl * 8;
        const float mean = (((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))) * inv;
        float q = 0.f;
        if (ok) {
            const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
            q = (a * a + b * b) + (c * c + d * d);
        }
        q = warp_sum(q);
        if (lane == 0) red[RG * 8 + rl * 8 + c0] = q;
        __syncthreads();
        if (bb < NB && idx < K4) {
            const float* r2 = red + RG * 8 + rl * 8;
            const float rstd = rsqrtf((((r2[0] + r2[1]) + (r2[2] + r2[3])) + ((r2[4] + r2[5]) + (r2[6] + r2[7]))) * inv + p.eps);
            float4 o = make_float4(0, 0, 0, 0);
            if (row_ok) {
                o.x = (v.x - mean) * rstd * lw.x + lb.x; o.y = (v.y - mean) * rstd * lw.y + lb.y;
                o.z = (v.z - mean) * rstd * lw.z + lb.z; o.w = (v.w - mean) * rstd * lw.w + lb.w;
            }
            reinterpret_cast<float4*>(xs + bb * K)[idx] = o;
        }
    }
}

// chunk partials of warps that hold no chunk (c0 >= K4 / 32) must read as zero: the fixed tree always adds eight of them
__device__ __forceinline__ void m2_clear_ln_red(float* red, int tid) {
    if (tid < 32) red[tid] = 0.f;
}

// Every poll below puts ALL of a thread's loads in flight before it looks at the first tag (a wait per value would serialise one L2
// round trip per value: measured +60 us / token).
__device__ __forceinline__ bool ll_tag_ok4(ll_t a, ll_t b, ll_t c, ll_t d, unsigned tag) {
    return (unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag && (unsigned)(c >> 32) == tag && (unsigned)(d >> 32) == tag;
}
__device__ __forceinline__ float ll_val(ll_t a) { return __uint_as_float((unsigned)a); }

template <int NB>
__device__ __forceinline__ void m2_stage_plain(const GemvParams& p, const ll_t* ll_in, unsigned tag, float* xs, int tid, int* err) {
    const int K4 = p.K >> 2;
    constexpr int U = 3;                               // float4 per thread per batch (K = 3072: 1.5 per decoder row)
    for (int e0 = tid; e0 < NB * K4; e0 += U * M2_THREADS) {
        ll_t w[U][4];
        bool on[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * M2_THREADS, bb = e / K4;
            on[u] = e < NB * K4 && bb < p.B;
        }
        long long spin = 0;
        while (true) {
            bool ok = true;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (on[u]) {
                    const int e = e0 + u * M2_THREADS, bb = e / K4, c = e - bb * K4;
                    const ll_t* src = ll_in + (long long)bb * p.K + c * 4;
                    ll_load2(src, w[u][0], w[u][1]);
                    ll_load2(src + 2, w[u][2], w[u][3]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (on[u]) ok = ok && ll_tag_ok4(w[u][0], w[u][1], w[u][2], w[u][3], tag);
            if (ok || !ll_spin_check(spin, err)) break;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * M2_THREADS;
            if (e < NB * K4) reinterpret_cast<float4*>(xs)[e] = on[u] ? make_float4(ll_val(w[u][0]), ll_val(w[u][1]), ll_val(w[u][2]), ll_val(w[u][3])) : make_float4(0, 0, 0, 0);
        }
    }
}

// ---- attention unit -------------------------------------------------------------------------------------------------------------
// (split s, head h, row r) exactly like decode_attention_body<16>: the same score chains, the same per-warp softmax statistics, the same
// four PV accumulation chains — only the operand sources differ: q and the newest K/V row are polled from the exchange buffers, the
// result leaves as tagged pairs (merged heads when one split covers the context, else a split partial that the split-0 CTA merges).
__device__ __forceinline__ void m2_attention_unit(const DecAttnParams& p, const MegaLL& ll, bool is_self, int s, int h, int r, int slot, int L, int P,
                                                  unsigned in_tag, unsigned out_tag, float* sc, float (*red)[64], float* stat, float* qs, float* kns,
                                                  float* vns, int tid, AttnRegs<M2_WARPS>& R, int* err) {
    constexpr int NW = M2_WARPS;
    constexpr int SC_ITERS = AttnRegs<NW>::SC_ITERS, PV_PRE = AttnRegs<NW>::PV_PRE;
    const int lane = tid & 31, warp = tid >> 5;
    const int d = p.H * 64;
    const int k_begin = s * p.chunk, k_end = min(L, k_begin + p.chunk);
    const int nk = k_end - k_begin;
    const int newest = is_self ? (L - 1 - k_begin) : -1;           // index inside this split of the key appended by THIS token (self only)
    ll_t* part = ll.part + (((long long)r * p.H + h) * ll.max_splits + s) * M2_PART;
    ll_t* att = ll.att + (long long)r * d + h * 64;
    if (nk <= 0) {                                                 // empty split (uniform across the CTA)
        if (p.n_splits == 1) { if (tid < 64) for (int rep = 0; rep < ll.reps; ++rep) ll_store(att + rep * ll.x_rep + tid, 0.f, out_tag); return; }
        if (tid < 64) ll_store(part + tid, 0.f, out_tag);
        if (tid == 0) { ll_store(part + 64, -INFINITY, out_tag); ll_store(part + 65, 0.f, out_tag); }
        stat[0] = -INFINITY; stat[1] = 0.f;
        return;
    }
    const int tok = (int)p.tok_stride;
    const float* vb = p.vc + (long long)slot * p.row_stride + h * 64 + (long long)k_begin * tok;
    const int sub = lane & 7, kq = lane >> 3;
    // q and (self-attention) the key / value row this token appended are polled ONCE per CTA — warp 14 takes q, warp 15 the new K | V
    // row — into shared memory; everybody reads them from there after one CTA barrier.  (All 512 threads polling their own copy cost
    // 16x the L2 polling traffic and, with the prefetched cache rows live in registers, pushed the kernel into spills.)
    const bool has_new = newest >= 0 && newest < nk;
    if (warp == NW - 2) {
        const float2 v = ll_wait2(ll.q + (long long)r * d + h * 64 + lane * 2, in_tag, err);
        qs[lane * 2] = v.x; qs[lane * 2 + 1] = v.y;
    } else if (warp == NW - 1 && has_new) {
        const ll_t* kn = ll.kvnew + (long long)r * 2 * d + h * 64;
        ll_t w[4];
        long long spin = 0;
        while (true) {
            ll_load2(kn + lane * 2, w[0], w[1]);
            ll_load2(kn + d + lane * 2, w[2], w[3]);
            if (ll_tag_ok4(w[0], w[1], w[2], w[3], in_tag) || !ll_spin_check(spin, err)) break;
        }
        kns[lane * 2] = ll_val(w[0]); kns[lane * 2 + 1] = ll_val(w[1]);
        vns[lane * 2] = ll_val(w[2]); vns[lane * 2 + 1] = ll_val(w[3]);
    }
    __syncthreads();
    const float4 q0 = *reinterpret_cast<const float4*>(qs + sub * 8), q1 = *reinterpret_cast<const float4*>(qs + sub * 8 + 4);
#pragma unroll
    for (int it = 0; it < SC_ITERS; ++it) {
        const int kk = it * 4 * NW + warp * 4 + kq;
        const bool is_new = kk == newest;
        const float4 a = is_new ? *reinterpret_cast<const float4*>(kns + sub * 8) : R.ka[it];
        const float4 b = is_new ? *reinterpret_cast<const float4*>(kns + sub * 8 + 4) : R.kb4[it];
        float dd = q0.x * a.x;
        dd = fmaf(q0.y, a.y, dd); dd = fmaf(q0.z, a.z, dd); dd = fmaf(q0.w, a.w, dd);
        dd = fmaf(q1.x, b.x, dd); dd = fmaf(q1.y, b.y, dd); dd = fmaf(q1.z, b.z, dd); dd = fmaf(q1.w, b.w, dd);
        dd += __shfl_xor_sync(0xffffffffu, dd, 1);
        dd += __shfl_xor_sync(0xffffffffu, dd, 2);
        dd += __shfl_xor_sync(0xffffffffu, dd, 4);
        if (kk < nk && sub == 0) sc[kk] = (R.kvalid[it] || is_new) ? dd : -INFINITY;
    }
    __syncthreads();
    if (warp < 4) {
        float pv[4];
        float m = -INFINITY;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = lane + 32 * t;
            pv[t] = i < nk ? sc[i] : -INFINITY;
            m = fmaxf(m, pv[t]);
        }
        m = warp_max(m);
        float l = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = lane + 32 * t;
            pv[t] = (i < nk && pv[t] != -INFINITY) ? expf(pv[t] - m) : 0.f;
            if (i < nk) l += pv[t];
        }
        l = warp_sum(l);
        if (tid == 0) { stat[0] = m; stat[1] = l; }
        float2 o = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < PV_PRE; ++i) {
            const int kk = warp + 4 * i;
            const float pk = __shfl_sync(0xffffffffu, pv[i >> 3], kk & 31);
            const float2 vv = kk == newest ? *reinterpret_cast<const float2*>(vns + lane * 2) : R.vpre[i];
            if (kk < nk) {
                o.x = fmaf(pk, vv.x, o.x);
                o.y = fmaf(pk, vv.y, o.y);
            }
        }
        if (nk > 64) {
#pragma unroll
            for (int t = 2; t < 4; ++t) {
                float2 vv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int kk = 32 * t + warp + 4 * i;
                    vv[i] = make_float2(0.f, 0.f);
                    if (kk < nk) vv[i] = kk == newest ? *reinterpret_cast<const float2*>(vns + lane * 2) : ldcg2(vb + kk * tok + lane * 2);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int kk = 32 * t + warp + 4 * i;
                    const float pk = __shfl_sync(0xffffffffu, pv[t], kk & 31);
                    if (kk < nk) {
                        o.x = fmaf(pk, vv[i].x, o.x);
                        o.y = fmaf(pk, vv[i].y, o.y);
                    }
                }
            }
        }
        red[warp][lane * 2] = o.x;
        red[warp][lane * 2 + 1] = o.y;
    }
    __syncthreads();
    if (tid < 64) {
        const float v = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
        if (p.n_splits == 1) {
            // one split holds the whole context: the merge degenerates to o / l (w = exp(m - m) = 1: the same bits as the general path)
            const float num = fmaf(1.f, v, 0.f), den = fmaf(1.f, stat[1], 0.f);
            const float o = (stat[1] > 0.f && den > 0.f) ? num / den : 0.f;
            for (int rep = 0; rep < ll.reps; ++rep) ll_store(att + rep * ll.x_rep + tid, o, out_tag);
        } else {
            ll_store(part + tid, v, out_tag);
            if (tid == 0) { ll_store(part + 64, stat[0], out_tag); ll_store(part + 65, stat[1], out_tag); }
        }
    }
}

// merge of the S split partials of (row r, head h) by the CTA that computed split 0: splits visited in index order,
// out = sum_s w_s o_s / sum_s w_s l_s with w_s = exp(m_s - max m) — decode_attention_merge's arithmetic, operands polled
__device__ __forceinline__ void m2_attention_merge(const DecAttnParams& p, const MegaLL& ll, int h, int r, unsigned tag, float* msh /* >= 2 * 32 floats of shared memory */,
                                                   int tid, int* err) {
    // called by the WHOLE CTA (uniform).  Threads 64 .. 64+S-1 poll the (m, l) pair of one split each into shared memory, threads 0..63
    // poll their output dim of every split (S <= 8: one batch; beyond that: rolled), then 64 threads combine.
    const int S = p.n_splits, d = p.H * 64;
    const ll_t* base = ll.part + ((long long)r * p.H + h) * ll.max_splits * M2_PART;
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + (tid - 64);
        if (tid >= 64 && tid < 96 && s < S) {
            const float2 mv = ll_wait2(base + (long long)s * M2_PART + 64, tag, err);
            if (s < 32) { msh[s] = mv.x; msh[32 + s] = mv.y; }
        }
    }
    float ov[8];
    if (tid < 64 && S <= 8) {
        ll_t oo[8];
        long long spin = 0;
        while (true) {
            bool ok = true;
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < S) asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(oo[s]) : "l"(base + (long long)s * M2_PART + tid) : "memory");
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < S) ok = ok && (unsigned)(oo[s] >> 32) == tag;
            if (ok || !ll_spin_check(spin, err)) break;
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) ov[s] = s < S ? ll_val(oo[s]) : 0.f;
    }
    __syncthreads();
    if (tid >= 64) return;
    float num = 0.f, den = 0.f;
    float mmax = -INFINITY;
    if (S <= 8) {
#pragma unroll
        for (int s = 0; s < 8; ++s) if (s < S) mmax = fmaxf(mmax, msh[s]);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (s < S && msh[32 + s] > 0.f) {
                const float w = expf(msh[s] - mmax);
                num = fmaf(w, ov[s], num);
                den = fmaf(w, msh[32 + s], den);
            }
        }
    } else {
#pragma unroll 1
        for (int s = 0; s < S; ++s) mmax = fmaxf(mmax, msh[s]);
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
            const float o1 = ll_wait1(base + (long long)s * M2_PART + tid, tag, err);
            if (msh[32 + s] > 0.f) {
                const float w = expf(msh[s] - mmax);
                num = fmaf(w, o1, num);
                den = fmaf(w, msh[32 + s], den);
            }
        }
    }
    const float o = den > 0.f ? num / den : 0.f;
    for (int rep = 0; rep < ll.reps; ++rep) ll_store(ll.att + rep * ll.x_rep + (long long)r * d + h * 64 + tid, o, tag);
}


// ---- K-split GEMV phase (MODE 1) ---------------------------------------------------------------------------------------------------
// The row-per-warp form spends its time in shared memory: every output row re-reads the whole activation vector next to its weight
// row (2 x K x 4 bytes per row through a 128 B / clk port), only rows-per-CTA of the 16 warps have work (6 of 16 for the d x d
// projections), and the activation takes a detour through shared memory behind two or three CTA barriers.  Here the CTA's weight slab
// [R rows x K] is split along K instead: thread t owns float4 column(s) kq of EVERY row, keeps its 4 activation values in registers
// (polled straight from the exchange buffer: no staging pass), and the R partial sums per thread are reduced by a transposing warp
// butterfly (9 shuffles per 8 rows) plus one pass over <= 16 warp partials.  Summation order is fixed by the thread mapping.
//   K4 = K / 4 float4 columns.  "grouped" (K4 a multiple of 32, <= 256: the d_model-wide inputs): G = 512 / K4 row groups, group rg
//   takes rows rg, rg + G, ... and group 0 polls the activation for everybody (one CTA barrier).  Otherwise one group, thread t owns
//   columns t and t + 512 (K4 <= 1024) and polls exactly what it multiplies.
__device__ __forceinline__ float m3_dot4(const float4 w, const float4 x) {
    float t0 = w.x * x.x; t0 = fmaf(w.y, x.y, t0);
    float t1 = w.z * x.z; t1 = fmaf(w.w, x.w, t1);
    return t0 + t1;
}
__device__ __forceinline__ float m3_tree8(const float* r) { return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7])); }

// 8 values per lane -> lane L (L % 4 == 0) holds the warp-wide sum of value index ((L >> 4) & 1) * 4 + ((L >> 3) & 1) * 2 + ((L >> 2) & 1)
__device__ __forceinline__ float m3_reduce8(const float (&a)[8], int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    float b[4], c[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h16 ? a[i] : a[i + 4], keep = h16 ? a[i + 4] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h8 ? b[i] : b[i + 2], keep = h8 ? b[i + 2] : b[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = h4 ? c[0] : c[1], keep = h4 ? c[1] : c[0];
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Thread -> weight-column mapping of a K-wide input; K takes two values per model (d_model, ffn_dim): computed once per kernel.
struct M3Map { int K4, G, rg, kq0, NS, npw, wi; bool grouped, in_group, active_warp; };
__device__ __forceinline__ M3Map m3_make_map(int K, int tid) {
    M3Map m;
    m.K4 = K >> 2;
    m.grouped = (m.K4 & 31) == 0 && m.K4 <= 256;
    m.G = m.grouped ? M2_THREADS / m.K4 : 1;
    m.rg = m.grouped ? tid / m.K4 : 0;
    m.kq0 = tid - m.rg * m.K4;
    m.in_group = m.rg < m.G;
    m.NS = m.grouped ? 1 : (m.K4 + M2_THREADS - 1) / M2_THREADS;
    m.npw = m.grouped ? m.K4 >> 5 : min(M2_WARPS, (m.K4 + 31) >> 5);
    m.wi = m.grouped ? (tid >> 5) - m.rg * m.npw : (tid >> 5);
    m.active_warp = m.grouped ? m.in_group : (tid >> 5) < m.npw;
    return m;
}

// One pass of <= 8 row slots: all weight loads first (branch-free: out-of-range rows re-read row 0 and are discarded by a select),
// then the products, then the transposing butterfly.  Everything a phase does is a chain of dependent latencies, so no load may sit
// behind a branch that waits for the previous row's arithmetic.
template <int NB, int NS>
__device__ __forceinline__ void m3_pass(const float* wb, int K, int R, int G, int rg, int p0, const int (&kqc)[2], const float4 (&xv)[NB][2],
                                        float* red_b0, float* red_b1, int wi, int lane) {
    float acc[NB][8];
    float4 w[8][NS];
    bool valid[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int lr = (p0 + i) * G + rg;
        valid[i] = lr < R;
        const float4* wrow = reinterpret_cast<const float4*>(wb + (valid[i] ? lr : 0) * K);
#pragma unroll
        for (int s = 0; s < NS; ++s) w[i][s] = wrow[kqc[s]];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float a = m3_dot4(w[i][0], xv[b][0]);
            if (NS > 1) a += m3_dot4(w[i][NS - 1], xv[b][NS - 1]);
            acc[b][i] = valid[i] ? a : 0.f;
        }
    }
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const int lr = (p0 + ridx) * G + rg;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const float v = m3_reduce8(acc[b], lane);
        if ((lane & 3) == 0 && lr < R) (b == 0 ? red_b0 : red_b1)[lr * M2_WARPS + wi] = v;
    }
}

template <int NB, typename SM>
__device__ __forceinline__ void m3_gemv_phase(const Mega2Params& mp, const Mega2Phase& ph2, SM& sm, const M3Map& mapd, const M3Map& mapf, int cta, int rep_off,
                                              int tid, unsigned g_idx, int par, unsigned in_tag, unsigned out_tag, int cur_pos, int* err,
                                              unsigned long long* trace /* null or [8] */) {
    const MegaPhase& ph = ph2.base;
    const GemvParams& g = ph.g;
    const int lane = tid & 31, warp = tid >> 5;
    const int buf = g_idx & 1;
    const int d = mp.d_model;
    const int K = g.K;
    const bool isd = K == d;
    const int K4 = K >> 2;
    const bool grouped = isd ? mapd.grouped : mapf.grouped, in_group = isd ? mapd.in_group : mapf.in_group, active_warp = isd ? mapd.active_warp : mapf.active_warp;
    const int G = isd ? mapd.G : mapf.G, rg = isd ? mapd.rg : mapf.rg, kq0 = isd ? mapd.kq0 : mapf.kq0, NS = isd ? mapd.NS : mapf.NS,
              npw = isd ? mapd.npw : mapf.npw, wi = isd ? mapd.wi : mapf.wi;
    // Every thread is past the barrier of the previous GEMV phase, i.e. nobody reads the other weight buffer any more: request the next
    // GEMV's slice right away (a whole phase of lead).  Thread 416 has no columns in the grouped d_model mapping (K4 = 192).
    if (tid == M2_THREADS - 96) prefetch_weights(ph.nx_W, ph.nx_ldw, ph.nx_N, ph.nx_K, sm.wbuf[buf ^ 1], &sm.mbar[buf ^ 1], cta, ph.nx_rpc);
    int r0, r1;
    cta_rows(g.N, cta, ph.rpc, r0, r1);
    const int R = r1 - r0;
    const bool ln = g.xmode == X_LAYERNORM;

    // ---- epilogue operands, fetched NOW by the threads that will finish the rows (warps 14 / 15: thread 448 + e finishes decoder row e / R,
    //      output row e % R), so that nothing but the sum itself is left behind the barrier ----
    const int e = tid - (M2_THREADS - 64);
    bool epi = false;
    int eb = 0, elr = 0, act = 0, ll_nrep = 0;
    float bias_v = 0.f, res_v = 0.f, alpha = 1.f;
    ll_t* ll_out = nullptr;
    long long ll_rs = 0;
    float* plain = nullptr;
    if (e >= 0 && e < R * NB) {
        eb = (NB > 1 && e >= R) ? 1 : 0;
        elr = e - eb * R;
        epi = eb < g.B;
        if (epi) {
            const int n = r0 + elr;
            const int si = (int)(g.nseg > 1 && n >= g.seg[1].n_begin) + (int)(g.nseg > 2 && n >= g.seg[2].n_begin);
            const GemvSeg& sg = g.seg[si];
            const int osel = ph2.out_sel[si], col = n - sg.n_begin;
            if (g.bias) bias_v = __ldg(g.bias + n);
            if (ph2.res_xraw) res_v = sm.xraw[eb * d + n];
            act = sg.act; alpha = sg.alpha;
            if (osel == LL_X || osel == LL_H) {
                const long long width = osel == LL_H ? g.N : d;
                ll_rs = osel == LL_H ? mp.ll.h_rep : mp.ll.x_rep;
                ll_out = (osel == LL_H ? mp.ll.h : mp.ll.x) + (long long)eb * width + col;
                ll_nrep = mp.ll.reps;
            } else if (osel != LL_NONE) {
                const long long width = osel == LL_K || osel == LL_V ? 2 * d : (osel == LL_LOGITS ? mp.V : d);
                ll_out = ll_buf(mp.ll, osel) + (long long)eb * width + (osel == LL_V ? d : 0) + col;
                ll_nrep = 1;
            }
            if (ph2.plain_out[si]) plain = sg.out + (long long)eb * sg.out_bs + (long long)cur_pos * sg.pos_stride + col;
        }
    }

    // ---- input: polled straight into registers ----
    float4 xv[NB][2];
#pragma unroll
    for (int b = 0; b < NB; ++b) { xv[b][0] = make_float4(0, 0, 0, 0); xv[b][1] = make_float4(0, 0, 0, 0); }
    bool weights_waited = false;
    if (R > 0) {
        const ll_t* in = (ph2.in_sel == LL_H ? mp.ll.h + (long long)rep_off * mp.ll.h_rep
                                             : (ph2.in_sel == LL_ATT ? mp.ll.att : mp.ll.x) + (long long)rep_off * mp.ll.x_rep);
        if (ln || grouped) {
            // one group polls (tid < K4), everybody else picks the values up from shared memory after the barrier
            const bool poller = rg == 0 && kq0 < K4;
            const bool has_col = in_group && kq0 < K4;
            float4 lw = make_float4(0, 0, 0, 0), lb = lw;
            if (ln && has_col) {
                lw = __ldg(reinterpret_cast<const float4*>(g.ln_w) + kq0);
                lb = __ldg(reinterpret_cast<const float4*>(g.ln_b) + kq0);
            }
            float4 raw[NB];
            {
                ll_t w[NB][4];
#pragma unroll
                for (int b = 0; b < NB; ++b) { w[b][0] = w[b][1] = w[b][2] = w[b][3] = 0; }
                if (poller) {
                    long long spin = 0;
                    bool first = true;
                    while (true) {
#pragma unroll
                        for (int b = 0; b < NB; ++b)
                            if (b < g.B) { const ll_t* src = in + (long long)b * K + kq0 * 4; ll_load2(src, w[b][0], w[b][1]); ll_load2(src + 2, w[b][2], w[b][3]); }
                        if (first) { wait_weights(&sm.mbar[buf], (g_idx >> 1) & 1, err); first = false; }      // overlaps the first round trip
                        bool ok = true;
#pragma unroll
                        for (int b = 0; b < NB; ++b)
                            if (b < g.B) ok = ok && ll_tag_ok4(w[b][0], w[b][1], w[b][2], w[b][3], in_tag);
                        if (ok || !ll_spin_check(spin, err)) break;
                    }
                    weights_waited = true;
                    if (trace && tid == 0) { trace[1] = (unsigned long long)clock64(); trace[7] = (unsigned long long)spin; }
                }
#pragma unroll
                for (int b = 0; b < NB; ++b)
                    raw[b] = (poller && b < g.B) ? make_float4(ll_val(w[b][0]), ll_val(w[b][1]), ll_val(w[b][2]), ll_val(w[b][3])) : make_float4(0, 0, 0, 0);
            }
            if (ln) {
                // LayerNorm with the reduction structure of gemv_stage_x / m2_stage_ln (32-float4 chunks per warp, 8 chunk partials, fixed tree):
                // the normalised activations carry the same bits as in the other token-loop drivers.
                const int npl = (K4 + 31) >> 5;
                const float inv = 1.0f / (float)K;
                if (warp < npl) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const float sc = warp_sum((raw[b].x + raw[b].y) + (raw[b].z + raw[b].w));
                        if (lane == 0) sm.ln_red[b * 8 + warp] = sc;
                        if (poller) reinterpret_cast<float4*>(sm.xraw + b * K)[kq0] = raw[b];
                    }
                }
                __syncthreads();
                float mean[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b) mean[b] = m3_tree8(sm.ln_red + b * 8) * inv;
                if (warp < npl) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        float q = 0.f;
                        if (poller && b < g.B) {
                            const float a0 = raw[b].x - mean[b], a1 = raw[b].y - mean[b], a2 = raw[b].z - mean[b], a3 = raw[b].w - mean[b];
                            q = (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
                        }
                        q = warp_sum(q);
                        if (lane == 0) sm.ln_red2[b * 8 + warp] = q;
                    }
                }
                __syncthreads();
                if (has_col) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (b < g.B) {
                            const float rstd = rsqrtf(m3_tree8(sm.ln_red2 + b * 8) * inv + g.eps);
                            const float4 v = poller ? raw[b] : reinterpret_cast<const float4*>(sm.xraw + b * K)[kq0];
                            xv[b][0].x = (v.x - mean[b]) * rstd * lw.x + lb.x; xv[b][0].y = (v.y - mean[b]) * rstd * lw.y + lb.y;
                            xv[b][0].z = (v.z - mean[b]) * rstd * lw.z + lb.z; xv[b][0].w = (v.w - mean[b]) * rstd * lw.w + lb.w;
                        }
                    }
                }
            } else {
                if (poller) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) reinterpret_cast<float4*>(sm.u.xs + b * K)[kq0] = raw[b];
                }
                __syncthreads();
                if (has_col) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) xv[b][0] = poller ? raw[b] : reinterpret_cast<const float4*>(sm.u.xs + b * K)[kq0];
                }
            }
        } else {
            // one group: every thread polls exactly the columns it multiplies, all loads in flight before the first tag check
            ll_t w[NB][2][4];
            bool on[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    on[b][s] = b < g.B && s < NS && kq0 + s * M2_THREADS < K4;
                    w[b][s][0] = w[b][s][1] = w[b][s][2] = w[b][s][3] = 0;
                }
            long long spin = 0;
            bool first = true;
            while (true) {
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int s = 0; s < 2; ++s)
                        if (on[b][s]) {
                            const ll_t* src = in + (long long)b * K + (kq0 + s * M2_THREADS) * 4;
                            ll_load2(src, w[b][s][0], w[b][s][1]);
                            ll_load2(src + 2, w[b][s][2], w[b][s][3]);
                        }
                if (first) { wait_weights(&sm.mbar[buf], (g_idx >> 1) & 1, err); first = false; }
                bool ok = true;
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int s = 0; s < 2; ++s)
                        if (on[b][s]) ok = ok && ll_tag_ok4(w[b][s][0], w[b][s][1], w[b][s][2], w[b][s][3], in_tag);
                if (ok || !ll_spin_check(spin, err)) break;
            }
            weights_waited = true;
            if (trace && tid == 0) { trace[1] = (unsigned long long)clock64(); trace[7] = (unsigned long long)spin; }
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    if (on[b][s]) xv[b][s] = make_float4(ll_val(w[b][s][0]), ll_val(w[b][s][1]), ll_val(w[b][s][2]), ll_val(w[b][s][3]));
        }
    }
    if (trace && tid == 0) trace[2] = (unsigned long long)clock64();
    if (!weights_waited) wait_weights(&sm.mbar[buf], (g_idx >> 1) & 1, err);      // every thread observes the phase of the mbarrier it will wait on next time
    if (trace && tid == 0) trace[3] = (unsigned long long)clock64();

    // ---- multiply + reduce ----
    if (R > 0 && active_warp && !(c_ll_debug & 4)) {
        const int Pn = G == 1 ? R : (G == 2 ? (R + 1) >> 1 : (R + G - 1) / G);       // row slots per thread (host guarantees <= M3_SLOTS)
        const float* wb = sm.wbuf[buf];
        int kqc[2];
        kqc[0] = kq0 < K4 ? kq0 : 0;
        kqc[1] = kq0 + M2_THREADS < K4 ? kq0 + M2_THREADS : 0;
        float* red0 = &sm.red[par][0][0][0];
        float* red1 = &sm.red[par][NB - 1][0][0];
#pragma unroll 1
        for (int p0 = 0; p0 < Pn; p0 += 8) {
            if (NS == 1) m3_pass<NB, 1>(wb, K, R, G, rg, p0, kqc, xv, red0, red1, wi, lane);
            else         m3_pass<NB, 2>(wb, K, R, G, rg, p0, kqc, xv, red0, red1, wi, lane);
        }
    }
    if (trace && tid == 0) trace[4] = (unsigned long long)clock64();
    asm volatile("cp.async.wait_all;" ::: "memory");   // next phase's descriptor (issued at the top of the phase loop)
    __syncthreads();
    if (trace && tid == 0) trace[5] = (unsigned long long)clock64();

    // ---- epilogue: sum of the warp partials (fixed tree over 16 slots, unused ones read as zero), bias / activation / residual, tagged store ----
    if (epi) {
        const float4* r4 = reinterpret_cast<const float4*>(&sm.red[par][eb][elr][0]);
        float4 q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = r4[i];
        float t[16] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w};
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = i < npw ? t[i] : 0.f;
        float v = (((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]))) + (((t[8] + t[9]) + (t[10] + t[11])) + ((t[12] + t[13]) + (t[14] + t[15])));
        v += bias_v;
        v = apply_act(v, act) * alpha;
        v += res_v;
        for (int rep = 0; rep < ll_nrep; ++rep) ll_store(ll_out + rep * ll_rs, v, out_tag);
        if (plain) {
            *plain = v;
            __threadfence();                           // K/V cache rows are read by LATER tokens through plain loads (see the header comment)
        }
        if (trace && e == 0) trace[6] = (unsigned long long)clock64();
    }
}

// TRACE = true is a separate instantiation for tools/mega2_trace.py: CTAs 0, 1, 100 and 140 stamp clock64 at four points of every phase
// of token `trace_step` (phase start, input staged, weights landed, rows / units done); the production kernel carries no stamp code.
// MODE 0: GEMV phases stage the activation in shared memory and give every warp whole output rows (the barrier kernel's gemv_dot).
// MODE 1: K-split GEMV phases (m3_*): every thread polls ITS k-slice of the activation straight into registers, multiplies it with
//         its slice of every weight row the CTA owns, and the CTA reduces the partial sums (warp butterfly + 16 warp partials).
template <int NB, bool TRACE, int MODE>
__global__ void __launch_bounds__(M2_THREADS, 1) decode_megakernel_ll(Mega2Params mp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    M2Smem& sm = *reinterpret_cast<M2Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int d = mp.d_model;
    int* const err = mp.error_flag;

    if (tid == 0) {
        sm.sample_params = mp.sample;
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    m2_clear_ln_red(sm.ln_red, tid);
    m2_clear_ln_red(sm.ln_red2, tid);
    __syncthreads();
    unsigned int g_idx = 0;          // running index of GEMV phases (selects weight buffer + mbarrier parity)
    if (tid == 0) {
        const MegaPhase* f = &mp.phases[0].base;
        prefetch_weights(f->g.W, f->g.ldw, f->g.N, f->g.K, sm.wbuf[0], &sm.mbar[0], cta, (f->g.N + G - 1) / G);
    }
    // CTA 0 publishes the residual stream left by the prefill (plain memory, written by an earlier kernel) and the first token header
    // under the tag the first phase of step 0 expects: "last phase of step -1"
    if (cta == 0) {
        const unsigned t0 = ll_tag(-1, mp.n_phases - 1);
        for (int rep = 0; rep < mp.ll.reps; ++rep)
            for (int i = tid; i < mp.rows * d; i += M2_THREADS) ll_store(mp.ll.x + rep * mp.ll.x_rep + i, mp.x_in[i], t0);
        if (tid == 0) {
            ll_store(mp.ll.hdr + 0, __int_as_float(ld_state(&mp.st->cur_len)), t0);
            ll_store(mp.ll.hdr + 1, __int_as_float(ld_state(&mp.st->all_finished)), t0);
        }
    }
    {
        const int* src = reinterpret_cast<const int*>(&mp.phases[0]);
        int* dst = reinterpret_cast<int*>(&sm.phase[0]);
        for (int i = tid; i < (int)(sizeof(Mega2Phase) / 4); i += M2_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    int cur = 0;
    AttnRegs<M2_WARPS> areg;
    const M3Map mapd = m3_make_map(mp.d_model, tid), mapf = m3_make_map(mp.ffn_dim > 0 ? mp.ffn_dim : mp.d_model, tid);
    const int rep_off = cta % mp.ll.reps;

    for (int step = 0; step < mp.max_steps; ++step) {
        if (tid == 0) {      // token header: written by the selection phase of the previous token (or the prologue above)
            const unsigned th = ll_tag(step - 1, mp.n_phases - 1);
            sm.ctrl[0] = __float_as_int(ll_wait1(mp.ll.hdr + 0, th, err));
            sm.ctrl[1] = __float_as_int(ll_wait1(mp.ll.hdr + 1, th, err));
            sm.ctrl[2] = *reinterpret_cast<volatile int*>(err);
            sm.ctrl[3] = ld_state(&mp.st->prompt_len);
            sm.ctrl[4] = mp.row_slot[0]; sm.ctrl[5] = mp.rows > 1 ? mp.row_slot[1] : 0;
        }
        __syncthreads();
        const int cur_pos = sm.ctrl[0] - 1, fin = sm.ctrl[1], e = sm.ctrl[2], P = sm.ctrl[3];
        if (fin || e) break;
        const int tslot = cta == 0 ? 0 : (cta == 1 ? 1 : (cta == 100 ? 2 : (cta == 140 ? 3 : -1)));
        const bool tracing = TRACE && mp.trace != nullptr && step == mp.trace_step && tslot >= 0 && tid == 0;
#define M2_TRACE(slot) do { if (tracing) { if (MODE == 1) { if (tslot < 2) mp.trace[((long long)tslot * mp.n_phases + pi) * 8 + ((slot) == 3 ? 5 : (slot))] = (unsigned long long)clock64(); } \
                                           else mp.trace[((long long)tslot * mp.n_phases + pi) * 4 + (slot)] = (unsigned long long)clock64(); } } while (0)
        for (int pi = 0; pi < mp.n_phases; ++pi) {
            M2_TRACE(0);
            const Mega2Phase& ph2 = sm.phase[cur];
            const MegaPhase& ph = ph2.base;
            const unsigned in_tag = pi == 0 ? ll_tag(step - 1, mp.n_phases - 1) : ll_tag(step, pi - 1);
            const unsigned out_tag = ll_tag(step, pi);
            // next phase's descriptor: global -> shared asynchronously, drained before the end-of-phase CTA barrier
            constexpr int DESC_WORDS = (int)(sizeof(Mega2Phase) / 4);
            const int nxt = cur == 2 ? 0 : cur + 1;
            // (K-split mode: by warps 12..15, which hold no columns of the d_model-wide inputs, so the pollers start polling at once)
            for (int i = MODE == 1 ? tid - 384 : tid; i >= 0 && i < DESC_WORDS; i += MODE == 1 ? 128 : M2_THREADS) {
                const int* src = reinterpret_cast<const int*>(&mp.phases[pi + 1 < mp.n_phases ? pi + 1 : 0]) + i;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(reinterpret_cast<int*>(&sm.phase[nxt]) + i)), "l"(src) : "memory");
            }
            if (ph.kind == 0 && MODE == 1) {
                // finer timeline than MODE 0: CTAs 0 and 1 only, 8 stamps per phase (0 start, 1 polled, 2 input ready, 3 weights landed,
                // 4 partial sums written, 5 past the barrier, 6 epilogue of output row 0 stored, 7 = failed polls of thread 0)
                unsigned long long* tr = nullptr;
                if (TRACE && mp.trace != nullptr && step == mp.trace_step && cta < 2) {
                    tr = mp.trace + ((long long)cta * mp.n_phases + pi) * 8;
                    if (tid == 0) tr[0] = (unsigned long long)clock64();
                }
                m3_gemv_phase<NB>(mp, ph2, sm, mapd, mapf, cta, rep_off, tid, g_idx, pi & 1, in_tag, out_tag, cur_pos, err, tr);
                ++g_idx;
            } else if (ph.kind == 0) {
                const GemvParams& g = ph.g;
                const int buf = g_idx & 1;
                int r0, r1;
                cta_rows(g.N, cta, ph.rpc, r0, r1);
                // bias of this warp's rows: lane j*NB + b holds it for the warp's j-th row (fetched while the input is still on its way)
                float bias_pref = 0.f;
                {
                    const int n = r0 + warp + (lane / NB) * M2_WARPS;
                    if (n < r1 && g.bias) bias_pref = __ldg(g.bias + n);
                }
                if (r0 < r1) {
                    const ll_t* in = ll_buf(mp.ll, ph2.in_sel) + (cta % mp.ll.reps) * (ph2.in_sel == LL_H ? mp.ll.h_rep : mp.ll.x_rep);
                    if (g.xmode == X_LAYERNORM) m2_stage_ln<NB>(g, in, in_tag, sm.u.xs, sm.xraw, sm.ln_red, tid, err);
                    else m2_stage_plain<NB>(g, in, in_tag, sm.u.xs, tid, err);
                }
                __syncthreads();
                M2_TRACE(1);
                // the next GEMV's weight slice is requested only now (after this phase's small latency-critical loads), by the last warp
                if (tid == M2_THREADS - 32) prefetch_weights(ph.nx_W, ph.nx_ldw, ph.nx_N, ph.nx_K, sm.wbuf[buf ^ 1], &sm.mbar[buf ^ 1], cta, ph.nx_rpc);
                wait_weights(&sm.mbar[buf], (g_idx >> 1) & 1, err);
                M2_TRACE(2);
                int j = 0;
#pragma unroll 1
                for (int n = r0 + warp; n < r1; n += M2_WARPS, ++j) {
                    const float bias_v = __shfl_sync(0xffffffffu, bias_pref, (j * NB) & 31);
                    const int si = (int)(g.nseg > 1 && n >= g.seg[1].n_begin) + (int)(g.nseg > 2 && n >= g.seg[2].n_begin);
                    const GemvSeg& sg = g.seg[si];
                    const int osel = ph2.out_sel[si];
                    const float mine = gemv_dot<NB, false>(g.K, sm.wbuf[buf] + (long long)(n - r0) * g.K, sm.u.xs, lane);
                    float v = 0.f;
                    if (lane < NB && lane < g.B) {
                        v = mine;
                        if (g.bias) v += bias_v;
                        v = apply_act(v, sg.act) * sg.alpha;
                        if (ph2.res_xraw) v += sm.xraw[lane * d + n];
                    }
                    if (osel == LL_X || osel == LL_H) {
                        // replicated buffers: lane l stores replica l / NB of decoder row l % NB (the value comes from lane l % NB)
                        const int b = lane % NB, rep = lane / NB;
                        const float vb = __shfl_sync(0xffffffffu, v, b);
                        if (rep < mp.ll.reps && b < g.B) {
                            const long long width = osel == LL_H ? g.N : d;
                            ll_store(ll_buf(mp.ll, osel) + rep * (osel == LL_H ? mp.ll.h_rep : mp.ll.x_rep) + (long long)b * width + (n - sg.n_begin), vb, out_tag);
                        }
                    } else if (lane < NB && lane < g.B) {
                        if (osel != LL_NONE) {         // the tagged copy first: it is what this token's next phase waits for
                            ll_t* out = ll_buf(mp.ll, osel);
                            const long long width = osel == LL_K || osel == LL_V ? 2 * d : (osel == LL_LOGITS ? mp.V : d);
                            const int col = osel == LL_V ? d + (n - sg.n_begin) : (n - sg.n_begin);
                            ll_store(out + (long long)lane * width + col, v, out_tag);
                        }
                        if (ph2.plain_out[si]) {
                            sg.out[(long long)lane * sg.out_bs + (long long)cur_pos * sg.pos_stride + (n - sg.n_begin)] = v;
                            // K/V cache rows are read by LATER tokens through plain loads: order them before this thread's next tagged
                            // store (the one of the following phase; this token's were issued above)
                            __threadfence();
                        }
                    }
                }
                ++g_idx;
            } else if (ph.kind == 1) {
                const DecAttnParams& a = ph.a;
                const bool is_self = a.fixed_len == 0;
                const int L = is_self ? cur_pos + 1 : a.fixed_len;
                const int units = a.rows * a.H * a.n_splits;
                for (int u = cta; u < units; u += G) {
                    const int hr = ph.magic_ns ? (int)__umulhi((unsigned)u, ph.magic_ns) : u;
                    const int s = u - hr * a.n_splits;
                    const int r = ph.magic_h ? (int)__umulhi((unsigned)hr, ph.magic_h) : hr;
                    const int h = hr - r * a.H;
                    const int slot = a.row_slot ? sm.ctrl[4 + r] : r;
                    // K/V of the cache first (they do not depend on this token's phases), q and the appended row are polled inside
                    decode_attention_load<M2_WARPS>(a, s, h, r, slot, L, P, tid, areg);
                    m2_attention_unit(a, mp.ll, is_self, s, h, r, slot, L, P, in_tag, out_tag, sm.u.attn.sc, sm.u.attn.red, sm.u.attn.stat, sm.u.attn.qs,
                                      sm.u.attn.kns, sm.u.attn.vns, tid, areg, err);
                    if (a.n_splits > 1 && s == 0) { __syncthreads(); m2_attention_merge(a, mp.ll, h, r, out_tag, sm.u.attn.sc, tid, err); }
                    __syncthreads();
                }
            } else {
                if (cta < sm.sample_params.cfg->B) {
                    if (tid == 0) { sm.sample_params.ll_in_tag = in_tag; sm.sample_params.ll_out_tag = out_tag; }
                    __syncthreads();
                    sample_body(sm.sample_params, cta, sm.u.sample);
                }
            }
            if (!(ph.kind == 0 && MODE == 1)) {           // (the K-split GEMV phase ends with its own barrier, before its epilogue)
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncthreads();                          // xs / attention scratch free for the next phase; the next descriptor has landed
                M2_TRACE(3);
            }
            cur = nxt;
        }
    }
#undef M2_TRACE
    // drain the weight prefetch that is still in flight so no bulk copy outlives the CTA
    wait_weights(&sm.mbar[g_idx & 1], (g_idx >> 1) & 1, err);
}

}  // namespace

size_t mega2_smem_bytes() { return sizeof(M2Smem) + 128; }

int mega2_set_poll_sleep(int ns) {
    MB_CUDA_CHECK(cudaMemcpyToSymbol(c_ll_sleep_ns, &ns, sizeof(int)));
    return 0;
}
int mega2_set_debug(int bits) {
    MB_CUDA_CHECK(cudaMemcpyToSymbol(c_ll_debug, &bits, sizeof(int)));
    return 0;
}

template <int NB, bool TRACE, int MODE>
static const void* m2_configure() {
    const void* fn = (const void*)decode_megakernel_ll<NB, TRACE, MODE>;
    static bool done = false;
    if (!done) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega2_smem_bytes()) != cudaSuccess) return nullptr;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, M2_THREADS, mega2_smem_bytes()) != cudaSuccess || per_sm < 1) return nullptr;
        done = true;
    }
    return fn;
}

int launch_megakernel2(const Mega2Params& mp, int grid, cudaStream_t stream) {
    MB_REQUIRE(mp.rows >= 1 && mp.rows <= M2_NB_MAX, "megakernel handles 1 or 2 decoder rows");
    MB_REQUIRE(mp.n_phases <= 126, "tag layout holds at most 126 phases per token");
    MB_REQUIRE(mp.d_model <= 1024, "residual scratch holds d_model <= 1024");
    const bool tr = mp.trace != nullptr, one = mp.rows == 1;
    const void* fn = nullptr;
    if (mp.gemv_mode == 1) {
        fn = tr ? (one ? m2_configure<1, true, 1>() : m2_configure<2, true, 1>()) : (one ? m2_configure<1, false, 1>() : m2_configure<2, false, 1>());
    } else {
        fn = tr ? (one ? m2_configure<1, true, 0>() : m2_configure<2, true, 0>()) : (one ? m2_configure<1, false, 0>() : m2_configure<2, false, 0>());
    }
    MB_REQUIRE(fn != nullptr, "dataflow megakernel does not fit on an SM");
    Mega2Params p = mp;
    void* args[] = {&p};
    // Cooperative launch for its co-residency guarantee: every CTA polls values that other CTAs produce, so all of them must be resident.
    MB_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(M2_THREADS), args, mega2_smem_bytes(), stream));
    ++g_launch_count;
    return 0;
}

// Limits of the K-split GEMV mapping for one phase (checked on the host when the phase table is built)
bool mega2_ksplit_ok(int N, int K, int rows, int grid) {
    const int K4 = K >> 2, rpc = (N + grid - 1) / grid;
    if ((K & 3) || K4 > 2 * M2_THREADS) return false;
    const bool grouped = (K4 & 31) == 0 && K4 <= 256;
    const int G = grouped ? M2_THREADS / K4 : 1;
    return (rpc + G - 1) / G <= M3_SLOTS && rpc <= M3_ROWS && rpc * rows <= 64;
}

}  // namespace mb200
