// Tensor-core flash attention (head_dim 64) for the dense phases: Whisper encoder self-attention (T = 512, HF modeling_whisper.py:286-358),
// the DiT blocks' +-128 band (osu_diffusion/utils/models.py:145-151) and causal / key-padded self-attention.  fp32 in, fp32 out, both
// contractions on tcgen05.mma kind::tf32 with the 3xTF32 split of gemm_tc.cu (x = hi + lo, hi.hi + hi.lo + lo.hi), accumulators in TMEM.
//
//   attn_prep_kernel   one pass over q | k | v (token-major, the projection GEMM's layout): scale + tf32 hi / lo split, written head-major
//                      as Q [2][B*H][Tq_pad][64], K [2][B*H][Tk_pad][64] and V TRANSPOSED VT [2][B*H][64][Tk_pad] (so that V is a K-major
//                      B operand of the second contraction), zero padded to whole tiles.
//   attention_tc_kernel  one CTA = 128 queries of one (batch, head), 64-key tiles, 192 threads:
//       warp 0 / lane 0 : TMA producer — Q once (4 x [128 x 32 floats]), then per KV tile 4 K tiles + 4 VT tiles (2-stage ring)
//       warp 1 / lane 0 : MMA issuer   — S = Q K^T (24 x tcgen05.mma 128x64x8, SS) into one of two 64-column TMEM buffers, issued one tile
//                         AHEAD; O_t = P V (24 x tcgen05.mma, TS: P is read from TMEM) into a third buffer
//       warps 2..5      : one thread per query row (its TMEM lane): tcgen05.ld S -> mask -> online softmax in registers -> P hi / lo back to
//                         TMEM (tcgen05.st) -> after the PV MMA: O <- O * alpha + O_t in registers -> normalise -> 256 contiguous bytes per row
//   Row max, row sum and the running output never leave the thread that owns the row: no shuffles, no shared-memory softmax.
// Every wait is bounded (error flag + fall through), like gemm_tc.cu.  Dense boolean masks and K/V gathered through kv_slot stay on the
// fp32 SIMT kernel (attention.cu).
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "kernels.h"

namespace mb200 {

namespace {

constexpr int FA_BM = 128, FA_BN = 64, FA_HD = 64, FA_THREADS = 320, FA_STAGES = 2;      // TMA warp + MMA warp + 8 softmax warps
constexpr int FA_Q_TILE = FA_BM * 32 * 4;                 // 16 KB: 128 rows x 32 floats
constexpr int FA_Q_BYTES = 4 * FA_Q_TILE;                 // hi d0-31 | hi d32-63 | lo d0-31 | lo d32-63
constexpr int FA_KV_TILE = FA_BN * 32 * 4;                // 8 KB: 64 rows x 32 floats
constexpr int FA_STAGE_BYTES = 8 * FA_KV_TILE;            // K: hi d0, hi d1, lo d0, lo d1 | VT: hi k0, hi k1, lo k0, lo k1
constexpr unsigned FA_TM_S0 = 0, FA_TM_S1 = 64, FA_TM_PHI = 128, FA_TM_PLO = 192, FA_TM_OT = 256, FA_TM_COLS = 512;

struct FaBarriers {
    unsigned long long q_full;
    unsigned long long kv_full[FA_STAGES], kv_empty[FA_STAGES];
    unsigned long long s_full[2];
    unsigned long long p_full;        // 256 softmax threads arrive
    unsigned long long o_full;
    unsigned int tmem_base;
    int pad;
    float xchg[2][2][FA_BM];          // [tile parity][column half][row]: row maxima of the two threads that share a query row
};

__device__ __forceinline__ unsigned fa_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool fa_wait(unsigned long long* bar, unsigned parity, int* err) {
    for (long long spin = 0; spin < (1ll << 22); ++spin) {
        unsigned ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(fa_s32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicExch(err, 5);
    return false;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (see gemm_tc.cu)
__device__ __forceinline__ unsigned long long fa_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)(1024 >> 4) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)2 << 61;
    return d;
}
__device__ __forceinline__ void fa_mma_ss(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fa_mma_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fa_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"l"(__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fa_tmem_ld32(unsigned (&v)[32], unsigned taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void fa_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fa_tmem_st32(unsigned taddr, const unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ float fa_rn_tf32(float x) {
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

struct FaParams {
    float* o; long long o_ld, o_bs;
    int B, H, Tq, Tk, Tq_pad, Tk_pad;
    int mask_mode, q_pos0, band;
    const unsigned char* key_valid; long long key_valid_ld;
};

__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_vt,
                    FaParams p, int* err) {
    extern __shared__ unsigned char fa_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* q_s = smem;
    unsigned char* kv_s = smem + FA_Q_BYTES;
    FaBarriers* bars = reinterpret_cast<FaBarriers*>(smem + FA_Q_BYTES + FA_STAGES * FA_STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * FA_BM, bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;

    // KV tile range the mask allows for this query tile
    int kt_begin = 0, kt_end = (p.Tk + FA_BN - 1) / FA_BN;
    if (p.mask_mode == MASK_CAUSAL) {
        const int last = p.q_pos0 + min(q0 + FA_BM - 1, p.Tq - 1);
        kt_end = min(kt_end, last / FA_BN + 1);
    } else if (p.mask_mode == MASK_BAND) {
        const int lo = q0 - p.band + 1, hi = min(q0 + FA_BM - 1, p.Tq - 1) + p.band;
        kt_begin = max(0, lo) / FA_BN;
        kt_end = min(kt_end, hi / FA_BN + 1);
    }
    const int ntiles = max(0, kt_end - kt_begin);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fa_s32(&bars->q_full)));
        for (int s = 0; s < FA_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fa_s32(&bars->kv_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fa_s32(&bars->kv_empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fa_s32(&bars->s_full[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fa_s32(&bars->s_full[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(fa_s32(&bars->p_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fa_s32(&bars->o_full)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fa_s32(&bars->tmem_base)), "r"(FA_TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0 && ntiles > 0) {
            // Q: rows bh * Tq_pad + q0 .. +127, halves d0-31 / d32-63, hi (plane 0) then lo (plane 1)
            const unsigned qb = fa_s32(&bars->q_full);
            asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(qb), "r"(FA_Q_BYTES) : "memory");
            const int qrow = bh * p.Tq_pad + q0;
            for (int t = 0; t < 4; ++t)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(fa_s32(q_s + t * FA_Q_TILE)), "l"(&map_q), "r"((t & 1) * 32), "r"(qrow), "r"(t >> 1), "r"(qb) : "memory");
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % FA_STAGES;
                const unsigned ph = (i / FA_STAGES) & 1;
                if (!fa_wait(&bars->kv_empty[s], ph ^ 1, err)) break;
                unsigned char* st = kv_s + s * FA_STAGE_BYTES;
                const unsigned fb = fa_s32(&bars->kv_full[s]);
                asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(fb), "r"(FA_STAGE_BYTES) : "memory");
                const int k0 = (kt_begin + i) * FA_BN;
                const int krow = bh * p.Tk_pad + k0;
                for (int t = 0; t < 4; ++t)      // K tiles: [64 keys x 32 dims]
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(fa_s32(st + t * FA_KV_TILE)), "l"(&map_k), "r"((t & 1) * 32), "r"(krow), "r"(t >> 1), "r"(fb) : "memory");
                for (int t = 0; t < 4; ++t)      // V^T tiles: [64 dims x 32 keys]
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(fa_s32(st + (4 + t) * FA_KV_TILE)), "l"(&map_vt), "r"(k0 + (t & 1) * 32), "r"(bh * FA_HD), "r"(t >> 1), "r"(fb) : "memory");
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && ntiles > 0) {
            // instruction descriptor: D fp32, A / B tf32, both K-major, N = 64, M = 128
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(FA_BN >> 3) << 17) | ((unsigned)(FA_BM >> 4) << 24);
            bool ok = fa_wait(&bars->q_full, 0, err);
            const unsigned qa = fa_s32(q_s);
            auto issue_s = [&](int i) {          // S_i = Q K_i^T into TMEM buffer i & 1
                const int s = i % FA_STAGES;
                ok = ok && fa_wait(&bars->kv_full[s], (i / FA_STAGES) & 1, err);
                if (!ok) return;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned kb = fa_s32(kv_s + s * FA_STAGE_BYTES);
                const unsigned acc = tmem + ((i & 1) ? FA_TM_S1 : FA_TM_S0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {         // 8 k-steps of 8 dims: tile ks / 4, 32-byte sub-step ks % 4
                    const unsigned qo = (ks >> 2) * FA_Q_TILE + (ks & 3) * 32, ko = (ks >> 2) * FA_KV_TILE + (ks & 3) * 32;
                    const unsigned long long q_hi = fa_desc(qa + qo), q_lo = fa_desc(qa + 2 * FA_Q_TILE + qo);
                    const unsigned long long k_hi = fa_desc(kb + ko), k_lo = fa_desc(kb + 2 * FA_KV_TILE + ko);
                    fa_mma_ss(acc, q_hi, k_hi, idesc, ks > 0 ? 1u : 0u);
                    fa_mma_ss(acc, q_hi, k_lo, idesc, 1u);
                    fa_mma_ss(acc, q_lo, k_hi, idesc, 1u);
                }
                fa_commit(&bars->s_full[i & 1]);
            };
            issue_s(0);
            for (int i = 0; i < ntiles && ok; ++i) {
                if (i + 1 < ntiles) issue_s(i + 1);      // the next tile's scores are computed while the softmax warps work on this one
                ok = ok && fa_wait(&bars->p_full, i & 1, err);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int s = i % FA_STAGES;
                const unsigned vb = fa_s32(kv_s + s * FA_STAGE_BYTES + 4 * FA_KV_TILE);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {         // 8 k-steps of 8 keys: P columns ks * 8 .., V^T tile ks / 4
                    const unsigned vo = (ks >> 2) * FA_KV_TILE + (ks & 3) * 32;
                    const unsigned long long v_hi = fa_desc(vb + vo), v_lo = fa_desc(vb + 2 * FA_KV_TILE + vo);
                    const unsigned p_hi = tmem + FA_TM_PHI + ks * 8, p_lo = tmem + FA_TM_PLO + ks * 8;
                    fa_mma_ts(tmem + FA_TM_OT, p_hi, v_hi, idesc, ks > 0 ? 1u : 0u);
                    fa_mma_ts(tmem + FA_TM_OT, p_hi, v_lo, idesc, 1u);
                    fa_mma_ts(tmem + FA_TM_OT, p_lo, v_hi, idesc, 1u);
                }
                fa_commit(&bars->kv_empty[s]);           // K / V of this stage are free once these MMAs have read them
                fa_commit(&bars->o_full);
            }
        }
    } else {
        // ---- softmax / output warps: TWO threads per query row (warps w and w + 4 share a TMEM lane group), 32 score / output columns each.
        //      One warp per scheduler left every dependent instruction exposed (ncu: 20 k cycles per 64-key tile, tensor pipe 9 % active);
        //      two warps per scheduler and half the columns per thread cut the per-tile chain four-fold.  Scores arrive in the log2 domain
        //      (scale * log2 e folded into Q by the prep pass): p = ex2(s - m) is one MUFU instruction.
        const int lg = warp & 3;                          // TMEM lane group this warp may access
        const int half = (warp - 2) >> 2;                 // column half of this thread
        const int row = lg * 32 + lane, q = q0 + row;
        const unsigned lane_addr = (unsigned)(lg * 32) << 16;
        const unsigned col0 = (unsigned)(half * 32);
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        bool ok = true;
        const unsigned char* kvalid = p.key_valid ? p.key_valid + (long long)b * p.key_valid_ld : nullptr;
        // keys this row may see form one interval [k_lo, k_hi) (none / causal / band); key padding is applied on top when present
        int k_lo = 0, k_hi = p.Tk;
        if (p.mask_mode == MASK_CAUSAL) k_hi = min(k_hi, p.q_pos0 + q + 1);
        else if (p.mask_mode == MASK_BAND) { k_lo = max(k_lo, q - p.band + 1); k_hi = min(k_hi, q + p.band + 1); }
        if (q >= p.Tq) k_hi = k_lo;
        for (int i = 0; i < ntiles && ok; ++i) {
            const int k0 = (kt_begin + i) * FA_BN + (int)col0;       // key of this thread's first column
            ok = fa_wait(&bars->s_full[i & 1], (i >> 1) & 1, err);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned sv[32];
            fa_tmem_ld32(sv, tmem + lane_addr + ((i & 1) ? FA_TM_S1 : FA_TM_S0) + col0);
            fa_tmem_ld_wait();
            const int c_lo = k_lo - k0, c_hi = k_hi - k0;             // allowed columns of this thread: [c_lo, c_hi)
            float mx = -INFINITY;
            if (__all_sync(0xffffffffu, c_lo <= 0 && c_hi >= 32 && kvalid == nullptr)) {
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(sv[c]));
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    bool allowed = c >= c_lo && c < c_hi;
                    if (kvalid) allowed = allowed && kvalid[min(k0 + c, p.Tk - 1)] != 0;
                    const float sc = allowed ? __uint_as_float(sv[c]) : -INFINITY;
                    sv[c] = __float_as_uint(sc);
                    mx = fmaxf(mx, sc);
                }
            }
            // the row's maximum over both column halves
            bars->xchg[i & 1][half][row] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(mx, bars->xchg[i & 1][half ^ 1][row]);
            const float m_new = fmaxf(m_run, mx);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far: every s is -inf, ex2(-inf - 0) = 0
            float alpha;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"(m_run - m_use));     // m_run = -inf -> 0 (o and l are 0 anyway)
            float psum = 0.f;
            unsigned ph[32], pl[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float pr;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pr) : "f"(__uint_as_float(sv[c]) - m_use));
                psum += pr;
                const float hi = fa_rn_tf32(pr);
                ph[c] = __float_as_uint(hi);
                pl[c] = __float_as_uint(fa_rn_tf32(pr - hi));
            }
            l_run = l_run * alpha + psum;
            m_run = m_new;
            // (the previous tile's PV MMA has been waited for below, so the P buffers are free)
            fa_tmem_st32(tmem + lane_addr + FA_TM_PHI + col0, ph);
            fa_tmem_st32(tmem + lane_addr + FA_TM_PLO + col0, pl);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(fa_s32(&bars->p_full)) : "memory");
            // O <- O * alpha + P V   (this thread's 32 output dims)
            ok = fa_wait(&bars->o_full, i & 1, err);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned ov[32];
            fa_tmem_ld32(ov, tmem + lane_addr + FA_TM_OT + col0);
            fa_tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = fmaf(o[j], alpha, __uint_as_float(ov[j]));
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        // row sum over both halves, then normalise; fully masked rows (left-pad queries) produce 0 like torch SDPA
        bars->xchg[ntiles & 1][half][row] = l_run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float l_tot = l_run + bars->xchg[ntiles & 1][half ^ 1][row];
        if (q < p.Tq) {
            const float inv = (ok && l_tot > 0.f) ? 1.0f / l_tot : 0.f;
            float4* orow = reinterpret_cast<float4*>(p.o + (long long)b * p.o_bs + (long long)q * p.o_ld + h * FA_HD + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) orow[j] = make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(FA_TM_COLS) : "memory");
}

// q | k | v (token-major, strided) -> scaled tf32 hi / lo planes, head-major, V transposed; zero padded to whole tiles.
// grid (ceil(max(Tq_pad, Tk_pad) / 64), B * H), 256 threads: each CTA handles 64 tokens of one (batch, head).
struct PrepParams {
    const float* q; long long q_ld, q_bs;
    const float* k; long long k_ld, k_bs;
    const float* v; long long v_ld, v_bs;
    float* qw; float* kw; float* vtw;             // workspaces: [2][BH][Tq_pad][64], [2][BH][Tk_pad][64], [2][BH][64][Tk_pad]
    int B, H, Tq, Tk, Tq_pad, Tk_pad;
    float scale;
};
__global__ void __launch_bounds__(256) attn_prep_kernel(PrepParams p) {
    __shared__ float vt_hi[64][65], vt_lo[64][65];
    const int t0 = blockIdx.x * 64, bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
    const int tid = threadIdx.x, r = tid >> 4, d4 = (tid & 15) * 4;
    const long long BH = (long long)p.B * p.H;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int row = r + rr * 16, t = t0 + row;
        if (t < p.Tq_pad) {
            float4 x = make_float4(0, 0, 0, 0);
            if (t < p.Tq) x = *reinterpret_cast<const float4*>(p.q + (long long)b * p.q_bs + (long long)t * p.q_ld + h * 64 + d4);
            x.x *= p.scale; x.y *= p.scale; x.z *= p.scale; x.w *= p.scale;
            float4 hi = make_float4(fa_rn_tf32(x.x), fa_rn_tf32(x.y), fa_rn_tf32(x.z), fa_rn_tf32(x.w));
            float4 lo = make_float4(fa_rn_tf32(x.x - hi.x), fa_rn_tf32(x.y - hi.y), fa_rn_tf32(x.z - hi.z), fa_rn_tf32(x.w - hi.w));
            float* dst = p.qw + ((long long)bh * p.Tq_pad + t) * 64 + d4;
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<float4*>(dst + BH * p.Tq_pad * 64) = lo;
        }
        if (t < p.Tk_pad) {
            float4 x = make_float4(0, 0, 0, 0), y = x;
            if (t < p.Tk) {
                x = *reinterpret_cast<const float4*>(p.k + (long long)b * p.k_bs + (long long)t * p.k_ld + h * 64 + d4);
                y = *reinterpret_cast<const float4*>(p.v + (long long)b * p.v_bs + (long long)t * p.v_ld + h * 64 + d4);
            }
            float4 hi = make_float4(fa_rn_tf32(x.x), fa_rn_tf32(x.y), fa_rn_tf32(x.z), fa_rn_tf32(x.w));
            float4 lo = make_float4(fa_rn_tf32(x.x - hi.x), fa_rn_tf32(x.y - hi.y), fa_rn_tf32(x.z - hi.z), fa_rn_tf32(x.w - hi.w));
            float* dst = p.kw + ((long long)bh * p.Tk_pad + t) * 64 + d4;
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<float4*>(dst + BH * p.Tk_pad * 64) = lo;
            const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float vh = fa_rn_tf32(yv[c]);
                vt_hi[d4 + c][row] = vh;
                vt_lo[d4 + c][row] = fa_rn_tf32(yv[c] - vh);
            }
        }
    }
    __syncthreads();
    if (t0 < p.Tk_pad) {
        // transposed write: dim = r + 16 * rr, 4 consecutive keys per thread
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int dim = r + rr * 16, kq = (tid & 15) * 4;
            float* dst = p.vtw + ((long long)bh * 64 + dim) * p.Tk_pad + t0 + kq;
            *reinterpret_cast<float4*>(dst) = make_float4(vt_hi[dim][kq], vt_hi[dim][kq + 1], vt_hi[dim][kq + 2], vt_hi[dim][kq + 3]);
            *reinterpret_cast<float4*>(dst + BH * 64 * p.Tk_pad) = make_float4(vt_lo[dim][kq], vt_lo[dim][kq + 1], vt_lo[dim][kq + 2], vt_lo[dim][kq + 3]);
        }
    }
}

typedef CUresult (*FaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
FaEncodeTiledFn fa_encode_fn() {
    static FaEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<FaEncodeTiledFn>(p);
    }
    return fn;
}
// [2 planes][rows][cols] fp32, cols contiguous; box {32, box_rows, 1}; 128-byte swizzle
int fa_make_map(CUtensorMap* out, const float* base, long long cols, long long rows, int box_rows) {
    FaEncodeTiledFn fn = fa_encode_fn();
    MB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * rows * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

}  // namespace

int g_attn_tc_enabled = 1;
int g_attn_tc_min_t = 256;        // below this many queries the launch is a handful of CTAs either way: stay on the SIMT kernel

size_t attn_tc_workspace_bytes(int B, int H, int Tq, int Tk) {
    const size_t tq = (size_t)(Tq + FA_BM - 1) / FA_BM * FA_BM, tk = (size_t)(Tk + FA_BN - 1) / FA_BN * FA_BN;
    return (size_t)B * H * 64 * 4 * 2 * (tq + 2 * tk) + 1024;
}

int AttnCtx::reserve(size_t bytes) {
    if (bytes <= ws_bytes) return 0;
    MB_REQUIRE(!frozen, "attention workspace is frozen (a CUDA graph holds its address) and too small for this launch");
    if (ws) cudaFree(ws);
    ws = nullptr; ws_bytes = 0;
    MB_CUDA_CHECK(cudaMalloc(&ws, bytes));
    ws_bytes = bytes;
    if (!err) { MB_CUDA_CHECK(cudaMalloc(&err, sizeof(int))); MB_CUDA_CHECK(cudaMemset(err, 0, sizeof(int))); }
    return 0;
}
void AttnCtx::destroy() {
    if (ws) cudaFree(ws);
    if (err) cudaFree(err);
    ws = nullptr; err = nullptr; ws_bytes = 0;
}
int AttnCtx::error() {
    if (!err) return 0;
    int e = 0;
    if (cudaMemcpy(&e, err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return e;
}

bool attn_tc_eligible(const AttentionParams& p, const AttnCtx* ctx) {
    if (!ctx || !g_attn_tc_enabled) return false;
    if (p.mask_mode == MASK_DENSE || p.kv_slot != nullptr) return false;
    if (p.Tq < g_attn_tc_min_t || p.Tk < 1) return false;
    if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.k) & 15) || (reinterpret_cast<uintptr_t>(p.v) & 15) ||
        (reinterpret_cast<uintptr_t>(p.o) & 15))
        return false;
    if ((p.q_bs % 4) || (p.k_bs % 4) || (p.v_bs % 4) || (p.o_bs % 4)) return false;
    return true;
}

int launch_attention_tc(const AttentionParams& p, cudaStream_t stream, AttnCtx* ctx) {
    const int Tq_pad = (p.Tq + FA_BM - 1) / FA_BM * FA_BM, Tk_pad = (p.Tk + FA_BN - 1) / FA_BN * FA_BN;
    const long long BH = (long long)p.B * p.H;
    { const int rs = ctx->reserve(attn_tc_workspace_bytes(p.B, p.H, p.Tq, p.Tk)); if (rs) return rs; }
    float* qw = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ctx->ws) + 1023) & ~uintptr_t(1023));
    float* kw = qw + 2 * BH * Tq_pad * 64;
    float* vtw = kw + 2 * BH * Tk_pad * 64;
    // softmax(scale q.k) = 2^(log2e scale q.k - max) / sum: the kernel works in the log2 domain (one ex2 per score)
    PrepParams pp{p.q, p.q_ld, p.q_bs, p.k, p.k_ld, p.k_bs, p.v, p.v_ld, p.v_bs, qw, kw, vtw, p.B, p.H, p.Tq, p.Tk, Tq_pad, Tk_pad, p.scale * 1.4426950408889634f};
    dim3 pgrid((unsigned)((std::max(Tq_pad, Tk_pad) + 63) / 64), (unsigned)BH);
    attn_prep_kernel<<<pgrid, 256, 0, stream>>>(pp);
    MB_LAUNCH_CHECK();
    CUtensorMap mq, mk, mvt;
    MB_REQUIRE(fa_make_map(&mq, qw, 64, BH * Tq_pad, FA_BM) == 0, "tensor map Q");
    MB_REQUIRE(fa_make_map(&mk, kw, 64, BH * Tk_pad, FA_BN) == 0, "tensor map K");
    MB_REQUIRE(fa_make_map(&mvt, vtw, Tk_pad, BH * 64, FA_HD) == 0, "tensor map V^T");
    static bool configured = false;
    const int smem = FA_Q_BYTES + FA_STAGES * FA_STAGE_BYTES + (int)sizeof(FaBarriers) + 1024;
    if (!configured) {
        MB_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    FaParams fp{p.o, p.o_ld, p.o_bs, p.B, p.H, p.Tq, p.Tk, Tq_pad, Tk_pad, p.mask_mode, p.q_pos0, p.band, p.key_valid, p.key_valid_ld};
    dim3 grid((unsigned)(Tq_pad / FA_BM), (unsigned)BH);
    attention_tc_kernel<<<grid, FA_THREADS, smem, stream>>>(mq, mk, mvt, fp, ctx->err);
    MB_LAUNCH_CHECK();
    g_launch_count += 2;
    return 0;
}

}  // namespace mb200
