// Tensor-core GEMM for the dense phases (encoder layers, conv stem, DiT blocks): tcgen05.mma kind::tf32 with fp32-grade
// accuracy through the 3xTF32 split
//        A.W^T  ~=  Ahi.Whi^T + Ahi.Wlo^T + Alo.Whi^T ,   x = xhi + xlo,  xhi = rn_tf32(x), xlo = rn_tf32(x - xhi).
// Both parts are rounded to nearest tf32 (cvt.rna) — weights once at load, activations by a tiny elementwise pass — because the
// tensor core would otherwise TRUNCATE the low 13 bits, and truncation bias adds up linearly over K.  Accumulation is fp32 in TMEM.  Measured error vs fp64 is ~1e-6 relative — the same order as an fp32 FMA chain of that
// length — which is what bit-exact greedy decoding against the fp32 reference needs; plain TF32 (1e-3) flips tokens.
//
// Structure (one 128x128 output tile per CTA, 192 threads):
//   warp 0 / lane 0 : TMA producer — 4 tiles per k-block (A, Alo, W, Wlo; 128 rows x 32 floats, SWIZZLE_128B) into a 3-stage ring
//   warp 1 / lane 0 : MMA issuer   — 12 x tcgen05.mma (128x128x8) per k-block into a 128-column fp32 TMEM accumulator,
//                     tcgen05.commit frees the stage / publishes the accumulator
//   warps 2..5      : epilogue     — tcgen05.ld (32 lanes x 32 columns per warp and pass) -> per-warp shared-memory transpose ->
//                     bias / activation / gate / residual (same GemmParams epilogue as gemm.cu) -> 128-byte coalesced stores
// Measured alternative (kept out): deriving hi/lo inside the kernel from raw tiles (half the L2->SM operand traffic, no mirror
// copies) was 10-15 % SLOWER — the split's shared-memory traffic competes with the MMA's own operand reads.
// A may be any RowMap (im2col-free conv over the padded buffer, batched rows) via a 3-D tensor map.
// Every wait is bounded: on a timeout the kernel sets an error flag and falls through, it can never hang the GPU.
#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>
#include <tuple>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

namespace mb200 {

int g_tc_enabled = 1;

namespace {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3, TC_THREADS = 192;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;          // 16 KB
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;         // A, Alo, W, Wlo

struct TcBarriers {
    unsigned long long full[TC_STAGES];
    unsigned long long empty[TC_STAGES];
    unsigned long long tmem_full;
    unsigned int tmem_base;
    int pad;
};

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool bar_wait(unsigned long long* bar, unsigned parity, int* err) {
    for (long long spin = 0; spin < (1ll << 22); ++spin) {
        unsigned ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicExch(err, 3);
    return false;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024B>>4 | version 1 | layout 2
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)(1024 >> 4) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)2 << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"l"(__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(unsigned (&v)[32], unsigned taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_alo,
                   const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo, GemmParams p, int a_rpb, int* err) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem + TC_STAGES * TC_STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * TC_BN;
    const long long m0 = (long long)blockIdx.y * TC_BM;
    const int nkb_all = (p.K + TC_BK - 1) / TC_BK;
    // Split-K (kernels.h): k-range z = k-blocks [z*kpb, (z+1)*kpb) is summed on its own.  Grid split (under-filled grids): CTA
    // blockIdx.z owns range z and stores raw partials, gemm.cu's reduce kernel adds them in order.  In-tile split (large M): this CTA
    // walks every range, range z accumulating into TMEM columns [z*128, z*128+128); the epilogue adds the S accumulators in the same
    // order.  Same partial sums, same additions, same bits.
    const int kpb = p.splitk > 1 ? p.k_per_split / TC_BK : nkb_all;
    const bool grid_split = p.split_mode == 1, tile_split = p.split_mode == 2;
    const int kb0 = grid_split ? blockIdx.z * kpb : 0;
    const int nkb = grid_split ? max(0, min(nkb_all - kb0, kpb)) : nkb_all;
    const int nacc = tile_split ? p.splitk : 1;
    const unsigned tmem_cols = nacc == 1 ? 128u : (nacc == 2 ? 256u : 512u);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars->full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars->empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars->tmem_full)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            const int a_b = a_rpb > 0 ? (int)(m0 / a_rpb) : 0;
            const int a_t = a_rpb > 0 ? (int)(m0 - (long long)a_b * a_rpb) : (int)m0;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % TC_STAGES;
                const unsigned ph = (kb / TC_STAGES) & 1;
                if (!bar_wait(&bars->empty[s], ph ^ 1, err)) break;
                unsigned char* st = smem + s * TC_STAGE_BYTES;
                const unsigned fb = s32(&bars->full[s]);
                asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(fb), "r"(TC_STAGE_BYTES) : "memory");
                const int k0 = (kb0 + kb) * TC_BK;
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(s32(st)), "l"(&map_a), "r"(k0), "r"(a_t), "r"(a_b), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(s32(st + TC_TILE_BYTES)), "l"(&map_alo), "r"(k0), "r"(a_t), "r"(a_b), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(s32(st + 2 * TC_TILE_BYTES)), "l"(&map_w), "r"(k0), "r"(n0), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(s32(st + 3 * TC_TILE_BYTES)), "l"(&map_wlo), "r"(k0), "r"(n0), "r"(fb) : "memory");
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, N = 128, M = 128
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TC_BN >> 3) << 17) | ((unsigned)(TC_BM >> 4) << 24);
            bool ok = true;
            for (int kb = 0; kb < nkb && ok; ++kb) {
                const int s = kb % TC_STAGES;
                const unsigned ph = (kb / TC_STAGES) & 1;
                ok = bar_wait(&bars->full[s], ph, err);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned base = s32(smem + s * TC_STAGE_BYTES);
                const int z = tile_split ? kb / kpb : 0, kr = tile_split ? kb - z * kpb : kb;      // accumulator and position inside its k-range
                const unsigned acc = tmem + (unsigned)(z * TC_BN);
#pragma unroll
                for (int sub = 0; sub < TC_BK / 8; ++sub) {
                    const unsigned off = sub * 32;       // 8 tf32 = 32 bytes along K inside the 128-byte swizzle atom
                    const unsigned long long a_hi = umma_desc(base + off), a_lo = umma_desc(base + TC_TILE_BYTES + off);
                    const unsigned long long w_hi = umma_desc(base + 2 * TC_TILE_BYTES + off), w_lo = umma_desc(base + 3 * TC_TILE_BYTES + off);
                    umma_tf32(acc, a_hi, w_hi, idesc, (kr > 0 || sub > 0) ? 1u : 0u);
                    umma_tf32(acc, a_hi, w_lo, idesc, 1u);
                    umma_tf32(acc, a_lo, w_hi, idesc, 1u);
                }
                umma_commit(&bars->empty[s]);            // frees the stage once these MMAs have read it
            }
            umma_commit(&bars->tmem_full);               // accumulator complete
        }
    } else {
        const bool ok = true;                             // (the wait for the accumulator comes after the first operand prefetch, below)
        const int lg = warp & 3;                          // TMEM lane group this warp may access
        static_assert(sizeof(TcBarriers) <= 256, "barrier block");
        // Epilogue: tcgen05.ld hands every thread 32 consecutive columns of ITS row, so storing straight from registers would make
        // each store instruction touch 32 different rows (32 sectors per instruction; measured: the epilogue, not the MMA pipe,
        // set the kernel's duration).  Each warp transposes its 32x32 block through shared memory instead (the pipeline stages
        // are idle by now: every TMA load and every MMA that reads them completed before tmem_full fired) and then walks the
        // block row by row with lane = column: bias / gate / residual loads and the C store are all 128-byte coalesced.
        float* stg = reinterpret_cast<float*>(smem) + lg * (32 * 33);                    // [32 rows][33] per warp
        struct RowPtrs { float* c; const float* r; const float* g; };
        // (its own shared-memory region behind the barriers: it is written while the main loop still owns the pipeline stages)
        RowPtrs* rows = reinterpret_cast<RowPtrs*>(smem + TC_STAGES * TC_STAGE_BYTES + 256) + lg * 32;  // this warp's 32 row bases
        {
            const long long m = m0 + lg * 32 + lane;
            RowPtrs rp{nullptr, nullptr, nullptr};
            if (ok && m < p.M) {
                if (grid_split) {
                    rp.c = p.splitk_ws + ((long long)blockIdx.z * p.M + m) * p.N;         // raw partial sums of this split
                } else {
                    rp.c = p.C.row(m);
                    rp.r = p.R.ptr ? p.R.row(m) : nullptr;
                    rp.g = p.gate ? p.gate + (m / p.gate_rpb) * p.gate_ld : nullptr;
                }
            }
            rows[lane] = rp;
        }
        __syncwarp();
        // Residual / gate operands of a 32-column chunk do not depend on the accumulator: all 32 rows' loads are issued in one batch
        // BEFORE the chunk's TMEM read (chunk 0: before the accumulator is even complete — the epilogue warps idle through the main
        // loop), instead of one dependent L2 round trip per row behind the transpose (ncu: 20 % of the kernel's samples sat on those loads).
        const bool has_r = !grid_split && p.R.ptr != nullptr, has_g = !grid_split && p.gate != nullptr;
        float rv[32], gv[32];
        auto fetch_rg = [&](int c0) {
            const int n = n0 + c0 + lane;
            const bool n_ok = n < p.N;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
                const RowPtrs rp = rows[rr];
                rv[rr] = (has_r && n_ok && rp.r) ? rp.r[n] : 0.f;
                gv[rr] = (has_g && n_ok && rp.g) ? __ldg(rp.g + n) : 1.f;
            }
        };
        if (has_r || has_g) fetch_rg(0);
        const bool ok2 = bar_wait(&bars->tmem_full, 0, err);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        (void)ok2;
#pragma unroll 1
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
            unsigned v[32];
            const unsigned taddr = tmem + ((unsigned)(lg * 32) << 16) + (unsigned)c0;
            tmem_ld32(v, taddr);
            for (int z = 1; z < nacc; ++z) {                                              // in-tile split: ((acc0 + acc1) + acc2) + acc3
                unsigned w[32];
                tmem_ld32(w, taddr + (unsigned)(z * TC_BN));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);     // bank (lane + j) % 32: conflict-free
            __syncwarp();
            const int n = n0 + c0 + lane;
            const bool n_ok = n < p.N;
            const float bias_v = (n_ok && p.bias) ? __ldg(p.bias + n) : 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
                const RowPtrs rp = rows[rr];                                              // broadcast
                if (rp.c && n_ok) {
                    float x = stg[rr * 33 + lane];
                    if (grid_split) { rp.c[n] = x; continue; }
                    if (p.bias) x += bias_v;
                    x = apply_act(x, p.act) * p.alpha;
                    if (has_g) x *= gv[rr];
                    if (has_r) x += rv[rr];
                    rp.c[n] = x;
                }
            }
            __syncwarp();                                                                 // block consumed before the next chunk overwrites it
            if ((has_r || has_g) && c0 + 32 < TC_BN) fetch_rg(c0 + 32);                   // next chunk's operands travel while its TMEM read / transpose run
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// x = hi + lo with hi = round-to-nearest tf32(x) and lo = round-to-nearest tf32(x - hi).  Rounding (not truncating) both parts
// matters: truncation errors all point toward zero and add up linearly over K (measured 8e-6 relative at K = 3072, 70x the
// fp32 FMA kernel); rounded parts leave unbiased ~2^-22 errors that grow like sqrt(K).
__device__ __forceinline__ float rn_tf32(float x) {
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__global__ void tf32_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 h, l;
    h.x = rn_tf32(v.x); h.y = rn_tf32(v.y); h.z = rn_tf32(v.z); h.w = rn_tf32(v.w);
    l.x = rn_tf32(v.x - h.x); l.y = rn_tf32(v.y - h.y); l.z = rn_tf32(v.z - h.z); l.w = rn_tf32(v.w - h.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows x K] fp32, K contiguous, optional batching: dims {K, rpb, batches}; box {32, 128, 1}; 128-byte swizzle; OOB reads give 0
int make_map(CUtensorMap* out, const float* base, long long K, long long rows_per_batch, long long ld, long long batches, long long bstride,
             int rank) {
    EncodeTiledFn fn = encode_fn();
    MB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows_per_batch, (cuuint64_t)batches};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(bstride > 0 ? bstride : ld * rows_per_batch) * 4};
    cuuint32_t box[3] = {TC_BK, TC_BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

}  // namespace

// S of an (N, K) problem on the tensor-core path: the split that fills the machine for ONE encoder window (M = 512 rows = 4 row
// tiles), at least 8 k-blocks per range, at most 4 ranges (4 x 128 TMEM columns).  A function of N and K only.
int gemm_splits_tc(int N, int K, int num_sms, int* k_per_split) {
    const int tiles_ref = 4 * ((N + TC_BN - 1) / TC_BN), nkb = (K + TC_BK - 1) / TC_BK;
    int splits = std::max(1, std::min({4, num_sms / std::max(1, tiles_ref), nkb / 8}));
    const int kpb = (nkb + splits - 1) / splits;                  // k-blocks per range
    splits = (nkb + kpb - 1) / kpb;                               // no empty range
    *k_per_split = kpb * TC_BK;
    return splits;
}

// One-time self test of the tensor-core path against a host fp64 product (grid split, in-tile split and unsplit shapes).  If the
// tcgen05 pipeline misbehaves on this driver / device the path is switched off LOUDLY and every GEMM stays on the fp32 SIMT kernel.
static int g_tc_tested = 0;
static void tc_self_test() {
    g_tc_tested = 1;
    GemmCtx ctx;
    ctx.num_sms = default_gemm_ctx()->num_sms;
    bool all_ok = true;
    const int shapes[3][3] = {{512, 128, 96}, {512, 128, 1024}, {2048, 256, 1024}};      // unsplit, grid split, in-tile split
    for (int t = 0; t < 3 && all_ok; ++t) {
        const int M = shapes[t][0], N = shapes[t][1], K = shapes[t][2];
        std::vector<float> a((size_t)M * K), w((size_t)N * K), c((size_t)M * N);
        unsigned s = 12345u + t;
        auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
        for (auto& v : a) v = rnd();
        for (auto& v : w) v = rnd();
        float *da = nullptr, *dw = nullptr, *dc = nullptr;
        bool ok = cudaMalloc(&da, a.size() * 4) == cudaSuccess && cudaMalloc(&dw, w.size() * 4) == cudaSuccess && cudaMalloc(&dc, c.size() * 4) == cudaSuccess;
        if (ok) {
            cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
            GemmParams g{};
            g.A = plain_map(da, K); g.W = dw; g.ldw = K; g.C = plain_map(dc, N); g.alpha = 1.f; g.gate_rpb = 1; g.M = M; g.N = N; g.K = K;
            ok = ctx.register_weight(dw, (long long)N * K) == 0 && launch_gemm_tc(g, nullptr, &ctx) == 0 && cudaDeviceSynchronize() == cudaSuccess &&
                 ctx.error() == 0;
            if (ok) {
                cudaMemcpy(c.data(), dc, c.size() * 4, cudaMemcpyDeviceToHost);
                double worst = 0;
                for (int m = 0; m < M; m += 37)
                    for (int n = 0; n < N; n += 5) {
                        double r = 0;
                        for (int k = 0; k < K; ++k) r += (double)a[(size_t)m * K + k] * (double)w[(size_t)n * K + k];
                        worst = std::max(worst, std::fabs(r - (double)c[(size_t)m * N + n]));
                    }
                ok = worst < 1e-3;
            }
        }
        if (da) cudaFree(da);
        if (dw) cudaFree(dw);
        if (dc) cudaFree(dc);
        all_ok = all_ok && ok;
    }
    ctx.destroy();
    if (!all_ok) {
        g_tc_enabled = 0;
        fprintf(stderr, "[mapperatorinator_b200] WARNING: tcgen05 GEMM self-test FAILED — tensor-core path disabled, using the fp32 SIMT GEMM\n");
        cudaGetLastError();
    }
}

bool tc_gemm_eligible(const GemmParams& p, GemmCtx* ctx) {
    if (g_tc_enabled && !g_tc_tested) tc_self_test();
    if (!g_tc_enabled || p.m_base != 0) return false;
    if (p.M < 512 || p.N < 64 || p.K < 32 || p.K % 4 != 0) return false;
    if (ctx->mirrors.find(p.W) == ctx->mirrors.end()) return false;
    if (p.A.rpb != 0 && (p.A.rpb % TC_BM) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(p.A.ptr) & 15) || (reinterpret_cast<uintptr_t>(p.W) & 15) || (p.A.ld % 4) || (p.ldw % 4)) return false;
    if (p.A.rpb != 0 && (p.A.bstride % 4)) return false;
    return true;
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t stream, GemmCtx* ctx) {
    const Tf32Mirror wm = ctx->mirrors.at(p.W);
    // extent of the buffer A rows live in (rows may overlap: im2col-free conv) and its lo mirror
    const long long batches = p.A.rpb ? (p.M + p.A.rpb - 1) / p.A.rpb : 1;
    const long long rpb = p.A.rpb ? p.A.rpb : p.M;
    const long long extent = (batches - 1) * (p.A.rpb ? p.A.bstride : 0) + (rpb - 1) * p.A.ld + p.K;
    const long long n4 = (extent + 3) / 4;
    dim3 grid((p.N + TC_BN - 1) / TC_BN, (unsigned)((p.M + TC_BM - 1) / TC_BM));
    GemmParams q = p;
    q.splitk = gemm_splits_tc(p.N, p.K, ctx->num_sms, &q.k_per_split);
    q.split_mode = 0; q.splitk_ws = nullptr;
    if (q.splitk > 1) {
        // grid split while the tiles alone would leave SMs idle, in-tile split otherwise: same bits either way
        const int tiles = (int)(grid.x * grid.y);
        q.split_mode = tiles * 4 <= ctx->num_sms * 3 ? 1 : 2;
    } else {
        q.splitk = 1;
    }
    {
        const int s = ctx->reserve(q.split_mode == 1 ? (size_t)q.splitk * p.M * p.N * sizeof(float) : 0, (size_t)n4 * 32);      // hi and lo halves
        if (s) return s;
    }
    if (q.split_mode == 1) { q.splitk_ws = ctx->splitk_ws; grid.z = q.splitk; }
    float* a_hi = ctx->a_split;
    float* a_lo = ctx->a_split + n4 * 4;
    tf32_split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(p.A.ptr, a_hi, a_lo, n4);
    MB_LAUNCH_CHECK();
    CUtensorMap ma, mal, mw, mwl;
    MB_REQUIRE(make_map(&ma, a_hi, p.K, rpb, p.A.ld, batches, p.A.rpb ? p.A.bstride : 0, 3) == 0, "tensor map Ahi");
    MB_REQUIRE(make_map(&mal, a_lo, p.K, rpb, p.A.ld, batches, p.A.rpb ? p.A.bstride : 0, 3) == 0, "tensor map Alo");
    MB_REQUIRE(make_map(&mw, wm.hi, p.K, p.N, p.ldw, 1, 0, 2) == 0, "tensor map Whi");
    MB_REQUIRE(make_map(&mwl, wm.lo, p.K, p.N, p.ldw, 1, 0, 2) == 0, "tensor map Wlo");
    static bool configured = false;
    const int smem = TC_STAGES * TC_STAGE_BYTES + 256 + 4 * 32 * 24 + 1024;      // stages | barriers (256 B) | epilogue row table | alignment slack
    if (!configured) {
        MB_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    gemm_tf32x3_kernel<<<grid, TC_THREADS, smem, stream>>>(ma, mal, mw, mwl, q, p.A.rpb, ctx->tc_err);
    MB_LAUNCH_CHECK();
    g_launch_count += 2;
    if (q.split_mode == 1) return launch_splitk_reduce(q, stream);
    return 0;
}

void GemmCtx::unregister_weight(const float* w) {
    auto it = mirrors.find(w);
    if (it == mirrors.end()) return;
    cudaFree(const_cast<float*>(it->second.hi));      // hi and lo share one allocation
    mirrors.erase(it);
}

// tf32 hi / lo mirror of a weight matrix (called once per weight at load; the owner unregisters it before freeing the weight)
int GemmCtx::register_weight(const float* w, long long numel) {
    unregister_weight(w);      // a recycled device address must never inherit a stale mirror
    float* buf = nullptr;
    const long long n4 = (numel + 3) / 4;
    MB_CUDA_CHECK(cudaMalloc(&buf, (size_t)n4 * 32));
    tf32_split_kernel<<<(unsigned)((n4 + 255) / 256), 256>>>(w, buf, buf + n4 * 4, n4);
    MB_LAUNCH_CHECK();
    mirrors[w] = Tf32Mirror{buf, buf + n4 * 4};
    return 0;
}

}  // namespace mb200
