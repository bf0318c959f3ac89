// Persistent token-loop megakernel: the WHOLE autoregressive decode of one generate() call (every layer of every token,
// the logits-processor chain and the token selection) is ONE cooperative launch of one CTA per SM.
//
// Why: at batch 1-2 a decoder step is 98 dependent micro-phases moving ~4.7 MB each (464 MB of fp32 weights per token).  As
// separate kernels every phase pays launch + DRAM-latency ramp (~2.5 us) for ~0.7 us of HBM time.  Here the phases are
// separated by a ~1 us grid barrier instead, and — because each CTA knows statically which weight rows it owns in the NEXT
// GEMV phase — those rows are pulled into shared memory by the TMA bulk-copy engine (cp.async.bulk + mbarrier
// complete_tx) while the current phase is still computing / waiting at its barrier.  HBM therefore streams weights
// continuously across phase boundaries (2 x 77 KB in flight per SM, 148 SMs => ~22 MB outstanding), which is what the
// weight-streaming roofline needs; the math runs out of shared memory.
//
// Same arithmetic, same order as the per-kernel path (decode_device.cuh bodies are shared), so tokens are bit-identical.
// Every cross-CTA value is read through L2 (ld.global.cg) — L1 is not coherent across SMs inside one launch.
// All spin loops are bounded: on a timeout the kernel raises `error_flag`, falls through every remaining barrier and exits,
// so a logic error can never hang the GPU.
#include "common.cuh"
#include "kernels.h"
#include "decode_device.cuh"

namespace mb200 {

namespace {

constexpr int MEGA_THREADS = SAMPLE_THREADS;            // 512
constexpr int MEGA_WARPS = MEGA_THREADS / 32;
constexpr int MEGA_NB = 2;                              // decoder rows handled (1 or 2)
constexpr int XS_FLOATS = MEGA_NB * 3072;

struct __align__(16) MegaSmem {
    float wbuf[2][MEGA_WBUF_FLOATS];
    union {
        float xs[XS_FLOATS];
        SampleSmem sample;
        struct { float sc[128]; float red[4][64]; float stat[2]; } attn;
    } u;
    MegaPhase phase;
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
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// rows [r0, r1) of an N-row GEMV owned by CTA `cta` of `G`
__device__ __forceinline__ void cta_rows(int N, int cta, int G, int& r0, int& r1) {
    const int rpc = (N + G - 1) / G;
    r0 = min(N, cta * rpc);
    r1 = min(N, r0 + rpc);
}

// thread 0: start streaming this CTA's weight rows of GEMV phase `ph` into wbuf[buf]
__device__ __forceinline__ void prefetch_weights(const MegaPhase* ph, float* dst, unsigned long long* bar, int cta, int G) {
    int r0, r1;
    cta_rows(ph->g.N, cta, G, r0, r1);
    const unsigned bytes = (unsigned)(r1 - r0) * (unsigned)ph->g.K * 4u;
    if (bytes == 0) { mbar_arrive(bar); return; }
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(dst, ph->g.W + (long long)r0 * ph->g.ldw, bytes, bar);
}

__device__ __forceinline__ bool wait_weights(unsigned long long* bar, unsigned parity, int* error_flag) {
    for (long long spin = 0; spin < (1ll << 24); ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
    }
    atomicExch(error_flag, 2);
    return false;
}

__device__ __forceinline__ void grid_sync(unsigned int* counter, unsigned int target, int* error_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v = 0;
        long long spin = 0;
        while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            if (++spin > (1ll << 26) || *reinterpret_cast<volatile int*>(error_flag) != 0) { atomicExch(error_flag, 1); break; }
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_megakernel(MegaParams mp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MegaSmem& sm = *reinterpret_cast<MegaSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;

    if (tid == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned int g_idx = 0;          // running index of GEMV phases (selects buffer + mbarrier parity)
    unsigned int sync_target = 0;
    if (tid == 0) prefetch_weights(&mp.phases[mp.first_gemv], sm.wbuf[0], &sm.mbar[0], cta, G);
    bool ok = true;

    for (int step = 0; step < mp.max_steps && ok; ++step) {
        if (ld_state(&mp.st->all_finished)) break;                 // uniform: written before the previous grid barrier
        const int cur_pos = ld_state(&mp.st->cur_len) - 1;
        const int P = ld_state(&mp.st->prompt_len);
        for (int pi = 0; pi < mp.n_phases && ok; ++pi) {
            // stage the phase descriptor in shared memory (immutable, so plain loads are fine)
            {
                const int* src = reinterpret_cast<const int*>(&mp.phases[pi]);
                int* dst = reinterpret_cast<int*>(&sm.phase);
                for (int i = tid; i < (int)(sizeof(MegaPhase) / 4); i += MEGA_THREADS) dst[i] = src[i];
            }
            __syncthreads();
            const MegaPhase& ph = sm.phase;
            if (ph.kind == 0) {
                const int buf = g_idx & 1;
                if (tid == 0) prefetch_weights(&mp.phases[ph.next_gemv], sm.wbuf[buf ^ 1], &sm.mbar[buf ^ 1], cta, G);
                gemv_stage_x<MEGA_NB>(ph.g, 0, sm.u.xs, tid, MEGA_THREADS);
                __syncthreads();
                ok = wait_weights(&sm.mbar[buf], (g_idx >> 1) & 1, mp.error_flag);
                int r0, r1;
                cta_rows(ph.g.N, cta, G, r0, r1);
                if (ok)
                    for (int n = r0 + warp; n < r1; n += MEGA_WARPS)
                        gemv_row<MEGA_NB, false>(ph.g, n, sm.wbuf[buf] + (long long)(n - r0) * ph.g.K, sm.u.xs, 0, lane, cur_pos);
                ++g_idx;
            } else if (ph.kind == 1) {
                const DecAttnParams& a = ph.a;
                const int L = a.fixed_len > 0 ? a.fixed_len : cur_pos + 1;
                const int units = a.rows * a.H * a.n_splits;
                for (int u = cta; u < units; u += G) {
                    const int s = u % a.n_splits, h = (u / a.n_splits) % a.H, r = u / (a.n_splits * a.H);
                    decode_attention_body<MEGA_WARPS>(a, s, h, r, L, P, sm.u.attn.sc, sm.u.attn.red, sm.u.attn.stat, tid);
                    __syncthreads();
                }
            } else {
                if (cta < mp.sample.cfg->B) sample_body(mp.sample, cta, sm.u.sample);
            }
            sync_target += G;
            grid_sync(mp.sync_counter, sync_target, mp.error_flag);
            if (*reinterpret_cast<volatile int*>(mp.error_flag) != 0) ok = false;
        }
    }
    // drain the weight prefetch that is still in flight so no bulk copy outlives the CTA
    if (ok) wait_weights(&sm.mbar[g_idx & 1], (g_idx >> 1) & 1, mp.error_flag);
}

}  // namespace

size_t mega_smem_bytes() { return sizeof(MegaSmem) + 128; }

int launch_megakernel(const MegaParams& mp, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        MB_CUDA_CHECK(cudaFuncSetAttribute(decode_megakernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem_bytes()));
        configured = true;
    }
    int per_sm = 0;
    MB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_megakernel, MEGA_THREADS, mega_smem_bytes()));
    MB_REQUIRE(per_sm >= 1, "megakernel does not fit on an SM");
    MegaParams p = mp;
    void* args[] = {&p};
    MB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)decode_megakernel, dim3(grid), dim3(MEGA_THREADS), args, mega_smem_bytes(), stream));
    ++g_launch_count;
    return 0;
}

}  // namespace mb200
