// Per-token decode path of the osuT5 decoder (reference: HF WhisperDecoderLayer x12 + proj_out called once per token by
// GenerationMixin._sample, ~330 launches + a host sync per token — SURVEY §3.2).  Here one token = 8 kernels per layer:
//   gemv[LN1 -> q|k|v, k/v written in place into the self cache]  ->  split-KV self attention  ->
//   gemv[out_proj + residual]  ->  gemv[LN2 -> cross q]  ->  split-KV cross attention (+ merge)  ->
//   gemv[out_proj + residual]  ->  gemv[LN3 -> fc1 + GELU]  ->  gemv[fc2 + residual]
// then gemv[final LN -> proj_out] and ONE kernel that runs the whole logits-processor chain (server.py:106-134 +
// HF min_new_tokens / top-k / top-p), selects the token, tests the EOS set, appends to `ids`, and writes the next step's
// embedding.  No host synchronisation per token; every step-varying scalar is read from GenState in device memory.
//
// The GEMVs are weight-streaming (HBM-bound): one warp per output row, float4 coalesced reads of the [N, K] row-major
// weight, activations for up to 8 batch rows staged in shared memory, fp32 accumulation in a fixed order.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "decode_device.cuh"

namespace mb200 {
namespace {

constexpr int GEMV_THREADS = 128, GEMV_WARPS = GEMV_THREADS / 32;

template <int NB>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_kernel(GemvParams p) {
    extern __shared__ __align__(16) float xs[];   // [NB][K] activations, then 32 floats of LayerNorm reduction scratch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int cur_pos = p.st ? p.st->cur_len - 1 : 0;
    for (int b0 = 0; b0 < p.B; b0 += NB) {
        gemv_stage_x<NB, GEMV_THREADS>(p, b0, xs, xs + NB * p.K, tid);
        __syncthreads();
        for (int n = blockIdx.x * GEMV_WARPS + warp; n < p.N; n += gridDim.x * GEMV_WARPS)
            gemv_row<NB, true>(p, n, p.W + (long long)n * p.ldw, xs, b0, lane, cur_pos);
        __syncthreads();
    }
}

template <int KMAX>
__global__ void __launch_bounds__(128, KMAX == 64 ? 6 : 4) decode_attention_kernel(DecAttnParams p) {
    __shared__ float sc[128];
    __shared__ float red[4][64];
    __shared__ float stat[2];
    pdl_launch_dependents();
    pdl_wait();
    const int L = p.fixed_len > 0 ? p.fixed_len : p.st->cur_len;
    const int P = p.st ? p.st->prompt_len : 0;
    const int r = blockIdx.z;
    const int slot = p.row_slot ? p.row_slot[r] : r;
    AttnRegs<4, KMAX> regs;
    decode_attention_load<4, KMAX>(p, blockIdx.x, blockIdx.y, r, slot, L, P, threadIdx.x, regs);
    decode_attention_body<4, KMAX>(p, blockIdx.x, blockIdx.y, r, slot, L, P, sc, red, stat, threadIdx.x, regs);
}

// batch form: one warp per (split, head, row) unit, 8 units per CTA (see decode_attention_warp_body)
__global__ void __launch_bounds__(256) decode_attention_warp_kernel(DecAttnParams p) {
    __shared__ float sc[8][128];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.x * 8 + warp;
    if (u >= p.rows * p.H * p.n_splits) return;
    const int L = p.fixed_len > 0 ? p.fixed_len : p.st->cur_len;
    const int P = p.st ? p.st->prompt_len : 0;
    const int hr = u / p.n_splits, s = u - hr * p.n_splits, r = hr / p.H, h = hr - r * p.H;
    decode_attention_warp_body(p, s, h, r, p.row_slot ? p.row_slot[r] : r, L, P, sc[warp], lane);
}

__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(SampleParams p) {
    __shared__ SampleSmem sm;
    pdl_launch_dependents();
    pdl_wait();
    if (p.st->all_finished) return;   // replays past the end of a call are no-ops (uniform across the grid)
    sample_body<SAMPLE_THREADS>(p, blockIdx.x, sm);
}

__global__ void prompt_scan_kernel(const long long* ids, long long ids_ld, int P, const unsigned char* vflags, int ts_start, int ts_end,
                                   int* last_ts) {
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    int lt = -1;
    for (int t = 0; t < P; ++t) {
        long long tok = ids[(long long)b * ids_ld + t];
        if (vflags[tok] & VF_SOS) lt = -1;
        else if (tok >= ts_start && tok < ts_end) lt = (int)(tok - ts_start);
    }
    last_ts[b] = lt;
}

__global__ void embed_kernel(const long long* ids, long long ids_ld, int P, const int* n_left_pad, int pos_rule_cumsum,
                             const float* tok_emb, const float* pos_emb, int d_model, float* x) {
    const int t = blockIdx.x, r = blockIdx.y;
    const long long tok = ids[(long long)r * ids_ld + t];
    int pos = t;
    if (pos_rule_cumsum && n_left_pad) pos = max(0, t - n_left_pad[r]);
    const float4* te = reinterpret_cast<const float4*>(tok_emb + tok * d_model);
    const float4* pe = reinterpret_cast<const float4*>(pos_emb + (long long)pos * d_model);
    float4* xo = reinterpret_cast<float4*>(x + ((long long)r * P + t) * d_model);
    for (int i = threadIdx.x; i < d_model / 4; i += blockDim.x) {
        float4 a = te[i], q = pe[i];
        xo[i] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
    }
}

static int g_prof_class = 0;

template <typename Kern, typename Params>
int launch_with_attrs(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, const Params& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    ++g_launch_count;
    const bool prof = g_prof.on && g_prof.n < 512;
    if (prof) MB_CUDA_CHECK(cudaEventRecord(g_prof.ev[2 * g_prof.n], stream));
    MB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
    if (prof) { MB_CUDA_CHECK(cudaEventRecord(g_prof.ev[2 * g_prof.n + 1], stream)); g_prof.cls[g_prof.n++] = g_prof_class; }
    return 0;
}

}  // namespace

int launch_gemv(const GemvParams& p, cudaStream_t stream, bool pdl) {
    MB_REQUIRE(p.K % 4 == 0 && p.ldw % 4 == 0, "GEMV K / ldw must be multiples of 4");
    MB_REQUIRE(p.xmode != X_LAYERNORM || p.K <= 1024, "fused LayerNorm prologue supports K <= 1024");
    if (p.B <= 0 || p.N <= 0) return 0;
    int nb = p.B >= 8 ? 8 : (p.B > 4 ? 8 : (p.B > 2 ? 4 : p.B));
    const size_t smem = ((size_t)nb * p.K + 32) * sizeof(float);
    const int blocks = (p.N + GEMV_WARPS - 1) / GEMV_WARPS;
    static bool configured = false;
    if (!configured) {
        MB_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MB_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MB_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MB_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    MB_REQUIRE(smem <= 200 * 1024, "GEMV activation tile does not fit shared memory");
    g_prof_class = 0;
    switch (nb) {
        case 1: return launch_with_attrs(gemv_kernel<1>, dim3(blocks), dim3(GEMV_THREADS), smem, stream, pdl, p);
        case 2: return launch_with_attrs(gemv_kernel<2>, dim3(blocks), dim3(GEMV_THREADS), smem, stream, pdl, p);
        case 4: return launch_with_attrs(gemv_kernel<4>, dim3(blocks), dim3(GEMV_THREADS), smem, stream, pdl, p);
        default: return launch_with_attrs(gemv_kernel<8>, dim3(blocks), dim3(GEMV_THREADS), smem, stream, pdl, p);
    }
}

int launch_decode_attention(const DecAttnParams& p, cudaStream_t stream, bool pdl) {
    MB_REQUIRE(p.chunk > 0 && p.chunk <= 128, "decode attention chunk must be in (0, 128]");
    MB_REQUIRE(p.out && p.ticket, "decode attention needs the merged-output buffer and its tickets");
    if (p.rows <= 0) return 0;
    g_prof_class = 1;
    // Default: one CTA per unit with everything prefetched (the megakernel's phase body).  MB200_ATTN_BATCH=1 selects the
    // one-warp-per-unit form for rows > 2 — the same arithmetic value for value (parity-tested), but measured SLOWER on B200
    // (B = 64: 3.04 vs 3.30 TB/s, B = 8: 0.66 vs 1.15 TB/s): kept as the starting point for a persistent multi-unit kernel.
    static const int batch_form = [] { const char* e = getenv("MB200_ATTN_BATCH"); return e ? atoi(e) : 0; }();
    if (p.rows > 2 && batch_form) {
        const int units = p.rows * p.H * p.n_splits;
        return launch_with_attrs(decode_attention_warp_kernel, dim3((units + 7) / 8), dim3(256), 0, stream, pdl, p);
    }
    // 64-key chunks (every split launch) take the instantiation with half the K registers: 6 resident CTAs per SM instead of 4
    if (p.chunk <= 64) return launch_with_attrs(decode_attention_kernel<64>, dim3(p.n_splits, p.H, p.rows), dim3(128), 0, stream, pdl, p);
    return launch_with_attrs(decode_attention_kernel<128>, dim3(p.n_splits, p.H, p.rows), dim3(128), 0, stream, pdl, p);
}

int launch_sample(const SampleParams& p, int B, cudaStream_t stream, bool pdl) {
    g_prof_class = 2;
    return launch_with_attrs(sample_kernel, dim3(B), dim3(SAMPLE_THREADS), 0, stream, pdl, p);
}

int launch_prompt_scan(const long long* ids, long long ids_ld, int B, int P, const unsigned char* vflags, int ts_start, int ts_end,
                       int* last_ts, cudaStream_t stream) {
    prompt_scan_kernel<<<B, 32, 0, stream>>>(ids, ids_ld, P, vflags, ts_start, ts_end, last_ts);
    MB_LAUNCH_CHECK();
    ++g_launch_count;
    return 0;
}

int launch_embed(const long long* ids, long long ids_ld, int rows, int B_ids, int P, const int* n_left_pad, int pos_rule_cumsum,
                 const float* tok_emb, const float* pos_emb, int d_model, float* x, cudaStream_t stream) {
    (void)B_ids;
    if (rows <= 0 || P <= 0) return 0;
    embed_kernel<<<dim3(P, rows), 128, 0, stream>>>(ids, ids_ld, P, n_left_pad, pos_rule_cumsum, tok_emb, pos_emb, d_model, x);
    MB_LAUNCH_CHECK();
    ++g_launch_count;
    return 0;
}

}  // namespace mb200
