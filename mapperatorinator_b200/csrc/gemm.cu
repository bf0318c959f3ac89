// fp32 GEMM with fused epilogues:  C = R + gate * ( act(A . W^T + bias) * alpha )
//
// Serves every dense projection of the hot path at fp32 accuracy (bit-exact-greedy parity needs fp32-grade sums;
// the reference's CPU path is fp32): encoder_embedder, conv1/conv2 (as im2col-free GEMMs over a padded
// token-major buffer, see RowMap), Whisper q/k/v/out/fc1/fc2 in encoder + prefill, DiT qkv/out/fc1/fc2/adaLN.
//   A: [M, K] via RowMap (K contiguous), W: [N, K] row-major (torch Linear layout), C/R: [M, N] via RowMap.
// Tile 128x128x16, 256 threads, 8x8 micro-tile, double-buffered shared memory with register prefetch.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace mb200 {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256, 2) gemm_f32_kernel(GemmParams p) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    // global->smem loader coordinates: two rows per operand per thread, one float4 along K each
    const int lrow = tid >> 2;           // 0..63
    const int lk = (tid & 3) * 4;        // 0,4,8,12
    const float* a_ptr[2];
    const float* w_ptr[2];
    bool a_ok[2], w_ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        long long m = m0 + lrow + h * 64;
        a_ok[h] = m < p.M;
        a_ptr[h] = a_ok[h] ? p.A.row(m) : p.A.ptr;
        int n = n0 + lrow + h * 64;
        w_ok[h] = n < p.N;
        w_ptr[h] = p.W + (long long)(w_ok[h] ? n : 0) * p.ldw;
    }

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rw[2];
    const int k_lo_ = p.splitk > 1 ? blockIdx.z * p.k_per_split : 0;
    const int k_hi_ = p.splitk > 1 ? min(p.K, k_lo_ + p.k_per_split) : p.K;
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int k = k_lo_ + k0 + lk;
            ra[h] = (a_ok[h] && k < k_hi_) ? __ldg(reinterpret_cast<const float4*>(a_ptr[h] + k)) : make_float4(0, 0, 0, 0);
            rw[h] = (w_ok[h] && k < k_hi_) ? __ldg(reinterpret_cast<const float4*>(w_ptr[h] + k)) : make_float4(0, 0, 0, 0);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int r = lrow + h * 64;
            As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk + 0][r] = rw[h].x; Bs[buf][lk + 1][r] = rw[h].y; Bs[buf][lk + 2][r] = rw[h].z; Bs[buf][lk + 3][r] = rw[h].w;
        }
    };

    const int k_lo = p.splitk > 1 ? blockIdx.z * p.k_per_split : 0;
    const int k_hi = p.splitk > 1 ? min(p.K, k_lo + p.k_per_split) : p.K;
    const int nk = (k_hi - k_lo + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue --------------------------------------------------------------------------------------------------
    if (p.splitk > 1) {      // raw partial sums; gemm_splitk_reduce_kernel finishes the job
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (m >= p.M) continue;
            float* wrow = p.splitk_ws + ((long long)blockIdx.z * p.M + m) * p.N;
#pragma unroll
            for (int jh = 0; jh < 2; ++jh)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int nn = n0 + jh * 64 + tx * 4 + j;
                    if (nn < p.N) wrow[nn] = acc[i][jh * 4 + j];
                }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
        float* crow = p.C.row(m);
        const float* rrow = p.R.ptr ? p.R.row(m) : nullptr;
        const float* grow = p.gate ? p.gate + (m / p.gate_rpb) * p.gate_ld : nullptr;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            int n = n0 + jh * 64 + tx * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int nn = n + j;
                if (nn >= p.N) continue;
                float v = acc[i][jh * 4 + j];
                if (p.bias) v += __ldg(p.bias + nn);
                v = apply_act(v, p.act) * p.alpha;
                if (grow) v *= __ldg(grow + nn);
                if (rrow) v += rrow[nn];
                crow[nn] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(GemmParams p) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)p.M * p.N) return;
    const long long m = idx / p.N;
    const int n = (int)(idx - m * p.N);
    float v = 0.f;
    for (int z = 0; z < p.splitk; ++z) v += p.splitk_ws[((long long)z * p.M + m) * p.N + n];   // fixed order
    if (p.bias) v += __ldg(p.bias + n);
    v = apply_act(v, p.act) * p.alpha;
    if (p.gate) v *= __ldg(p.gate + (m / p.gate_rpb) * p.gate_ld + n);
    if (p.R.ptr) v += p.R.row(m)[n];
    p.C.row(m)[n] = v;
}

// ---- skinny variant: M <= 64 rows (decoder prefill of a 17-50 token prompt) ------------------------------------------------
// The 128x128 tile kernel above would run such a problem on N/128 = 6..24 CTAs with a latency-bound K loop (measured 115 us
// per projection, 11.5 ms per prefill).  Here the weight matrix is the streamed operand: one CTA owns 16 weight rows, lanes
// own the (<= 64) activation rows, the activation chunk sits transposed in shared memory and every weight value is a
// shared-memory broadcast.  No cross-lane reduction; fixed k order.
constexpr int SK_ROWS = 16, SK_KC = 128, SK_MPAD = 65;

__global__ void __launch_bounds__(128) gemm_skinny_kernel(GemmParams p) {
    __shared__ float xT[SK_KC][SK_MPAD];                 // [k][m], padded: conflict-free transposed stores
    __shared__ __align__(16) float ws[SK_ROWS][SK_KC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * SK_ROWS;
    float acc[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[h][r] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += SK_KC) {
        const int kc = min(SK_KC, p.K - k0);             // multiple of 4
        for (int idx = tid; idx < 64 * (SK_KC / 4); idx += 128) {
            const int m = idx / (SK_KC / 4), kq = (idx - m * (SK_KC / 4)) * 4;
            float4 v = make_float4(0, 0, 0, 0);
            if (m < p.M && kq < kc) v = __ldg(reinterpret_cast<const float4*>(p.A.row(m) + k0 + kq));
            xT[kq + 0][m] = v.x; xT[kq + 1][m] = v.y; xT[kq + 2][m] = v.z; xT[kq + 3][m] = v.w;
        }
        for (int idx = tid; idx < SK_ROWS * (SK_KC / 4); idx += 128) {
            const int r = idx / (SK_KC / 4), kq = (idx - r * (SK_KC / 4)) * 4;
            float4 v = make_float4(0, 0, 0, 0);
            if (n0 + r < p.N && kq < kc) v = __ldg(reinterpret_cast<const float4*>(p.W + (long long)(n0 + r) * p.ldw + k0 + kq));
            *reinterpret_cast<float4*>(&ws[r][kq]) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < SK_KC; kk += 4) {
            float4 w4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) w4[r] = *reinterpret_cast<const float4*>(&ws[warp * 4 + r][kk]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x0 = xT[kk + j][lane], x1 = xT[kk + j][lane + 32];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float wv = j == 0 ? w4[r].x : (j == 1 ? w4[r].y : (j == 2 ? w4[r].z : w4[r].w));
                    acc[0][r] = fmaf(wv, x0, acc[0][r]);
                    acc[1][r] = fmaf(wv, x1, acc[1][r]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = lane + 32 * h;
        if (m >= p.M) continue;
        float* crow = p.C.row(m);
        const float* rrow = p.R.ptr ? p.R.row(m) : nullptr;
        const float* grow = p.gate ? p.gate + (m / p.gate_rpb) * p.gate_ld : nullptr;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = n0 + warp * 4 + r;
            if (n >= p.N) continue;
            float v = acc[h][r];
            if (p.bias) v += __ldg(p.bias + n);
            v = apply_act(v, p.act) * p.alpha;
            if (grow) v *= __ldg(grow + n);
            if (rrow) v += rrow[n];
            crow[n] = v;
        }
    }
}

}  // namespace

// fixed-size split-K workspace, allocated once (captured graphs hold this pointer); null if `need` does not fit
float* splitk_workspace(size_t need) {
    static float* ws = nullptr;
    static const size_t ws_bytes = (size_t)64 << 20;
    if (!ws && cudaMalloc(&ws, ws_bytes) != cudaSuccess) { ws = nullptr; return nullptr; }
    return need <= ws_bytes ? ws : nullptr;
}

int launch_splitk_reduce(const GemmParams& q, cudaStream_t stream) {
    gemm_splitk_reduce_kernel<<<(unsigned)(((long long)q.M * q.N + 255) / 256), 256, 0, stream>>>(q);
    MB_LAUNCH_CHECK();
    ++g_launch_count;
    return 0;
}

int launch_gemm(const GemmParams& p, cudaStream_t stream) {
    MB_REQUIRE(p.K % 4 == 0, "GEMM K must be a multiple of 4 (float4 loads)");
    MB_REQUIRE(p.A.ld % 4 == 0 && p.ldw % 4 == 0, "GEMM operand row strides must be multiples of 4 floats");
    MB_REQUIRE((reinterpret_cast<uintptr_t>(p.A.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0,
               "GEMM operands must be 16-byte aligned");
    if (p.M <= 0 || p.N <= 0) return 0;
    if (tc_gemm_eligible(p) && launch_gemm_tc(p, stream) == 0) return 0;      // tensor cores (3xTF32); falls through on failure
    static const int small_m_mode = [] { const char* e = getenv("MB200_SMALLM"); return e ? atoi(e) : 2; }();   // 0 tile, 1 skinny, 2 split-K
    if (p.M <= 64 && small_m_mode == 1) {
        gemm_skinny_kernel<<<(p.N + SK_ROWS - 1) / SK_ROWS, 128, 0, stream>>>(p);
        MB_LAUNCH_CHECK();
        ++g_launch_count;
        return 0;
    }
    const int tiles_m = (int)((p.M + BM - 1) / BM), tiles_n = (p.N + BN - 1) / BN;
    if ((long long)tiles_m * tiles_n < 100 && small_m_mode == 2 && p.K >= 64) {
        // fewer output tiles than SMs (decoder prefill, a single encoder window, DiT conditioning): split K over enough CTAs to
        // fill the machine — these problems are weight-streaming bound, not FLOP bound
        const int tiles = tiles_m * tiles_n;
        int splits = std::max(1, std::min((148 + tiles - 1) / tiles, p.K / 32));
        const int kps = ((p.K + splits - 1) / splits + BK - 1) / BK * BK;
        splits = (p.K + kps - 1) / kps;
        float* ws = splits > 1 ? splitk_workspace((size_t)splits * p.M * p.N * sizeof(float)) : nullptr;
        if (ws) {
            GemmParams q = p;
            q.splitk_ws = ws; q.splitk = splits; q.k_per_split = kps;
            gemm_f32_kernel<<<dim3(tiles_n, tiles_m, splits), 256, 0, stream>>>(q);
            MB_LAUNCH_CHECK();
            ++g_launch_count;
            return launch_splitk_reduce(q, stream);
        }
    }
    dim3 grid((p.N + BN - 1) / BN, (unsigned)((p.M + BM - 1) / BM));
    gemm_f32_kernel<<<grid, 256, 0, stream>>>(p);
    MB_LAUNCH_CHECK();
    ++g_launch_count;
    return 0;
}

}  // namespace mb200
