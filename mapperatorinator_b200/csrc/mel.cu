// Fused raw-PCM -> (log-)mel spectrogram: centre padding + Hann window + 1024-point STFT + |.|^2 + mel filterbank in
// ONE kernel; the complex spectrum never touches HBM (reference: two cuDNN conv1d of 513x1024 taps + a cuBLAS matmul in
// nnAudio, or cuFFT + matmul in torchaudio — osuT5/osuT5/model/spectrogram.py:38-61,79-83).
//   HBM traffic = PCM in (523 776 B/window) + mel out (1 589 248 B/window at 388 mels), the algorithmic minimum.
// One CTA = 16 consecutive frames of one window: their 2944 samples are staged once in shared memory (hop 128 => 8x
// reuse), frames are transformed two at a time as the real/imag parts of one complex 1024-point Stockham radix-4 FFT
// (5 passes, ping-pong in shared memory), split with the conjugate-symmetry identity, squared, and projected through the
// mel filterbank held in CSR form (each FFT bin feeds <= 2 triangles), written coalesced as (B, frames, n_mels).
#include <vector>
#include "common.cuh"
#include "kernels.h"

namespace mb200 {

struct MelPlan {
    int n_fft, hop, n_mels, n_freq, pad_reflect, log_scale;
    float2* d_twiddle = nullptr;   // exp(-2*pi*i*k/n_fft), k < n_fft
    int* d_start = nullptr;        // [n_mels] first bin of the filter support
    int* d_count = nullptr;        // [n_mels] support length
    int* d_offset = nullptr;       // [n_mels] offset into d_weights
    float* d_weights = nullptr;
};

namespace {

constexpr int NFFT = 1024, HOP = 128, FRAMES_PER_CTA = 16, NFREQ = NFFT / 2 + 1;
constexpr int STAGE_SAMPLES = (FRAMES_PER_CTA - 1) * HOP + NFFT;   // 2944

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(256) mel_kernel(const float* __restrict__ pcm, long long pcm_ld, int n_samples, int n_frames,
                                                  float* __restrict__ mel, long long mel_ld, long long mel_bs,
                                                  const float2* __restrict__ g_tw, const int* __restrict__ f_start,
                                                  const int* __restrict__ f_count, const int* __restrict__ f_offset,
                                                  const float* __restrict__ f_weights, int n_mels, int pad_reflect, int log_scale) {
    __shared__ __align__(16) float xs[STAGE_SAMPLES];
    __shared__ __align__(16) float2 bufA[NFFT];
    __shared__ __align__(16) float2 bufB[NFFT];
    __shared__ __align__(16) float2 tw[NFFT];
    __shared__ float pw[2][NFREQ + 3];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * FRAMES_PER_CTA;
    const float* x = pcm + (long long)b * pcm_ld;

    for (int i = tid; i < NFFT; i += 256) tw[i] = g_tw[i];
    // stage samples: xs[j] = padded[f0*HOP + j], padded index q maps to original q - NFFT/2
    for (int j = tid; j < STAGE_SAMPLES; j += 256) {
        long long q = (long long)f0 * HOP + j - NFFT / 2;
        float v = 0.f;
        if (q >= 0 && q < n_samples) v = x[q];
        else if (pad_reflect) {
            long long r = q < 0 ? -q : 2LL * (n_samples - 1) - q;
            if (r >= 0 && r < n_samples) v = x[r];
        }
        xs[j] = v;
    }
    __syncthreads();

    for (int pair = 0; pair < FRAMES_PER_CTA / 2; ++pair) {
        const int fa = pair * 2, fb = fa + 1;
        if (f0 + fa >= n_frames) break;
        // windowed load: z[n] = hann[n] * (x_a[n] + i x_b[n]),  hann[n] = 0.5 - 0.5 cos(2 pi n / N) = 0.5 - 0.5 Re(tw[n])
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int n = tid + r * 256;
            float w = 0.5f - 0.5f * tw[n].x;
            bufA[n] = make_float2(w * xs[fa * HOP + n], w * xs[fb * HOP + n]);
        }
        __syncthreads();
        // 5 Stockham radix-4 passes, p = 1, 4, 16, 64, 256
        float2* src = bufA;
        float2* dst = bufB;
#pragma unroll
        for (int pass = 0; pass < 5; ++pass) {
            const int p = 1 << (2 * pass);
            const int i = tid, t = 256;
            const int k = i & (p - 1);
            const int j = ((i - k) << 2) + k;
            const int tstep = 256 / p;                     // twiddle index = k * N / (4 p)
            float2 u0 = src[i];
            float2 u1 = cmul(tw[k * tstep], src[i + t]);
            float2 u2 = cmul(tw[2 * k * tstep], src[i + 2 * t]);
            float2 u3 = cmul(tw[3 * k * tstep], src[i + 3 * t]);
            float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y);
            float2 v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
            float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
            float2 d13 = make_float2(u1.x - u3.x, u1.y - u3.y);
            float2 v3 = make_float2(d13.y, -d13.x);        // (u1 - u3) * (-i)
            dst[j] = make_float2(v0.x + v2.x, v0.y + v2.y);
            dst[j + p] = make_float2(v1.x + v3.x, v1.y + v3.y);
            dst[j + 2 * p] = make_float2(v0.x - v2.x, v0.y - v2.y);
            dst[j + 3 * p] = make_float2(v1.x - v3.x, v1.y - v3.y);
            __syncthreads();
            float2* tmp = src; src = dst; dst = tmp;
        }
        // src now holds Z = FFT(x_a + i x_b).  Split and take |X|^2 the way both references do (sqrt, then square).
        for (int kf = tid; kf < NFREQ; kf += 256) {
            float2 z1 = src[kf];
            float2 z2 = src[(NFFT - kf) & (NFFT - 1)];
            float ar = 0.5f * (z1.x + z2.x), ai = 0.5f * (z1.y - z2.y);
            float br = 0.5f * (z1.y + z2.y), bi = -0.5f * (z1.x - z2.x);
            float ma = sqrtf(ar * ar + ai * ai), mb = sqrtf(br * br + bi * bi);
            pw[0][kf] = ma * ma;
            pw[1][kf] = mb * mb;
        }
        __syncthreads();
        // mel projection, one thread per (frame, filter)
        for (int idx = tid; idx < 2 * n_mels; idx += 256) {
            const int which = idx >= n_mels;
            const int m = idx - which * n_mels;
            const int f = f0 + fa + which;
            if (f >= n_frames) continue;
            const int s = f_start[m], c = f_count[m];
            const float* wv = f_weights + f_offset[m];
            float acc = 0.f;
            for (int q = 0; q < c; ++q) acc = fmaf(wv[q], pw[which][s + q], acc);
            if (log_scale) acc = log1pf(acc);
            mel[(long long)b * mel_bs + (long long)f * mel_ld + m] = acc;
        }
        __syncthreads();
    }
}

}  // namespace

int mel_plan_create(MelPlan** out, int n_fft, int hop, int n_mels, int pad_reflect, int log_scale, const float* basis) {
    MB_REQUIRE(n_fft == NFFT && hop == HOP, "mel kernel is specialised for n_fft 1024 / hop 128 (every shipped config)");
    MelPlan* p = new MelPlan();
    p->n_fft = n_fft; p->hop = hop; p->n_mels = n_mels; p->n_freq = NFREQ; p->pad_reflect = pad_reflect; p->log_scale = log_scale;
    std::vector<float2> tw(NFFT);
    for (int k = 0; k < NFFT; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)NFFT;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<int> start(n_mels), count(n_mels), offset(n_mels);
    std::vector<float> weights;
    for (int m = 0; m < n_mels; ++m) {
        int lo = -1, hi = -1;
        for (int k = 0; k < NFREQ; ++k)
            if (basis[(size_t)m * NFREQ + k] != 0.f) { if (lo < 0) lo = k; hi = k; }
        start[m] = lo < 0 ? 0 : lo;
        count[m] = lo < 0 ? 0 : hi - lo + 1;
        offset[m] = (int)weights.size();
        for (int k = 0; k < count[m]; ++k) weights.push_back(basis[(size_t)m * NFREQ + start[m] + k]);
    }
    if (weights.empty()) weights.push_back(0.f);
    MB_CUDA_CHECK(cudaMalloc(&p->d_twiddle, sizeof(float2) * NFFT));
    MB_CUDA_CHECK(cudaMalloc(&p->d_start, sizeof(int) * n_mels));
    MB_CUDA_CHECK(cudaMalloc(&p->d_count, sizeof(int) * n_mels));
    MB_CUDA_CHECK(cudaMalloc(&p->d_offset, sizeof(int) * n_mels));
    MB_CUDA_CHECK(cudaMalloc(&p->d_weights, sizeof(float) * weights.size()));
    MB_CUDA_CHECK(cudaMemcpy(p->d_twiddle, tw.data(), sizeof(float2) * NFFT, cudaMemcpyHostToDevice));
    MB_CUDA_CHECK(cudaMemcpy(p->d_start, start.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    MB_CUDA_CHECK(cudaMemcpy(p->d_count, count.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    MB_CUDA_CHECK(cudaMemcpy(p->d_offset, offset.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    MB_CUDA_CHECK(cudaMemcpy(p->d_weights, weights.data(), sizeof(float) * weights.size(), cudaMemcpyHostToDevice));
    *out = p;
    return 0;
}

void mel_plan_destroy(MelPlan* p) {
    if (!p) return;
    cudaFree(p->d_twiddle); cudaFree(p->d_start); cudaFree(p->d_count); cudaFree(p->d_offset); cudaFree(p->d_weights);
    delete p;
}

int launch_mel(const MelPlan* plan, const float* pcm, long long pcm_ld, int B, int n_samples, float* mel, long long mel_ld,
               long long mel_bs, cudaStream_t stream) {
    if (B <= 0) return 0;
    const int n_frames = n_samples / plan->hop + 1;
    MB_REQUIRE(!plan->pad_reflect || n_samples > NFFT / 2, "reflect padding needs more than n_fft/2 samples");
    dim3 grid((n_frames + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA, B);
    mel_kernel<<<grid, 256, 0, stream>>>(pcm, pcm_ld, n_samples, n_frames, mel, mel_ld, mel_bs, plan->d_twiddle, plan->d_start,
                                         plan->d_count, plan->d_offset, plan->d_weights, plan->n_mels, plan->pad_reflect,
                                         plan->log_scale);
    MB_LAUNCH_CHECK();
    ++g_launch_count;
    return 0;
}

}  // namespace mb200
