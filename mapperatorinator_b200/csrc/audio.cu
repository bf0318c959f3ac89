// Audio ingest: interleaved 16-bit PCM at the file's rate -> mono float32 at the model rate, on the device.
// Replaces the CPU tail of `load_audio_file` (osuT5/osuT5/dataset/data_utils.py:80-101): pydub `set_frame_rate` = audioop.ratecv,
// `set_channels(1)` = audioop.tomono(0.5, 0.5), int16 -> float32, `normalize_audio_samples` (:132-137).  File decoding (ffmpeg) stays
// with the caller.  The arithmetic is audioop's, operation by operation, so the result is bit-identical (tests/test_audio_ingest.py):
//   ratecv is a linear interpolator with a running phase d; output frame k reads input frames n - 2 and n - 1,
//   n = 1 + ceil(k * inrate / outrate), d = (n - 1) * outrate - k * inrate (rates divided by their gcd), and
//   out = (int)(((double)prev * d + (double)cur * (outrate - d)) / (double)outrate) >> 16 on samples widened to 32 bits (s << 16);
//   tomono = floor(l * 0.5 + r * 0.5) after audioop's fbound clamp.
// HBM-bound byte work (31.7 MB in, 11.5 MB out for a 180 s stereo 44.1 kHz song): one thread per output frame, neighbouring threads
// read neighbouring input frames (a warp covers ~88 consecutive frames = 352 contiguous bytes), grid = a multiple of the SM count,
// block maximum by shuffles + one atomicMax per block for the peak.  Doubles go through the _rn intrinsics: no FMA contraction.
#include <cstdint>

#include "common.cuh"
#include "kernels.h"

namespace mb200 {
namespace {

struct AudioParams {
    const short* pcm; long long n_in; int ch;
    int irate, orate;            // divided by their gcd; irate == orate: copy
    float* out; long long n_out;
    int* peak;                   // device int, zero on entry: max |sample| of the mono int16 signal
};

__device__ __forceinline__ int ratecv_sample(int prev16, int cur16, int d, int orate) {
    const double prev = (double)(prev16 * 65536), cur = (double)(cur16 * 65536);      // GETSAMPLE32: s << 16
    const double v = __ddiv_rn(__dadd_rn(__dmul_rn(prev, (double)d), __dmul_rn(cur, (double)(orate - d))), (double)orate);
    return __double2int_rz(v) >> 16;                                                   // (int) cast, SETSAMPLE32: >> 16
}

__global__ void __launch_bounds__(256) audio_resample_kernel(AudioParams p) {
    int local_max = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < p.n_out; k += (long long)gridDim.x * blockDim.x) {
        int l, r;
        if (p.irate == p.orate) {
            l = p.pcm[k * p.ch];
            r = p.ch == 2 ? p.pcm[k * p.ch + 1] : l;
        } else {
            const long long n = 1 + (k * p.irate + p.orate - 1) / p.orate;            // input frames consumed when output k is emitted
            const int d = (int)((n - 1) * p.orate - k * p.irate);
            const short* cur = p.pcm + (n - 1) * p.ch;
            const bool has_prev = n >= 2;
            const int pl = has_prev ? cur[-p.ch] : 0, cl = cur[0];
            l = ratecv_sample(pl, cl, d, p.orate);
            r = l;
            if (p.ch == 2) {
                const int pr = has_prev ? cur[-1] : 0, cr = cur[1];
                r = ratecv_sample(pr, cr, d, p.orate);
            }
        }
        int s = l;
        if (p.ch == 2) {                                                                // audioop.tomono(0.5, 0.5)
            double v = __dadd_rn(__dmul_rn((double)l, 0.5), __dmul_rn((double)r, 0.5));
            if (v > 32767.0) v = 32767.0; else if (v < -32767.0) v = -32768.0;          // fbound
            s = __double2int_rd(v);                                                      // floor
        }
        p.out[k] = (float)s;
        local_max = max(local_max, abs(s));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    __shared__ int wmax[8];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = local_max;
    __syncthreads();
    if (threadIdx.x == 0) {
        int m = wmax[0];
        for (int w = 1; w < 8; ++w) m = max(m, wmax[w]);
        if (m > 0) atomicMax(p.peak, m);
    }
}

__global__ void __launch_bounds__(256) audio_normalize_kernel(float* x, long long n, const int* peak) {
    const int pk = *peak;
    if (pk <= 0) return;
    const float fp = (float)pk;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) x[k] = __fdiv_rn(x[k], fp);
}

long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

}  // namespace

long long audio_out_frames(long long n_in, int in_rate, int out_rate) {
    if (n_in <= 0 || in_rate <= 0 || out_rate <= 0) return 0;
    if (in_rate == out_rate) return n_in;
    const long long g = gcd_ll(in_rate, out_rate), i = in_rate / g, o = out_rate / g;
    return (n_in - 1) * o / i + 1;
}

int launch_audio_ingest(const short* pcm, long long n_frames, int channels, int in_rate, int out_rate, int normalize, float* out, int* scratch,
                        int num_sms, cudaStream_t stream) {
    MB_REQUIRE(channels == 1 || channels == 2, "audio ingest handles mono or stereo PCM (pydub set_channels(1) supports nothing else either)");
    MB_REQUIRE(in_rate > 0 && out_rate > 0, "sample rates must be positive");
    const long long n_out = audio_out_frames(n_frames, in_rate, out_rate);
    if (n_out <= 0) return 0;
    const long long g = gcd_ll(in_rate, out_rate);
    MB_REQUIRE(in_rate / g < (1 << 20) && out_rate / g < (1 << 20), "sample-rate ratio too fine for the 32-bit phase");
    MB_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(int), stream));
    AudioParams p{pcm, n_frames, channels, (int)(in_rate / g), (int)(out_rate / g), out, n_out, scratch};
    const int grid = (int)std::min<long long>((n_out + 255) / 256, (long long)std::max(1, num_sms) * 8);     // a multiple of the SM count once the signal is long enough
    audio_resample_kernel<<<grid, 256, 0, stream>>>(p);
    MB_LAUNCH_CHECK();
    ++g_launch_count;
    if (normalize) {
        audio_normalize_kernel<<<grid, 256, 0, stream>>>(out, n_out, scratch);
        MB_LAUNCH_CHECK();
        ++g_launch_count;
    }
    return 0;
}

}  // namespace mb200
