// Stage (iii): DiT forward + the ancestral sampling loop, entirely on the device.
//   reference: DiT.forward_with_cfg (osu_diffusion/utils/models.py:281-317), GaussianDiffusion.p_mean_variance / p_sample
//   (osu_diffusion/utils/diffusion/gaussian_diffusion.py:273-369, 420-467), the slider-free denoised_fn and in-paint mask of
//   DiffisionPipeline.sample_part (diffusion_pipeline.py:203-234).
// The conditioning vector silu(t_emb + y_emb) depends on the step only through t, so every adaLN modulation
// (12 blocks x 6 vectors + final 2) is computed for ALL steps up front with a handful of GEMMs; a step is then
//   first-layer embed -> 12 x [LN-modulate, qkv GEMM, band attention, gated out_proj GEMM, LN-modulate, fc1 GEMM + tanh-GELU,
//   gated fc2 GEMM] -> final LN-modulate -> 4-channel GEMM -> fused CFG-mix + learned-range variance + x0 + in-paint +
//   clamp(-2,2) + posterior mean + noise update, with no host round trip (the reference does ~25 tiny launches and several
//   numpy->tensor table gathers per step).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/mapperatorinator_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace mb200;

namespace {

struct DevBufD {
    void* p = nullptr; size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        MB_CUDA_CHECK(cudaMalloc(&p, need));
        bytes = need;
        return 0;
    }
    float* f() const { return reinterpret_cast<float*>(p); }
    ~DevBufD() { if (p) cudaFree(p); }
};

#define MB_TRY(expr) do { int _s = (expr); if (_s) return _s; } while (0)

// A0[(n,t), :] = [cos|sin (x0*512*f) (128) , cos|sin (x1*512*f) (128) , c[n, :, t] (E)]   (FirstLayer.forward, models.py:204-209)
__global__ void dit_embed_kernel(const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ freqs, int N, int T, int C,
                                 int E, int FD, float* __restrict__ a0) {
    const int t = blockIdx.x, n = blockIdx.y;
    const int half_n = N / 2 > 0 ? N / 2 : 1;
    const int src = n % half_n;                      // forward_with_cfg feeds cat([half, half]) (models.py:306-307)
    const int K = C * FD + E, halfd = FD / 2;
    float* row = a0 + ((long long)n * T + t) * K;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        float v;
        if (j < C * FD) {
            int ch = j / FD, k = j - ch * FD;
            float xv = x[((long long)src * C + ch) * T + t] * 512.0f;
            float arg = xv * freqs[k < halfd ? k : k - halfd];
            v = k < halfd ? cosf(arg) : sinf(arg);
        } else {
            v = c[((long long)n * E + (j - C * FD)) * T + t];
        }
        row[j] = v;
    }
}

// timestep_embedding(t, 256): row r -> [cos(t f) | sin(t f)]
__global__ void dit_temb_kernel(const float* __restrict__ tvals, const float* __restrict__ freqs, int FD, float* __restrict__ out) {
    const int r = blockIdx.x, halfd = FD / 2;
    const float t = tvals[r];
    for (int j = threadIdx.x; j < FD; j += blockDim.x) {
        float arg = t * freqs[j < halfd ? j : j - halfd];
        out[(long long)r * FD + j] = j < halfd ? cosf(arg) : sinf(arg);
    }
}

// b[r, :] = silu(te[r, :] + ye[r % N, :])   (DiT.forward: b = t + y, then adaLN_modulation[0] = SiLU)
__global__ void dit_cond_kernel(const float* __restrict__ te, const float* __restrict__ ye, int N, int d, float* __restrict__ b) {
    const int r = blockIdx.x;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        float v = te[(long long)r * d + j] + ye[(long long)(r % N) * d + j];
        b[(long long)r * d + j] = v / (1.0f + expf(-v));
    }
}

// forward_with_cfg output assembly: out[n, ch, t] (models.py:312-317) from out4[(n,t), 4]
__global__ void dit_cfg_out_kernel(const float* __restrict__ o4, int N, int T, float cfg_scale, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * T) return;
    const int n = idx / T, t = idx - n * T, half_n = N / 2;
    for (int ch = 0; ch < 2; ++ch) {
        float cond = o4[((long long)(n % half_n) * T + t) * 4 + ch];
        float unc = o4[((long long)(half_n + n % half_n) * T + t) * 4 + ch];
        out[((long long)n * 4 + ch) * T + t] = unc + cfg_scale * (cond - unc);
        out[((long long)n * 4 + 2 + ch) * T + t] = o4[((long long)n * T + t) * 4 + 2 + ch];
    }
}

struct StepConst { float sqrt_recip, sqrt_recipm1, min_log, max_log, coef1, coef2, nonzero; };

// one p_sample update (gaussian_diffusion.py:312-358, 454-466) for every (n, ch, t)
__global__ void dit_update_kernel(const float* __restrict__ o4, const float* __restrict__ x, const float* __restrict__ z,
                                  const unsigned char* __restrict__ inpaint, const float* __restrict__ noise, int N, int T, float cfg_scale,
                                  StepConst sc, float* __restrict__ x_new) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * 2 * T) return;
    const int n = idx / (2 * T), rem = idx - n * 2 * T, ch = rem / T, t = rem - ch * T, half_n = N / 2;
    const float cond = o4[((long long)(n % half_n) * T + t) * 4 + ch];
    const float unc = o4[((long long)(half_n + n % half_n) * T + t) * 4 + ch];
    const float eps = unc + cfg_scale * (cond - unc);
    const float v = o4[((long long)n * T + t) * 4 + 2 + ch];
    const float frac = (v + 1.0f) / 2.0f;
    const float logvar = frac * sc.max_log + (1.0f - frac) * sc.min_log;
    const float xt = x[idx];
    float x0 = sc.sqrt_recip * xt - sc.sqrt_recipm1 * eps;
    if (inpaint && !inpaint[idx]) x0 = z[idx];
    x0 = fminf(fmaxf(x0, -2.0f), 2.0f);
    const float mean = sc.coef1 * x0 + sc.coef2 * xt;
    x_new[idx] = mean + sc.nonzero * expf(0.5f * logvar) * noise[idx];
}

}  // namespace

struct mb200_dit {
    mb200_dit_config cfg;
    std::unordered_map<std::string, std::vector<float>> host_w;
    bool finalized = false;
    DevBufD arena;
    std::unordered_map<std::string, const float*> w;   // device pointers by reference name
    const float *pos_freqs = nullptr, *t_freqs = nullptr;
    DevBufD a0, x, h, qkv, att, ffn, o4, temb, te1, te, ye1, ye, bcond, mods, fmod, tvals, state0, state1;
    int mod_steps = 0;
    std::vector<const float*> tc_weights;
};

extern "C" int mb200_dit_create(mb200_dit** out, const mb200_dit_config* cfg) {
    MB_REQUIRE(out && cfg, "null argument");
    MB_REQUIRE(cfg->hidden == cfg->heads * 64, "kernels are specialised for head_dim 64 (DiT-B: 768 / 12)");
    MB_REQUIRE(cfg->hidden % 128 == 0 && cfg->hidden <= 1024, "hidden must be a multiple of 128 and <= 1024");
    MB_REQUIRE(cfg->in_channels == 2, "the position DiT has 2 input channels");
    MB_REQUIRE(cfg->class_size % 4 == 0 && cfg->context_size % 4 == 0, "class_size / context_size must be multiples of 4");
    mb200_dit* d = new mb200_dit();
    d->cfg = *cfg;
    *out = d;
    return 0;
}

extern "C" void mb200_dit_destroy(mb200_dit* d) {
    if (!d) return;
    for (const float* w : d->tc_weights) tc_unregister_weight(w);
    delete d;
}

extern "C" int mb200_dit_set_weight(mb200_dit* d, const char* name, const float* data, int64_t numel) {
    MB_REQUIRE(d && name && data && !d->finalized, "bad argument / state");
    d->host_w[name] = std::vector<float>(data, data + numel);
    return 0;
}

extern "C" int mb200_dit_finalize(mb200_dit* dd) {
    MB_REQUIRE(dd && !dd->finalized, "bad state");
    const auto& c = dd->cfg;
    const int d = c.hidden;
    std::vector<std::string> names = {
        "context_embedder.mlp.0.weight", "context_embedder.mlp.0.bias", "t_embedder.mlp.0.weight", "t_embedder.mlp.0.bias",
        "t_embedder.mlp.2.weight", "t_embedder.mlp.2.bias", "y_embedder.class_embedding.0.weight", "y_embedder.class_embedding.0.bias",
        "y_embedder.class_embedding.2.weight", "y_embedder.class_embedding.2.bias", "final_layer.adaLN_modulation.1.weight",
        "final_layer.adaLN_modulation.1.bias", "final_layer.linear.weight", "final_layer.linear.bias"};
    for (int i = 0; i < c.depth; ++i) {
        std::string p = "blocks." + std::to_string(i) + ".";
        for (const char* s : {"attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias", "mlp.fc1.weight",
                              "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "adaLN_modulation.1.weight", "adaLN_modulation.1.bias"})
            names.push_back(p + s);
    }
    std::vector<float> pack;
    std::unordered_map<std::string, size_t> offs, sizes;
    for (const auto& n : names) {
        MB_REQUIRE(dd->host_w.count(n) == 1, "missing DiT weight " + n);
        std::vector<float> v = dd->host_w.at(n);
        // nn.MultiheadAttention scales q by head_dim^-0.5 = 1/8 (a power of two): fold it into the q rows, bit-exact
        if (n.find("attn.in_proj_weight") != std::string::npos) for (size_t i = 0; i < (size_t)d * d; ++i) v[i] *= 0.125f;
        if (n.find("attn.in_proj_bias") != std::string::npos) for (int i = 0; i < d; ++i) v[i] *= 0.125f;
        size_t off = (pack.size() + 63) & ~size_t(63);
        pack.resize(off + v.size());
        std::copy(v.begin(), v.end(), pack.begin() + off);
        offs[n] = off; sizes[n] = v.size();
    }
    // sinusoid frequency tables, same fp32 chain as positional_embedding.timestep_embedding (:40-46)
    auto freqs = [&](int dim) {
        std::vector<float> f(dim / 2);
        const float neg_log = (float)(-std::log(10000.0));
        for (int k = 0; k < dim / 2; ++k) f[k] = std::exp(neg_log * (float)k / (float)(dim / 2));
        return f;
    };
    for (auto pr : {std::make_pair(std::string("__pos_freqs"), c.pos_freq_dim), std::make_pair(std::string("__t_freqs"), c.t_freq_dim)}) {
        std::vector<float> v = freqs(pr.second);
        size_t off = (pack.size() + 63) & ~size_t(63);
        pack.resize(off + v.size());
        std::copy(v.begin(), v.end(), pack.begin() + off);
        offs[pr.first] = off;
    }
    MB_TRY(dd->arena.ensure(pack.size() * 4));
    MB_CUDA_CHECK(cudaMemcpy(dd->arena.p, pack.data(), pack.size() * 4, cudaMemcpyHostToDevice));
    for (auto& kv : offs) dd->w[kv.first] = dd->arena.f() + kv.second;
    // tf32 "lo" mirrors for the tensor-core GEMMs (every 2-D weight of the blocks and the embedders)
    for (const auto& n : names)
        if (n.find("weight") != std::string::npos && sizes.at(n) >= (size_t)64 * 32) {
            dd->tc_weights.push_back(dd->w[n]);
            MB_TRY(tc_register_weight(dd->w[n], (long long)sizes.at(n)));
        }
    MB_CUDA_CHECK(cudaDeviceSynchronize());
    dd->pos_freqs = dd->w["__pos_freqs"]; dd->t_freqs = dd->w["__t_freqs"];
    MB_REQUIRE(dd->host_w.at("context_embedder.mlp.0.weight").size() == (size_t)d * (c.in_channels * c.pos_freq_dim + c.context_size),
               "context_embedder shape mismatch");
    dd->host_w.clear();
    dd->finalized = true;
    return 0;
}

namespace {

GemmParams gb(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc, const float* bias, int M, int N, int K) {
    GemmParams g{};
    g.A = plain_map(A, lda); g.W = W; g.ldw = ldw; g.C = plain_map(C, ldc); g.bias = bias; g.act = ACT_NONE; g.alpha = 1.f;
    g.gate = nullptr; g.gate_ld = 0; g.gate_rpb = 1; g.R = RowMap{nullptr, 0, 0, 0}; g.M = M; g.N = N; g.K = K;
    return g;
}

int ln_mod(const float* x, float* y, const float* shift, const float* scale, long long mod_ld, int T, int rows, int d, cudaStream_t st) {
    LayerNormParams p{};
    p.x = x; p.ldx = d; p.y = y; p.ldy = d; p.weight = nullptr; p.bias = nullptr; p.shift = shift; p.scale = scale; p.mod_ld = mod_ld;
    p.rows_per_batch = T; p.rows = rows; p.dim = d; p.eps = 1e-6f;
    return launch_layernorm(p, st);
}

// modulation vectors for `steps` timesteps: mods[l][step*N + n][6d], fmod[step*N + n][2d]
int prepare_conditioning(mb200_dit* dd, const float* tvals_host, int steps, int N, const float* y, cudaStream_t st) {
    const auto& c = dd->cfg;
    const int d = c.hidden, RS = steps * N;
    MB_TRY(dd->tvals.ensure((size_t)RS * 4)); MB_TRY(dd->temb.ensure((size_t)RS * c.t_freq_dim * 4));
    MB_TRY(dd->te1.ensure((size_t)RS * d * 4)); MB_TRY(dd->te.ensure((size_t)RS * d * 4));
    MB_TRY(dd->ye1.ensure((size_t)N * d * 4)); MB_TRY(dd->ye.ensure((size_t)N * d * 4));
    MB_TRY(dd->bcond.ensure((size_t)RS * d * 4));
    MB_TRY(dd->mods.ensure((size_t)c.depth * RS * 6 * d * 4)); MB_TRY(dd->fmod.ensure((size_t)RS * 2 * d * 4));
    MB_CUDA_CHECK(cudaMemcpyAsync(dd->tvals.p, tvals_host, (size_t)RS * 4, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    dit_temb_kernel<<<RS, 128, 0, st>>>(dd->tvals.f(), dd->t_freqs, c.t_freq_dim, dd->temb.f());
    MB_LAUNCH_CHECK();
    auto& w = dd->w;
    {
        GemmParams g = gb(dd->temb.f(), c.t_freq_dim, w["t_embedder.mlp.0.weight"], c.t_freq_dim, dd->te1.f(), d, w["t_embedder.mlp.0.bias"], RS, d, c.t_freq_dim);
        g.act = ACT_SILU;
        MB_TRY(launch_gemm(g, st));
        MB_TRY(launch_gemm(gb(dd->te1.f(), d, w["t_embedder.mlp.2.weight"], d, dd->te.f(), d, w["t_embedder.mlp.2.bias"], RS, d, d), st));
        GemmParams gy = gb(y, c.class_size, w["y_embedder.class_embedding.0.weight"], c.class_size, dd->ye1.f(), d,
                           w["y_embedder.class_embedding.0.bias"], N, d, c.class_size);
        gy.act = ACT_SILU;
        MB_TRY(launch_gemm(gy, st));
        MB_TRY(launch_gemm(gb(dd->ye1.f(), d, w["y_embedder.class_embedding.2.weight"], d, dd->ye.f(), d, w["y_embedder.class_embedding.2.bias"], N, d, d), st));
    }
    dit_cond_kernel<<<RS, 128, 0, st>>>(dd->te.f(), dd->ye.f(), N, d, dd->bcond.f());
    MB_LAUNCH_CHECK();
    for (int l = 0; l < c.depth; ++l) {
        std::string p = "blocks." + std::to_string(l) + ".adaLN_modulation.1.";
        MB_TRY(launch_gemm(gb(dd->bcond.f(), d, w[p + "weight"], d, dd->mods.f() + (size_t)l * RS * 6 * d, 6 * d, w[p + "bias"], RS, 6 * d, d), st));
    }
    MB_TRY(launch_gemm(gb(dd->bcond.f(), d, w["final_layer.adaLN_modulation.1.weight"], d, dd->fmod.f(), 2 * d,
                          w["final_layer.adaLN_modulation.1.bias"], RS, 2 * d, d), st));
    dd->mod_steps = steps;
    return 0;
}

int ensure_work(mb200_dit* dd, int N, int T) {
    const auto& c = dd->cfg;
    const int d = c.hidden;
    const size_t R = (size_t)N * T;
    const int K0 = c.in_channels * c.pos_freq_dim + c.context_size;
    MB_TRY(dd->a0.ensure(R * K0 * 4)); MB_TRY(dd->x.ensure(R * d * 4)); MB_TRY(dd->h.ensure(R * d * 4));
    MB_TRY(dd->qkv.ensure(R * 3 * d * 4)); MB_TRY(dd->att.ensure(R * d * 4)); MB_TRY(dd->ffn.ensure(R * d * c.mlp_ratio * 4));
    MB_TRY(dd->o4.ensure(R * 4 * 4));
    return 0;
}

// one DiT forward at conditioning row block `step` -> o4[(n,t), 4]
int dit_forward(mb200_dit* dd, const float* xstate, const float* cctx, int N, int T, int step, int steps_total, const mb200_dit_mask* mask,
                cudaStream_t st) {
    const auto& c = dd->cfg;
    const int d = c.hidden, f = d * c.mlp_ratio, R = N * T, RS = steps_total * N;
    const int K0 = c.in_channels * c.pos_freq_dim + c.context_size;
    auto& w = dd->w;
    dit_embed_kernel<<<dim3(T, N), 128, 0, st>>>(xstate, cctx, dd->pos_freqs, N, T, c.in_channels, c.context_size, c.pos_freq_dim, dd->a0.f());
    MB_LAUNCH_CHECK();
    MB_TRY(launch_gemm(gb(dd->a0.f(), K0, w["context_embedder.mlp.0.weight"], K0, dd->x.f(), d, w["context_embedder.mlp.0.bias"], R, d, K0), st));
    for (int l = 0; l < c.depth; ++l) {
        std::string p = "blocks." + std::to_string(l) + ".";
        const float* mod = dd->mods.f() + ((size_t)l * RS + (size_t)step * N) * 6 * d;   // rows n = 0..N-1 of this step
        MB_TRY(ln_mod(dd->x.f(), dd->h.f(), mod + 0, mod + d, 6 * d, T, R, d, st));
        MB_TRY(launch_gemm(gb(dd->h.f(), d, w[p + "attn.in_proj_weight"], d, dd->qkv.f(), 3 * d, w[p + "attn.in_proj_bias"], R, 3 * d, d), st));
        AttentionParams a{};
        a.q = dd->qkv.f(); a.q_ld = 3 * d; a.q_bs = (long long)T * 3 * d;
        a.k = dd->qkv.f() + d; a.k_ld = 3 * d; a.k_bs = a.q_bs;
        a.v = dd->qkv.f() + 2 * d; a.v_ld = 3 * d; a.v_bs = a.q_bs;
        a.o = dd->att.f(); a.o_ld = d; a.o_bs = (long long)T * d;
        a.B = N; a.H = c.heads; a.Tq = T; a.Tk = T; a.scale = 1.f;
        a.mask_mode = mask ? mask->mask_mode : MASK_NONE; a.band = mask ? mask->band : 0; a.dense = mask ? mask->dense_mask : nullptr;
        MB_TRY(launch_attention(a, st));
        {
            GemmParams g = gb(dd->att.f(), d, w[p + "attn.out_proj.weight"], d, dd->x.f(), d, w[p + "attn.out_proj.bias"], R, d, d);
            g.gate = mod + 2 * d; g.gate_ld = 6 * d; g.gate_rpb = T; g.R = plain_map(dd->x.f(), d);
            MB_TRY(launch_gemm(g, st));
        }
        MB_TRY(ln_mod(dd->x.f(), dd->h.f(), mod + 3 * d, mod + 4 * d, 6 * d, T, R, d, st));
        {
            GemmParams g = gb(dd->h.f(), d, w[p + "mlp.fc1.weight"], d, dd->ffn.f(), f, w[p + "mlp.fc1.bias"], R, f, d);
            g.act = ACT_GELU_TANH;
            MB_TRY(launch_gemm(g, st));
        }
        {
            GemmParams g = gb(dd->ffn.f(), f, w[p + "mlp.fc2.weight"], f, dd->x.f(), d, w[p + "mlp.fc2.bias"], R, d, f);
            g.gate = mod + 5 * d; g.gate_ld = 6 * d; g.gate_rpb = T; g.R = plain_map(dd->x.f(), d);
            MB_TRY(launch_gemm(g, st));
        }
    }
    const float* fm = dd->fmod.f() + (size_t)step * N * 2 * d;
    MB_TRY(ln_mod(dd->x.f(), dd->h.f(), fm, fm + d, 2 * d, T, R, d, st));
    MB_TRY(launch_gemm(gb(dd->h.f(), d, w["final_layer.linear.weight"], d, dd->o4.f(), 4, w["final_layer.linear.bias"], R, 4, d), st));
    return 0;
}

int check_shapes(mb200_dit* d, int N, int T) {
    MB_REQUIRE(d && d->finalized, "DiT not finalized");
    MB_REQUIRE(N >= 2 && N % 2 == 0 && N <= d->cfg.max_batch, "forward_with_cfg needs an even batch (cond | uncond) within max_batch");
    MB_REQUIRE(T >= 1 && T <= d->cfg.max_seq_len, "sequence longer than max_seq_len");
    return 0;
}

}  // namespace

extern "C" int mb200_dit_forward_with_cfg(mb200_dit* d, const float* x, const int32_t* t, const float* c, const float* y, int32_t N, int32_t T,
                                          float cfg_scale, const mb200_dit_mask* mask, float* out, void* stream) {
    MB_TRY(check_shapes(d, N, T));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<float> tv(N);
    for (int i = 0; i < N; ++i) tv[i] = (float)t[i];
    MB_TRY(ensure_work(d, N, T));
    MB_TRY(prepare_conditioning(d, tv.data(), 1, N, y, st));
    MB_TRY(dit_forward(d, x, c, N, T, 0, 1, mask, st));
    dit_cfg_out_kernel<<<(N * T + 255) / 256, 256, 0, st>>>(d->o4.f(), N, T, cfg_scale, out);
    MB_LAUNCH_CHECK();
    return 0;
}

extern "C" int mb200_dit_sample_loop(mb200_dit* d, const float* z, const float* c, const float* y, const uint8_t* inpaint, int32_t N, int32_t T,
                                     float cfg_scale, const mb200_dit_mask* mask, const float* schedule, int32_t steps, const float* noise,
                                     float* out, void* stream) {
    MB_TRY(check_shapes(d, N, T));
    MB_REQUIRE(z && c && y && schedule && noise && out && steps >= 1, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<float> tv((size_t)steps * N);
    for (int k = 0; k < steps; ++k)
        for (int n = 0; n < N; ++n) tv[(size_t)k * N + n] = schedule[(size_t)k * 8 + 0];
    MB_TRY(ensure_work(d, N, T));
    const size_t state_bytes = (size_t)N * 2 * T * 4;
    MB_TRY(d->state0.ensure(state_bytes)); MB_TRY(d->state1.ensure(state_bytes));
    MB_TRY(prepare_conditioning(d, tv.data(), steps, N, y, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(d->state0.p, z, state_bytes, cudaMemcpyDeviceToDevice, st));
    float* cur = d->state0.f();
    float* nxt = d->state1.f();
    const int total = N * 2 * T;
    for (int k = 0; k < steps; ++k) {
        MB_TRY(dit_forward(d, cur, c, N, T, k, steps, mask, st));
        const float* s = schedule + (size_t)k * 8;
        StepConst sc{s[1], s[2], s[3], s[4], s[5], s[6], s[7]};
        float* dst = (k == steps - 1) ? out : nxt;
        dit_update_kernel<<<(total + 255) / 256, 256, 0, st>>>(d->o4.f(), cur, z, inpaint, noise + (size_t)k * total, N, T, cfg_scale, sc, dst);
        MB_LAUNCH_CHECK();
        std::swap(cur, nxt);
    }
    return 0;
}
