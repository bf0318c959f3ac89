// Host runtime of the osuT5 stage: weights, resident encoder/cross-KV slots, self-KV arena, encoder pass, decoder
// prefill, CUDA-graph token loop.  Mirrors what `server.model_generate` drives through HF `generate`
// (osuT5/osuT5/inference/server.py:83-156) — see include/mapperatorinator_b200.h for the boundary.
#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/mapperatorinator_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace mb200;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need, bool zero = false) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        MB_CUDA_CHECK(cudaMalloc(&p, need));
        bytes = need;
        if (zero) MB_CUDA_CHECK(cudaMemset(p, 0, need));
        return 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
    ~DevBuf() { if (p) cudaFree(p); }
};

struct LayerW {   // device pointers into the weight arena
    const float *ln1_w, *ln1_b, *wqkv, *bqkv, *wo, *bo;
    const float *ln2_w, *ln2_b, *wq_c, *bq_c, *wkv_c, *bkv_c, *wo_c, *bo_c;   // decoder cross attention only
    const float *ln3_w, *ln3_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};

}  // namespace

struct mb200_model {
    mb200_model_config cfg;
    MelPlan* mel = nullptr;
    std::unordered_map<std::string, std::vector<float>> host_w;   // until finalize()
    bool finalized = false;

    DevBuf arena;                       // all packed weights
    const float *emb_w = nullptr, *emb_b = nullptr, *conv1_w = nullptr, *conv1_b = nullptr, *conv2_w = nullptr, *conv2_b = nullptr,
                *enc_pos = nullptr, *enc_ln_w = nullptr, *enc_ln_b = nullptr;
    const float *tok_emb = nullptr, *dec_pos = nullptr, *dec_ln_w = nullptr, *dec_ln_b = nullptr, *proj_out = nullptr;
    std::vector<LayerW> enc, dec;

    DevBuf cross_kv;                    // [dec_layers][max_windows][Ts][2d]   (k | v per token)
    DevBuf self_kv;                     // [dec_layers][max_rows][tgt][2d]
    int max_rows = 0;

    // encoder workspaces (chunk of windows)
    int enc_chunk = 0;
    DevBuf w_mel, w_embpad, w_c1pad, w_x, w_h, w_qkv, w_attn, w_ffn;
    // decoder workspaces
    DevBuf p_x, p_h, p_q, p_attn, p_ffn;          // prefill, sized rows * P
    DevBuf d_x, d_q, d_h, d_parto, d_partml, d_logits;   // decode step
    DevBuf d_attn, d_ticket;            // merged attention heads [rows, d] and the per-(row, head) arrival counters of the split merge
    DevBuf g_state, g_cfg, g_vflags, g_ids, g_prefill_ids, g_keyvalid, g_leftpad, g_rowslot, g_finished, g_lastts, g_lastscores;
    int* h_flag = nullptr;              // pinned
    // graph keys carry B as well as rows: the captured sample kernel's grid is dim3(B), and a CFG call (B = 1, rows = 2) must
    // never replay the graph of a plain batch-2 call (B = 2, rows = 2)
    std::map<std::tuple<int, int, int>, cudaGraphExec_t> graphs;
    std::map<std::tuple<int, int, int>, long long> graph_nodes;   // (rows, B, n_splits_self) -> token-step graph
    bool use_pdl = false;
    cudaStream_t cap_stream = nullptr;
    std::map<std::tuple<int, int, int, int>, int> prefill_seen;                                   // (rows, B, P, position rule)
    std::map<std::tuple<int, int, int, int>, std::pair<cudaGraphExec_t, long long>> prefill_graphs;   // -> graph + node count
    // persistent megakernel path
    int enc_graph = 1;                  // replay single-window encodes as a CUDA graph (option "enc_graph")
    int trace_cta = 0;                  // CTA whose phases the dataflow megakernel's TRACE instantiation stamps
    int use_mega = 2;                   // 0 = CUDA-graph replay per token, 1 = grid-barrier megakernel, 2 = dataflow (tagged-pair) megakernel
    DevBuf ll_arena;                    // exchange buffers of the dataflow megakernel (rows <= 2)
    MegaLL ll{};
    size_t ll_bytes = 0;
    std::map<std::tuple<int, int, int>, std::pair<DevBuf*, int>> mega2_phases;   // (rows, B, n_splits_self) -> device phase table
    int num_sms = 0;                    // 0 = no cooperative launch -> no megakernel
    int num_sms_phys = 0;
    DevBuf g_megasync;                  // [0] grid-barrier counter, [8] error flag
    DevBuf mega_trace;                  // optional per-phase clock64 stamps (option "mega_trace")
    cudaEvent_t mega_ev[2] = {nullptr, nullptr};
    double mega_ms = 0.0; long long mega_launches = 0, mega_tokens = 0;   // CUDA-event time of every megakernel launch
    std::map<std::pair<int, int>, std::pair<DevBuf*, int>> mega_phases;   // (rows, n_splits_self) -> device phase table
    DevBuf w_pcm;                       // single-window encode: engine-owned copy of the window's PCM (the captured graph reads it)
    std::map<int, std::pair<cudaGraphExec_t, long long>> enc_graphs;   // slot -> captured single-window encode (the drop-in per-call pattern), node count
    std::map<int, int> enc_seen;
    AttnCtx attn;                       // tensor-core attention scratch of the encoder (head-major tf32 copies of q | k | v^T)
    GemmCtx gemm;                       // this engine's GEMM scratch: split-K planes, tf32 activation copies, weight mirrors, error flag

    int d() const { return cfg.d_model; }
    int Ts() const { return cfg.src_seq_len / 2; }
    size_t cross_layer_stride() const { return (size_t)cfg.max_windows * Ts() * 2 * d(); }
    size_t self_layer_stride() const { return (size_t)max_rows * cfg.tgt_seq_len * 2 * d(); }
};

namespace {

std::vector<std::string> required_names(const mb200_model_config& c) {
    std::vector<std::string> n = {
        "encoder_embedder.weight", "encoder_embedder.bias", "decoder_embedder.weight",
        "transformer.model.encoder.conv1.weight", "transformer.model.encoder.conv1.bias",
        "transformer.model.encoder.conv2.weight", "transformer.model.encoder.conv2.bias",
        "transformer.model.encoder.embed_positions.weight",
        "transformer.model.encoder.layer_norm.weight", "transformer.model.encoder.layer_norm.bias",
        "transformer.model.decoder.embed_positions.weight",
        "transformer.model.decoder.layer_norm.weight", "transformer.model.decoder.layer_norm.bias",
        "transformer.proj_out.weight"};
    auto attn = [&](const std::string& p) {
        n.push_back(p + "q_proj.weight"); n.push_back(p + "q_proj.bias"); n.push_back(p + "k_proj.weight");
        n.push_back(p + "v_proj.weight"); n.push_back(p + "v_proj.bias");
        n.push_back(p + "out_proj.weight"); n.push_back(p + "out_proj.bias");
    };
    auto ln = [&](const std::string& p) { n.push_back(p + "weight"); n.push_back(p + "bias"); };
    for (int side = 0; side < 2; ++side) {
        int L = side == 0 ? c.encoder_layers : c.decoder_layers;
        for (int i = 0; i < L; ++i) {
            std::string p = std::string("transformer.model.") + (side == 0 ? "encoder" : "decoder") + ".layers." + std::to_string(i) + ".";
            attn(p + "self_attn."); ln(p + "self_attn_layer_norm.");
            if (side == 1) { attn(p + "encoder_attn."); ln(p + "encoder_attn_layer_norm."); }
            n.push_back(p + "fc1.weight"); n.push_back(p + "fc1.bias"); n.push_back(p + "fc2.weight"); n.push_back(p + "fc2.bias");
            ln(p + "final_layer_norm.");
        }
    }
    return n;
}

GemmParams gemm_base(const RowMap& A, const float* W, long long ldw, const RowMap& C, const float* bias, int M, int N, int K) {
    GemmParams g{};
    g.A = A; g.W = W; g.ldw = ldw; g.C = C; g.bias = bias; g.act = ACT_NONE; g.alpha = 1.f;
    g.gate = nullptr; g.gate_ld = 0; g.gate_rpb = 1; g.R = RowMap{nullptr, 0, 0, 0};
    g.M = M; g.N = N; g.K = K;
    return g;
}

#define MB_TRY(expr) do { int _s = (expr); if (_s) return _s; } while (0)

int layernorm(const float* x, float* y, const float* w, const float* b, int rows, int dim, float eps, cudaStream_t st) {
    LayerNormParams p{};
    p.x = x; p.ldx = dim; p.y = y; p.ldy = dim; p.weight = w; p.bias = b; p.shift = nullptr; p.scale = nullptr; p.mod_ld = 0;
    p.rows_per_batch = 1; p.rows = rows; p.dim = dim; p.eps = eps;
    return launch_layernorm(p, st);
}

}  // namespace

// =====================================================================================================================
extern "C" int mb200_model_create(mb200_model** out, const mb200_model_config* cfg, const float* mel_basis_host) {
    MB_REQUIRE(cfg && out, "null argument");
    MB_REQUIRE(cfg->d_model == cfg->heads * 64, "kernels are specialised for head_dim 64 (whisper-small: 768 / 12)");
    MB_REQUIRE(cfg->d_model % 128 == 0 && cfg->d_model <= 1024, "d_model must be a multiple of 128 and <= 1024");
    MB_REQUIRE(cfg->vocab_size_out <= 4096, "fused sampling kernel holds the vocabulary in shared memory (<= 4096 ids)");
    MB_REQUIRE(cfg->mel.n_mels % 4 == 0 && cfg->ffn_dim % 4 == 0, "n_mels / ffn_dim must be multiples of 4");
    mb200_model* m = new mb200_model();
    m->cfg = *cfg;
    int s = mel_plan_create(&m->mel, cfg->mel.n_fft, cfg->mel.hop_length, cfg->mel.n_mels, cfg->mel.pad_reflect, cfg->mel.log_scale,
                            mel_basis_host);
    if (s) { delete m; return s; }
    if (cudaMallocHost(&m->h_flag, 64) != cudaSuccess) { delete m; set_last_error("cudaMallocHost failed"); return 1; }
    {
        int dev = 0, coop = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev);
        m->num_sms_phys = m->num_sms;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (!coop) m->num_sms = 0;
    }
    *out = m;
    return 0;
}

extern "C" void mb200_model_destroy(mb200_model* m) {
    if (!m) return;
    for (auto& g : m->graphs) cudaGraphExecDestroy(g.second);
    for (auto& g : m->prefill_graphs) cudaGraphExecDestroy(g.second.first);
    for (auto& g : m->enc_graphs) cudaGraphExecDestroy(g.second.first);
    m->gemm.destroy();
    m->attn.destroy();
    for (auto& kv : m->mega_phases) delete kv.second.first;
    for (auto& kv : m->mega2_phases) delete kv.second.first;
    for (auto& e : m->mega_ev) if (e) cudaEventDestroy(e);
    if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
    mel_plan_destroy(m->mel);
    if (m->h_flag) cudaFreeHost(m->h_flag);
    delete m;
}

extern "C" int mb200_model_set_weight(mb200_model* m, const char* name, const float* data, int64_t numel) {
    MB_REQUIRE(m && name && data, "null argument");
    MB_REQUIRE(!m->finalized, "model already finalized");
    m->host_w[name] = std::vector<float>(data, data + numel);
    return 0;
}

extern "C" int mb200_model_finalize(mb200_model* m) {
    MB_REQUIRE(m && !m->finalized, "bad model state");
    const auto& c = m->cfg;
    const int d = c.d_model, f = c.ffn_dim;
    for (const auto& n : required_names(c))
        MB_REQUIRE(m->host_w.count(n) == 1, std::string("missing weight ") + n);
    auto W = [&](const std::string& n) -> const std::vector<float>& { return m->host_w.at(n); };
    auto expect = [&](const std::string& n, size_t numel) -> int {
        MB_REQUIRE(W(n).size() == numel, "wrong element count for " + n);
        return 0;
    };
    MB_TRY(expect("encoder_embedder.weight", (size_t)d * c.mel.n_mels));
    MB_TRY(expect("decoder_embedder.weight", (size_t)c.vocab_size_in * d));
    MB_TRY(expect("transformer.proj_out.weight", (size_t)c.vocab_size_out * d));
    MB_TRY(expect("transformer.model.encoder.conv1.weight", (size_t)d * d * 3));
    MB_TRY(expect("transformer.model.encoder.embed_positions.weight", (size_t)(c.src_seq_len / 2) * d));
    MB_TRY(expect("transformer.model.decoder.embed_positions.weight", (size_t)c.tgt_seq_len * d));

    // ---- pack on the host ----
    std::vector<float> pack;
    auto push = [&](const std::vector<float>& v) -> size_t {
        size_t off = (pack.size() + 63) & ~size_t(63);   // 256-byte alignment
        pack.resize(off + v.size());
        std::copy(v.begin(), v.end(), pack.begin() + off);
        return off;
    };
    auto conv_pack = [&](const std::vector<float>& w) {   // [co][ci][3] -> [co][tap*C + ci]
        std::vector<float> o((size_t)d * 3 * d);
        for (int co = 0; co < d; ++co)
            for (int ci = 0; ci < d; ++ci)
                for (int j = 0; j < 3; ++j) o[(size_t)co * 3 * d + (size_t)j * d + ci] = w[((size_t)co * d + ci) * 3 + j];
        return o;
    };
    const float qs = 0.125f;   // head_dim^-0.5 for head_dim 64: a power of two, so folding it into Wq/bq is bit-exact
    auto qkv_pack = [&](const std::string& p, std::vector<float>& wq, std::vector<float>& bq) {
        wq.assign((size_t)3 * d * d, 0.f); bq.assign((size_t)3 * d, 0.f);
        const auto &q = W(p + "q_proj.weight"), &k = W(p + "k_proj.weight"), &v = W(p + "v_proj.weight");
        const auto &qb = W(p + "q_proj.bias"), &vb = W(p + "v_proj.bias");
        for (size_t i = 0; i < (size_t)d * d; ++i) { wq[i] = q[i] * qs; wq[(size_t)d * d + i] = k[i]; wq[(size_t)2 * d * d + i] = v[i]; }
        for (int i = 0; i < d; ++i) { bq[i] = qb[i] * qs; bq[2 * d + i] = vb[i]; }
    };
    struct Off { size_t v[20]; };
    std::vector<Off> eo(c.encoder_layers), dof(c.decoder_layers);
    size_t o_emb_w = push(W("encoder_embedder.weight")), o_emb_b = push(W("encoder_embedder.bias"));
    size_t o_c1w = push(conv_pack(W("transformer.model.encoder.conv1.weight"))), o_c1b = push(W("transformer.model.encoder.conv1.bias"));
    size_t o_c2w = push(conv_pack(W("transformer.model.encoder.conv2.weight"))), o_c2b = push(W("transformer.model.encoder.conv2.bias"));
    size_t o_epos = push(W("transformer.model.encoder.embed_positions.weight"));
    size_t o_elnw = push(W("transformer.model.encoder.layer_norm.weight")), o_elnb = push(W("transformer.model.encoder.layer_norm.bias"));
    size_t o_tok = push(W("decoder_embedder.weight")), o_dpos = push(W("transformer.model.decoder.embed_positions.weight"));
    size_t o_dlnw = push(W("transformer.model.decoder.layer_norm.weight")), o_dlnb = push(W("transformer.model.decoder.layer_norm.bias"));
    size_t o_proj = push(W("transformer.proj_out.weight"));
    for (int side = 0; side < 2; ++side) {
        int L = side == 0 ? c.encoder_layers : c.decoder_layers;
        for (int i = 0; i < L; ++i) {
            std::string p = std::string("transformer.model.") + (side == 0 ? "encoder" : "decoder") + ".layers." + std::to_string(i) + ".";
            Off& o = side == 0 ? eo[i] : dof[i];
            std::vector<float> wq, bq;
            qkv_pack(p + "self_attn.", wq, bq);
            o.v[0] = push(W(p + "self_attn_layer_norm.weight")); o.v[1] = push(W(p + "self_attn_layer_norm.bias"));
            o.v[2] = push(wq); o.v[3] = push(bq);
            o.v[4] = push(W(p + "self_attn.out_proj.weight")); o.v[5] = push(W(p + "self_attn.out_proj.bias"));
            if (side == 1) {
                qkv_pack(p + "encoder_attn.", wq, bq);
                o.v[6] = push(W(p + "encoder_attn_layer_norm.weight")); o.v[7] = push(W(p + "encoder_attn_layer_norm.bias"));
                o.v[8] = push(std::vector<float>(wq.begin(), wq.begin() + (size_t)d * d));
                o.v[9] = push(std::vector<float>(bq.begin(), bq.begin() + d));
                o.v[10] = push(std::vector<float>(wq.begin() + (size_t)d * d, wq.end()));
                o.v[11] = push(std::vector<float>(bq.begin() + d, bq.end()));
                o.v[12] = push(W(p + "encoder_attn.out_proj.weight")); o.v[13] = push(W(p + "encoder_attn.out_proj.bias"));
            }
            o.v[14] = push(W(p + "final_layer_norm.weight")); o.v[15] = push(W(p + "final_layer_norm.bias"));
            o.v[16] = push(W(p + "fc1.weight")); o.v[17] = push(W(p + "fc1.bias"));
            o.v[18] = push(W(p + "fc2.weight")); o.v[19] = push(W(p + "fc2.bias"));
            MB_REQUIRE(W(p + "fc1.weight").size() == (size_t)f * d, "wrong fc1 shape");
        }
    }
    MB_TRY(m->arena.ensure(pack.size() * sizeof(float)));
    MB_CUDA_CHECK(cudaMemcpy(m->arena.p, pack.data(), pack.size() * sizeof(float), cudaMemcpyHostToDevice));
    const float* base = m->arena.as<float>();
    m->emb_w = base + o_emb_w; m->emb_b = base + o_emb_b; m->conv1_w = base + o_c1w; m->conv1_b = base + o_c1b;
    m->conv2_w = base + o_c2w; m->conv2_b = base + o_c2b; m->enc_pos = base + o_epos; m->enc_ln_w = base + o_elnw; m->enc_ln_b = base + o_elnb;
    m->tok_emb = base + o_tok; m->dec_pos = base + o_dpos; m->dec_ln_w = base + o_dlnw; m->dec_ln_b = base + o_dlnb; m->proj_out = base + o_proj;
    auto fill = [&](const Off& o, bool cross) {
        LayerW l{};
        l.ln1_w = base + o.v[0]; l.ln1_b = base + o.v[1]; l.wqkv = base + o.v[2]; l.bqkv = base + o.v[3]; l.wo = base + o.v[4]; l.bo = base + o.v[5];
        if (cross) {
            l.ln2_w = base + o.v[6]; l.ln2_b = base + o.v[7]; l.wq_c = base + o.v[8]; l.bq_c = base + o.v[9];
            l.wkv_c = base + o.v[10]; l.bkv_c = base + o.v[11]; l.wo_c = base + o.v[12]; l.bo_c = base + o.v[13];
        }
        l.ln3_w = base + o.v[14]; l.ln3_b = base + o.v[15]; l.fc1_w = base + o.v[16]; l.fc1_b = base + o.v[17];
        l.fc2_w = base + o.v[18]; l.fc2_b = base + o.v[19];
        return l;
    };
    for (auto& o : eo) m->enc.push_back(fill(o, false));
    for (auto& o : dof) m->dec.push_back(fill(o, true));
    m->host_w.clear();
    // tf32 "lo" mirrors of the weights that feed large (tensor-core) GEMMs: encoder stem + layers, cross K|V projections
    {
        const long long dd = (long long)d * d;
        auto reg = [&](const float* w, long long n) -> int { return m->gemm.register_weight(w, n); };
        MB_TRY(reg(m->emb_w, (long long)d * c.mel.n_mels));
        MB_TRY(reg(m->conv1_w, 3 * dd)); MB_TRY(reg(m->conv2_w, 3 * dd));
        for (const auto& l : m->enc) {
            MB_TRY(reg(l.wqkv, 3 * dd)); MB_TRY(reg(l.wo, dd));
            MB_TRY(reg(l.fc1_w, (long long)f * d)); MB_TRY(reg(l.fc2_w, (long long)f * d));
        }
        for (const auto& l : m->dec) MB_TRY(reg(l.wkv_c, 2 * dd));
        MB_CUDA_CHECK(cudaDeviceSynchronize());
        // Scratch for the largest encoder chunk, allocated once: captured prefill graphs hold the split-K workspace pointer, so
        // nothing in the context may move after this point.
        const size_t chunk = (size_t)std::min(std::max(1, c.max_windows), 16);
        const size_t a_floats = std::max({chunk * (size_t)(c.src_seq_len / 2) * f, chunk * (size_t)(c.src_seq_len + 2) * d,
                                          chunk * (size_t)c.src_seq_len * c.mel.n_mels});
        m->gemm.num_sms = std::max(1, m->num_sms_phys);
        MB_TRY(m->gemm.reserve((size_t)64 << 20, a_floats * 8 + 1024));
        m->gemm.frozen = true;
        MB_TRY(m->attn.reserve(attn_tc_workspace_bytes((int)chunk, c.heads, c.src_seq_len / 2, c.src_seq_len / 2)));
        m->attn.frozen = true;
    }

    // ---- resident state ----
    m->max_rows = std::max(1, c.max_batch);
    MB_TRY(m->cross_kv.ensure((size_t)c.decoder_layers * m->cross_layer_stride() * sizeof(float)));
    MB_TRY(m->self_kv.ensure((size_t)c.decoder_layers * m->self_layer_stride() * sizeof(float)));
    MB_TRY(m->g_state.ensure(sizeof(GenState)));
    MB_TRY(m->g_cfg.ensure(sizeof(SampleConfig)));
    MB_TRY(m->g_vflags.ensure(c.vocab_size_in));
    MB_TRY(m->g_ids.ensure((size_t)m->max_rows * c.tgt_seq_len * sizeof(long long)));
    MB_TRY(m->g_prefill_ids.ensure((size_t)m->max_rows * c.tgt_seq_len * sizeof(long long)));
    MB_TRY(m->g_keyvalid.ensure((size_t)m->max_rows * c.tgt_seq_len));
    MB_TRY(m->g_leftpad.ensure(m->max_rows * sizeof(int)));
    MB_TRY(m->g_rowslot.ensure(m->max_rows * sizeof(int)));
    MB_TRY(m->g_finished.ensure(m->max_rows));
    MB_TRY(m->g_lastts.ensure(m->max_rows * sizeof(int)));
    MB_TRY(m->g_lastscores.ensure((size_t)2 * m->max_rows * c.vocab_size_out * sizeof(float)));
    MB_TRY(m->d_x.ensure((size_t)m->max_rows * d * sizeof(float)));
    MB_TRY(m->d_q.ensure((size_t)m->max_rows * d * sizeof(float)));
    MB_TRY(m->d_h.ensure((size_t)m->max_rows * f * sizeof(float)));
    MB_TRY(m->d_logits.ensure((size_t)m->max_rows * c.vocab_size_out * sizeof(float)));
    {   // prefill activations for the largest possible call, so captured prefill graphs never see a reallocation
        const size_t RPmax = (size_t)m->max_rows * c.tgt_seq_len;
        MB_TRY(m->p_x.ensure(RPmax * d * 4)); MB_TRY(m->p_h.ensure(RPmax * d * 4)); MB_TRY(m->p_q.ensure(RPmax * d * 4));
        MB_TRY(m->p_attn.ensure(RPmax * d * 4)); MB_TRY(m->p_ffn.ensure(RPmax * f * 4));
    }
    {   // split-KV partials sized for the longest possible context so captured graphs never see a reallocation
        const int max_splits = std::max((c.tgt_seq_len + 63) / 64, (c.src_seq_len / 2 + 63) / 64);
        MB_TRY(m->d_parto.ensure((size_t)m->max_rows * c.heads * max_splits * 64 * sizeof(float)));
        MB_TRY(m->d_partml.ensure((size_t)m->max_rows * c.heads * max_splits * 2 * sizeof(float)));
        MB_TRY(m->d_attn.ensure((size_t)m->max_rows * d * sizeof(float)));
        MB_TRY(m->d_ticket.ensure((size_t)m->max_rows * c.heads * sizeof(int)));
        MB_CUDA_CHECK(cudaMemset(m->d_ticket.p, 0, (size_t)m->max_rows * c.heads * sizeof(int)));
    }
    {   // exchange buffers of the dataflow megakernel: 8-byte {value | tag} pairs, two decoder rows
        const int R2 = 2, max_splits = std::max((c.tgt_seq_len + 63) / 64, (c.src_seq_len / 2 + 63) / 64);
        auto al = [](size_t n) { return (n + 15) & ~size_t(15); };
        const size_t n_x = al((size_t)R2 * d), n_kv = al((size_t)R2 * 2 * d), n_h = al((size_t)R2 * f), n_l = al((size_t)R2 * c.vocab_size_out),
                     n_p = al((size_t)R2 * c.heads * max_splits * 66), n_hdr = 16;
        const size_t RM = MEGA_LL_MAX_REPS;
        m->ll_bytes = ((2 * RM + 1) * n_x + n_kv + RM * n_h + n_l + n_p + n_hdr) * 8;
        MB_TRY(m->ll_arena.ensure(m->ll_bytes, true));
        unsigned long long* b = m->ll_arena.as<unsigned long long>();
        m->ll.x = b; b += RM * n_x; m->ll.att = b; b += RM * n_x; m->ll.q = b; b += n_x; m->ll.kvnew = b; b += n_kv; m->ll.h = b; b += RM * n_h;
        m->ll.logits = b; b += n_l; m->ll.part = b; b += n_p; m->ll.hdr = b;
        m->ll.max_splits = max_splits;
        m->ll.reps = 8; m->ll.x_rep = (long long)n_x; m->ll.h_rep = (long long)n_h;
    }
    m->finalized = true;
    return 0;
}

// =====================================================================================================================
// encoder
// =====================================================================================================================
static int encode_chunk(mb200_model* m, const float* pcm, int n, int slot_begin, float* enc_out, cudaStream_t st) {
    const auto& c = m->cfg;
    const int d = c.d_model, f = c.ffn_dim, H = c.heads, T2 = c.src_seq_len, T = T2 / 2, nm = c.mel.n_mels;
    const int n_samples = (c.src_seq_len - 1) * c.mel.hop_length;
    float* mel = m->w_mel.as<float>();
    float* embp = m->w_embpad.as<float>();
    float* c1p = m->w_c1pad.as<float>();
    float* x = m->w_x.as<float>();
    float* h = m->w_h.as<float>();
    float* qkv = m->w_qkv.as<float>();
    float* att = m->w_attn.as<float>();
    float* ffn = m->w_ffn.as<float>();
    const long long padrow = (long long)(T2 + 2) * d;

    MB_TRY(launch_mel(m->mel, pcm, n_samples, n, n_samples, mel, nm, (long long)T2 * nm, st));
    // encoder_embedder (modeling_mapperatorinator.py:433) written token-major into the zero-padded conv input
    {
        GemmParams g = gemm_base(plain_map(mel, nm), m->emb_w, nm, batched_map(embp + d, d, T2, padrow), m->emb_b, n * T2, d, nm);
        MB_TRY(launch_gemm(g, st, &m->gemm));
    }
    // conv1 k3 p1 + GELU  (row (b,t) of the padded buffer spans taps t-1, t, t+1 contiguously: K = 3d)
    {
        GemmParams g = gemm_base(batched_map(embp, d, T2, padrow), m->conv1_w, 3 * d, batched_map(c1p + d, d, T2, padrow), m->conv1_b,
                                 n * T2, d, 3 * d);
        g.act = ACT_GELU_ERF;
        MB_TRY(launch_gemm(g, st, &m->gemm));
    }
    // conv2 k3 s2 p1 + GELU + frozen positions
    {
        GemmParams g = gemm_base(batched_map(c1p, 2 * d, T, padrow), m->conv2_w, 3 * d, plain_map(x, d), m->conv2_b, n * T, d, 3 * d);
        g.act = ACT_GELU_ERF;
        g.R = batched_map(m->enc_pos, d, T, 0);
        MB_TRY(launch_gemm(g, st, &m->gemm));
    }
    const int rows = n * T;
    for (int l = 0; l < c.encoder_layers; ++l) {
        const LayerW& w = m->enc[l];
        MB_TRY(layernorm(x, h, w.ln1_w, w.ln1_b, rows, d, 1e-5f, st));
        MB_TRY(launch_gemm(gemm_base(plain_map(h, d), w.wqkv, d, plain_map(qkv, 3 * d), w.bqkv, rows, 3 * d, d), st, &m->gemm));
        AttentionParams a{};
        a.q = qkv; a.q_ld = 3 * d; a.q_bs = (long long)T * 3 * d;
        a.k = qkv + d; a.k_ld = 3 * d; a.k_bs = a.q_bs;
        a.v = qkv + 2 * d; a.v_ld = 3 * d; a.v_bs = a.q_bs;
        a.o = att; a.o_ld = d; a.o_bs = (long long)T * d;
        a.B = n; a.H = H; a.Tq = T; a.Tk = T; a.scale = 1.f; a.mask_mode = MASK_NONE;
        MB_TRY(launch_attention(a, st, &m->attn));
        {
            GemmParams g = gemm_base(plain_map(att, d), w.wo, d, plain_map(x, d), w.bo, rows, d, d);
            g.R = plain_map(x, d);
            MB_TRY(launch_gemm(g, st, &m->gemm));
        }
        MB_TRY(layernorm(x, h, w.ln3_w, w.ln3_b, rows, d, 1e-5f, st));
        {
            GemmParams g = gemm_base(plain_map(h, d), w.fc1_w, d, plain_map(ffn, f), w.fc1_b, rows, f, d);
            g.act = ACT_GELU_ERF;
            MB_TRY(launch_gemm(g, st, &m->gemm));
        }
        {
            GemmParams g = gemm_base(plain_map(ffn, f), w.fc2_w, f, plain_map(x, d), w.fc2_b, rows, d, f);
            g.R = plain_map(x, d);
            MB_TRY(launch_gemm(g, st, &m->gemm));
        }
    }
    float* enc = enc_out ? enc_out : h;
    MB_TRY(layernorm(x, enc, m->enc_ln_w, m->enc_ln_b, rows, d, 1e-5f, st));
    // cross-attention K|V of every decoder layer, straight into the resident slots
    for (int l = 0; l < c.decoder_layers; ++l) {
        float* dst = m->cross_kv.as<float>() + (size_t)l * m->cross_layer_stride() + (size_t)slot_begin * T * 2 * d;
        MB_TRY(launch_gemm(gemm_base(plain_map(enc, d), m->dec[l].wkv_c, d, plain_map(dst, 2 * d), m->dec[l].bkv_c, rows, 2 * d, d), st, &m->gemm));
    }
    return 0;
}

extern "C" int mb200_model_encode(mb200_model* m, const float* pcm, int32_t n_windows, int32_t slot_begin, float* enc_out, void* stream) {
    MB_REQUIRE(m && m->finalized, "model not finalized");
    MB_REQUIRE(slot_begin >= 0 && slot_begin + n_windows <= m->cfg.max_windows, "encoder slots out of range (raise max_windows)");
    cudaStream_t st = (cudaStream_t)stream;
    const auto& c = m->cfg;
    const int d = c.d_model, T2 = c.src_seq_len, T = T2 / 2;
    const int chunk = std::min(n_windows, 16);
    if (chunk > m->enc_chunk) {
        MB_TRY(m->w_mel.ensure((size_t)chunk * T2 * c.mel.n_mels * 4));
        MB_TRY(m->w_embpad.ensure((size_t)chunk * (T2 + 2) * d * 4, true));
        MB_TRY(m->w_c1pad.ensure((size_t)chunk * (T2 + 2) * d * 4, true));
        MB_TRY(m->w_x.ensure((size_t)chunk * T * d * 4));
        MB_TRY(m->w_h.ensure((size_t)chunk * T * d * 4));
        MB_TRY(m->w_qkv.ensure((size_t)chunk * T * 3 * d * 4));
        MB_TRY(m->w_attn.ensure((size_t)chunk * T * d * 4));
        MB_TRY(m->w_ffn.ensure((size_t)chunk * T * c.ffn_dim * 4));
        m->enc_chunk = chunk;
        for (auto& g : m->enc_graphs) cudaGraphExecDestroy(g.second.first);      // the captured encodes point into the old workspaces
        m->enc_graphs.clear(); m->enc_seen.clear();
    }
    const long long n_samples = (long long)(c.src_seq_len - 1) * c.mel.hop_length;
    if (n_windows == 1 && enc_out == nullptr && m->enc_graph) {
        // The drop-in call pattern (`server.model_generate` encodes its one window per call): ~200 launches of 5-60 us.  The first call of
        // a slot runs eagerly, from the second on the same sequence is replayed as one CUDA graph over an engine-owned copy of the PCM.
        auto seen = m->enc_seen.find(slot_begin);
        if (seen == m->enc_seen.end()) {
            m->enc_seen[slot_begin] = 1;
            return encode_chunk(m, pcm, 1, slot_begin, nullptr, st);
        }
        MB_TRY(m->w_pcm.ensure((size_t)n_samples * sizeof(float)));
        auto git = m->enc_graphs.find(slot_begin);
        if (git == m->enc_graphs.end()) {
            if (!m->cap_stream) MB_CUDA_CHECK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
            MB_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaGraph_t graph;
            MB_CUDA_CHECK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
            const long long before = g_launch_count;
            int rc = encode_chunk(m, m->w_pcm.as<float>(), 1, slot_begin, nullptr, m->cap_stream);
            cudaError_t e = cudaStreamEndCapture(m->cap_stream, &graph);
            const long long nodes = g_launch_count - before;
            g_launch_count = before;
            if (rc) return rc;
            MB_CUDA_CHECK(e);
            cudaGraphExec_t exec;
            MB_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            git = m->enc_graphs.emplace(slot_begin, std::make_pair(exec, nodes)).first;
        }
        MB_CUDA_CHECK(cudaMemcpyAsync(m->w_pcm.p, pcm, (size_t)n_samples * sizeof(float), cudaMemcpyDeviceToDevice, st));
        MB_CUDA_CHECK(cudaGraphLaunch(git->second.first, st));
        g_launch_count += git->second.second;
        return 0;
    }
    for (int i = 0; i < n_windows; i += m->enc_chunk) {
        int n = std::min(m->enc_chunk, n_windows - i);
        MB_TRY(encode_chunk(m, pcm + (long long)i * n_samples, n, slot_begin + i, enc_out ? enc_out + (long long)i * T * d : nullptr, st));
    }
    return 0;
}

// =====================================================================================================================
// decoder prefill (shared by generate and forward_logits)
// =====================================================================================================================
// ids_dev [rows, P] int64, keyvalid_dev [rows, P], row_slot dev; afterwards p_x holds the final hidden states [rows*P, d]
// (pre final-LN) and the self cache holds positions [0, P).
static int decoder_prefill(mb200_model* m, int rows, int P, const long long* ids_dev, int pos_rule, cudaStream_t st) {
    const auto& c = m->cfg;
    const int d = c.d_model, f = c.ffn_dim, H = c.heads, T = c.src_seq_len / 2;
    const size_t RP = (size_t)rows * P;
    MB_TRY(m->p_x.ensure(RP * d * 4)); MB_TRY(m->p_h.ensure(RP * d * 4)); MB_TRY(m->p_q.ensure(RP * d * 4));
    MB_TRY(m->p_attn.ensure(RP * d * 4)); MB_TRY(m->p_ffn.ensure(RP * f * 4));
    float *x = m->p_x.as<float>(), *h = m->p_h.as<float>(), *q = m->p_q.as<float>(), *att = m->p_attn.as<float>(), *ffn = m->p_ffn.as<float>();
    const unsigned char* kv_valid = m->g_keyvalid.as<unsigned char>();
    const int* row_slot = m->g_rowslot.as<int>();
    MB_TRY(launch_embed(ids_dev, P, rows, rows, P, m->g_leftpad.as<int>(), pos_rule, m->tok_emb, m->dec_pos, d, x, st));
    const long long self_row = (long long)c.tgt_seq_len * 2 * d;
    for (int l = 0; l < c.decoder_layers; ++l) {
        const LayerW& w = m->dec[l];
        float* skv = m->self_kv.as<float>() + (size_t)l * m->self_layer_stride();
        const float* ckv = m->cross_kv.as<float>() + (size_t)l * m->cross_layer_stride();
        MB_TRY(layernorm(x, h, w.ln1_w, w.ln1_b, (int)RP, d, 1e-5f, st));
        MB_TRY(launch_gemm(gemm_base(plain_map(h, d), w.wqkv, d, plain_map(q, d), w.bqkv, (int)RP, d, d), st, &m->gemm));
        MB_TRY(launch_gemm(gemm_base(plain_map(h, d), w.wqkv + (size_t)d * d, d, batched_map(skv, 2 * d, P, self_row), w.bqkv + d, (int)RP,
                                     2 * d, d), st, &m->gemm));
        AttentionParams a{};
        a.q = q; a.q_ld = d; a.q_bs = (long long)P * d;
        a.k = skv; a.k_ld = 2 * d; a.k_bs = self_row;
        a.v = skv + d; a.v_ld = 2 * d; a.v_bs = self_row;
        a.o = att; a.o_ld = d; a.o_bs = (long long)P * d;
        a.B = rows; a.H = H; a.Tq = P; a.Tk = P; a.scale = 1.f; a.mask_mode = MASK_CAUSAL; a.q_pos0 = 0;
        a.key_valid = kv_valid; a.key_valid_ld = c.tgt_seq_len;
        MB_TRY(launch_attention(a, st));
        {
            GemmParams g = gemm_base(plain_map(att, d), w.wo, d, plain_map(x, d), w.bo, (int)RP, d, d);
            g.R = plain_map(x, d);
            MB_TRY(launch_gemm(g, st, &m->gemm));
        }
        MB_TRY(layernorm(x, h, w.ln2_w, w.ln2_b, (int)RP, d, 1e-5f, st));
        MB_TRY(launch_gemm(gemm_base(plain_map(h, d), w.wq_c, d, plain_map(q, d), w.bq_c, (int)RP, d, d), st, &m->gemm));
        AttentionParams ca{};
        ca.q = q; ca.q_ld = d; ca.q_bs = (long long)P * d;
        ca.k = ckv; ca.k_ld = 2 * d; ca.k_bs = (long long)T * 2 * d;
        ca.v = ckv + d; ca.v_ld = 2 * d; ca.v_bs = ca.k_bs;
        ca.o = att; ca.o_ld = d; ca.o_bs = (long long)P * d;
        ca.B = rows; ca.H = H; ca.Tq = P; ca.Tk = T; ca.scale = 1.f; ca.mask_mode = MASK_NONE; ca.kv_slot = row_slot;
        MB_TRY(launch_attention(ca, st));
        {
            GemmParams g = gemm_base(plain_map(att, d), w.wo_c, d, plain_map(x, d), w.bo_c, (int)RP, d, d);
            g.R = plain_map(x, d);
            MB_TRY(launch_gemm(g, st, &m->gemm));
        }
        MB_TRY(layernorm(x, h, w.ln3_w, w.ln3_b, (int)RP, d, 1e-5f, st));
        {
            GemmParams g = gemm_base(plain_map(h, d), w.fc1_w, d, plain_map(ffn, f), w.fc1_b, (int)RP, f, d);
            g.act = ACT_GELU_ERF;
            MB_TRY(launch_gemm(g, st, &m->gemm));
        }
        {
            GemmParams g = gemm_base(plain_map(ffn, f), w.fc2_w, f, plain_map(x, d), w.fc2_b, (int)RP, d, f);
            g.R = plain_map(x, d);
            MB_TRY(launch_gemm(g, st, &m->gemm));
        }
    }
    return 0;
}

// =====================================================================================================================
// token step (captured into a CUDA graph)
// =====================================================================================================================
static GemvParams gemv_base(int xmode, const float* W, long long ldw, const float* bias, int K, int N, int B, const GenState* stt) {
    GemvParams g{};
    g.xmode = xmode; g.W = W; g.ldw = ldw; g.bias = bias; g.K = K; g.N = N; g.B = B; g.st = stt; g.eps = 1e-5f; g.nseg = 1;
    g.seg[0] = GemvSeg{nullptr, 0, 0, 0, N, 1.f, ACT_NONE};
    return g;
}

static GemvParams final_logits_params(mb200_model* m, int rows, const float* x, long long x_ld) {
    const auto& c = m->cfg;
    GemvParams g = gemv_base(X_LAYERNORM, m->proj_out, c.d_model, nullptr, c.d_model, c.vocab_size_out, rows, m->g_state.as<GenState>());
    g.x = x; g.x_ld = x_ld; g.ln_w = m->dec_ln_w; g.ln_b = m->dec_ln_b;
    g.seg[0].out = m->d_logits.as<float>(); g.seg[0].out_bs = c.vocab_size_out;
    return g;
}

static int final_logits(mb200_model* m, int rows, const float* x, long long x_ld, cudaStream_t st, bool pdl) {
    return launch_gemv(final_logits_params(m, rows, x, x_ld), st, pdl);
}

static SampleParams sample_params(mb200_model* m, int rows) {
    const auto& c = m->cfg;
    SampleParams s{};
    s.logits = m->d_logits.as<float>(); s.logits_ld = c.vocab_size_out;
    s.cfg = m->g_cfg.as<SampleConfig>(); s.st = m->g_state.as<GenState>(); s.vflags = m->g_vflags.as<unsigned char>();
    s.ids = m->g_ids.as<long long>(); s.finished = m->g_finished.as<unsigned char>(); s.last_ts = m->g_lastts.as<int>();
    s.last_scores = m->g_lastscores.as<float>(); s.n_left_pad = m->g_leftpad.as<int>();
    s.tok_emb = m->tok_emb; s.pos_emb = m->dec_pos; s.d_model = c.d_model;
    s.x_out = m->d_x.as<float>(); s.x_ld = c.d_model; s.rows = rows;
    return s;
}

// split-KV layout of the self-attention cache: one 128-key split while the context fits, 64-key splits beyond
static int self_splits(int max_length) { return max_length <= 128 ? 1 : (max_length + 63) / 64; }

// Either launches the 98 micro-phases of one token on `st` (eager / graph capture), or — when `collect` is given — records
// them as phase descriptors for the persistent megakernel.  One definition, so both paths run the same arithmetic.
static int token_step(mb200_model* m, int rows, int B, int n_splits_self, cudaStream_t st, bool pdl,
                      std::vector<MegaPhase>* collect = nullptr) {
    auto emit_gemv = [&](const GemvParams& g) -> int {
        if (!collect) return launch_gemv(g, st, pdl);
        MegaPhase ph{}; ph.kind = 0; ph.g = g; collect->push_back(ph);
        return 0;
    };
    auto emit_attn = [&](const DecAttnParams& a) -> int {
        if (!collect) return launch_decode_attention(a, st, pdl);
        MegaPhase ph{}; ph.kind = 1; ph.a = a;
        ph.magic_ns = a.n_splits > 1 ? (unsigned)((1ull << 32) / (unsigned)a.n_splits + 1) : 0u;       // 0 = divisor 1 (2^32 + 1 does not fit)
        ph.magic_h = a.H > 1 ? (unsigned)((1ull << 32) / (unsigned)a.H + 1) : 0u;
        collect->push_back(ph);
        return 0;
    };
    const auto& c = m->cfg;
    const int d = c.d_model, f = c.ffn_dim, H = c.heads, T = c.src_seq_len / 2;
    const GenState* gs = m->g_state.as<GenState>();
    float *x = m->d_x.as<float>(), *q = m->d_q.as<float>(), *hh = m->d_h.as<float>();
    float *po = m->d_parto.as<float>(), *pml = m->d_partml.as<float>(), *attn = m->d_attn.as<float>();
    int* ticket = m->d_ticket.as<int>();
    const int chunk = 64, n_splits_cross = (T + chunk - 1) / chunk;
    const long long self_row = (long long)c.tgt_seq_len * 2 * d;
    for (int l = 0; l < c.decoder_layers; ++l) {
        const LayerW& w = m->dec[l];
        float* skv = m->self_kv.as<float>() + (size_t)l * m->self_layer_stride();
        const float* ckv = m->cross_kv.as<float>() + (size_t)l * m->cross_layer_stride();
        {   // LN1 -> q | k | v  (k, v land in the self cache at position cur_len - 1)
            GemvParams g = gemv_base(X_LAYERNORM, w.wqkv, d, w.bqkv, d, 3 * d, rows, gs);
            g.x = x; g.x_ld = d; g.ln_w = w.ln1_w; g.ln_b = w.ln1_b;
            g.nseg = 3;
            g.seg[0] = GemvSeg{q, d, 0, 0, d, 1.f, ACT_NONE};
            g.seg[1] = GemvSeg{skv, self_row, 2 * d, d, 2 * d, 1.f, ACT_NONE};
            g.seg[2] = GemvSeg{skv + d, self_row, 2 * d, 2 * d, 3 * d, 1.f, ACT_NONE};
            MB_TRY(emit_gemv(g));
        }
        {
            DecAttnParams a{};
            a.q = q; a.q_ld = d; a.kc = skv; a.vc = skv + d; a.row_stride = self_row; a.tok_stride = 2 * d; a.row_slot = nullptr;
            a.fixed_len = 0; a.st = gs; a.key_valid = m->g_keyvalid.as<unsigned char>(); a.key_valid_ld = c.tgt_seq_len;
            a.part_o = po; a.part_ml = pml; a.rows = rows; a.H = H; a.n_splits = n_splits_self;
            a.chunk = n_splits_self == 1 ? 128 : chunk;      // contexts up to 128 tokens: one split per head, no merge step
            a.out = attn; a.out_ld = d; a.ticket = ticket;
            MB_TRY(emit_attn(a));
        }
        {   // out_proj + residual (the heads were merged by the attention phase)
            GemvParams g = gemv_base(X_PLAIN, w.wo, d, w.bo, d, d, rows, gs);
            g.x = attn; g.x_ld = d;
            g.seg[0].out = x; g.seg[0].out_bs = d; g.R = x; g.r_ld = d;
            MB_TRY(emit_gemv(g));
        }
        {   // LN2 -> cross q
            GemvParams g = gemv_base(X_LAYERNORM, w.wq_c, d, w.bq_c, d, d, rows, gs);
            g.x = x; g.x_ld = d; g.ln_w = w.ln2_w; g.ln_b = w.ln2_b;
            g.seg[0].out = q; g.seg[0].out_bs = d;
            MB_TRY(emit_gemv(g));
        }
        {
            DecAttnParams a{};
            a.q = q; a.q_ld = d; a.kc = ckv; a.vc = ckv + d; a.row_stride = (long long)T * 2 * d; a.tok_stride = 2 * d;
            a.row_slot = m->g_rowslot.as<int>(); a.fixed_len = T; a.st = gs; a.key_valid = nullptr;
            a.part_o = po; a.part_ml = pml; a.rows = rows; a.H = H; a.n_splits = n_splits_cross; a.chunk = chunk;
            a.out = attn; a.out_ld = d; a.ticket = ticket;
            MB_TRY(emit_attn(a));
        }
        {
            GemvParams g = gemv_base(X_PLAIN, w.wo_c, d, w.bo_c, d, d, rows, gs);
            g.x = attn; g.x_ld = d;
            g.seg[0].out = x; g.seg[0].out_bs = d; g.R = x; g.r_ld = d;
            MB_TRY(emit_gemv(g));
        }
        {   // LN3 -> fc1 + GELU
            GemvParams g = gemv_base(X_LAYERNORM, w.fc1_w, d, w.fc1_b, d, f, rows, gs);
            g.x = x; g.x_ld = d; g.ln_w = w.ln3_w; g.ln_b = w.ln3_b;
            g.seg[0] = GemvSeg{hh, f, 0, 0, f, 1.f, ACT_GELU_ERF};
            MB_TRY(emit_gemv(g));
        }
        {   // fc2 + residual
            GemvParams g = gemv_base(X_PLAIN, w.fc2_w, f, w.fc2_b, f, d, rows, gs);
            g.x = hh; g.x_ld = f;
            g.seg[0].out = x; g.seg[0].out_bs = d; g.R = x; g.r_ld = d;
            MB_TRY(emit_gemv(g));
        }
    }
    MB_TRY(emit_gemv(final_logits_params(m, rows, x, d)));
    if (collect) {
        MegaPhase ph{}; ph.kind = 2; collect->push_back(ph);
        const int n = (int)collect->size();
        for (int i = 0; i < n; ++i) {          // next GEMV phase after i, wrapping into the next token
            int j = (i + 1) % n;
            while ((*collect)[j].kind != 0) j = (j + 1) % n;
            (*collect)[i].next_gemv = j;
            (*collect)[i].nx_W = (*collect)[j].g.W; (*collect)[i].nx_ldw = (*collect)[j].g.ldw;
            (*collect)[i].nx_N = (*collect)[j].g.N; (*collect)[i].nx_K = (*collect)[j].g.K;
            (*collect)[i].nx_rpc = ((*collect)[j].g.N + m->num_sms - 1) / m->num_sms;
            (*collect)[i].rpc = (*collect)[i].kind == 0 ? ((*collect)[i].g.N + m->num_sms - 1) / m->num_sms : 0;
        }
        return 0;
    }
    MB_TRY(launch_sample(sample_params(m, rows), B, st, pdl));
    return 0;
}

// The persistent path: all remaining tokens of the call in one cooperative launch (rows <= 2, weight slices must fit).
static int run_megakernel(mb200_model* m, int rows, int B, int n_splits_self, int max_steps, cudaStream_t st) {
    auto key = std::make_pair(rows, n_splits_self);
    auto it = m->mega_phases.find(key);
    if (it == m->mega_phases.end()) {
        std::vector<MegaPhase> phases;
        MB_TRY(token_step(m, rows, B, n_splits_self, st, false, &phases));
        DevBuf* buf = new DevBuf();
        MB_TRY(buf->ensure(phases.size() * sizeof(MegaPhase)));
        MB_CUDA_CHECK(cudaMemcpy(buf->p, phases.data(), phases.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice));
        it = m->mega_phases.emplace(key, std::make_pair(buf, (int)phases.size())).first;
    }
    MB_TRY(m->g_megasync.ensure(64));
    MB_CUDA_CHECK(cudaMemsetAsync(m->g_megasync.p, 0, 64, st));
    MegaParams mp{};
    mp.phases = it->second.first->as<MegaPhase>(); mp.n_phases = it->second.second; mp.first_gemv = 0;
    mp.sample = sample_params(m, rows); mp.st = m->g_state.as<GenState>();
    mp.sync_counter = m->g_megasync.as<unsigned int>(); mp.error_flag = m->g_megasync.as<int>() + 8;
    mp.max_steps = max_steps; mp.row_slot = m->g_rowslot.as<int>();
    mp.trace = m->mega_trace.p ? m->mega_trace.as<unsigned long long>() : nullptr; mp.trace_step = 8;
    if (!m->mega_ev[0]) { MB_CUDA_CHECK(cudaEventCreate(&m->mega_ev[0])); MB_CUDA_CHECK(cudaEventCreate(&m->mega_ev[1])); }
    MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag + 2, &m->g_state.as<GenState>()->cur_len, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaEventRecord(m->mega_ev[0], st));
    MB_TRY(launch_megakernel(mp, m->num_sms, st));
    MB_CUDA_CHECK(cudaEventRecord(m->mega_ev[1], st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag + 1, m->g_megasync.as<int>() + 8, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag + 3, &m->g_state.as<GenState>()->cur_len, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    {
        float ms = 0.f;
        MB_CUDA_CHECK(cudaEventElapsedTime(&ms, m->mega_ev[0], m->mega_ev[1]));
        m->mega_ms += ms; m->mega_launches += 1; m->mega_tokens += m->h_flag[3] - m->h_flag[2];
    }
    MB_REQUIRE(m->h_flag[1] == 0, m->h_flag[1] == 1 ? "megakernel grid barrier timed out" : "megakernel weight copy timed out");
    return 0;
}

// The dataflow path: same phase list, each phase annotated with the exchange buffers it reads / writes.
static int run_megakernel2(mb200_model* m, int rows, int B, int n_splits_self, int max_steps, cudaStream_t st) {
    auto key = std::make_tuple(rows, B, n_splits_self);
    auto it = m->mega2_phases.find(key);
    if (it == m->mega2_phases.end()) {
        std::vector<MegaPhase> phases;
        MB_TRY(token_step(m, rows, B, n_splits_self, st, false, &phases));
        std::vector<Mega2Phase> p2(phases.size());
        const float *dx = m->d_x.as<float>(), *dq = m->d_q.as<float>(), *dh = m->d_h.as<float>(), *datt = m->d_attn.as<float>(),
                    *dlog = m->d_logits.as<float>();
        for (size_t i = 0; i < phases.size(); ++i) {
            Mega2Phase& q = p2[i];
            q = Mega2Phase{};
            q.base = phases[i];
            if (phases[i].kind != 0) continue;
            const GemvParams& g = phases[i].g;
            q.in_sel = g.xmode == X_LAYERNORM ? LL_X : (g.x == datt ? LL_ATT : (g.x == dh ? LL_H : LL_NONE));
            MB_REQUIRE(q.in_sel != LL_NONE && (g.xmode != X_LAYERNORM || g.x == dx), "dataflow megakernel: unknown GEMV input buffer");
            q.res_xraw = g.R != nullptr;
            MB_REQUIRE(!g.R || g.R == dx, "dataflow megakernel: residual must be the residual stream");
            for (int sgi = 0; sgi < g.nseg; ++sgi) {
                const GemvSeg& sg = g.seg[sgi];
                if (sg.pos_stride != 0) { q.out_sel[sgi] = sgi == 1 ? LL_K : LL_V; q.plain_out[sgi] = 1; MB_REQUIRE(sgi >= 1, "cache segment order"); }
                else if (sg.out == dq) q.out_sel[sgi] = LL_Q;
                else if (sg.out == dx) q.out_sel[sgi] = LL_X;
                else if (sg.out == dh) q.out_sel[sgi] = LL_H;
                else if (sg.out == dlog) q.out_sel[sgi] = LL_LOGITS;
                else MB_REQUIRE(false, "dataflow megakernel: unknown GEMV output buffer");
                const int os = q.out_sel[sgi];
                const unsigned long long* ob = os == LL_X ? m->ll.x : os == LL_H ? m->ll.h : os == LL_Q ? m->ll.q : os == LL_LOGITS ? m->ll.logits : m->ll.kvnew;
                q.out_off[sgi] = (long long)(ob - m->ll.x) + (os == LL_V ? m->cfg.d_model : 0);
                q.out_rs[sgi] = os == LL_X ? m->ll.x_rep : (os == LL_H ? m->ll.h_rep : 0);
                q.out_bw[sgi] = os == LL_H ? g.N : (os == LL_K || os == LL_V ? 2 * m->cfg.d_model : (os == LL_LOGITS ? m->cfg.vocab_size_out : m->cfg.d_model));
            }
            {
                const unsigned long long* ib = q.in_sel == LL_H ? m->ll.h : (q.in_sel == LL_ATT ? m->ll.att : m->ll.x);
                q.in_off = (long long)(ib - m->ll.x);
                q.in_rs = q.in_sel == LL_H ? m->ll.h_rep : m->ll.x_rep;
                const int rpc = (g.N + m->num_sms - 1) / m->num_sms;
                q.n_active = (g.N + rpc - 1) / rpc;
            }
        }
        DevBuf* buf = new DevBuf();
        MB_TRY(buf->ensure(p2.size() * sizeof(Mega2Phase)));
        MB_CUDA_CHECK(cudaMemcpy(buf->p, p2.data(), p2.size() * sizeof(Mega2Phase), cudaMemcpyHostToDevice));
        it = m->mega2_phases.emplace(key, std::make_pair(buf, (int)p2.size())).first;
    }
    MB_TRY(m->g_megasync.ensure(64));
    MB_CUDA_CHECK(cudaMemsetAsync(m->g_megasync.p, 0, 64, st));
    MB_CUDA_CHECK(cudaMemsetAsync(m->ll_arena.p, 0, m->ll_bytes, st));        // tag 0 = "nothing here yet"
    Mega2Params mp{};
    mp.phases = it->second.first->as<Mega2Phase>(); mp.n_phases = it->second.second;
    mp.sample = sample_params(m, rows); mp.st = m->g_state.as<GenState>();
    mp.ll = m->ll; mp.error_flag = m->g_megasync.as<int>() + 8;
    mp.sample.ll_logits = m->ll.logits; mp.sample.ll_x_out = m->ll.x; mp.sample.ll_hdr = m->ll.hdr; mp.sample.ll_err = mp.error_flag;
    mp.sample.ll_reps = m->ll.reps; mp.sample.ll_x_rep = m->ll.x_rep;
    mp.max_steps = max_steps; mp.row_slot = m->g_rowslot.as<int>(); mp.x_in = m->d_x.as<float>();
    mp.rows = rows; mp.d_model = m->cfg.d_model; mp.V = m->cfg.vocab_size_out; mp.ffn_dim = m->cfg.ffn_dim; mp.trace_cta = m->trace_cta;
    mp.trace = m->mega_trace.p ? m->mega_trace.as<unsigned long long>() : nullptr; mp.trace_step = 8;
    if (!m->mega_ev[0]) { MB_CUDA_CHECK(cudaEventCreate(&m->mega_ev[0])); MB_CUDA_CHECK(cudaEventCreate(&m->mega_ev[1])); }
    MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag + 2, &m->g_state.as<GenState>()->cur_len, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaEventRecord(m->mega_ev[0], st));
    MB_TRY(launch_megakernel2(mp, m->num_sms, st));
    MB_CUDA_CHECK(cudaEventRecord(m->mega_ev[1], st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag + 1, m->g_megasync.as<int>() + 8, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag + 3, &m->g_state.as<GenState>()->cur_len, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    {
        float ms = 0.f;
        MB_CUDA_CHECK(cudaEventElapsedTime(&ms, m->mega_ev[0], m->mega_ev[1]));
        m->mega_ms += ms; m->mega_launches += 1; m->mega_tokens += m->h_flag[3] - m->h_flag[2];
    }
    MB_REQUIRE(m->h_flag[1] == 0, m->h_flag[1] == 2 ? "dataflow megakernel: weight copy timed out" : "dataflow megakernel: a wait for tagged data timed out");
    return 0;
}

static bool mega_eligible(const mb200_model* m, int rows) {
    if (!m->use_mega || rows > 2 || m->num_sms <= 0) return false;
    const auto& c = m->cfg;
    const int G = m->num_sms;
    auto fits = [&](int N, int K) { return (size_t)((N + G - 1) / G) * K <= (size_t)MEGA_WBUF_FLOATS; };
    return c.ffn_dim <= 3072 && fits(3 * c.d_model, c.d_model) && fits(c.d_model, c.d_model) && fits(c.ffn_dim, c.d_model) &&
           fits(c.d_model, c.ffn_dim) && fits(c.vocab_size_out, c.d_model);
}

static SampleConfig make_sample_config(const mb200_generate_params* gp, int B, bool use_cfg, int V, int ids_ld) {
    SampleConfig sc{};
    sc.B = B; sc.use_cfg = use_cfg ? 1 : 0; sc.cfg_scale = gp->cfg_scale; sc.V = V; sc.ts_start = gp->time_shift_start; sc.ts_end = gp->time_shift_end;
    sc.timeshift_bias = gp->timeshift_bias; sc.types_first = gp->types_first; sc.temperature = gp->temperature;
    sc.n_cond = gp->n_cond;
    for (int i = 0; i < 3; ++i) { sc.cond_temp[i] = gp->cond_temp[i]; sc.cond_offset[i] = gp->cond_offset[i]; sc.cond_flag[i] = gp->cond_flag[i]; }
    sc.lookback_on = gp->lookback_on; sc.lookback_start = gp->lookback_start; sc.lookback_end = gp->lookback_end;
    sc.do_sample = gp->do_sample; sc.top_k = gp->top_k; sc.top_p = gp->top_p; sc.top_p_cut = gp->top_p_cut; sc.seed = gp->seed; sc.pad_id = gp->pad_token_id;
    sc.pos_rule_cumsum = gp->position_rule; sc.ids_ld = ids_ld;
    return sc;
}

// =====================================================================================================================
extern "C" int mb200_model_generate(mb200_model* m, const int32_t* slots, int32_t B, const int64_t* prompt, const uint8_t* prompt_mask,
                                    int32_t P, const int64_t* neg_prompt, const uint8_t* neg_mask, const uint8_t* vflags,
                                    const mb200_generate_params* gp, int64_t* out_ids, int32_t* out_len, void* stream) {
    MB_REQUIRE(m && m->finalized, "model not finalized");
    MB_REQUIRE(slots && prompt && vflags && gp && out_ids && out_len, "null argument");
    const auto& c = m->cfg;
    cudaStream_t st = (cudaStream_t)stream;
    const bool use_cfg = neg_prompt != nullptr;
    const int rows = use_cfg ? 2 * B : B;
    MB_REQUIRE(B >= 1 && rows <= m->max_rows, "batch exceeds max_batch (rows double under classifier-free guidance)");
    MB_REQUIRE(P >= 1 && P < gp->max_length && gp->max_length <= c.tgt_seq_len, "need 1 <= prompt_len < max_length <= tgt_seq_len");
    for (int b = 0; b < B; ++b) MB_REQUIRE(slots[b] >= 0 && slots[b] < c.max_windows, "encoder slot out of range");
    const int d = c.d_model, H = c.heads, T = c.src_seq_len / 2, V = c.vocab_size_out;
    const int ids_ld = c.tgt_seq_len;

    // ---- host-side staging of the call state ----
    std::vector<long long> pre((size_t)rows * P), idsrow((size_t)B * ids_ld, (long long)gp->pad_token_id);
    std::vector<unsigned char> kv((size_t)rows * ids_ld, 1);
    std::vector<int> leftpad(rows, 0), rowslot(rows);
    for (int r = 0; r < rows; ++r) {
        const int b = r % B;
        const bool neg_row = use_cfg && r < B;   // first half carries the negative prompt (modeling_mapperatorinator.py:243-245)
        const int64_t* src = neg_row ? neg_prompt : prompt;
        const uint8_t* msk = neg_row ? (neg_mask ? neg_mask : prompt_mask) : prompt_mask;
        int npad = 0; bool seen = false;
        for (int t = 0; t < P; ++t) {
            long long tok = src[(size_t)b * P + t];
            MB_REQUIRE(tok >= 0 && tok < c.vocab_size_in, "prompt token id out of range");
            pre[(size_t)r * P + t] = tok;
            unsigned char ok = msk ? (msk[(size_t)b * P + t] != 0) : 1;
            kv[(size_t)r * ids_ld + t] = ok;
            if (!ok && !seen) ++npad; else seen = true;
        }
        leftpad[r] = npad;
        rowslot[r] = slots[b];
    }
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < P; ++t) idsrow[(size_t)b * ids_ld + t] = prompt[(size_t)b * P + t];
    GenState gs{};
    gs.cur_len = P; gs.prompt_len = P; gs.max_length = gp->max_length; gs.min_new_tokens = gp->min_new_tokens;
    const SampleConfig sc = make_sample_config(gp, B, use_cfg, V, ids_ld);

    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_prefill_ids.p, pre.data(), pre.size() * 8, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_ids.p, idsrow.data(), idsrow.size() * 8, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_keyvalid.p, kv.data(), kv.size(), cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_leftpad.p, leftpad.data(), rows * 4, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_rowslot.p, rowslot.data(), rows * 4, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_vflags.p, vflags, c.vocab_size_in, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_state.p, &gs, sizeof(gs), cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_cfg.p, &sc, sizeof(sc), cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemsetAsync(m->g_finished.p, 0, m->max_rows, st));
    MB_CUDA_CHECK(cudaMemsetAsync(m->d_ticket.p, 0, (size_t)m->max_rows * m->cfg.heads * sizeof(int), st));   // self-resetting; cleared in case a previous call aborted
    MB_CUDA_CHECK(cudaStreamSynchronize(st));   // host vectors go out of scope; the copies above are from pageable memory

    // partial buffers for the split-KV attentions
    const int chunk = 64;
    const int n_splits_self = self_splits(gp->max_length);
    (void)H; (void)T;

    MB_TRY(launch_prompt_scan(m->g_ids.as<long long>(), ids_ld, B, P, m->g_vflags.as<unsigned char>(), sc.ts_start, sc.ts_end,
                              m->g_lastts.as<int>(), st));
    // prefill + first token: ~230 small launches.  The first call of a given (rows, P) shape runs eagerly (it may allocate);
    // from the second call on the same sequence is replayed as one CUDA graph (sequential windows reuse a few prompt lengths).
    {
        auto run_prefill = [&](cudaStream_t s) -> int {
            MB_TRY(decoder_prefill(m, rows, P, m->g_prefill_ids.as<long long>(), gp->position_rule, s));
            MB_TRY(final_logits(m, rows, m->p_x.as<float>() + (size_t)(P - 1) * d, (long long)P * d, s, false));
            MB_TRY(launch_sample(sample_params(m, rows), B, s, false));
            return 0;
        };
        const auto pkey = std::make_tuple(rows, (int)B, (int)P, (int)gp->position_rule);
        auto seen = m->prefill_seen.find(pkey);
        if (seen == m->prefill_seen.end()) {
            m->prefill_seen[pkey] = 1;
            MB_TRY(run_prefill(st));
        } else {
            auto git = m->prefill_graphs.find(pkey);
            if (git == m->prefill_graphs.end()) {
                if (!m->cap_stream) MB_CUDA_CHECK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
                MB_CUDA_CHECK(cudaStreamSynchronize(st));
                cudaGraph_t graph;
                MB_CUDA_CHECK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
                const long long before = g_launch_count;
                int s = run_prefill(m->cap_stream);
                cudaError_t e = cudaStreamEndCapture(m->cap_stream, &graph);
                const long long nodes = g_launch_count - before;
                g_launch_count = before;
                if (s) return s;
                MB_CUDA_CHECK(e);
                cudaGraphExec_t exec;
                MB_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
                cudaGraphDestroy(graph);
                if (m->prefill_graphs.size() >= 48) {      // bounded cache: real songs see many prompt lengths; drop everything and re-learn
                    for (auto& g : m->prefill_graphs) cudaGraphExecDestroy(g.second.first);
                    m->prefill_graphs.clear();
                    m->prefill_seen.clear();
                }
                git = m->prefill_graphs.emplace(pkey, std::make_pair(exec, nodes)).first;
            }
            MB_CUDA_CHECK(cudaGraphLaunch(git->second.first, st));
            g_launch_count += git->second.second;
        }
    }

    // ---- token loop, persistent path: every remaining token in ONE cooperative launch ----
    if (mega_eligible(m, rows) && gp->max_length - (P + 1) > 0) {
        bool dataflow = m->use_mega >= 2;
        if (dataflow) {     // every projection of this model must fit the K-split thread mapping, else the grid-barrier kernel takes the call
            const int G = m->num_sms;
            dataflow = mega2_ksplit_ok(3 * c.d_model, c.d_model, rows, G) && mega2_ksplit_ok(c.d_model, c.d_model, rows, G) &&
                       mega2_ksplit_ok(c.ffn_dim, c.d_model, rows, G) && mega2_ksplit_ok(c.d_model, c.ffn_dim, rows, G) &&
                       mega2_ksplit_ok(c.vocab_size_out, c.d_model, rows, G);
        }
        // (one 256-key attention unit per head for contexts of 129..256 tokens was tried: 365 vs 344 us / token against three 64-key units +
        //  merge — the V rows beyond the first 64 keys are fetched inside the PV loop, and the wider unit costs instructions in EVERY unit)
        if (dataflow) MB_TRY(run_megakernel2(m, rows, B, n_splits_self, gp->max_length - (P + 1), st));
        else MB_TRY(run_megakernel(m, rows, B, n_splits_self, gp->max_length - (P + 1), st));
        MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag, &m->g_state.as<GenState>()->cur_len, 4, cudaMemcpyDeviceToHost, st));
        MB_CUDA_CHECK(cudaStreamSynchronize(st));
        const int Lm = *m->h_flag;
        MB_CUDA_CHECK(cudaMemcpy2DAsync(out_ids, (size_t)Lm * 8, m->g_ids.p, (size_t)ids_ld * 8, (size_t)Lm * 8, B, cudaMemcpyDeviceToHost, st));
        MB_CUDA_CHECK(cudaStreamSynchronize(st));
        *out_len = Lm;
        return 0;
    }
    // ---- token loop, graph path: one graph replay per token, flag polled every few tokens ----
    auto key = std::make_tuple(rows, (int)B, n_splits_self);
    auto it = m->graphs.find(key);
    if (it == m->graphs.end()) {
        // capture on an engine-owned stream (the caller's stream may be the legacy default stream, which cannot capture)
        cudaGraph_t graph;
        if (!m->cap_stream) MB_CUDA_CHECK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
        MB_CUDA_CHECK(cudaStreamSynchronize(st));
        MB_CUDA_CHECK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
        const long long before = g_launch_count;
        int s = token_step(m, rows, B, n_splits_self, m->cap_stream, m->use_pdl);
        cudaError_t e = cudaStreamEndCapture(m->cap_stream, &graph);
        m->graph_nodes[key] = g_launch_count - before;
        g_launch_count = before;   // captured, not launched; replays are counted below
        if (s) return s;
        MB_CUDA_CHECK(e);
        cudaGraphExec_t exec;
        MB_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        it = m->graphs.emplace(key, exec).first;
    }
    int remaining = gp->max_length - (P + 1);
    const int burst_default = 16;
    int produced = 1;
    while (remaining > 0) {
        int burst = std::min(remaining, burst_default);
        // no EOS is possible before min_new_tokens are out, so the first poll can wait until then
        if (gp->min_new_tokens > produced) burst = std::min(remaining, std::max(burst, gp->min_new_tokens - produced));
        for (int i = 0; i < burst; ++i) MB_CUDA_CHECK(cudaGraphLaunch(it->second, st));
        g_launch_count += (long long)burst * m->graph_nodes[key];
        remaining -= burst; produced += burst;
        MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag, &m->g_state.as<GenState>()->all_finished, 4, cudaMemcpyDeviceToHost, st));
        MB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (*m->h_flag) break;
    }
    MB_CUDA_CHECK(cudaMemcpyAsync(m->h_flag, &m->g_state.as<GenState>()->cur_len, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    const int L = *m->h_flag;
    MB_CUDA_CHECK(cudaMemcpy2DAsync(out_ids, (size_t)L * 8, m->g_ids.p, (size_t)ids_ld * 8, (size_t)L * 8, B, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    *out_len = L;
    return 0;
}

extern "C" int mb200_model_forward_logits(mb200_model* m, const int32_t* slots, int32_t B, const int64_t* ids, const uint8_t* mask,
                                          int32_t len, int32_t position_rule, float* logits_out, void* stream) {
    MB_REQUIRE(m && m->finalized, "model not finalized");
    MB_REQUIRE(B >= 1 && B <= m->max_rows && len >= 1 && len <= m->cfg.tgt_seq_len, "bad batch / length");
    const auto& c = m->cfg;
    cudaStream_t st = (cudaStream_t)stream;
    const int ids_ld = c.tgt_seq_len, d = c.d_model;
    std::vector<unsigned char> kv((size_t)B * ids_ld, 1);
    std::vector<int> leftpad(B, 0), rowslot(B);
    std::vector<long long> pre((size_t)B * len);
    for (int b = 0; b < B; ++b) {
        int npad = 0; bool seen = false;
        for (int t = 0; t < len; ++t) {
            long long tok = ids[(size_t)b * len + t];
            MB_REQUIRE(tok >= 0 && tok < c.vocab_size_in, "token id out of range");
            pre[(size_t)b * len + t] = tok;
            unsigned char ok = mask ? (mask[(size_t)b * len + t] != 0) : 1;
            kv[(size_t)b * ids_ld + t] = ok;
            if (!ok && !seen) ++npad; else seen = true;
        }
        leftpad[b] = npad; rowslot[b] = slots[b];
    }
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_prefill_ids.p, pre.data(), pre.size() * 8, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_keyvalid.p, kv.data(), kv.size(), cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_leftpad.p, leftpad.data(), B * 4, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_rowslot.p, rowslot.data(), B * 4, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    MB_TRY(decoder_prefill(m, B, len, m->g_prefill_ids.as<long long>(), position_rule, st));
    const int R = B * len;
    MB_TRY(layernorm(m->p_x.as<float>(), m->p_h.as<float>(), m->dec_ln_w, m->dec_ln_b, R, d, 1e-5f, st));
    MB_TRY(launch_gemm(gemm_base(plain_map(m->p_h.as<float>(), d), m->proj_out, d, plain_map(logits_out, c.vocab_size_out), nullptr, R,
                                 c.vocab_size_out, d), st, &m->gemm));
    return 0;
}

// test / tuning hooks (not part of the reference-facing boundary)
extern "C" int mb200_model_set_option(mb200_model* m, const char* name, int value) {
    MB_REQUIRE(m && name, "null argument");
    if (!strcmp(name, "pdl")) {
        if (m->use_pdl != (value != 0)) {
            for (auto& g : m->graphs) cudaGraphExecDestroy(g.second);
            m->graphs.clear();
        }
        m->use_pdl = value != 0;
        return 0;
    }
    if (!strcmp(name, "mega")) { m->use_mega = value; return 0; }
    if (!strcmp(name, "ll_sleep")) return mega2_set_poll_sleep(value);
    if (!strcmp(name, "enc_graph")) { m->enc_graph = value; return 0; }
    if (!strcmp(name, "trace_cta")) { m->trace_cta = value; return 0; }
    if (!strcmp(name, "ll_debug")) return mega2_set_debug(value);
    if (!strcmp(name, "ll_reps")) {
        MB_REQUIRE(value >= 1 && value <= MEGA_LL_MAX_REPS && value * 2 <= 32, "ll_reps must be in [1, 16]");
        m->ll.reps = value;
        return 0;
    }
    if (!strcmp(name, "mega_trace")) {
        if (value) { MB_TRY(m->mega_trace.ensure(128 * 16 * 8)); MB_CUDA_CHECK(cudaMemset(m->mega_trace.p, 0, 128 * 16 * 8)); }
        return 0;
    }
    set_last_error(std::string("unknown option ") + name);
    return 2;
}

// Measurement hook for bench.py: replays the token step EAGERLY `iters` times on the state left by the last generate() call,
// bracketing every decode-path launch with CUDA events on the launching stream.  out_us[0..2] = device microseconds per token
// spent in {gemv, split-KV attention, logits/sample} kernels, out_us[3] = their launch counts packed as gemv*1e6 + attn*1e3 + sample.
extern "C" int mb200_model_profile_step(mb200_model* m, int32_t rows, int32_t B, int32_t max_length, int32_t iters, float* out_us, void* stream) {
    MB_REQUIRE(m && m->finalized && out_us && iters >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    GenState gs{};
    MB_CUDA_CHECK(cudaMemcpy(&gs, m->g_state.p, sizeof(gs), cudaMemcpyDeviceToHost));
    const int cur0 = gs.prompt_len + 1;
    double acc[3] = {0, 0, 0}; long long cnt[3] = {0, 0, 0};
    if (!g_prof.created) { for (auto& e : g_prof.ev) MB_CUDA_CHECK(cudaEventCreate(&e)); g_prof.created = true; }
    const int n_splits_self = self_splits(max_length);
    for (int it = 0; it < iters; ++it) {
        gs.cur_len = cur0 + it; gs.all_finished = 0; gs.n_finished = 0; gs.ticket = 0; gs.max_length = m->cfg.tgt_seq_len; gs.min_new_tokens = 0;
        MB_CUDA_CHECK(cudaMemcpy(m->g_state.p, &gs, sizeof(gs), cudaMemcpyHostToDevice));
        MB_CUDA_CHECK(cudaMemset(m->g_finished.p, 0, m->max_rows));
        g_prof.n = 0; g_prof.on = true;
        int s = token_step(m, rows, B, n_splits_self, st, false);
        g_prof.on = false;
        if (s) return s;
        MB_CUDA_CHECK(cudaStreamSynchronize(st));
        for (int i = 0; i < g_prof.n; ++i) {
            float ms = 0.f;
            MB_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
            acc[g_prof.cls[i]] += ms * 1000.0; cnt[g_prof.cls[i]]++;
        }
    }
    for (int c = 0; c < 3; ++c) out_us[c] = (float)(acc[c] / iters);
    out_us[3] = (float)((cnt[0] / iters) * 1000000LL + (cnt[1] / iters) * 1000LL + (cnt[2] / iters));
    return 0;
}

// debug: copy the megakernel phase trace (option "mega_trace") to host: out[n_phases][16] SM-cycle stamps
extern "C" int mb200_model_read_trace(mb200_model* m, uint64_t* out, int32_t n_phases) {
    MB_REQUIRE(m && out && m->mega_trace.p && n_phases <= 128, "trace not enabled");
    MB_CUDA_CHECK(cudaDeviceSynchronize());
    MB_CUDA_CHECK(cudaMemcpy(out, m->mega_trace.p, (size_t)n_phases * 16 * 8, cudaMemcpyDeviceToHost));
    return 0;
}

// Measurement hook: CUDA-event totals of the persistent token-loop kernel since the last reset:
// out[0] = launches, out[1] = total device milliseconds, out[2] = tokens decoded inside those launches.
extern "C" int mb200_model_mega_stats(mb200_model* m, double* out, int32_t reset) {
    MB_REQUIRE(m && out, "null argument");
    out[0] = (double)m->mega_launches; out[1] = m->mega_ms; out[2] = (double)m->mega_tokens;
    if (reset) { m->mega_launches = 0; m->mega_ms = 0.0; m->mega_tokens = 0; }
    return 0;
}

// Parity hook for the fused logits-processor chain (tests; not part of the reference-facing boundary): runs ONE selection step of
// `sample_body` on caller-supplied logits.  logits: DEVICE [rows, V] (rows = 2B under CFG, negative-prompt rows first);
// ids: HOST [B, L] the tokens so far (prompt + generated); `step` / `has_last_scores` select the look-back-bias state left by the
// previous call (scores are double-buffered by step parity, exactly as in generation).  Outputs: scores_out DEVICE [B, V] = the scores
// the selection sees (-inf where MinNewTokens / MonotonicTimeShift / LookbackBias / top-k / top-p removed the id), chosen_out HOST [B].
extern "C" int mb200_model_logits_chain(mb200_model* m, const float* logits, int32_t B, int32_t use_cfg, const int64_t* ids, int32_t L,
                                        int32_t prompt_len, const uint8_t* vflags, const mb200_generate_params* gp, int32_t step,
                                        int32_t has_last_scores, float* scores_out, int64_t* chosen_out, void* stream) {
    MB_REQUIRE(m && m->finalized && logits && ids && vflags && gp && scores_out && chosen_out, "null argument");
    const auto& c = m->cfg;
    const int rows = use_cfg ? 2 * B : B, V = c.vocab_size_out, ids_ld = c.tgt_seq_len;
    MB_REQUIRE(B >= 1 && rows <= m->max_rows && L >= 1 && L < ids_ld, "bad batch / length");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<long long> idsrow((size_t)B * ids_ld, (long long)gp->pad_token_id);
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < L; ++t) idsrow[(size_t)b * ids_ld + t] = ids[(size_t)b * L + t];
    GenState gs{};
    gs.cur_len = L; gs.prompt_len = prompt_len; gs.max_length = gp->max_length; gs.min_new_tokens = gp->min_new_tokens;
    gs.step = step; gs.has_last_scores = has_last_scores;
    const SampleConfig sc = make_sample_config(gp, B, use_cfg != 0, V, ids_ld);
    std::vector<int> zeros(rows, 0);
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_ids.p, idsrow.data(), idsrow.size() * 8, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_vflags.p, vflags, c.vocab_size_in, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_state.p, &gs, sizeof(gs), cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_cfg.p, &sc, sizeof(sc), cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->g_leftpad.p, zeros.data(), rows * 4, cudaMemcpyHostToDevice, st));
    MB_CUDA_CHECK(cudaMemsetAsync(m->g_finished.p, 0, m->max_rows, st));
    MB_CUDA_CHECK(cudaMemcpyAsync(m->d_logits.p, logits, (size_t)rows * V * 4, cudaMemcpyDeviceToDevice, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    MB_TRY(launch_prompt_scan(m->g_ids.as<long long>(), ids_ld, B, L, m->g_vflags.as<unsigned char>(), sc.ts_start, sc.ts_end, m->g_lastts.as<int>(), st));
    SampleParams sp = sample_params(m, rows);
    sp.dbg_scores = scores_out;
    MB_TRY(launch_sample(sp, B, st, false));
    std::vector<long long> chosen(B);
    MB_CUDA_CHECK(cudaMemcpy2DAsync(chosen.data(), 8, m->g_ids.as<long long>() + L, (size_t)ids_ld * 8, 8, B, cudaMemcpyDeviceToHost, st));
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) chosen_out[b] = chosen[b];
    return 0;
}
