// Host-side engine of the SD v1.x U-Net denoise step (include/sd_unet_b200.h): parameter store (diffusers names),
// activation buffers, and the kernel schedule of UNet2DConditionModel.forward expressed as
// launches of unet_gemm.cu (tcgen05 GEMM / implicit-GEMM conv) and unet_ops.cu (norms, softmax, elementwise).
// Everything is allocated and described once in finalize(); forward() only enqueues, so a step is CUDA-graph capturable.
#include "../../include/sd_unet_b200.h"
#include "tc_common.cuh"
#include "unet_gemm.h"
#include "unet_ops.h"
#include "unet_attn.h"
#include <cuda_bf16.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iterator>
#include <map>
#include <string>
#include <vector>

namespace {
thread_local char g_sd_err[512] = "";
void sd_err(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_sd_err, sizeof(g_sd_err), fmt, ap); va_end(ap);
}
#define SD_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { sd_err("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return (int)e_; } } while (0)

using bf16 = __nv_bfloat16;
using uce::GemmDesc;
struct Act { bf16* p; int n, h, w, c; long pixels() const { return (long)n * h * w; } };
struct Weight { std::vector<long> shape; bf16* b = nullptr; float* f = nullptr; long elems = 0; int kind = 0; };
// kind: 6 conv_in fp32 [ci][ky][kx][co]; 0 fp32 as is, 1 bf16 as is ([N,K] linear / 1x1 conv), 2 conv3x3 -> bf16 tap-major, 3 attention q/k/v (pad heads, bf16),
//       4 attention out projection (pad head columns, bf16), 5 conv_out fp32 [o][ky][kx][c]

int pad64(int x) { return (x + 63) / 64 * 64; }
}  // namespace

namespace uce { void sd_set_error(const char* msg) { snprintf(g_sd_err, sizeof(g_sd_err), "%s", msg); } }   // for vae_engine.cu: one sd_last_error() for both engines

struct sd_unet {
    sd_unet_config cfg;
    int device, NB, H, W, sm_count = 148, n_split = 0, n_fused_attn = 0;
    bool finalized = false;
    std::map<std::string, Weight> w;
    std::map<std::string, std::vector<long>> expected;          // name -> shape
    std::vector<void*> allocs;
    std::vector<std::function<int(cudaStream_t)>> ops;         // one denoise step
    std::vector<std::function<int(cudaStream_t)>> ctx_ops;     // depend on the text context (and attn2.to_k/to_v) only: run once per prompt
    bool ctx_set = false, ctx_dirty = true;
    std::vector<uce::TembJob> temb_jobs; uce::TembJob* temb_jobs_dev = nullptr; int temb_channels = 0;   // all time_emb_proj of a call: one launch
    std::map<std::string, Act> taps;
    // fixed I/O buffers
    float* x_in = nullptr; float* ctx_f32 = nullptr; bf16* ctx = nullptr; float* eps = nullptr;
    float* d_t = nullptr; float* h_t = nullptr;
    float* splitk_ws = nullptr; size_t splitk_cap = 0;
    static constexpr int MAX_GN = 256; int n_gn = 0;
    float* gn_stats = nullptr; float* S_scratch = nullptr; bf16* P_scratch = nullptr;
    bf16* temb_tap = nullptr;

    std::map<std::string, bf16*> qk_pool;          // self-attention block -> [2 HD, C] padded weights, to_q rows then to_k rows
    template <typename T> int alloc(T** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
        if (e != cudaSuccess) { sd_err("cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return (int)e; }
        allocs.push_back(q); *p = (T*)q; return 0;
    }
};

namespace {

// ------------------------------------------------------------------------------------------ parameter inventory
void expect_resnet(sd_unet* u, const std::string& p, int cin, int cout) {
    auto& e = u->expected; const int temb = u->cfg.temb_dim;
    e[p + ".norm1.weight"] = {cin}; e[p + ".norm1.bias"] = {cin};
    e[p + ".conv1.weight"] = {cout, cin, 3, 3}; e[p + ".conv1.bias"] = {cout};
    e[p + ".time_emb_proj.weight"] = {cout, temb}; e[p + ".time_emb_proj.bias"] = {cout};
    e[p + ".norm2.weight"] = {cout}; e[p + ".norm2.bias"] = {cout};
    e[p + ".conv2.weight"] = {cout, cout, 3, 3}; e[p + ".conv2.bias"] = {cout};
    if (cin != cout) { e[p + ".conv_shortcut.weight"] = {cout, cin, 1, 1}; e[p + ".conv_shortcut.bias"] = {cout}; }
}
void expect_tx(sd_unet* u, const std::string& p, int c) {
    auto& e = u->expected; const int ctx = u->cfg.cross_attention_dim;
    e[p + ".norm.weight"] = {c}; e[p + ".norm.bias"] = {c};
    e[p + ".proj_in.weight"] = {c, c, 1, 1}; e[p + ".proj_in.bias"] = {c};
    const std::string b = p + ".transformer_blocks.0";
    for (int i = 1; i <= 2; ++i) {
        const int kd = (i == 1) ? c : ctx;
        const std::string a = b + ".attn" + std::to_string(i);
        e[b + ".norm" + std::to_string(i) + ".weight"] = {c}; e[b + ".norm" + std::to_string(i) + ".bias"] = {c};
        e[a + ".to_q.weight"] = {c, c}; e[a + ".to_k.weight"] = {c, kd}; e[a + ".to_v.weight"] = {c, kd};
        e[a + ".to_out.0.weight"] = {c, c}; e[a + ".to_out.0.bias"] = {c};
    }
    e[b + ".norm3.weight"] = {c}; e[b + ".norm3.bias"] = {c};
    e[b + ".ff.net.0.proj.weight"] = {8 * c, c}; e[b + ".ff.net.0.proj.bias"] = {8 * c};
    e[b + ".ff.net.2.weight"] = {c, 4 * c}; e[b + ".ff.net.2.bias"] = {c};
    e[p + ".proj_out.weight"] = {c, c, 1, 1}; e[p + ".proj_out.bias"] = {c};
}
void build_inventory(sd_unet* u) {
    const sd_unet_config& c = u->cfg; auto& e = u->expected;
    const int* ch = c.block_out_channels; const int nl = c.n_levels, lpb = c.layers_per_block;
    e["conv_in.weight"] = {ch[0], c.in_channels, 3, 3}; e["conv_in.bias"] = {ch[0]};
    e["time_embedding.linear_1.weight"] = {c.temb_dim, ch[0]}; e["time_embedding.linear_1.bias"] = {c.temb_dim};
    e["time_embedding.linear_2.weight"] = {c.temb_dim, c.temb_dim}; e["time_embedding.linear_2.bias"] = {c.temb_dim};
    std::vector<int> skip{ch[0]};
    int cur = ch[0];
    for (int i = 0; i < nl; ++i) {
        for (int j = 0; j < lpb; ++j) {
            expect_resnet(u, "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), cur, ch[i]);
            cur = ch[i];
            if (c.down_has_attn[i]) expect_tx(u, "down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), cur);
            skip.push_back(cur);
        }
        if (i < nl - 1) {
            const std::string p = "down_blocks." + std::to_string(i) + ".downsamplers.0.conv";
            e[p + ".weight"] = {cur, cur, 3, 3}; e[p + ".bias"] = {cur};
            skip.push_back(cur);
        }
    }
    expect_resnet(u, "mid_block.resnets.0", cur, cur); expect_tx(u, "mid_block.attentions.0", cur); expect_resnet(u, "mid_block.resnets.1", cur, cur);
    for (int i = 0; i < nl; ++i) {
        const int cout = ch[nl - 1 - i];
        for (int j = 0; j < lpb + 1; ++j) {
            const int cin = cur + skip.back(); skip.pop_back();
            expect_resnet(u, "up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), cin, cout);
            cur = cout;
            if (c.up_has_attn[i]) expect_tx(u, "up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), cur);
        }
        if (i < nl - 1) {
            const std::string p = "up_blocks." + std::to_string(i) + ".upsamplers.0.conv";
            e[p + ".weight"] = {cur, cur, 3, 3}; e[p + ".bias"] = {cur};
        }
    }
    e["conv_norm_out.weight"] = {ch[0]}; e["conv_norm_out.bias"] = {ch[0]};
    e["conv_out.weight"] = {c.out_channels, ch[0], 3, 3}; e["conv_out.bias"] = {c.out_channels};
}

bool ends_with(const std::string& s, const char* suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

int weight_kind(const std::string& name, const std::vector<long>& shp) {
    if (name == "conv_in.weight") return 6;
    if (name == "conv_out.weight") return 5;
    if (shp.size() == 4) return shp[2] == 3 ? 2 : 1;
    if (shp.size() == 2) {
        if (name.find(".attn") != std::string::npos) {
            if (ends_with(name, "to_q.weight") || ends_with(name, "to_k.weight") || ends_with(name, "to_v.weight")) return 3;
            if (ends_with(name, "to_out.0.weight")) return 4;
        }
        return 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ schedule builder
struct Builder {
    sd_unet* u; int rc = 0;
    explicit Builder(sd_unet* uu) : u(uu) {}
    Act act(int n, int h, int w, int c) { Act a{nullptr, n, h, w, c}; if (!rc) rc = u->alloc(&a.p, (size_t)a.pixels() * c); return a; }
    const Weight& W(const std::string& n) { return u->w.at(n); }
    bool to_ctx = false;      // route the next pushes to the per-prompt list (cross-attention K / V^T of the text context)
    void push(std::function<int(cudaStream_t)> f) { (to_ctx ? u->ctx_ops : u->ops).push_back(std::move(f)); }
    void gemm(GemmDesc g) {
        if (!getenv("UCE_NO_PAIR") && uce::gemm_enable_pair(&g) < 0) { rc = rc ? rc : SD_E_STATE; sd_err("tensor map encode failed (pair)"); return; }
        int ks = uce::gemm_choose_ksplit(g, u->sm_count);
        while (ks > 1 && (size_t)ks * g.M * g.N > u->splitk_cap) --ks;
        if (ks > 1) {                         // every split GEMM shares one scratch: the schedule is a single in-order stream
            g.ksplit = ks; g.splitk_ws = u->splitk_ws;
            ++u->n_split;
        }
        if (!getenv("UCE_NO_TMA_EPI") && uce::gemm_enable_tma_epilogue(&g) < 0) { rc = rc ? rc : SD_E_STATE; sd_err("tensor map encode failed (epilogue)"); return; }
        g.stages = uce::gemm_choose_stages(g, u->sm_count, &g.katoms);
        push([g](cudaStream_t st) { return uce::gemm_launch(g, st); });
    }

    // out[M,N] = A[M,K] . Wt[N,K]^T (+bias) (+residual)
    void linear(const bf16* A, long M, int K, const std::string& wname, const float* bias, const bf16* residual, void* out, bool out_fp32, int N_override = 0) {
        const Weight& w = W(wname);
        const int N = N_override ? N_override : (int)(w.elems / K);
        GemmDesc g;
        if (uce::gemm_desc_linear(&g, A, K, 0, 0, w.b, K, 0, 0, (int)M, N, K, 1, 1, 0, 0)) { rc = rc ? rc : SD_E_STATE; sd_err("tensor map encode failed for %s", wname.c_str()); return; }
        g.out = out; g.out_fp32 = out_fp32; g.ldo = N; g.bias = bias; g.residual = residual; g.ldr = N;
        gemm(g);
    }
    void conv3(const Act& x, const std::string& wname, const float* bias, const float* rowbias, const bf16* residual, const Act& y, int stride) {
        const Weight& w = W(wname);
        GemmDesc g;
        const int r = uce::gemm_desc_conv(&g, x.p, x.n, x.h, x.w, x.c, w.b, y.c, 3, stride);
        if (r) { rc = rc ? rc : SD_E_STATE; sd_err("conv descriptor failed (%d) for %s: %dx%dx%d -> %d", r, wname.c_str(), x.h, x.w, x.c, y.c); return; }
        g.out = y.p; g.out_fp32 = 0; g.ldo = y.c; g.bias = bias; g.rowbias = rowbias; g.residual = residual; g.ldr = y.c;
        gemm(g);
    }
    void groupnorm(const Act& x, const Act& y, const std::string& p, float eps, int silu) {
        const float* ga = W(p + ".weight").f; const float* be = W(p + ".bias").f;
        const int G = u->cfg.norm_groups;
        if (u->n_gn >= sd_unet::MAX_GN) { rc = rc ? rc : SD_E_STATE; sd_err("too many GroupNorm calls"); return; }
        float* stats = u->gn_stats + (size_t)(u->n_gn++) * uce::op_groupnorm_ws_floats(u->NB, G);      // own workspace (per-CTA partial statistics; the apply kernel adds them)
        push([=](cudaStream_t st) { return uce::op_groupnorm(x.p, y.p, x.n, x.h * x.w, x.c, G, stats, ga, be, eps, silu, st); });
    }
    void layernorm(const bf16* x, bf16* y, long rows, int C, const std::string& p) {
        const float* ga = W(p + ".weight").f; const float* be = W(p + ".bias").f;
        push([=](cudaStream_t st) { return uce::op_layernorm(x, y, rows, C, ga, be, 1e-5f, st); });
    }

    Act resnet(const std::string& p, const Act& x, int cout, const bf16* st_emb) {
        Act a1 = act(x.n, x.h, x.w, x.c);
        groupnorm(x, a1, p + ".norm1", 1e-5f, 1);
        float* tproj = nullptr; if (!rc) rc = u->alloc(&tproj, (size_t)x.n * cout);
        (void)st_emb;         // projected for ALL resnets by the single op pushed after the time-embedding MLP (op_temb_proj_all)
        u->temb_jobs.push_back(uce::TembJob{W(p + ".time_emb_proj.weight").b, W(p + ".time_emb_proj.bias").f, tproj, cout, u->temb_channels});
        u->temb_channels += cout;
        Act h1 = act(x.n, x.h, x.w, cout);
        conv3(a1, p + ".conv1.weight", W(p + ".conv1.bias").f, tproj, nullptr, h1, 1);
        Act a2 = act(x.n, x.h, x.w, cout);
        groupnorm(h1, a2, p + ".norm2", 1e-5f, 1);
        const bf16* res = x.p;
        if (x.c != cout) {
            Act sc = act(x.n, x.h, x.w, cout);
            linear(x.p, x.pixels(), x.c, p + ".conv_shortcut.weight", W(p + ".conv_shortcut.bias").f, nullptr, sc.p, false);
            res = sc.p;
        }
        Act out = act(x.n, x.h, x.w, cout);
        conv3(a2, p + ".conv2.weight", W(p + ".conv2.bias").f, nullptr, res, out, 1);
        return out;
    }

    // softmax(q k^T / sqrt(dh)) v for all (image, head) pairs; q [NB*L, HD], keys/values from kv_src [NB*Lk, kdim]
    void attention(const std::string& a, const bf16* q_src, int C, long L, const bf16* kv_src, int Lk, int kdim, bf16* h /*residual, in place*/) {
        const int NB = u->NB, heads = u->cfg.heads, dh = C / heads, dhp = pad64(dh), HD = heads * dhp;
        const long M = (long)NB * L;
        const int Lkp = (Lk + 7) / 8 * 8;
        bf16 *q = nullptr, *k = nullptr, *vt = nullptr, *o = nullptr;
        // self-attention: q and k are projections of the same rows and their padded weights are adjacent (sd_unet_set_weight): ONE GEMM
        // writes [q | k] rows of 2 HD columns and the attention kernel reads both halves through strided tensor maps
        const bool qk_merged = kv_src == q_src && kdim == C && (long)Lk == L && W(a + ".to_k.weight").b == W(a + ".to_q.weight").b + (size_t)HD * C &&
                               !getenv("UCE_NO_QK_MERGE");
        const long ldq = qk_merged ? 2L * HD : HD, ldk = ldq;
        if (qk_merged) {
            if (!rc) rc = u->alloc(&q, (size_t)M * 2 * HD);
            k = q ? q + HD : nullptr;
        } else {
            if (!rc) rc = u->alloc(&q, (size_t)M * HD);
            if (!rc) rc = u->alloc(&k, (size_t)NB * Lk * HD);
        }
        if (!rc) rc = u->alloc(&vt, (size_t)NB * HD * Lkp);
        if (!rc) rc = u->alloc(&o, (size_t)M * HD);
        linear(q_src, M, C, a + ".to_q.weight", nullptr, nullptr, q, false, qk_merged ? 2 * HD : HD);
        // K and V^T of the TEXT context do not change between denoise steps (generate-images-sd.py:37-42 runs all steps of a row
        // with one prompt embedding): those projections go to the per-prompt list (sd_unet_set_context), not to the step.
        to_ctx = (kv_src == u->ctx);
        if (!qk_merged) linear(kv_src, (long)NB * Lk, kdim, a + ".to_k.weight", nullptr, nullptr, k, false, HD);
        {   // V^T[b] [HD, Lk] = Wv_pad [HD, kdim] . kv_src[b]^T
            GemmDesc g;
            if (uce::gemm_desc_linear(&g, W(a + ".to_v.weight").b, kdim, 0, 0, kv_src, kdim, (long)Lk * kdim, 0, HD, Lk, kdim, NB, 1, 0, 1)) { rc = rc ? rc : SD_E_STATE; to_ctx = false; return; }
            g.out = vt; g.out_fp32 = 0; g.ldo = Lkp; g.out_b1_stride = (long)HD * Lkp;
            gemm(g);
        }
        to_ctx = false;
        const bool fused = uce::attn_fused_supported(dhp) && !getenv("UCE_NO_FLASH");
        if (fused) {   // flash-style kernel: no score matrix in HBM
            uce::AttnDesc ad;
            if (uce::attn_desc_make(&ad, q, k, vt, o, NB, heads, dhp, L, Lk, Lkp, 1.f / sqrtf((float)dh), ldq, ldk)) { rc = rc ? rc : SD_E_STATE; sd_err("attention tensor maps failed"); return; }
            push([ad](cudaStream_t st) { return uce::attn_launch(ad, st); });
            ++u->n_fused_attn;
        } else {
        float* S = u->S_scratch; bf16* P = u->P_scratch;
        {   // S[b,h] [L, Lk] = q[b,:,h] k[b,:,h]^T / sqrt(dh)
            GemmDesc g;
            if (uce::gemm_desc_linear(&g, q, ldq, dhp, L * ldq, k, ldk, dhp, (long)Lk * ldk, (int)L, Lk, dhp, heads, NB, 1, 1)) { rc = rc ? rc : SD_E_STATE; return; }
            g.alpha = 1.f / sqrtf((float)dh);
            g.out = S; g.out_fp32 = 1; g.ldo = Lkp; g.out_b1_stride = L * Lkp; g.out_b2_stride = (long)heads * L * Lkp;
            gemm(g);
        }
        const long rows = (long)NB * heads * L;
        push([=](cudaStream_t st) { return uce::op_softmax(S, Lkp, P, Lkp, rows, Lk, st); });
        {   // o[b,:,h] [L, dhp] = P[b,h] [L, Lk] . Vt[b,h] [dhp, Lk]^T
            GemmDesc g;
            if (uce::gemm_desc_linear(&g, P, Lkp, L * Lkp, (long)heads * L * Lkp, vt, Lkp, (long)dhp * Lkp, (long)HD * Lkp, (int)L, dhp, Lk, heads, NB, 1, 1)) { rc = rc ? rc : SD_E_STATE; return; }
            g.out = o; g.out_fp32 = 0; g.ldo = HD; g.out_b1_stride = dhp; g.out_b2_stride = L * HD;
            gemm(g);
        }
        }
        // h += o Wo^T + bias   (in place: every element is read then written by the same thread)
        {
            const Weight& w = W(a + ".to_out.0.weight");
            GemmDesc g;
            if (uce::gemm_desc_linear(&g, o, HD, 0, 0, w.b, HD, 0, 0, (int)M, C, HD, 1, 1, 0, 0)) { rc = rc ? rc : SD_E_STATE; return; }
            g.out = h; g.out_fp32 = 0; g.ldo = C; g.bias = W(a + ".to_out.0.bias").f; g.residual = h; g.ldr = C;
            gemm(g);
        }
    }

    Act transformer(const std::string& p, const Act& x) {
        const int C = x.c, NB = x.n; const long L = (long)x.h * x.w, M = NB * L;
        Act g0 = act(x.n, x.h, x.w, C);
        groupnorm(x, g0, p + ".norm", 1e-6f, 0);
        Act h = act(x.n, x.h, x.w, C);
        linear(g0.p, M, C, p + ".proj_in.weight", W(p + ".proj_in.bias").f, nullptr, h.p, false);
        const std::string b = p + ".transformer_blocks.0";
        Act n = act(x.n, x.h, x.w, C);
        layernorm(h.p, n.p, M, C, b + ".norm1");
        attention(b + ".attn1", n.p, C, L, n.p, (int)L, C, h.p);
        layernorm(h.p, n.p, M, C, b + ".norm2");
        attention(b + ".attn2", n.p, C, L, u->ctx, u->cfg.context_len, u->cfg.cross_attention_dim, h.p);
        layernorm(h.p, n.p, M, C, b + ".norm3");
        bf16 *ff1 = nullptr, *gg = nullptr;
        if (!rc) rc = u->alloc(&ff1, (size_t)M * 8 * C);
        if (!rc) rc = u->alloc(&gg, (size_t)M * 4 * C);
        linear(n.p, M, C, b + ".ff.net.0.proj.weight", W(b + ".ff.net.0.proj.bias").f, nullptr, ff1, false);
        push([=](cudaStream_t st) { return uce::op_geglu(ff1, gg, M, 4 * C, st); });
        linear(gg, M, 4 * C, b + ".ff.net.2.weight", W(b + ".ff.net.2.bias").f, h.p, h.p, false);
        Act out = act(x.n, x.h, x.w, C);
        linear(h.p, M, C, p + ".proj_out.weight", W(p + ".proj_out.bias").f, x.p, out.p, false);
        return out;
    }
};

__global__ void f32_to_bf16_kernel(const float* x, bf16* y, long n) {
    uce::pdl_launch(); uce::pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2bfloat16(x[i]);
}
// the timestep of an eager call travels BY VALUE as a kernel argument
__global__ void set_scalar_kernel(float* dst, float v) { *dst = v; }
__global__ void timestep_from_dev_kernel(const float* t, int dim, int NB, bf16* out) {
    uce::pdl_launch(); uce::pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim / 2;
    if (i >= half) return;
    const float a = t[0] * expf(-logf(10000.f) * (float)i / (float)half);
    const bf16 c = __float2bfloat16(cosf(a)), s = __float2bfloat16(sinf(a));
    for (int n = 0; n < NB; ++n) { out[(long)n * dim + i] = c; out[(long)n * dim + half + i] = s; }
}

int build_schedule(sd_unet* u) {
    const sd_unet_config& c = u->cfg;
    const int NB = u->NB, H = u->H, W = u->W, nl = c.n_levels, lpb = c.layers_per_block;
    const int* ch = c.block_out_channels;
    Builder B(u);
    int rc;
    if ((rc = u->alloc(&u->x_in, (size_t)NB * c.in_channels * H * W))) return rc;
    if ((rc = u->alloc(&u->ctx_f32, (size_t)NB * c.context_len * c.cross_attention_dim))) return rc;
    if ((rc = u->alloc(&u->ctx, (size_t)NB * c.context_len * c.cross_attention_dim))) return rc;
    if ((rc = u->alloc(&u->eps, (size_t)NB * c.out_channels * H * W))) return rc;
    if ((rc = u->alloc(&u->d_t, 4))) return rc;
    SD_CUDA(cudaMallocHost((void**)&u->h_t, sizeof(float)));
    {
        const size_t gn_floats = (size_t)sd_unet::MAX_GN * uce::op_groupnorm_ws_floats(NB, c.norm_groups);
        if ((rc = u->alloc(&u->gn_stats, gn_floats))) return rc;
        SD_CUDA(cudaMemset(u->gn_stats, 0, gn_floats * sizeof(float)));      // (nothing in it needs a defined start any more; kept so that tools see initialised memory)
    }
    u->splitk_cap = (size_t)(3 * u->sm_count) * 128 * 128;      // gemm_choose_ksplit targets ~2 CTAs per SM of 128 x 128 partial tiles
    if ((rc = u->alloc(&u->splitk_ws, u->splitk_cap))) return rc;
    {   // attention scratch: largest (L x Lk) over the attention levels
        size_t mx = 0;
        for (int i = 0; i < nl; ++i) {
            const long L = (long)(H >> i) * (W >> i);
            const bool has = c.down_has_attn[i] || c.up_has_attn[nl - 1 - i] || (i == nl - 1);
            if (has) mx = std::max(mx, (size_t)L * std::max<long>((L + 7) / 8 * 8, (c.context_len + 7) / 8 * 8));
        }
        if ((rc = u->alloc(&u->S_scratch, (size_t)NB * c.heads * mx))) return rc;
        if ((rc = u->alloc(&u->P_scratch, (size_t)NB * c.heads * mx))) return rc;
    }
    // ---- inputs ----
    {
        float* cf = u->ctx_f32; bf16* cb = u->ctx; const long n = (long)NB * c.context_len * c.cross_attention_dim;
        B.to_ctx = true;
        B.push([=](cudaStream_t st) { return (int)uce::launch_k(f32_to_bf16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 1, cf, cb, n); });
        B.to_ctx = false;
    }
    // ---- time embedding ----
    bf16 *te0 = nullptr, *t1 = nullptr, *t1s = nullptr, *temb = nullptr, *st_emb = nullptr;
    if ((rc = u->alloc(&te0, (size_t)NB * ch[0]))) return rc;
    if ((rc = u->alloc(&t1, (size_t)NB * c.temb_dim))) return rc;
    if ((rc = u->alloc(&t1s, (size_t)NB * c.temb_dim))) return rc;
    if ((rc = u->alloc(&temb, (size_t)NB * c.temb_dim))) return rc;
    if ((rc = u->alloc(&st_emb, (size_t)NB * c.temb_dim))) return rc;
    {
        const float* dt = u->d_t; const int d0 = ch[0];
        B.push([=](cudaStream_t st) { return (int)uce::launch_k(timestep_from_dev_kernel, dim3((d0 / 2 + 127) / 128), dim3(128), 0, st, 1, dt, d0, NB, te0); });
    }
    B.linear(te0, NB, ch[0], "time_embedding.linear_1.weight", B.W("time_embedding.linear_1.bias").f, nullptr, t1, false);
    { const long n = (long)NB * c.temb_dim; B.push([=](cudaStream_t st) { return uce::op_silu(t1, t1s, n, st); }); }
    B.linear(t1s, NB, c.temb_dim, "time_embedding.linear_2.weight", B.W("time_embedding.linear_2.bias").f, nullptr, temb, false);
    { const long n = (long)NB * c.temb_dim; B.push([=](cudaStream_t st) { return uce::op_silu(temb, st_emb, n, st); }); }
    u->temb_tap = temb;
    // every resnet's time_emb_proj in one launch (the job table is filled while the resnets below are built, uploaded at the end)
    B.push([u, st_emb](cudaStream_t st) { return uce::op_temb_proj_all(u->temb_jobs_dev, (int)u->temb_jobs.size(), u->temb_channels, st_emb, u->NB, u->cfg.temb_dim, st); });
    // ---- conv_in ----
    Act h = B.act(NB, H, W, ch[0]);
    {
        const float* xin = u->x_in; const float* w = B.W("conv_in.weight").f; const float* bi = B.W("conv_in.bias").f; const int c0 = ch[0];
        if (c.in_channels != 4) { sd_err("conv_in kernel supports 4 input channels"); return SD_E_ARG; }
        B.push([=](cudaStream_t st) { return uce::op_conv_in(xin, w, bi, h.p, NB, H, W, c0, st); });
    }
    u->taps["conv_in"] = h;
    std::vector<Act> skips{h};
    for (int i = 0; i < nl; ++i) {
        for (int j = 0; j < lpb; ++j) {
            h = B.resnet("down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), h, ch[i], st_emb);
            if (c.down_has_attn[i]) h = B.transformer("down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), h);
            skips.push_back(h);
            u->taps["down." + std::to_string(i) + "." + std::to_string(j)] = h;
        }
        if (i < nl - 1) {
            const std::string p = "down_blocks." + std::to_string(i) + ".downsamplers.0.conv";
            Act d = B.act(NB, h.h / 2, h.w / 2, h.c);
            B.conv3(h, p + ".weight", B.W(p + ".bias").f, nullptr, nullptr, d, 2);
            h = d; skips.push_back(h);
        }
    }
    h = B.resnet("mid_block.resnets.0", h, h.c, st_emb);
    h = B.transformer("mid_block.attentions.0", h);
    h = B.resnet("mid_block.resnets.1", h, h.c, st_emb);
    u->taps["mid"] = h;
    for (int i = 0; i < nl; ++i) {
        const int cout = ch[nl - 1 - i];
        for (int j = 0; j < lpb + 1; ++j) {
            const Act s = skips.back(); skips.pop_back();
            Act cat = B.act(NB, h.h, h.w, h.c + s.c);
            { const Act hh = h; B.push([=](cudaStream_t st) { return uce::op_concat_c(hh.p, s.p, cat.p, hh.pixels(), hh.c, s.c, st); }); }
            h = B.resnet("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), cat, cout, st_emb);
            if (c.up_has_attn[i]) h = B.transformer("up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), h);
            u->taps["up." + std::to_string(i) + "." + std::to_string(j)] = h;
        }
        if (i < nl - 1) {
            const std::string p = "up_blocks." + std::to_string(i) + ".upsamplers.0.conv";
            Act up = B.act(NB, h.h * 2, h.w * 2, h.c);
            { const Act hh = h; B.push([=](cudaStream_t st) { return uce::op_upsample2x(hh.p, up.p, hh.n, hh.h, hh.w, hh.c, st); }); }
            Act o = B.act(NB, up.h, up.w, up.c);
            B.conv3(up, p + ".weight", B.W(p + ".bias").f, nullptr, nullptr, o, 1);
            h = o;
        }
    }
    Act a = B.act(NB, H, W, ch[0]);
    B.groupnorm(h, a, "conv_norm_out", 1e-5f, 1);
    {
        const float* w = B.W("conv_out.weight").f; const float* bi = B.W("conv_out.bias").f; float* e = u->eps; const int c0 = ch[0];
        if (c.out_channels != 4) { sd_err("conv_out kernel supports 4 output channels"); return SD_E_ARG; }
        B.push([=](cudaStream_t st) { return uce::op_conv_out(a.p, w, bi, e, NB, H, W, c0, st); });
    }
    if (!B.rc && !u->temb_jobs.empty()) {      // job table of the batched time-embedding projection
        if ((rc = u->alloc(&u->temb_jobs_dev, u->temb_jobs.size()))) return rc;
        SD_CUDA(cudaMemcpy(u->temb_jobs_dev, u->temb_jobs.data(), u->temb_jobs.size() * sizeof(uce::TembJob), cudaMemcpyHostToDevice));
    }
    return B.rc;
}

}  // namespace

extern "C" {

const char* sd_last_error(void) { return g_sd_err; }

int sd_unet_create(int device, const sd_unet_config* cfg, int batch, int H, int W, sd_unet** out) {
    if (!cfg || !out || batch <= 0 || H <= 0 || W <= 0 || cfg->n_levels < 1 || cfg->n_levels > 4) { sd_err("sd_unet_create: bad argument"); return SD_E_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); sd_err("CUDA device %d not available", device); return SD_E_DEVICE; }
    cudaDeviceProp prop;
    SD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { sd_err("device %d is sm_%d%d, need sm_100", device, prop.major, prop.minor); return SD_E_DEVICE; }
    if ((H % (1 << (cfg->n_levels - 1))) || (W % (1 << (cfg->n_levels - 1)))) { sd_err("latent size must be divisible by 2^(levels-1)"); return SD_E_ARG; }
    for (int i = 0; i < cfg->n_levels; ++i)
        if (cfg->block_out_channels[i] % 8 || (cfg->block_out_channels[i] / cfg->heads) * cfg->heads != cfg->block_out_channels[i]) { sd_err("channels must be multiples of 8 and of heads"); return SD_E_ARG; }
    SD_CUDA(cudaSetDevice(device));
    sd_unet* u = new sd_unet();
    u->cfg = *cfg; u->device = device; u->NB = batch; u->H = H; u->W = W; u->sm_count = prop.multiProcessorCount;
    build_inventory(u);
    *out = u;
    return 0;
}

int sd_unet_destroy(sd_unet* u) {
    if (!u) return 0;
    cudaSetDevice(u->device);
    cudaDeviceSynchronize();
    for (void* p : u->allocs) cudaFree(p);
    if (u->h_t) cudaFreeHost(u->h_t);
    delete u;
    return 0;
}

int sd_unet_set_weight(sd_unet* u, const char* name, const float* data, const long* shape, int ndim) {
    if (!u || !name || !data || !shape) { sd_err("sd_unet_set_weight: bad argument"); return SD_E_ARG; }
    auto it = u->expected.find(name);
    if (it == u->expected.end()) { sd_err("unknown parameter '%s'", name); return SD_E_WEIGHT; }
    std::vector<long> shp(shape, shape + ndim);
    if (shp != it->second) { sd_err("parameter '%s': shape mismatch", name); return SD_E_WEIGHT; }
    SD_CUDA(cudaSetDevice(u->device));
    long n = 1; for (long d : shp) n *= d;
    const int kind = weight_kind(name, shp);
    const int heads = u->cfg.heads;
    Weight& w = u->w[name];
    const bool fresh = w.elems == 0;
    u->ctx_dirty = true;        // cached K / V^T of the text context may depend on this parameter (attn2.to_k / to_v)
    std::vector<float> tmp;
    const float* src = data; long out_n = n;
    if (kind == 2) {            // [Cout][Cin][3][3] -> [Cout][ky][kx][Cin]
        const long co = shp[0], ci = shp[1];
        tmp.resize(n);
        for (long o = 0; o < co; ++o) for (long c = 0; c < ci; ++c) for (int k = 0; k < 9; ++k) tmp[(o * 9 + k) * ci + c] = data[(o * ci + c) * 9 + k];
        src = tmp.data();
    } else if (kind == 5) {     // conv_out [4][Cin][3][3] -> [4][ky][kx][Cin]
        const long co = shp[0], ci = shp[1];
        tmp.resize(n);
        for (long o = 0; o < co; ++o) for (long c = 0; c < ci; ++c) for (int k = 0; k < 9; ++k) tmp[(o * 9 + k) * ci + c] = data[(o * ci + c) * 9 + k];
        src = tmp.data();
    } else if (kind == 6) {     // conv_in [Cout][Cin][3][3] -> [Cin][ky][kx][Cout]
        const long co = shp[0], ci = shp[1];
        tmp.resize(n);
        for (long o = 0; o < co; ++o) for (long c = 0; c < ci; ++c) for (int k = 0; k < 9; ++k) tmp[(c * 9 + k) * co + o] = data[(o * ci + c) * 9 + k];
        src = tmp.data();
    } else if (kind == 3) {     // [C, K] -> [heads*dhp, K], zero rows for the head padding
        const long C = shp[0], K = shp[1]; const int dh = (int)(C / heads), dhp = pad64(dh);
        out_n = (long)heads * dhp * K; tmp.assign(out_n, 0.f);
        for (int h = 0; h < heads; ++h) for (int d = 0; d < dh; ++d) memcpy(&tmp[((long)h * dhp + d) * K], &data[((long)h * dh + d) * K], K * sizeof(float));
        src = tmp.data();
    } else if (kind == 4) {     // [C, C] -> [C, heads*dhp], zero columns for the head padding
        const long C = shp[0]; const int dh = (int)(C / heads), dhp = pad64(dh); const long HD = (long)heads * dhp;
        out_n = C * HD; tmp.assign(out_n, 0.f);
        for (long r = 0; r < C; ++r) for (int h = 0; h < heads; ++h) memcpy(&tmp[r * HD + (long)h * dhp], &data[r * C + (long)h * dh], dh * sizeof(float));
        src = tmp.data();
    }
    const bool as_bf16 = (kind >= 1 && kind <= 4);
    if (fresh) {
        w.shape = shp; w.kind = kind; w.elems = out_n;
        // the padded to_q and to_k of a SELF-attention share one allocation, q rows first: one GEMM over [2 HD, C] then projects both
        // (attention() checks the adjacency before it relies on it)
        const std::string nm(name);
        const size_t pq = nm.rfind(".attn1.to_q.weight"), pk = nm.rfind(".attn1.to_k.weight");
        if (kind == 3 && (pq != std::string::npos || pk != std::string::npos)) {
            const std::string key = nm.substr(0, pq != std::string::npos ? pq : pk);
            bf16*& pool = u->qk_pool[key];
            if (!pool) {
                int rc = u->alloc(&pool, (size_t)2 * out_n);
                if (rc) return rc;
                SD_CUDA(cudaMemset(pool, 0, (size_t)2 * out_n * sizeof(bf16)));
            }
            w.b = pool + (pk != std::string::npos ? out_n : 0);
        } else {
            int rc = as_bf16 ? u->alloc(&w.b, (size_t)out_n) : u->alloc(&w.f, (size_t)out_n);
            if (rc) return rc;
        }
    }
    if (as_bf16) {
        std::vector<bf16> hb(out_n);
        for (long i = 0; i < out_n; ++i) hb[i] = __float2bfloat16(src[i]);
        SD_CUDA(cudaMemcpy(w.b, hb.data(), out_n * sizeof(bf16), cudaMemcpyHostToDevice));
    } else {
        SD_CUDA(cudaMemcpy(w.f, src, out_n * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

int sd_unet_finalize(sd_unet* u) {
    if (!u) return SD_E_ARG;
    if (u->finalized) return 0;
    for (auto& kv : u->expected)
        if (!u->w.count(kv.first)) { sd_err("parameter '%s' was never set", kv.first.c_str()); return SD_E_WEIGHT; }
    SD_CUDA(cudaSetDevice(u->device));
    int rc = build_schedule(u);
    if (rc) return rc;
    u->finalized = true;
    return 0;
}

static int run_ctx_ops(sd_unet* u, cudaStream_t st) {
    int i = 0;
    for (auto& op : u->ctx_ops) {
        int rc = op(st);
        if (rc) { sd_err("U-Net context op %d failed: %s", i, rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "descriptor error"); return rc; }
        ++i;
    }
    u->ctx_dirty = false;
    return 0;
}

int sd_unet_set_context(sd_unet* u, const float* ctx, void* stream) {
    if (!u || !ctx) { sd_err("sd_unet_set_context: bad argument"); return SD_E_ARG; }
    if (!u->finalized) { sd_err("sd_unet_set_context before sd_unet_finalize"); return SD_E_STATE; }
    SD_CUDA(cudaSetDevice(u->device));
    cudaStream_t st = (cudaStream_t)stream;
    const sd_unet_config& c = u->cfg;
    SD_CUDA(cudaMemcpyAsync(u->ctx_f32, ctx, (size_t)u->NB * c.context_len * c.cross_attention_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
    u->ctx_set = true;
    return run_ctx_ops(u, st);
}

int sd_unet_forward(sd_unet* u, const float* x, float t, const float* ctx, float* eps, void* stream) {
    if (!u || !x || !eps) { sd_err("sd_unet_forward: bad argument"); return SD_E_ARG; }
    if (!u->finalized) { sd_err("sd_unet_forward before sd_unet_finalize"); return SD_E_STATE; }
    if (!ctx && !u->ctx_set) { sd_err("sd_unet_forward: ctx is NULL and sd_unet_set_context was never called"); return SD_E_STATE; }
    SD_CUDA(cudaSetDevice(u->device));
    cudaStream_t st = (cudaStream_t)stream;
    const sd_unet_config& c = u->cfg;
    // The timestep reaches the kernels through a device scalar so that a captured CUDA graph can be replayed for every step.
    // While capturing, the graph gets a copy node that reads the pinned scalar when the REPLAY executes (sd_unet_set_timestep
    // before each replay; the caller must not change it again before that replay has started).  Eager calls must not use that
    // scalar: the host runs many calls ahead of the device, and a copy that executes later would pick up the timestep of a later
    // call (found by tests/test_unet_gpu.py::test_fifty_step_guided_loop...: 51 calls enqueued back to back drifted 2.5x more than
    // the same calls issued one at a time).
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    SD_CUDA(cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone) {
        *u->h_t = t;
        SD_CUDA(cudaMemcpyAsync(u->d_t, u->h_t, sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        set_scalar_kernel<<<1, 1, 0, st>>>(u->d_t, t);
        SD_CUDA(cudaGetLastError());
    }
    SD_CUDA(cudaMemcpyAsync(u->x_in, x, (size_t)u->NB * c.in_channels * u->H * u->W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (ctx) {                       // context given with the call: (re)compute its K / V^T now
        int rc = sd_unet_set_context(u, ctx, stream);
        if (rc) return rc;
    } else if (u->ctx_dirty) {       // an attn2.to_k / to_v weight was overwritten since: recompute from the stored context
        int rc = run_ctx_ops(u, st);
        if (rc) return rc;
    }
    int i = 0;
    for (auto& op : u->ops) {
        int rc = op(st);
        if (rc) { sd_err("U-Net op %d failed: %s", i, rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "descriptor error"); return rc; }
        ++i;
    }
    SD_CUDA(cudaMemcpyAsync(eps, u->eps, (size_t)u->NB * c.out_channels * u->H * u->W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int sd_unet_set_timestep(sd_unet* u, float t) {
    if (!u || !u->h_t) { sd_err("sd_unet_set_timestep: engine not finalized"); return SD_E_STATE; }
    *u->h_t = t;
    return 0;
}

int sd_cfg_step(const float* eps2, long n, float gs, float* eps_out, const float* h1, const float* h2, const float* h3, const float c[4],
                float cx, float ce, const float* x_in, float* x_out, void* stream) {
    if (!eps2 || !c || !x_in || !x_out || n <= 0) { sd_err("sd_cfg_step: bad argument"); return SD_E_ARG; }
    int rc = uce::op_cfg_step(eps2, n, gs, eps_out, h1, h2, h3, c[0], c[1], c[2], c[3], cx, ce, x_in, x_out, (cudaStream_t)stream);
    if (rc) sd_err("cfg_step launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    return rc;
}

int sd_unet_inventory(const sd_unet_config* cfg, int index, char* name, size_t cap, long shape[4], int* ndim) {
    if (!cfg || cfg->n_levels < 1 || cfg->n_levels > 4 || cfg->layers_per_block < 1) { sd_err("sd_unet_inventory: bad argument"); return SD_E_ARG; }
    sd_unet tmp;                            // host-only: the inventory depends on the configuration alone
    tmp.cfg = *cfg;
    build_inventory(&tmp);
    const int n = (int)tmp.expected.size();
    if (index < 0) return n;
    if (index >= n || !name || !shape || !ndim) { sd_err("sd_unet_inventory: bad index / output"); return SD_E_ARG; }
    auto it = tmp.expected.begin();
    std::advance(it, index);
    if (it->first.size() + 1 > cap) { sd_err("sd_unet_inventory: name buffer too small"); return SD_E_ARG; }
    memcpy(name, it->first.c_str(), it->first.size() + 1);
    *ndim = (int)it->second.size();
    for (int i = 0; i < *ndim; ++i) shape[i] = it->second[i];
    return n;
}

int sd_unet_launch_count(sd_unet* u) { return u ? (int)u->ops.size() : SD_E_ARG; }
int sd_unet_context_launch_count(sd_unet* u) { return u ? (int)u->ctx_ops.size() : SD_E_ARG; }

int sd_unet_read_tap(sd_unet* u, const char* name, float* out, size_t cap, int dims[4]) {
    if (!u || !name || !out) return SD_E_ARG;
    SD_CUDA(cudaSetDevice(u->device));
    SD_CUDA(cudaDeviceSynchronize());
    if (std::string(name) == "temb") {
        const size_t n = (size_t)u->NB * u->cfg.temb_dim;
        if (cap < n) { sd_err("tap buffer too small"); return SD_E_ARG; }
        std::vector<bf16> hb(n);
        SD_CUDA(cudaMemcpy(hb.data(), u->temb_tap, n * sizeof(bf16), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) out[i] = __bfloat162float(hb[i]);
        if (dims) { dims[0] = u->NB; dims[1] = u->cfg.temb_dim; dims[2] = 1; dims[3] = 1; }
        return 0;
    }
    auto it = u->taps.find(name);
    if (it == u->taps.end()) { sd_err("unknown tap '%s'", name); return SD_E_ARG; }
    const Act a = it->second;
    const size_t n = (size_t)a.pixels() * a.c;
    if (cap < n) { sd_err("tap buffer too small"); return SD_E_ARG; }
    std::vector<bf16> hb(n);
    SD_CUDA(cudaMemcpy(hb.data(), a.p, n * sizeof(bf16), cudaMemcpyDeviceToHost));
    for (int nn = 0; nn < a.n; ++nn) for (int y = 0; y < a.h; ++y) for (int x = 0; x < a.w; ++x) for (int c = 0; c < a.c; ++c)
        out[(((size_t)nn * a.c + c) * a.h + y) * a.w + x] = __bfloat162float(hb[(((size_t)nn * a.h + y) * a.w + x) * a.c + c]);
    if (dims) { dims[0] = a.n; dims[1] = a.c; dims[2] = a.h; dims[3] = a.w; }
    return 0;
}

}  // extern "C"
