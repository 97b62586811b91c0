// Host-side engine of the SD v1.x VAE decoder (include/sd_vae_b200.h): parameter store (diffusers names of pipe.vae), a pooled
// set of activation buffers, and the kernel schedule of AutoencoderKL.decode — post_quant_conv, Decoder.conv_in, the mid block
// (resnet, one-head attention over the h*w tokens, resnet), the up blocks (layers_per_block + 1 resnets, nearest 2x upsample +
// conv), GroupNorm + SiLU, conv_out — expressed as launches of the kernels the U-Net step already uses: unet_gemm.cu (tcgen05 GEMM /
// implicit-GEMM convolution) and unet_ops.cu (GroupNorm, softmax, upsample, the 4-channel edge convolutions).  The reference reaches
// this through `pipe(...)` (evalscripts/generate-images-sd.py:37-46; explicit in evalscripts/concept_algebra.py:126-135).
// Everything is allocated and described once in finalize(); decode() only enqueues.
#include "../../include/sd_vae_b200.h"
#include "tc_common.cuh"
#include "unet_gemm.h"
#include "unet_ops.h"
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

namespace uce { void sd_set_error(const char* msg); }       // unet_engine.cu: the message sd_last_error() returns

namespace {
void vae_err(const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    uce::sd_set_error(buf);
}
#define VAE_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { vae_err("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return (int)e_; } } while (0)

using bf16 = __nv_bfloat16;
using uce::GemmDesc;
struct Act { bf16* p; int n, h, w, c; long pixels() const { return (long)n * h * w; } size_t elems() const { return (size_t)pixels() * c; } };
struct Weight { std::vector<long> shape; bf16* b = nullptr; float* f = nullptr; long elems = 0; };
enum Kind { K_F32 = 0, K_BF16 = 1, K_CONV3 = 2, K_CONV_IN = 3, K_CONV_OUT = 4 };
}  // namespace

struct sd_vae {
    sd_vae_config cfg;
    int device, NB, h, w, H, W, sm_count = 148;
    bool finalized = false;
    std::map<std::string, Weight> wt;
    std::map<std::string, std::vector<long>> expected;
    std::map<std::string, std::vector<float>> host_keep;        // fp32 host copies of the tensors the attention bias fold needs
    std::vector<void*> allocs;
    std::vector<std::pair<size_t, void*>> pool;                 // released activation buffers (bytes, pointer)
    std::vector<std::function<int(cudaStream_t)>> ops;
    std::map<std::string, Act> taps;
    float *z_in = nullptr, *z_pq = nullptr, *img4 = nullptr;    // latents, after post_quant_conv, conv_out result [NB,4,H,W]
    float* attn_bias = nullptr;                                 // to_out.0.bias + to_out.0.weight . to_v.bias
    float* splitk_ws = nullptr; size_t splitk_cap = 0;
    static constexpr int MAX_GN = 64; int n_gn = 0;
    float* gn_stats = nullptr; float* S_scratch = nullptr; bf16* P_scratch = nullptr;
    size_t act_bytes = 0;

    template <typename T> int alloc(T** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
        if (e != cudaSuccess) { vae_err("cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return (int)e; }
        allocs.push_back(q); *p = (T*)q; return 0;
    }
};

namespace {

// ------------------------------------------------------------------------------------------ parameter inventory
void expect_resnet(sd_vae* v, const std::string& p, int cin, int cout) {
    auto& e = v->expected;
    e[p + ".norm1.weight"] = {cin}; e[p + ".norm1.bias"] = {cin};
    e[p + ".conv1.weight"] = {cout, cin, 3, 3}; e[p + ".conv1.bias"] = {cout};
    e[p + ".norm2.weight"] = {cout}; e[p + ".norm2.bias"] = {cout};
    e[p + ".conv2.weight"] = {cout, cout, 3, 3}; e[p + ".conv2.bias"] = {cout};
    if (cin != cout) { e[p + ".conv_shortcut.weight"] = {cout, cin, 1, 1}; e[p + ".conv_shortcut.bias"] = {cout}; }
}
void build_inventory(sd_vae* v) {
    const sd_vae_config& c = v->cfg; auto& e = v->expected;
    const int nl = c.n_levels, top = c.block_out_channels[nl - 1], lat = c.latent_channels;
    e["post_quant_conv.weight"] = {lat, lat, 1, 1}; e["post_quant_conv.bias"] = {lat};
    e["decoder.conv_in.weight"] = {top, lat, 3, 3}; e["decoder.conv_in.bias"] = {top};
    expect_resnet(v, "decoder.mid_block.resnets.0", top, top);
    const std::string a = "decoder.mid_block.attentions.0";
    e[a + ".group_norm.weight"] = {top}; e[a + ".group_norm.bias"] = {top};
    for (const char* n : {"to_q", "to_k", "to_v", "to_out.0"}) { e[a + "." + n + ".weight"] = {top, top}; e[a + "." + n + ".bias"] = {top}; }
    expect_resnet(v, "decoder.mid_block.resnets.1", top, top);
    int cur = top;
    for (int i = 0; i < nl; ++i) {
        const int cout = c.block_out_channels[nl - 1 - i];
        for (int j = 0; j < c.layers_per_block + 1; ++j) { expect_resnet(v, "decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), cur, cout); cur = cout; }
        if (i != nl - 1) {
            const std::string p = "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv";
            e[p + ".weight"] = {cout, cout, 3, 3}; e[p + ".bias"] = {cout};
        }
    }
    e["decoder.conv_norm_out.weight"] = {cur}; e["decoder.conv_norm_out.bias"] = {cur};
    e["decoder.conv_out.weight"] = {c.out_channels, cur, 3, 3}; e["decoder.conv_out.bias"] = {c.out_channels};
}

Kind weight_kind(const std::string& name, const std::vector<long>& shp) {
    if (name == "decoder.conv_in.weight") return K_CONV_IN;
    if (name == "decoder.conv_out.weight") return K_CONV_OUT;
    if (name.rfind("post_quant_conv", 0) == 0) return K_F32;
    if (shp.size() == 4) return shp[2] == 3 ? K_CONV3 : K_BF16;
    if (shp.size() == 2) return K_BF16;
    return K_F32;
}

// ------------------------------------------------------------------------------------------ the two kernels of the edges
// post_quant_conv on the scaled latents: y[n,o,p] = b[o] + sum_c W[o,c] * x[n,c,p] / scaling_factor   (1x1, 4 -> 4, fp32 NCHW)
__global__ void __launch_bounds__(256) post_quant_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                         float inv_sf, float* __restrict__ y, int NB, long HW) {
    uce::pdl_launch(); uce::pdl_wait();
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)NB * HW) return;
    const long n = idx / HW, p = idx % HW;
    float in[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) in[c] = x[(n * 4 + c) * HW + p] * inv_sf;
#pragma unroll
    for (int o = 0; o < 4; ++o) y[(n * 4 + o) * HW + p] = b[o] + w[o * 4] * in[0] + w[o * 4 + 1] * in[1] + w[o * 4 + 2] * in[2] + w[o * 4 + 3] * in[3];
}
// rgb8[n,y,x,c] = round_half_even(clamp(img[n,c,y,x] / 2 + 0.5, 0, 1) * 255)   (img has 4 planes per image; the 4th is padding)
__global__ void __launch_bounds__(256) to_rgb8_kernel(const float* __restrict__ img4, unsigned char* __restrict__ rgb, int NB, long HW) {
    uce::pdl_launch(); uce::pdl_wait();
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)NB * HW) return;
    const long n = idx / HW, p = idx % HW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = fminf(fmaxf(img4[(n * 4 + c) * HW + p] * 0.5f + 0.5f, 0.f), 1.f);
        rgb[idx * 3 + c] = (unsigned char)rintf(v * 255.f);
    }
}

// ------------------------------------------------------------------------------------------ schedule builder
struct Builder {
    sd_vae* v; int rc = 0;
    explicit Builder(sd_vae* vv) : v(vv) {}
    // The schedule is one in-order stream, so a buffer can be handed out again as soon as every op that reads it has been PUSHED.
    Act act(int n, int h, int w, int c) {
        Act a{nullptr, n, h, w, c};
        const size_t bytes = a.elems() * sizeof(bf16);
        for (size_t i = 0; i < v->pool.size(); ++i)
            if (v->pool[i].first == bytes) { a.p = (bf16*)v->pool[i].second; v->pool.erase(v->pool.begin() + i); return a; }
        if (!rc) { rc = v->alloc(&a.p, a.elems()); v->act_bytes += bytes; }
        return a;
    }
    void release(const Act& a) { if (a.p) v->pool.emplace_back(a.elems() * sizeof(bf16), (void*)a.p); }
    const Weight& W(const std::string& n) { return v->wt.at(n); }
    void push(std::function<int(cudaStream_t)> f) { v->ops.push_back(std::move(f)); }
    void fail(const char* what, const std::string& name) { if (!rc) { rc = SD_E_STATE; vae_err("%s failed for %s", what, name.c_str()); } }
    // weights_b = false: the B operand is an ACTIVATION (attention scores / values).  The CTA-pair kernel requests its B tiles before
    // the programmatic-dependency wait and ignores batch coordinates, and the split-K grid reuses grid.z: both are for weight operands
    // only, so those GEMMs stay on the plain kernel (which orders a batched B after the wait).
    void gemm(GemmDesc g, bool weights_b = true) {
        if (weights_b && !getenv("UCE_NO_PAIR") && uce::gemm_enable_pair(&g) < 0) { fail("tensor map encode (pair)", ""); return; }
        int ks = weights_b ? uce::gemm_choose_ksplit(g, v->sm_count) : 1;
        while (ks > 1 && (size_t)ks * g.M * g.N > v->splitk_cap) --ks;
        if (ks > 1) { g.ksplit = ks; g.splitk_ws = v->splitk_ws; }
        if (!getenv("UCE_NO_TMA_EPI") && uce::gemm_enable_tma_epilogue(&g) < 0) { fail("tensor map encode (epilogue)", ""); return; }
        g.stages = uce::gemm_choose_stages(g, v->sm_count, &g.katoms);
        push([g](cudaStream_t st) { return uce::gemm_launch(g, st); });
    }
    // out[M,N] = A[M,K] . Wt[N,K]^T (+bias) (+residual)
    void linear(const bf16* A, long M, int K, const std::string& wname, const float* bias, const bf16* residual, bf16* out) {
        const Weight& w = W(wname);
        const int N = (int)(w.elems / K);
        GemmDesc g;
        if (uce::gemm_desc_linear(&g, A, K, 0, 0, w.b, K, 0, 0, (int)M, N, K, 1, 1, 0, 0)) { fail("tensor map encode", wname); return; }
        g.out = out; g.out_fp32 = 0; g.ldo = N; g.bias = bias; g.residual = residual; g.ldr = N;
        gemm(g);
    }
    void conv3(const Act& x, const std::string& p, const bf16* residual, const Act& y) {
        GemmDesc g;
        if (uce::gemm_desc_conv(&g, x.p, x.n, x.h, x.w, x.c, W(p + ".weight").b, y.c, 3, 1)) { fail("conv descriptor", p); return; }
        g.out = y.p; g.out_fp32 = 0; g.ldo = y.c; g.bias = W(p + ".bias").f; g.residual = residual; g.ldr = y.c;
        gemm(g);
    }
    void groupnorm(const Act& x, const Act& y, const std::string& p, int silu) {
        const float* ga = W(p + ".weight").f; const float* be = W(p + ".bias").f;
        const int G = v->cfg.norm_groups;
        if (v->n_gn >= sd_vae::MAX_GN) { fail("GroupNorm slot", p); return; }
        float* stats = v->gn_stats + (size_t)(v->n_gn++) * uce::op_groupnorm_ws_floats(v->NB, G);      // own workspace (per-CTA partial statistics; the apply kernel adds them)
        push([=](cudaStream_t st) { return uce::op_groupnorm(x.p, y.p, x.n, x.h * x.w, x.c, G, stats, ga, be, 1e-6f, silu, st); });
    }
    // ResnetBlock2D without time embedding (temb_channels=None in the VAE), eps 1e-6.  Consumes x.
    Act resnet(const std::string& p, const Act& x, int cout) {
        Act a1 = act(x.n, x.h, x.w, x.c);
        groupnorm(x, a1, p + ".norm1", 1);
        Act h1 = act(x.n, x.h, x.w, cout);
        conv3(a1, p + ".conv1", nullptr, h1);
        release(a1);
        Act a2 = act(x.n, x.h, x.w, cout);
        groupnorm(h1, a2, p + ".norm2", 1);
        release(h1);
        Act res = x;
        if (x.c != cout) {
            res = act(x.n, x.h, x.w, cout);
            linear(x.p, x.pixels(), x.c, p + ".conv_shortcut.weight", W(p + ".conv_shortcut.bias").f, nullptr, res.p);
            release(x);
        }
        Act out = act(x.n, x.h, x.w, cout);
        conv3(a2, p + ".conv2", res.p, out);
        release(a2); release(res);
        return out;
    }
    // Attention block of the mid level: GroupNorm, one head of width C over the h*w tokens, residual.  Consumes x.
    //   softmax(q k^T / sqrt(C)) (v0 + 1 bv^T) Wo^T + bo  =  softmax(.) v0 Wo^T + (bo + Wo bv)   (rows of the softmax sum to 1),
    // so V^T is produced directly by a GEMM without its bias and the folded bias (sd_vae::attn_bias) is added by the last one.
    Act attention(const std::string& a, const Act& x) {
        const int NB = x.n, C = x.c; const long L = (long)x.h * x.w, M = NB * L;
        const long Lp = (L + 7) / 8 * 8;
        Act g0 = act(x.n, x.h, x.w, C);
        groupnorm(x, g0, a + ".group_norm", 0);
        Act q = act(x.n, x.h, x.w, C), k = act(x.n, x.h, x.w, C), o = act(x.n, x.h, x.w, C);
        bf16* vt = nullptr; if (!rc) rc = v->alloc(&vt, (size_t)NB * C * Lp);
        linear(g0.p, M, C, a + ".to_q.weight", W(a + ".to_q.bias").f, nullptr, q.p);
        linear(g0.p, M, C, a + ".to_k.weight", W(a + ".to_k.bias").f, nullptr, k.p);
        {   // V^T[b] [C, L] = Wv [C, C] . g0[b]^T
            GemmDesc g;
            if (uce::gemm_desc_linear(&g, W(a + ".to_v.weight").b, C, 0, 0, g0.p, C, L * C, 0, C, (int)L, C, NB, 1, 0, 1)) { fail("tensor map encode", a + ".to_v"); return x; }
            g.out = vt; g.out_fp32 = 0; g.ldo = Lp; g.out_b1_stride = (long)C * Lp;
            gemm(g, false);
        }
        release(g0);
        float* S = v->S_scratch; bf16* P = v->P_scratch;
        {   // S[b] [L, L] = q[b] k[b]^T / sqrt(C)
            GemmDesc g;
            if (uce::gemm_desc_linear(&g, q.p, C, C, L * C, k.p, C, C, L * C, (int)L, (int)L, C, 1, NB, 1, 1)) { fail("tensor map encode", a + " scores"); return x; }
            g.alpha = 1.f / sqrtf((float)C);
            g.out = S; g.out_fp32 = 1; g.ldo = Lp; g.out_b1_stride = L * Lp; g.out_b2_stride = L * Lp;
            gemm(g, false);
        }
        push([=](cudaStream_t st) { return uce::op_softmax(S, Lp, P, Lp, M, (int)L, st); });
        {   // o[b] [L, C] = P[b] [L, L] . Vt[b] [C, L]^T
            GemmDesc g;
            if (uce::gemm_desc_linear(&g, P, Lp, L * Lp, L * Lp, vt, Lp, (long)C * Lp, (long)C * Lp, (int)L, C, (int)L, 1, NB, 1, 1)) { fail("tensor map encode", a + " PV"); return x; }
            g.out = o.p; g.out_fp32 = 0; g.ldo = C; g.out_b1_stride = C; g.out_b2_stride = L * C;
            gemm(g, false);
        }
        release(q); release(k);
        Act out = act(x.n, x.h, x.w, C);
        linear(o.p, M, C, a + ".to_out.0.weight", v->attn_bias, x.p, out.p);
        release(o); release(x);
        return out;
    }
};

int build_schedule(sd_vae* v) {
    const sd_vae_config& c = v->cfg;
    const int NB = v->NB, nl = c.n_levels, top = c.block_out_channels[nl - 1];
    int rc;
    Builder B(v);
    const long hw = (long)v->h * v->w, HW = (long)v->H * v->W;
    if ((rc = v->alloc(&v->z_in, (size_t)NB * 4 * hw))) return rc;
    if ((rc = v->alloc(&v->z_pq, (size_t)NB * 4 * hw))) return rc;
    if ((rc = v->alloc(&v->img4, (size_t)NB * 4 * HW))) return rc;
    {
        const size_t gn_floats = (size_t)sd_vae::MAX_GN * uce::op_groupnorm_ws_floats(NB, c.norm_groups);
        if ((rc = v->alloc(&v->gn_stats, gn_floats))) return rc;
        VAE_CUDA(cudaMemset(v->gn_stats, 0, gn_floats * sizeof(float)));      // (nothing in it needs a defined start any more; kept so that tools see initialised memory)
    }
    v->splitk_cap = (size_t)(3 * v->sm_count) * 128 * 128;
    if ((rc = v->alloc(&v->splitk_ws, v->splitk_cap))) return rc;
    {
        const size_t Lp = (size_t)(hw + 7) / 8 * 8;
        if ((rc = v->alloc(&v->S_scratch, (size_t)NB * hw * Lp))) return rc;
        if ((rc = v->alloc(&v->P_scratch, (size_t)NB * hw * Lp))) return rc;
    }
    {   // folded bias of the attention output projection
        const std::string a = "decoder.mid_block.attentions.0";
        const std::vector<float>&wo = v->host_keep.at(a + ".to_out.0.weight"), &bo = v->host_keep.at(a + ".to_out.0.bias"), &bv = v->host_keep.at(a + ".to_v.bias");
        std::vector<float> fb(top);
        for (int i = 0; i < top; ++i) {
            double s = bo[i];
            for (int j = 0; j < top; ++j) s += (double)wo[(size_t)i * top + j] * bv[j];
            fb[i] = (float)s;
        }
        if ((rc = v->alloc(&v->attn_bias, (size_t)top))) return rc;
        VAE_CUDA(cudaMemcpy(v->attn_bias, fb.data(), top * sizeof(float), cudaMemcpyHostToDevice));
    }
    {
        const float* zi = v->z_in; float* zo = v->z_pq; const float* w = B.W("post_quant_conv.weight").f; const float* b = B.W("post_quant_conv.bias").f;
        const float inv_sf = 1.f / c.scaling_factor;
        B.push([=](cudaStream_t st) { return (int)uce::launch_k(post_quant_kernel, dim3((unsigned)((NB * hw + 255) / 256)), dim3(256), 0, st, 1, zi, w, b, inv_sf, zo, NB, hw); });
    }
    Act h = B.act(NB, v->h, v->w, top);
    {
        const float* zi = v->z_pq; const float* w = B.W("decoder.conv_in.weight").f; const float* bi = B.W("decoder.conv_in.bias").f;
        const int hh = v->h, ww = v->w;
        B.push([=](cudaStream_t st) { return uce::op_conv_in(zi, w, bi, h.p, NB, hh, ww, top, st); });
    }
    h = B.resnet("decoder.mid_block.resnets.0", h, top);
    h = B.attention("decoder.mid_block.attentions.0", h);
    h = B.resnet("decoder.mid_block.resnets.1", h, top);
    auto tap = [&](const std::string& name, const Act& a) {      // a tapped buffer must never return to the pool: copy it out
        if (!getenv("UCE_VAE_TAPS")) return;
        Act t{nullptr, a.n, a.h, a.w, a.c};
        if (!B.rc) B.rc = v->alloc(&t.p, t.elems());
        const Act src = a;
        B.push([=](cudaStream_t st) { return (int)cudaMemcpyAsync(t.p, src.p, src.elems() * sizeof(bf16), cudaMemcpyDeviceToDevice, st); });
        v->taps[name] = t;
    };
    tap("mid", h);
    for (int i = 0; i < nl; ++i) {
        const int cout = c.block_out_channels[nl - 1 - i];
        for (int j = 0; j < c.layers_per_block + 1; ++j) h = B.resnet("decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), h, cout);
        if (i != nl - 1) {
            Act up = B.act(NB, h.h * 2, h.w * 2, h.c);
            { const Act hh = h; B.push([=](cudaStream_t st) { return uce::op_upsample2x(hh.p, up.p, hh.n, hh.h, hh.w, hh.c, st); }); }
            B.release(h);
            Act o = B.act(NB, up.h, up.w, up.c);
            B.conv3(up, "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv", nullptr, o);
            B.release(up);
            h = o;
        }
        tap("up." + std::to_string(i), h);
    }
    Act a = B.act(NB, h.h, h.w, h.c);
    B.groupnorm(h, a, "decoder.conv_norm_out", 1);
    B.release(h);
    {
        const float* w = B.W("decoder.conv_out.weight").f; const float* bi = B.W("decoder.conv_out.bias").f; float* img = v->img4;
        const int HH = v->H, WW = v->W, cc = a.c;
        B.push([=](cudaStream_t st) { return uce::op_conv_out(a.p, w, bi, img, NB, HH, WW, cc, st); });
    }
    return B.rc;
}

}  // namespace

extern "C" {

int sd_vae_create(int device, const sd_vae_config* cfg, int batch, int h, int w, sd_vae** out) {
    if (!cfg || !out || batch <= 0 || h <= 0 || w <= 0 || cfg->n_levels < 1 || cfg->n_levels > 4 || cfg->layers_per_block < 1) { vae_err("sd_vae_create: bad argument"); return SD_E_ARG; }
    *out = nullptr;
    if (cfg->latent_channels != 4 || cfg->out_channels != 3) { vae_err("the edge convolutions support 4 latent and 3 image channels"); return SD_E_ARG; }
    if (!(cfg->scaling_factor > 0.f)) { vae_err("scaling_factor must be positive"); return SD_E_ARG; }
    for (int i = 0; i < cfg->n_levels; ++i)
        if (cfg->block_out_channels[i] % 8 || cfg->block_out_channels[i] % cfg->norm_groups) { vae_err("channels must be multiples of 8 and of norm_groups"); return SD_E_ARG; }
    if (cfg->block_out_channels[cfg->n_levels - 1] % 64) { vae_err("the attention width (last block_out_channels) must be a multiple of 64"); return SD_E_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); vae_err("CUDA device %d not available", device); return SD_E_DEVICE; }
    cudaDeviceProp prop;
    VAE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { vae_err("device %d is sm_%d%d, need sm_100", device, prop.major, prop.minor); return SD_E_DEVICE; }
    VAE_CUDA(cudaSetDevice(device));
    sd_vae* v = new sd_vae();
    v->cfg = *cfg; v->device = device; v->NB = batch; v->h = h; v->w = w;
    v->H = h << (cfg->n_levels - 1); v->W = w << (cfg->n_levels - 1); v->sm_count = prop.multiProcessorCount;
    build_inventory(v);
    *out = v;
    return 0;
}

int sd_vae_destroy(sd_vae* v) {
    if (!v) return 0;
    cudaSetDevice(v->device);
    cudaDeviceSynchronize();
    for (void* p : v->allocs) cudaFree(p);
    delete v;
    return 0;
}

int sd_vae_set_weight(sd_vae* v, const char* name, const float* data, const long* shape, int ndim) {
    if (!v || !name || !data || !shape) { vae_err("sd_vae_set_weight: bad argument"); return SD_E_ARG; }
    if (v->finalized) { vae_err("sd_vae_set_weight after sd_vae_finalize"); return SD_E_STATE; }
    auto it = v->expected.find(name);
    if (it == v->expected.end()) { vae_err("unknown parameter '%s'", name); return SD_E_WEIGHT; }
    std::vector<long> shp(shape, shape + ndim);
    if (shp != it->second) { vae_err("parameter '%s': shape mismatch", name); return SD_E_WEIGHT; }
    VAE_CUDA(cudaSetDevice(v->device));
    long n = 1; for (long d : shp) n *= d;
    const Kind kind = weight_kind(name, shp);
    const std::string nm(name);
    if (nm == "decoder.mid_block.attentions.0.to_out.0.weight" || nm == "decoder.mid_block.attentions.0.to_out.0.bias" ||
        nm == "decoder.mid_block.attentions.0.to_v.bias")
        v->host_keep[nm].assign(data, data + n);
    Weight& w = v->wt[nm];
    const bool fresh = w.elems == 0;
    std::vector<float> tmp;
    const float* src = data; long out_n = n;
    if (kind == K_CONV3) {                  // [Cout][Cin][3][3] -> [Cout][ky][kx][Cin]
        const long co = shp[0], ci = shp[1];
        tmp.resize(n);
        for (long o = 0; o < co; ++o) for (long c = 0; c < ci; ++c) for (int k = 0; k < 9; ++k) tmp[(o * 9 + k) * ci + c] = data[(o * ci + c) * 9 + k];
        src = tmp.data();
    } else if (kind == K_CONV_OUT) {        // [3][Cin][3][3] -> [4][ky][kx][Cin], the 4th output plane is zero padding
        const long co = shp[0], ci = shp[1];
        out_n = 4 * 9 * ci; tmp.assign(out_n, 0.f);
        for (long o = 0; o < co; ++o) for (long c = 0; c < ci; ++c) for (int k = 0; k < 9; ++k) tmp[(o * 9 + k) * ci + c] = data[(o * ci + c) * 9 + k];
        src = tmp.data();
    } else if (kind == K_CONV_IN) {         // [Cout][4][3][3] -> [4][ky][kx][Cout]
        const long co = shp[0], ci = shp[1];
        tmp.resize(n);
        for (long o = 0; o < co; ++o) for (long c = 0; c < ci; ++c) for (int k = 0; k < 9; ++k) tmp[(c * 9 + k) * co + o] = data[(o * ci + c) * 9 + k];
        src = tmp.data();
    } else if (nm == "decoder.conv_out.bias") {
        out_n = 4; tmp.assign(4, 0.f);
        for (long i = 0; i < n; ++i) tmp[i] = data[i];
        src = tmp.data();
    }
    const bool as_bf16 = (kind == K_BF16 || kind == K_CONV3);
    if (fresh) {
        w.shape = shp; w.elems = out_n;
        int rc = as_bf16 ? v->alloc(&w.b, (size_t)out_n) : v->alloc(&w.f, (size_t)out_n);
        if (rc) return rc;
    }
    if (as_bf16) {
        std::vector<bf16> hb(out_n);
        for (long i = 0; i < out_n; ++i) hb[i] = __float2bfloat16(src[i]);
        VAE_CUDA(cudaMemcpy(w.b, hb.data(), out_n * sizeof(bf16), cudaMemcpyHostToDevice));
    } else {
        VAE_CUDA(cudaMemcpy(w.f, src, out_n * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

int sd_vae_finalize(sd_vae* v) {
    if (!v) return SD_E_ARG;
    if (v->finalized) return 0;
    for (auto& kv : v->expected)
        if (!v->wt.count(kv.first)) { vae_err("parameter '%s' was never set", kv.first.c_str()); return SD_E_WEIGHT; }
    VAE_CUDA(cudaSetDevice(v->device));
    int rc = build_schedule(v);
    if (rc) return rc;
    v->finalized = true;
    return 0;
}

int sd_vae_decode(sd_vae* v, const float* latents, float* image, unsigned char* rgb8, void* stream) {
    if (!v || !latents || (!image && !rgb8)) { vae_err("sd_vae_decode: bad argument"); return SD_E_ARG; }
    if (!v->finalized) { vae_err("sd_vae_decode before sd_vae_finalize"); return SD_E_STATE; }
    VAE_CUDA(cudaSetDevice(v->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long hw = (long)v->h * v->w, HW = (long)v->H * v->W;
    VAE_CUDA(cudaMemcpyAsync(v->z_in, latents, (size_t)v->NB * 4 * hw * sizeof(float), cudaMemcpyDeviceToDevice, st));
    int i = 0;
    for (auto& op : v->ops) {
        int rc = op(st);
        if (rc) { vae_err("VAE decoder op %d failed: %s", i, rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "descriptor error"); return rc; }
        ++i;
    }
    if (image)      // the first three of every image's four planes
        VAE_CUDA(cudaMemcpy2DAsync(image, 3 * HW * sizeof(float), v->img4, 4 * HW * sizeof(float), 3 * HW * sizeof(float), v->NB, cudaMemcpyDeviceToDevice, st));
    if (rgb8) {
        cudaError_t e = uce::launch_k(to_rgb8_kernel, dim3((unsigned)((v->NB * HW + 255) / 256)), dim3(256), 0, st, 1, (const float*)v->img4, rgb8, v->NB, HW);
        if (e != cudaSuccess) { vae_err("to_rgb8 launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    return 0;
}

int sd_vae_launch_count(sd_vae* v) { return v ? (int)v->ops.size() : SD_E_ARG; }

int sd_vae_inventory(const sd_vae_config* cfg, int index, char* name, size_t cap, long shape[4], int* ndim) {
    if (!cfg || cfg->n_levels < 1 || cfg->n_levels > 4 || cfg->layers_per_block < 1) { vae_err("sd_vae_inventory: bad argument"); return SD_E_ARG; }
    sd_vae tmp;                             // host-only: the inventory depends on the configuration alone
    tmp.cfg = *cfg;
    build_inventory(&tmp);
    const int n = (int)tmp.expected.size();
    if (index < 0) return n;
    if (index >= n || !name || !shape || !ndim) { vae_err("sd_vae_inventory: bad index / output"); return SD_E_ARG; }
    auto it = tmp.expected.begin();
    std::advance(it, index);
    if (it->first.size() + 1 > cap) { vae_err("sd_vae_inventory: name buffer too small"); return SD_E_ARG; }
    memcpy(name, it->first.c_str(), it->first.size() + 1);
    *ndim = (int)it->second.size();
    for (int i = 0; i < *ndim; ++i) shape[i] = it->second[i];
    return n;
}

int sd_vae_read_tap(sd_vae* v, const char* name, float* out, size_t cap, int dims[4]) {
    if (!v || !name || !out) return SD_E_ARG;
    VAE_CUDA(cudaSetDevice(v->device));
    VAE_CUDA(cudaDeviceSynchronize());
    auto it = v->taps.find(name);
    if (it == v->taps.end()) { vae_err("unknown tap '%s' (taps are recorded only when UCE_VAE_TAPS is set before finalize)", name); return SD_E_ARG; }
    const Act a = it->second;
    const size_t n = a.elems();
    if (cap < n) { vae_err("tap buffer too small"); return SD_E_ARG; }
    std::vector<bf16> hb(n);
    VAE_CUDA(cudaMemcpy(hb.data(), a.p, n * sizeof(bf16), cudaMemcpyDeviceToHost));
    for (int nn = 0; nn < a.n; ++nn) for (int y = 0; y < a.h; ++y) for (int x = 0; x < a.w; ++x) for (int c = 0; c < a.c; ++c)
        out[(((size_t)nn * a.c + c) * a.h + y) * a.w + x] = __bfloat162float(hb[(((size_t)nn * a.h + y) * a.w + x) * a.c + c]);
    if (dims) { dims[0] = a.n; dims[1] = a.c; dims[2] = a.h; dims[3] = a.w; }
    return 0;
}

}  // extern "C"
