// CLIP image tower + zero-shot scoring in fp32 (include/clip_vision_b200.h).
//
// The classifier of the debias loop — transformers.pipeline("zero-shot-image-classification", "openai/clip-vit-base-patch32"),
// trainscripts/uce_sd_debias.py:245-250, called on every generated image at :27 — as CUDA kernels: CLIPImageProcessor (antialiased
// bicubic resize, uint8 rounding, normalisation), the ViT (patch embedding as a GEMM over gathered patches, class token, position
// embeddings, pre-LayerNorm, the transformer layers shared with the text tower in clip_kernels.cuh — full attention here —, post-LayerNorm
// of the class token, visual projection) and the scaled cosine logits against the text tower's rows.  Oracle: oracle/clip_zero_shot_oracle.py,
// pinned to transformers' CLIPModel / CLIPImageProcessor.
#include "../../include/clip_vision_b200.h"
#include "clip_kernels.cuh"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

namespace {
using namespace clipk;

thread_local char g_verr[512] = "";
void set_verr(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_verr, sizeof(g_verr), fmt, ap);
    va_end(ap);
}
#define CV_CUDA(expr)                                                                                               \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess) { set_verr("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); return (int)_e; } \
    } while (0)

}  // namespace

struct clipv_enc {
    int device = 0, S = 0, patch = 0, D = 0, heads = 0, layers = 0, F = 0, E = 0, Dt = 0, max_batch = 0, T = 0, np = 0;
    bool finalized = false;
    float *patch_w = nullptr, *cls = nullptr, *pos = nullptr, *pre_w = nullptr, *pre_b = nullptr, *post_w = nullptr, *post_b = nullptr;
    float *vproj = nullptr, *tproj = nullptr, *zero_bias = nullptr;
    float logit_scale = 0.f;
    std::vector<Layer> L;
    std::map<std::string, int> seen;
    std::vector<void*> allocs;
    float *patches = nullptr, *ptok = nullptr, *x = nullptr, *h = nullptr, *qkv = nullptr, *att = nullptr, *ff = nullptr, *pooled = nullptr, *tfeat = nullptr;
    int* zero_idx = nullptr;
    // resize tables of the last (H -> S) geometry: taps per output coordinate
    int tab_H = 0, taps = 0;
    int* tab_min = nullptr; float* tab_w = nullptr; float* tmp_h = nullptr; size_t tmp_cap = 0;
    float* meanstd = nullptr;
    int launches = 0;
    template <typename T> int alloc(T** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
        if (e != cudaSuccess) { set_verr("cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return (int)e; }
        allocs.push_back(q); *p = (T*)q; return 0;
    }
};

namespace {

// patches[(b * np + p), (c * patch + ky) * patch + kx] = pixel_values[b, c, py * patch + ky, px * patch + kx]  (Conv2d weight order)
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ pv, float* __restrict__ out, int S, int patch, int grid) {
    const int row = blockIdx.x, np = grid * grid, b = row / np, p = row % np, py = p / grid, px = p % grid;
    const int cols = 3 * patch * patch;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
        const int c = i / (patch * patch), r = i % (patch * patch), ky = r / patch, kx = r % patch;
        out[(size_t)row * cols + i] = pv[(((size_t)b * 3 + c) * S + py * patch + ky) * S + px * patch + kx];
    }
}

// x[b, 0, :] = class_embedding + pos[0];  x[b, 1 + p, :] = patch_tokens[b * np + p, :] + pos[1 + p]
__global__ void __launch_bounds__(256) assemble_kernel(const float* __restrict__ ptok, const float* __restrict__ cls, const float* __restrict__ pos,
                                                       float* __restrict__ x, int T, int D) {
    const int row = blockIdx.x, b = row / T, t = row % T;
    const float* src = t == 0 ? cls : ptok + ((size_t)b * (T - 1) + (t - 1)) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) x[(size_t)row * D + i] = src[i] + pos[(size_t)t * D + i];
}

// antialiased bicubic resize, horizontal pass: tmp[b, y, ox, c] = sum_j w[ox][j] * img[b, y, xmin[ox] + j, c]   (uint8 in, fp32 out)
__global__ void __launch_bounds__(256) resize_h_kernel(const unsigned char* __restrict__ img, float* __restrict__ tmp, int H, int S, int taps,
                                                       const int* __restrict__ xmin, const float* __restrict__ w, long total) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % 3); long r = idx / 3;
    const int ox = (int)(r % S); r /= S;
    const int y = (int)(r % H); const int b = (int)(r / H);
    const unsigned char* row = img + (((size_t)b * H + y) * H) * 3 + c;
    const float* ww = w + (size_t)ox * taps;
    const int x0 = xmin[ox];
    float acc = 0.f;
    for (int j = 0; j < taps; ++j) {
        const int xx = x0 + j;
        if (xx < H) acc = fmaf(ww[j], (float)row[(size_t)xx * 3], acc);
    }
    tmp[idx] = acc;
}
// vertical pass + uint8 rounding + rescale + normalisation: pv[b, c, oy, ox] = ((round_clamp(sum_j w[oy][j] tmp[b, ymin + j, ox, c]) / 255) - mean[c]) / std[c]
__global__ void __launch_bounds__(256) resize_v_kernel(const float* __restrict__ tmp, float* __restrict__ pv, int H, int S, int taps,
                                                       const int* __restrict__ ymin, const float* __restrict__ w, const float* __restrict__ meanstd, long total) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox = (int)(idx % S); long r = idx / S;
    const int oy = (int)(r % S); r /= S;
    const int c = (int)(r % 3); const int b = (int)(r / 3);
    const float* ww = w + (size_t)oy * taps;
    const int y0 = ymin[oy];
    float acc = 0.f;
    for (int j = 0; j < taps; ++j) {
        const int yy = y0 + j;
        if (yy < H) acc = fmaf(ww[j], tmp[(((size_t)b * H + yy) * S + ox) * 3 + c], acc);
    }
    acc = fminf(fmaxf(rintf(acc), 0.f), 255.f);
    pv[idx] = (acc / 255.f - meanstd[c]) / meanstd[3 + c];
}
// same geometry (H == S): no resampling, only the rescale + normalisation
__global__ void __launch_bounds__(256) normalize_u8_kernel(const unsigned char* __restrict__ img, float* __restrict__ pv, int S, const float* __restrict__ meanstd, long total) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox = (int)(idx % S); long r = idx / S;
    const int oy = (int)(r % S); r /= S;
    const int c = (int)(r % 3); const int b = (int)(r / 3);
    pv[idx] = ((float)img[(((size_t)b * S + oy) * S + ox) * 3 + c] / 255.f - meanstd[c]) / meanstd[3 + c];
}

// logits[i, j] = scale * <a_i, b_j> / (|a_i| |b_j|); one warp per (i, j)
__global__ void __launch_bounds__(256) cos_logits_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int nA, int nB, int E, float scale) {
    const int w = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= nA * nB) return;
    const int i = w / nB, j = w % nB;
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int k = lane; k < E; k += 32) { const float a = A[(size_t)i * E + k], b = B[(size_t)j * E + k]; ab = fmaf(a, b, ab); aa = fmaf(a, a, aa); bb = fmaf(b, b, bb); }
    ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
    if (lane == 0) out[w] = scale * ab * rsqrtf(aa) * rsqrtf(bb);
}

// torch's antialiased bicubic (a = -0.5) in the precision aten uses for float input (aten UpSampleKernel.cpp, _compute_indices_weights_aa)
float cubic_aa(float x) {
    const float a = -0.5f;
    x = std::fabs(x);
    if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
    if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
    return 0.0f;
}

int build_tables(clipv_enc* e, int H) {
    if (e->tab_H == H) return 0;
    const int S = e->S;
    const float scale = (float)H / (float)S, support = scale >= 1.0f ? 2.0f * scale : 2.0f, inv = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const int taps = (int)std::ceil(support) * 2 + 1;
    std::vector<int> mn(S); std::vector<float> w((size_t)S * taps, 0.f);
    for (int i = 0; i < S; ++i) {
        const float center = scale * ((float)i + 0.5f);
        const long x0 = std::max<long>((long)(center - support + 0.5f), 0);
        const long xs = std::min<long>(std::min<long>((long)(center + support + 0.5f), H) - x0, taps);
        float tot = 0.f;
        for (long j = 0; j < xs; ++j) { const float v = cubic_aa(((float)(j + x0) - center + 0.5f) * inv); w[(size_t)i * taps + j] = v; tot += v; }
        mn[i] = (int)x0;
        for (long j = 0; j < xs; ++j) w[(size_t)i * taps + j] /= tot;
    }
    if (e->tab_min) { cudaFree(e->tab_min); cudaFree(e->tab_w); e->tab_min = nullptr; e->tab_w = nullptr; }
    CV_CUDA(cudaMalloc((void**)&e->tab_min, S * sizeof(int)));
    CV_CUDA(cudaMalloc((void**)&e->tab_w, (size_t)S * taps * sizeof(float)));
    CV_CUDA(cudaMemcpy(e->tab_min, mn.data(), S * sizeof(int), cudaMemcpyHostToDevice));
    CV_CUDA(cudaMemcpy(e->tab_w, w.data(), (size_t)S * taps * sizeof(float), cudaMemcpyHostToDevice));
    e->tab_H = H; e->taps = taps;
    return 0;
}

}  // namespace

extern "C" {

const char* clipv_last_error(void) { return g_verr; }

int clipv_create(int device, int image_size, int patch, int width, int heads, int layers, int ffn, int proj_dim, int text_width, int max_batch, clipv_enc** out) {
    if (!out || image_size <= 0 || patch <= 0 || image_size % patch || width <= 0 || heads <= 0 || width % heads || width % 4 || ffn % 4 || layers <= 0 ||
        proj_dim <= 0 || text_width <= 0 || text_width % 4 || max_batch <= 0 || width > 32 * 4 * LN_MAXV || (patch * patch * 3) % 4) {
        set_verr("clipv_create: bad argument");
        return CLIPV_E_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); set_verr("clipv_create: CUDA device %d not available", device); return CLIPV_E_STATE; }
    CV_CUDA(cudaSetDevice(device));
    clipv_enc* e = new (std::nothrow) clipv_enc();
    if (!e) { set_verr("out of host memory"); return CLIPV_E_STATE; }
    e->device = device; e->S = image_size; e->patch = patch; e->D = width; e->heads = heads; e->layers = layers; e->F = ffn; e->E = proj_dim; e->Dt = text_width;
    e->max_batch = max_batch;
    const int g = image_size / patch;
    e->np = g * g; e->T = e->np + 1;
    e->L.resize(layers);
    const size_t D = width, F = ffn, M = (size_t)max_batch * e->T, PC = (size_t)3 * patch * patch;
    int rc = 0;
#define A_(p, n) if (!rc) rc = e->alloc(&(p), (n))
    A_(e->patch_w, D * PC); A_(e->cls, D); A_(e->pos, (size_t)e->T * D); A_(e->pre_w, D); A_(e->pre_b, D); A_(e->post_w, D); A_(e->post_b, D);
    A_(e->vproj, (size_t)proj_dim * D); A_(e->tproj, (size_t)proj_dim * text_width);
    const size_t zb = std::max<size_t>(std::max<size_t>(D, proj_dim), 16);
    A_(e->zero_bias, zb);
    for (auto& L : e->L) {
        A_(L.ln1_w, D); A_(L.ln1_b, D); A_(L.qkv_w, 3 * D * D); A_(L.qkv_b, 3 * D); A_(L.out_w, D * D); A_(L.out_b, D);
        A_(L.ln2_w, D); A_(L.ln2_b, D); A_(L.fc1_w, F * D); A_(L.fc1_b, F); A_(L.fc2_w, D * F); A_(L.fc2_b, D);
    }
    A_(e->patches, (size_t)max_batch * e->np * PC); A_(e->ptok, (size_t)max_batch * e->np * D);
    A_(e->x, M * D); A_(e->h, M * D); A_(e->qkv, M * 3 * D); A_(e->att, M * D); A_(e->ff, M * F); A_(e->pooled, (size_t)2 * max_batch * D);
    A_(e->tfeat, (size_t)256 * proj_dim); A_(e->zero_idx, (size_t)max_batch); A_(e->meanstd, 8);
#undef A_
    if (!rc) {
        cudaMemset(e->zero_bias, 0, zb * sizeof(float));
        cudaMemset(e->zero_idx, 0, (size_t)max_batch * sizeof(int));
    }
    if (rc) { clipv_destroy(e); return rc; }
    *out = e;
    return 0;
}

int clipv_destroy(clipv_enc* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (void* p : e->allocs) cudaFree(p);
    if (e->tab_min) cudaFree(e->tab_min);
    if (e->tab_w) cudaFree(e->tab_w);
    if (e->tmp_h) cudaFree(e->tmp_h);
    delete e;
    return 0;
}

int clipv_set_weight(clipv_enc* e, const char* name, const float* data, size_t n) {
    if (!e || !name || !data) { set_verr("clipv_set_weight: bad argument"); return CLIPV_E_ARG; }
    CV_CUDA(cudaSetDevice(e->device));
    const std::string s(name);
    const size_t D = e->D, F = e->F;
    float* dst = nullptr; size_t want = 0, off = 0;
    if (s == "logit_scale") {
        if (n != 1) { set_verr("clipv_set_weight: logit_scale has one element"); return CLIPV_E_ARG; }
        e->logit_scale = data[0]; e->seen[s] = 1;
        return 0;
    }
    if (s == "visual_projection.weight") { dst = e->vproj; want = (size_t)e->E * D; }
    else if (s == "text_projection.weight") { dst = e->tproj; want = (size_t)e->E * e->Dt; }
    else if (s.compare(0, 13, "vision_model.") == 0) {
        const std::string t = s.substr(13);
        if (t == "embeddings.patch_embedding.weight") { dst = e->patch_w; want = D * 3 * e->patch * e->patch; }
        else if (t == "embeddings.class_embedding") { dst = e->cls; want = D; }
        else if (t == "embeddings.position_embedding.weight") { dst = e->pos; want = (size_t)e->T * D; }
        else if (t == "pre_layrnorm.weight") { dst = e->pre_w; want = D; } else if (t == "pre_layrnorm.bias") { dst = e->pre_b; want = D; }
        else if (t == "post_layernorm.weight") { dst = e->post_w; want = D; } else if (t == "post_layernorm.bias") { dst = e->post_b; want = D; }
        else if (t.compare(0, 15, "encoder.layers.") == 0) {
            const size_t dot = t.find('.', 15);
            const int li = atoi(t.substr(15, dot - 15).c_str());
            if (dot == std::string::npos || li < 0 || li >= e->layers) { set_verr("clipv_set_weight: layer index out of range in '%s'", name); return CLIPV_E_ARG; }
            dst = layer_param(e->L[li], t.substr(dot + 1), D, F, &want, &off);
        }
    }
    if (!dst) { set_verr("clipv_set_weight: unknown parameter '%s'", name); return CLIPV_E_ARG; }
    if (n != want) { set_verr("clipv_set_weight: '%s' has %zu elements, expected %zu", name, n, want); return CLIPV_E_ARG; }
    CV_CUDA(cudaMemcpy(dst + off, data, n * sizeof(float), cudaMemcpyHostToDevice));
    e->seen[s] = 1;
    return 0;
}

int clipv_finalize(clipv_enc* e) {
    if (!e) return CLIPV_E_ARG;
    const size_t want = 7 + 3 + (size_t)e->layers * 16;
    if (e->seen.size() != want) { set_verr("clipv_finalize: %zu of %zu parameters uploaded", e->seen.size(), want); return CLIPV_E_STATE; }
    e->finalized = true;
    return 0;
}

int clipv_preprocess_u8(clipv_enc* e, const unsigned char* images, int batch, int H, const float* mean, const float* std_, float* pixel_values, void* stream) {
    if (!e || !images || !mean || !std_ || !pixel_values || batch <= 0 || H <= 0) { set_verr("clipv_preprocess_u8: bad argument"); return CLIPV_E_ARG; }
    CV_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    const float ms[6] = {mean[0], mean[1], mean[2], std_[0], std_[1], std_[2]};
    CV_CUDA(cudaMemcpyAsync(e->meanstd, ms, sizeof(ms), cudaMemcpyHostToDevice, st));
    CV_CUDA(cudaStreamSynchronize(st));                    // `ms` is a stack buffer
    const int S = e->S;
    e->launches = 0;
    if (H == S) {
        const long total = (long)batch * 3 * S * S;
        normalize_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(images, pixel_values, S, e->meanstd, total);
        CV_CUDA(cudaGetLastError()); ++e->launches;
        return 0;
    }
    int rc = build_tables(e, H);
    if (rc) return rc;
    const size_t need = (size_t)batch * H * S * 3;
    if (need > e->tmp_cap) {
        if (e->tmp_h) { CV_CUDA(cudaStreamSynchronize(st)); cudaFree(e->tmp_h); e->tmp_h = nullptr; }
        CV_CUDA(cudaMalloc((void**)&e->tmp_h, need * sizeof(float)));
        e->tmp_cap = need;
    }
    const long t1 = (long)batch * H * S * 3, t2 = (long)batch * 3 * S * S;
    resize_h_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(images, e->tmp_h, H, S, e->taps, e->tab_min, e->tab_w, t1);
    CV_CUDA(cudaGetLastError()); ++e->launches;
    resize_v_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(e->tmp_h, pixel_values, H, S, e->taps, e->tab_min, e->tab_w, e->meanstd, t2);
    CV_CUDA(cudaGetLastError()); ++e->launches;
    return 0;
}

int clipv_image_features(clipv_enc* e, const float* pixel_values, int batch, float* features, void* stream) {
    if (!e || !pixel_values || !features || batch <= 0) { set_verr("clipv_image_features: bad argument"); return CLIPV_E_ARG; }
    if (!e->finalized) { set_verr("clipv_image_features before clipv_finalize"); return CLIPV_E_STATE; }
    if (batch > e->max_batch) { set_verr("clipv_image_features: batch %d exceeds max_batch %d", batch, e->max_batch); return CLIPV_E_ARG; }
    CV_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int D = e->D, T = e->T, np = e->np, PC = 3 * e->patch * e->patch, M = batch * T;
    e->launches = 0;
    im2col_kernel<<<batch * np, 256, 0, st>>>(pixel_values, e->patches, e->S, e->patch, e->S / e->patch);
    CV_CUDA(cudaGetLastError()); ++e->launches;
    CV_CUDA(launch_linear(0, e->patches, e->patch_w, e->zero_bias, nullptr, e->ptok, batch * np, D, PC, st)); ++e->launches;
    assemble_kernel<<<M, 256, 0, st>>>(e->ptok, e->cls, e->pos, e->x, T, D);
    CV_CUDA(cudaGetLastError()); ++e->launches;
    ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(e->x, e->x, M, D, e->pre_w, e->pre_b, 1e-5f);          // in place: a warp owns its row
    CV_CUDA(cudaGetLastError()); ++e->launches;
    CV_CUDA(run_layers(e->L.data(), e->layers, e->x, e->h, e->qkv, e->att, e->ff, batch, T, D, e->heads, e->F, /*causal=*/0, e->device, st, &e->launches));
    select_rows_kernel<<<batch, 256, 0, st>>>(e->x, e->zero_idx, e->pooled, T, D);                 // the class token of every image
    CV_CUDA(cudaGetLastError()); ++e->launches;
    float* pl = e->pooled + (size_t)e->max_batch * D;
    ln_kernel<<<(batch + 7) / 8, 256, 0, st>>>(e->pooled, pl, batch, D, e->post_w, e->post_b, 1e-5f);
    CV_CUDA(cudaGetLastError()); ++e->launches;
    CV_CUDA(launch_linear(0, pl, e->vproj, e->zero_bias, nullptr, features, batch, e->E, D, st)); ++e->launches;
    return 0;
}

int clipv_logits(clipv_enc* e, const float* image_features, int n_images, const float* text_rows, int n_texts, float* logits, void* stream) {
    if (!e || !image_features || !text_rows || !logits || n_images <= 0 || n_texts <= 0 || n_texts > 256) { set_verr("clipv_logits: bad argument (at most 256 texts)"); return CLIPV_E_ARG; }
    if (!e->finalized) { set_verr("clipv_logits before clipv_finalize"); return CLIPV_E_STATE; }
    CV_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    CV_CUDA(launch_linear(0, text_rows, e->tproj, e->zero_bias, nullptr, e->tfeat, n_texts, e->E, e->Dt, st));
    const long threads = (long)n_images * n_texts * 32;
    cos_logits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(image_features, e->tfeat, logits, n_images, n_texts, e->E, expf(e->logit_scale));
    CV_CUDA(cudaGetLastError());
    return 0;
}

int clipv_launch_count(clipv_enc* e) { return e ? e->launches : CLIPV_E_ARG; }

}  // extern "C"
