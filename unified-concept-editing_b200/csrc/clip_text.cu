// Batched CLIP text encoder in fp32 (include/clip_text_b200.h).
//
// What the reference runs once per concept — pipe.encode_prompt -> transformers CLIPTextModel.forward, trainscripts/uce_sd_erase.py:29-33 —
// as ONE forward over all distinct prompts, plus the row select of :34-42.  fp32 SIMT kernels: the edit is a function of these rows and the
// reference computes them in fp32 (uce_sd_erase.py:117), so this path is held to fp32 parity (oracle/clip_text_oracle.py, pinned to the
// transformers implementation), not to the bf16 storage of the U-Net engine.
//
//   embed        x[b,t,:] = token_embedding[ids[b,t]] + position_embedding[t]
//   layernorm    one warp per row, row in registers, two-pass variance
//   linear       out = epilogue(X W^T + bias): 64 x 64 x 16 tiles, 4 x 4 register blocking; epilogues: none | quick-GELU | + residual
//   attention    one CTA per (head, prompt): K and V of the head in shared memory, thread = query row, causal mask, q scaled by dh^-0.5
//   select       rows_out[b,:] = hidden[b, row_index[b], :]
#include "../../include/clip_text_b200.h"
#include "clip_kernels.cuh"
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local char g_err[512] = "";
void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
#define CT_CUDA(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess) { set_err("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); return (int)_e; } \
    } while (0)

using clipk::Layer;

}  // namespace

struct clipt_enc {
    int device = 0, vocab = 0, D = 0, heads = 0, layers = 0, F = 0, max_pos = 0, max_batch = 0;
    bool finalized = false;
    float *tok = nullptr, *pos = nullptr, *fln_w = nullptr, *fln_b = nullptr;
    std::vector<Layer> L;
    std::map<std::string, int> seen;
    std::vector<void*> allocs;
    int* ids_dev = nullptr; int* idx_dev = nullptr;
    int *ids_pin = nullptr, *idx_pin = nullptr;
    float *x = nullptr, *h = nullptr, *qkv = nullptr, *att = nullptr, *ff = nullptr;
    int launches = 0;
    template <typename T> int alloc(T** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
        if (e != cudaSuccess) { set_err("cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return (int)e; }
        allocs.push_back(q); *p = (T*)q; return 0;
    }
};

namespace {
using namespace clipk;

int forward(clipt_enc* e, const int* ids_host, int batch, int T, cudaStream_t st) {
    if (!e || !ids_host || batch <= 0 || T <= 0) { set_err("clip text: bad argument"); return CLIPT_E_ARG; }
    if (!e->finalized) { set_err("clip text: forward before clipt_finalize"); return CLIPT_E_STATE; }
    if (batch > e->max_batch || T > e->max_pos) { set_err("clip text: batch %d / T %d exceed the encoder's limits (%d / %d)", batch, T, e->max_batch, e->max_pos); return CLIPT_E_ARG; }
    CT_CUDA(cudaSetDevice(e->device));
    const int M = batch * T, D = e->D;
    e->launches = 0;
    // the pinned staging is reused by every call: wait until the previous call's copy has executed (a forward is milliseconds of work)
    CT_CUDA(cudaStreamSynchronize(st));
    memcpy(e->ids_pin, ids_host, (size_t)M * sizeof(int));
    CT_CUDA(cudaMemcpyAsync(e->ids_dev, e->ids_pin, (size_t)M * sizeof(int), cudaMemcpyHostToDevice, st));
    embed_kernel<<<M, 256, 0, st>>>(e->ids_dev, e->tok, e->pos, e->x, T, D, e->vocab);
    CT_CUDA(cudaGetLastError()); ++e->launches;
    CT_CUDA(run_layers(e->L.data(), e->layers, e->x, e->h, e->qkv, e->att, e->ff, batch, T, D, e->heads, e->F, /*causal=*/1, e->device, st, &e->launches));
    ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(e->x, e->h, M, D, e->fln_w, e->fln_b, 1e-5f);
    CT_CUDA(cudaGetLastError()); ++e->launches;
    return 0;
}

}  // namespace

extern "C" {

const char* clipt_last_error(void) { return g_err; }

int clipt_create(int device, int vocab, int width, int heads, int layers, int ffn, int max_pos, int max_batch, clipt_enc** out) {
    if (!out || vocab <= 0 || width <= 0 || heads <= 0 || layers <= 0 || ffn <= 0 || max_pos <= 0 || max_batch <= 0 || width % heads || width % 4 || ffn % 4 ||
        width > 32 * 4 * LN_MAXV) {
        set_err("clipt_create: bad argument (width must be a multiple of 4 and of heads, <= %d)", 32 * 4 * LN_MAXV);
        return CLIPT_E_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); set_err("clipt_create: CUDA device %d not available", device); return CLIPT_E_STATE; }
    CT_CUDA(cudaSetDevice(device));
    clipt_enc* e = new (std::nothrow) clipt_enc();
    if (!e) { set_err("out of host memory"); return CLIPT_E_STATE; }
    e->device = device; e->vocab = vocab; e->D = width; e->heads = heads; e->layers = layers; e->F = ffn; e->max_pos = max_pos; e->max_batch = max_batch;
    e->L.resize(layers);
    const size_t D = width, F = ffn, M = (size_t)max_batch * max_pos;
    int rc = 0;
#define A_(p, n) if (!rc) rc = e->alloc(&(p), (n))
    A_(e->tok, (size_t)vocab * D); A_(e->pos, (size_t)max_pos * D); A_(e->fln_w, D); A_(e->fln_b, D);
    for (auto& L : e->L) {
        A_(L.ln1_w, D); A_(L.ln1_b, D); A_(L.qkv_w, 3 * D * D); A_(L.qkv_b, 3 * D); A_(L.out_w, D * D); A_(L.out_b, D);
        A_(L.ln2_w, D); A_(L.ln2_b, D); A_(L.fc1_w, F * D); A_(L.fc1_b, F); A_(L.fc2_w, D * F); A_(L.fc2_b, D);
    }
    A_(e->ids_dev, M); A_(e->idx_dev, (size_t)max_batch);
    A_(e->x, M * D); A_(e->h, M * D); A_(e->qkv, M * 3 * D); A_(e->att, M * D); A_(e->ff, M * F);
#undef A_
    if (!rc && cudaMallocHost((void**)&e->ids_pin, M * sizeof(int)) != cudaSuccess) { set_err("cudaMallocHost failed"); rc = CLIPT_E_STATE; }
    if (!rc && cudaMallocHost((void**)&e->idx_pin, (size_t)max_batch * sizeof(int)) != cudaSuccess) { set_err("cudaMallocHost failed"); rc = CLIPT_E_STATE; }
    if (rc) { clipt_destroy(e); return rc; }
    *out = e;
    return 0;
}

int clipt_destroy(clipt_enc* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (void* p : e->allocs) cudaFree(p);
    if (e->ids_pin) cudaFreeHost(e->ids_pin);
    if (e->idx_pin) cudaFreeHost(e->idx_pin);
    delete e;
    return 0;
}

int clipt_set_weight(clipt_enc* e, const char* name, const float* data, size_t n) {
    if (!e || !name || !data) { set_err("clipt_set_weight: bad argument"); return CLIPT_E_ARG; }
    CT_CUDA(cudaSetDevice(e->device));
    const std::string s(name);
    const size_t D = e->D, F = e->F;
    float* dst = nullptr; size_t want = 0, off = 0;
    const std::string pre = "text_model.";
    if (s.compare(0, pre.size(), pre) != 0) { set_err("clipt_set_weight: unknown parameter '%s'", name); return CLIPT_E_ARG; }
    const std::string t = s.substr(pre.size());
    if (t == "embeddings.token_embedding.weight") { dst = e->tok; want = (size_t)e->vocab * D; }
    else if (t == "embeddings.position_embedding.weight") { dst = e->pos; want = (size_t)e->max_pos * D; }
    else if (t == "final_layer_norm.weight") { dst = e->fln_w; want = D; }
    else if (t == "final_layer_norm.bias") { dst = e->fln_b; want = D; }
    else if (t.compare(0, 15, "encoder.layers.") == 0) {
        const size_t dot = t.find('.', 15);
        const int li = atoi(t.substr(15, dot - 15).c_str());
        if (dot == std::string::npos || li < 0 || li >= e->layers) { set_err("clipt_set_weight: layer index out of range in '%s'", name); return CLIPT_E_ARG; }
        dst = layer_param(e->L[li], t.substr(dot + 1), D, F, &want, &off);
    }
    if (!dst) { set_err("clipt_set_weight: unknown parameter '%s'", name); return CLIPT_E_ARG; }
    if (n != want) { set_err("clipt_set_weight: '%s' has %zu elements, expected %zu", name, n, want); return CLIPT_E_ARG; }
    CT_CUDA(cudaMemcpy(dst + off, data, n * sizeof(float), cudaMemcpyHostToDevice));
    e->seen[s] = 1;
    return 0;
}

int clipt_finalize(clipt_enc* e) {
    if (!e) return CLIPT_E_ARG;
    const size_t want = 4 + (size_t)e->layers * 16;
    if (e->seen.size() != want) { set_err("clipt_finalize: %zu of %zu parameters uploaded", e->seen.size(), want); return CLIPT_E_STATE; }
    e->finalized = true;
    return 0;
}

int clipt_encode(clipt_enc* e, const int* input_ids, int batch, int T, float* hidden_out, void* stream) {
    if (!hidden_out) { set_err("clipt_encode: bad argument"); return CLIPT_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = forward(e, input_ids, batch, T, st);
    if (rc) return rc;
    CT_CUDA(cudaMemcpyAsync(hidden_out, e->h, (size_t)batch * T * e->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int clipt_concept_rows(clipt_enc* e, const int* input_ids, const int* row_index, int batch, int T, float* rows_out, void* stream) {
    if (!rows_out || !row_index) { set_err("clipt_concept_rows: bad argument"); return CLIPT_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = forward(e, input_ids, batch, T, st);
    if (rc) return rc;
    memcpy(e->idx_pin, row_index, (size_t)batch * sizeof(int));
    CT_CUDA(cudaMemcpyAsync(e->idx_dev, e->idx_pin, (size_t)batch * sizeof(int), cudaMemcpyHostToDevice, st));
    select_rows_kernel<<<batch, 256, 0, st>>>(e->h, e->idx_dev, rows_out, T, e->D);
    CT_CUDA(cudaGetLastError()); ++e->launches;
    return 0;
}

int clipt_launch_count(clipt_enc* e) { return e ? e->launches : CLIPT_E_ARG; }

}  // extern "C"
