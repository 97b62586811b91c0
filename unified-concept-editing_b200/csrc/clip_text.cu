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

struct Layer {
    float *ln1_w = nullptr, *ln1_b = nullptr, *qkv_w = nullptr, *qkv_b = nullptr, *out_w = nullptr, *out_b = nullptr;
    float *ln2_w = nullptr, *ln2_b = nullptr, *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;
};

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

// ------------------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) embed_kernel(const int* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                                                    float* __restrict__ x, int T, int D, int vocab) {
    const int row = blockIdx.x, t = row % T;
    int id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4* a = reinterpret_cast<const float4*>(tok + (size_t)id * D);
    const float4* b = reinterpret_cast<const float4*>(pos + (size_t)t * D);
    float4* o = reinterpret_cast<float4*>(x + (size_t)row * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
        const float4 u = a[i], v = b[i];
        o[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row; D % 4 == 0, D <= 4096 (row cached in registers: up to 32 float4 per lane)
constexpr int LN_MAXV = 8;
__global__ void __launch_bounds__(256) ln_kernel(const float* __restrict__ x, float* __restrict__ y, long rows, int D, const float* __restrict__ w,
                                                 const float* __restrict__ b, float eps) {
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = D / 4;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int k = lane + 32 * i;
        v[i] = (k < nv) ? xr[k] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (lane + 32 * i < nv) {
            const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
            ss += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
    const float rstd = rsqrtf(warp_sum(ss) / D + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * D);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int k = lane + 32 * i;
        if (k < nv) {
            const float4 g = reinterpret_cast<const float4*>(w)[k], be = reinterpret_cast<const float4*>(b)[k];
            yr[k] = make_float4((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y,
                                (v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w);
        }
    }
}

// out[M,N] = epi(X[M,K] W[N,K]^T + bias[N]) (+ R[M,N]);  EPI 0 none, 1 quick-GELU (v * sigmoid(1.702 v)), 2 + residual
constexpr int LB = 64, LK = 16;
template <int EPI>
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                                                     const float* __restrict__ R, float* __restrict__ out, int M, int N, int K) {
    __shared__ float As[LK][LB + 4];
    __shared__ float Bs[LK][LB + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * LB, n0 = blockIdx.x * LB;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // each thread stages one float4 of A and one of B per k-tile: row = tid / 4, k quad = tid % 4
    const int lr = tid >> 2, lq = (tid & 3) * 4;
    for (int k0 = 0; k0 < K; k0 += LK) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + lr < M && k0 + lq < K) a = *reinterpret_cast<const float4*>(X + (size_t)(m0 + lr) * K + k0 + lq);
        if (n0 + lr < N && k0 + lq < K) b = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * K + k0 + lq);
        As[lq][lr] = a.x; As[lq + 1][lr] = a.y; As[lq + 2][lr] = a.z; As[lq + 3][lr] = a.w;
        Bs[lq][lr] = b.x; Bs[lq + 1][lr] = b.y; Bs[lq + 2][lr] = b.z; Bs[lq + 3][lr] = b.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < LK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] + bias[gn];
            if (EPI == 1) v = v / (1.f + expf(-1.702f * v));
            if (EPI == 2) v += R[(size_t)gm * N + gn];
            out[(size_t)gm * N + gn] = v;
        }
    }
}

// causal self-attention of one (head, prompt): qkv [B*T, 3D] rows (q | k | v), out [B*T, D].  thread = query row.
// shared: Qs, Ks, Vs [T][dh + 1] each, S [T][T + 1]
__global__ void __launch_bounds__(128) attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int D, int dh, float scale) {
    extern __shared__ float sm[];
    const int head = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    float* Qs = sm; float* Ks = Qs + (size_t)T * (dh + 1); float* Vs = Ks + (size_t)T * (dh + 1); float* S = Vs + (size_t)T * (dh + 1);
    const float* base = qkv + (size_t)b * T * 3 * D + (size_t)head * dh;
    for (int idx = tid; idx < T * dh; idx += blockDim.x) {
        const int t = idx / dh, d = idx % dh;
        Qs[t * (dh + 1) + d] = base[(size_t)t * 3 * D + d] * scale;
        Ks[t * (dh + 1) + d] = base[(size_t)t * 3 * D + D + d];
        Vs[t * (dh + 1) + d] = base[(size_t)t * 3 * D + 2 * D + d];
    }
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) {
        const float* q = Qs + (size_t)t * (dh + 1);
        float* s = S + (size_t)t * (T + 1);
        float m = -INFINITY;
        for (int j = 0; j <= t; ++j) {
            float a = 0.f;
            for (int d = 0; d < dh; ++d) a = fmaf(q[d], Ks[j * (dh + 1) + d], a);
            s[j] = a;
            m = fmaxf(m, a);
        }
        float l = 0.f;
        for (int j = 0; j <= t; ++j) { const float p = expf(s[j] - m); s[j] = p; l += p; }
        const float inv = 1.f / l;
        float* o = out + ((size_t)b * T + t) * D + (size_t)head * dh;
        for (int d = 0; d < dh; ++d) {
            float a = 0.f;
            for (int j = 0; j <= t; ++j) a = fmaf(s[j], Vs[j * (dh + 1) + d], a);
            o[d] = a * inv;
        }
    }
}

__global__ void __launch_bounds__(256) select_rows_kernel(const float* __restrict__ h, const int* __restrict__ idx, float* __restrict__ rows, int T, int D) {
    const int b = blockIdx.x;
    int t = idx[b];
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    const float4* s = reinterpret_cast<const float4*>(h + ((size_t)b * T + t) * D);
    float4* o = reinterpret_cast<float4*>(rows + (size_t)b * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) o[i] = s[i];
}

int linear(clipt_enc* e, int epi, const float* X, const float* W, const float* bias, const float* R, float* out, int M, int N, int K, cudaStream_t st) {
    dim3 grid((N + LB - 1) / LB, (M + LB - 1) / LB);
    if (epi == 0) linear_kernel<0><<<grid, 256, 0, st>>>(X, W, bias, R, out, M, N, K);
    else if (epi == 1) linear_kernel<1><<<grid, 256, 0, st>>>(X, W, bias, R, out, M, N, K);
    else linear_kernel<2><<<grid, 256, 0, st>>>(X, W, bias, R, out, M, N, K);
    CT_CUDA(cudaGetLastError());
    ++e->launches;
    return 0;
}

int forward(clipt_enc* e, const int* ids_host, int batch, int T, cudaStream_t st) {
    if (!e || !ids_host || batch <= 0 || T <= 0) { set_err("clip text: bad argument"); return CLIPT_E_ARG; }
    if (!e->finalized) { set_err("clip text: forward before clipt_finalize"); return CLIPT_E_STATE; }
    if (batch > e->max_batch || T > e->max_pos) { set_err("clip text: batch %d / T %d exceed the encoder's limits (%d / %d)", batch, T, e->max_batch, e->max_pos); return CLIPT_E_ARG; }
    CT_CUDA(cudaSetDevice(e->device));
    const int M = batch * T, D = e->D, dh = D / e->heads;
    e->launches = 0;
    // the pinned staging is reused by every call: wait until the previous call's copy has executed (a forward is milliseconds of work)
    CT_CUDA(cudaStreamSynchronize(st));
    memcpy(e->ids_pin, ids_host, (size_t)M * sizeof(int));
    CT_CUDA(cudaMemcpyAsync(e->ids_dev, e->ids_pin, (size_t)M * sizeof(int), cudaMemcpyHostToDevice, st));
    embed_kernel<<<M, 256, 0, st>>>(e->ids_dev, e->tok, e->pos, e->x, T, D, e->vocab);
    CT_CUDA(cudaGetLastError()); ++e->launches;
    const size_t attn_smem = ((size_t)3 * T * (dh + 1) + (size_t)T * (T + 1)) * sizeof(float);
    static thread_local size_t attn_conf[64] = {0};
    if (attn_smem > 48 * 1024 && attn_conf[e->device & 63] < attn_smem) {
        CT_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
        attn_conf[e->device & 63] = attn_smem;
    }
    int rc;
    for (int l = 0; l < e->layers; ++l) {
        const Layer& L = e->L[l];
        ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(e->x, e->h, M, D, L.ln1_w, L.ln1_b, 1e-5f);
        CT_CUDA(cudaGetLastError()); ++e->launches;
        if ((rc = linear(e, 0, e->h, L.qkv_w, L.qkv_b, nullptr, e->qkv, M, 3 * D, D, st))) return rc;
        attn_kernel<<<dim3(e->heads, batch), 128, attn_smem, st>>>(e->qkv, e->att, T, D, dh, 1.0f / sqrtf((float)dh));
        CT_CUDA(cudaGetLastError()); ++e->launches;
        if ((rc = linear(e, 2, e->att, L.out_w, L.out_b, e->x, e->x, M, D, D, st))) return rc;          // x += out_proj(attn)
        ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(e->x, e->h, M, D, L.ln2_w, L.ln2_b, 1e-5f);
        CT_CUDA(cudaGetLastError()); ++e->launches;
        if ((rc = linear(e, 1, e->h, L.fc1_w, L.fc1_b, nullptr, e->ff, M, e->F, D, st))) return rc;
        if ((rc = linear(e, 2, e->ff, L.fc2_w, L.fc2_b, e->x, e->x, M, D, e->F, st))) return rc;          // x += fc2(quick_gelu(fc1))
    }
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
        Layer& L = e->L[li];
        const std::string r = t.substr(dot + 1);
        if (r == "layer_norm1.weight") { dst = L.ln1_w; want = D; } else if (r == "layer_norm1.bias") { dst = L.ln1_b; want = D; }
        else if (r == "layer_norm2.weight") { dst = L.ln2_w; want = D; } else if (r == "layer_norm2.bias") { dst = L.ln2_b; want = D; }
        else if (r == "self_attn.q_proj.weight") { dst = L.qkv_w; want = D * D; off = 0; } else if (r == "self_attn.k_proj.weight") { dst = L.qkv_w; want = D * D; off = D * D; }
        else if (r == "self_attn.v_proj.weight") { dst = L.qkv_w; want = D * D; off = 2 * D * D; }
        else if (r == "self_attn.q_proj.bias") { dst = L.qkv_b; want = D; off = 0; } else if (r == "self_attn.k_proj.bias") { dst = L.qkv_b; want = D; off = D; }
        else if (r == "self_attn.v_proj.bias") { dst = L.qkv_b; want = D; off = 2 * D; }
        else if (r == "self_attn.out_proj.weight") { dst = L.out_w; want = D * D; } else if (r == "self_attn.out_proj.bias") { dst = L.out_b; want = D; }
        else if (r == "mlp.fc1.weight") { dst = L.fc1_w; want = F * D; } else if (r == "mlp.fc1.bias") { dst = L.fc1_b; want = F; }
        else if (r == "mlp.fc2.weight") { dst = L.fc2_w; want = D * F; } else if (r == "mlp.fc2.bias") { dst = L.fc2_b; want = D; }
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
