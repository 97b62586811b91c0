// Kernels shared by the CLIP text tower (clip_text.cu) and the CLIP vision tower (clip_vision.cu): fp32 SIMT, see the two files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

namespace clipk {

struct Layer {
    float *ln1_w = nullptr, *ln1_b = nullptr, *qkv_w = nullptr, *qkv_b = nullptr, *out_w = nullptr, *out_b = nullptr;
    float *ln2_w = nullptr, *ln2_b = nullptr, *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;
};

static __global__ void __launch_bounds__(256) embed_kernel(const int* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
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
static __global__ void __launch_bounds__(256) ln_kernel(const float* __restrict__ x, float* __restrict__ y, long rows, int D, const float* __restrict__ w,
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
static __global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
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

// self-attention of one (head, sequence), causal (text tower) or full (vision tower): qkv [B*T, 3D] rows (q | k | v), out [B*T, D].  thread = query row.
// shared: Qs, Ks, Vs [T][dh + 1] each, S [T][T + 1]
static __global__ void __launch_bounds__(128) attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int D, int dh, float scale, int causal) {
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
        const int jl = causal ? t : T - 1;                  // last key this query sees
        float m = -INFINITY;
        for (int j = 0; j <= jl; ++j) {
            float a = 0.f;
            for (int d = 0; d < dh; ++d) a = fmaf(q[d], Ks[j * (dh + 1) + d], a);
            s[j] = a;
            m = fmaxf(m, a);
        }
        float l = 0.f;
        for (int j = 0; j <= jl; ++j) { const float p = expf(s[j] - m); s[j] = p; l += p; }
        const float inv = 1.f / l;
        float* o = out + ((size_t)b * T + t) * D + (size_t)head * dh;
        for (int d = 0; d < dh; ++d) {
            float a = 0.f;
            for (int j = 0; j <= jl; ++j) a = fmaf(s[j], Vs[j * (dh + 1) + d], a);
            o[d] = a * inv;
        }
    }
}

static __global__ void __launch_bounds__(256) select_rows_kernel(const float* __restrict__ h, const int* __restrict__ idx, float* __restrict__ rows, int T, int D) {
    const int b = blockIdx.x;
    int t = idx[b];
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    const float4* s = reinterpret_cast<const float4*>(h + ((size_t)b * T + t) * D);
    float4* o = reinterpret_cast<float4*>(rows + (size_t)b * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) o[i] = s[i];
}


inline cudaError_t launch_linear(int epi, const float* X, const float* W, const float* bias, const float* R, float* out, int M, int N, int K, cudaStream_t st) {
    dim3 grid((N + LB - 1) / LB, (M + LB - 1) / LB);
    if (epi == 0) linear_kernel<0><<<grid, 256, 0, st>>>(X, W, bias, R, out, M, N, K);
    else if (epi == 1) linear_kernel<1><<<grid, 256, 0, st>>>(X, W, bias, R, out, M, N, K);
    else linear_kernel<2><<<grid, 256, 0, st>>>(X, W, bias, R, out, M, N, K);
    return cudaGetLastError();
}

inline size_t attn_smem_bytes(int T, int dh) { return ((size_t)3 * T * (dh + 1) + (size_t)T * (T + 1)) * sizeof(float); }

// L pre-LayerNorm transformer layers on x [batch * T, D] in place (h, qkv, att, ff: scratch).  Returns the first CUDA error; counts launches.
inline cudaError_t run_layers(const Layer* layers, int n_layers, float* x, float* h, float* qkv, float* att, float* ff, int batch, int T, int D,
                              int heads, int F, int causal, int device, cudaStream_t st, int* launches) {
    const int M = batch * T, dh = D / heads;
    const size_t smem = attn_smem_bytes(T, dh);
    static thread_local size_t conf[64] = {0};
    if (smem > 48 * 1024 && conf[device & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        conf[device & 63] = smem;
    }
    cudaError_t e;
    for (int l = 0; l < n_layers; ++l) {
        const Layer& L = layers[l];
        ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, h, M, D, L.ln1_w, L.ln1_b, 1e-5f);
        if ((e = cudaGetLastError()) != cudaSuccess) return e; ++*launches;
        if ((e = launch_linear(0, h, L.qkv_w, L.qkv_b, nullptr, qkv, M, 3 * D, D, st)) != cudaSuccess) return e; ++*launches;
        attn_kernel<<<dim3(heads, batch), 128, smem, st>>>(qkv, att, T, D, dh, 1.0f / sqrtf((float)dh), causal);
        if ((e = cudaGetLastError()) != cudaSuccess) return e; ++*launches;
        if ((e = launch_linear(2, att, L.out_w, L.out_b, x, x, M, D, D, st)) != cudaSuccess) return e; ++*launches;      // x += out_proj(attn)
        ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, h, M, D, L.ln2_w, L.ln2_b, 1e-5f);
        if ((e = cudaGetLastError()) != cudaSuccess) return e; ++*launches;
        if ((e = launch_linear(1, h, L.fc1_w, L.fc1_b, nullptr, ff, M, F, D, st)) != cudaSuccess) return e; ++*launches;
        if ((e = launch_linear(2, ff, L.fc2_w, L.fc2_b, x, x, M, D, F, st)) != cudaSuccess) return e; ++*launches;     // x += fc2(quick_gelu(fc1))
    }
    return cudaSuccess;
}

// maps "<i>.<rest>" of an encoder.layers.* parameter name to its slot; returns nullptr for unknown names.  want / off in floats.
inline float* layer_param(Layer& L, const std::string& r, size_t D, size_t F, size_t* want, size_t* off) {
    *off = 0;
    if (r == "layer_norm1.weight") { *want = D; return L.ln1_w; } if (r == "layer_norm1.bias") { *want = D; return L.ln1_b; }
    if (r == "layer_norm2.weight") { *want = D; return L.ln2_w; } if (r == "layer_norm2.bias") { *want = D; return L.ln2_b; }
    if (r == "self_attn.q_proj.weight") { *want = D * D; return L.qkv_w; } if (r == "self_attn.k_proj.weight") { *want = D * D; *off = D * D; return L.qkv_w; }
    if (r == "self_attn.v_proj.weight") { *want = D * D; *off = 2 * D * D; return L.qkv_w; }
    if (r == "self_attn.q_proj.bias") { *want = D; return L.qkv_b; } if (r == "self_attn.k_proj.bias") { *want = D; *off = D; return L.qkv_b; }
    if (r == "self_attn.v_proj.bias") { *want = D; *off = 2 * D; return L.qkv_b; }
    if (r == "self_attn.out_proj.weight") { *want = D * D; return L.out_w; } if (r == "self_attn.out_proj.bias") { *want = D; return L.out_b; }
    if (r == "mlp.fc1.weight") { *want = F * D; return L.fc1_w; } if (r == "mlp.fc1.bias") { *want = F; return L.fc1_b; }
    if (r == "mlp.fc2.weight") { *want = D * F; return L.fc2_w; } if (r == "mlp.fc2.bias") { *want = D; return L.fc2_b; }
    return nullptr;
}

}  // namespace clipk
