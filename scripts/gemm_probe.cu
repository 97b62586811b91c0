// Stand-alone timing probe of the U-Net GEMM kernel on representative SD-1.4 shapes (not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DUG_TRACE -lcuda scripts/gemm_probe.cu -o gpurun_out/gemm_probe
// Prints, per shape: L2-warm and L2-flushed time per launch (CUDA events), and clock64 stamps of CTA (0,0,0).
#include "../unified-concept-editing_b200/csrc/unet_gemm.cu"
#include <vector>
#include <cstdlib>
using namespace uce;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Shape { const char* name; int conv; int NB, H, W, Cin, Cout, stride; int M, N, K; int bias, residual; };

int main(int argc, char** argv) {
    int sm = 0; CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    const Shape shapes[] = {
        {"lin 128x1280x1280", 0, 0, 0, 0, 0, 0, 0, 128, 1280, 1280, 1, 0},
        {"lin 512x1280x1280", 0, 0, 0, 0, 0, 0, 0, 512, 1280, 1280, 1, 1},
        {"lin 2048x640x640", 0, 0, 0, 0, 0, 0, 0, 2048, 640, 640, 1, 1},
        {"lin 8192x320x320", 0, 0, 0, 0, 0, 0, 0, 8192, 320, 320, 1, 1},
        {"lin 8192x2560x320 (ff1)", 0, 0, 0, 0, 0, 0, 0, 8192, 2560, 320, 1, 0},
        {"lin 8192x320x1280 (ff2)", 0, 0, 0, 0, 0, 0, 0, 8192, 320, 1280, 1, 1},
        {"conv 64x64 320->320", 1, 2, 64, 64, 320, 320, 1, 0, 0, 0, 1, 1},
        {"conv 32x32 640->640", 1, 2, 32, 32, 640, 640, 1, 0, 0, 0, 1, 1},
        {"conv 16x16 1280->1280", 1, 2, 16, 16, 1280, 1280, 1, 0, 0, 0, 1, 1},
        {"conv 8x8 1280->1280", 1, 2, 8, 8, 1280, 1280, 1, 0, 0, 0, 1, 1},
        {"conv 32x32 1280->640", 1, 2, 32, 32, 1280, 640, 1, 0, 0, 0, 1, 1},
        {"conv 64x64 640->320", 1, 2, 64, 64, 640, 320, 1, 0, 0, 0, 1, 1},
    };
    const size_t flush_bytes = 512u << 20;
    void* flush; CK(cudaMalloc(&flush, flush_bytes));
    float* ws; const size_t ws_cap = (size_t)3 * sm * 128 * 128; CK(cudaMalloc(&ws, ws_cap * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (const Shape& s : shapes) {
        const int M = s.conv ? s.NB * (s.H / s.stride) * (s.W / s.stride) : s.M;
        const int N = s.conv ? s.Cout : s.N;
        const long K = s.conv ? 9L * s.Cin : s.K;
        __nv_bfloat16 *A, *B, *out, *res; float* bias;
        const size_t a_el = s.conv ? (size_t)s.NB * s.H * s.W * s.Cin : (size_t)M * K;
        CK(cudaMalloc(&A, a_el * 2)); CK(cudaMalloc(&B, (size_t)N * K * 2)); CK(cudaMalloc(&out, (size_t)M * N * 2)); CK(cudaMalloc(&res, (size_t)M * N * 2));
        CK(cudaMalloc(&bias, N * 4));
        CK(cudaMemset(A, 0, a_el * 2)); CK(cudaMemset(B, 0, (size_t)N * K * 2)); CK(cudaMemset(res, 0, (size_t)M * N * 2)); CK(cudaMemset(bias, 0, N * 4));
        GemmDesc g;
        int rc = s.conv ? gemm_desc_conv(&g, A, s.NB, s.H, s.W, s.Cin, B, s.Cout, 3, s.stride)
                        : gemm_desc_linear(&g, A, K, 0, 0, B, K, 0, 0, M, N, (int)K, 1, 1, 0, 0);
        if (rc) { printf("%s: descriptor failed %d\n", s.name, rc); continue; }
        g.out = out; g.out_fp32 = 0; g.ldo = N; g.bias = s.bias ? bias : nullptr; g.residual = s.residual ? res : nullptr; g.ldr = N;
        if (!getenv("UCE_NO_PAIR")) gemm_enable_pair(&g);
        int ks = gemm_choose_ksplit(g, sm);
        while (ks > 1 && (size_t)ks * g.M * g.N > ws_cap) --ks;
        if (argc > 1) ks = atoi(argv[1]) > 0 ? atoi(argv[1]) : ks;
        if (ks > 1) { g.ksplit = ks; g.splitk_ws = ws; }
        if (!getenv("UCE_NO_TMA_EPI")) gemm_enable_tma_epilogue(&g);
        g.stages = gemm_choose_stages(g, sm, &g.katoms);
        const int ctas = (g.pair ? ((N + g.bn - 1) / g.bn) * ((g.m_tiles + 1) / 2 * 2) : ((N + 127) / 128) * g.m_tiles) * (ks > 1 ? ks : 1);
        for (int i = 0; i < 3; ++i) { int lr = gemm_launch(g, 0); if (lr) { printf("%s: launch failed %d\n", s.name, lr); break; } }
        CK(cudaDeviceSynchronize());
        const int R = 20;
        CK(cudaEventRecord(e0)); for (int i = 0; i < R; ++i) gemm_launch(g, 0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float warm; CK(cudaEventElapsedTime(&warm, e0, e1)); warm = warm * 1000 / R;
        float cold = 0;
        for (int i = 0; i < 5; ++i) {
            CK(cudaMemsetAsync(flush, i, flush_bytes));
            CK(cudaEventRecord(e0)); gemm_launch(g, 0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float t; CK(cudaEventElapsedTime(&t, e0, e1)); cold += t * 1000 / 5;
        }
        long long tr[16]; CK(cudaMemcpyFromSymbol(tr, ug_trace, sizeof(tr)));
        const double gf = 2.0 * M * N * (double)K * 1e-9;
        printf("%-28s M=%5d N=%4d K=%5ld bn=%3d ctas=%4d ks=%2d st=%d | warm %7.1f us (%6.0f TF/s) cold %7.1f us | cta0 cycles: setup %lld, 1st data %lld, mma done %lld, acc seen %lld, epi done %lld, exit %lld\n",
               s.name, M, N, K, g.pair ? g.bn : 128, ctas, ks, g.stages, warm, gf / warm * 1e-3, cold,
               tr[1] - tr[0], tr[3] - tr[0], tr[4] - tr[0], tr[5] - tr[0], tr[6] - tr[0], tr[7] - tr[0]);
        cudaFree(A); cudaFree(B); cudaFree(out); cudaFree(res); cudaFree(bias);
    }
    return 0;
}
