// Numerical check of the U-Net GEMM kernels (single-CTA and CTA-pair, register and TMA epilogues, split-K) against a
// straightforward CPU evaluation, on small shapes.  Not part of the library:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fopenmp -lcuda scripts/gemm_check.cu -o scripts/gemm_check.bin
#include "../unified-concept-editing_b200/csrc/unet_gemm.cu"
#include <vector>
#include <cstdlib>
#include <cmath>
using namespace uce;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static uint32_t rng = 12345u;
static float frand() { rng = rng * 1664525u + 1013904223u; return ((rng >> 8) & 0xFFFF) / 65536.f - 0.5f; }
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Case { const char* name; int conv, NB, H, W, Cin, Cout; int M, N, K; int bias, rowbias, residual, inplace, force_ks, pair, tma; };

int main() {
    int sm = 0; CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    const Case cases[] = {
        {"lin 512x256x192 reg", 0, 0, 0, 0, 0, 0, 512, 256, 192, 1, 0, 1, 0, 0, 0, 0},
        {"lin 512x256x192 pair reg-epi", 0, 0, 0, 0, 0, 0, 512, 256, 192, 1, 0, 1, 0, 0, 1, 0},
        {"lin 512x256x192 pair tma bias", 0, 0, 0, 0, 0, 0, 512, 256, 192, 1, 0, 0, 0, 0, 1, 1},
        {"lin 512x256x192 pair tma res", 0, 0, 0, 0, 0, 0, 512, 256, 192, 0, 0, 1, 0, 0, 1, 1},
        {"lin 512x256x192 pair tma res inplace", 0, 0, 0, 0, 0, 0, 512, 256, 192, 1, 0, 1, 1, 0, 1, 1},
        {"lin 300x320x320 pair tma (ragged M, N=320)", 0, 0, 0, 0, 0, 0, 300, 320, 320, 1, 0, 1, 0, 0, 1, 1},
        {"lin 512x640x2560 pair tma split3", 0, 0, 0, 0, 0, 0, 512, 640, 2560, 1, 0, 1, 0, 3, 1, 1},
        {"conv 16x16 64->64 pair tma all", 1, 2, 16, 16, 64, 64, 0, 0, 0, 1, 1, 1, 0, 0, 1, 1},
        {"conv 16x16 64->64 reg all", 1, 2, 16, 16, 64, 64, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0},
        {"conv 32x32 128->320 pair tma all", 1, 2, 32, 32, 128, 320, 0, 0, 0, 1, 1, 1, 0, 0, 1, 1},
        {"conv 16x16 256->128 pair tma split4", 1, 2, 16, 16, 256, 128, 0, 0, 0, 1, 1, 1, 0, 4, 1, 1},
        {"conv 8x8 128->128 pair tma (2 img/tile)", 1, 4, 8, 8, 128, 128, 0, 0, 0, 1, 1, 1, 0, 0, 1, 1},
    };
    float* ws; const size_t ws_cap = (size_t)3 * sm * 128 * 128; CK(cudaMalloc(&ws, ws_cap * 4));
    int bad = 0;
    for (const Case& c : cases) {
        const int M = c.conv ? c.NB * c.H * c.W : c.M, N = c.conv ? c.Cout : c.N;
        const long K = c.conv ? 9L * c.Cin : c.K;
        const size_t a_el = c.conv ? (size_t)c.NB * c.H * c.W * c.Cin : (size_t)M * K;
        std::vector<float> A(a_el), B((size_t)N * K), R((size_t)M * N), bias(N), rb((size_t)(c.conv ? c.NB : 1) * N);
        for (auto& x : A) x = bf(frand()); for (auto& x : B) x = bf(frand() * 0.2f); for (auto& x : R) x = bf(frand() * 2.f);
        for (auto& x : bias) x = frand(); for (auto& x : rb) x = frand();
        std::vector<__nv_bfloat16> Ah(a_el), Bh((size_t)N * K), Rh((size_t)M * N);
        for (size_t i = 0; i < a_el; ++i) Ah[i] = __float2bfloat16(A[i]);
        for (size_t i = 0; i < B.size(); ++i) Bh[i] = __float2bfloat16(B[i]);
        for (size_t i = 0; i < R.size(); ++i) Rh[i] = __float2bfloat16(R[i]);
        __nv_bfloat16 *dA, *dB, *dO, *dR; float *dbias, *drb;
        CK(cudaMalloc(&dA, a_el * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dO, R.size() * 2)); CK(cudaMalloc(&dR, R.size() * 2));
        CK(cudaMalloc(&dbias, N * 4)); CK(cudaMalloc(&drb, rb.size() * 4));
        CK(cudaMemcpy(dA, Ah.data(), a_el * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bh.data(), B.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dR, Rh.data(), R.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(drb, rb.data(), rb.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dO, 0xFF, R.size() * 2));
        GemmDesc g;
        int rc = c.conv ? gemm_desc_conv(&g, dA, c.NB, c.H, c.W, c.Cin, dB, c.Cout, 3, 1) : gemm_desc_linear(&g, dA, K, 0, 0, dB, K, 0, 0, M, N, (int)K, 1, 1, 0, 0);
        if (rc) { printf("%s: descriptor failed\n", c.name); ++bad; continue; }
        __nv_bfloat16* out = c.inplace ? dR : dO;
        g.out = out; g.out_fp32 = 0; g.ldo = N; g.bias = c.bias ? dbias : nullptr; g.rowbias = (c.conv && c.rowbias) ? drb : nullptr;
        g.residual = c.residual ? dR : nullptr; g.ldr = N;
        if (c.pair) gemm_enable_pair(&g);
        int ks = c.force_ks ? c.force_ks : 1;
        if (ks > 1) { g.ksplit = ks; g.splitk_ws = ws; }
        if (c.tma) gemm_enable_tma_epilogue(&g);
        g.stages = gemm_choose_stages(g, sm, &g.katoms);
        rc = gemm_launch(g, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (rc || e != cudaSuccess) { printf("%s: launch rc=%d sync=%s\n", c.name, rc, cudaGetErrorString(e)); ++bad; break; }
        std::vector<__nv_bfloat16> Oh(R.size());
        CK(cudaMemcpy(Oh.data(), out, R.size() * 2, cudaMemcpyDeviceToHost));
        double num = 0, den = 0; double worst = 0; long worst_i = -1;
#pragma omp parallel for reduction(+ : num, den)
        for (long row = 0; row < M; ++row) {
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                if (c.conv) {
                    const int img = (int)(row / (c.H * c.W)), hh = (int)(row / c.W) % c.H, ww = (int)(row % c.W);
                    for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
                        const int y = hh + ky - 1, x = ww + kx - 1;
                        if (y < 0 || y >= c.H || x < 0 || x >= c.W) continue;
                        const float* ap = &A[(((size_t)img * c.H + y) * c.W + x) * c.Cin];
                        const float* bp = &B[(size_t)n * K + (size_t)(ky * 3 + kx) * c.Cin];
                        for (int ci = 0; ci < c.Cin; ++ci) acc += (double)ap[ci] * bp[ci];
                    }
                    if (c.rowbias) acc += rb[(size_t)img * N + n];
                } else {
                    const float* ap = &A[(size_t)row * K]; const float* bp = &B[(size_t)n * K];
                    for (long k = 0; k < K; ++k) acc += (double)ap[k] * bp[k];
                }
                if (c.bias) acc += bias[n];
                if (c.residual) acc += R[(size_t)row * N + n];
                const double got = __bfloat162float(Oh[(size_t)row * N + n]);
                const double d = got - acc;
                num += d * d; den += acc * acc;
#pragma omp critical
                if (std::fabs(d) > worst || std::isnan(got)) { worst = std::isnan(got) ? 1e30 : std::fabs(d); worst_i = row * N + n; }
            }
        }
        const double rel = std::sqrt(num / (den + 1e-30));
        printf("%-46s pair=%d bn=%3d tma_epi=%d ks=%d  rel err %.3e  worst |d| %.3e at row %ld col %ld  %s\n", c.name, g.pair, g.pair ? g.bn : 128, g.tma_epi, ks, rel,
               worst, worst_i / N, worst_i % N, rel < 4e-3 ? "ok" : "FAIL");
        if (!(rel < 4e-3)) ++bad;
        cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dR); cudaFree(dbias); cudaFree(drb);
    }
    printf("%s\n", bad ? "GEMM CHECK FAILED" : "gemm check passed");
    return bad ? 1 : 0;
}
