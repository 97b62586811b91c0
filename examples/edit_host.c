/*
 * examples/edit_host.c — a plain-C consumer of the drop-in boundary (include/uce_b200.h), no Python, no torch:
 * the whole edit of trainscripts/uce_sd_erase.py:45-88 from HOST buffers — guide outputs, accumulation and solve
 * (uce_edit_host_f32), then the safetensors artifact (uce_artifact_write_f32) — for a small synthetic problem.
 *
 *   gcc -std=c11 -O2 -Iinclude examples/edit_host.c -Lunified-concept-editing_b200 -luce_b200 \
 *       -Wl,-rpath,$PWD/unified-concept-editing_b200 -lm -o /tmp/edit_host && /tmp/edit_host /tmp/edited.safetensors
 *
 * Exit status: 0 edited and written; 3 no sm_100 device (the library has no CPU path: UCE_E_NO_DEVICE); 1 any other failure.
 * The closed form is checked on the host for the first projection:  W_new (lam I + C^T S C) = lam W_old + W_old G^T S C.
 */
#include "uce_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define K 64
#define N_EDIT 3
#define N_ROWS 8            /* 3 edit rows + 5 preserve rows */
#define N_LAYERS 2

static float rnd(unsigned *s) { *s = *s * 1664525u + 1013904223u; return ((float)(*s >> 8) / 8388608.0f) - 1.0f; }

int main(int argc, char **argv) {
    const char *out_path = argc > 1 ? argv[1] : "edited.safetensors";
    const int d[N_LAYERS] = {40, 24};
    const char *names[N_LAYERS] = {"down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight",
                                   "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_v.weight"};
    static float C[N_ROWS * K], G[N_EDIT * K], scales[N_ROWS];
    float *W_old[N_LAYERS], *W_new[N_LAYERS];
    unsigned seed = 7;
    if (uce_abi_version() != UCE_B200_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    for (int i = 0; i < N_ROWS * K; ++i) C[i] = 3.0f * rnd(&seed);        /* edit rows first, then preserve rows */
    for (int i = 0; i < N_EDIT * K; ++i) G[i] = 3.0f * rnd(&seed);        /* guide row of every edit row */
    for (int i = 0; i < N_ROWS; ++i) scales[i] = 1.0f;
    for (int l = 0; l < N_LAYERS; ++l) {
        W_old[l] = malloc(sizeof(float) * d[l] * K); W_new[l] = malloc(sizeof(float) * d[l] * K);
        for (int i = 0; i < d[l] * K; ++i) W_old[l][i] = 0.03f * rnd(&seed);
    }
    uce_ws *ws = NULL;
    int rc = uce_ws_create(0, K, N_ROWS, &ws);
    if (rc == UCE_E_NO_DEVICE) { fprintf(stderr, "no B200: %s\n", uce_last_error()); return 3; }
    if (rc) { fprintf(stderr, "uce_ws_create: %d %s\n", rc, uce_last_error()); return 1; }
    const float lamb = 0.5f;
    rc = uce_edit_host_f32(ws, C, G, scales, N_ROWS, N_EDIT, lamb, (const float *const *)W_old, W_new, d, N_LAYERS);
    if (rc) { fprintf(stderr, "uce_edit_host_f32: %d %s\n", rc, uce_last_error()); return 1; }

    /* residual of the normal equations for projection 0, in double */
    double worst = 0.0, scale = 0.0;
    for (int r = 0; r < d[0]; ++r)
        for (int c = 0; c < K; ++c) {
            double lhs = lamb * W_new[0][r * K + c], rhs = lamb * W_old[0][r * K + c];
            for (int i = 0; i < N_ROWS; ++i) {
                const float *g = i < N_EDIT ? &G[i * K] : &C[i * K];            /* preserve rows guide themselves */
                double wn = 0.0, wo = 0.0;
                for (int k = 0; k < K; ++k) { wn += (double)W_new[0][r * K + k] * C[i * K + k]; wo += (double)W_old[0][r * K + k] * g[k]; }
                lhs += scales[i] * wn * C[i * K + c]; rhs += scales[i] * wo * C[i * K + c];
            }
            if (fabs(lhs - rhs) > worst) worst = fabs(lhs - rhs);
            if (fabs(rhs) > scale) scale = fabs(rhs);
        }
    printf("normal-equation residual %.3e (relative to %.3e)\n", worst, scale);
    const long rows[N_LAYERS] = {d[0], d[1]}, cols[N_LAYERS] = {K, K};
    rc = uce_artifact_write_f32(out_path, N_LAYERS, names, (const float *const *)W_new, rows, cols);
    if (rc) { fprintf(stderr, "uce_artifact_write_f32: %d %s\n", rc, uce_last_error()); return 1; }
    uce_ws_destroy(ws);
    for (int l = 0; l < N_LAYERS; ++l) { free(W_old[l]); free(W_new[l]); }
    return worst <= 1e-4 * scale ? 0 : 1;
}
