/*
 * nbgpu_bind.c -- the file a maintainer of the reference adds to its tree to run on libnbldpc_b200.so.
 *
 * It is compiled AGAINST THE REFERENCE'S OWN HEADERS (struct.h: code_t, table_t, decoder_t) and the C ABI
 * (include/nbldpc_b200.h); nothing in the product library depends on it.  Two things live here:
 *
 *   nbgpu_bind_code       code_t + table_t (as LoadCode / LoadTables filled them, NB_LDPC.c:122-123)  ->  nbgpu_code
 *   nbgpu_CheckPassLogEMS drop-in for boundary 1 (include/bubble_decoder.h:17): same signature, same in/out through
 *                         decoder->M_VtoC_* / decoder->M_CtoV_*, the check node itself runs on the GPU
 *                         (nbgpu_check_node).  Building the reference with -DCheckPassLogEMS=nbgpu_CheckPassLogEMS
 *                         swaps it in without touching a line of NB_LDPC.c.
 *
 * oracle/Makefile builds oracle/_ref/essai_gpucn this way (unmodified reference main() + this file + the library);
 * tests/test_gpu_parity.py::test_reference_main_with_gpu_check_node compares its console with the stock binary's.
 * The per-frame boundary (NB_LDPC.c:266-474 -> nbgpu_decode_noisy) is shown in INTEGRATION.md section 3 and is what
 * csrc/nbldpc_mc.c implements.
 */
#include "struct.h"            /* the reference's: code_t, table_t, decoder_t */
#include "nbldpc_b200.h"
#include <stdio.h>
#include <stdlib.h>

/* flatten the reference's row-pointer arrays once, after LoadCode/LoadTables (NB_LDPC.c:122-123) */
nbgpu_code *nbgpu_bind_code(const code_t *code, const table_t *table)
{
    const int E = code->nbBranch, q = code->GF, lq = code->logGF;
    int m, t, e = 0, a, b;
    int *col = malloc(sizeof(int) * E), *val = malloc(sizeof(int) * E);
    int *bin = malloc(sizeof(int) * q * lq), *add = malloc(sizeof(int) * q * q), *mul = malloc(sizeof(int) * q * q), *dv = malloc(sizeof(int) * q * q);
    nbgpu_code *c = NULL;
    if (!col || !val || !bin || !add || !mul || !dv) { fprintf(stderr, "nbgpu_bind_code: out of memory\n"); return NULL; }
    for (m = 0; m < code->M; m++)
        for (t = 0; t < code->rowDegree[m]; t++, e++) { col[e] = code->mat[m][t]; val[e] = code->matValue[m][t]; }
    for (a = 0; a < q; a++) {
        for (b = 0; b < lq; b++) bin[a * lq + b] = table->BINGF[a][b];
        for (b = 0; b < q; b++) { add[a * q + b] = table->ADDGF[a][b]; mul[a * q + b] = table->MULGF[a][b]; dv[a * q + b] = table->DIVGF[a][b]; }
    }
    if (nbgpu_code_from_arrays(&c, code->N, code->M, q, code->rowDegree, col, val, bin, add, mul, dv)) {
        fprintf(stderr, "nbgpu_bind_code: %s\n", nbgpu_last_error(NULL));
        c = NULL;
    }
    free(col); free(val); free(bin); free(add); free(mul); free(dv);
    return c;
}

/* one context per process, created by the first call with that call's parameters (the reference keeps them fixed for a run) */
static nbgpu_code *g_code;
static nbgpu_ctx *g_ctx;
static float *g_vllr, *g_cllr;
static int *g_vgf, *g_cgf;

static void nbgpu_unbind(void)
{
    nbgpu_destroy(g_ctx); nbgpu_code_free(g_code);
    free(g_vllr); free(g_cllr); free(g_vgf); free(g_cgf);
    g_ctx = NULL; g_code = NULL;
}

/* void CheckPassLogEMS(int node, decoder_t *decoder, code_t *code, table_t *table, int NbOper, float offset), bubble_decoder.c:72 */
void nbgpu_CheckPassLogEMS(int node, decoder_t *decoder, code_t *code, table_t *table, int NbOper, float offset)
{
    const int q = code->GF, n_m = decoder->nbMax, dc = code->rowDegree[node];
    int t, k;
    if (!g_ctx) {
        nbgpu_params p = { 0 };
        int dcmax = 0, m;
        for (m = 0; m < code->M; m++) if (code->rowDegree[m] > dcmax) dcmax = code->rowDegree[m];
        p.n_m = n_m; p.nb_oper = NbOper; p.nb_iter_max = 2; p.offset = offset; p.ecn_kind = 0; p.early_stop = 1;
        g_code = nbgpu_bind_code(code, table);
        if (!g_code || nbgpu_create(&g_ctx, g_code, &p, 0, 1)) { fprintf(stderr, "nbgpu_CheckPassLogEMS: %s\n", nbgpu_last_error(NULL)); exit(EXIT_FAILURE); }
        g_vllr = malloc(sizeof(float) * dcmax * n_m); g_vgf = malloc(sizeof(int) * dcmax * n_m);
        g_cllr = malloc(sizeof(float) * dcmax * q); g_cgf = malloc(sizeof(int) * dcmax * q);
        if (!g_vllr || !g_vgf || !g_cllr || !g_cgf) { fprintf(stderr, "nbgpu_CheckPassLogEMS: out of memory\n"); exit(EXIT_FAILURE); }
        atexit(nbgpu_unbind);
    }
    for (t = 0; t < dc; t++)
        for (k = 0; k < n_m; k++) { g_vllr[t * n_m + k] = decoder->M_VtoC_LLR[t][k]; g_vgf[t * n_m + k] = decoder->M_VtoC_GF[t][k]; }
    if (nbgpu_check_node(g_ctx, node, g_vllr, g_vgf, g_cllr, g_cgf, 1)) { fprintf(stderr, "nbgpu_CheckPassLogEMS: %s\n", nbgpu_last_error(g_ctx)); exit(EXIT_FAILURE); }
    for (t = 0; t < dc; t++)
        for (k = 0; k < q; k++) { decoder->M_CtoV_LLR[t][k] = g_cllr[t * q + k]; decoder->M_CtoV_GF[t][k] = g_cgf[t * q + k]; }
}
