/*
 * nbldpc_mc.c -- Monte-Carlo driver in plain C on top of the C ABI (include/nbldpc_b200.h).
 *
 * Stand-in for the reference's main() (NB_LDPC.c:63-604): same seven positional arguments, same console
 * and results-file lines, same drand48 frame stream (so FER/BER/avr_it are the reference's, frame for
 * frame), but the frames are decoded in batches on the GPU:
 *
 *   nbldpc_mc NbMonteCarlo NbIterMax FileMatrix EbN NbMax Offset NbOper [batch [device [ecn]]]
 *
 * batch  frames per nbgpu_decode_noisy call (default 256); device = CUDA device (default 0);
 * ecn    0 = CheckPassLogEMS (default), 1 = syndrome_ems with the default parameters of nbgpu_params.
 * Differences to the reference, all deliberate: no getchar() at exit, errors are reported instead of
 * exit()ing inside library calls, the '\r' progress line is refreshed once per batch.
 */
#include "nbldpc_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static int fail(const char *what, const nbgpu_ctx *ctx)
{
    fprintf(stderr, "nbldpc_mc: %s: %s\n", what, nbgpu_last_error(ctx));
    return EXIT_FAILURE;
}

int main(int argc, char *argv[])
{
    if (argc < 8) {
        printf("File:\n %s\n ", argv[0]);
        printf("Usage: NbMonteCarlo NbIterMax FileMatrix EbN NbMax Offset NbOper [batch [device [ecn]]]\n"
               "  NbMonteCarlo : # simulated frames\n  NbIterMax    : # of maximum decoding iteration\n"
               "  FileMatrix   : File name of the parity-check matrix (UBS or KN alist)\n  EbN          : Eb/No (dB)\n"
               "  NbMax        : size of truncated messages\n  Offset       : offset correction factor (0.4 -- 1)\n"
               "  NbOper       : Maximum number of operations for sorting\n");
        return EXIT_FAILURE;
    }
    const int NbMonteCarlo = atoi(argv[1]), NbIterMax = atoi(argv[2]);
    const char *FileMatrix = argv[3];
    const float EbN = (float)atof(argv[4]);
    const int NbMax = atoi(argv[5]);
    const float offset = (float)atof(argv[6]);
    const int NbOper = atoi(argv[7]);
    int B = argc > 8 ? atoi(argv[8]) : 256;
    const int device = argc > 9 ? atoi(argv[9]) : 0, ecn = argc > 10 ? atoi(argv[10]) : 0;
    if (B < 1) B = 1;
    if (B > NbMonteCarlo && NbMonteCarlo > 0) B = NbMonteCarlo;

    printf(" Monte-Carlo simulation of Non-Binary LDPC decoder \n\n");                    /* NB_LDPC.c:112-120 */
    printf("Simulation parameters:\n");
    printf("\n\t NbMonteCarlo     : %d", NbMonteCarlo);
    printf("\n\t NbIterMax        : %d", NbIterMax);
    printf("\n\t FileMatrix       : %s", FileMatrix);
    printf("\n\t Eb/No (dB)       : %g", EbN);
    printf("\n\t NbMax            : %d", NbMax);
    printf("\n\t Offset           : %g", offset);
    printf("\n\t NbOper           : %d\n", NbOper);

    nbgpu_code *code = NULL;
    if (nbgpu_code_load(&code, FileMatrix, NBGPU_ALIST_AUTO)) return fail("LoadCode", NULL);
    int info[10];
    nbgpu_code_info(code, info);
    const int N = info[0], M = info[1], K = info[2], GF = info[3], logGF = info[4];
    printf("LDPC code parameters: \n");                                                      /* init.c:267 */
    printf(" \t N \t:%d \n \t K \t:%d \n \t M\t:%d \n \t CR\t:%g \n \t GF \t:%d \n \t logGF \t:%d\n", N, K, M,
           nbgpu_code_rate(code), GF, logGF);
    if (nbgpu_code_prepare_encoder(code)) return fail("GaussianElimination", NULL);

    const char *note = "FB30";                                                               /* NB_LDPC.c:129-136 */
    printf("\n\t Note             : %s\n", note);
    char file_name[256];
    snprintf(file_name, sizeof file_name, "./data/results_N%d_CR%0.2f_GF%d_IT%d_Offset%0.1f_nm%d_%s.txt", N,
             nbgpu_code_rate(code), GF, NbIterMax, offset, NbMax, note);
    time_t now = time(NULL);
    printf("Simulation started at time: %s \n", ctime(&now));

    nbgpu_params p;
    memset(&p, 0, sizeof p);
    p.n_m = NbMax; p.nb_oper = NbOper; p.nb_iter_max = NbIterMax; p.offset = offset; p.ecn_kind = ecn; p.early_stop = 1;
    nbgpu_ctx *ctx = NULL;
    if (nbgpu_create(&ctx, code, &p, device, B)) return fail("AllocateDecoder", NULL);

    const size_t fl = (size_t)N * logGF;
    float *noisy = malloc(sizeof(float) * fl * B);
    int *bits = malloc(sizeof(int) * fl * B), *cw = malloc(sizeof(int) * N);
    int *decide = malloc(sizeof(int) * (size_t)N * B), *synd = malloc(sizeof(int) * B), *iters = malloc(sizeof(int) * B);
    if (!noisy || !bits || !cw || !decide || !synd || !iters) { fprintf(stderr, "nbldpc_mc: out of memory\n"); return EXIT_FAILURE; }
    nbgpu_host_register(noisy, sizeof(float) * fl * B);
    nbgpu_rng rng;
    nbgpu_rng_reference_default(&rng);            /* the reference never seeds drand48 (NB_LDPC.c:88 seeds rand() only) */
    const float sigma = nbgpu_sigma(code, EbN);
    long stats[6] = { 0, 0, 0, 0, 0, 0 };
    int nb = 1;
    while (nb <= NbMonteCarlo && !stats[5]) {
        const int b = (NbMonteCarlo - nb + 1 < B) ? NbMonteCarlo - nb + 1 : B;
        for (int f = 0; f < b; f++) {                                                        /* NB_LDPC.c:252-261 */
            if (nbgpu_random_codeword(code, &rng, cw, bits + fl * f)) return fail("Encoding", NULL);
            nbgpu_awgn_bpsk_noise(code, &rng, bits + fl * f, EbN, noisy + fl * f);
        }
        if (nbgpu_decode_noisy(ctx, noisy, sigma, b, decide, synd, iters)) return fail("decode", ctx);     /* :266-474 */
        if (nbgpu_accumulate_stats(code, bits, decide, synd, iters, b, stats)) return fail("statistics", NULL);
        nb += b;
        const long n = stats[0];
        printf("\r<%ld> FER= %ld / %ld = %f BER= %ld / x = %f  avr_it=%.2f", stats[2], stats[1], n, (double)stats[1] / n, stats[3],
               (double)stats[3] / ((double)n * K * logGF), (double)stats[4] / n);            /* :498-500 */
        fflush(stdout);
    }
    /* 'nb' of the reference after its loop: the frame of the 40th error, else NbMonteCarlo + 1 (:250, :506) */
    const long nb_file = stats[5] ? stats[0] : (long)NbMonteCarlo + 1;
    printf(" \n results are printed in file %s \n", file_name);
    now = time(NULL);
    const char *ts = ctime(&now);
    FILE *op = fopen(file_name, "a");
    if (!op) printf(" \n !! file not found \n ");
    else {                                                                                   /* :576-577 */
        fprintf(op, " SNR:%.2f: \t FER= %ld / %ld = %f ", EbN, stats[1], nb_file, (double)stats[1] / nb_file);
        fprintf(op, " \t BER= %ld / x = \t %f  avr_it= \t %.2f \t time: %s", stats[3],
                (double)stats[3] / ((double)nb_file * K * logGF), (double)stats[4] / nb_file, ts);
        fclose(op);
    }
    printf(" \n results printed \n ");
    printf("\n");
    printf("Simulation complete at time: %s", ts);
    nbgpu_host_unregister(noisy);
    free(noisy); free(bits); free(cw); free(decide); free(synd); free(iters);
    nbgpu_destroy(ctx);
    nbgpu_code_free(code);
    return EXIT_SUCCESS;
}
