/*
 * oracle/nbldpc_oracle_synd.c -- TEST INFRASTRUCTURE ONLY; included by nbldpc_oracle.c (NBO_HAVE_SYND).
 *
 * CPU restatement of the syndrome-based check node of the reference (syndrome_decoder.c): configuration
 * table (build_config_table -> gen_config_table2, sort_config_table), presorting_mvc, syndrome_ems,
 * sorting and bayes.  The reference ships this path with its call site commented out
 * (NB_LDPC.c:189-201, 388); what is pinned here is the behaviour of the functions themselves, checked
 * against the compiled reference through oracle/_ref/libref.so (tests/test_oracle_pinned.py).
 */

/* factorial / combin, syndrome_decoder.c:2122-2139 */
static int nbo_fact(int p) { int f = 1, i; for (i = 1; i <= p; i++) f *= i; return f; }
static int nbo_combin(int n, int r) { return (n < r) ? 0 : nbo_fact(n) / (nbo_fact(r) * nbo_fact(n - r)); }

/* stable insertion sort by LLR, syndrome_decoder.c:1315-1334 (strict '>' while shifting) */
typedef struct { int GF; float LLR; int config; } nbo_synd;
static void nbo_sorting(nbo_synd *s, int n)
{
    int i, j;
    for (j = 1; j < n; j++) {
        nbo_synd t = s[j];
        i = j - 1;
        while (i >= 0 && s[i].LLR > t.LLR) { s[i + 1] = s[i]; i--; }
        s[i + 1] = t;
    }
}

/* build_config_table (syndrome_decoder.c:1542-1574) = gen_config_table2 (:1661-1767, including its
 * four-deviation block with d_4 = 2) into a zero-filled table, then sort_config_table (:2285-2371:
 * cost = sum over deviating edges of (deviation + 3.0*edge), stable insertion sort), then the
 * truncation of NB_LDPC.c:198-201 restricted to the generated size.  Returns malloc'd [size][dc]. */
int *nbo_build_config_table(int dc, int d1, int d2, int d3, int trunc, int *size_out)
{
    const int d4 = 2;
    int cap = 1 + dc * d1 + nbo_combin(dc, 2) * d2 * d2 + nbo_combin(dc, 3) * d3 * d3 * d3;   /* compute_config_table_size, :1523 */
    int gen = 1 + dc * d1 + nbo_combin(dc, 2) * (d2 * (d2 + 1) / 2) + nbo_combin(dc, 3) * (d3 * (d3 + 1) * (d3 + 2) / 6)
              + nbo_combin(dc, 4) * 8;
    if (gen > cap) cap = gen;                        /* the reference would overrun its table here; never for the shipped shapes */
    int *tab = calloc((size_t)cap * dc, sizeof(int));
    int cfg = 1, i, j, k, l, m, n, o, p;
    for (i = 0; i < dc; i++) for (j = 0; j < d1; j++) { tab[cfg * dc + i] = j + 1; cfg++; }
    for (i = 0; i < dc - 1; i++) for (j = i + 1; j < dc; j++)
        for (k = 0; k < d2; k++) for (l = 0; l < d2; l++)
            if (k + l < d2) { tab[cfg * dc + i] = k + 1; tab[cfg * dc + j] = l + 1; cfg++; }
    for (i = 0; i < dc - 2; i++) for (j = i + 1; j < dc - 1; j++) for (k = j + 1; k < dc; k++)
        for (l = 0; l < d3; l++) for (m = 0; m < d3; m++) for (n = 0; n < d3; n++)
            if (l + m + n < d3) { tab[cfg * dc + i] = l + 1; tab[cfg * dc + j] = m + 1; tab[cfg * dc + k] = n + 1; cfg++; }
    for (o = 0; o < dc - 3; o++) for (i = o + 1; i < dc - 2; i++) for (j = i + 1; j < dc - 1; j++) for (k = j + 1; k < dc; k++)
        for (l = 0; l < d4; l++) for (m = 0; m < d4; m++) for (n = 0; n < d4; n++) for (p = 0; p < d4; p++)
            if (l + m + n < d4) {
                tab[cfg * dc + i] = l + 1; tab[cfg * dc + j] = m + 1; tab[cfg * dc + k] = n + 1; tab[cfg * dc + o] = p + 1; cfg++;
            }
    /* sort_config_table */
    nbo_synd *s = malloc(sizeof(nbo_synd) * cfg);
    for (i = 0; i < cfg; i++) {
        float cost = 0;
        for (j = 0; j < dc; j++) if (tab[i * dc + j] > 0) cost = cost + tab[i * dc + j] + 3.0 * j;   /* float <- double sum, :2298 */
        s[i].LLR = cost; s[i].config = i; s[i].GF = 0;
    }
    nbo_sorting(s, cfg);
    int size = (trunc > 0 && trunc < cfg) ? trunc : cfg;
    int *out = malloc(sizeof(int) * (size_t)size * dc);
    for (i = 0; i < size; i++) memcpy(out + i * dc, tab + s[i].config * dc, sizeof(int) * dc);
    free(s); free(tab);
    *size_out = size;
    return out;
}

/* bayes, syndrome_decoder.c:2142-2211: double arguments, float locals, double constants */
static double nbo_bayes(double M1, double M2)
{
    float dif, min;
    if (M1 < M2) { min = M1; dif = M2 - M1; }
    else { min = M2; dif = M1 - M2; }
    if (dif < 0.1) min = 0.5 * min;
    else if (dif < 0.2) min = 0.75 * min;
    else if (dif < 1) min = 0.825 * min;
    else if (dif < 2) min = 0.9375 * min;
    return min;
}

/* selection used by presorting_mvc (:318-339, :416-431): n rounds of "strict '<' from 10000, mask with 15000" */
static void nbo_presort_order(float *key, int n, int *order)
{
    int i, j, index_min = 0;
    for (i = 0; i < n; i++) {
        float min = 10000.0f;
        for (j = 0; j < n; j++) if (key[j] < min) { min = key[j]; index_min = j; }
        key[index_min] = 15000.0f;
        order[i] = index_min;
    }
}

/* syndrome_ems, syndrome_decoder.c:26-284 (with presorting_mvc :289-496, border = 4 as at :56).
 * vllr/vgf[dc][n_m]: M_VtoC_LLR/GF on entry; cllr/cgf[dc][GF]: M_CtoV_LLR/GF on return. */
void nbo_check_node_syndrome(const nbo_code *c, int node, const float *vllr, const int *vgf,
                             float *cllr, int *cgf, int n_m, const int *cfg, int cfg_size,
                             float offset, int n_cv)
{
    enum { DCMAX = 32, NM = 64 };
    const int GF = c->GF, dc = c->row_deg[node], border = 4;
    const int *h = c->val + c->row_ptr[node];
    float L[DCMAX][NM], TL[DCMAX][NM], key[DCMAX];
    int G[DCMAX][NM], TG[DCMAX][NM], order[DCMAX], order2[DCMAX], tmp_order[DCMAX];
    int i, j, t, k, d;
    for (t = 0; t < dc; t++) for (k = 0; k < n_m; k++) {                           /* rotation in, :41-48 */
        L[t][k] = vllr[t * n_m + k];
        G[t][k] = c->mulgf[vgf[t * n_m + k] * GF + h[t]];
    }
    /* presorting_mvc: edges ascending by their 2nd LLR ... */
    for (i = 0; i < dc; i++) key[i] = L[i][1];
    nbo_presort_order(key, dc, order);
    for (i = 0; i < dc; i++) { memcpy(TL[i], L[order[i]], sizeof(float) * n_m); memcpy(TG[i], G[order[i]], sizeof(int) * n_m); }
    for (i = 0; i < dc; i++) { memcpy(L[i], TL[i], sizeof(float) * n_m); memcpy(G[i], TG[i], sizeof(int) * n_m); }
    /* ... then the first 'border' edges by their 3rd LLR, :377-469 */
    for (i = 0; i < border; i++) key[i] = L[i][2];
    nbo_presort_order(key, border, order2);
    for (i = 0; i < border; i++) { memcpy(TL[i], L[order2[i]], sizeof(float) * n_m); memcpy(TG[i], G[order2[i]], sizeof(int) * n_m); }
    for (i = 0; i < border; i++) { memcpy(L[i], TL[i], sizeof(float) * n_m); memcpy(G[i], TG[i], sizeof(int) * n_m); }
    for (i = 0; i < border; i++) tmp_order[i] = order[order2[i]];
    for (i = 0; i < border; i++) order[i] = tmp_order[i];

    /* syndromes, :64-77: LLR accumulated edge 0 -> dc-1 from 0, GF by ADDGF */
    nbo_synd *S = malloc(sizeof(nbo_synd) * cfg_size);
    for (i = 0; i < cfg_size; i++) {
        S[i].LLR = 0; S[i].GF = 0; S[i].config = i;
        for (j = 0; j < dc; j++) {
            S[i].LLR = S[i].LLR + L[j][cfg[i * dc + j]];
            S[i].GF = c->addgf[S[i].GF * GF + G[j][cfg[i * dc + j]]];
        }
    }
    nbo_sorting(S, cfg_size);                                                       /* :81 */

    float *D_llr = malloc(sizeof(float) * cfg_size);
    int *D_gf = malloc(sizeof(int) * cfg_size);
    int *updated = malloc(sizeof(int) * GF);
    float (*OL)[256] = malloc(sizeof(float) * 256 * DCMAX);
    for (d = 0; d < dc; d++) {                                                      /* decorrelator, :93-228 */
        int n = 0;
        for (j = 0; j < cfg_size; j++)
            if (cfg[S[j].config * dc + d] == 0) { D_llr[n] = S[j].LLR; D_gf[n] = c->addgf[S[j].GF * GF + G[d][0]]; n++; }
        for (j = 0; j < GF; j++) { updated[j] = 0; OL[d][j] = 1500.0f; }
        for (j = 0; j < n; j++) {
            if (updated[D_gf[j]] == 1) OL[d][D_gf[j]] = nbo_bayes(D_llr[j], OL[d][D_gf[j]]);
            else { OL[d][D_gf[j]] = D_llr[j]; updated[D_gf[j]] = 1; }
        }
        float sat = D_llr[n_cv - 1 + 3 * d];                                        /* :195 (index must be < n) */
        for (j = 0; j < GF; j++) if (OL[d][j] > sat) OL[d][j] = sat + offset;
    }
    /* reorder (:234-253) and rotation out (:260-266): M_CtoV_GF[t][k] = DIVGF[k][h_t] */
    for (i = 0; i < dc; i++)
        for (j = 0; j < GF; j++) { cllr[order[i] * GF + j] = OL[i][j]; }
    for (t = 0; t < dc; t++) for (k = 0; k < GF; k++) cgf[t * GF + k] = c->divgf[k * GF + h[t]];
    free(S); free(D_llr); free(D_gf); free(updated); free(OL);
}
