/*
 * polar_oracle_trellis.c -- CPU restatement of the deletion-channel path: BinaryTrellis / CollectionOfBinaryTrellises
 * under BinaryPolarEncoderDecoder.decode (uniform a-priori distribution), plus the Guardbands helpers.
 *
 * TEST INFRASTRUCTURE ONLY (see polar_oracle.c).  Follows the reference line by line:
 *   buildTrellis_uniformInput_deletion     VectorDistributions/BinaryTrellis.py:309-438
 *   __miusPlusTransform                    VectorDistributions/BinaryTrellis.py:206-258
 *   calcMarginalizedProbabilities          VectorDistributions/BinaryTrellis.py:260-278
 *   calcNormalizationVector / normalize    VectorDistributions/BinaryTrellis.py:280-306
 *   collection transforms and collapse     VectorDistributions/CollectionOfBinaryTrellises.py:55-103
 *   recursion                              BinaryPolarEncoderDecoder.py:223-325
 *   memoryless arithmetic after collapse   VectorDistributions/BinaryMemorylessVectorDistribution.py:15-87
 *
 * The reference keeps vertices and edges in Python dicts and iterates them in INSERTION order; floating-point sums
 * over edges therefore depend on the order in which vertices and edges were first created.  This restatement keeps the
 * same insertion-ordered lists (vertices per layer, incoming / outgoing edges per vertex), so every sum is taken in the
 * reference's order and results are bit-identical (pinned by tests/golden/trellis.npz from the live reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int from_v, to_v, label; /* indices into the vertex lists of layers l and l+1 */
    double p;
} tedge;

typedef struct {
    int vpos;
    double vprob;
    int nout, nin, cap_out, cap_in;
    int *out, *in; /* edge ids, insertion order */
} tvert;

typedef struct {
    int len;       /* number of inputs = number of edge layers; vertex layers 0..len */
    int *nv, *capv;
    tvert **v;     /* v[layer][k], insertion order */
    int *ne, *cape;
    tedge **e;     /* e[layer][id]: edges leaving layer `layer` */
} trellis;

static void *xm(size_t b) {
    void *p = calloc(1, b ? b : 1);
    if (!p) abort();
    return p;
}

static trellis *t_new(int len) {
    trellis *t = xm(sizeof *t);
    t->len = len;
    t->nv = xm(sizeof(int) * (len + 1));
    t->capv = xm(sizeof(int) * (len + 1));
    t->v = xm(sizeof(tvert *) * (len + 1));
    t->ne = xm(sizeof(int) * (len + 1));
    t->cape = xm(sizeof(int) * (len + 1));
    t->e = xm(sizeof(tedge *) * (len + 1));
    return t;
}

static void t_free(trellis *t) {
    if (!t) return;
    for (int l = 0; l <= t->len; ++l) {
        for (int k = 0; k < t->nv[l]; ++k) {
            free(t->v[l][k].out);
            free(t->v[l][k].in);
        }
        free(t->v[l]);
        free(t->e[l]);
    }
    free(t->nv), free(t->capv), free(t->v), free(t->ne), free(t->cape), free(t->e);
    free(t);
}

/* __getVertexAndAddIfNeeded, BinaryTrellis.py:154-161: returns the index of the vertex in its layer's ordered list */
static int t_vertex(trellis *t, int layer, int vpos) {
    for (int k = 0; k < t->nv[layer]; ++k)
        if (t->v[layer][k].vpos == vpos) return k;
    if (t->nv[layer] == t->capv[layer]) {
        t->capv[layer] = t->capv[layer] ? 2 * t->capv[layer] : 8;
        t->v[layer] = realloc(t->v[layer], sizeof(tvert) * t->capv[layer]);
        if (!t->v[layer]) abort();
    }
    tvert *nvx = &t->v[layer][t->nv[layer]];
    memset(nvx, 0, sizeof *nvx);
    nvx->vpos = vpos;
    nvx->vprob = -1.0;
    return t->nv[layer]++;
}

static void push_int(int **a, int *n, int *cap, int x) {
    if (*n == *cap) {
        *cap = *cap ? 2 * *cap : 8;
        *a = realloc(*a, sizeof(int) * *cap);
        if (!*a) abort();
    }
    (*a)[(*n)++] = x;
}

static void t_set_vertex_prob(trellis *t, int layer, int vpos, double p) {
    const int k = t_vertex(t, layer, vpos); /* may move the layer's array: index first */
    t->v[layer][k].vprob = p;
}

/* addToEdgeProb, BinaryTrellis.py:128-136 + __getEdgeAndAddIfNeeded :163-175 */
static void t_add_edge_prob(trellis *t, int layer, int from_vpos, int to_vpos, int label, double p) {
    const int fk = t_vertex(t, layer, from_vpos);
    const int tk = t_vertex(t, layer + 1, to_vpos);
    tvert *fv = &t->v[layer][fk];
    int id = -1;
    for (int k = 0; k < fv->nout; ++k) {
        const tedge *e = &t->e[layer][fv->out[k]];
        if (e->to_v == tk && e->label == label) {
            id = fv->out[k];
            break;
        }
    }
    if (id < 0) {
        if (t->ne[layer] == t->cape[layer]) {
            t->cape[layer] = t->cape[layer] ? 2 * t->cape[layer] : 16;
            t->e[layer] = realloc(t->e[layer], sizeof(tedge) * t->cape[layer]);
            if (!t->e[layer]) abort();
        }
        id = t->ne[layer]++;
        tedge *e = &t->e[layer][id];
        e->from_v = fk;
        e->to_v = tk;
        e->label = label;
        e->p = 0.0;
        fv = &t->v[layer][fk];
        push_int(&fv->out, &fv->nout, &fv->cap_out, id);
        tvert *tv = &t->v[layer + 1][tk];
        push_int(&tv->in, &tv->nin, &tv->cap_in, id);
    }
    t->e[layer][id].p += p;
}

static double comb_exact(int n, int k) {
    double r = 1.0; /* exact for the small arguments used here (scipy.special.comb(exact=True) is an int) */
    for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    return floor(r + 0.5);
}

/* buildTrellis_uniformInput_deletion, BinaryTrellis.py:309-438 (trimmedZerosAtEdges as given) */
static trellis *t_build(const uint8_t *rw, int rlen, int codewordLength, double d, int trimmed, int ones) {
    trellis *t = t_new(codewordLength);
    const int deletionCount = codewordLength + 2 * ones - rlen;
    const double inputProb[2] = {0.5, 0.5};
    if (ones > 0) {
        const int m = ones < rlen ? ones : rlen;
        for (int i = 0; i < 1 + m; ++i)
            t_set_vertex_prob(t, 0, i, comb_exact(ones, i) * pow(1.0 - d, i) * pow(d, ones - i));
    } else {
        t_set_vertex_prob(t, 0, 0, 1.0);
    }
    if (ones > 0) {
        const int m = ones < rlen ? ones : rlen;
        for (int i = rlen; i > rlen - m - 1; --i) {
            const int j = rlen - i;
            t_set_vertex_prob(t, codewordLength, i, comb_exact(ones, j) * pow(1.0 - d, j) * pow(d, ones - j));
        }
    } else {
        t_set_vertex_prob(t, codewordLength, rlen, 1.0);
    }
    for (int l = 0; l < codewordLength; ++l) {
        int vmin, vmax;
        if (ones > 0) {
            vmin = l + ones - deletionCount > 0 ? l + ones - deletionCount : 0;
            vmax = l + ones < rlen ? l + ones : rlen;
        } else {
            vmin = l - deletionCount > 0 ? l - deletionCount : 0;
            vmax = l < rlen ? l : rlen;
        }
        for (int vpos = vmin; vpos <= vmax; ++vpos) {
            if (vpos < rlen) {
                const int label = rw[vpos];
                t_add_edge_prob(t, l, vpos, vpos + 1, label, inputProb[label] * (1.0 - d));
            }
            if (l + 1 + ones - deletionCount <= vpos) {
                for (int label = 0; label < 2; ++label) {
                    double p;
                    if (!trimmed || label == 1 || (vpos > 0 && vpos < rlen))
                        p = inputProb[label] * d;
                    else
                        p = inputProb[label];
                    t_add_edge_prob(t, l, vpos, vpos, label, p);
                }
            }
        }
    }
    return t;
}

/* __miusPlusTransform, BinaryTrellis.py:206-258; decision == NULL: minus */
static trellis *t_transform(const trellis *t, const int64_t *decision) {
    trellis *nt = t_new(t->len / 2);
    for (int k = 0; k < t->nv[0]; ++k) t_set_vertex_prob(nt, 0, t->v[0][k].vpos, t->v[0][k].vprob);
    for (int k = 0; k < t->nv[t->len]; ++k) t_set_vertex_prob(nt, t->len / 2, t->v[t->len][k].vpos, t->v[t->len][k].vprob);
    for (int ml = 1; ml <= t->len; ml += 2) {
        for (int wk = 0; wk < t->nv[ml]; ++wk) {
            const tvert *w = &t->v[ml][wk];
            for (int a = 0; a < w->nin; ++a) {
                const tedge *ein = &t->e[ml - 1][w->in[a]];
                for (int b = 0; b < w->nout; ++b) {
                    const tedge *eout = &t->e[ml][w->out[b]];
                    const int u_vpos = t->v[ml - 1][ein->from_v].vpos, v_vpos = t->v[ml + 1][eout->to_v].vpos;
                    const int x0 = ein->label, x1 = eout->label;
                    const double np_ = ein->p * eout->p;
                    const int mlabel = x0 != x1 ? 1 : 0;
                    if (!decision) {
                        t_add_edge_prob(nt, (ml - 1) / 2, u_vpos, v_vpos, mlabel, np_);
                    } else {
                        if (mlabel != (int)decision[(ml - 1) / 2]) continue;
                        t_add_edge_prob(nt, (ml - 1) / 2, u_vpos, v_vpos, x1, np_);
                    }
                }
            }
        }
    }
    return nt;
}

/* calcMarginalizedProbabilities(normalize=False), BinaryTrellis.py:260-278 */
static void t_marginal_unnormalized(const trellis *t, double *m) {
    m[0] = m[1] = 0.0;
    const double s = 1.0;
    for (int k = 0; k < t->nv[0]; ++k) {
        const tvert *v = &t->v[0][k];
        for (int a = 0; a < v->nout; ++a) {
            const tedge *e = &t->e[0][v->out[a]];
            m[e->label] += v->vprob * e->p * t->v[1][e->to_v].vprob / s;
        }
    }
}

/* calcNormalizationVector + normalize, BinaryTrellis.py:280-306 */
static void t_normalize(trellis *t) {
    for (int i = 0; i < t->len; ++i) {
        double tp[2] = {0.0, 0.0};
        for (int k = 0; k < t->nv[i]; ++k) {
            const tvert *v = &t->v[i][k];
            for (int a = 0; a < v->nout; ++a) {
                const tedge *e = &t->e[i][v->out[a]];
                tp[e->label] += e->p;
            }
        }
        double nrm = tp[0] > tp[1] ? tp[0] : tp[1]; /* np.maximum */
        if (nrm == 0) nrm = 1;
        for (int k = 0; k < t->nv[i]; ++k) {
            const tvert *v = &t->v[i][k];
            for (int a = 0; a < v->nout; ++a) t->e[i][v->out[a]].p /= nrm;
        }
    }
}

/* ---- memoryless arithmetic after the collapse (BinaryMemorylessVectorDistribution.py:15-87) ---------------------- */
static void m_minus(const double *p, int len, double *o) {
    for (int h = 0; h < len / 2; ++h) {
        const double *a = p + 4 * h, *b = a + 2;
        o[2 * h] = a[0] * b[0] + a[1] * b[1];
        o[2 * h + 1] = a[0] * b[1] + a[1] * b[0];
    }
}
static void m_plus(const double *p, int len, const int64_t *u, double *o) {
    for (int h = 0; h < len / 2; ++h) {
        const double *a = p + 4 * h, *b = a + 2;
        if (u[h] == 0) {
            o[2 * h] = a[0] * b[0];
            o[2 * h + 1] = a[1] * b[1];
        } else {
            o[2 * h] = a[1] * b[0];
            o[2 * h + 1] = a[0] * b[1];
        }
    }
}
static void m_normalize(double *p, int len) {
    for (int i = 0; i < len; ++i) {
        double t = p[2 * i] > p[2 * i + 1] ? p[2 * i] : p[2 * i + 1];
        if (t == 0) t = 1;
        p[2 * i] /= t;
        p[2 * i + 1] /= t;
    }
}

typedef struct {
    const uint8_t *frozen;
    const double *r;
    int64_t *info;
    int info_idx;
    double *collapse; /* optional: concatenation of the collapsed (unnormalised) vectors in visiting order */
    int collapse_pos;
} tctx;

/* memoryless recursion, BinaryPolarEncoderDecoder.py:245-325 with a uniform prior (leaf marginal of the prior = 0.5) */
static void rec_mem(tctx *c, double *xy, int len, int u0, int64_t *out) {
    if (len == 1) {
        int64_t bit;
        if (c->frozen[u0]) {
            bit = 0.5 >= c->r[u0] ? 0 : 1;
        } else {
            const double s = xy[0] + xy[1];
            double m0 = 0.5, m1 = 0.5;
            if (s > 0.0) {
                m0 = xy[0] / s;
                m1 = xy[1] / s;
            }
            bit = m0 >= m1 ? 0 : 1;
            c->info[c->info_idx++] = bit;
        }
        out[0] = bit;
        return;
    }
    const int half = len / 2;
    double *ch = xm(sizeof(double) * 2 * half);
    int64_t *mb = xm(sizeof(int64_t) * half), *pb = xm(sizeof(int64_t) * half);
    m_minus(xy, len, ch);
    m_normalize(ch, half);
    rec_mem(c, ch, half, u0, mb);
    m_plus(xy, len, mb, ch);
    m_normalize(ch, half);
    rec_mem(c, ch, half, u0 + half, pb);
    for (int h = 0; h < half; ++h) {
        out[2 * h] = (mb[h] + pb[h]) % 2;
        out[2 * h + 1] = pb[h];
    }
    free(ch), free(mb), free(pb);
}

/* collection recursion, CollectionOfBinaryTrellises.py:55-82 under BinaryPolarEncoderDecoder.py:276-325 */
static void rec_coll(tctx *c, trellis **tr, int ntr, int len, int u0, int64_t *out) {
    const int half = len / 2;
    const int tlen = len / ntr;     /* current trellis length */
    const int sub = half / ntr;     /* decision bits per trellis */
    int64_t *mb = xm(sizeof(int64_t) * half), *pb = xm(sizeof(int64_t) * half);
    for (int pass = 0; pass < 2; ++pass) {
        int64_t *bits = pass == 0 ? mb : pb;
        const int uu = pass == 0 ? u0 : u0 + half;
        if (half > ntr) {
            trellis **ch = xm(sizeof(trellis *) * ntr);
            for (int i = 0; i < ntr; ++i) {
                ch[i] = t_transform(tr[i], pass == 0 ? NULL : mb + (size_t)i * sub);
                t_normalize(ch[i]);
            }
            rec_coll(c, ch, ntr, half, uu, bits);
            for (int i = 0; i < ntr; ++i) t_free(ch[i]);
            free(ch);
        } else { /* tlen == 2: collapse to a memoryless vector of length ntr, CollectionOfBinaryTrellises.py:68-82 */
            double *pr = xm(sizeof(double) * 2 * ntr);
            for (int i = 0; i < ntr; ++i) {
                trellis *tt = t_transform(tr[i], pass == 0 ? NULL : mb + (size_t)i * sub);
                t_marginal_unnormalized(tt, pr + 2 * i);
                t_free(tt);
            }
            if (c->collapse) {
                memcpy(c->collapse + c->collapse_pos, pr, sizeof(double) * 2 * ntr);
                c->collapse_pos += 2 * ntr;
            }
            m_normalize(pr, ntr);
            rec_mem(c, pr, ntr, uu, bits);
            free(pr);
        }
    }
    (void)tlen;
    for (int h = 0; h < half; ++h) {
        out[2 * h] = (mb[h] + pb[h]) % 2;
        out[2 * h + 1] = pb[h];
    }
    free(mb), free(pb);
}

/* BinaryPolarEncoderDecoder.decode over buildCollectionOfBinaryTrellises_uniformInput_deletion (uniform prior).
 * sub_bits [T][maxlen], sub_len [T]: the trimmed sub-words (Guardbands.removeDeletionGuardBands), T = 2^(n-n0).
 * collapse (optional) receives the 2^n0 collapsed unnormalised vectors [2^n0][T][2] in visiting order. */
int po_trellis_decode(int n, int n0, const uint8_t *frozen, const double *r, const uint8_t *sub_bits, const int32_t *sub_len,
                      int maxlen, double deletionProb, int ones, int64_t *cw_out, int64_t *info_out, double *collapse) {
    if (n0 < 1 || n0 > n) return -1;
    const int N = 1 << n, T = 1 << (n - n0), tl = 1 << n0;
    trellis **tr = xm(sizeof(trellis *) * T);
    for (int i = 0; i < T; ++i) tr[i] = t_build(sub_bits + (size_t)i * maxlen, sub_len[i], tl, deletionProb, 1, ones);
    tctx c = {frozen, r, info_out, 0, collapse, 0};
    if (N == 1) return -1;
    if (T == 1 && tl == 1) return -1;
    rec_coll(&c, tr, T, N, 0, cw_out);
    for (int i = 0; i < T; ++i) t_free(tr[i]);
    free(tr);
    return 0;
}

/* ---- Guardbands.py ------------------------------------------------------------------------------------------------ */
/* trimZerosAtEdges, Guardbands.py:66-93: returns the new length, *start = first kept index */
static int trim_zeros(const uint8_t *w, int len, int *start) {
    int first = -1, last = -1;
    for (int i = 0; i < len; ++i)
        if (w[i] == 1) {
            first = i;
            break;
        }
    if (first < 0) {
        *start = 0;
        return 0;
    }
    for (int i = len - 1; i >= 0; --i)
        if (w[i] == 1) {
            last = i;
            break;
        }
    *start = first;
    return last - first + 1;
}

static void remove_gb(const uint8_t *w, int len, int n, int n0, uint8_t *sub_bits, int32_t *sub_len, int maxlen, int *idx, int *overflow) {
    int st;
    const int tl = trim_zeros(w, len, &st);
    const uint8_t *tw = w + st;
    if (n <= n0) {
        sub_len[*idx] = tl;
        if (tl > maxlen) *overflow = 1;
        memcpy(sub_bits + (size_t)(*idx) * maxlen, tw, tl < maxlen ? tl : maxlen);
        ++*idx;
        return;
    }
    remove_gb(tw, tl / 2, n - 1, n0, sub_bits, sub_len, maxlen, idx, overflow);
    remove_gb(tw + tl / 2, tl - tl / 2, n - 1, n0, sub_bits, sub_len, maxlen, idx, overflow);
}

/* removeDeletionGuardBands, Guardbands.py:47-63.  Returns 0, or 1 when a sub-word is longer than maxlen. */
int po_remove_guard_bands(const uint8_t *received, int len, int n, int n0, uint8_t *sub_bits, int32_t *sub_len, int maxlen) {
    int idx = 0, overflow = 0;
    memset(sub_bits, 0, (size_t)maxlen << (n - n0));
    remove_gb(received, len, n, n0, sub_bits, sub_len, maxlen, &idx, &overflow);
    return overflow;
}

static int add_gb(const uint8_t *enc, int len, int n, int n0, double xi, int ones, uint8_t *out) {
    if (n <= n0) {
        int o = 0;
        for (int i = 0; i < ones; ++i) out[o++] = 1;
        memcpy(out + o, enc, len);
        o += len;
        for (int i = 0; i < ones; ++i) out[o++] = 1;
        return o;
    }
    const int ln = (int)floor(pow(2.0, (1.0 - xi) * (n - 1)));
    int o = add_gb(enc, len / 2, n - 1, n0, xi, ones, out);
    memset(out + o, 0, ln);
    o += ln;
    o += add_gb(enc + len / 2, len - len / 2, n - 1, n0, xi, ones, out + o);
    return o;
}

/* addDeletionGuardBands, Guardbands.py:4-44; out must hold the result (query with out == NULL is not supported: size it
 * as N + 2^(n-n0) 2 ones + sum of guard bands <= 4 N + ...; the Python wrapper computes the exact size) */
int po_add_guard_bands(const uint8_t *enc, int n, int n0, double xi, int ones, uint8_t *out) {
    return add_gb(enc, 1 << n, n, n0, xi, ones, out);
}
