/*
 * polar_oracle_list.c -- CPU restatement of QaryPolarEncoderDecoder.listDecode (SC-list decoding with
 * Rate-0 / Rep / Rate-1 / SPC fast nodes), linear probability domain, float64.
 *
 * TEST INFRASTRUCTURE ONLY (see polar_oracle.c).  Follows the reference line by line:
 *   listDecode            QaryPolarEncoderDecoder.py:118-227
 *   recursiveListDecode   QaryPolarEncoderDecoder.py:403-757 (Rate-0 :495, Rep :521, Rate-1 :581, SPC :631, general :684)
 *   pickLeastReliableIndices / reliability / forkIndices / forkIndicesSpc   :759-820
 *   normalize             :867-872
 *   f / g / sum-normalise VectorDistributions/QaryMemorylessVectorDistribution.py:26-118
 *
 * One thing in the reference is NOT a function of its inputs alone: np.argpartition leaves the kept
 * candidates in an implementation-defined order (x86-simd-sort fully sorts small arrays on AVX-512 hosts,
 * introselect elsewhere), and ties between equal metrics are broken arbitrarily.  This restatement fixes
 * the order every AVX-512 numpy >= 2.0 produces on tie-free inputs: kept candidates ascending by metric,
 * ties by candidate index; pickLeastReliableIndices returns ascending (score, index).  The golden vectors
 * (tests/golden/scl.npz, continuous channels, tie-free) pin exactly this behaviour against the live reference.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int q, N, k, L, use_log;
    const uint8_t *frozen;
    const int64_t *frozenValues; /* [N-k] */
    int fv_pos;
    const int64_t *actualInfo;   /* [k] (mandatory in the reference, :484-487) */
    double *prob;                /* prob_list */
    int nprob;
    double actual_prob;
    int64_t *infoList;           /* [L*q][k] */
    int failed;
} lctx;

typedef struct {
    int n;          /* list size */
    int seg;
    int64_t *enc;   /* [n][seg] */
    int *omap;      /* [n] */
    int64_t *aenc;  /* [seg] actual encoded vector */
} lres;

static void *xmalloc(size_t b) {
    void *p = malloc(b ? b : 1);
    if (!p) abort();
    return p;
}

static void free_res(lres *r) {
    free(r->enc);
    free(r->omap);
    free(r->aenc);
}

/* polarTransformOfQudits, QaryPolarEncoderDecoder.py:1136-1154 */
static void pt_qudits(int q, const int64_t *x, int len, int64_t *u) {
    if (len == 1) {
        u[0] = x[0];
        return;
    }
    int half = len / 2;
    int64_t *v = xmalloc(sizeof(int64_t) * len);
    for (int i = 0; i < half; ++i) {
        v[i] = (x[2 * i] + x[2 * i + 1]) % q;
        v[half + i] = (q - x[2 * i + 1]) % q;
    }
    pt_qudits(q, v, half, u);
    pt_qudits(q, v + half, half, u + half);
    free(v);
}

static void q_minus_l(int q, const double *p, int len, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int s = 0; s < q; ++s) d[s] = 0.0;
        for (int x1 = 0; x1 < q; ++x1)
            for (int x2 = 0; x2 < q; ++x2) d[(x1 + x2) % q] += a[x1] * b[x2];
    }
}
static void q_plus_l(int q, const double *p, int len, const int64_t *u, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int u2 = 0; u2 < q; ++u2) d[u2] = 0.0 + a[(u[h] + u2) % q] * b[(q - u2) % q];
    }
}
static void q_norm_l(int q, double *p, int len) {
    for (int i = 0; i < len; ++i) {
        double *row = p + (size_t)i * q, t = 0;
        for (int x = 0; x < q; ++x) t += row[x];
        if (t != 0)
            for (int x = 0; x < q; ++x) row[x] /= t;
    }
}

/* ---- log domain (use_log=True): the `if self.use_log` branches of listDecode / recursiveListDecode and of
 * QaryMemorylessVectorDistribution; logaddexp / logsumexp as restated in polar_oracle.c ---- */
double po_logaddexp(double x, double y);
double po_logsumexp(const double *a, int n);
static void q_minus_ll(int q, const double *p, int len, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int s = 0; s < q; ++s) d[s] = -INFINITY;
        for (int x1 = 0; x1 < q; ++x1)
            for (int x2 = 0; x2 < q; ++x2) d[(x1 + x2) % q] = po_logaddexp(d[(x1 + x2) % q], a[x1] + b[x2]);
    }
}
static void q_plus_ll(int q, const double *p, int len, const int64_t *u, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int u2 = 0; u2 < q; ++u2) d[u2] = po_logaddexp(-INFINITY, a[(u[h] + u2) % q] + b[(q - u2) % q]);
    }
}
static void q_norm_ll(int q, double *p, int len) {
    for (int i = 0; i < len; ++i) {
        double *row = p + (size_t)i * q, t = po_logsumexp(row, q);
        if (t != -INFINITY)
            for (int x = 0; x < q; ++x) row[x] -= t;
    }
}
/* np.sum of a float64 vector: numpy's pairwise summation (loops_utils.h.src), checked against np.sum for n = 1 .. 2048 */
static double np_sum(const double *a, ptrdiff_t n) {
    if (n < 8) {
        double r = 0.;
        for (ptrdiff_t i = 0; i < n; ++i) r += a[i];
        return r;
    } else if (n <= 128) {
        double r[8];
        ptrdiff_t i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    ptrdiff_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_sum(a, n2) + np_sum(a + n2, n - n2);
}

/* np.product([probs[j, cw[j]] for j in range(seg)]): sequential product in index order; log domain: np.sum of the list */
static double path_product(const lctx *c, const double *probs, const int64_t *cw, int seg) {
    const int q = c->q;
    if (c->use_log) {
        double *t = xmalloc(sizeof(double) * seg);
        for (int j = 0; j < seg; ++j) t[j] = probs[(size_t)j * q + cw[j]];
        double r = np_sum(t, seg);
        free(t);
        return r;
    }
    double r = probs[cw[0]];
    for (int j = 1; j < seg; ++j) r = r * probs[(size_t)j * q + cw[j]];
    return r;
}
static double comb(const lctx *c, double a, double b) { return c->use_log ? a + b : a * b; }

/* normalize, :867-872: divide by the maximum */
static double normalize_probs(const lctx *c, double *p, int n) {
    double mx = p[0];
    for (int i = 1; i < n; ++i)
        if (p[i] > mx) mx = p[i];
    for (int i = 0; i < n; ++i) p[i] = c->use_log ? p[i] - mx : p[i] / mx;
    return mx;
}

/* prune rule (:446-451 etc.): keep = min(#nonzero, L) largest; ascending (metric, index) order */
static int prune(const lctx *lc, const double *m, int C, int L, int *keep) {
    int nz = 0;
    for (int c = 0; c < C; ++c) nz += lc->use_log ? !(isinf(m[c]) && m[c] < 0) : (m[c] != 0); /* np.isneginf / np.count_nonzero */
    int ns = nz < L ? nz : L;
    /* full ascending argsort by (m, index): C is at most L*q^3 */
    int *ord = xmalloc(sizeof(int) * C);
    for (int c = 0; c < C; ++c) ord[c] = c;
    for (int a = 1; a < C; ++a) { /* insertion sort: stable */
        int v = ord[a], b = a - 1;
        while (b >= 0 && m[ord[b]] > m[v]) {
            ord[b + 1] = ord[b];
            --b;
        }
        ord[b + 1] = v;
    }
    for (int t = 0; t < ns; ++t) keep[t] = ord[C - ns + t];
    free(ord);
    return ns;
}

/* reliability (linear), :763-768: second largest / largest */
static double reliability(int q, int use_log, const double *row) {
    double m1 = use_log ? -INFINITY : -1, m2 = m1; /* m1 largest, m2 second largest (with multiplicity) */
    for (int x = 0; x < q; ++x) {
        double v = row[x];
        if (v > m1) {
            m2 = m1;
            m1 = v;
        } else if (v > m2) {
            m2 = v;
        }
    }
    return use_log ? m2 - m1 : m2 / m1;
}

/* pickLeastReliableIndices, :759-761: the `num` largest scores, ascending (score, index) */
static void pick_least_reliable(int q, int use_log, const double *probs, int seg, int num, int *out) {
    double *sc = xmalloc(sizeof(double) * seg);
    int *ord = xmalloc(sizeof(int) * seg);
    for (int j = 0; j < seg; ++j) {
        sc[j] = reliability(q, use_log, probs + (size_t)j * q);
        ord[j] = j;
    }
    for (int a = 1; a < seg; ++a) {
        int v = ord[a], b = a - 1;
        while (b >= 0 && sc[ord[b]] > sc[v]) {
            ord[b + 1] = ord[b];
            --b;
        }
        ord[b + 1] = v;
    }
    for (int t = 0; t < num; ++t) out[t] = ord[seg - num + t];
    free(sc);
    free(ord);
}

static int argmax_row(int q, const double *row) {
    int b = 0;
    for (int x = 1; x < q; ++x)
        if (row[x] > row[b]) b = x;
    return b;
}
static double max_row(int q, const double *row) {
    double m = row[0];
    for (int x = 1; x < q; ++x)
        if (row[x] > m) m = row[x];
    return m;
}

static int ipow(int b, int e) {
    int r = 1;
    while (e--) r *= b;
    return r;
}

/* recursiveListDecode, :403-757.  xy: [inList][seg][q]; axy: [seg][q] (actual path). */
static lres rec_list(lctx *c, int uIndex, int *infoIdx, const double *xy, int inList, const double *axy, int seg) {
    const int q = c->q, k = c->k, L = c->L;
    lres R;
    memset(&R, 0, sizeof R);
    R.seg = seg;
    int numInfo = 0;
    for (int j = 0; j < seg; ++j) numInfo += !c->frozen[uIndex + j];
    const size_t vsz = (size_t)seg * q;

    if (seg == 1 || c->failed) { /* leaves are only reachable for N == 1; not restated */
        c->failed = 1;
        return R;
    }
    /* ---------------- Rate-0, :495-518 ---------------- */
    if (numInfo == 0) {
        int64_t *fv = xmalloc(sizeof(int64_t) * seg), *ev = xmalloc(sizeof(int64_t) * seg);
        for (int j = 0; j < seg; ++j) fv[j] = c->frozenValues[c->fv_pos++];
        pt_qudits(q, fv, seg, ev);
        for (int i = 0; i < inList; ++i) c->prob[i] = comb(c, c->prob[i], path_product(c, xy + i * vsz, ev, seg));
        double nw = normalize_probs(c, c->prob, inList);
        if (c->use_log)
            c->actual_prob += path_product(c, axy, ev, seg) - nw;
        else
            c->actual_prob *= path_product(c, axy, ev, seg) / nw;
        R.n = inList;
        R.enc = xmalloc(sizeof(int64_t) * inList * seg);
        R.omap = xmalloc(sizeof(int) * inList);
        for (int i = 0; i < inList; ++i) {
            memcpy(R.enc + (size_t)i * seg, ev, sizeof(int64_t) * seg);
            R.omap[i] = i;
        }
        R.aenc = ev;
        free(fv);
        return R;
    }
    /* ---------------- Rep, :521-578 ---------------- */
    if (numInfo == 1) {
        int kpos = 0;
        while (c->frozen[uIndex + kpos]) ++kpos;
        int64_t *in = xmalloc(sizeof(int64_t) * q * seg), *es = xmalloc(sizeof(int64_t) * q * seg);
        for (int j = 0; j < seg; ++j) {
            if (j != kpos) {
                int64_t v = c->frozenValues[c->fv_pos++];
                for (int s = 0; s < q; ++s) in[(size_t)s * seg + j] = v;
            } else {
                for (int s = 0; s < q; ++s) in[(size_t)s * seg + j] = s;
            }
        }
        for (int s = 0; s < q; ++s) pt_qudits(q, in + (size_t)s * seg, seg, es + (size_t)s * seg);
        int C = inList * q;
        double *np_ = xmalloc(sizeof(double) * C);
        for (int i = 0; i < inList; ++i)
            for (int s = 0; s < q; ++s) {
                if (s > 0) memcpy(c->infoList + (size_t)(s * inList + i) * k, c->infoList + (size_t)i * k, sizeof(int64_t) * k);
                c->infoList[(size_t)(s * inList + i) * k + *infoIdx] = s;
                np_[s * inList + i] = comb(c, c->prob[i], path_product(c, xy + i * vsz, es + (size_t)s * seg, seg));
            }
        int ns = C;
        int *keep = xmalloc(sizeof(int) * C);
        if (C > L) {
            ns = prune(c, np_, C, L, keep);
            int64_t *tmp = xmalloc(sizeof(int64_t) * (size_t)ns * k);
            for (int t = 0; t < ns; ++t) memcpy(tmp + (size_t)t * k, c->infoList + (size_t)keep[t] * k, sizeof(int64_t) * k);
            memcpy(c->infoList, tmp, sizeof(int64_t) * (size_t)ns * k);
            for (size_t z = (size_t)ns * k; z < (size_t)L * q * k; ++z) c->infoList[z] = -1;
            free(tmp);
        } else {
            for (int t = 0; t < C; ++t) keep[t] = t;
        }
        R.n = ns;
        R.enc = xmalloc(sizeof(int64_t) * ns * seg);
        R.omap = xmalloc(sizeof(int) * ns);
        for (int t = 0; t < ns; ++t) {
            memcpy(R.enc + (size_t)t * seg, es + (size_t)(keep[t] / inList) * seg, sizeof(int64_t) * seg);
            R.omap[t] = keep[t] % inList;
            c->prob[t] = np_[keep[t]];
        }
        double nw = normalize_probs(c, c->prob, ns);
        R.aenc = xmalloc(sizeof(int64_t) * seg);
        memcpy(R.aenc, es + (size_t)c->actualInfo[*infoIdx] * seg, sizeof(int64_t) * seg);
        if (c->use_log)
            c->actual_prob += path_product(c, axy, R.aenc, seg) - nw;
        else
            c->actual_prob *= path_product(c, axy, R.aenc, seg) / nw;
        *infoIdx += 1;
        free(in), free(es), free(np_), free(keep);
        return R;
    }
    /* ---------------- Rate-1 (:581-628) and SPC (:631-682) ---------------- */
    if (numInfo == seg || numInfo == seg - 1) {
        const int spc = numInfo == seg - 1;
        const int nfork = spc ? 3 : 2, npick = spc ? 4 : 2;
        const int fs = ipow(q, nfork);
        const int64_t frozenValue = spc ? c->frozenValues[c->fv_pos] : 0;
        int C = inList * fs;
        int64_t *ev = xmalloc(sizeof(int64_t) * (size_t)C * seg);
        double *np_ = xmalloc(sizeof(double) * C);
        for (int i = 0; i < inList; ++i) {
            const double *pr = xy + i * vsz;
            int pick[4];
            pick_least_reliable(q, c->use_log, pr, seg, npick, pick);
            /* constant positions: argmax; base_prob = cur * prod(max) in index order (:785-788, :814-817) */
            uint8_t *isf = xmalloc(seg);
            memset(isf, 0, seg);
            for (int t = 0; t < npick; ++t) isf[pick[t]] = 1;
            int64_t *base = xmalloc(sizeof(int64_t) * seg);
            double prodmax = c->use_log ? 0.0 : 1.0; /* log: builtin sum() starts from 0 */
            int first = c->use_log ? 0 : 1;
            int64_t sumconst = 0;
            for (int j = 0; j < seg; ++j)
                if (!isf[j]) {
                    base[j] = argmax_row(q, pr + (size_t)j * q);
                    sumconst += base[j];
                    double mv = max_row(q, pr + (size_t)j * q);
                    prodmax = first ? mv : comb(c, prodmax, mv);
                    first = 0;
                }
            double base_prob = comb(c, c->prob[i], prodmax); /* np.product([]) == 1.0, sum([]) == 0 */
            int64_t delta = ((frozenValue - sumconst) % q + q) % q;
            for (int f = 0; f < fs; ++f) {
                int64_t *row = ev + (size_t)(i * fs + f) * seg;
                for (int j = 0; j < seg; ++j) row[j] = isf[j] ? 0 : base[j];
                int digits[3], rem = f;
                for (int t = nfork - 1; t >= 0; --t) { /* itertools.product: first index outermost */
                    digits[t] = rem % q;
                    rem /= q;
                }
                int64_t sf = 0;
                for (int t = 0; t < nfork; ++t) {
                    row[pick[t]] = digits[t];
                    sf += digits[t];
                }
                double pr_f;
                if (spc) {
                    row[pick[3]] = ((delta - sf) % q + q) % q;
                    pr_f = pr[(size_t)pick[0] * q + row[pick[0]]];
                    if (c->use_log) pr_f = 0. + pr_f; /* np.sum(axis=1) */
                    for (int t = 1; t < 4; ++t) pr_f = comb(c, pr_f, pr[(size_t)pick[t] * q + row[pick[t]]]);
                } else {
                    pr_f = comb(c, c->use_log ? 0. + pr[(size_t)pick[0] * q + digits[0]] : pr[(size_t)pick[0] * q + digits[0]],
                                pr[(size_t)pick[1] * q + digits[1]]);
                }
                np_[i * fs + f] = comb(c, pr_f, base_prob);
            }
            free(isf), free(base);
        }
        if (spc) c->fv_pos++;
        int ns = C;
        int *keep = xmalloc(sizeof(int) * C);
        if (C > L)
            ns = prune(c, np_, C, L, keep);
        else
            for (int t = 0; t < C; ++t) keep[t] = t;
        R.n = ns;
        R.enc = xmalloc(sizeof(int64_t) * ns * seg);
        R.omap = xmalloc(sizeof(int) * ns);
        for (int t = 0; t < ns; ++t) {
            memcpy(R.enc + (size_t)t * seg, ev + (size_t)keep[t] * seg, sizeof(int64_t) * seg);
            R.omap[t] = keep[t] / fs;
            c->prob[t] = np_[keep[t]];
        }
        double nw = normalize_probs(c, c->prob, ns);
        /* actual path */
        int64_t *au = xmalloc(sizeof(int64_t) * seg);
        if (spc) {
            au[0] = frozenValue;
            for (int j = 1; j < seg; ++j) au[j] = c->actualInfo[*infoIdx + j - 1];
        } else {
            for (int j = 0; j < seg; ++j) au[j] = c->actualInfo[*infoIdx + j];
        }
        R.aenc = xmalloc(sizeof(int64_t) * seg);
        pt_qudits(q, au, seg, R.aenc);
        if (c->use_log)
            c->actual_prob += path_product(c, axy, R.aenc, seg) - nw;
        else
            c->actual_prob *= path_product(c, axy, R.aenc, seg) / nw;
        /* informationList update, :618-620 / :672-674 */
        int64_t *tmp = xmalloc(sizeof(int64_t) * (size_t)ns * k), *uu = xmalloc(sizeof(int64_t) * seg);
        for (int t = 0; t < ns; ++t) memcpy(tmp + (size_t)t * k, c->infoList + (size_t)R.omap[t] * k, sizeof(int64_t) * k);
        memcpy(c->infoList, tmp, sizeof(int64_t) * (size_t)ns * k);
        for (int t = 0; t < ns; ++t) {
            pt_qudits(q, R.enc + (size_t)t * seg, seg, uu);
            for (int j = 0; j < numInfo; ++j) c->infoList[(size_t)t * k + *infoIdx + j] = uu[j + spc];
        }
        for (size_t z = (size_t)ns * k; z < (size_t)L * q * k; ++z) c->infoList[z] = -1;
        *infoIdx += numInfo;
        free(tmp), free(uu), free(au), free(ev), free(np_), free(keep);
        return R;
    }
    /* ---------------- general node, :684-757 ---------------- */
    const int half = seg / 2;
    const size_t hsz = (size_t)half * q;
    double *mv = xmalloc(sizeof(double) * (size_t)L * hsz);
    double *amv = xmalloc(sizeof(double) * hsz);
    void (*qm)(int, const double *, int, double *) = c->use_log ? q_minus_ll : q_minus_l;
    void (*qp)(int, const double *, int, const int64_t *, double *) = c->use_log ? q_plus_ll : q_plus_l;
    void (*qn)(int, double *, int) = c->use_log ? q_norm_ll : q_norm_l;
    for (int i = 0; i < inList; ++i) {
        qm(q, xy + i * vsz, seg, mv + i * hsz);
        qn(q, mv + i * hsz, half);
    }
    qm(q, axy, seg, amv);
    qn(q, amv, half);
    lres M = rec_list(c, uIndex, infoIdx, mv, inList, amv, half);
    if (c->failed) {
        free(mv), free(amv), free_res(&M);
        return R;
    }
    for (int i = 0; i < M.n; ++i) {
        qp(q, xy + (size_t)M.omap[i] * vsz, seg, M.enc + (size_t)i * half, mv + i * hsz);
        qn(q, mv + i * hsz, half);
    }
    qp(q, axy, seg, M.aenc, amv);
    qn(q, amv, half);
    lres P = rec_list(c, uIndex + half, infoIdx, mv, M.n, amv, half);
    free(mv), free(amv);
    if (c->failed) {
        free_res(&M), free_res(&P);
        return R;
    }
    R.n = P.n;
    R.enc = xmalloc(sizeof(int64_t) * P.n * seg);
    R.omap = xmalloc(sizeof(int) * P.n);
    R.aenc = xmalloc(sizeof(int64_t) * seg);
    for (int i = 0; i < P.n; ++i) {
        int mi = P.omap[i];
        for (int h = 0; h < half; ++h) {
            int64_t m = M.enc[(size_t)mi * half + h], p = P.enc[(size_t)i * half + h];
            R.enc[(size_t)i * seg + 2 * h] = (m + p) % q;
            R.enc[(size_t)i * seg + 2 * h + 1] = ((-p) % q + q) % q;
        }
        R.omap[i] = M.omap[mi];
    }
    for (int h = 0; h < half; ++h) {
        R.aenc[2 * h] = (M.aenc[h] + P.aenc[h]) % q;
        R.aenc[2 * h + 1] = ((-P.aenc[h]) % q + q) % q;
    }
    free_res(&M), free_res(&P);
    return R;
}

/*
 * listDecode, :118-227, with actualInformation given (genie selection, the only working caller is ir(), :856).
 * Outputs: info_out[k]; prob_result (ProbResult value 0..5); final list for set-parity checks:
 * list_size, list_info[L][k], list_prob[L], actual_prob_out.  Returns 0, or -1 on unsupported shapes.
 */
static int list_decode(int use_log, int q, int N, int L, const uint8_t *frozen, const double *xyprobs, const int64_t *frozenValues,
                       const int64_t *actualInfo, int64_t *info_out, int *prob_result, int *list_size, int64_t *list_info,
                       double *list_prob, double *actual_prob_out) {
    if (q < 2 || N < 2 || L < 1 || !actualInfo) return -1;
    lctx c;
    memset(&c, 0, sizeof c);
    c.q = q;
    c.use_log = use_log;
    c.N = N;
    c.L = L;
    c.frozen = frozen;
    c.frozenValues = frozenValues;
    c.actualInfo = actualInfo;
    int k = 0;
    for (int i = 0; i < N; ++i) k += !frozen[i];
    c.k = k;
    c.prob = xmalloc(sizeof(double) * (size_t)L * q * q * q + 64);
    c.prob[0] = use_log ? 0.0 : 1.0;
    c.nprob = 1;
    c.actual_prob = use_log ? 0.0 : 1.0;
    c.infoList = xmalloc(sizeof(int64_t) * (size_t)L * q * (k ? k : 1));
    for (size_t z = 0; z < (size_t)L * q * k; ++z) c.infoList[z] = -1;
    int infoIdx = 0;
    lres R = rec_list(&c, 0, &infoIdx, xyprobs, 1, xyprobs, N);
    if (c.failed) {
        free(c.prob), free(c.infoList), free_res(&R);
        return -1;
    }
    double maxp = c.prob[0], minp = c.prob[0];
    for (int i = 1; i < R.n; ++i) {
        if (c.prob[i] > maxp) maxp = c.prob[i];
        if (c.prob[i] < minp) minp = c.prob[i];
    }
    int found = -1;
    for (int i = 0; i < R.n && found < 0; ++i)
        if (memcmp(c.infoList + (size_t)i * k, actualInfo, sizeof(int64_t) * k) == 0) found = i;
    if (found >= 0) {
        memcpy(info_out, c.infoList + (size_t)found * k, sizeof(int64_t) * k);
        *prob_result = c.prob[found] == maxp ? 0 : 1;
    } else {
        memcpy(info_out, c.infoList, sizeof(int64_t) * k);
        *prob_result = c.actual_prob > maxp ? 2 : c.actual_prob == maxp ? 3 : c.actual_prob >= minp ? 4 : 5;
    }
    if (list_size) *list_size = R.n;
    if (list_info) memcpy(list_info, c.infoList, sizeof(int64_t) * (size_t)R.n * k);
    if (list_prob) memcpy(list_prob, c.prob, sizeof(double) * R.n);
    if (actual_prob_out) *actual_prob_out = c.actual_prob;
    free(c.prob), free(c.infoList), free_res(&R);
    return 0;
}

int po_list_decode(int q, int N, int L, const uint8_t *frozen, const double *xyprobs, const int64_t *frozenValues,
                   const int64_t *actualInfo, int64_t *info_out, int *prob_result, int *list_size, int64_t *list_info,
                   double *list_prob, double *actual_prob_out) {
    return list_decode(0, q, N, L, frozen, xyprobs, frozenValues, actualInfo, info_out, prob_result, list_size, list_info, list_prob,
                       actual_prob_out);
}
/* use_log=True: xyprobs hold natural logarithms; list metrics and actual_prob are log values */
int po_list_decode_log(int q, int N, int L, const uint8_t *frozen, const double *xyprobs, const int64_t *frozenValues,
                       const int64_t *actualInfo, int64_t *info_out, int *prob_result, int *list_size, int64_t *list_info,
                       double *list_prob, double *actual_prob_out) {
    return list_decode(1, q, N, L, frozen, xyprobs, frozenValues, actualInfo, info_out, prob_result, list_size, list_info, list_prob,
                       actual_prob_out);
}

int po_list_decode_batch(int B, int q, int N, int L, const uint8_t *frozen, const double *xyprobs,
                         const int64_t *frozenValues, const int64_t *actualInfo, int64_t *info_out, int *prob_result) {
    int k = 0;
    for (int i = 0; i < N; ++i) k += !frozen[i];
    for (int b = 0; b < B; ++b) {
        int rc = po_list_decode(q, N, L, frozen, xyprobs + (size_t)b * N * q, frozenValues + (size_t)b * (N - k),
                                actualInfo + (size_t)b * k, info_out + (size_t)b * k, prob_result + b, NULL, NULL, NULL,
                                NULL);
        if (rc) return rc;
    }
    return 0;
}
