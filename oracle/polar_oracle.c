/*
 * polar_oracle.c -- CPU restatement of the reference's polar encode / SC decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (polarcub_b200/) may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker or as the timed CPU baseline.
 *
 * It follows the reference (benjilieber/polarcub, pure Python, float64 probability domain) operation
 * by operation, so results are bit-identical to the Python when compiled WITHOUT floating-point
 * contraction (-ffp-contract=off: the Python rounds every * and + separately).
 * Parity is PINNED: tests/test_oracle_golden.py checks every function here against golden vectors
 * produced by running the live reference in the build container (oracle/gen_golden.py).
 *
 * Reference lines followed (file:line relative to the reference root):
 *   binary recursion      BinaryPolarEncoderDecoder.py:223-325
 *   binary f / g / norm   VectorDistributions/BinaryMemorylessVectorDistribution.py:15-87
 *   inverse transform     BinaryPolarEncoderDecoder.py:494-516
 *   q-ary recursion       QaryPolarEncoderDecoder.py:318-401
 *   q-ary f / g / norm    VectorDistributions/QaryMemorylessVectorDistribution.py:26-118
 *   q-ary inverse         QaryPolarEncoderDecoder.py:1136-1154
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* scratch arena: the recursion allocates one child vector per level, like the Python allocates objects */
typedef struct {
    char *base;
    size_t cap, top;
} arena_t;

static void *arena_get(arena_t *a, size_t bytes) {
    bytes = (bytes + 15u) & ~(size_t)15u;
    if (a->top + bytes > a->cap) return NULL;
    void *p = a->base + a->top;
    a->top += bytes;
    return p;
}

/* ------------------------------------------------------------------------------------------------ */
/* Binary memoryless vector distribution (probs[len][2], float64)                                     */

/* BinaryMemorylessVectorDistribution.py:15-29 */
static void bin_minus(const double *p, int len, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + 4 * h, *b = p + 4 * h + 2;
        o[2 * h + 0] = a[0] * b[0] + a[1] * b[1];
        o[2 * h + 1] = a[0] * b[1] + a[1] * b[0];
    }
}

/* BinaryMemorylessVectorDistribution.py:31-47 */
static void bin_plus(const double *p, int len, const int64_t *u, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + 4 * h, *b = p + 4 * h + 2;
        if (u[h] == 0) {
            o[2 * h + 0] = a[0] * b[0];
            o[2 * h + 1] = a[1] * b[1];
        } else {
            o[2 * h + 0] = a[1] * b[0];
            o[2 * h + 1] = a[0] * b[1];
        }
    }
}

/* calcNormalizationVector + normalize, BinaryMemorylessVectorDistribution.py:71-87 */
static void bin_normalize(double *p, int len) {
    for (int i = 0; i < len; ++i) {
        double t = p[2 * i] > p[2 * i + 1] ? p[2 * i] : p[2 * i + 1]; /* np.maximum on non-NaN input */
        if (t == 0) t = 1;
        p[2 * i] /= t;
        p[2 * i + 1] /= t;
    }
}

/* BinaryMemorylessVectorDistribution.py:52-69 */
static void bin_marginal(const double *p, double *m) {
    double s = 0.0;
    s += p[0];
    s += p[1];
    if (s > 0.0) {
        m[0] = p[0] / s;
        m[1] = p[1] / s;
    } else {
        m[0] = 0.5;
        m[1] = 0.5;
    }
}

typedef struct {
    const uint8_t *frozen;  /* [N] 1 = frozen */
    const double *r;        /* [N] randomlyGeneratedNumbers, BinaryPolarEncoderDecoder.py:33-44 */
    int64_t *info;          /* [k] read when encoding, written when decoding */
    double *marg;           /* optional [N][2]: marginalizedUProbs (genie capture, :268-273) */
    double *lvl1_minus;     /* optional [N/2][2]: first xy minus child after normalisation */
    double *lvl1_plus;      /* optional [N/2][2]: first xy plus child after normalisation */
    int uidx, iidx, top_len;
    arena_t ar;
    int oom;
} bin_ctx;

/* BinaryPolarEncoderDecoder.py:223-325.  x = prior tree, xy = posterior tree (NULL when encoding). */
static void bin_rec(bin_ctx *c, int len, const double *x, const double *xy, int64_t *enc) {
    if (c->oom) return;
    if (len == 1) {
        double m[2];
        if (!c->frozen[c->uidx]) {
            if (xy) {
                bin_marginal(xy, m);
                c->info[c->iidx] = (m[0] >= m[1]) ? 0 : 1; /* :252, tie -> 0 */
            }
            enc[0] = c->info[c->iidx];
            c->iidx += 1;
        } else {
            bin_marginal(x, m);
            enc[0] = (m[0] >= c->r[c->uidx]) ? 0 : 1; /* :259-262 */
        }
        if (c->marg) {
            bin_marginal(xy ? xy : x, m);
            c->marg[2 * c->uidx] = m[0];
            c->marg[2 * c->uidx + 1] = m[1];
        }
        c->uidx += 1;
        return;
    }
    int half = len / 2;
    size_t mark = c->ar.top;
    double *xc = arena_get(&c->ar, sizeof(double) * 2 * half);
    double *xyc = xy ? arena_get(&c->ar, sizeof(double) * 2 * half) : NULL;
    int64_t *em = arena_get(&c->ar, sizeof(int64_t) * half);
    int64_t *ep = arena_get(&c->ar, sizeof(int64_t) * half);
    if (!xc || (xy && !xyc) || !em || !ep) {
        c->oom = 1;
        return;
    }
    bin_minus(x, len, xc);
    bin_normalize(xc, half);
    if (xy) {
        bin_minus(xy, len, xyc);
        bin_normalize(xyc, half);
        if (len == c->top_len && c->lvl1_minus) memcpy(c->lvl1_minus, xyc, sizeof(double) * 2 * half);
    }
    bin_rec(c, half, xc, xyc, em);
    bin_plus(x, len, em, xc);
    bin_normalize(xc, half);
    if (xy) {
        bin_plus(xy, len, em, xyc);
        bin_normalize(xyc, half);
        if (len == c->top_len && c->lvl1_plus) memcpy(c->lvl1_plus, xyc, sizeof(double) * 2 * half);
    }
    bin_rec(c, half, xc, xyc, ep);
    for (int h = 0; h < half; ++h) { /* :321-323 */
        enc[2 * h] = (em[h] + ep[h]) % 2;
        enc[2 * h + 1] = ep[h];
    }
    c->ar.top = mark;
}

static int bin_run(int N, const uint8_t *frozen, const double *r, const double *x, const double *xy, int64_t *info,
                   int64_t *cw, double *marg, double *l1m, double *l1p) {
    bin_ctx c;
    memset(&c, 0, sizeof c);
    c.frozen = frozen;
    c.r = r;
    c.info = info;
    c.marg = marg;
    c.lvl1_minus = l1m;
    c.lvl1_plus = l1p;
    c.top_len = N;
    c.ar.cap = (size_t)N * 64 + 4096; /* sum over levels of (2*16 + 2*8) * len/2 < 48 N */
    c.ar.base = malloc(c.ar.cap);
    if (!c.ar.base) return -2;
    bin_rec(&c, N, x, xy, cw);
    free(c.ar.base);
    return c.oom ? -2 : 0;
}

/* BinaryPolarEncoderDecoder.encode, :46-69.  info[k] in, cw[N] out. */
int po_bin_encode(int N, const uint8_t *frozen, const double *r, const double *xprobs, const int64_t *info,
                  int64_t *cw, double *marg) {
    return bin_run(N, frozen, r, xprobs, NULL, (int64_t *)info, cw, marg, NULL, NULL);
}

/* BinaryPolarEncoderDecoder.decode, :71-99.  cw[N], info[k] out. */
int po_bin_decode(int N, const uint8_t *frozen, const double *r, const double *xprobs, const double *xyprobs,
                  int64_t *cw, int64_t *info, double *marg, double *lvl1_minus, double *lvl1_plus) {
    return bin_run(N, frozen, r, xprobs, xyprobs, info, cw, marg, lvl1_minus, lvl1_plus);
}

/* B frames, same code; xy is [B][N][2], x is [N][2] (shared prior, BinaryPolarEncoderDecoder.py:346) */
int po_bin_decode_batch(int B, int N, int k, const uint8_t *frozen, const double *r, const double *xprobs,
                        const double *xyprobs, int64_t *cw, int64_t *info) {
    for (int b = 0; b < B; ++b) {
        int rc = po_bin_decode(N, frozen, r, xprobs, xyprobs + (size_t)b * N * 2, cw + (size_t)b * N,
                               info + (size_t)b * k, NULL, NULL, NULL);
        if (rc) return rc;
    }
    return 0;
}

int po_bin_encode_batch(int B, int N, int k, const uint8_t *frozen, const double *r, const double *xprobs,
                        const int64_t *info, int64_t *cw) {
    for (int b = 0; b < B; ++b) {
        int rc = po_bin_encode(N, frozen, r, xprobs, info + (size_t)b * k, cw + (size_t)b * N, NULL);
        if (rc) return rc;
    }
    return 0;
}

/* polarTransformOfBits, BinaryPolarEncoderDecoder.py:494-516 (x -> u).  Works in place on a copy. */
static void bits_rec(const int64_t *x, int len, int64_t *u, int64_t *tmp) {
    if (len == 1) {
        u[0] = x[0];
        return;
    }
    int half = len / 2;
    int64_t *v1 = tmp, *v2 = tmp + half;
    for (int i = 0; i < half; ++i) {
        v1[i] = (x[2 * i] + x[2 * i + 1]) % 2;
        v2[i] = x[2 * i + 1];
    }
    bits_rec(v1, half, u, tmp + len);
    bits_rec(v2, half, u + half, tmp + len);
}

int po_polar_transform_bits(int N, const int64_t *x, int64_t *u) {
    int64_t *tmp = malloc(sizeof(int64_t) * 2 * (size_t)N + 64);
    if (!tmp) return -2;
    bits_rec(x, N, u, tmp);
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* q-ary memoryless vector distribution (probs[len][q], float64, linear domain)                        */

/* QaryMemorylessVectorDistribution.py:26-43 (linear branch): accumulation order x1 outer, x2 inner */
static void q_minus(int q, const double *p, int len, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int s = 0; s < q; ++s) d[s] = 0.0;
        for (int x1 = 0; x1 < q; ++x1)
            for (int x2 = 0; x2 < q; ++x2) {
                int u1 = (x1 + x2) % q;
                d[u1] += a[x1] * b[x2];
            }
    }
}

/* QaryMemorylessVectorDistribution.py:45-64 (linear branch) */
static void q_plus(int q, const double *p, int len, const int64_t *u, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int u2 = 0; u2 < q; ++u2) {
            int x1 = (int)((u[h] + u2) % q);
            int x2 = (q - u2) % q;
            d[u2] = 0.0 + a[x1] * b[x2];
        }
    }
}

/* normalize(None), QaryMemorylessVectorDistribution.py:92-118: divide by the row sum unless it is 0 */
static void q_normalize(int q, double *p, int len) {
    for (int i = 0; i < len; ++i) {
        double *row = p + (size_t)i * q;
        double t = 0;
        for (int x = 0; x < q; ++x) t += row[x]; /* builtin sum(): ((0 + p0) + p1) + ... */
        if (t != 0)
            for (int x = 0; x < q; ++x) row[x] /= t;
    }
}

/* QaryMemorylessVectorDistribution.py:69-90 */
static void q_marginal(int q, const double *p, double *m) {
    double s = 0;
    for (int x = 0; x < q; ++x) s += p[x];
    if (s > 0.0)
        for (int x = 0; x < q; ++x) m[x] = p[x] / s;
    else
        for (int x = 0; x < q; ++x) m[x] = 1.0 / q;
}

/* ---- log domain (use_log=True): numpy.logaddexp and scipy.special.logsumexp (scipy >= 1.15) restated ---- */
/* npy_logaddexp (numpy/_core/src/npymath/npy_math_internal.h.src) */
double po_logaddexp(double x, double y) {
    if (x == y) return x + 0.693147180559945309417232121458176568; /* NPY_LOGE2: also covers equal infinities */
    double tmp = x - y;
    if (tmp > 0) return x + log1p(exp(-tmp));
    if (tmp <= 0) return y + log1p(exp(tmp));
    return tmp; /* NaN */
}
/* scipy.special._logsumexp._logsumexp for a real vector, b=None: the maximal elements are taken out of the sum (m of them),
 * out = log1p(sum_{others} exp(a - a_max) / m) + log(m) + a_max; an infinite result falls back to log(sum(exp(a))) */
double po_logsumexp(const double *a, int n) {
    double mx = a[0];
    for (int i = 1; i < n; ++i)
        if (a[i] > mx) mx = a[i];
    double m = 0, s = 0;
    for (int i = 0; i < n; ++i) {
        if (a[i] == mx)
            m += 1.0;
        else
            s += exp(a[i] - mx);
    }
    if (s != 0) s = s / m;
    double out = log1p(s) + log(m) + mx;
    if (!isfinite(out)) {
        double t = 0;
        for (int i = 0; i < n; ++i) t += exp(a[i]);
        out = log(t);
    }
    return out;
}

/* QaryMemorylessVectorDistribution.py:26-43, log branch */
static void q_minus_log(int q, const double *p, int len, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int s = 0; s < q; ++s) d[s] = -INFINITY;
        for (int x1 = 0; x1 < q; ++x1)
            for (int x2 = 0; x2 < q; ++x2) {
                int u1 = (x1 + x2) % q;
                d[u1] = po_logaddexp(d[u1], a[x1] + b[x2]);
            }
    }
}
/* :45-64, log branch */
static void q_plus_log(int q, const double *p, int len, const int64_t *u, double *o) {
    int half = len / 2;
    for (int h = 0; h < half; ++h) {
        const double *a = p + (size_t)2 * h * q, *b = a + q;
        double *d = o + (size_t)h * q;
        for (int u2 = 0; u2 < q; ++u2) {
            int x1 = (int)((u[h] + u2) % q);
            int x2 = (q - u2) % q;
            d[u2] = po_logaddexp(-INFINITY, a[x1] + b[x2]);
        }
    }
}
/* :92-118, log branch */
static void q_normalize_log(int q, double *p, int len) {
    for (int i = 0; i < len; ++i) {
        double *row = p + (size_t)i * q;
        double t = po_logsumexp(row, q);
        if (t != -INFINITY)
            for (int x = 0; x < q; ++x) row[x] -= t;
    }
}
/* :69-90, log branch */
static void q_marginal_log(int q, const double *p, double *m) {
    double s = po_logsumexp(p, q);
    if (s > -INFINITY)
        for (int x = 0; x < q; ++x) m[x] = p[x] - s;
    else
        for (int x = 0; x < q; ++x) m[x] = -log((double)q);
}

typedef struct {
    int q, use_log;
    const uint8_t *frozen;
    int64_t *info;
    double *marg; /* optional [N][q] */
    double *lvl1_minus, *lvl1_plus;
    int uidx, iidx, top_len;
    arena_t ar;
    int oom;
} q_ctx;

/* QaryPolarEncoderDecoder.py:318-401 */
static void q_rec(q_ctx *c, int len, const double *x, const double *xy, int64_t *enc) {
    int q = c->q;
    if (c->oom) return;
    if (len == 1) {
        double m[64];
        if (!c->frozen[c->uidx]) {
            if (xy) {
                (c->use_log ? q_marginal_log : q_marginal)(q, xy, m);
                int best = 0; /* np.argmax: first maximum */
                for (int s = 1; s < q; ++s)
                    if (m[s] > m[best]) best = s;
                c->info[c->iidx] = best;
            }
            enc[0] = c->info[c->iidx];
            c->iidx += 1;
        } else {
            enc[0] = 0; /* :351 */
        }
        if (c->marg) {
            (c->use_log ? q_marginal_log : q_marginal)(q, xy ? xy : x, m);
            memcpy(c->marg + (size_t)c->uidx * q, m, sizeof(double) * q);
        }
        c->uidx += 1;
        return;
    }
    int half = len / 2;
    size_t mark = c->ar.top;
    double *xc = arena_get(&c->ar, sizeof(double) * q * half);
    double *xyc = xy ? arena_get(&c->ar, sizeof(double) * q * half) : NULL;
    int64_t *em = arena_get(&c->ar, sizeof(int64_t) * half);
    int64_t *ep = arena_get(&c->ar, sizeof(int64_t) * half);
    if (!xc || (xy && !xyc) || !em || !ep) {
        c->oom = 1;
        return;
    }
    void (*qm)(int, const double *, int, double *) = c->use_log ? q_minus_log : q_minus;
    void (*qp)(int, const double *, int, const int64_t *, double *) = c->use_log ? q_plus_log : q_plus;
    void (*qn)(int, double *, int) = c->use_log ? q_normalize_log : q_normalize;
    qm(q, x, len, xc);
    qn(q, xc, half);
    if (xy) {
        qm(q, xy, len, xyc);
        qn(q, xyc, half);
        if (len == c->top_len && c->lvl1_minus) memcpy(c->lvl1_minus, xyc, sizeof(double) * q * half);
    }
    q_rec(c, half, xc, xyc, em);
    qp(q, x, len, em, xc);
    qn(q, xc, half);
    if (xy) {
        qp(q, xy, len, em, xyc);
        qn(q, xyc, half);
        if (len == c->top_len && c->lvl1_plus) memcpy(c->lvl1_plus, xyc, sizeof(double) * q * half);
    }
    q_rec(c, half, xc, xyc, ep);
    for (int h = 0; h < half; ++h) { /* :397-399 */
        enc[2 * h] = (em[h] + ep[h]) % q;
        enc[2 * h + 1] = (-ep[h] + q) % q;
    }
    c->ar.top = mark;
}

static int q_run(int q, int use_log, int N, const uint8_t *frozen, const double *x, const double *xy, int64_t *info, int64_t *cw,
                 double *marg, double *l1m, double *l1p) {
    if (q < 2 || q > 64) return -1;
    q_ctx c;
    memset(&c, 0, sizeof c);
    c.q = q;
    c.use_log = use_log;
    c.frozen = frozen;
    c.info = info;
    c.marg = marg;
    c.lvl1_minus = l1m;
    c.lvl1_plus = l1p;
    c.top_len = N;
    c.ar.cap = (size_t)N * (16 * q + 32) + 4096;
    c.ar.base = malloc(c.ar.cap);
    if (!c.ar.base) return -2;
    q_rec(&c, N, x, xy, cw);
    free(c.ar.base);
    return c.oom ? -2 : 0;
}

/* QaryPolarEncoderDecoder.encode, :65-88 */
int po_q_encode(int q, int N, const uint8_t *frozen, const double *xprobs, const int64_t *info, int64_t *cw) {
    return q_run(q, 0, N, frozen, xprobs, NULL, (int64_t *)info, cw, NULL, NULL, NULL);
}

/* QaryPolarEncoderDecoder.decode, :90-116 (the reference returns only `information`; cw is extra) */
int po_q_decode(int q, int N, const uint8_t *frozen, const double *xprobs, const double *xyprobs, int64_t *cw,
                int64_t *info, double *marg, double *lvl1_minus, double *lvl1_plus) {
    return q_run(q, 0, N, frozen, xprobs, xyprobs, info, cw, marg, lvl1_minus, lvl1_plus);
}

/* the same with use_log=True: xprobs / xyprobs hold natural logarithms (-inf for 0) */
int po_q_decode_log(int q, int N, const uint8_t *frozen, const double *xprobs, const double *xyprobs, int64_t *cw,
                    int64_t *info, double *marg, double *lvl1_minus, double *lvl1_plus) {
    return q_run(q, 1, N, frozen, xprobs, xyprobs, info, cw, marg, lvl1_minus, lvl1_plus);
}

int po_q_decode_batch(int B, int q, int N, int k, const uint8_t *frozen, const double *xprobs, const double *xyprobs,
                      int64_t *cw, int64_t *info) {
    for (int b = 0; b < B; ++b) {
        int rc = po_q_decode(q, N, frozen, xprobs, xyprobs + (size_t)b * N * q, cw + (size_t)b * N,
                             info + (size_t)b * k, NULL, NULL, NULL);
        if (rc) return rc;
    }
    return 0;
}

/* polarTransformOfQudits, QaryPolarEncoderDecoder.py:1136-1154 */
static void qudits_rec(int q, const int64_t *x, int len, int64_t *u, int64_t *tmp) {
    if (len == 1) {
        u[0] = x[0];
        return;
    }
    int half = len / 2;
    int64_t *v1 = tmp, *v2 = tmp + half;
    for (int i = 0; i < half; ++i) {
        v1[i] = (x[2 * i] + x[2 * i + 1]) % q;
        v2[i] = (q - x[2 * i + 1]) % q;
    }
    qudits_rec(q, v1, half, u, tmp + len);
    qudits_rec(q, v2, half, u + half, tmp + len);
}

int po_polar_transform_qudits(int q, int N, const int64_t *x, int64_t *u) {
    int64_t *tmp = malloc(sizeof(int64_t) * 2 * (size_t)N + 64);
    if (!tmp) return -2;
    qudits_rec(q, x, N, u, tmp);
    free(tmp);
    return 0;
}
