// tv_construct.cu -- HOST code: the Tal-Vardy degrading construction for binary-input memoryless channels with a uniform
// input, i.e. the Pe vector of calcFrozenSet_degradingUpgrading(n, L, eps, xDistribution=None, xyDistribution)
// (ScalarDistributions/BinaryMemorylessDistribution.py:620-680) -- the code-construction step in front of the encode /
// decode path (634 s in the reference at N = 1024, L = 100).
//
// Every step follows the reference operation by operation so that the float64 results are the reference's bit for bit:
//   minusTransform / plusTransform        BinaryMemorylessDistribution.py:259-285 (double loops, products rounded one by one)
//   mergeEquivalentSymbols                :168-208  removeZeroProbOutput :93-110, sortProbs :112-166 (two STABLE sorts on the
//                                         keys p1/(p0+p1) and -p0/(p0+p1)), math.isclose merge, normalize :89-91 -- whose
//                                         `sum(sum(probs, []))` is CPython >= 3.12's COMPENSATED (Neumaier) float sum
//   degrade                               :287-345 with _calcKey_degrade :499-511, hxgiveny / eta :452-474 (glibc log2) and
//                                         the array heap + linked list of UpgradingDegrading/LinkedListHeap.py:4-175, whose
//                                         sift-up swaps on EQUAL keys and whose sift-down prefers the left child on ties
//   errorProb                             :39-45
// No device code in this file; nvcc only forwards it to the host compiler (no FMA contraction on x86-64 without -march).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

// one output symbol of a binary channel; head / tail: the chain (through a shared `next` array) of the original output
// letters merged into it -- the reference's `auxiliary` sets, used only by the q-ary construction (-1: not tracked)
struct Pair {
    double p0, p1;
    int head = -1, tail = -1;
    Pair() : p0(0.0), p1(0.0) {}
    Pair(double a, double b) : p0(a), p1(b) {}
};
// aux union: a |= b
inline void aux_join(Pair &a, const Pair &b, std::vector<int> *next) {
    if (!next || b.head < 0) return;
    if (a.head < 0) {
        a.head = b.head;
    } else {
        (*next)[a.tail] = b.head;
    }
    a.tail = b.tail;
}

// builtin sum() of a list of exact Python floats with the default integer start (Python/bltinmodule.c, 3.12+): 0 + v[0],
// then Neumaier's compensated additions, the compensation added once at the end
double py_sum(const std::vector<double> &v) {
    if (v.empty()) return 0.0;
    double f = 0.0 + v[0], c = 0.0;
    for (size_t i = 1; i < v.size(); ++i) {
        const double x = v[i], t = f + x;
        if (std::fabs(f) >= std::fabs(x))
            c += (f - t) + x;
        else
            c += (x - t) + f;
        f = t;
    }
    if (c != 0.0 && std::isfinite(c)) f += c;
    return f;
}

// math.isclose(a, b) with the default rel_tol = 1e-09, abs_tol = 0.0 (Modules/mathmodule.c)
bool py_isclose(double a, double b) {
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b)) return false;
    const double diff = std::fabs(b - a), rel = 1e-09;
    return (diff <= std::fabs(rel * b)) || (diff <= std::fabs(rel * a)) || (diff <= 0.0);
}

double eta(double p) {  // :452-460
    p = std::min(1.0, p);
    return p == 0.0 ? 0.0 : -p * std::log2(p);
}
double hxgiveny(double d0, double d1) {  // :471-473
    const double py = d0 + d1;
    return py * (eta(d0 / py) + eta(d1 / py));
}
// :499-511: hxgiveny(merged) - hxgiveny(left) - hxgiveny(center); hl / hc are the two symbols' own (cached) terms -- the same
// function of the same operands, so caching changes no bit
double key_degrade(const Pair &l, double hl, const Pair &c, double hc) {
    return hxgiveny(l.p0 + c.p0, l.p1 + c.p1) - hl - hc;
}

void merge_equivalent_symbols(std::vector<Pair> &probs, std::vector<int> *next) {
    // removeZeroProbOutput
    std::vector<Pair> kept;
    kept.reserve(probs.size());
    for (const Pair &p : probs)
        if (p.p0 + p.p1 > 0.0) kept.push_back(p);
    // sortProbs: ascending p(x=0|y) split in two halves for numerical stability, each a stable sort on its key
    struct Keyed {
        double key;
        Pair p;
    };
    std::vector<Keyed> zero_more, one_more;
    for (const Pair &p : kept) {
        const double s = p.p0 + p.p1;
        if (p.p0 / s > 0.5)
            zero_more.push_back({p.p1 / s, p});
        else
            one_more.push_back({-p.p0 / s, p});
    }
    auto by_key = [](const Keyed &a, const Keyed &b) { return a.key < b.key; };
    std::stable_sort(zero_more.begin(), zero_more.end(), by_key);
    std::stable_sort(one_more.begin(), one_more.end(), by_key);
    std::vector<Pair> sorted;
    sorted.reserve(kept.size());
    for (const Keyed &k : zero_more) sorted.push_back(k.p);
    for (const Keyed &k : one_more) sorted.push_back(k.p);
    // merge symbols whose normalised pairs are close to the running merged symbol
    std::vector<Pair> merged;
    if (!sorted.empty()) merged.push_back(sorted[0]);
    for (size_t i = 1; i < sorted.size(); ++i) {
        const Pair &p = sorted[i];
        Pair &prev = merged.back();
        const double s = p.p0 + p.p1, sp = prev.p0 + prev.p1;
        const bool close = py_isclose(p.p0 / s, prev.p0 / sp) && py_isclose(p.p1 / s, prev.p1 / sp);
        if (!close) {
            merged.push_back(p);
        } else {
            prev.p0 += p.p0;
            prev.p1 += p.p1;
            aux_join(prev, p, next);
        }
    }
    // normalize
    std::vector<double> flat;
    flat.reserve(2 * merged.size());
    for (const Pair &p : merged) {
        flat.push_back(p.p0);
        flat.push_back(p.p1);
    }
    const double total = py_sum(flat);
    for (Pair &p : merged) {
        p.p0 = p.p0 / total;
        p.p1 = p.p1 / total;
    }
    probs.swap(merged);
}

// LinkedListHeap.py: a binary min-heap in an array whose elements are also a doubly linked list in symbol order
struct Heap {
    struct El {
        double key;
        Pair d;
        int left, right, at;
    };
    std::vector<El> el;
    std::vector<int> arr;  // heap array of element ids

    void swap_els(int a, int b) {
        std::swap(el[a].at, el[b].at);
        arr[el[a].at] = a;
        arr[el[b].at] = b;
    }
    void up(int e) {
        for (;;) {
            const int pi = (el[e].at + 1) / 2 - 1;
            if (pi == -1) break;
            const int par = arr[pi];
            if (el[par].key < el[e].key) break;  // equal keys DO swap
            swap_els(e, par);
        }
    }
    void down(int e) {
        for (;;) {
            const int li = 2 * (el[e].at + 1) - 1, ri = li + 1;
            double mk = el[e].key;
            int mc = -1;
            if (li < (int)arr.size() && el[arr[li]].key < mk) {
                mk = el[arr[li]].key;
                mc = arr[li];
            }
            if (ri < (int)arr.size() && el[arr[ri]].key < mk) {
                mk = el[arr[ri]].key;
                mc = arr[ri];
            }
            if (mc < 0) break;
            swap_els(e, mc);
        }
    }
    void insert_at_tail(double key, const Pair &d) {
        const int id = (int)el.size();
        el.push_back({key, d, id - 1, -1, (int)arr.size()});
        if (id > 0) el[id - 1].right = id;
        arr.push_back(id);
        up(id);
    }
    int extract_min() {
        const int e = arr[0];
        const int l = el[e].left, r = el[e].right;
        if (l >= 0) el[l].right = r;
        if (r >= 0) el[r].left = l;
        const int last = arr.back();
        arr.pop_back();
        arr[0] = last;
        el[last].at = 0;
        down(last);
        return e;
    }
    void update_key(int e, double k) {
        const double old = el[e].key;
        el[e].key = k;
        if (old < k)
            down(e);
        else if (old > k)
            up(e);
    }
};

void degrade(std::vector<Pair> &probs, int L, std::vector<int> *next = nullptr) {  // :287-345
    merge_equivalent_symbols(probs, next);
    const double inf = std::numeric_limits<double>::infinity();
    Heap h;
    h.el.reserve(probs.size());
    h.arr.reserve(probs.size());
    std::vector<double> keys(probs.size()), hs(probs.size());
    for (size_t i = 0; i < probs.size(); ++i) hs[i] = hxgiveny(probs[i].p0, probs[i].p1);
    for (size_t i = 0; i < probs.size(); ++i) keys[i] = i == 0 ? inf : key_degrade(probs[i - 1], hs[i - 1], probs[i], hs[i]);
    for (size_t i = 0; i < probs.size(); ++i) h.insert_at_tail(keys[i], probs[i]);
    while ((int)h.arr.size() > L) {
        const int top = h.extract_min();
        const int l = h.el[top].left, r = h.el[top].right;
        h.el[l].d.p0 += h.el[top].d.p0;
        h.el[l].d.p1 += h.el[top].d.p1;
        aux_join(h.el[l].d, h.el[top].d, next);
        hs[l] = hxgiveny(h.el[l].d.p0, h.el[l].d.p1);
        const int ll = h.el[l].left;
        if (ll >= 0) h.update_key(l, key_degrade(h.el[ll].d, hs[ll], h.el[l].d, hs[l]));
        if (r >= 0) h.update_key(r, key_degrade(h.el[l].d, hs[l], h.el[r].d, hs[r]));
    }
    std::vector<Pair> out;
    for (int e = probs.empty() ? -1 : 0; e >= 0; e = h.el[e].right) out.push_back(h.el[e].d);  // the head is never extracted (key inf)
    probs.swap(out);
}

std::vector<Pair> minus_transform(const std::vector<Pair> &p) {  // :259-270
    std::vector<Pair> o;
    o.reserve(p.size() * p.size());
    for (const Pair &a : p)
        for (const Pair &b : p) o.push_back(Pair(a.p0 * b.p0 + a.p1 * b.p1, a.p0 * b.p1 + a.p1 * b.p0));
    return o;
}
std::vector<Pair> plus_transform(const std::vector<Pair> &p) {  // :272-285
    std::vector<Pair> o;
    o.reserve(2 * p.size() * p.size());
    for (const Pair &a : p)
        for (const Pair &b : p) {
            o.push_back(Pair(a.p0 * b.p0, a.p1 * b.p1));
            o.push_back(Pair(a.p1 * b.p0, a.p0 * b.p1));
        }
    return o;
}

}  // namespace

extern "C" int pc_tv_degrade_pe(int n, int L, const double *h_table, int Y, double *h_pe, int threads) {
    PC_REQUIRE(n >= 0 && n <= 20, "n must be in [0,20]");
    PC_REQUIRE(L >= 1, "L must be positive");
    PC_REQUIRE(h_table && Y >= 1 && h_pe, "null table / output");
    std::vector<std::vector<Pair>> cur(1);
    for (int y = 0; y < Y; ++y) cur[0].push_back(Pair(h_table[2 * y], h_table[2 * y + 1]));
    for (int m = 1; m <= n; ++m) {
        std::vector<std::vector<Pair>> nxt(2 * cur.size());
        // the children of different parents are independent: a static split over host threads changes no result
        const int T = std::max(1, std::min<int>(threads, (int)cur.size()));
        auto work = [&](int t) {
            for (size_t i = t; i < cur.size(); i += T) {
                nxt[2 * i] = minus_transform(cur[i]);
                degrade(nxt[2 * i], L);
                nxt[2 * i + 1] = plus_transform(cur[i]);
                degrade(nxt[2 * i + 1], L);
            }
        };
        if (T == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back(work, t);
            for (auto &x : th) x.join();
        }
        cur.swap(nxt);
    }
    for (size_t i = 0; i < cur.size(); ++i) {  // errorProb :39-45
        double s = 0.0;
        for (const Pair &p : cur[i]) s += std::min(p.p0, p.p1);
        h_pe[i] = s;
    }
    return PC_OK;
}

// ---- q-ary channels: QaryMemorylessDistribution.degrade = degrade_dynamic (ScalarDistributions/QaryMemorylessDistribution.py
// :215-260): q-1 one-hot binary channels (:98-153), each degraded to M = floor(L^(1/(q-1)) + eps) letters with the binary
// routine above while tracking which original letters every new letter absorbed, the product alphabet of the q-1 degraded
// channels, zero-probability letters dropped, and a normalisation whose sum runs over the SORTED probabilities (:736-751).
namespace {

typedef std::vector<double> QSym;  // q probabilities of one output letter

std::vector<QSym> q_minus(const std::vector<QSym> &p, int q) {  // :182-196
    std::vector<QSym> o;
    o.reserve(p.size() * p.size());
    for (const QSym &a : p)
        for (const QSym &b : p) {
            QSym t(q, 0.0);
            for (int x1 = 0; x1 < q; ++x1)
                for (int x2 = 0; x2 < q; ++x2) t[(x1 + x2) % q] += a[x1] * b[x2];
            o.push_back(std::move(t));
        }
    return o;
}
std::vector<QSym> q_plus(const std::vector<QSym> &p, int q) {  // :198-212
    std::vector<QSym> o;
    o.reserve(p.size() * p.size() * q);
    for (const QSym &a : p)
        for (const QSym &b : p)
            for (int u1 = 0; u1 < q; ++u1) {
                QSym t(q, 0.0);
                for (int u2 = 0; u2 < q; ++u2) t[u2] += a[(u1 - u2 + q) % q] * b[u2];
                o.push_back(std::move(t));
            }
    return o;
}

void q_degrade(std::vector<QSym> &probs, int q, int L) {
    const int Yold = (int)probs.size();
    // one-hot binary channels, :98-153
    std::vector<double> marg(q, 0.0), pgt(q, 0.0);
    for (int x = 0; x < q; ++x) {
        double t = 0.0;
        for (const QSym &y : probs) t += y[x];
        marg[x] = t;
    }
    for (int x = q - 2; x >= 0; --x) pgt[x] = pgt[x + 1] + marg[x + 1];
    std::vector<std::vector<Pair>> onehot(q - 1);
    for (int j = 0; j < q - 1; ++j) onehot[j].reserve(Yold);
    for (int y = 0; y < Yold; ++y) {
        double prev0 = 0.0, prev1 = 0.0;
        for (int j = q - 2; j >= 0; --j) {
            const double pb1 = probs[y][j];
            const double pb0 = j == q - 2 ? probs[y][j + 1] : prev1 + prev0;
            prev0 = pb0, prev1 = pb1;
            Pair pr = j == 0 ? Pair(pb0, pb1) : Pair(pb0 / pgt[j - 1], pb1 / pgt[j - 1]);
            pr.head = pr.tail = y;
            onehot[j].push_back(pr);
        }
    }
    const int M = (int)std::floor(std::pow((double)L, 1.0 / (q - 1)) + std::numeric_limits<double>::epsilon());  // :753-755
    // degrade each one-hot channel; letters of zero probability in a channel keep the default mapping 0
    std::vector<std::vector<int>> mapped(q - 1, std::vector<int>(Yold, 0));
    std::vector<int> mult(q - 1, 1);
    long newsize = 1;
    for (int x = 0; x < q - 1; ++x) {
        std::vector<int> next(Yold, -1);
        degrade(onehot[x], M, &next);
        for (int yi = 0; yi < (int)onehot[x].size(); ++yi)
            for (int y = onehot[x][yi].head; y >= 0; y = next[y]) mapped[x][y] = yi;
        if (x > 0) mult[x] = mult[x - 1] * (int)onehot[x - 1].size();
        newsize *= (long)onehot[x].size();
    }
    std::vector<QSym> out((size_t)newsize, QSym(q, 0.0));
    for (int y = 0; y < Yold; ++y) {
        int ynew = 0;
        for (int x = 0; x < q - 1; ++x) ynew += mapped[x][y] * mult[x];
        for (int x = 0; x < q; ++x) out[ynew][x] += probs[y][x];
    }
    // removeZeroProbOutput :708-715 (builtin sum of non-negative floats: positive iff any entry is)
    std::vector<QSym> kept;
    for (QSym &t : out)
        if (py_sum(t) > 0.0) kept.push_back(std::move(t));
    // normalize :736-751: plain running sum over the ascending probabilities
    std::vector<double> flat;
    for (const QSym &t : kept)
        for (double v : t) flat.push_back(v);
    std::sort(flat.begin(), flat.end());
    double total = 0.0;
    for (double v : flat) total += v;
    for (QSym &t : kept)
        for (double &v : t) v /= total;
    probs.swap(kept);
}

double q_error_prob(const std::vector<QSym> &probs) {  // :53-61
    double total = 0.0;
    for (const QSym &t : probs) {
        QSym s(t);
        std::sort(s.begin(), s.end());
        s.pop_back();
        total += py_sum(s);
    }
    return total;
}

}  // namespace

extern "C" int pc_tv_degrade_pe_qary(int q, int n, int L, const double *h_table, int Y, double *h_pe, int threads) {
    PC_REQUIRE(q >= 2 && q <= 16, "alphabet size must be in [2,16]");
    PC_REQUIRE(n >= 0 && n <= 20, "n must be in [0,20]");
    PC_REQUIRE(L >= 1, "L must be positive");
    PC_REQUIRE(h_table && Y >= 1 && h_pe, "null table / output");
    std::vector<std::vector<QSym>> cur(1);
    for (int y = 0; y < Y; ++y) cur[0].push_back(QSym(h_table + (size_t)y * q, h_table + (size_t)(y + 1) * q));
    for (int m = 1; m <= n; ++m) {
        std::vector<std::vector<QSym>> nxt(2 * cur.size());
        const int T = std::max(1, std::min<int>(threads, (int)cur.size()));
        auto work = [&](int t) {
            for (size_t i = t; i < cur.size(); i += T) {
                nxt[2 * i] = q_minus(cur[i], q);
                q_degrade(nxt[2 * i], q, L);
                nxt[2 * i + 1] = q_plus(cur[i], q);
                q_degrade(nxt[2 * i + 1], q, L);
            }
        };
        if (T == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back(work, t);
            for (auto &x : th) x.join();
        }
        cur.swap(nxt);
    }
    for (size_t i = 0; i < cur.size(); ++i) h_pe[i] = q_error_prob(cur[i]);
    return PC_OK;
}
