// scl.cu -- batched SC-list decoding with Rate-0 / Rep / Rate-1 / SPC fast nodes, float64 linear domain,
// arithmetic and decisions identical to QaryPolarEncoderDecoder.listDecode.
//
// Replaces QaryPolarEncoderDecoder.listDecode (QaryPolarEncoderDecoder.py:118-227), recursiveListDecode (:403-757:
// Rate-0 :495, Rep :521, Rate-1 :581, SPC :631, general :684), pickLeastReliableIndices / reliability /
// forkIndices / forkIndicesSpc (:759-820), normalize (:867-872) and the per-path QaryMemorylessVectorDistribution
// arithmetic (QaryMemorylessVectorDistribution.py:26-118).  Binary SCL is q = 2.
//
// Organisation (not a port of the recursion):
//  * the recursion is flattened on the host into a list of ops (MINUS, PLUS, COMBINE and the four fast nodes);
//    it depends on the frozen set only, so every frame of a batch executes the same op list;
//  * one FRAME PER LANE: per-level path vectors live in a global scratch laid out [element][symbol][lane][slot]
//    (slot innermost: the lazy "copy" of a path is just the parent index kept in a per-level map, and the
//    gather through that map stays inside the lane's own 64-byte chunk, so no path vector is ever copied);
//  * path metrics follow the reference's sequential float64 products exactly (np.product is a left-to-right
//    product; np.argpartition's candidate order is fixed to ascending (metric, index), see oracle header);
//  * informationList is not maintained while decoding: a path's information is the inverse transform of its root
//    codeword, recovered once for the selected path (and for all paths only when the caller asks for the list);
//  * node-local codewords of the frozen values and of the genie ("actual") path are produced by a masked
//    butterfly at kernel start.
#include <algorithm>
#include <map>
#include <mutex>

#include "qlog_arith.cuh"
#include "scl_tables.cuh"

namespace pc {

int qsc_ingest_launch(int n, int q, int64_t frames, int64_t Bpad, const double *in, double *out, cudaStream_t st);
int byte_egress_launch(bool bitrev, int n, int R, int64_t frames, int64_t Bpad, const uint8_t *in_t, uint8_t *out,
                       cudaStream_t st);
// scl_path.cu: path-per-lane binary decoder (the default for q = 2)
bool sclp_supported(const pc_plan *plan, int L);
size_t sclp_workspace_bytes(const pc_plan *plan, int L, int64_t B, bool want_list);
size_t sclp_workspace_bytes_packed(const pc_plan *plan, int L, int64_t B, bool want_list);
int64_t sclp_wave_frames(const pc_plan *plan, int L);
int sclp_decode_bytes(const pc_plan *plan, const SclTables *T, int L, const double *d_xy, const uint8_t *d_fv, const uint8_t *d_ainfo,
                      int64_t B, uint8_t *d_info, int32_t *d_res, int32_t *d_lsize, double *d_lprob, double *d_aprob,
                      uint8_t *d_linfo, void *ws, size_t ws_bytes, cudaStream_t st);
int sclp_decode_packed(const pc_plan *plan, const SclTables *T, int L, const double *d_xy, const uint8_t *d_y, const double *h_table,
                       int Y, const uint32_t *d_fvp, const uint32_t *d_ainfo, int64_t B, uint32_t *d_info, int32_t *d_res,
                       int32_t *d_lsize, double *d_lprob, double *d_aprob, uint32_t *d_linfo, void *ws, size_t ws_bytes,
                       cudaStream_t st);


struct SclParams {
    int n, k, L, n_ops, nfrozen;
    int64_t frames, Bpad;
    const SclOp *ops;
    const int32_t *a_src, *f_src, *info_src;
    const int8_t *node_level;
    const uint8_t *rep_coef;
    const double *in_t;        // [N][q][Bpad] channel probabilities, natural order
    const uint8_t *fv;         // [frames][N-k] frozen values (caller layout)
    const uint8_t *ainfo;      // [frames][k] actual information (caller layout)
    double *vals;              // [warps] x { V: [N-2][q][32][L], VA: [N-2][q][32] }
    uint8_t *codes;            // [warps] x { R: [4N-4][32][L], RA: [4N-4][32], A: [N][32], F: [N][32] }
    uint8_t *info_t;           // [k][Bpad] selected information
    int32_t *result;           // [Bpad] ProbResult
    // optional final-list outputs (tests): list_size [Bpad], list_prob [L][Bpad], actual_prob [Bpad], list_info [L][k][Bpad]
    int32_t *list_size;
    double *list_prob, *actual_prob;
    uint8_t *list_info;
};

// LOG: the log domain (QaryPolarEncoderDecoder(..., use_log=True)): inputs, path vectors and metrics are natural logarithms;
// products become sums (np.sum: numpy's pairwise order), the transforms go through logaddexp / logsumexp (qlog_arith.cuh)
template <int Q, bool LOG = false>
__global__ void __launch_bounds__(SCL_THREADS) scl_decode_kernel(const SclParams p) {
    constexpr int LQ = Q <= 3 ? SCL_LMAX : 8;  // largest list this instantiation is sized for
    const int n = p.n, N = 1 << n, L = p.L, k = p.k;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (SCL_THREADS / 32) + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * (SCL_THREADS / 32);
    const int64_t groups = (p.frames + 31) / 32;
    const int64_t v_elems = (int64_t)(N - 2) * Q * 32;  // per warp, times L for V
    double *Vb = p.vals + (int64_t)warp_global * v_elems * (L + 1);
    double *VAb = Vb + v_elems * L;
    const int64_t r_pos = (int64_t)4 * N - 4;
    uint8_t *Rb = p.codes + (int64_t)warp_global * ((r_pos * 32) * (L + 1) + (int64_t)2 * N * 32);
    uint8_t *RAb = Rb + r_pos * 32 * L;
    uint8_t *Ab = RAb + r_pos * 32;
    uint8_t *Fb = Ab + (int64_t)N * 32;

    for (int64_t grp = warp_global; grp < groups; grp += warps_total) {
        const int64_t col = grp * 32 + lane;
        const bool live = col < p.frames;
        const int64_t fr = live ? col : 0;
        const double *CH = p.in_t + col;
        // ---- accessors -------------------------------------------------------------------------------
        auto vidx = [&](int l, int h, int x) -> int64_t { return ((int64_t)(((1 << l) - 2) + h) * Q + x) * 32 + lane; };
        auto ldV = [&](int l, int h, int x, int slot) -> double {
            return l == n ? CH[(int64_t)(h * Q + x) * p.Bpad] : Vb[vidx(l, h, x) * L + slot];
        };
        auto ldVA = [&](int l, int h, int x) -> double {
            return l == n ? CH[(int64_t)(h * Q + x) * p.Bpad] : VAb[vidx(l, h, x)];
        };
        auto ridx = [&](int l, int c, int pos) -> int64_t { return ((int64_t)(2 * ((1 << l) - 2) + c * (1 << l)) + pos) * 32 + lane; };
        auto A = [&](int pos) -> uint8_t & { return Ab[(int64_t)pos * 32 + lane]; };
        auto F = [&](int pos) -> uint8_t & { return Fb[(int64_t)pos * 32 + lane]; };

        // ---- node-local codewords of the genie path (A) and of the frozen values (F): masked butterfly ----
        for (int pos = 0; pos < N; ++pos) {
            const int sa = p.a_src[pos], sf = p.f_src[pos];
            A(pos) = sa >= 0 ? p.ainfo[fr * k + sa] : p.fv[fr * p.nfrozen + (~sa)];
            F(pos) = sf >= 0 ? p.fv[fr * p.nfrozen + sf] : (uint8_t)0;
        }
        for (int t = 0; t < n; ++t) {
            const int s = 1 << t;
            for (int pos = 0; pos < N; ++pos) {
                if ((pos & s) || p.node_level[pos] <= t) continue;
                int a = A(pos), b = A(pos + s);
                int sum = a + b;
                A(pos) = (uint8_t)(sum >= Q ? sum - Q : sum);
                A(pos + s) = (uint8_t)(b ? Q - b : 0);
                a = F(pos), b = F(pos + s);
                sum = a + b;
                F(pos) = (uint8_t)(sum >= Q ? sum - Q : sum);
                F(pos + s) = (uint8_t)(b ? Q - b : 0);
            }
        }

        double prob[SCL_LMAX];
        double actual_prob = LOG ? 0.0 : 1.0;
        uint8_t omap[25][2][SCL_LMAX];
        int nl[25][2];
        int nin[25];
        prob[0] = LOG ? 0.0 : 1.0;
        nin[n] = 1;
        nl[n][0] = 1;

        // divide the list by its maximum (normalize, :867-872); returns the maximum
        auto normalize = [&](int cnt) -> double {
            double mx = prob[0];
            for (int t = 1; t < cnt; ++t)
                if (prob[t] > mx) mx = prob[t];
            for (int t = 0; t < cnt; ++t) prob[t] = LOG ? prob[t] - mx : prob[t] / mx;
            return mx;
        };
        // keep the min(#nonzero, L) largest candidates, ascending (metric, index)
        auto prune = [&](const double *m, int C, int *keep) -> int {
            int nz = 0;
            for (int c = 0; c < C; ++c) nz += LOG ? (m[c] != -INFINITY) : (m[c] != 0.0);  // np.isneginf / np.count_nonzero
            const int ns = nz < L ? nz : L;
            int cnt = 0;
            for (int c = 0; c < C; ++c) {
                if (cnt == ns) {
                    if (ns == 0 || !(m[c] >= m[keep[0]])) continue;
                    for (int t = 0; t + 1 < cnt; ++t) keep[t] = keep[t + 1];  // drop the smallest
                    --cnt;
                }
                int t = cnt - 1;
                while (t >= 0 && m[keep[t]] > m[c]) {
                    keep[t + 1] = keep[t];
                    --t;
                }
                keep[t + 1] = c;
                ++cnt;
            }
            return ns;
        };

        for (int oi = 0; oi < p.n_ops; ++oi) {
            const SclOp op = p.ops[oi];
            const int l = op.l, size = 1 << l, half = size >> 1;
            if (op.kind == OP_MINUS || op.kind == OP_PLUS) {
                const bool plus = op.kind == OP_PLUS;
                const int cnt = plus ? nl[l - 1][0] : nin[l];
                nin[l - 1] = cnt;
                for (int h = 0; h < half; ++h) {
                    for (int t = 0; t <= cnt; ++t) {  // t == cnt: the genie path
                        const bool act = t == cnt;
                        const int src = act ? 0 : (plus ? omap[l - 1][0][t] : t);
                        double a[Q], b[Q], d[Q];
#pragma unroll
                        for (int x = 0; x < Q; ++x) {
                            a[x] = act ? ldVA(l, h, x) : ldV(l, h, x, src);
                            b[x] = act ? ldVA(l, h + half, x) : ldV(l, h + half, x, src);
                        }
                        const int u1 = !plus ? 0 : (act ? RAb[ridx(l - 1, 0, h)] : Rb[ridx(l - 1, 0, h) * L + t]);
                        if (LOG) {
                            q_node_log<Q>(a, b, plus, u1, d);
                        } else {
                            if (!plus) {  // QaryMemorylessVectorDistribution.py:36-42
#pragma unroll
                                for (int x = 0; x < Q; ++x) d[x] = 0.0;
#pragma unroll
                                for (int x1 = 0; x1 < Q; ++x1)
#pragma unroll
                                    for (int x2 = 0; x2 < Q; ++x2)
                                        d[(x1 + x2) % Q] = __dadd_rn(d[(x1 + x2) % Q], __dmul_rn(a[x1], b[x2]));
                            } else {  // :56-62
#pragma unroll
                                for (int u2 = 0; u2 < Q; ++u2) {
                                    double av = a[0];
#pragma unroll
                                    for (int x = 1; x < Q; ++x)
                                        if ((u1 + u2) % Q == x) av = a[x];
                                    d[u2] = __dadd_rn(0.0, __dmul_rn(av, b[(Q - u2) % Q]));
                                }
                            }
                            double tsum = 0.0;
#pragma unroll
                            for (int x = 0; x < Q; ++x) tsum = __dadd_rn(tsum, d[x]);
                            if (tsum != 0.0) {
#pragma unroll
                                for (int x = 0; x < Q; ++x) d[x] = d[x] / tsum;
                            }
                        }
#pragma unroll
                        for (int x = 0; x < Q; ++x) {
                            if (act)
                                VAb[vidx(l - 1, h, x)] = d[x];
                            else
                                Vb[vidx(l - 1, h, x) * L + t] = d[x];
                        }
                    }
                }
                continue;
            }
            if (op.kind == OP_COMBINE) {  // :726-754, natural order: [m + p, -p]
                const int cnt = nl[l - 1][1];
                for (int t = 0; t < cnt; ++t) omap[l][op.c][t] = omap[l - 1][0][omap[l - 1][1][t]];
                nl[l][op.c] = cnt;
                for (int h = 0; h < half; ++h) {
                    for (int t = 0; t <= cnt; ++t) {
                        const bool act = t == cnt;
                        const int mi = act ? 0 : omap[l - 1][1][t];
                        const int m = act ? RAb[ridx(l - 1, 0, h)] : Rb[ridx(l - 1, 0, h) * L + mi];
                        const int pp = act ? RAb[ridx(l - 1, 1, h)] : Rb[ridx(l - 1, 1, h) * L + t];
                        int s = m + pp;
                        s = s >= Q ? s - Q : s;
                        const uint8_t lo = (uint8_t)s, hi = (uint8_t)(pp ? Q - pp : 0);
                        if (act) {
                            RAb[ridx(l, op.c, h)] = lo;
                            RAb[ridx(l, op.c, h + half)] = hi;
                        } else {
                            Rb[ridx(l, op.c, h) * L + t] = lo;
                            Rb[ridx(l, op.c, h + half) * L + t] = hi;
                        }
                    }
                }
                continue;
            }
            // ------------------------------- fast nodes ---------------------------------------------------
            const int cnt = nin[l];
            const int i0 = op.i;
            // the genie path's node codeword (natural order) is A[i0 .. i0+size)
            auto aval = [&](int j) -> double {  // reference order j, natural position rev(j)
                const int pos = bitrev_n((uint32_t)j, l);
                const int sym = A(i0 + pos);
                double v = ldVA(l, pos, 0);
#pragma unroll
                for (int x = 1; x < Q; ++x)
                    if (sym == x) v = ldVA(l, pos, x);
                return v;
            };
            double aprod = 1.0;
            if (LOG) aprod = np_sum_pow2(aval, size);
            for (int j = 0; j < size; ++j) {
                const int pos = bitrev_n((uint32_t)j, l);
                if (!LOG) {
                    const double v = aval(j);
                    aprod = j == 0 ? v : __dmul_rn(aprod, v);
                }
                RAb[ridx(l, op.c, pos)] = (uint8_t)A(i0 + pos);
            }
            auto ldsym = [&](int pos, int slot, int sym) -> double {
                double v = ldV(l, pos, 0, slot);
#pragma unroll
                for (int x = 1; x < Q; ++x)
                    if (sym == x) v = ldV(l, pos, x, slot);
                return v;
            };
            int nout = cnt;
            if (op.kind == OP_RATE0) {  // :495-518
                for (int t = 0; t < cnt; ++t) {
                    double pr = 1.0;
                    if (LOG) pr = np_sum_pow2([&](int j) { const int pos = bitrev_n((uint32_t)j, l); return ldsym(pos, t, F(i0 + pos)); }, size);
                    for (int j = 0; j < size; ++j) {
                        const int pos = bitrev_n((uint32_t)j, l);
                        const int sym = F(i0 + pos);
                        if (!LOG) {
                            const double v = ldsym(pos, t, sym);
                            pr = j == 0 ? v : __dmul_rn(pr, v);
                        }
                        Rb[ridx(l, op.c, pos) * L + t] = (uint8_t)sym;
                    }
                    prob[t] = LOG ? prob[t] + pr : __dmul_rn(prob[t], pr);
                    omap[l][op.c][t] = (uint8_t)t;
                }
            } else if (op.kind == OP_REP) {  // :521-578
                double cand[LQ * Q];
                int keep[LQ * Q];
                const uint8_t *coef = p.rep_coef + op.coef_off;
                for (int t = 0; t < cnt; ++t)
                    for (int s = 0; s < Q; ++s) {
                        auto rval = [&](int j) -> double {
                            const int pos = bitrev_n((uint32_t)j, l);
                            return ldsym(pos, t, (F(i0 + pos) + s * coef[pos]) % Q);
                        };
                        double pr = 1.0;
                        if (LOG) {
                            pr = np_sum_pow2(rval, size);
                        } else {
                            for (int j = 0; j < size; ++j) {
                                const double v = rval(j);
                                pr = j == 0 ? v : __dmul_rn(pr, v);
                            }
                        }
                        cand[s * cnt + t] = LOG ? prob[t] + pr : __dmul_rn(prob[t], pr);
                    }
                const int C = cnt * Q;
                if (C > L) {
                    nout = prune(cand, C, keep);
                } else {
                    for (int c = 0; c < C; ++c) keep[c] = c;
                    nout = C;
                }
                for (int t = 0; t < nout; ++t) {
                    const int s = keep[t] / cnt;
                    omap[l][op.c][t] = (uint8_t)(keep[t] % cnt);
                    prob[t] = cand[keep[t]];
                    for (int pos = 0; pos < size; ++pos)
                        Rb[ridx(l, op.c, pos) * L + t] = (uint8_t)((F(i0 + pos) + s * coef[pos]) % Q);
                }
            } else {  // Rate-1 :581-628 and SPC :631-682
                const bool spc = op.kind == OP_SPC;
                const int nfork = spc ? 3 : 2, npick = spc ? 4 : 2;
                const int fs = spc ? Q * Q * Q : Q * Q;
                const int fval = spc ? p.fv[fr * p.nfrozen + op.fv_idx] : 0;
                double cand[LQ * Q * Q * Q];
                int keep[SCL_LMAX];
                int16_t pick[SCL_LMAX][4];
                uint8_t delta[SCL_LMAX];
                for (int t = 0; t < cnt; ++t) {
                    // pickLeastReliableIndices: the npick largest (score, j), ascending
                    const double lowest = LOG ? -INFINITY : -1.0;
                    double sc[4] = {lowest, lowest, lowest, lowest};
                    int sj[4] = {0, 0, 0, 0};
                    double prodmax = LOG ? 0.0 : 1.0;  // log: builtin sum() starts from 0
                    for (int j = 0; j < size; ++j) {
                        const int pos = bitrev_n((uint32_t)j, l);
                        double m1 = lowest, m2 = lowest;
#pragma unroll
                        for (int x = 0; x < Q; ++x) {
                            const double v = ldV(l, pos, x, t);
                            if (v > m1) {
                                m2 = m1;
                                m1 = v;
                            } else if (v > m2) {
                                m2 = v;
                            }
                        }
                        const double s = LOG ? m2 - m1 : m2 / m1;  // reliability, :763-768
                        if (s >= sc[0]) {  // enters the top-npick buffer (sc[npick-1] is the largest)
                            int w = 0;
                            while (w + 1 < npick && s >= sc[w + 1]) {
                                sc[w] = sc[w + 1];
                                sj[w] = sj[w + 1];
                                ++w;
                            }
                            sc[w] = s;
                            sj[w] = j;
                        }
                    }
                    // sc[] / sj[] hold the npick largest ascending in slots [0, npick) -- but unfilled slots start at
                    // -1 and are pushed out by the first npick positions (size >= npick always holds)
                    int sumconst = 0;
                    bool first = !LOG;
                    for (int j = 0; j < size; ++j) {
                        bool forked = false;
                        for (int w = 0; w < npick; ++w) forked |= (sj[w] == j);
                        if (forked) continue;
                        const int pos = bitrev_n((uint32_t)j, l);
                        double mv = ldV(l, pos, 0, t);
                        int am = 0;
#pragma unroll
                        for (int x = 1; x < Q; ++x) {
                            const double v = ldV(l, pos, x, t);
                            if (v > mv) {
                                mv = v;
                                am = x;
                            }
                        }
                        sumconst += am;
                        prodmax = first ? mv : (LOG ? prodmax + mv : __dmul_rn(prodmax, mv));
                        first = false;
                    }
                    const double base_prob = LOG ? prob[t] + prodmax : __dmul_rn(prob[t], prodmax);
                    for (int w = 0; w < npick; ++w) pick[t][w] = (int16_t)sj[w];
                    const int dl = ((fval - sumconst) % Q + Q) % Q;
                    delta[t] = (uint8_t)dl;
                    for (int f = 0; f < fs; ++f) {
                        int dg[3], rem = f, sf = 0;
                        for (int w = nfork - 1; w >= 0; --w) {
                            dg[w] = rem % Q;
                            rem /= Q;
                        }
                        double pf = 1.0;
                        for (int w = 0; w < nfork; ++w) {
                            const double v = ldsym(bitrev_n((uint32_t)sj[w], l), t, dg[w]);
                            pf = LOG ? (w == 0 ? 0.0 + v : pf + v) : (w == 0 ? v : __dmul_rn(pf, v));
                            sf += dg[w];
                        }
                        if (spc) {
                            const int dep = ((dl - sf) % Q + Q) % Q;
                            const double vd = ldsym(bitrev_n((uint32_t)sj[3], l), t, dep);
                            pf = LOG ? pf + vd : __dmul_rn(pf, vd);
                        }
                        cand[t * fs + f] = LOG ? pf + base_prob : __dmul_rn(pf, base_prob);
                    }
                }
                const int C = cnt * fs;
                int keepall[SCL_LMAX];
                if (C > L) {
                    nout = prune(cand, C, keep);
                } else {
                    for (int c = 0; c < C; ++c) keepall[c] = c;
                    nout = C;
                }
                double newprob[SCL_LMAX];
                for (int t = 0; t < nout; ++t) {
                    const int cidx = C > L ? keep[t] : keepall[t];
                    const int src = cidx / fs, f = cidx % fs;
                    newprob[t] = cand[cidx];
                    omap[l][op.c][t] = (uint8_t)src;
                    int dg[3], rem = f, sf = 0;
                    for (int w = nfork - 1; w >= 0; --w) {
                        dg[w] = rem % Q;
                        rem /= Q;
                    }
                    for (int w = 0; w < nfork; ++w) sf += dg[w];
                    for (int j = 0; j < size; ++j) {
                        const int pos = bitrev_n((uint32_t)j, l);
                        int sym = -1;
                        for (int w = 0; w < nfork; ++w)
                            if (pick[src][w] == j) sym = dg[w];
                        if (spc && pick[src][3] == j) sym = ((delta[src] - sf) % Q + Q) % Q;
                        if (sym < 0) {
                            double mv = ldV(l, pos, 0, src);
                            sym = 0;
#pragma unroll
                            for (int x = 1; x < Q; ++x) {
                                const double v = ldV(l, pos, x, src);
                                if (v > mv) {
                                    mv = v;
                                    sym = x;
                                }
                            }
                        }
                        Rb[ridx(l, op.c, pos) * L + t] = (uint8_t)sym;
                    }
                }
                for (int t = 0; t < nout; ++t) prob[t] = newprob[t];
            }
            nl[l][op.c] = nout;
            const double nw = normalize(nout);
            actual_prob = LOG ? actual_prob + (aprod - nw) : __dmul_rn(actual_prob, aprod / nw);
        }

        // ---- final selection (listDecode :172-213): the genie path is in the list iff a root codeword equals it ----
        const int cnt = nl[n][0];
        int found = -1;
        for (int t = 0; t < cnt && found < 0; ++t) {
            bool eq = true;
            for (int pos = 0; pos < N && eq; ++pos) eq = Rb[ridx(n, 0, pos) * L + t] == RAb[ridx(n, 0, pos)];
            if (eq) found = t;
        }
        double maxp = prob[0], minp = prob[0];
        for (int t = 1; t < cnt; ++t) {
            maxp = prob[t] > maxp ? prob[t] : maxp;
            minp = prob[t] < minp ? prob[t] : minp;
        }
        int res;
        if (found >= 0)
            res = prob[found] == maxp ? 0 : 1;
        else
            res = actual_prob > maxp ? 2 : (actual_prob == maxp ? 3 : (actual_prob >= minp ? 4 : 5));
        const int sel = found >= 0 ? found : 0;
        p.result[col] = res;
        if (p.list_size) {
            p.list_size[col] = cnt;
            p.actual_prob[col] = actual_prob;
            for (int t = 0; t < L; ++t) p.list_prob[(int64_t)t * p.Bpad + col] = t < cnt ? prob[t] : (LOG ? -INFINITY : 0.0);
        }
        // information of a path = gather of T(root codeword) (u = T(x) in natural order)
        const int npaths = p.list_info ? cnt : 1;
        for (int pi = 0; pi < npaths; ++pi) {
            const int t = p.list_info ? pi : sel;
            for (int pos = 0; pos < N; ++pos) A(pos) = Rb[ridx(n, 0, pos) * L + t];
            for (int tt = 0; tt < n; ++tt) {
                const int s = 1 << tt;
                for (int pos = 0; pos < N; ++pos) {
                    if (pos & s) continue;
                    const int a = A(pos), b = A(pos + s);
                    int sum = a + b;
                    A(pos) = (uint8_t)(sum >= Q ? sum - Q : sum);
                    A(pos + s) = (uint8_t)(b ? Q - b : 0);
                }
            }
            for (int j = 0; j < k; ++j) {
                const uint8_t v = A(p.info_src[j]);
                if (p.list_info) p.list_info[((int64_t)pi * k + j) * p.Bpad + col] = v;
                if (t == sel) p.info_t[(int64_t)j * p.Bpad + col] = v;
            }
        }
    }
}

struct SclLayout {
    int64_t chunk, Bpad;
    size_t off_in, off_info, off_res, off_vals, off_codes, off_lsize, off_lprob, off_aprob, off_linfo, total;
    int grid;
};

static SclLayout scl_layout(const pc_plan *plan, int L, int64_t chunk, bool want_list) {
    SclLayout Y;
    const int64_t N = plan->N, q = plan->q, k = plan->k > 0 ? plan->k : 1;
    Y.chunk = chunk;
    Y.Bpad = round_up(chunk, 32);
    const int64_t blocks = (Y.Bpad + SCL_THREADS - 1) / SCL_THREADS;
    const int64_t gmax = (int64_t)num_sms() * 2;
    Y.grid = (int)(blocks < gmax ? blocks : gmax);
    const int64_t warps = (int64_t)Y.grid * (SCL_THREADS / 32);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o += align256(bytes + 256);
        return at;
    };
    Y.off_in = take((size_t)N * q * Y.Bpad * 8);
    Y.off_info = take((size_t)k * Y.Bpad);
    Y.off_res = take((size_t)Y.Bpad * 4);
    Y.off_vals = take((size_t)warps * (N > 2 ? N - 2 : 1) * q * 32 * (L + 1) * 8);
    Y.off_codes = take((size_t)warps * (((4 * N - 4) * 32) * (L + 1) + 2 * N * 32));
    Y.off_lsize = take(want_list ? (size_t)Y.Bpad * 4 : 0);
    Y.off_lprob = take(want_list ? (size_t)L * Y.Bpad * 8 : 0);
    Y.off_aprob = take(want_list ? (size_t)Y.Bpad * 8 : 0);
    Y.off_linfo = take(want_list ? (size_t)L * k * Y.Bpad : 0);
    Y.total = o;
    return Y;
}

__global__ void scl_list_egress_kernel(int L, int k, int64_t frames, int64_t Bpad, const int32_t *ls_t, const double *lp_t,
                                       const double *ap_t, int32_t *ls, double *lp, double *ap) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    ls[f] = ls_t[f];
    ap[f] = ap_t[f];
    for (int t = 0; t < L; ++t) lp[f * L + t] = lp_t[(int64_t)t * Bpad + f];
}

template <int Q>
static int scl_launch(const SclParams &p, int grid, cudaStream_t st, bool use_log) {
    prof_mark(st);
    if (use_log)
        scl_decode_kernel<Q, true><<<grid, SCL_THREADS, 0, st>>>(p);
    else
        scl_decode_kernel<Q, false><<<grid, SCL_THREADS, 0, st>>>(p);
    prof_mark(st);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

}  // namespace pc

extern "C" {

static size_t scl_workspace_common(const pc_plan *plan, int L, int64_t B, int want_list, int use_log) {
    if (!plan || B <= 0 || L < 1 || L > pc::SCL_LMAX) return 256;
    if (!use_log && pc::sclp_supported(plan, L)) return pc::sclp_workspace_bytes(plan, L, B, want_list != 0);
    int64_t chunk = pc::round_up(B, 32);
    const int64_t cap = (int64_t)pc::num_sms() * 2 * pc::SCL_THREADS * 4;
    if (chunk > cap) chunk = cap;
    return pc::scl_layout(plan, L, chunk, want_list != 0).total;
}
size_t pc_scl_workspace_bytes(const pc_plan *plan, int L, int64_t B, int want_list) {
    return scl_workspace_common(plan, L, B, want_list, 0);
}
size_t pc_scl_workspace_bytes_log(const pc_plan *plan, int L, int64_t B, int want_list) {
    return scl_workspace_common(plan, L, B, want_list, 1);
}

int64_t pc_scl_wave_frames(const pc_plan *plan, int L) {
    if (!plan || L < 1 || L > pc::SCL_LMAX) return 0;
    if (pc::sclp_supported(plan, L)) return pc::sclp_wave_frames(plan, L);
    return (int64_t)pc::num_sms() * 2 * pc::SCL_THREADS;
}

static int scl_decode_common(const pc_plan *plan, int L, const double *d_xy, const uint8_t *d_frozen_values,
                             const uint8_t *d_actual_info, int64_t B, uint8_t *d_info, int32_t *d_prob_result,
                             int32_t *d_list_size, double *d_list_prob, double *d_actual_prob, uint8_t *d_list_info,
                             void *d_workspace, size_t workspace_bytes, void *stream, int use_log);

/* see include/polarcub_b200.h */
int pc_scl_decode_probs(const pc_plan *plan, int L, const double *d_xy, const uint8_t *d_frozen_values,
                        const uint8_t *d_actual_info, int64_t B, uint8_t *d_info, int32_t *d_prob_result,
                        int32_t *d_list_size, double *d_list_prob, double *d_actual_prob, uint8_t *d_list_info,
                        void *d_workspace, size_t workspace_bytes, void *stream) {
    return scl_decode_common(plan, L, d_xy, d_frozen_values, d_actual_info, B, d_info, d_prob_result, d_list_size, d_list_prob,
                             d_actual_prob, d_list_info, d_workspace, workspace_bytes, stream, 0);
}
int pc_scl_decode_logprobs(const pc_plan *plan, int L, const double *d_xy_log, const uint8_t *d_frozen_values,
                           const uint8_t *d_actual_info, int64_t B, uint8_t *d_info, int32_t *d_prob_result,
                           int32_t *d_list_size, double *d_list_prob, double *d_actual_prob, uint8_t *d_list_info,
                           void *d_workspace, size_t workspace_bytes, void *stream) {
    return scl_decode_common(plan, L, d_xy_log, d_frozen_values, d_actual_info, B, d_info, d_prob_result, d_list_size,
                             d_list_prob, d_actual_prob, d_list_info, d_workspace, workspace_bytes, stream, 1);
}

static int scl_decode_common(const pc_plan *plan, int L, const double *d_xy, const uint8_t *d_frozen_values,
                             const uint8_t *d_actual_info, int64_t B, uint8_t *d_info, int32_t *d_prob_result,
                             int32_t *d_list_size, double *d_list_prob, double *d_actual_prob, uint8_t *d_list_info,
                             void *d_workspace, size_t workspace_bytes, void *stream, int use_log) {
    using namespace pc;
    PC_REQUIRE(plan != nullptr, "plan is null");
    PC_REQUIRE(plan->q == 2 || plan->q == 3 || plan->q == 4 || plan->q == 5, "SCL is built for q in {2,3,4,5}");
    PC_REQUIRE(plan->n >= 1 && plan->n <= 16, "SCL needs 2 <= N <= 65536");
    PC_REQUIRE(L >= 1 && L <= SCL_LMAX, "list size must be in [1,32]");
    PC_REQUIRE(L * plan->q * plan->q * plan->q <= SCL_LMAX * 27 + 0 && (plan->q <= 3 || L <= 8), "L too large for this q");
    PC_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return PC_OK;
    PC_REQUIRE(d_xy && d_actual_info && d_prob_result && d_workspace && (d_info || plan->k == 0), "null buffer");
    PC_REQUIRE(d_frozen_values || plan->k == plan->N, "frozen values missing");
    PC_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
    const bool want_list = d_list_size != nullptr;
    if (want_list) PC_REQUIRE(d_list_prob && d_actual_prob, "list outputs incomplete");
    SclTables *T = scl_tables(plan);
    if (!T) return PC_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    if (!use_log && sclp_supported(plan, L))  // q = 2: one path per lane, 32 / L frames per warp (scl_path.cu)
        return sclp_decode_bytes(plan, T, L, d_xy, d_frozen_values, d_actual_info, B, d_info, d_prob_result, d_list_size,
                                 d_list_prob, d_actual_prob, d_list_info, d_workspace, workspace_bytes, st);
    int64_t chunk = round_up(B, 32);
    const int64_t cap = (int64_t)num_sms() * 2 * SCL_THREADS * 4;
    if (chunk > cap) chunk = cap;
    while (chunk > 32 && scl_layout(plan, L, chunk, want_list).total > workspace_bytes) chunk = round_up(chunk / 2, 32);
    SclLayout Y = scl_layout(plan, L, chunk, want_list);
    if (Y.total > workspace_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed for a 32-frame chunk", workspace_bytes, Y.total);
        return PC_ERR_NOMEM;
    }
    const int N = plan->N, q = plan->q, k = plan->k;
    char *base = (char *)d_workspace;
    SclParams p{};
    p.n = plan->n;
    p.k = k;
    p.L = L;
    p.n_ops = (int)T->ops.size();
    p.nfrozen = N - k;
    p.Bpad = Y.Bpad;
    p.ops = T->d_ops;
    p.a_src = T->d_a_src;
    p.f_src = T->d_f_src;
    p.info_src = T->d_info_src;
    p.node_level = T->d_node_level;
    p.rep_coef = T->d_rep_coef;
    p.in_t = (const double *)(base + Y.off_in);
    p.vals = (double *)(base + Y.off_vals);
    p.codes = (uint8_t *)(base + Y.off_codes);
    p.info_t = (uint8_t *)(base + Y.off_info);
    p.result = (int32_t *)(base + Y.off_res);
    if (want_list) {
        p.list_size = (int32_t *)(base + Y.off_lsize);
        p.list_prob = (double *)(base + Y.off_lprob);
        p.actual_prob = (double *)(base + Y.off_aprob);
        p.list_info = d_list_info ? (uint8_t *)(base + Y.off_linfo) : nullptr;
    }
    for (int64_t f0 = 0; f0 < B; f0 += chunk) {
        const int64_t frames = (B - f0) < chunk ? (B - f0) : chunk;
        const int64_t tiles = (frames + 31) / 32;
        p.frames = frames;
        p.fv = d_frozen_values ? d_frozen_values + f0 * (N - k) : nullptr;
        p.ainfo = d_actual_info + f0 * k;
        int rc = qsc_ingest_launch(plan->n, q, frames, Y.Bpad, d_xy + f0 * N * q, (double *)p.in_t, st);
        if (rc) return rc;
        const int64_t blocks = (tiles * 32 + SCL_THREADS - 1) / SCL_THREADS;
        const int grid = (int)(blocks < Y.grid ? blocks : Y.grid);
        switch (q) {
            case 2: rc = scl_launch<2>(p, grid, st, use_log != 0); break;
            case 3: rc = scl_launch<3>(p, grid, st, use_log != 0); break;
            case 4: rc = scl_launch<4>(p, grid, st, use_log != 0); break;
            default: rc = scl_launch<5>(p, grid, st, use_log != 0); break;
        }
        if (rc) return rc;
        if (k > 0) {
            rc = byte_egress_launch(false, plan->n, k, frames, Y.Bpad, p.info_t, d_info + f0 * k, st);
            if (rc) return rc;
        }
        PC_CUDA(cudaMemcpyAsync(d_prob_result + f0, p.result, sizeof(int32_t) * frames, cudaMemcpyDeviceToDevice, st));
        if (want_list) {
            scl_list_egress_kernel<<<(unsigned)((frames + 255) / 256), 256, 0, st>>>(
                L, k, frames, Y.Bpad, p.list_size, p.list_prob, p.actual_prob, d_list_size + f0, d_list_prob + f0 * L,
                d_actual_prob + f0);
            PC_LAUNCH_CHECK();
            if (d_list_info && k > 0) {
                rc = byte_egress_launch(false, plan->n, L * k, frames, Y.Bpad, p.list_info, d_list_info + f0 * L * k, st);
                if (rc) return rc;
            }
        }
    }
    return PC_OK;
}

/* see include/polarcub_b200.h */
size_t pc_scl_workspace_bytes_packed(const pc_plan *plan, int L, int64_t B, int want_list) {
    if (!plan || B <= 0 || L < 1 || L > pc::SCL_LMAX || !pc::sclp_supported(plan, L)) return 256;
    return pc::sclp_workspace_bytes_packed(plan, L, B, want_list != 0);
}

int pc_scl_decode_packed(const pc_plan *plan, int L, const double *d_xy, const uint8_t *d_y, const double *h_table, int Y,
                         const uint32_t *d_frozen_packed, const uint32_t *d_actual_info_packed, int64_t B,
                         uint32_t *d_info_packed, int32_t *d_prob_result, int32_t *d_list_size, double *d_list_prob,
                         double *d_actual_prob, uint32_t *d_list_info_packed, void *d_workspace, size_t workspace_bytes,
                         void *stream) {
    using namespace pc;
    PC_REQUIRE(plan != nullptr, "plan is null");
    PC_REQUIRE(plan->q == 2, "the packed list decoder is binary (q = 2)");
    PC_REQUIRE(L >= 1 && L <= SCL_LMAX, "list size must be in [1,32]");
    PC_REQUIRE(sclp_supported(plan, L), "block length not supported by the packed list decoder (2 <= N <= 8192)");
    PC_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return PC_OK;
    PC_REQUIRE((d_xy != nullptr) != (d_y != nullptr), "exactly one of d_xy / d_y must be given");
    if (d_y) PC_REQUIRE(h_table != nullptr && Y >= 1 && Y <= 256, "channel table missing or larger than 256 rows");
    if (d_y) PC_REQUIRE(((uintptr_t)d_y & 3) == 0, "d_y must be 4-byte aligned");
    PC_REQUIRE(d_actual_info_packed && d_prob_result && d_workspace && (d_info_packed || plan->k == 0), "null buffer");
    PC_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (d_list_size) PC_REQUIRE(d_list_prob && d_actual_prob, "list outputs incomplete");
    SclTables *T = scl_tables(plan);
    if (!T) return PC_ERR_CUDA;
    return sclp_decode_packed(plan, T, L, d_xy, d_y, h_table, Y, d_frozen_packed, d_actual_info_packed, B, d_info_packed,
                              d_prob_result, d_list_size, d_list_prob, d_actual_prob, d_list_info_packed, d_workspace,
                              workspace_bytes, (cudaStream_t)stream);
}

int pc_scl_decode_symbols(const pc_plan *plan, int L, const uint8_t *d_y, const double *h_table, int Y,
                          const uint32_t *d_frozen_packed, const uint32_t *d_actual_info_packed, int64_t B,
                          uint32_t *d_info_packed, int32_t *d_prob_result, void *d_workspace, size_t workspace_bytes,
                          void *stream) {
    return pc_scl_decode_packed(plan, L, nullptr, d_y, h_table, Y, d_frozen_packed, d_actual_info_packed, B, d_info_packed,
                                d_prob_result, nullptr, nullptr, nullptr, nullptr, d_workspace, workspace_bytes, stream);
}

}  // extern "C"
