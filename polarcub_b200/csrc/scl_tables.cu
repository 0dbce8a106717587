// scl_tables.cu -- host-flattened op list and per-plan tables of the SC-list decoders (scl.cu, scl_warp.cu, scl_path.cu).
//
// The recursion of QaryPolarEncoderDecoder.recursiveListDecode (QaryPolarEncoderDecoder.py:403-757: Rate-0 :495, Rep :521,
// Rate-1 :581, SPC :631, general :684) depends on the frozen set only, so it is flattened once per plan into a list of ops
// that every frame of a batch executes.  Host code only (no kernels).
#include <algorithm>
#include <map>
#include <mutex>

#include "scl_tables.cuh"

namespace pc {

static std::mutex g_scl_mu;
static std::map<const pc_plan *, SclTables *> g_scl_tables;

static void scl_build(const pc_plan *p, SclTables &T, int i, int l, int c, int &info_idx, int &fv_idx) {
    const int size = 1 << l, q = p->q;
    int ninfo = 0;
    for (int j = i; j < i + size; ++j) ninfo += !p->frozen_mask[j];
    SclOp op{};
    op.l = (int8_t)l;
    op.c = (int8_t)c;
    op.i = i;
    op.info_idx = info_idx;
    op.fv_idx = fv_idx;
    auto mark = [&]() {
        for (int j = i; j < i + size; ++j) T.node_level[j] = (int8_t)l;
    };
    if (ninfo == 0) {
        op.kind = OP_RATE0;
        for (int j = 0; j < size; ++j) {
            T.a_src[i + j] = ~(fv_idx + j);
            T.f_src[i + j] = fv_idx + j;
        }
        fv_idx += size;
        mark();
        T.ops.push_back(op);
    } else if (ninfo == 1) {
        op.kind = OP_REP;
        int kpos = 0;
        while (p->frozen_mask[i + kpos]) ++kpos;
        op.kpos = kpos;
        int f = fv_idx;
        for (int j = 0; j < size; ++j) {
            if (j == kpos) {
                T.a_src[i + j] = info_idx;
                T.f_src[i + j] = -1;
                T.info_src[info_idx] = i + j;
            } else {
                T.a_src[i + j] = ~f;
                T.f_src[i + j] = f;
                ++f;
            }
        }
        // natural-order T(e_kpos) mod q: T([a;b]) = [T(a)+T(b), -T(b)]
        std::vector<int> cf(size, 0);
        cf[kpos] = 1;
        for (int s = 1; s < size; s <<= 1)
            for (int b = 0; b < size; b += 2 * s)
                for (int j = b; j < b + s; ++j) {
                    const int x = cf[j], y = cf[j + s];
                    cf[j] = (x + y) % q;
                    cf[j + s] = (q - y) % q;
                }
        op.coef_off = (int32_t)T.rep_coef.size();
        for (int j = 0; j < size; ++j) T.rep_coef.push_back((uint8_t)cf[j]);
        if (q == 2) {  // reference order: position j of the node is natural position bitrev(j, l)
            op.coefw_off = (int32_t)T.rep_coef_words.size();
            const int words = size >= 32 ? size / 32 : 1;
            for (int w = 0; w < words; ++w) {
                uint32_t v = 0;
                for (int b = 0; b < 32 && 32 * w + b < size; ++b) {
                    const int j = 32 * w + b;
                    int r = 0;
                    for (int t = 0; t < l; ++t) r |= ((j >> t) & 1) << (l - 1 - t);
                    v |= (uint32_t)(cf[r] & 1) << b;
                }
                T.rep_coef_words.push_back(v);
            }
        }
        fv_idx += size - 1;
        info_idx += 1;
        mark();
        T.ops.push_back(op);
    } else if (ninfo == size) {
        op.kind = OP_RATE1;
        for (int j = 0; j < size; ++j) {
            T.a_src[i + j] = info_idx + j;
            T.f_src[i + j] = -1;
            T.info_src[info_idx + j] = i + j;
        }
        info_idx += size;
        mark();
        T.ops.push_back(op);
    } else if (ninfo == size - 1) {
        // SPC.  The reference treats the frozen value as u[first of the segment] wherever the frozen index
        // really is (QaryPolarEncoderDecoder.py:637, :662, :673): u' = [frozenValue, info...]; reproduced here.
        op.kind = OP_SPC;
        T.a_src[i] = ~fv_idx;
        T.f_src[i] = -1;
        for (int j = 1; j < size; ++j) {
            T.a_src[i + j] = info_idx + j - 1;
            T.f_src[i + j] = -1;
            T.info_src[info_idx + j - 1] = i + j;
        }
        fv_idx += 1;
        info_idx += size - 1;
        mark();
        T.ops.push_back(op);
    } else {
        op.kind = OP_MINUS;
        T.ops.push_back(op);
        scl_build(p, T, i, l - 1, 0, info_idx, fv_idx);
        op.kind = OP_PLUS;
        T.ops.push_back(op);
        scl_build(p, T, i + size / 2, l - 1, 1, info_idx, fv_idx);
        op.kind = OP_COMBINE;
        T.ops.push_back(op);
    }
}

template <class T>
static cudaError_t upload(T *&dst, const std::vector<T> &v) {
    cudaError_t e = cudaMalloc((void **)&dst, sizeof(T) * (v.size() ? v.size() : 1));
    if (e != cudaSuccess) return e;
    if (v.size()) e = cudaMemcpy(dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
    return e;
}

SclTables *scl_tables(const pc_plan *p) {
    std::lock_guard<std::mutex> lk(g_scl_mu);
    auto it = g_scl_tables.find(p);
    if (it != g_scl_tables.end()) return it->second;
    SclTables *T = new SclTables();
    T->a_src.assign(p->N, 0);
    T->f_src.assign(p->N, -1);
    T->node_level.assign(p->N, 0);
    T->info_src.assign(p->k > 0 ? p->k : 1, 0);
    int ii = 0, fi = 0;
    scl_build(p, *T, 0, p->n, 0, ii, fi);
    if (p->q == 2) {
        const int N = p->N, NW = N >= 32 ? N / 32 : 1;
        T->perm.assign(N, 0);
        for (int i = 0; i < N; ++i) {
            const int l = T->node_level[i], size = 1 << l, i0 = i & ~(size - 1), j = i - i0;
            int r = 0;
            for (int t = 0; t < l; ++t) r |= ((j >> t) & 1) << (l - 1 - t);
            T->perm[i] = i0 + r;
        }
        for (const SclOp &o : T->ops)
            T->ops2.push_back(make_uint2((uint32_t)o.kind | (uint32_t)o.l << 3 | (uint32_t)o.c << 7 | (uint32_t)o.i << 8,
                                         (uint32_t)(o.fv_idx & 0xffff) | (uint32_t)(o.kind == OP_REP ? o.coefw_off : 0) << 16));
        auto is_minus = [&](size_t a, int l) {
            return a < T->ops2.size() && (T->ops2[a].x & 7) == OP_MINUS && (int)((T->ops2[a].x >> 3) & 15) == l;
        };
        const char *mdv = getenv("PC_SCLW_MAXDEPTH");
        const int maxdepth = mdv && *mdv ? atoi(mdv) : 2;
        for (size_t a = 0; a < T->ops2.size(); ++a) {
            uint2 o = T->ops2[a];
            const int kind = o.x & 7, l = (o.x >> 3) & 15;
            if ((kind == OP_MINUS || kind == OP_PLUS) && maxdepth > 1) {
                if (l >= 7 && is_minus(a + 1, l - 1)) {
                    o.x |= 1u << 30;
                    a += 1;
                }
            }
            T->ops3.push_back(o);
        }
        {   // scl_path.cu: layout flags from a static walk (the list holds one path until the first forking node)
            bool single = true;
            std::vector<char> lay(p->n + 1, 0);  // 0 per-path, 1 shared (single path), 2 dual (two variants of a shared parent)
            std::vector<uint4> tmp;
            for (const SclOp &o : T->ops) {
                uint32_t x = (uint32_t)o.kind | (uint32_t)o.l << 3 | (uint32_t)o.c << 7;
                const bool chan = o.l == p->n;
                if (chan) x |= SCLP_CHAN;
                if (o.kind == OP_MINUS || o.kind == OP_PLUS) {
                    if (!chan && lay[o.l] == 1) x |= SCLP_SSRC;
                    if (!chan && lay[o.l] == 2) x |= SCLP_DSRC;
                    if (single) x |= SCLP_SDST;
                    const bool dual = o.kind == OP_PLUS && !single && (chan || lay[o.l] == 1);
                    if (dual) x |= SCLP_DUAL;
                    lay[o.l - 1] = single ? 1 : (dual ? 2 : 0);
                } else if (o.kind != OP_COMBINE) {
                    if (!chan && lay[o.l] == 1) x |= SCLP_SSRC;
                    if (!chan && lay[o.l] == 2) x |= SCLP_DSRC;
                    if (single) x |= SCLP_SDST;
                    if (o.kind != OP_RATE0) single = false;
                    ++T->n_leaf;
                } else if (single) {
                    x |= SCLP_SDST;
                }
                tmp.push_back(make_uint4(x, (uint32_t)o.i, (uint32_t)o.fv_idx, (uint32_t)(o.kind == OP_REP ? o.coefw_off : 0)));
            }
            const char *nf = getenv("PC_SCLP_NOFUSE");
            const bool fuse = !(nf && *nf && atoi(nf));
            for (size_t a = 0; a < tmp.size(); ++a) {
                uint4 o = tmp[a];
                const int kind = o.x & 7, l = (o.x >> 3) & 15;
                if (fuse && (kind == OP_MINUS || kind == OP_PLUS) && !(o.x & SCLP_DUAL) && l >= 3 && a + 1 < tmp.size() &&
                    (tmp[a + 1].x & 7) == OP_MINUS && (int)((tmp[a + 1].x >> 3) & 15) == l - 1) {
                    o.x |= SCLP_FUSED;
                    ++a;
                }
                T->opsP.push_back(o);
            }
        }
        T->stage_mask.assign((size_t)(p->n > 0 ? p->n : 1) * NW, 0u);
        for (int t = 0; t < p->n; ++t)
            for (int pos = 0; pos < N; ++pos)
                if (!(pos & (1 << t)) && T->node_level[pos] > t) T->stage_mask[(size_t)t * NW + (pos >> 5)] |= 1u << (pos & 31);
    }
    if (upload(T->d_ops, T->ops) != cudaSuccess || upload(T->d_a_src, T->a_src) != cudaSuccess ||
        upload(T->d_f_src, T->f_src) != cudaSuccess || upload(T->d_info_src, T->info_src) != cudaSuccess ||
        upload(T->d_node_level, T->node_level) != cudaSuccess || upload(T->d_rep_coef, T->rep_coef) != cudaSuccess ||
        upload(T->d_rep_coef_words, T->rep_coef_words) != cudaSuccess || upload(T->d_stage_mask, T->stage_mask) != cudaSuccess ||
        upload(T->d_perm, T->perm) != cudaSuccess || upload(T->d_ops2, T->ops2) != cudaSuccess ||
        upload(T->d_ops3, T->ops3) != cudaSuccess || upload(T->d_opsP, T->opsP) != cudaSuccess) {
        set_error("scl tables: device upload failed");
        return nullptr;
    }
    g_scl_tables[p] = T;
    return T;
}

void scl_tables_release(const pc_plan *p) {
    std::lock_guard<std::mutex> lk(g_scl_mu);
    auto it = g_scl_tables.find(p);
    if (it == g_scl_tables.end()) return;
    SclTables *T = it->second;
    cudaFree(T->d_ops), cudaFree(T->d_a_src), cudaFree(T->d_f_src), cudaFree(T->d_info_src);
    cudaFree(T->d_node_level), cudaFree(T->d_rep_coef);
    cudaFree(T->d_rep_coef_words), cudaFree(T->d_stage_mask), cudaFree(T->d_perm), cudaFree(T->d_ops2), cudaFree(T->d_ops3), cudaFree(T->d_opsP);
    delete T;
    g_scl_tables.erase(it);
}

}  // namespace pc
