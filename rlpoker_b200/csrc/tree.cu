// rlp_tree_upload: validate a flattened game, derive the device layout, copy it to HBM.
#include <algorithm>
#include <cstdlib>
#include <memory>
#include <new>
#include <numeric>

#include "tree.cuh"
#include "fused.cuh"

namespace rlp {
thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};
thread_local bool g_dry_run = false;

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load();
    if (v < 0) {
        const char *e = getenv("RLP_PDL");        // opt-in: measured no gain inside CUDA graphs (profiles/README.md)
        v = (e && e[0] == '1') ? 1 : 0;
        g_pdl.store(v);
    }
    return v == 1;
}
void pdl_disable() { g_pdl.store(0); }

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace rlp

using rlp::fail;

namespace rlp {
int build_tiles(const rlp_tree_desc *d, rlp_tree *t, const std::vector<int32_t> &colbase,
                const std::vector<int32_t> &srank, const std::vector<int32_t> &krank,
                const std::vector<int64_t> &koff_col, const std::vector<int64_t> &koff_iset, cudaStream_t stream);
}

int64_t rlp_tree::bytes() const {
    return (int64_t)(first_child.bytes() + colbase.bytes() + srank.bytes() + krank.bytes() + player.bytes() +
                     cprob.bytes() + pay1.bytes() + pay2.bytes() + node_rec.bytes() + koff_col.bytes() + koff_iset.bytes() +
                     iset_size.bytes() + iset_level.bytes() + iset_player.bytes() + col_off.bytes() +
                     iset_node_off.bytes() + iset_nodes.bytes() + d_canon_of_col.bytes()) + tiles.bytes();
}

extern "C" const char *rlp_last_error(void) { return rlp::g_error; }
extern "C" int rlp_version(void) { return 1; }
extern "C" uint64_t rlp_kernel_launches(void) { return rlp::g_launches.load(); }

extern "C" int rlp_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(RLP_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

extern "C" void rlp_cfr_free(rlp_cfr *c);
extern "C" void rlp_tree_free(rlp_tree *t) {
    if (!t) return;
    if (t->scratch) rlp_cfr_free(t->scratch);
    delete t;
}
extern "C" int64_t rlp_tree_bytes(const rlp_tree *t) { return t ? t->bytes() : 0; }

extern "C" int64_t rlp_tree_stat(const rlp_tree *t, int which) {
    if (!t) return -1;
    switch (which) {
        case 0: return t->tiles.enabled ? 1 : 0;          // tiled sweep in use (else per-level kernels)
        case 1: return t->tiles.num_tiles;                // upper tiles
        case 2: return t->tiles.num_micro;                // micro subtrees
        case 3: return t->tiles.num_classes;              // distinct micro shapes
        case 4: return (int64_t)t->tiles.jit.size();      // shapes with a compiled straight-line kernel
        case 5: return t->zero_sum ? 1 : 0;
        case 9: return t->tiles.upper_jit ? 1 : 0;        // upper tiles run as straight-line kernels
        case 6: return t->Kmax;
        case 7: return t->tiles.block;
        case 8: return (int64_t)t->tiles.smem_bytes;
        case 10: return rlp::fused_stat(t, 0);            // group-fused persistent kernel in use
        case 11: return rlp::fused_stat(t, 1);            // its micro batches per iteration
        case 12: return rlp::fused_stat(t, 2);            // its upper groups per iteration
        case 13: return rlp::fused_stat(t, 3);            // its grid (CTAs, all co-resident)
        case 14: return rlp::fused_stat(t, 4);            // public-state components (sharding unit)
        case 15: return rlp::fused_stat(t, 5);            // group-fused best-response kernel in use
        default: return -1;
    }
}

extern "C" int rlp_tree_upload(const rlp_tree_desc *d, void *stream_, rlp_tree **out) {
    if (!d || !out) return fail(RLP_ERR_INVALID, "null argument");
    *out = nullptr;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t N = d->num_nodes;
    const int32_t I = d->num_infosets;
    const int32_t L = d->num_levels;
    if (N < 1 || N >= (int64_t)INT32_MAX || I < 0 || L < 1) return fail(RLP_ERR_INVALID, "bad sizes");
    if (!d->first_child || !d->player || !d->infoset || !d->cprob || !d->u1 || !d->u2 || !d->act_off || !d->level_off)
        return fail(RLP_ERR_INVALID, "null array in rlp_tree_desc");
    if (d->level_off[0] != 0 || d->level_off[1] != 1 || d->level_off[L] != N)
        return fail(RLP_ERR_INVALID, "level_off must start at the root and end at num_nodes");
    if (d->first_child[0] != 1 && N > 1) return fail(RLP_ERR_INVALID, "first_child[0] must be 1 (breadth-first order)");

    std::unique_ptr<rlp_tree> t(new (std::nothrow) rlp_tree);
    if (!t) return fail(RLP_ERR_NOMEM, "out of host memory");
    t->N = N; t->I = I; t->L = L;
    t->level_off.assign(d->level_off, d->level_off + L + 1);
    t->act_off.assign(d->act_off, d->act_off + I + 1);
    t->Q = I ? d->act_off[I] : 0;

    // ---- validation + information-set sizes ------------------------------------------------
    std::vector<int32_t> size(I, 0), level(I, -1);
    std::vector<int8_t> owner(I, 0);
    bool zs = true;
    int32_t max_act = 0, max_chance = 0;
    for (int32_t lv = 0; lv < L; ++lv) {
        if (d->level_off[lv + 1] < d->level_off[lv]) return fail(RLP_ERR_INVALID, "level_off not monotone");
        for (int64_t i = d->level_off[lv]; i < d->level_off[lv + 1]; ++i) {
            int pl = d->player[i];
            int64_t nc = (int64_t)d->first_child[i + 1] - d->first_child[i];
            if (nc < 0 || d->first_child[i] < 1 || d->first_child[i + 1] > N)
                return fail(RLP_ERR_INVALID, "first_child out of range at node %lld", (long long)i);
            if (pl == -1) {
                if (nc != 0) return fail(RLP_ERR_INVALID, "terminal node %lld has children", (long long)i);
                if (d->u2[i] != -d->u1[i]) zs = false;
                continue;
            }
            if (pl < 0 || pl > 2) return fail(RLP_ERR_INVALID, "player %d at node %lld", pl, (long long)i);
            if (nc == 0) return fail(RLP_ERR_INVALID, "non-terminal node %lld has no children", (long long)i);
            if (d->first_child[i] < d->level_off[lv + 1] || d->first_child[i + 1] > d->level_off[std::min(lv + 2, L)])
                return fail(RLP_ERR_INVALID, "children of node %lld are not on the next level", (long long)i);
            if (pl == 0) { max_chance = std::max<int32_t>(max_chance, (int32_t)nc); continue; }
            int32_t is = d->infoset[i];
            if (is < 0 || is >= I) return fail(RLP_ERR_INVALID, "information set %d at node %lld", is, (long long)i);
            if (nc != d->act_off[is + 1] - d->act_off[is])
                return fail(RLP_ERR_INVALID, "node %lld disagrees with its information set on #actions", (long long)i);
            if (nc > rlp::kMaxActions) return fail(RLP_ERR_UNSUPPORTED, "more than %d actions", rlp::kMaxActions);
            if (size[is] == 0) { level[is] = lv; owner[is] = (int8_t)pl; }
            else if (level[is] != lv || owner[is] != pl)
                return fail(RLP_ERR_INVALID, "information set %d spans depths or players", is);
            size[is]++;
            max_act = std::max<int32_t>(max_act, (int32_t)nc);
        }
    }
    // information sets without nodes are legal: a deal-sharded rank sees only part of each set
    t->zero_sum = zs;
    t->max_act = max_act;
    t->max_chance = max_chance;

    // ---- internal order: descending size, stable ---------------------------------------------
    t->iset_of_srank.resize(I);
    std::iota(t->iset_of_srank.begin(), t->iset_of_srank.end(), 0);
    std::stable_sort(t->iset_of_srank.begin(), t->iset_of_srank.end(),
                     [&](int32_t a, int32_t b) { return size[a] > size[b]; });
    t->srank_of_iset.resize(I);
    t->col_off_h.assign(I + 1, 0);
    std::vector<int32_t> size_s(I), level_s(I);
    std::vector<int8_t> owner_s(I);
    for (int32_t s = 0; s < I; ++s) {
        int32_t is = t->iset_of_srank[s];
        t->srank_of_iset[is] = s;
        size_s[s] = size[is]; level_s[s] = level[is]; owner_s[s] = owner[is];
        t->col_off_h[s + 1] = t->col_off_h[s] + (d->act_off[is + 1] - d->act_off[is]);
    }
    t->canon_of_col.resize(t->Q);
    for (int32_t s = 0; s < I; ++s) {
        int32_t is = t->iset_of_srank[s];
        for (int64_t a = 0; a < t->col_off_h[s + 1] - t->col_off_h[s]; ++a)
            t->canon_of_col[t->col_off_h[s] + a] = d->act_off[is] + a;
    }
    t->Kmax = I ? size_s[0] : 0;
    // koff[k]: start of the k-th jagged diagonal
    std::vector<int64_t> koff_col(t->Kmax + 1, 0), koff_iset(t->Kmax + 1, 0);
    {
        int32_t alive = I;   // information sets with size > k
        for (int32_t k = 0; k < t->Kmax; ++k) {
            while (alive > 0 && size_s[alive - 1] <= k) --alive;
            koff_col[k + 1] = koff_col[k] + t->col_off_h[alive];
            koff_iset[k + 1] = koff_iset[k] + alive;
        }
    }
    t->NC = koff_col[t->Kmax];
    t->ND = koff_iset[t->Kmax];

    // ---- per-node tables -----------------------------------------------------------------------
    std::vector<int32_t> colbase(N, -1), srank(N, -1), krank(N, -1), seen(I, 0);
    std::vector<int64_t> node_off(I + 1, 0);
    for (int32_t s = 0; s < I; ++s) node_off[s + 1] = node_off[s] + size_s[s];
    std::vector<int32_t> nodes(t->ND);
    for (int64_t i = 0; i < N; ++i) {
        if (d->player[i] <= 0) continue;
        int32_t s = t->srank_of_iset[d->infoset[i]];
        srank[i] = s;
        colbase[i] = (int32_t)t->col_off_h[s];
        krank[i] = seen[s];
        nodes[node_off[s] + seen[s]] = (int32_t)i;
        seen[s]++;
    }
    // leading levels one block can sweep
    int32_t top = 0;
    while (top < L && d->level_off[top + 1] - d->level_off[top] <= 4096) ++top;
    t->top_levels = top;

    // ---- upload ----------------------------------------------------------------------------------
    std::vector<int32_t> fc(d->first_child, d->first_child + N + 1);
    std::vector<int8_t> pl(d->player, d->player + N);
    std::vector<double> cp(d->cprob, d->cprob + N), p1(d->u1, d->u1 + N);
    RLP_CUDA(t->first_child.upload(fc, stream));
    RLP_CUDA(t->player.upload(pl, stream));
    RLP_CUDA(t->colbase.upload(colbase, stream));
    RLP_CUDA(t->srank.upload(srank, stream));
    RLP_CUDA(t->krank.upload(krank, stream));
    {
        std::vector<int4> rec((size_t)N);
        for (int64_t i = 0; i < N; ++i) {
            uint32_t leaf = 0;          // bit a: child a is a leaf (first 16 children; wider nodes look the child up)
            for (int a = 0; a < fc[i + 1] - fc[i] && a < 16; ++a)
                if (pl[fc[i] + a] < 0) leaf |= 1u << a;
            rec[i] = make_int4(fc[i], (int)((uint32_t)(fc[i + 1] - fc[i]) | ((uint32_t)(pl[i] + 1) << 8) | (leaf << 16)),
                               colbase[i], srank[i]);
        }
        RLP_CUDA(t->node_rec.upload(rec, stream));
        RLP_CUDA(rlp::sync_stream(stream));
    }
    t->first_child_h = fc;
    t->player_h = pl;
    RLP_CUDA(t->cprob.upload(cp, stream));
    RLP_CUDA(t->pay1.upload(p1, stream));
    if (!zs) {
        std::vector<double> p2(d->u2, d->u2 + N);
        RLP_CUDA(t->pay2.upload(p2, stream));
        RLP_CUDA(rlp::sync_stream(stream));
    }
    RLP_CUDA(t->koff_col.upload(koff_col, stream));
    RLP_CUDA(t->koff_iset.upload(koff_iset, stream));
    RLP_CUDA(t->iset_size.upload(size_s, stream));
    RLP_CUDA(t->iset_level.upload(level_s, stream));
    RLP_CUDA(t->iset_player.upload(owner_s, stream));
    t->iset_player_h = owner_s;
    RLP_CUDA(t->col_off.upload(t->col_off_h, stream));
    RLP_CUDA(t->iset_node_off.upload(node_off, stream));
    RLP_CUDA(t->iset_nodes.upload(nodes, stream));
    RLP_CUDA(t->d_canon_of_col.upload(t->canon_of_col, stream));
    RLP_CUDA(rlp::sync_stream(stream));     // host vectors die at scope exit
    {
        const char *no_tiles = getenv("RLP_NO_TILES");
        if (!(no_tiles && no_tiles[0] == '1')) {
            int rc = rlp::build_tiles(d, t.get(), colbase, srank, krank, koff_col, koff_iset, stream);
            if (rc) return rc;
        }
    }
    *out = t.release();
    return RLP_OK;
}

// Host-only planning: everything rlp_tree_upload derives (validation, information-set layout, two-tier tiling, shape
// classes, generated kernel sources compiled with NVRTC) without a GPU.  Lets the CPU test-suite cover the host logic.
namespace rlp { extern thread_local int g_jit_compiled, g_jit_failed; thread_local std::string *g_plan_source_out = nullptr; }

// Planning mode plus the generated source of the group-fused kernel (empty when the tree does not qualify): the
// number of bytes needed (incl. the terminating NUL) is returned in *len; buf (nullable) receives up to cap bytes.
extern "C" int rlp_tree_plan_fused_source(const rlp_tree_desc *d, int64_t *stats16, char *buf, int64_t cap, int64_t *len) {
    std::string src;
    rlp::g_plan_source_out = &src;
    int64_t tmp[16];
    int rc = rlp_tree_plan(d, stats16 ? stats16 : tmp);
    rlp::g_plan_source_out = nullptr;
    if (rc) return rc;
    if (len) *len = (int64_t)src.size() + 1;
    if (buf && cap > 0) {
        size_t n = std::min<size_t>((size_t)cap - 1, src.size());
        memcpy(buf, src.data(), n);
        buf[n] = 0;
    }
    return RLP_OK;
}

namespace rlp { const std::vector<int64_t> &lane_record_bounds(rlp_tree *t, int variant); }

// Host-only: the per-level record capacities of the level-by-level sampled lanes for ONE lane (both traversals), i.e.
// out[e] = most records with e explored ancestors; zero-padded to 33 entries.
extern "C" int rlp_tree_plan_lane_records(const rlp_tree_desc *d, int variant, int64_t *out33) {
    if (!d || !out33 || (variant != 0 && variant != 1)) return fail(RLP_ERR_INVALID, "bad argument");
    rlp::g_dry_run = true;
    rlp_tree *t = nullptr;
    int rc = rlp_tree_upload(d, nullptr, &t);
    rlp::g_dry_run = false;
    if (rc) return rc;
    const std::vector<int64_t> &w = rlp::lane_record_bounds(t, variant);
    for (int k = 0; k < 33; ++k) out33[k] = k < (int)w.size() ? w[k] : 0;
    delete t;
    return RLP_OK;
}

extern "C" int rlp_tree_plan(const rlp_tree_desc *d, int64_t *stats16);
extern "C" int rlp_tree_plan(const rlp_tree_desc *d, int64_t *stats16) {
    if (!d || !stats16) return fail(RLP_ERR_INVALID, "null argument");
    rlp::g_dry_run = true;
    rlp::g_jit_compiled = rlp::g_jit_failed = 0;
    rlp_tree *t = nullptr;
    int rc = rlp_tree_upload(d, nullptr, &t);
    rlp::g_dry_run = false;
    if (rc) return rc;
    const TileSet &ts = t->tiles;
    if (rlp::g_plan_source_out) {
        const std::string *src = rlp::fused_source(t);
        *rlp::g_plan_source_out = src ? *src : std::string();
    }
    int64_t v[16] = {ts.enabled ? 1 : 0, ts.num_tiles, ts.num_micro, ts.num_classes, ts.micro_nodes, ts.micro_nt,
                     ts.micro_term, ts.acc_a_count, ts.acc_b_count, ts.max_nodes, ts.max_nt, (int64_t)ts.smem_bytes,
                     rlp::g_jit_compiled, rlp::g_jit_failed, t->Kmax, t->zero_sum ? 1 : 0};
    for (int k = 0; k < 16; ++k) stats16[k] = v[k];
    delete t;
    return RLP_OK;
}
