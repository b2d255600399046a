// Tiled sweep: host-side tiling of a flattened game + the kernel that sweeps one tile per CTA in shared
// memory (forward reach, backward counterfactual values, regret / own-reach contributions).
// Same arithmetic, operand for operand, as the level sweep in cfr.cu (and therefore as the reference's
// cfr_recursive, rlpoker/cfr/cfr.py:158-255); only where the intermediates live changes.
#include <algorithm>

#include "cfr.cuh"

using rlp::fail;

namespace {

constexpr size_t kSmemSmall = 110 * 1024;    // two CTAs per SM
constexpr size_t kSmemLarge = 220 * 1024;    // one CTA per SM

size_t tile_smem(int64_t n_nodes, int64_t n_nt, int64_t n_cp, bool zs) {
    size_t b = 0;
    b += sizeof(double) * (size_t)n_nodes * (zs ? 1 : 2);   // val1 (val2)
    b += sizeof(double) * 3 * (size_t)n_nt;                 // p1, p2, static chance reach
    b += sizeof(double) * (size_t)n_cp;                     // chance probabilities
    b += (sizeof(TileRec) + sizeof(TileAux)) * (size_t)n_nt;
    b += sizeof(unsigned short) * (size_t)n_nodes;
    return b + 64;
}

// ------------------------------------------------------------------------------------ kernel
struct TileView {
    const TileDesc *desc;
    const unsigned short *ntof, *term_lid;
    const TileRec *rec;
    const TileAux *aux;
    const double *pc, *pay1, *pay2, *cprob, *sigma;
    double *contrib_r, *contrib_o, *root_val;
    const IterState *iter;
    int max_nodes, max_nt, max_cp;
};

constexpr int kRegAct = 4;      // information sets up to this wide keep their strategy row in registers

// One CTA sweeps one tile.  Ownership is static: non-terminal rank i belongs to thread i % blockDim, slot
// i / blockDim (< R), so the strategy rows of a thread's nodes are fetched ONCE into registers while the
// topology is being staged, and both passes then run out of shared memory and registers only -- no global
// load sits on the level-to-level dependency chain.
template <bool ZS, int R, bool WIDE>
__global__ void __launch_bounds__(1024) k_tile_sweep(TileView v) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileDesc td;
    double *val1 = reinterpret_cast<double *>(smem_raw);
    double *val2 = val1 + (ZS ? 0 : v.max_nodes);
    double *sp1 = val2 + v.max_nodes;
    double *sp2 = sp1 + v.max_nt;
    double *spc = sp2 + v.max_nt;
    double *scp = spc + v.max_nt;
    TileAux *saux = reinterpret_cast<TileAux *>(scp + v.max_cp);
    TileRec *rec = reinterpret_cast<TileRec *>(saux + v.max_nt);
    unsigned short *ntof = reinterpret_cast<unsigned short *>(rec + v.max_nt);

    const int tid = threadIdx.x, bd = blockDim.x;
    if (tid < (int)(sizeof(TileDesc) / sizeof(int)))
        reinterpret_cast<int *>(&td)[tid] = reinterpret_cast<const int *>(v.desc + blockIdx.x)[tid];
    __syncthreads();
    const double w = v.iter->linear ? (double)v.iter->t : 1.0;
    const int n_nt = td.n_nt;

    // ---- stage the tile; the owner threads fetch their strategy rows as soon as they know where -------
    double sg[R][kRegAct];
    {
        const TileAux *gaux = v.aux + td.nt_off;
        const TileRec *grec = v.rec + td.nt_off;
        const double *gpc = v.pc + td.nt_off;
#pragma unroll
        for (int m = 0; m < R; ++m) {
            int i = m * bd + tid;
            if (i < n_nt) {
                TileAux ax = gaux[i];
                TileRec r = grec[i];
                saux[i] = ax;
                rec[i] = r;
                spc[i] = gpc[i];
                if (!WIDE && r.player > 0) {
                    const double *row = v.sigma + ax.colbase;
#pragma unroll
                    for (int k = 0; k < kRegAct; ++k) sg[m][k] = k < r.nchild ? row[k] : 0.0;
                }
            }
        }
        const unsigned short *gnt = v.ntof + td.node_off;
        for (int i = tid; i < td.n_nodes; i += bd) ntof[i] = gnt[i];
        const unsigned short *glid = v.term_lid + td.term_off;
        const double *gp1 = v.pay1 + td.term_off;
        for (int i = tid; i < td.n_term; i += bd) {
            val1[glid[i]] = gp1[i];
            if (!ZS) val2[glid[i]] = v.pay2[td.term_off + i];
        }
        const double *gcp = v.cprob + td.cp_off;
        for (int i = tid; i < td.n_cp; i += bd) scp[i] = gcp[i];
        if (tid == 0) { sp1[0] = td.root_p1; sp2[0] = td.root_p2; }
    }
    __syncthreads();

    // ---- forward: strategy reach of every non-terminal node (cfr.py:208,217) ----------------------
    for (int l = 0; l + 1 < td.n_levels; ++l) {
        const int lo = td.lvl_nt[l], hi = td.lvl_nt[l + 1];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int i = m * bd + tid;
            if (i < lo || i >= hi) continue;
            TileRec r = rec[i];
            double a = sp1[i], b = sp2[i];
            if (r.player == 0) {
                for (int k = 0; k < r.nchild; ++k) {
                    unsigned short c = ntof[r.first_child + k];
                    if (c != 0xFFFF) { sp1[c] = a; sp2[c] = b; }
                }
            } else if constexpr (WIDE) {
                const double *row = v.sigma + saux[i].colbase;
                for (int k = 0; k < r.nchild; ++k) {
                    unsigned short c = ntof[r.first_child + k];
                    if (c == 0xFFFF) continue;
                    double e = row[k];
                    sp1[c] = r.player == 1 ? e * a : a;
                    sp2[c] = r.player == 2 ? e * b : b;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kRegAct; ++k) {
                    if (k >= r.nchild) break;
                    unsigned short c = ntof[r.first_child + k];
                    if (c == 0xFFFF) continue;
                    double e = sg[m][k];
                    sp1[c] = r.player == 1 ? e * a : a;
                    sp2[c] = r.player == 2 ? e * b : b;
                }
            }
        }
        __syncthreads();
    }

    // ---- backward: values, and the contributions of every decision node (cfr.py:177-186,204-242) ---
    for (int l = td.n_levels - 1; l >= 0; --l) {
        const int lo = td.lvl_nt[l], hi = td.lvl_nt[l + 1];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int i = m * bd + tid;
            if (i < lo || i >= hi) continue;
            TileRec r = rec[i];
            const int fc = r.first_child, n = r.nchild, pl = r.player;
            double a1 = 0.0, a2 = 0.0;
            if (pl == 0) {
                const double *e = scp + r.cp_off;
                for (int k = 0; k < n; ++k) {
                    a1 = a1 + e[k] * val1[fc + k];
                    if (!ZS) a2 = a2 + e[k] * val2[fc + k];
                }
                val1[r.lid] = a1;
                if (!ZS) val2[r.lid] = a2;
                continue;
            }
            TileAux ax = saux[i];
            double va[WIDE ? 1 : kRegAct];
            if constexpr (WIDE) {
                const double *row = v.sigma + ax.colbase;
                for (int k = 0; k < n; ++k) {
                    a1 = a1 + row[k] * val1[fc + k];
                    if (!ZS) a2 = a2 + row[k] * val2[fc + k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < kRegAct; ++k) {
                    if (k < n) {
                        double x = val1[fc + k];
                        a1 = a1 + sg[m][k] * x;
                        if (!ZS) {
                            double y = val2[fc + k];
                            a2 = a2 + sg[m][k] * y;
                            va[k] = pl == 1 ? x : y;
                        } else {
                            va[k] = pl == 1 ? x : -x;
                        }
                    }
                }
            }
            val1[r.lid] = a1;
            if (!ZS) val2[r.lid] = a2;
            double c = spc[i], x1 = sp1[i], x2 = sp2[i];
            double opp = pl == 1 ? c * x2 : c * x1;
            double own = pl == 1 ? x1 : x2;
            double vp = pl == 1 ? a1 : (ZS ? -a1 : a2);
            double *dst = v.contrib_r + ax.slot_r;
            if constexpr (WIDE) {
                for (int k = 0; k < n; ++k) {
                    double x = pl == 1 ? val1[fc + k] : (ZS ? -val1[fc + k] : val2[fc + k]);
                    dst[k] = w * (x - vp) * opp;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kRegAct; ++k)
                    if (k < n) dst[k] = w * (va[k] - vp) * opp;
            }
            v.contrib_o[ax.slot_o] = w * own;
        }
        __syncthreads();
    }
    if (tid == 0) v.root_val[blockIdx.x] = val1[0];
}

}  // namespace

// ------------------------------------------------------------------------------------ host
namespace rlp {

// Called by rlp_tree_upload once the per-node information-set tables exist.  Leaves t->tiles disabled
// (level-sweep fallback) when the tree cannot be tiled under an all-chance trunk.
int build_tiles(const rlp_tree_desc *d, rlp_tree *t, const std::vector<int32_t> &colbase,
                const std::vector<int32_t> &srank, const std::vector<int32_t> &krank,
                const std::vector<int64_t> &koff_col, const std::vector<int64_t> &koff_iset, cudaStream_t stream) {
    const int64_t N = d->num_nodes;
    const int32_t *fc = d->first_child;
    const int8_t *pl = d->player;
    TileSet &ts = t->tiles;
    ts.enabled = false;
    if (N < 2) return RLP_OK;
    const bool zs = t->zero_sum;

    std::vector<int32_t> parent(N, -1);
    for (int64_t i = 0; i < N; ++i)
        for (int32_t c = fc[i]; c < fc[i + 1]; ++c) parent[c] = (int32_t)i;
    std::vector<int32_t> cn(N, 1), cnt(N, 0), cce(N, 0);
    std::vector<uint8_t> height(N, 0);
    bool wide = false;
    for (int64_t i = N - 1; i >= 0; --i) {
        if (pl[i] >= 0) cnt[i] += 1;
        int nc = fc[i + 1] - fc[i];
        if (nc > 255) wide = true;
        if (pl[i] == 0) cce[i] += nc;
        if (i > 0) {
            int32_t p = parent[i];
            cn[p] += cn[i]; cnt[p] += cnt[i]; cce[p] += cce[i];
            height[p] = std::max<int>(height[p], std::min(254, height[i] + 1));
        }
    }
    if (wide) return RLP_OK;

    std::vector<uint8_t> status(N);        // 0 trunk, 1 tile root, 2 inside a tile
    std::vector<int32_t> roots;
    bool ok = false;
    size_t budget_used = 0;
    for (size_t budget : {kSmemSmall, kSmemLarge}) {
        auto fits = [&](int64_t v) {
            return pl[v] >= 0 && cn[v] <= 65535 && cce[v] <= 65535 && height[v] + 1 <= kMaxTileLevels &&
                   tile_smem(cn[v], cnt[v], cce[v], zs) <= budget;
        };
        roots.clear();
        ok = true;
        for (int64_t v = 0; v < N && ok; ++v) {
            uint8_t ps = v == 0 ? 0 : status[parent[v]];
            if (ps != 0) { status[v] = 2; continue; }
            if (fits(v)) { status[v] = 1; roots.push_back((int32_t)v); continue; }
            status[v] = 0;
            if (pl[v] > 0) ok = false;            // a decision node above the cut: not handled by the tiled path
        }
        if (ok && !roots.empty()) { budget_used = budget; break; }
        ok = false;
    }
    if (!ok) return RLP_OK;

    // static chance reach of every node (pi_c does not depend on the strategy): cp * pi_c, top-down
    std::vector<double> pcs(N);
    pcs[0] = 1.0;
    for (int64_t i = 0; i < N; ++i)
        for (int32_t c = fc[i]; c < fc[i + 1]; ++c) pcs[c] = pl[i] == 0 ? d->cprob[c] * pcs[i] : pcs[i];

    const size_t T = roots.size();
    std::vector<TileDesc> desc(T);
    int64_t tot_nodes = 0, tot_nt = 0, tot_term = 0, tot_cp = 0;
    for (size_t k = 0; k < T; ++k) {
        int32_t r = roots[k];
        tot_nodes += cn[r]; tot_nt += cnt[r]; tot_term += cn[r] - cnt[r]; tot_cp += cce[r];
    }
    std::vector<unsigned short> ntof(tot_nodes), term_lid(tot_term);
    std::vector<TileRec> rec(tot_nt);
    std::vector<TileAux> aux(tot_nt);
    std::vector<double> pc(tot_nt), pay1(tot_term), pay2(zs ? 0 : tot_term), cprob(tot_cp);
    std::vector<int32_t> local;          // local id -> global id
    std::vector<int32_t> ldepth;
    int64_t o_node = 0, o_nt = 0, o_term = 0, o_cp = 0;
    int max_nodes = 0, max_nt = 0, max_cp = 0;
    for (size_t k = 0; k < T; ++k) {
        const int32_t r = roots[k];
        TileDesc &td = desc[k];
        memset(&td, 0, sizeof(td));
        td.node_off = o_node; td.nt_off = o_nt; td.term_off = o_term; td.cp_off = o_cp;
        td.n_nodes = cn[r]; td.n_nt = cnt[r]; td.n_term = cn[r] - cnt[r];
        td.root_global = r; td.root_p1 = 1.0; td.root_p2 = 1.0;
        local.clear(); ldepth.clear();
        local.push_back(r); ldepth.push_back(0);
        for (size_t q = 0; q < local.size(); ++q) {
            int32_t g = local[q];
            for (int32_t c = fc[g]; c < fc[g + 1]; ++c) { local.push_back(c); ldepth.push_back(ldepth[q] + 1); }
        }
        // children of local node q start at local id first_child_local: running count in BFS order
        int nt = 0, term = 0, cpo = 0, next_child = 1, level = -1;
        for (size_t q = 0; q < local.size(); ++q) {
            int32_t g = local[q];
            if (pl[g] < 0) {
                ntof[o_node + q] = 0xFFFF;
                term_lid[o_term + term] = (unsigned short)q;
                pay1[o_term + term] = d->u1[g];
                if (!zs) pay2[o_term + term] = d->u2[g];
                ++term;
                continue;
            }
            while (level < ldepth[q]) td.lvl_nt[++level] = nt;
            ntof[o_node + q] = (unsigned short)nt;
            int nc = fc[g + 1] - fc[g];
            TileRec &rc = rec[o_nt + nt];
            rc.first_child = (unsigned short)next_child;
            rc.lid = (unsigned short)q;
            rc.nchild = (unsigned char)nc;
            rc.player = (unsigned char)pl[g];
            rc.cp_off = 0;
            TileAux &ax = aux[o_nt + nt];
            ax.colbase = ax.slot_r = ax.slot_o = ax.pad = 0;
            if (pl[g] == 0) {
                rc.cp_off = (unsigned short)cpo;
                for (int32_t c = fc[g]; c < fc[g + 1]; ++c) cprob[o_cp + cpo++] = d->cprob[c];
            } else {
                ax.colbase = colbase[g];
                ax.slot_r = (int)(koff_col[krank[g]] + colbase[g]);
                ax.slot_o = (int)(koff_iset[krank[g]] + srank[g]);
            }
            pc[o_nt + nt] = pcs[g];
            next_child += nc;
            ++nt;
        }
        td.n_levels = level + 1;
        for (int l = level + 1; l <= kMaxTileLevels; ++l) td.lvl_nt[l] = nt;
        max_nodes = std::max(max_nodes, td.n_nodes);
        max_nt = std::max(max_nt, td.n_nt);
        max_cp = std::max(max_cp, cpo);
        td.n_cp = cpo;
        o_node += td.n_nodes; o_nt += td.n_nt; o_term += td.n_term; o_cp += cpo;
    }
    if (koff_col.back() >= (int64_t)INT32_MAX) return RLP_OK;     // slots are 32-bit

    ts.num_tiles = (int)T;
    ts.max_nodes = max_nodes;
    ts.max_nt = max_nt;
    ts.max_cp = max_cp;
    ts.smem_bytes = tile_smem(max_nodes, max_nt, max_cp, zs);
    if (ts.smem_bytes > 226 * 1024) return RLP_OK;               // (max_nodes, max_nt) came from different tiles
    (void)budget_used;
    // smallest block that keeps every thread at <= 2 owned nodes, else 1024 threads with up to 4
    ts.block = 1024;
    for (int b : {128, 256, 512}) if ((max_nt + b - 1) / b <= 2) { ts.block = b; break; }
    ts.slots = (max_nt + ts.block - 1) / ts.block;
    if (ts.slots > 4) return RLP_OK;
    ts.wide = t->max_act > kRegAct;
    RLP_CUDA(ts.desc.upload(desc, stream));
    RLP_CUDA(ts.ntof.upload(ntof, stream));
    RLP_CUDA(ts.term_lid.upload(term_lid, stream));
    RLP_CUDA(ts.rec.upload(rec, stream));
    RLP_CUDA(ts.aux.upload(aux, stream));
    RLP_CUDA(ts.pc.upload(pc, stream));
    RLP_CUDA(ts.pay1.upload(pay1, stream));
    RLP_CUDA(ts.pay2.upload(pay2, stream));
    RLP_CUDA(ts.cprob.upload(cprob, stream));
    RLP_CUDA(ts.root_val.alloc(T));
    RLP_CUDA(cudaStreamSynchronize(stream));
    ts.enabled = true;
    return RLP_OK;
}

// One tiled sweep (replaces the per-level forward/backward launches of enqueue_sweep).
template <bool ZS, int R, bool WIDE>
static int launch_variant(const TileSet &ts, const TileView &v, cudaStream_t s) {
    static thread_local bool attr_set = false;
    if (!attr_set) {
        RLP_CUDA(cudaFuncSetAttribute(k_tile_sweep<ZS, R, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_set = true;
    }
    k_tile_sweep<ZS, R, WIDE><<<ts.num_tiles, ts.block, ts.smem_bytes, s>>>(v);
    RLP_LAUNCH_CHECK();
    return RLP_OK;
}

template <bool ZS, bool WIDE>
static int launch_slots(const TileSet &ts, const TileView &v, cudaStream_t s) {
    switch (ts.slots) {
        case 1: return launch_variant<ZS, 1, WIDE>(ts, v, s);
        case 2: return launch_variant<ZS, 2, WIDE>(ts, v, s);
        case 3: return launch_variant<ZS, 3, WIDE>(ts, v, s);
        default: return launch_variant<ZS, 4, WIDE>(ts, v, s);
    }
}

int launch_tile_sweep(rlp_cfr *c, cudaStream_t s, const IterState *iter) {
    rlp_tree *t = c->tree;
    TileSet &ts = t->tiles;
    TileView v;
    v.desc = ts.desc.p; v.ntof = ts.ntof.p; v.term_lid = ts.term_lid.p; v.rec = ts.rec.p; v.aux = ts.aux.p;
    v.pc = ts.pc.p; v.pay1 = ts.pay1.p; v.pay2 = ts.pay2.p; v.cprob = ts.cprob.p; v.sigma = c->sigma.p;
    v.contrib_r = c->contrib_r.p; v.contrib_o = c->contrib_o.p; v.root_val = ts.root_val.p;
    v.iter = iter ? iter : c->iter.p; v.max_nodes = ts.max_nodes; v.max_nt = ts.max_nt; v.max_cp = ts.max_cp;
    if (t->zero_sum) return ts.wide ? launch_slots<true, true>(ts, v, s) : launch_slots<true, false>(ts, v, s);
    return ts.wide ? launch_slots<false, true>(ts, v, s) : launch_slots<false, false>(ts, v, s);
}

}  // namespace rlp
