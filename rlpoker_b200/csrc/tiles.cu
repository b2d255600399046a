// Tiled sweep: host-side two-tier tiling of a flattened game (see tiles.cuh) and the kernels that sweep it.
// Same arithmetic, operand for operand, as the level sweep in cfr.cu (and therefore as the reference's
// cfr_recursive, rlpoker/cfr/cfr.py:158-255); only where the intermediates live changes.
#include <algorithm>
#include <map>
#include <string>

#include "cfr.cuh"
#include "fused.cuh"

using rlp::fail;

namespace {

constexpr size_t kSmemSmall = 48 * 1024;     // several CTAs per SM
constexpr size_t kSmemMid = 110 * 1024;      // two CTAs per SM
constexpr size_t kSmemLarge = 220 * 1024;    // one CTA per SM
constexpr int kRegAct = 4;      // information sets up to this wide keep their strategy row in registers

size_t tile_smem(int64_t n_nodes, int64_t n_nt, int64_t n_cp, bool zs) {
    size_t b = 0;
    b += sizeof(double) * (size_t)n_nodes * (zs ? 1 : 2);   // val1 (val2)
    b += sizeof(double) * 3 * (size_t)n_nt;                 // p1, p2, static chance reach
    b += sizeof(double) * (size_t)n_cp;                     // chance probabilities
    b += (sizeof(TileRec) + sizeof(TileAux)) * (size_t)n_nt;
    b += sizeof(unsigned short) * (size_t)n_nodes;
    return b + 64;
}

// ------------------------------------------------------------------------------------ upper tiles
struct TileView {
    const TileDesc *desc;
    const unsigned short *ntof, *term_lid, *portal_lid;
    const TileRec *rec;
    const TileAux *aux;
    const double *pc, *pay1, *pay2, *cprob, *sigma;
    double *contrib_r, *contrib_o, *root_val;
    double *portal_p1, *portal_p2;
    const double *portal_v1, *portal_v2;
    const IterState *iter;
    int max_nodes, max_nt, max_cp;
    long long nd;           // contrib_r is action-major [action][nd], indexed by the contrib_o slot
};

// One CTA sweeps one upper tile.  Ownership is static: non-terminal rank i belongs to thread i % blockDim, slot
// i / blockDim (< R), so the strategy rows of a thread's nodes are fetched ONCE into registers while the
// topology is being staged, and the passes then run out of shared memory and registers only.
// PHASE 0: forward only, reach written to the portal leaves.  PHASE 1: forward + backward, portal leaves read
// the values the micro kernel left; contributions of the tile's decision nodes are scattered.
template <bool ZS, int R, bool WIDE, int PHASE>
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
    for (int i = tid; i < (int)(sizeof(TileDesc) / sizeof(int)); i += bd)
        reinterpret_cast<int *>(&td)[i] = reinterpret_cast<const int *>(v.desc + blockIdx.x)[i];
    __syncthreads();
    rlp::pdl_launch_dependents();
    const int n_nt = td.n_nt;
    const bool idle = PHASE == 0 && td.n_portal == 0;

    // ---- stage the static part of the tile (topology, payoffs) while the previous kernel may still run ------
    double sg[R][kRegAct];
    TileAux my_aux[R];
    if (!idle) {
        const TileAux *gaux = v.aux + td.nt_off;
        const TileRec *grec = v.rec + td.nt_off;
        const double *gpc = v.pc + td.nt_off;
#pragma unroll
        for (int m = 0; m < R; ++m) {
            int i = m * bd + tid;
            if (i < n_nt) {
                my_aux[m] = gaux[i];
                saux[i] = my_aux[m];
                rec[i] = grec[i];
                spc[i] = gpc[i];
            }
        }
        const unsigned short *gnt = v.ntof + td.node_off;
        for (int i = tid; i < td.n_nodes; i += bd) ntof[i] = gnt[i];
        if (PHASE == 1) {
            const unsigned short *glid = v.term_lid + td.term_off;
            const double *gp1 = v.pay1 + td.term_off;
            for (int i = tid; i < td.n_term; i += bd) {
                val1[glid[i]] = gp1[i];
                if (!ZS) val2[glid[i]] = v.pay2[td.term_off + i];
            }
            const double *gcp = v.cprob + td.cp_off;
            for (int i = tid; i < td.n_cp; i += bd) scp[i] = gcp[i];
        }
        if (tid == 0) { sp1[0] = td.root_p1; sp2[0] = td.root_p2; }
    }
    // ---- everything below reads what the previous kernels wrote (strategy, portal values, iteration counter) ---
    rlp::pdl_wait();
    if (idle) return;
    const int tl_slot = rlp_tl_begin(v.iter, PHASE == 0 ? 0 : 2);
    const double w = v.iter->linear ? (double)v.iter->t : 1.0;
#pragma unroll
    for (int m = 0; m < R; ++m) {
        int i = m * bd + tid;
        if (!WIDE && i < n_nt) {
            TileRec r = rec[i];
            if (r.player > 0) {
                const double *row = v.sigma + my_aux[m].colbase;
#pragma unroll
                for (int k = 0; k < kRegAct; ++k) sg[m][k] = k < r.nchild ? row[k] : 0.0;
            }
        }
    }
    if (PHASE == 1) {
        const unsigned short *plid = v.portal_lid + td.portal_off;
        for (int i = tid; i < td.n_portal; i += bd) {
            val1[plid[i]] = v.portal_v1[(long long)i * gridDim.x + blockIdx.x];
            if (!ZS) val2[plid[i]] = v.portal_v2[(long long)i * gridDim.x + blockIdx.x];
        }
    }
    __syncthreads();

    // ---- forward: strategy reach of every non-terminal node (cfr.py:208,217) ----------------------
    for (int l = 0; l < td.n_levels; ++l) {
        const int lo = td.lvl_nt[l], hi = td.lvl_nt[l + 1];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int i = m * bd + tid;
            if (i < lo || i >= hi) continue;
            TileRec r = rec[i];
            const double a = sp1[i], b = sp2[i];
            const double *row = (WIDE && r.player > 0) ? v.sigma + saux[i].colbase : nullptr;
            auto put = [&](int k, double e) {
                unsigned short c = ntof[r.first_child + k];
                if (c == 0xFFFF) return;
                double x = r.player == 1 ? e * a : a;
                double y = r.player == 2 ? e * b : b;
                if (c & 0x8000) {
                    if (PHASE == 0) {
                        long long g = (long long)(c & 0x7FFF) * gridDim.x + blockIdx.x;     // [portal rank][tile]
                        v.portal_p1[g] = x;
                        v.portal_p2[g] = y;
                    }
                } else {
                    sp1[c] = x;
                    sp2[c] = y;
                }
            };
            if (r.player == 0) {
                for (int k = 0; k < r.nchild; ++k) put(k, 1.0);
            } else if constexpr (WIDE) {
                for (int k = 0; k < r.nchild; ++k) put(k, row[k]);
            } else {
#pragma unroll
                for (int k = 0; k < kRegAct; ++k)
                    if (k < r.nchild) put(k, sg[m][k]);
            }
        }
        __syncthreads();
    }
    if (PHASE == 0) { rlp_tl_end(v.iter, tl_slot); return; }

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
            double *dst = v.contrib_r + ax.slot_o;
            if constexpr (WIDE) {
                for (int k = 0; k < n; ++k) {
                    double x = pl == 1 ? val1[fc + k] : (ZS ? -val1[fc + k] : val2[fc + k]);
                    dst[(long long)k * v.nd] = w * (x - vp) * opp;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kRegAct; ++k)
                    if (k < n) dst[(long long)k * v.nd] = w * (va[k] - vp) * opp;
            }
            v.contrib_o[ax.slot_o] = w * own;
        }
        __syncthreads();
    }
    if (tid == 0) v.root_val[blockIdx.x] = val1[0];
    rlp_tl_end(v.iter, tl_slot);
}

// ------------------------------------------------------------------------------------ micro subtrees

constexpr int kMicroBlock = 64;

// One thread = one micro subtree.  Everything the sweep needs from HBM/L2 -- information-set columns,
// contribution slots, the strategy probability of every edge, the payoff of every leaf -- is fetched up front in
// two batches of independent loads into a thread-interleaved (bank-conflict-free) slice of shared memory; the
// forward and backward passes over the <= 8 decision nodes then touch shared memory only, so no global load
// sits on the node-to-node dependency chain.
struct MicroView {
    const MicroTmpl *tmpl;
    const unsigned short *cls;
    const int *colbase, *slot_r, *slot_o;      // [micro_nt][P]
    const int *pidx;                           // [P] where this instance's portal lives ([portal rank][tile] layout)
    const int *ridx;                           // [P] where its root reach is (its chance parent's slot, if any)
    const double *pay1, *pay2;                 // [micro_term][P]
    const double *pc, *sigma;               // sigma here is the ACTION-MAJOR copy [action][internal set]
    const double *portal_p1, *portal_p2;
    double *portal_v1, *portal_v2;
    double *contrib_r, *contrib_o;
    const IterState *iter;
    int P, num_classes, mn, mnt;
    long long nd, num_isets;
};

template <bool ZS>
__global__ void __launch_bounds__(kMicroBlock) k_micro(MicroView v) {
    extern __shared__ __align__(16) unsigned char micro_smem[];
    __shared__ MicroTmpl stm[kMicroSmemClasses];
    const int tid = threadIdx.x;
    const int mn = v.mn, mnt = v.mnt;
    double *ssig = reinterpret_cast<double *>(micro_smem);            // [mn][bd]  strategy prob of the edge INTO node j
    double *sv1 = ssig + mn * kMicroBlock;                            // [mn][bd]  value of node j (player 1)
    double *sv2 = sv1 + mn * kMicroBlock;                             // [mn][bd]  (non-zero-sum only)
    double *sq1 = sv2 + (ZS ? 0 : mn * kMicroBlock);                  // [mnt][bd] reach by non-terminal rank
    double *sq2 = sq1 + mnt * kMicroBlock;
    int *scb = reinterpret_cast<int *>(sq2 + mnt * kMicroBlock);      // [mnt][bd] information-set column base
    int *ssr = scb + mnt * kMicroBlock;                               // [mnt][bd] slot in contrib_r
    int *sso = ssr + mnt * kMicroBlock;                               // [mnt][bd] slot in contrib_o
    const bool in_smem = v.num_classes <= kMicroSmemClasses;
    if (in_smem) {
        const int words = v.num_classes * (int)(sizeof(MicroTmpl) / sizeof(int));
        for (int i = tid; i < words; i += kMicroBlock)
            reinterpret_cast<int *>(stm)[i] = reinterpret_cast<const int *>(v.tmpl)[i];
        __syncthreads();
    }
    rlp::pdl_launch_dependents();
    const long long p = (long long)blockIdx.x * kMicroBlock + tid;
    if (p >= v.P) { rlp::pdl_wait(); return; }
    const long long P = v.P;
    const MicroTmpl *tm = in_smem ? &stm[v.cls[p]] : v.tmpl + v.cls[p];
    const int n = tm->n, n_nt = tm->n_nt;
#define AT(arr, i) arr[(i) * kMicroBlock + tid]
    // batch 1: per-decision-node integers (coalesced: slot-major layout)
#pragma unroll 4
    for (int r = 0; r < n_nt; ++r) {
        AT(scb, r) = v.colbase[r * P + p];
        AT(ssr, r) = v.slot_r[r * P + p];
        AT(sso, r) = v.slot_o[r * P + p];
    }
    const double pcv = v.pc[p];
#pragma unroll 4
    for (int j = 1; j < n; ++j) {
        const int rk = tm->rank[j];
        if (rk & 0x80) {
            AT(sv1, j) = v.pay1[(rk & 0x7F) * P + p];
            if (!ZS) AT(sv2, j) = v.pay2[(rk & 0x7F) * P + p];
        }
    }
    rlp::pdl_wait();                 // below: strategy, portal reach, iteration counter
    const int tl_slot = rlp_tl_begin(v.iter, 1);
    const double w = v.iter->linear ? (double)v.iter->t : 1.0;
    const int pi = v.pidx[p], ri = v.ridx[p];
    AT(sq1, 0) = v.portal_p1[ri];
    AT(sq2, 0) = v.portal_p2[ri];
    // batch 2: strategy probability of every edge, payoff of every leaf (independent loads)
#pragma unroll 4
    for (int j = 1; j < n; ++j) AT(ssig, j) = v.sigma[(long long)tm->par_slot[j] * v.num_isets + AT(scb, tm->par_rank[j])];
    // forward (cfr.py:208,217)
    for (int r = 0; r + 1 < n_nt; ++r) {
        const int j = tm->nt_node[r];
        const int pl = tm->player[j], fc = tm->first_child[j], nc = tm->nchild[j];
        const double a = AT(sq1, r), b = AT(sq2, r);
        for (int k = 0; k < nc; ++k) {
            const int rk = tm->rank[fc + k];
            if (rk & 0x80) continue;
            const double e = AT(ssig, fc + k);
            AT(sq1, rk) = pl == 1 ? e * a : a;
            AT(sq2, rk) = pl == 2 ? e * b : b;
        }
    }
    // backward (cfr.py:204-242)
    for (int r = n_nt - 1; r >= 0; --r) {
        const int j = tm->nt_node[r];
        const int pl = tm->player[j], fc = tm->first_child[j], nc = tm->nchild[j];
        double a1 = 0.0, a2 = 0.0;
        for (int k = 0; k < nc; ++k) {
            const double e = AT(ssig, fc + k);
            a1 = a1 + e * AT(sv1, fc + k);
            if (!ZS) a2 = a2 + e * AT(sv2, fc + k);
        }
        AT(sv1, j) = a1;
        if (!ZS) AT(sv2, j) = a2;
        const double x1 = AT(sq1, r), x2 = AT(sq2, r);
        const double opp = pl == 1 ? pcv * x2 : pcv * x1;
        const double own = pl == 1 ? x1 : x2;
        const double vp = pl == 1 ? a1 : (ZS ? -a1 : a2);
        double *dst = v.contrib_r + AT(sso, r);
        for (int k = 0; k < nc; ++k) {
            const double x = pl == 1 ? AT(sv1, fc + k) : (ZS ? -AT(sv1, fc + k) : AT(sv2, fc + k));
            dst[(long long)k * v.nd] = w * (x - vp) * opp;
        }
        v.contrib_o[AT(sso, r)] = w * own;
    }
    v.portal_v1[pi] = AT(sv1, 0);
    if (!ZS) v.portal_v2[pi] = AT(sv2, 0);
    rlp_tl_end(v.iter, tl_slot);
#undef AT
}

// Chance nodes whose children are all portals (in Leduc: the board deal): value = sum over children, in child
// order, of chance probability x the value the micro kernel left at the portal.  One thread per (node, tile).
template <bool ZS>
__global__ void __launch_bounds__(128)
k_fanin(const FanDesc *__restrict__ fd, int F, long long T, const double *__restrict__ cp, const double *__restrict__ pv1,
        const double *__restrict__ pv2, double *__restrict__ out1, double *__restrict__ out2) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)F * T) return;
    const int f = (int)(gid / T);
    const long long i = gid - (long long)f * T;
    const FanDesc d = fd[f];
    double a1 = 0.0, a2 = 0.0;
    int c = 0;
    for (; c + 8 <= d.nchild; c += 8) {          // 16 (24) independent loads in flight, then the ordered adds
        double e[8], x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            e[j] = cp[(long long)(d.e0 + c + j) * T + i];
            x[j] = pv1[(long long)(d.k0 + c + j) * T + i];
            if (!ZS) y[j] = pv2[(long long)(d.k0 + c + j) * T + i];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a1 = a1 + e[j] * x[j];
            if (!ZS) a2 = a2 + e[j] * y[j];
        }
    }
    for (; c < d.nchild; ++c) {
        const double e = cp[(long long)(d.e0 + c) * T + i];
        a1 = a1 + e * pv1[(long long)(d.k0 + c) * T + i];
        if (!ZS) a2 = a2 + e * pv2[(long long)(d.k0 + c) * T + i];
    }
    out1[gid] = a1;
    if (!ZS) out2[gid] = a2;
}

__global__ void k_fill(double *p, long long n, double x) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = x;
}

}  // namespace

// ------------------------------------------------------------------------------------ host
namespace rlp {

int jit_micro_kernel(const MicroTmpl &tm, bool zs, cudaLibrary_t *lib_out, cudaKernel_t *kernel, std::string *log_out);
int jit_upper_kernels(const UpperShape &sh, bool zs, cudaLibrary_t *lib_out, cudaKernel_t *k0, cudaKernel_t *k1, std::string *log_out);

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
    if (koff_col.back() >= (int64_t)INT32_MAX) return RLP_OK;     // slots are 32-bit
    const bool zs = t->zero_sum;

    std::vector<int32_t> parent(N, -1);
    for (int64_t i = 0; i < N; ++i)
        for (int32_t c = fc[i]; c < fc[i + 1]; ++c) parent[c] = (int32_t)i;

    // ---- micro subtrees: maximal, <= 32 nodes, <= 8 non-terminals, no chance node, <= 8 children --------
    std::vector<int32_t> cn(N, 1), cnt(N, 0);
    std::vector<uint8_t> bad(N, 0);
    bool wide = false;
    for (int64_t i = N - 1; i >= 0; --i) {
        int nc = fc[i + 1] - fc[i];
        if (nc > 255) wide = true;
        if (pl[i] >= 0) cnt[i] += 1;
        if (pl[i] == 0 || nc > kMicroAct) bad[i] = 1;
        if (i > 0) {
            int32_t p = parent[i];
            cn[p] = std::min(1 << 20, cn[p] + cn[i]);
            cnt[p] = std::min(1 << 20, cnt[p] + cnt[i]);
            bad[p] |= bad[i];
        }
    }
    if (wide) return RLP_OK;
    auto micro_ok = [&](int64_t v) { return pl[v] > 0 && !bad[v] && cn[v] <= kMicroNodes && cnt[v] <= kMicroNt; };
    std::vector<uint8_t> kind(N, 0);       // 0 upper, 1 micro root, 2 inside a micro subtree
    for (int64_t v = 0; v < N; ++v) {
        uint8_t pk = v == 0 ? 0 : kind[parent[v]];
        if (pk != 0) kind[v] = 2;
        else if (micro_ok(v)) kind[v] = 1;
    }

    // ---- upper tree statistics (micro roots count as leaves) -------------------------------------------
    std::vector<int32_t> un(N, 0), unt(N, 0), uce(N, 0);
    std::vector<uint8_t> height(N, 0);
    for (int64_t i = N - 1; i >= 0; --i) {
        if (kind[i] == 2) continue;
        un[i] += 1;
        if (kind[i] == 0 && pl[i] >= 0) unt[i] += 1;
        if (kind[i] == 0 && pl[i] == 0) uce[i] += fc[i + 1] - fc[i];
        if (i > 0) {
            int32_t p = parent[i];
            un[p] = std::min(1 << 24, un[p] + un[i]);
            unt[p] += unt[i]; uce[p] += uce[i];
            height[p] = std::max<int>(height[p], std::min(254, height[i] + 1));
        }
    }

    // ---- upper tile roots: maximal upper subtrees that fit; everything above must be chance ---------------
    std::vector<uint8_t> status(N, 0);     // upper nodes: 0 trunk, 1 tile root, 2 inside an upper tile
    std::vector<int32_t> roots;
    bool ok = false;
    for (size_t budget : {kSmemSmall, kSmemMid, kSmemLarge}) {
        auto fits = [&](int64_t v) {
            return kind[v] == 0 && pl[v] >= 0 && un[v] <= 32767 && uce[v] <= 65535 && height[v] + 1 <= kMaxTileLevels &&
                   tile_smem(un[v], unt[v], uce[v], zs) <= budget;
        };
        roots.clear();
        ok = true;
        for (int64_t v = 0; v < N && ok; ++v) {
            if (kind[v] == 2) continue;
            uint8_t ps = v == 0 ? 0 : status[parent[v]];
            if (ps != 0) { status[v] = 2; continue; }
            if (kind[v] == 1) { status[v] = 0; continue; }      // a portal hanging off the trunk
            if (pl[v] == 0) { status[v] = 0; continue; }         // chance nodes below the trunk extend it: the cut goes
                                                                // as deep as it can, giving more and smaller tiles
            if (fits(v)) { status[v] = 1; roots.push_back((int32_t)v); continue; }
            status[v] = 0;
            if (pl[v] > 0) ok = false;            // a decision node above the cut: not handled by the tiled path
        }
        if (ok) break;
    }
    if (!ok) return RLP_OK;

    // static chance reach of every node (pi_c does not depend on the strategy): cp * pi_c, top-down
    std::vector<double> pcs(N);
    pcs[0] = 1.0;
    for (int64_t i = 0; i < N; ++i)
        for (int32_t c = fc[i]; c < fc[i + 1]; ++c) pcs[c] = pl[i] == 0 ? d->cprob[c] * pcs[i] : pcs[i];

    // ---- build the upper tiles; portals are numbered tile-major ----------------------------------------------
    const size_t T = roots.size();
    std::vector<TileDesc> desc(T);
    std::vector<unsigned short> ntof, term_lid, portal_lid;
    std::vector<TileRec> rec;
    std::vector<TileAux> aux;
    std::vector<double> pc, pay1, pay2, cprob;
    std::vector<int32_t> portal_node;      // portal index -> global node id of the micro root
    std::vector<int32_t> prank(N, -1);     // micro root -> rank of its portal inside its tile
    std::vector<int32_t> local, ldepth;
    int max_nodes = 0, max_nt = 0, max_cp = 0;
    for (size_t k = 0; k < T; ++k) {
        const int32_t r = roots[k];
        TileDesc &td = desc[k];
        memset(&td, 0, sizeof(td));
        td.node_off = (long long)ntof.size(); td.nt_off = (long long)rec.size();
        td.term_off = (long long)term_lid.size(); td.cp_off = (long long)cprob.size();
        td.portal_off = (long long)portal_node.size();
        td.root_global = r; td.root_p1 = 1.0; td.root_p2 = 1.0;
        local.clear(); ldepth.clear();
        local.push_back(r); ldepth.push_back(0);
        for (size_t q = 0; q < local.size(); ++q) {
            int32_t g = local[q];
            if (kind[g] == 1) continue;                        // portal leaf: do not descend
            for (int32_t c = fc[g]; c < fc[g + 1]; ++c) { local.push_back(c); ldepth.push_back(ldepth[q] + 1); }
        }
        int nt = 0, cpo = 0, next_child = 1, level = -1, nportal = 0, nterm = 0;
        for (size_t q = 0; q < local.size(); ++q) {
            int32_t g = local[q];
            if (kind[g] == 1) {
                ntof.push_back((unsigned short)(0x8000 | nportal));
                portal_lid.push_back((unsigned short)q);
                prank[g] = nportal;
                portal_node.push_back(g);
                ++nportal;
                continue;
            }
            if (pl[g] < 0) {
                ntof.push_back(0xFFFF);
                term_lid.push_back((unsigned short)q);
                pay1.push_back(d->u1[g]);
                if (!zs) pay2.push_back(d->u2[g]);
                ++nterm;
                continue;
            }
            while (level < ldepth[q]) td.lvl_nt[++level] = nt;
            ntof.push_back((unsigned short)nt);
            int nc = fc[g + 1] - fc[g];
            TileRec rc;
            rc.first_child = (unsigned short)next_child;
            rc.lid = (unsigned short)q;
            rc.nchild = (unsigned char)nc;
            rc.player = (unsigned char)pl[g];
            rc.cp_off = 0;
            TileAux ax{0, 0, 0, 0};
            if (pl[g] == 0) {
                rc.cp_off = (unsigned short)cpo;
                for (int32_t c = fc[g]; c < fc[g + 1]; ++c) { cprob.push_back(d->cprob[c]); ++cpo; }
            } else {
                ax.colbase = colbase[g];
                ax.slot_r = (int)(koff_col[krank[g]] + colbase[g]);
                ax.slot_o = (int)(koff_iset[krank[g]] + srank[g]);
            }
            rec.push_back(rc);
            aux.push_back(ax);
            pc.push_back(pcs[g]);
            next_child += nc;
            ++nt;
        }
        if (nportal > 0x7FFF || nt > 0x7FFF) return RLP_OK;
        td.n_nodes = (int)local.size(); td.n_nt = nt; td.n_term = nterm; td.n_cp = cpo; td.n_portal = nportal;
        td.n_levels = level + 1;
        for (int l = level + 1; l <= kMaxTileLevels; ++l) td.lvl_nt[l] = nt;
        max_nodes = std::max(max_nodes, td.n_nodes);
        max_nt = std::max(max_nt, td.n_nt);
        max_cp = std::max(max_cp, cpo);
    }
    // portals that hang directly off the (all-chance) trunk: reach (1, 1) for ever, value unused
    for (int64_t v = 0; v < N; ++v)
        if (kind[v] == 1 && status[v] == 0) portal_node.push_back((int32_t)v);

    // ---- micro instances: one per portal, shapes deduplicated into templates ----------------------------------
    const int64_t P = (int64_t)portal_node.size();
    // portal reach / value arrays are laid out [portal rank inside its tile][tile] so that the tile-side kernels
    // (one thread or CTA per tile) touch consecutive words; trunk portals follow.
    int kp = 0;
    for (size_t k = 0; k < T; ++k) kp = std::max(kp, desc[k].n_portal);
    const int64_t portal_len = (int64_t)kp * (int64_t)T + (P - (T ? desc[T - 1].portal_off + desc[T - 1].n_portal : 0));
    if (portal_len >= (int64_t)INT32_MAX) return RLP_OK;
    std::vector<int> pidx(P);
    for (size_t k = 0; k < T; ++k)
        for (int q = 0; q < desc[k].n_portal; ++q) pidx[desc[k].portal_off + q] = (int)((int64_t)q * (int64_t)T + (int64_t)k);
    {
        int64_t first_trunk = T ? desc[T - 1].portal_off + desc[T - 1].n_portal : 0;
        for (int64_t p = first_trunk; p < P; ++p) pidx[p] = (int)((int64_t)kp * (int64_t)T + (p - first_trunk));
    }
    std::map<std::string, int> class_of;
    std::vector<MicroTmpl> tmpl;
    std::vector<unsigned short> m_class(P);
    int mnt = 0, mterm = 0, mnodes = 0;
    std::vector<int32_t> loc;
    for (int64_t p = 0; p < P; ++p) {
        loc.clear();
        loc.push_back(portal_node[p]);
        for (size_t q = 0; q < loc.size(); ++q)
            for (int32_t c = fc[loc[q]]; c < fc[loc[q] + 1]; ++c) loc.push_back(c);
        std::string sig;
        sig.reserve(loc.size() * 2);
        for (int32_t g : loc) { sig.push_back((char)(pl[g] + 2)); sig.push_back((char)(fc[g + 1] - fc[g])); }
        auto it = class_of.find(sig);
        int cls;
        if (it == class_of.end()) {
            cls = (int)tmpl.size();
            if (cls >= 65535) return RLP_OK;
            class_of.emplace(sig, cls);
            MicroTmpl tm;
            memset(&tm, 0, sizeof(tm));
            int nt = 0, term = 0, next_child = 1;
            for (size_t q = 0; q < loc.size(); ++q) {
                int32_t g = loc[q];
                int nc = fc[g + 1] - fc[g];
                tm.player[q] = (unsigned char)(pl[g] < 0 ? 0xFF : pl[g]);
                tm.nchild[q] = (unsigned char)nc;
                tm.first_child[q] = (unsigned char)next_child;
                if (pl[g] < 0) tm.rank[q] = (unsigned char)(0x80 | term++);
                else {
                    for (int k = 0; k < nc; ++k) { tm.par_rank[next_child + k] = (unsigned char)nt; tm.par_slot[next_child + k] = (unsigned char)k; }
                    tm.nt_node[nt] = (unsigned char)q; tm.rank[q] = (unsigned char)nt++;
                }
                next_child += nc;
            }
            tm.n = (unsigned char)loc.size(); tm.n_nt = (unsigned char)nt; tm.n_term = (unsigned char)term;
            tmpl.push_back(tm);
            mnt = std::max(mnt, nt);
            mterm = std::max(mterm, term);
            mnodes = std::max(mnodes, (int)loc.size());
        } else {
            cls = it->second;
        }
        m_class[p] = (unsigned short)cls;
    }
    std::vector<int> m_colbase((size_t)mnt * P, 0), m_slot_r((size_t)mnt * P, 0), m_slot_o((size_t)mnt * P, 0);
    std::vector<double> m_pay1((size_t)mterm * P, 0.0), m_pay2(zs ? 0 : (size_t)mterm * P, 0.0), m_pc(P);
    for (int64_t p = 0; p < P; ++p) {
        loc.clear();
        loc.push_back(portal_node[p]);
        for (size_t q = 0; q < loc.size(); ++q)
            for (int32_t c = fc[loc[q]]; c < fc[loc[q] + 1]; ++c) loc.push_back(c);
        int nt = 0, term = 0;
        for (int32_t g : loc) {
            if (pl[g] < 0) {
                m_pay1[(size_t)term * P + p] = d->u1[g];
                if (!zs) m_pay2[(size_t)term * P + p] = d->u2[g];
                ++term;
            } else {
                m_colbase[(size_t)nt * P + p] = srank[g];          // internal information-set id (sigma is read action-major)
                m_slot_r[(size_t)nt * P + p] = (int)(koff_col[krank[g]] + colbase[g]);
                m_slot_o[(size_t)nt * P + p] = (int)(koff_iset[krank[g]] + srank[g]);
                ++nt;
            }
        }
        m_pc[p] = pcs[portal_node[p]];
    }

    // ---- launch geometry of the upper tiles -------------------------------------------------------------------
    ts.num_tiles = (int)T;
    ts.max_nodes = max_nodes; ts.max_nt = max_nt; ts.max_cp = max_cp;
    ts.smem_bytes = T ? tile_smem(max_nodes, max_nt, max_cp, zs) : 0;
    if (ts.smem_bytes > 226 * 1024) return RLP_OK;               // (max_nodes, max_nt) came from different tiles
    ts.block = 1024;
    for (int b : {64, 128, 256, 512})
        if ((max_nt + b - 1) / b <= 2) { ts.block = b; break; }
    ts.slots = std::max(1, (max_nt + ts.block - 1) / ts.block);
    if (ts.slots > 4) return RLP_OK;
    ts.wide = t->max_act > kRegAct;
    ts.num_micro = (int)P; ts.num_classes = (int)tmpl.size(); ts.micro_nt = mnt; ts.micro_term = mterm; ts.micro_nodes = mnodes;
    ts.micro_smem = (size_t)kMicroBlock * ((size_t)mnodes * 8 * (zs ? 2 : 3) + (size_t)mnt * (16 + 12));

    RLP_CUDA(ts.desc.upload(desc, stream));
    RLP_CUDA(ts.ntof.upload(ntof, stream));
    RLP_CUDA(ts.term_lid.upload(term_lid, stream));
    RLP_CUDA(ts.portal_lid.upload(portal_lid, stream));
    RLP_CUDA(ts.rec.upload(rec, stream));
    RLP_CUDA(ts.aux.upload(aux, stream));
    RLP_CUDA(ts.pc.upload(pc, stream));
    RLP_CUDA(ts.pay1.upload(pay1, stream));
    RLP_CUDA(ts.pay2.upload(pay2, stream));
    RLP_CUDA(ts.cprob.upload(cprob, stream));
    RLP_CUDA(ts.root_val.alloc(T));
    RLP_CUDA(ts.tmpl.upload(tmpl, stream));
    RLP_CUDA(ts.m_class.upload(m_class, stream));
    RLP_CUDA(ts.m_colbase.upload(m_colbase, stream));
    RLP_CUDA(ts.m_slot_r.upload(m_slot_r, stream));
    RLP_CUDA(ts.m_slot_o.upload(m_slot_o, stream));
    RLP_CUDA(ts.m_pay1.upload(m_pay1, stream));
    RLP_CUDA(ts.m_pay2.upload(m_pay2, stream));
    RLP_CUDA(ts.m_pc.upload(m_pc, stream));
    std::vector<int> ridx(pidx);
    RLP_CUDA(ts.m_pidx.upload(pidx, stream));
    RLP_CUDA(ts.m_ridx.upload(ridx, stream));
    RLP_CUDA(ts.portal_p1.alloc((size_t)portal_len)); RLP_CUDA(ts.portal_p2.alloc((size_t)portal_len));
    RLP_CUDA(ts.portal_v1.alloc((size_t)portal_len)); RLP_CUDA(ts.portal_v2.alloc(zs ? 0 : (size_t)portal_len));
    if (portal_len > 0 && !rlp::g_dry_run) {
        unsigned g = (unsigned)((portal_len + 255) / 256);
        k_fill<<<g, 256, 0, stream>>>(ts.portal_p1.p, portal_len, 1.0); RLP_LAUNCH_CHECK();
        k_fill<<<g, 256, 0, stream>>>(ts.portal_p2.p, portal_len, 1.0); RLP_LAUNCH_CHECK();
    }
    RLP_CUDA(rlp::sync_stream(stream));
    // ---- shape-specialised kernels (jit.cu): all classes or none ------------------------------------------------
    ts.jit.clear();
    if (P > 0 && tmpl.size() <= 8) {
        std::vector<std::vector<int>> members(tmpl.size());
        if (tmpl.size() > 1)
            for (int64_t p = 0; p < P; ++p) members[m_class[p]].push_back((int)p);
        bool all = true;
        for (size_t c = 0; c < tmpl.size() && all; ++c) {
            std::unique_ptr<JitClass> jc(new JitClass);
            std::string log;
            int rc = jit_micro_kernel(tmpl[c], zs, &jc->lib, &jc->kernel, &log);
            if (rc) return rc;
            if (!jc->kernel) { all = false; if (!rlp::g_dry_run) break; else continue; }
            if (tmpl.size() > 1) {
                jc->count = (int)members[c].size();
                RLP_CUDA(jc->inst.upload(members[c], stream));
            } else {
                jc->count = (int)P;
            }
            ts.jit.push_back(std::move(jc));
        }
        if (!all) ts.jit.clear();
        RLP_CUDA(rlp::sync_stream(stream));
    }
    // ---- information sets by where their nodes live: A = only inside micro subtrees, B = the rest ----------------
    {
        std::vector<uint8_t> in_a((size_t)t->I, 1);
        for (int64_t g = 0; g < N; ++g)
            if (pl[g] > 0 && kind[g] == 0) in_a[srank[g]] = 0;
        std::vector<int> la, lb;
        for (int32_t r = 0; r < t->I; ++r) (in_a[r] ? la : lb).push_back(r);
        ts.acc_a_count = (int)la.size();
        ts.acc_b_count = (int)lb.size();
        RLP_CUDA(ts.acc_list_a.upload(la, stream));
        RLP_CUDA(ts.acc_list_b.upload(lb, stream));
        RLP_CUDA(rlp::sync_stream(stream));
    }
    // ---- upper tiles of ONE shape: straight-line kernels, one thread per tile ---------------------------------------
    ts.upper_jit.reset();
    if (T > 0 && max_nodes <= 640 && (!ts.jit.empty() || rlp::g_dry_run)) {
        UpperShape sh;
        std::vector<std::vector<int32_t>> locals(T);
        bool same = true;
        auto is_fan = [&](int32_t g) {          // chance node of the upper tree whose children are all portals
            if (kind[g] != 0 || pl[g] != 0 || fc[g + 1] - fc[g] < 2) return false;
            for (int32_t c = fc[g]; c < fc[g + 1]; ++c) if (kind[c] != 1) return false;
            return true;
        };
        for (size_t k = 0; k < T && same; ++k) {
            std::vector<int32_t> &lc = locals[k];
            lc.push_back(roots[k]);
            for (size_t q = 0; q < lc.size(); ++q) {
                int32_t g = lc[q];
                if (kind[g] == 1 || is_fan(g)) continue;
                for (int32_t c = fc[g]; c < fc[g + 1]; ++c) lc.push_back(c);
            }
            UpperShape cur;
            int next_child = 1, nfan = 0;
            for (int32_t g : lc) {
                bool fan = is_fan(g);
                signed char kd = kind[g] == 1 ? 3 : (fan ? 4 : (signed char)pl[g]);
                int nc = (kind[g] == 1 || fan) ? 0 : fc[g + 1] - fc[g];
                cur.kind.push_back(kd); cur.nchild.push_back(nc); cur.first_child.push_back(next_child);
                cur.portal_rank.push_back(kind[g] == 1 ? prank[g] : -1);
                cur.fan_index.push_back(fan ? nfan++ : -1);
                cur.fan_k0.push_back(fan ? prank[fc[g]] : -1);
                next_child += nc;
            }
            if (k == 0) sh = cur;
            else same = cur.kind == sh.kind && cur.nchild == sh.nchild && cur.portal_rank == sh.portal_rank &&
                        cur.fan_k0 == sh.fan_k0;
            if (same && k > 0)                       // fan nodes must also agree on their fan-out
                for (size_t j = 0; j < lc.size() && same; ++j)
                    if (sh.kind[j] == 4) same = fc[lc[j] + 1] - fc[lc[j]] == fc[locals[0][j] + 1] - fc[locals[0][j]];
        }
        if (same) {
            std::unique_ptr<UpperJit> uj(new UpperJit);
            std::string log;
            int rc = jit_upper_kernels(sh, zs, &uj->lib, &uj->k0, &uj->k1, &log);
            if (rc) return rc;
            if ((uj->k0 && uj->k1) || rlp::g_dry_run) {
                const int n = (int)sh.kind.size();
                int nd = 0, ne = 0, ntm = 0;
                for (int j = 0; j < n; ++j) {
                    if (sh.kind[j] == 1 || sh.kind[j] == 2) ++nd;
                    if (sh.kind[j] == -1) ++ntm;
                    if (sh.kind[j] == 0) ne += sh.nchild[j];
                }
                std::vector<int> ucb((size_t)nd * T), usr((size_t)nd * T), uso((size_t)nd * T);
                std::vector<double> upc((size_t)nd * T), ucp((size_t)ne * T), up1((size_t)ntm * T), up2(zs ? 0 : (size_t)ntm * T);
                std::vector<long long> upo(T);
                for (size_t k = 0; k < T; ++k) {
                    int di = 0, e = 0, tt = 0;
                    for (int j = 0; j < n; ++j) {
                        int32_t g = locals[k][j];
                        if (sh.kind[j] == 1 || sh.kind[j] == 2) {
                            ucb[(size_t)di * T + k] = colbase[g];
                            usr[(size_t)di * T + k] = (int)(koff_col[krank[g]] + colbase[g]);
                            uso[(size_t)di * T + k] = (int)(koff_iset[krank[g]] + srank[g]);
                            upc[(size_t)di * T + k] = pcs[g];
                            ++di;
                        } else if (sh.kind[j] == -1) {
                            up1[(size_t)tt * T + k] = d->u1[g];
                            if (!zs) up2[(size_t)tt * T + k] = d->u2[g];
                            ++tt;
                        } else if (sh.kind[j] == 0) {
                            for (int32_t c = fc[g]; c < fc[g + 1]; ++c) ucp[(size_t)(e++) * T + k] = d->cprob[c];
                        }
                    }
                    upo[k] = desc[k].portal_off;
                }
                uj->T = (long long)T;
                // fan-in chance nodes: descriptors, their chance probabilities, and where their portals read reach
                std::vector<FanDesc> fans;
                int fan_edges = 0;
                for (int j = 0; j < n; ++j)
                    if (sh.kind[j] == 4) {
                        int32_t g0 = locals[0][j];
                        FanDesc fdsc{sh.fan_k0[j], fc[g0 + 1] - fc[g0], fan_edges, 0};
                        fan_edges += fdsc.nchild;
                        fans.push_back(fdsc);
                    }
                std::vector<double> fcp((size_t)fan_edges * T);
                if (!fans.empty()) {
                    for (size_t k = 0; k < T; ++k) {
                        int f = 0;
                        for (int j = 0; j < n; ++j)
                            if (sh.kind[j] == 4) {
                                int32_t g = locals[k][j];
                                const FanDesc &fdsc = fans[f++];
                                for (int c = 0; c < fdsc.nchild; ++c) {
                                    fcp[(size_t)(fdsc.e0 + c) * T + k] = d->cprob[fc[g] + c];
                                    ridx[desc[k].portal_off + fdsc.k0 + c] = (int)((int64_t)fdsc.k0 * (int64_t)T + (int64_t)k);
                                }
                            }
                    }
                    RLP_CUDA(ts.m_ridx.upload(ridx, stream));
                }
                uj->num_fan = (int)fans.size();
                RLP_CUDA(uj->fan.upload(fans, stream));
                RLP_CUDA(uj->fan_cp.upload(fcp, stream));
                RLP_CUDA(uj->fan_v1.alloc(fans.size() * T));
                RLP_CUDA(uj->fan_v2.alloc(zs ? 0 : fans.size() * T));
                RLP_CUDA(uj->colbase.upload(ucb, stream)); RLP_CUDA(uj->slot_r.upload(usr, stream));
                RLP_CUDA(uj->slot_o.upload(uso, stream)); RLP_CUDA(uj->pcs.upload(upc, stream));
                RLP_CUDA(uj->cprob.upload(ucp, stream)); RLP_CUDA(uj->pay1.upload(up1, stream));
                RLP_CUDA(uj->pay2.upload(up2, stream)); RLP_CUDA(uj->poff.upload(upo, stream));
                RLP_CUDA(rlp::sync_stream(stream));
                ts.upper_jit = std::move(uj);
                // group-fused persistent kernel (fused.cu) when the tree has the group structure
                if (tmpl.size() == 1 && (!ts.jit.empty() || rlp::g_dry_run)) {
                    rlp::FusedCtx cx{d, t, &srank, &krank, &portal_node, &desc, &locals, &sh, &fans, &tmpl[0], &m_pay1,
                                     &pcs, &ridx, &fcp, kp};
                    int frc = rlp::plan_fused(cx, stream);
                    if (frc) return frc;
                }
            }
        }
    }
    ts.enabled = true;
    return RLP_OK;
}

static int launch_upper_jit(rlp_cfr *c, TileSet &ts, int phase, const IterState *it, cudaStream_t s) {
    UpperJit &u = *ts.upper_jit;
    long long T = u.T, nd = c->tree->ND;
    void *args[] = {&u.colbase.p, &u.slot_r.p, &u.slot_o.p, &u.pcs.p, &u.cprob.p, &u.pay1.p, &u.pay2.p, &u.poff.p,
                    &c->sigma.p, &ts.portal_p1.p, &ts.portal_p2.p, &ts.portal_v1.p, &ts.portal_v2.p,
                    &c->contrib_r.p, &c->contrib_o.p, &ts.root_val.p, &it, &T, &u.fan_v1.p, &u.fan_v2.p, &nd};
    if (phase == 1 && u.num_fan > 0) {
        long long total = (long long)u.num_fan * T;
        unsigned g = (unsigned)((total + 127) / 128);
        if (c->tree->zero_sum)
            k_fanin<true><<<g, 128, 0, s>>>(u.fan.p, u.num_fan, T, u.fan_cp.p, ts.portal_v1.p, ts.portal_v2.p, u.fan_v1.p, u.fan_v2.p);
        else
            k_fanin<false><<<g, 128, 0, s>>>(u.fan.p, u.num_fan, T, u.fan_cp.p, ts.portal_v1.p, ts.portal_v2.p, u.fan_v1.p, u.fan_v2.p);
        RLP_LAUNCH_CHECK();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((T + 63) / 64)); cfg.blockDim = dim3(64); cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = rlp::pdl_enabled() ? 1 : 0;
    RLP_CUDA(cudaLaunchKernelExC(&cfg, (const void *)(phase == 0 ? u.k0 : u.k1), args));
    rlp::g_launches.fetch_add(1, std::memory_order_relaxed);
    return RLP_OK;
}

template <bool ZS, int R, bool WIDE, int PHASE>
static int launch_variant(const TileSet &ts, const TileView &v, cudaStream_t s) {
    // the attribute is per device (context), not per thread: one bit per device ordinal
    static std::atomic<uint64_t> attr_set{0};
    int dev = 0;
    RLP_CUDA(cudaGetDevice(&dev));
    if (!(attr_set.load() >> (dev & 63) & 1)) {
        RLP_CUDA(cudaFuncSetAttribute(k_tile_sweep<ZS, R, WIDE, PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_set.fetch_or(1ull << (dev & 63));
    }
    RLP_CUDA(rlp::launch_pdl(k_tile_sweep<ZS, R, WIDE, PHASE>, dim3(ts.num_tiles), dim3(ts.block), ts.smem_bytes, s, v));
    rlp::g_launches.fetch_add(1, std::memory_order_relaxed);
    return RLP_OK;
}

template <bool ZS, bool WIDE, int PHASE>
static int launch_slots(const TileSet &ts, const TileView &v, cudaStream_t s) {
    switch (ts.slots) {
        case 1: return launch_variant<ZS, 1, WIDE, PHASE>(ts, v, s);
        case 2: return launch_variant<ZS, 2, WIDE, PHASE>(ts, v, s);
        case 3: return launch_variant<ZS, 3, WIDE, PHASE>(ts, v, s);
        default: return launch_variant<ZS, 4, WIDE, PHASE>(ts, v, s);
    }
}

template <int PHASE>
static int launch_upper(rlp_tree *t, const TileSet &ts, const TileView &v, cudaStream_t s) {
    if (t->zero_sum) return ts.wide ? launch_slots<true, true, PHASE>(ts, v, s) : launch_slots<true, false, PHASE>(ts, v, s);
    return ts.wide ? launch_slots<false, true, PHASE>(ts, v, s) : launch_slots<false, false, PHASE>(ts, v, s);
}

// One tiled sweep: [upper tiles, reach to portals] -> [micro subtrees] -> [upper tiles, values from portals].
int launch_tile_stage(rlp_cfr *c, cudaStream_t s, const IterState *iter, int stage, int *launches);

int launch_tile_sweep(rlp_cfr *c, cudaStream_t s, const IterState *iter, int *launches, cudaEvent_t *marks) {
    int n = 0;
    for (int stage = 0; stage < 3; ++stage) {
        if (marks) RLP_CUDA(cudaEventRecord(marks[stage], s));
        int rc = launch_tile_stage(c, s, iter, stage, &n);
        if (rc) return rc;
    }
    if (marks) RLP_CUDA(cudaEventRecord(marks[3], s));
    if (launches) *launches = n;
    return RLP_OK;
}

// stage 0: reach down to the portals; 1: micro subtrees; 2: values up from the portals + upper contributions
int launch_tile_stage(rlp_cfr *c, cudaStream_t s, const IterState *iter, int stage, int *launches) {
    rlp_tree *t = c->tree;
    TileSet &ts = t->tiles;
    TileView v;
    v.desc = ts.desc.p; v.ntof = ts.ntof.p; v.term_lid = ts.term_lid.p; v.portal_lid = ts.portal_lid.p;
    v.rec = ts.rec.p; v.aux = ts.aux.p;
    v.pc = ts.pc.p; v.pay1 = ts.pay1.p; v.pay2 = ts.pay2.p; v.cprob = ts.cprob.p; v.sigma = c->sigma.p;
    v.contrib_r = c->contrib_r.p; v.contrib_o = c->contrib_o.p; v.root_val = ts.root_val.p;
    v.portal_p1 = ts.portal_p1.p; v.portal_p2 = ts.portal_p2.p; v.portal_v1 = ts.portal_v1.p; v.portal_v2 = ts.portal_v2.p;
    v.iter = iter ? iter : c->iter.p; v.max_nodes = ts.max_nodes; v.max_nt = ts.max_nt; v.max_cp = ts.max_cp;
    v.nd = t->ND;
    int n = 0;
    if (stage == 0 && ts.num_tiles > 0 && ts.num_micro > 0) {
        int rc = ts.upper_jit ? launch_upper_jit(c, ts, 0, v.iter, s) : launch_upper<0>(t, ts, v, s);
        if (rc) return rc;
        ++n;
    }
    if (stage == 1 && ts.num_micro > 0 && !ts.jit.empty()) {
        const IterState *it = v.iter;
        long long P = ts.num_micro;
        for (auto &jc : ts.jit) {
            if (jc->count == 0) continue;
            const int *inst = jc->inst.n ? jc->inst.p : nullptr;
            int count = jc->count;
            long long nd = t->ND, ni = t->I;
            void *args[] = {&ts.m_colbase.p, &ts.m_slot_r.p, &ts.m_slot_o.p, &ts.m_pidx.p, &ts.m_ridx.p, &ts.m_pay1.p, &ts.m_pay2.p, &ts.m_pc.p,
                            &c->sigma_am.p, &ts.portal_p1.p, &ts.portal_p2.p, &ts.portal_v1.p, &ts.portal_v2.p,
                            &c->contrib_r.p, &c->contrib_o.p, &it, &P, &inst, &count, &nd, &ni};
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)((count + 127) / 128)); cfg.blockDim = dim3(128); cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = rlp::pdl_enabled() ? 1 : 0;
            RLP_CUDA(cudaLaunchKernelExC(&cfg, (const void *)jc->kernel, args));
            rlp::g_launches.fetch_add(1, std::memory_order_relaxed);
            ++n;
        }
    } else if (stage == 1 && ts.num_micro > 0) {
        MicroView m;
        m.tmpl = ts.tmpl.p; m.cls = ts.m_class.p; m.colbase = ts.m_colbase.p; m.slot_r = ts.m_slot_r.p;
        m.slot_o = ts.m_slot_o.p; m.pidx = ts.m_pidx.p; m.ridx = ts.m_ridx.p; m.pay1 = ts.m_pay1.p; m.pay2 = ts.m_pay2.p; m.pc = ts.m_pc.p; m.sigma = c->sigma_am.p; m.nd = t->ND; m.num_isets = t->I;
        m.portal_p1 = ts.portal_p1.p; m.portal_p2 = ts.portal_p2.p; m.portal_v1 = ts.portal_v1.p;
        m.portal_v2 = ts.portal_v2.p; m.contrib_r = c->contrib_r.p; m.contrib_o = c->contrib_o.p;
        m.iter = v.iter; m.P = ts.num_micro; m.num_classes = ts.num_classes; m.mn = ts.micro_nodes; m.mnt = ts.micro_nt;
        unsigned g = (unsigned)((ts.num_micro + kMicroBlock - 1) / kMicroBlock);
        static std::atomic<uint64_t> attr_set{0};      // per device ordinal (the attribute is per context)
        int dev = 0;
        RLP_CUDA(cudaGetDevice(&dev));
        if (!(attr_set.load() >> (dev & 63) & 1)) {
            RLP_CUDA(cudaFuncSetAttribute(k_micro<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            RLP_CUDA(cudaFuncSetAttribute(k_micro<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set.fetch_or(1ull << (dev & 63));
        }
        if (t->zero_sum) RLP_CUDA(rlp::launch_pdl(k_micro<true>, dim3(g), dim3(kMicroBlock), ts.micro_smem, s, m));
        else RLP_CUDA(rlp::launch_pdl(k_micro<false>, dim3(g), dim3(kMicroBlock), ts.micro_smem, s, m));
        rlp::g_launches.fetch_add(1, std::memory_order_relaxed);
        ++n;
    }
    if (stage == 2 && ts.num_tiles > 0) {
        int rc = ts.upper_jit ? launch_upper_jit(c, ts, 1, v.iter, s) : launch_upper<1>(t, ts, v, s);
        if (rc) return rc;
        ++n;
    }
    if (launches) *launches += n;
    return RLP_OK;
}

}  // namespace rlp
