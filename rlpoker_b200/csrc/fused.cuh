// Interface between the tiler (tiles.cu), the solver (cfr.cu) and the group-fused persistent kernel (fused.cu).
#pragma once
#include <string>
#include <vector>

#include "tree.cuh"

struct rlp_cfr;

namespace rlp {

// Host-side views of what build_tiles derived (all valid only during the call).
struct FusedCtx {
    const rlp_tree_desc *d;
    rlp_tree *t;
    const std::vector<int32_t> *srank, *krank;             // per node: internal information set, rank inside it
    const std::vector<int32_t> *portal_node;               // [P] global node id of every micro root
    const std::vector<TileDesc> *desc;                     // [T]
    const std::vector<std::vector<int32_t>> *locals;       // [T][shape node] global node ids in upper-shape order
    const UpperShape *sh;
    const std::vector<FanDesc> *fans;
    const MicroTmpl *tm;
    const std::vector<double> *m_pay1;                     // [leaf][P]
    const std::vector<double> *pcs;                        // [N] static chance reach
    const std::vector<int> *ridx;                          // [P] index of the root reach in the portal reach arrays
    const std::vector<double> *fan_cp;                     // [fan edge][T] chance probabilities of the fan-in nodes
    int kp;                                                // most portals of any tile
};

int plan_fused(const FusedCtx &cx, cudaStream_t stream);
bool fused_available(const rlp_cfr *c);
int fused_stat(const rlp_tree *t, int which);
const std::string *fused_source(const rlp_tree *t);
bool fused_br_available(const rlp_cfr *c);
int fused_br(rlp_cfr *c, const double *sigma_int, double *out2_dev, cudaStream_t s);
int fused_run(rlp_cfr *c, int64_t t0, int64_t n_iters, int linear_weight, cudaStream_t s, unsigned long long *tl,
              long long tl_cap);

}  // namespace rlp
