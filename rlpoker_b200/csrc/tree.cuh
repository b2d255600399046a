// Device-resident flattened game and the layout derived for the level sweep.
//
// HBM layout (all structure-of-arrays, node order = breadth-first so each depth is one range):
//   first_child[N+1] i32   children of i are first_child[i] .. first_child[i+1]-1
//   player[N]        i8    -1 terminal, 0 chance, 1, 2
//   colbase[N]       i32   decision nodes: first internal pair index of the node's information set
//   srank[N]         i32   decision nodes: internal (size-sorted) index of the information set
//   krank[N]         i32   decision nodes: rank of the node inside its information set (ascending id)
//   cprob[N]         f64   probability of the chance edge into the node
//   pay1[N] (pay2)   f64   terminal utilities (pay2 only for non-zero-sum games)
// Information sets are renumbered by DESCENDING size ("internal" order) so that the k-th members
// of all information sets that have one form a prefix: contributions are scattered into a
// jagged-diagonal scratch  contrib[koff[k] + column]  which the accumulate kernel streams through
// with unit stride while adding in ascending node order (= the reference's visit order).
#pragma once
#include "common.cuh"
#include "tiles.cuh"

struct rlp_cfr;

struct rlp_tree {
    int64_t N = 0;
    int32_t I = 0;
    int32_t L = 0;
    int64_t Q = 0;
    int64_t ND = 0;        // decision nodes
    int64_t NC = 0;        // sum over decision nodes of #actions (regret-contribution slots)
    int32_t Kmax = 0;      // largest information set
    int32_t max_act = 0;
    int32_t max_chance = 0;   // widest chance node
    bool zero_sum = false;
    int32_t top_levels = 0;   // leading levels small enough to be swept by one block
    std::vector<int64_t> level_off;
    std::vector<int64_t> act_off;        // canonical [I+1]
    std::vector<int32_t> iset_of_srank;  // internal -> canonical information set
    std::vector<int32_t> srank_of_iset;  // canonical -> internal
    std::vector<int64_t> canon_of_col;   // internal pair -> canonical pair
    std::vector<int64_t> col_off_h;      // internal [I+1]
    std::vector<int8_t> iset_player_h;   // internal [I]: owner of every information set
    std::vector<int32_t> first_child_h;  // host copies of the structure (frontier bounds of the sampled lanes)
    std::vector<int8_t> player_h;
    std::vector<int64_t> lane_width[2];  // [variant][level]: most frontier records one lane (both traversals) can hold; lazily filled

    rlp::DevBuf<int32_t> first_child, colbase, srank, krank;
    rlp::DevBuf<int8_t> player;
    rlp::DevBuf<double> cprob, pay1, pay2;
    rlp::DevBuf<int4> node_rec;                   // [N] {first_child, #children | (player + 1) << 8 | leaf mask of the first 16 children << 16, colbase, srank}
    rlp::DevBuf<int64_t> koff_col, koff_iset;     // [Kmax+1]
    rlp::DevBuf<int32_t> iset_size, iset_level;   // [I] internal order
    rlp::DevBuf<int8_t> iset_player;              // [I]
    rlp::DevBuf<int64_t> col_off;                 // [I+1]
    rlp::DevBuf<int64_t> iset_node_off;           // [I+1] internal order
    rlp::DevBuf<int32_t> iset_nodes;              // [ND]
    rlp::DevBuf<int64_t> d_canon_of_col;          // [Q]
    rlp_cfr *scratch = nullptr;                   // solver used by the stand-alone entry points (rlp_best_response ...)
    TileSet tiles;                                // subtree tiles for the shared-memory sweep (may be disabled)

    int64_t bytes() const;
};
