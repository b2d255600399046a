// Subtree tiles: the part of the tree below a cut is split into subtrees ("tiles") small enough for one
// CTA to sweep entirely in shared memory -- forward reach and backward values never touch HBM; only the
// per-decision-node contributions (regret terms, own reach) leave the SM.
//
// A tile is stored in tile-local breadth-first order (children of a node stay contiguous).  Per tile, in HBM:
//   ntof[n_nodes]   u16  non-terminal rank of every local node (0xFFFF = terminal)
//   rec[n_nt]       8 B  per non-terminal: first child (local id), own local id, #children, player,
//                        offset of its chance probabilities
//   aux[n_nt]       16 B per non-terminal: internal pair base of its information set, its slots in the two
//                        jagged-diagonal contribution arrays
//   pc[n_nt]        f64  chance reach of the node (static: it does not depend on the strategy)
//   term_lid/pay    per terminal: local id and player-1 payoff (player-2 payoff only for non-zero-sum games)
//   cprob           chance probabilities of the tile's chance edges
#pragma once
#include "common.cuh"

constexpr int kMaxTileLevels = 32;

struct TileDesc {
    long long node_off;     // into ntof
    long long nt_off;       // into rec / aux / pc
    long long term_off;     // into term_lid / pay
    long long cp_off;       // into cprob
    int n_nodes, n_nt, n_term, n_levels;
    int root_global;        // breadth-first id of the tile root
    int n_cp;               // chance edges inside the tile
    double root_p1, root_p2;            // strategy reach of the root (1.0 below an all-chance trunk)
    int lvl_nt[kMaxTileLevels + 1];     // non-terminal rank ranges per local level
};

struct TileRec {            // 8 bytes
    unsigned short first_child, lid;
    unsigned char nchild, player;
    unsigned short cp_off;
};

struct TileAux {            // 16 bytes
    int colbase, slot_r, slot_o, pad;
};

struct TileSet {
    bool enabled = false;
    int num_tiles = 0;
    int max_nodes = 0, max_nt = 0, max_cp = 0;
    int slots = 1;          // owned non-terminal nodes per thread (template parameter R)
    bool wide = false;      // some information set is wider than the register-cached strategy row
    size_t smem_bytes = 0;
    int block = 512;
    rlp::DevBuf<TileDesc> desc;
    rlp::DevBuf<unsigned short> ntof, term_lid;
    rlp::DevBuf<TileRec> rec;
    rlp::DevBuf<TileAux> aux;
    rlp::DevBuf<double> pc, pay1, pay2, cprob, root_val;
    int64_t bytes() const {
        return (int64_t)(desc.bytes() + ntof.bytes() + term_lid.bytes() + rec.bytes() + aux.bytes() + pc.bytes() +
                         pay1.bytes() + pay2.bytes() + cprob.bytes() + root_val.bytes());
    }
};
