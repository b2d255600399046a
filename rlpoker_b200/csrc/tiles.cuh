// Two-tier tiling of the tree below an all-chance trunk.
//
//  * MICRO subtrees: maximal subtrees with <= 32 nodes, <= 8 decision nodes, no chance node (in Leduc: every
//    23-node betting round after the board card).  One THREAD sweeps one micro subtree, forward and backward,
//    with no barrier at all; its scratch (reach and values of <= 8 nodes) sits in a thread-interleaved slice of
//    shared memory.  Subtrees of equal shape share one template; per-instance data (information-set columns,
//    contribution slots, payoffs) is stored slot-major so a warp reads consecutive words.
//  * UPPER tiles: what remains above the micro subtrees is cut into maximal subtrees that fit one CTA's shared
//    memory (in Leduc: the round-1 betting nodes and board-deal chance nodes of one (hole1, hole2) deal); the
//    micro roots appear in them as "portal" leaves.  Phase 0 sweeps reach down to the portals, the micro kernel
//    runs, phase 1 redoes the (tiny) forward pass and sweeps values up from the portals.
//
// Reach and node values never touch HBM except at the portals (24 B per micro subtree); only the per-decision-node
// contributions (regret terms, own reach) leave the SM, into the jagged-diagonal scratch the accumulate kernel
// streams.
#pragma once
#include <memory>
#include <string>

#include "common.cuh"

constexpr int kMaxTileLevels = 32;
constexpr int kMicroNodes = 32;     // nodes per micro subtree
constexpr int kMicroNt = 8;         // decision nodes per micro subtree
constexpr int kMicroAct = 8;        // widest node inside a micro subtree
constexpr int kMicroSmemClasses = 12;

struct TileDesc {
    long long node_off;     // into ntof
    long long nt_off;       // into rec / aux / pc
    long long term_off;     // into term_lid / pay
    long long cp_off;       // into cprob
    long long portal_off;   // first portal (= micro instance) index of this tile
    int n_nodes, n_nt, n_term, n_levels;
    int root_global;        // breadth-first id of the tile root
    int n_cp;               // chance edges inside the tile
    int n_portal, pad;
    double root_p1, root_p2;            // strategy reach of the root (1.0 below an all-chance trunk)
    int lvl_nt[kMaxTileLevels + 1];     // non-terminal rank ranges per local level
    int pad2;
};

struct TileRec {            // 8 bytes
    unsigned short first_child, lid;
    unsigned char nchild, player;
    unsigned short cp_off;
};

struct TileAux {            // 16 bytes
    int colbase, slot_r, slot_o, pad;
};

// Shape of a micro subtree in local breadth-first order.
struct MicroTmpl {
    unsigned char n, n_nt, n_term, pad;
    unsigned char player[kMicroNodes];
    unsigned char first_child[kMicroNodes];
    unsigned char nchild[kMicroNodes];
    unsigned char rank[kMicroNodes];        // non-terminal rank, or 0x80 | terminal rank
    unsigned char nt_node[kMicroNt];        // local id of the r-th non-terminal
    unsigned char par_rank[kMicroNodes];    // non-terminal rank of the parent
    unsigned char par_slot[kMicroNodes];    // which child of the parent
};

// Shape of an upper tile in local breadth-first order (host side; input of the straight-line generator).
struct UpperShape {
    // -1 terminal, 0 chance, 1 / 2 decision node of that player, 3 portal leaf,
    // 4 "fan-in" leaf: a chance node whose children are all portals (summed by k_fanin, not by the tile kernel)
    std::vector<signed char> kind;
    std::vector<int> first_child, nchild;
    std::vector<int> portal_rank;       // kind 3: rank of the portal inside its tile
    std::vector<int> fan_index, fan_k0; // kind 4: index of the fan-in node, portal rank of its first child
};

struct FanDesc { int k0, nchild, e0, pad; };

// All upper tiles share one shape and were compiled to straight-line code: one thread per tile.
struct UpperJit {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k0 = nullptr, k1 = nullptr;
    long long T = 0;
    rlp::DevBuf<int> colbase, slot_r, slot_o;       // [decision][T]
    rlp::DevBuf<double> pcs, cprob, pay1, pay2;     // [decision][T], [chance edge][T], [terminal][T]
    rlp::DevBuf<long long> poff;                    // [T]
    int num_fan = 0;                                // fan-in chance nodes per tile
    rlp::DevBuf<FanDesc> fan;                       // [num_fan]
    rlp::DevBuf<double> fan_cp;                     // [fan chance edge][T]
    rlp::DevBuf<double> fan_v1, fan_v2;             // [num_fan][T]
    ~UpperJit() { if (lib) cudaLibraryUnload(lib); }
};

// A shape class whose sweep was compiled to straight-line code (jit.cu).
struct JitClass {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    int count = 0;
    rlp::DevBuf<int> inst;      // instance (= portal) indices of this class; empty when the class covers 0..P-1
    ~JitClass() { if (lib) cudaLibraryUnload(lib); }
};

// Group-fused persistent iteration kernel (fused.cu): device tables + the compiled kernel.
struct FusedPlan;
void fused_plan_free(FusedPlan *);
struct FusedPlanDeleter { void operator()(FusedPlan *p) const { fused_plan_free(p); } };

struct TileSet {
    bool enabled = false;
    std::unique_ptr<FusedPlan, FusedPlanDeleter> fused;   // set when the tree qualifies (see fused.cu)
    // upper tiles
    int num_tiles = 0;
    int max_nodes = 0, max_nt = 0, max_cp = 0;
    int slots = 1;          // owned non-terminal nodes per thread (template parameter R)
    bool wide = false;      // some information set is wider than the register-cached strategy row
    size_t smem_bytes = 0;
    int block = 128;
    rlp::DevBuf<TileDesc> desc;
    rlp::DevBuf<unsigned short> ntof, term_lid, portal_lid;
    rlp::DevBuf<TileRec> rec;
    rlp::DevBuf<TileAux> aux;
    rlp::DevBuf<double> pc, pay1, pay2, cprob, root_val;
    // micro subtrees (instance p == portal p)
    int num_micro = 0, num_classes = 0, micro_nt = 0, micro_term = 0, micro_nodes = 0;
    size_t micro_smem = 0;
    rlp::DevBuf<MicroTmpl> tmpl;
    rlp::DevBuf<unsigned short> m_class;            // [P]
    rlp::DevBuf<int> m_colbase, m_slot_r, m_slot_o; // [micro_nt][P]
    rlp::DevBuf<int> m_pidx;                        // [P] index of the instance's portal in the portal arrays
    rlp::DevBuf<int> m_ridx;                        // [P] where to read its root reach (the chance parent's slot when it has one)
    rlp::DevBuf<double> m_pay1, m_pay2;             // [micro_term][P]
    rlp::DevBuf<double> m_pc;                       // [P]
    rlp::DevBuf<double> portal_p1, portal_p2, portal_v1, portal_v2;   // [P]
    std::vector<std::unique_ptr<JitClass>> jit;     // one per class when every class compiled, else empty
    rlp::DevBuf<int> acc_list_a, acc_list_b;        // internal information-set ids: only-micro / touching upper tiles
    int acc_a_count = 0, acc_b_count = 0;
    std::unique_ptr<UpperJit> upper_jit;            // set when all upper tiles share one compiled shape
    int64_t bytes() const {
        return (int64_t)(desc.bytes() + ntof.bytes() + term_lid.bytes() + portal_lid.bytes() + rec.bytes() +
                         aux.bytes() + pc.bytes() + pay1.bytes() + pay2.bytes() + cprob.bytes() + root_val.bytes() +
                         tmpl.bytes() + m_class.bytes() + m_pidx.bytes() + m_ridx.bytes() + m_colbase.bytes() + m_slot_r.bytes() + m_slot_o.bytes() +
                         m_pay1.bytes() + m_pay2.bytes() + m_pc.bytes() + portal_p1.bytes() + portal_p2.bytes() +
                         portal_v1.bytes() + portal_v2.bytes());
    }
};
