// Group-fused persistent CFR iteration (replaces, for trees that qualify, the five-kernel pipeline of cfr.cu /
// tiles.cu: upper_k0 -> micro_k -> k_fanin -> upper_k1 -> k_accumulate_rm4 x2).  Same arithmetic, operand for
// operand, as those kernels and therefore as the reference's cfr_recursive (rlpoker/cfr/cfr.py:158-255) and
// compute_regret_matching (rlpoker/cfr/cfr_util.py:14-58): regrets, action counts and strategies stay bit-identical.
//
// Idea.  All nodes of an information set of player p inside the micro subtrees differ only in what p cannot see, so
// the micro subtrees fall into GROUPS per player: the G subtrees of a group share ALL of p's information sets, and the
// position of a subtree inside the group is the rank of its nodes inside those information sets (ascending node id =
// the order in which the reference's recursion visits them).  One thread evaluates one subtree (straight-line code
// generated for the shape, NVRTC), the G threads of a group leave p's regret terms in shared memory, and one thread
// per (information set, action) adds them IN MEMBER ORDER onto the regret and applies regret matching -- no
// contribution scratch in HBM/L2, no separate accumulate kernel, no atomics.  Every subtree is evaluated twice (once
// in a player-1 group, once in a player-2 group); that costs ~1.3x the arithmetic and removes ~100 MB of traffic per
// iteration at Leduc-13.  The upper tiles (round-1 betting of one deal) are grouped the same way.
//
// One cooperative persistent kernel runs all iterations of a call:
//     prologue  reach of both players from the tile roots to the portals            | grid barrier
//     phase A   micro groups of both players: evaluate, accumulate, regret matching | grid barrier
//     phase B   upper groups: fan-in of the portal values, evaluate the tile, accumulate + regret matching of the
//               round-1 information sets, then the reach of the group's player to the portals for the NEXT
//               iteration (it depends only on that player's own new strategy)      | grid barrier
// The strategy is double buffered by iteration parity (a group rewrites its player's strategy while groups of the
// other player still read the old one) in a "fused" information-set numbering chosen so that the opponent-strategy
// gathers of a group hit consecutive words.
#include <dlfcn.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <sstream>
#include <string>

#include "cfr.cuh"
#include "fused.cuh"

using rlp::fail;

namespace rlp {
int compile_and_load_source(const std::string &src, cudaLibrary_t *lib_out, std::string *log_out);
}

// ---- parameter block shared with the generated source (every field is 8 bytes: no padding questions) ----------
#define RLP_FUSED_FIELDS(X)                                                                                        \
    X(const int *, mem_opp0) X(const int *, mem_opp1)         /* [opp decision][slot] fused ids of the opponent's sets */ \
    X(const int *, mem_pat0) X(const int *, mem_pat1)         /* [slot] payoff pattern */                          \
    X(const int *, mem_ridx0) X(const int *, mem_ridx1)       /* [slot] where the OPPONENT's reach of the subtree root lives */ \
    X(const int *, grp_ridx0) X(const int *, grp_ridx1)       /* [group] where the group's own player's reach lives */   \
    X(const int *, mem_pidx0) X(const int *, mem_pidx1)       /* [slot] where its value goes (tile * KP + portal rank) */ \
    X(const double *, mem_pc0) X(const double *, mem_pc1)     /* [slot] static chance reach */                      \
    X(const int *, grp_fid0) X(const int *, grp_fid1)         /* [group][own decision] fused id */                   \
    X(const int *, grp_col0) X(const int *, grp_col1)         /* [group][own decision] first internal pair */        \
    X(const int *, grp_iset0) X(const int *, grp_iset1)       /* [group][own decision] internal information set */   \
    X(const double *, paytab)                                 /* [pattern][leaf] */                                 \
    X(const signed char *, paytab8)                           /* [pattern][16-byte rows]: small integer payoffs */    \
    X(const int *, tg_tile0) X(const int *, tg_tile1)         /* [upper group][member] tile */                      \
    X(const int *, tg_fid0) X(const int *, tg_fid1)           /* [upper group][own decision] */                     \
    X(const int *, tg_col0) X(const int *, tg_col1)                                                                \
    X(const int *, tg_iset0) X(const int *, tg_iset1)                                                              \
    X(const int *, u_fid)                                     /* [decision][tile] fused id */                       \
    X(const double *, u_pcs) X(const double *, u_cprob) X(const double *, u_pay1)                                   \
    X(const int *, fan_desc) X(const double *, fan_cp)        /* FanDesc[num_fan] as 4 ints; [fan edge][tile] */     \
    X(double *, R) X(double *, S) X(double *, sigma_row)      /* solver state, internal pair order */               \
    X(double *, sigF0) X(double *, sigF1)                     /* strategy, [action][fused id], two buffers */        \
    X(double *, portal_p1) X(double *, portal_p2)                                                                   \
    X(double *, pvl0) X(double *, pvl1)                       /* subtree values, this GPU, by iteration parity: [pass][upper group][portal][member] */ \
    X(double *, pvA0) X(double *, pvA1) X(double *, pvA2) X(double *, pvA3) X(double *, pvA4) X(double *, pvA5)     \
    X(double *, pvA6) X(double *, pvA7)                       /* even-iteration buffer of every rank (peer mapped) */ \
    X(double *, pvB0) X(double *, pvB1) X(double *, pvB2) X(double *, pvB3) X(double *, pvB4) X(double *, pvB5)     \
    X(double *, pvB6) X(double *, pvB7)                       /* odd-iteration buffer of every rank */               \
    X(unsigned long long *, xfp0) X(unsigned long long *, xfp1) X(unsigned long long *, xfp2)                        \
    X(unsigned long long *, xfp3) X(unsigned long long *, xfp4) X(unsigned long long *, xfp5)                        \
    X(unsigned long long *, xfp6) X(unsigned long long *, xfp7) /* flag arrays of every rank (peer mapped) */        \
    X(unsigned long long *, xfl)                              /* this rank's flag array: xfl[r] written by rank r */  \
    X(long long, nranks) X(long long, rank) X(long long, xbase)                                                     \
    X(unsigned char *, flags) X(unsigned int *, bar) X(unsigned long long *, tl)                                    \
    X(const double *, br_sig)                                 /* best response: strategy, [action][fused id] */      \
    X(double *, br_reach1) X(double *, br_reach2)             /* reach of player 1 / 2 to the portals under br_sig */  \
    X(double *, br_pv1) X(double *, br_pv2)                   /* subtree values for responder 1 / 2 */                \
    X(double *, br_top1) X(double *, br_top2)                 /* values of the trunk nodes and tile roots, by node id */ \
    X(double *, br_out) X(int *, br_astar)                    /* {BR1, BR2}; chosen action per internal information set */ \
    X(const int *, tile_root) X(const int *, g_first_child) X(const signed char *, g_player) X(long long, n_top)       \
    X(long long, n_tl) X(long long, tl0) X(long long, tl1) X(long long, tl2) X(long long, tl3) X(long long, tl4)       \
    X(long long, tl5) X(long long, tl6) X(long long, tl7) X(long long, tl8) /* node ranges of the top levels */        \
    X(long long, num_isets) X(long long, first_group0) X(long long, end_group0) X(long long, first_group1)          \
    X(long long, end_group1) X(long long, tl_cap)

struct FusedParams {
#define X(type, name) type name;
    RLP_FUSED_FIELDS(X)
#undef X
};
static_assert(sizeof(FusedParams) % 8 == 0, "8-byte fields only");

#define RLP_STR2(x) #x
#define X(type, name) RLP_STR2(type) " " #name ";\n"
static const char *kFusedFieldsText = RLP_FUSED_FIELDS(X);
#undef X

struct FusedPlan {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    FusedParams host{};                 // tree part filled; solver part (R, S, ...) filled per solver
    rlp::DevBuf<int> mem_opp[2], mem_pat[2], mem_ridx[2], mem_pidx[2], grp_fid[2], grp_col[2], grp_iset[2], grp_ridx[2];
    rlp::DevBuf<double> mem_pc[2], paytab, reach[2];
    rlp::DevBuf<signed char> paytab8;
    rlp::DevBuf<int> tg_tile[2], tg_fid[2], tg_col[2], tg_iset[2], u_fid;
    rlp::DevBuf<int> fid_of_iset;       // [I] internal -> fused
    rlp::DevBuf<int> tile_root;         // [T] node id of every tile root
    cudaKernel_t br_kernel = nullptr;   // best response on the groups (null when the tree does not qualify)
    long long n_top = 0;                // nodes 0 .. n_top-1 are the all-chance trunk and the tile roots
    int kp = 0, nug[2] = {0, 0};
    rlp::DevBuf<unsigned int> bar;
    long long NF = 0;
    int n_batches[2] = {0, 0}, n_ugroups[2] = {0, 0};
    int max_act = 0;
    int grid = 0, block = 256;
    int num_components = 0;
    std::vector<int> group_comp[2];     // component of every group, fused order (multi-GPU sharding by public state)
    std::vector<int> h_grp_iset[2], h_tg_iset[2];   // host copies: internal information sets of the groups
    int mo[2] = {0, 0};                 // own decisions per micro group
    size_t pv_len = 0;                  // doubles in one pass's subtree-value array (T * KP)
    rlp::DevBuf<double> pvbuf;          // single GPU: [2 passes][T * KP]
    std::string source;
    ~FusedPlan() { if (lib) cudaLibraryUnload(lib); }
};
void fused_plan_free(FusedPlan *p) { delete p; }

namespace {

constexpr int kMaxBlock = 256;     // upper bound of the CTA size (shared-memory arrays are sized from the actual one)

// ------------------------------------------------------------------------------------------------ source generator
struct Shape {
    // micro
    const MicroTmpl *tm;
    std::vector<int> dec[2];            // decision ranks of player 1 / 2
    // upper
    const UpperShape *sh;
    std::vector<int> udec[2];           // decision indices (shape order among decision nodes) of player 1 / 2
    std::vector<int> dec_idx, term_idx, edge_idx, parent, slot;
    int n_udec = 0, n_uterm = 0, n_uedge = 0;
};

struct Consts {
    int BLK, MINB;                      // CTA size, resident CTAs per SM asked of the compiler
    int TEAM;                           // threads that work on one micro unit together: 32 (a warp, __syncwarp) or BLK
    int NPUSH;                          // CTAs that copy this rank's subtree values to the peers (multi-GPU)
    int BR_OK, MAXA;                    // the best-response kernel is generated too; widest decision node
    int PREFETCH;                       // micro phase: fetch the next unit's static indices while the current one runs
    int XFENCE_ALL;                     // multi-GPU: every CTA fences its peer stores at system scope (else only the last arriver)
    int G[2], GPC[2], NG[2], GB[2], NUG[2];
    long long NS[2];                    // slots = NG * G
    long long NF, T;
    int NT, KP, F;
    int NCH;                            // children of every fan-in node when they all agree, else 0
    int PAY8;                           // payoffs are integers in [-127, 127]: 1-byte pattern table, rows of PAYROW bytes
    int PAYROW;
    std::vector<double> fan_cps;        // one chance probability per fan-in node when it is the same for all edges / tiles
    int sm_r_n, sm_o_n;                 // doubles in the kernel's sm_r / sm_o arrays (contiguous: the fan-in stages through both)
    long long PVLEN;                    // doubles in one pass's subtree-value array (T * KP)
    std::vector<FanDesc> fans;
    int GBS, MBS, UT_OWN, UT_N;         // layout of the cached per-unit tables (ints): strides and the offset of the own-decision part
};

std::string S_(const char *p, int j) { return std::string(p) + std::to_string(j); }

// (own decision jj, action a) enumeration of a player's regret terms
struct Terms { std::vector<int> cj, ca, coff, na; int kr = 0; };

Terms micro_terms(const Shape &s, int pl) {
    Terms t;
    for (size_t jj = 0; jj < s.dec[pl].size(); ++jj) {
        int node = s.tm->nt_node[s.dec[pl][jj]];
        t.coff.push_back(t.kr);
        t.na.push_back(s.tm->nchild[node]);
        for (int a = 0; a < s.tm->nchild[node]; ++a) { t.cj.push_back((int)jj); t.ca.push_back(a); ++t.kr; }
    }
    return t;
}

Terms upper_terms(const Shape &s, int pl) {
    Terms t;
    const UpperShape &sh = *s.sh;
    std::vector<int> node_of_dec;
    for (size_t j = 0; j < sh.kind.size(); ++j) if (sh.kind[j] == 1 || sh.kind[j] == 2) node_of_dec.push_back((int)j);
    for (size_t jj = 0; jj < s.udec[pl].size(); ++jj) {
        int node = node_of_dec[s.udec[pl][jj]];
        t.coff.push_back(t.kr);
        t.na.push_back(sh.nchild[node]);
        for (int a = 0; a < sh.nchild[node]; ++a) { t.cj.push_back((int)jj); t.ca.push_back(a); ++t.kr; }
    }
    return t;
}

void emit_table(std::ostringstream &o, const char *name, int pl, const std::vector<int> &v) {
    o << "__constant__ int " << name << pl << "[" << std::max<size_t>(1, v.size()) << "] = {";
    for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
    if (v.empty()) o << "0";
    o << "};\n";
}

// One micro subtree, evaluated by one thread for the groups of player `pl` (0-based).
void gen_micro_eval(std::ostringstream &o, const Shape &s, const Consts &c, int pl) {
    const MicroTmpl &tm = *s.tm;
    const int n = tm.n, n_nt = tm.n_nt;
    const int opp = 1 - pl;
    std::vector<int> own_j(n_nt, -1), opp_j(n_nt, -1);
    for (size_t j = 0; j < s.dec[pl].size(); ++j) own_j[s.dec[pl][j]] = (int)j;
    for (size_t j = 0; j < s.dec[opp].size(); ++j) opp_j[s.dec[opp][j]] = (int)j;
    Terms t = micro_terms(s, pl);
    // Static per-subtree / per-group indices: fetched one unit AHEAD by the caller (MIds), so the strategy and reach loads
    // below can issue as soon as the unit starts.
    o << "__device__ __forceinline__ MIds micro_ids" << pl << "(const FusedParams* __restrict__ P, long long g, int k, bool ok,\n"
         "    long long ag, int aj, bool acc) {\n"
         "  MIds id;\n"
         "  const long long slot = ok ? g * " << c.G[pl] << " + k : 0;\n"
         "  const long long gg = ok ? g : 0;\n"
         "  id.pat = P->mem_pat" << pl << "[slot]; id.ri = P->mem_ridx" << pl << "[slot]; id.ro = P->grp_ridx" << pl << "[gg];\n"
         "  id.pcv = P->mem_pc" << pl << "[slot];\n"
         "  id.pi = P->mem_pidx" << pl << "[slot];\n";
    for (int r = 0; r < n_nt; ++r) {
        if (own_j[r] >= 0)
            o << "  id.f[" << r << "] = P->grp_fid" << pl << "[gg * " << s.dec[pl].size() << " + " << own_j[r] << "];\n";
        else
            o << "  id.f[" << r << "] = P->mem_opp" << pl << "[" << opp_j[r] << "LL * " << c.NS[pl] << "LL + slot];\n";
    }
    o << "  const long long agg = acc ? ag : 0;\n"
         "  id.acol = P->grp_col" << pl << "[agg * " << s.dec[pl].size() << " + aj];\n"
         "  id.afid = P->grp_fid" << pl << "[agg * " << s.dec[pl].size() << " + aj];\n"
         "  return id;\n}\n";
    o << "__device__ __forceinline__ void micro_eval" << pl << "(const FusedParams* __restrict__ P, const MIds& id, int tix, double w,\n"
         "    const double* sig_in, double* sm_r, double* sm_o, int par) {\n"
         "  const int pat = id.pat, ri = id.ri, ro = id.ro;\n"
         "  const double pcv = id.pcv;\n";
    o << "  const int pi = id.pi;\n";
    for (int r = 0; r < n_nt; ++r) o << "  const int f" << r << " = id.f[" << r << "];\n";
    if (c.PAY8) {           // one or two 16-byte loads fetch all leaf payoffs of the subtree's pattern (exact: small integers)
        o << "  const int4* pp = reinterpret_cast<const int4*>(P->paytab8 + (long long)pat * " << c.PAYROW << ");\n";
        for (int q = 0; q < c.PAYROW / 16; ++q) o << "  const int4 pw" << q << " = __ldg(pp + " << q << ");\n";
        for (int j = 0; j < n; ++j)
            if (tm.rank[j] & 0x80) {
                int tr = tm.rank[j] & 0x7F, q = tr / 16, wd = (tr % 16) / 4, by = tr % 4;
                o << "  const double x" << j << " = (double)(int)(signed char)(pw" << q << "." << "xyzw"[wd] << " >> " << 8 * by << ");\n";
            }
    } else {
        o << "  const double* pp = P->paytab + (long long)pat * " << c.NT << ";\n";
        for (int j = 0; j < n; ++j)
            if (tm.rank[j] & 0x80) o << "  const double x" << j << " = pp[" << (tm.rank[j] & 0x7F) << "];\n";
    }
    // a player's reach of a portal depends only on that player's own upper group: one value per (reach slot, group)
    if (pl == 0) o << "  const double q1_0 = __ldcg(P->portal_p1 + ro), q2_0 = __ldcg(P->portal_p2 + ri);\n";
    else o << "  const double q1_0 = __ldcg(P->portal_p1 + ri), q2_0 = __ldcg(P->portal_p2 + ro);\n";
    for (int j = 1; j < n; ++j)
        o << "  const double s" << j << " = __ldcg(sig_in + " << (int)tm.par_slot[j] << "LL * " << c.NF << "LL + f" << (int)tm.par_rank[j] << ");\n";
    for (int r = 0; r < n_nt; ++r) {                    // forward (cfr.py:208,217)
        int j = tm.nt_node[r], p = tm.player[j];
        for (int k = 0; k < tm.nchild[j]; ++k) {
            int ch = tm.first_child[j] + k;
            if (tm.rank[ch] & 0x80) continue;
            int rc = tm.rank[ch];
            o << "  const double q1_" << rc << " = " << (p == 1 ? S_("s", ch) + " * " : std::string()) << "q1_" << r << ";\n";
            o << "  const double q2_" << rc << " = " << (p == 2 ? S_("s", ch) + " * " : std::string()) << "q2_" << r << ";\n";
        }
    }
    for (int r = n_nt - 1; r >= 0; --r) {               // backward (cfr.py:204-242)
        int j = tm.nt_node[r], p = tm.player[j], fc = tm.first_child[j], nc = tm.nchild[j];
        o << "  double a" << j << " = 0.0;\n";
        for (int k = 0; k < nc; ++k) o << "  a" << j << " = a" << j << " + s" << (fc + k) << " * x" << (fc + k) << ";\n";
        o << "  const double x" << j << " = a" << j << ";\n";
        if (p != pl + 1) continue;
        int jj = own_j[r];
        o << "  {\n    const double opp = " << (p == 1 ? "pcv * q2_" : "pcv * q1_") << r << ";\n"
          << "    const double own = " << (p == 1 ? "q1_" : "q2_") << r << ";\n"
          << "    const double vp = " << (p == 1 ? S_("x", j) : "-" + S_("x", j)) << ";\n";
        for (int k = 0; k < nc; ++k) {
            std::string va = p == 1 ? S_("x", fc + k) : "-" + S_("x", fc + k);
            o << "    sm_r[" << (t.coff[jj] + k) << " * SMS + tix] = w * (" << va << " - vp) * opp;\n";
        }
        o << "    sm_o[" << jj << " * SMS + tix] = w * own;\n  }\n";
    }
    // The subtree value goes to this pass's array (laid out for this player's upper groups: [group][portal][member], so
    // the G threads of a group store consecutive words) in the buffer of this iteration's parity -- of EVERY rank: the
    // peers' buffers are mapped (CUDA IPC), the stores travel over NVLink while the phase is still running.
    o << "  { double* const* pvp = par ? &P->pvB0 : &P->pvA0; const int nr = (int)P->nranks;\n"
         "    for (int r = 0; r < nr; ++r) pvp[r][pi] = x0; }\n";
    o << "}\n";
}

// One upper tile, evaluated by one thread for the upper groups of player `pl`: values up from the portals / fan-in
// sums, contributions of pl's decision nodes into shared memory.
void gen_upper_eval(std::ostringstream &o, const Shape &s, const Consts &c, int pl) {
    const UpperShape &sh = *s.sh;
    const int n = (int)sh.kind.size();
    Terms t = upper_terms(s, pl);
    std::vector<int> own_jj(s.n_udec, -1);
    for (size_t jj = 0; jj < s.udec[pl].size(); ++jj) own_jj[s.udec[pl][jj]] = (int)jj;
    // everything the tile needs from global memory EXCEPT the fan-in sums: loaded at the start of the unit so that the
    // loads overlap the fan-in instead of following it
    o << "struct UIn" << pl << " {\n";
    for (int j = 0; j < n; ++j) {
        if (s.dec_idx[j] >= 0 && sh.kind[j] == pl + 1) o << "  double pc" << j << ";\n";
        if (s.edge_idx[j] >= 0) o << "  double e" << j << ";\n";
        if (s.term_idx[j] >= 0) o << "  double x" << j << ";\n";
        if (j >= 1 && s.parent[j] >= 0 && s.dec_idx[s.parent[j]] >= 0) o << "  double s" << j << ";\n";
        if (sh.kind[j] == 3) o << "  double x" << j << ";\n";
    }
    o << "  double pad_;\n};\n"
         "__device__ __forceinline__ void upper_load" << pl << "(const FusedParams* __restrict__ P, long long i, int m, const int* sm_ut, const double* sig_in, const double* blk, UIn" << pl << "& in) {\n";
    for (int j = 0; j < n; ++j)
        if (s.dec_idx[j] >= 0) {
            o << "  const int f" << j << " = sm_ut[" << c.GBS + s.dec_idx[j] * c.GBS << " + m];\n";
            if (sh.kind[j] == pl + 1) o << "  in.pc" << j << " = P->u_pcs[" << s.dec_idx[j] << "LL * " << c.T << "LL + i];\n";
        }
    for (int j = 0; j < n; ++j) {
        if (s.edge_idx[j] >= 0) o << "  in.e" << j << " = P->u_cprob[" << s.edge_idx[j] << "LL * " << c.T << "LL + i];\n";
        if (s.term_idx[j] >= 0) o << "  in.x" << j << " = P->u_pay1[" << s.term_idx[j] << "LL * " << c.T << "LL + i];\n";
    }
    for (int j = 1; j < n; ++j)
        if (s.parent[j] >= 0 && s.dec_idx[s.parent[j]] >= 0)
            o << "  in.s" << j << " = __ldcg(sig_in + " << s.slot[j] << "LL * " << c.NF << "LL + f" << s.parent[j] << ");\n";
    for (int j = 0; j < n; ++j)
        if (sh.kind[j] == 3) o << "  in.x" << j << " = __ldcg(blk + " << sh.portal_rank[j] * c.GB[pl] << " + m);\n";
    o << "  in.pad_ = 0.0;\n}\n";
    o << "__device__ __forceinline__ void upper_eval" << pl << "(int m, double w, const UIn" << pl << "& in, const double* sm_fan, double* sm_r, double* sm_o) {\n";
    for (int j = 0; j < n; ++j) {
        if (s.dec_idx[j] >= 0 && sh.kind[j] == pl + 1) o << "  const double pc" << j << " = in.pc" << j << ";\n";
        if (s.edge_idx[j] >= 0) o << "  const double e" << j << " = in.e" << j << ";\n";
        if (s.term_idx[j] >= 0 || sh.kind[j] == 3) o << "  const double x" << j << " = in.x" << j << ";\n";
        if (j >= 1 && s.parent[j] >= 0 && s.dec_idx[s.parent[j]] >= 0) o << "  const double s" << j << " = in.s" << j << ";\n";
        if (sh.kind[j] == 4) o << "  const double x" << j << " = sm_fan[m * " << c.F << " + " << sh.fan_index[j] << "];\n";
    }
    o << "  const double q1_0 = 1.0, q2_0 = 1.0;\n";
    for (int j = 0; j < n; ++j) {                       // forward: reach of the tile's own decision nodes
        int kd = sh.kind[j];
        if (kd == -1 || kd == 3 || kd == 4) continue;
        for (int k = 0; k < sh.nchild[j]; ++k) {
            int ch = sh.first_child[j] + k;
            if (sh.kind[ch] == -1 || sh.kind[ch] == 3 || sh.kind[ch] == 4) continue;
            o << "  const double q1_" << ch << " = " << (kd == 1 ? S_("s", ch) + " * " : std::string()) << S_("q1_", j) << ";\n"
              << "  const double q2_" << ch << " = " << (kd == 2 ? S_("s", ch) + " * " : std::string()) << S_("q2_", j) << ";\n";
        }
    }
    for (int j = n - 1; j >= 0; --j) {                  // backward
        int kd = sh.kind[j];
        if (kd == -1 || kd == 3 || kd == 4) continue;
        int fc = sh.first_child[j], nc = sh.nchild[j];
        o << "  double a" << j << " = 0.0;\n";
        for (int k = 0; k < nc; ++k)
            o << "  a" << j << " = a" << j << " + " << (kd == 0 ? S_("e", fc + k) : S_("s", fc + k)) << " * x" << (fc + k) << ";\n";
        o << "  const double x" << j << " = a" << j << ";\n";
        if (kd != pl + 1) continue;
        int jj = own_jj[s.dec_idx[j]];
        o << "  {\n    const double opp = pc" << j << (kd == 1 ? " * q2_" : " * q1_") << j << ";\n"
          << "    const double own = " << (kd == 1 ? "q1_" : "q2_") << j << ";\n"
          << "    const double vp = " << (kd == 1 ? S_("x", j) : "-" + S_("x", j)) << ";\n";
        for (int k = 0; k < nc; ++k) {
            std::string va = kd == 1 ? S_("x", fc + k) : "-" + S_("x", fc + k);
            o << "    sm_r[" << (t.coff[jj] + k) << " * " << c.GB[pl] << " + m] = w * (" << va << " - vp) * opp;\n";
        }
        o << "    sm_o[" << jj << " * " << c.GB[pl] << " + m] = w * own;\n  }\n";
    }
    o << "  (void)x0; (void)q1_0; (void)q2_0;\n}\n";
}

// Reach of player `pl` from the tile root to every portal / fan-in node of one tile, from pl's strategy rows in
// shared memory (sm_s[term index] = probability of that (own decision, action)).
void gen_upper_reach(std::ostringstream &o, const Shape &s, const Consts &c, int pl) {
    const UpperShape &sh = *s.sh;
    const int n = (int)sh.kind.size();
    Terms t = upper_terms(s, pl);
    std::vector<int> own_jj(s.n_udec, -1);
    for (size_t jj = 0; jj < s.udec[pl].size(); ++jj) own_jj[s.udec[pl][jj]] = (int)jj;
    o << "__device__ __forceinline__ void upper_reach" << pl << "(double* out, int u, const double* sm_s) {\n"
         "  const double q_0 = 1.0;\n";
    for (int j = 0; j < n; ++j) {
        int kd = sh.kind[j];
        if (kd == -1 || kd == 3 || kd == 4) continue;
        for (int k = 0; k < sh.nchild[j]; ++k) {
            int ch = sh.first_child[j] + k;
            if (sh.kind[ch] == -1) continue;
            std::string v = kd == pl + 1 ? "sm_s[" + std::to_string(t.coff[own_jj[s.dec_idx[j]]] + k) + "] * " + S_("q_", j) : S_("q_", j);
            if (sh.kind[ch] == 4) o << "  out[" << sh.fan_k0[ch] * c.NUG[pl] << " + u] = " << v << ";\n";
            else if (sh.kind[ch] == 3) o << "  out[" << sh.portal_rank[ch] * c.NUG[pl] << " + u] = " << v << ";\n";
            else o << "  const double q_" << ch << " = " << v << ";\n";
        }
    }
    o << "}\n";
}

// ---- best response (rlpoker/best_response.py:9-141) on the same groups ---------------------------------------------------
// Responder p's information sets are exactly the groups of p: the G threads of a group hold the G nodes of every
// information set of the group, so Q[I, a] = sum over the nodes of I (ascending) of v(h.a) is an ordered sum over the
// group's threads through shared memory, the arg-max is taken redundantly by every thread, and v(h) = v(h.a*).
// v(z) = reach(z) * u_p(z) with reach = static chance reach x the opponent's strategy along the path (the responder's
// own actions count 1), as in csrc/br.cu; the association of the reach products differs from the level sweep's, so
// values agree with it (and with the reference) to a few ulp, not bit for bit -- the tests hold 1e-12.
void emit_payoffs(std::ostringstream &o, const MicroTmpl &tm, const Consts &c) {
    const int n = tm.n;
    if (c.PAY8) {
        o << "  const int4* pp = reinterpret_cast<const int4*>(P->paytab8 + (long long)pat * " << c.PAYROW << ");\n";
        for (int q = 0; q < c.PAYROW / 16; ++q) o << "  const int4 pw" << q << " = __ldg(pp + " << q << ");\n";
        for (int j = 0; j < n; ++j)
            if (tm.rank[j] & 0x80) {
                int tr = tm.rank[j] & 0x7F, q = tr / 16, wd = (tr % 16) / 4, by = tr % 4;
                o << "  const double x" << j << " = (double)(int)(signed char)(pw" << q << "." << "xyzw"[wd] << " >> " << 8 * by << ");\n";
            }
    } else {
        o << "  const double* pp = P->paytab + (long long)pat * " << c.NT << ";\n";
        for (int j = 0; j < n; ++j)
            if (tm.rank[j] & 0x80) o << "  const double x" << j << " = pp[" << (tm.rank[j] & 0x7F) << "];\n";
    }
}

void gen_micro_br(std::ostringstream &o, const Shape &s, const Consts &c, int pl) {
    const MicroTmpl &tm = *s.tm;
    const int n = tm.n, n_nt = tm.n_nt, opp = 1 - pl;
    std::vector<int> own_j(n_nt, -1), opp_j(n_nt, -1);
    for (size_t j = 0; j < s.dec[pl].size(); ++j) own_j[s.dec[pl][j]] = (int)j;
    for (size_t j = 0; j < s.dec[opp].size(); ++j) opp_j[s.dec[opp][j]] = (int)j;
    const int G = c.G[pl], GPT = c.GPC[pl], MO = (int)s.dec[pl].size(), A = c.MAXA;
    o << "__device__ __forceinline__ void micro_br" << pl << "(const FusedParams* __restrict__ P, long long unit, double* sm_q) {\n"
         "  const int tix = threadIdx.x % TEAM;\n"
         "  const int gl = tix / " << G << ", k = tix - gl * " << G << ";\n"
         "  const long long g = unit * " << GPT << " + gl;\n"
         "  const bool ok = gl < " << GPT << " && g < " << c.NG[pl] << "LL;\n"
         "  const int glr = gl < " << GPT << " ? gl : 0;          // column block read by the reductions (idle lanes stay in bounds)\n"
         "  const long long slot = ok ? g * " << G << " + k : 0;\n"
         "  const int pat = P->mem_pat" << pl << "[slot], ri = P->mem_ridx" << pl << "[slot], pi = P->mem_pidx" << pl << "[slot];\n"
         "  const double pcv = P->mem_pc" << pl << "[slot];\n";
    for (int r = 0; r < n_nt; ++r)
        if (opp_j[r] >= 0) o << "  const int f" << r << " = P->mem_opp" << pl << "[" << opp_j[r] << "LL * " << c.NS[pl] << "LL + slot];\n";
    emit_payoffs(o, tm, c);
    o << "  const double rho_0 = pcv * __ldcg(P->br_reach" << opp + 1 << " + ri);\n";
    for (int j = 1; j < n; ++j)
        if (opp_j[tm.par_rank[j]] >= 0)
            o << "  const double s" << j << " = __ldcg(P->br_sig + " << (int)tm.par_slot[j] << "LL * " << c.NF << "LL + f" << (int)tm.par_rank[j] << ");\n";
    for (int r = 0; r < n_nt; ++r) {                    // forward: reach (chance x opponent) of the non-terminal nodes
        int j = tm.nt_node[r];
        for (int kk = 0; kk < tm.nchild[j]; ++kk) {
            int ch = tm.first_child[j] + kk;
            if (tm.rank[ch] & 0x80) continue;
            o << "  const double rho_" << (int)tm.rank[ch] << " = rho_" << r << (opp_j[r] >= 0 ? " * " + S_("s", ch) : std::string()) << ";\n";
        }
    }
    for (int r = n_nt - 1; r >= 0; --r) {               // backward
        int j = tm.nt_node[r], fc = tm.first_child[j], nc = tm.nchild[j];
        for (int kk = 0; kk < nc; ++kk) {
            int ch = fc + kk;
            if (tm.rank[ch] & 0x80)
                o << "  const double v" << ch << " = (rho_" << r << (opp_j[r] >= 0 ? " * " + S_("s", ch) : std::string()) << ") * "
                  << (pl == 0 ? "" : "-") << "x" << ch << ";\n";
            else
                o << "  const double v" << ch << " = V" << (int)tm.rank[ch] << ";\n";
        }
        if (opp_j[r] >= 0) {
            o << "  double V" << r << " = 0.0;\n";
            for (int kk = 0; kk < nc; ++kk) o << "  V" << r << " = V" << r << " + v" << (fc + kk) << ";\n";
        } else {
            const int jj = own_j[r];
            o << "  if (ok) {\n";
            for (int kk = 0; kk < nc; ++kk) o << "    sm_q[" << (jj * A + kk) << " * SMS + tix] = v" << (fc + kk) << ";\n";
            o << "  }\n  TEAM_SYNC();\n  double V" << r << ";\n  {\n    int best = 0; double bq = 0.0;\n";
            for (int kk = 0; kk < nc; ++kk) {
                o << "    { double q = 0.0; const double* col = sm_q + " << (jj * A + kk) << " * SMS + glr * " << G << ";\n"
                     "      for (int m = 0; m < " << G << "; ++m) q = q + col[m];\n"
                     "      if (" << (kk == 0 ? "true" : "q > bq") << ") { bq = q; best = " << kk << "; } }\n";
            }
            o << "    V" << r << " = v" << fc;
            for (int kk = 1; kk < nc; ++kk) o << ";\n    if (best == " << kk << ") V" << r << " = v" << (fc + kk);
            o << ";\n    if (ok && k == 0) P->br_astar[P->grp_iset" << pl << "[g * " << MO << " + " << jj << "]] = best;\n  }\n";
        }
    }
    o << "  if (ok) P->br_pv1[pi] = V0;\n  TEAM_SYNC();\n}\n";
}

// One upper group of the responder: fan-in (plain ordered sums of the subtree values), then every tile bottom-up with the
// responder's round-1 information sets reduced across the tiles of the group.  Executed by the whole CTA (block barriers
// inside); thread m < GB works on tile m.
void gen_upper_br(std::ostringstream &o, const Shape &s, const Consts &c, int pl) {
    const UpperShape &sh = *s.sh;
    const int n = (int)sh.kind.size(), opp = 1 - pl, GB = c.GB[pl], A = c.MAXA, MB = (int)s.udec[pl].size();
    std::vector<int> own_jj(s.n_udec, -1);
    for (size_t jj = 0; jj < s.udec[pl].size(); ++jj) own_jj[s.udec[pl][jj]] = (int)jj;
    o << "__device__ __noinline__ void upper_br" << pl << "(const FusedParams* __restrict__ P, int u, double* sm_fan, double* sm_q, const int* sm_ut) {\n"
         "  const int tid = threadIdx.x;\n"
         "  const double* pvl = P->br_pv1 + " << (long long)pl * c.PVLEN << "LL + (long long)u * " << (long long)c.KP * GB << "LL;\n"
         "  const int* fd = P->fan_desc;\n"
         "  for (int x = tid; x < " << GB * c.F << "; x += BLOCK) {\n"
         "    const int mm = x / " << std::max(1, c.F) << ", f = x - mm * " << std::max(1, c.F) << ";\n"
         "    const int k0 = fd[4 * f], nch = fd[4 * f + 1];\n"
         "    const double* pv = pvl + (long long)k0 * " << GB << " + mm;\n"
         "    double acc = 0.0;\n"
      << (c.NCH > 0 && c.NCH <= 32
              ? "    double vv[" + std::to_string(c.NCH) + "];\n#pragma unroll\n    for (int ch = 0; ch < " + std::to_string(c.NCH) +
                    "; ++ch) vv[ch] = __ldcg(pv + ch * " + std::to_string(GB) + ");\n#pragma unroll\n    for (int ch = 0; ch < " + std::to_string(c.NCH) +
                    "; ++ch) acc = acc + vv[ch];\n    (void)nch;\n"
              : std::string("    for (int ch = 0; ch < nch; ++ch) acc = acc + __ldcg(pv + ch * ") + std::to_string(GB) + ");\n")
      << "    sm_fan[x] = acc;\n"
         "  }\n"
         "  __syncthreads();\n"
         "  const bool act = tid < " << GB << ";\n"
         "  const int m = act ? tid : 0;\n"
         "  const long long i = sm_ut[m];\n";
    for (int j = 0; j < n; ++j)
        if (s.dec_idx[j] >= 0) {
            o << "  const double pc" << j << " = P->u_pcs[" << s.dec_idx[j] << "LL * " << c.T << "LL + i];\n";
            if (sh.kind[j] == opp + 1) o << "  const int f" << j << " = sm_ut[" << c.GBS + s.dec_idx[j] * c.GBS << " + m];\n";
        }
    for (int j = 0; j < n; ++j)
        if (s.term_idx[j] >= 0) o << "  const double x" << j << " = P->u_pay1[" << s.term_idx[j] << "LL * " << c.T << "LL + i];\n";
    for (int j = 1; j < n; ++j)
        if (s.parent[j] >= 0 && sh.kind[s.parent[j]] == opp + 1)
            o << "  const double s" << j << " = __ldcg(P->br_sig + " << s.slot[j] << "LL * " << c.NF << "LL + f" << s.parent[j] << ");\n";
    o << "  const double qo_0 = 1.0;\n";
    for (int j = 0; j < n; ++j) {
        int kd = sh.kind[j];
        if (kd != 1 && kd != 2) continue;
        for (int kk = 0; kk < sh.nchild[j]; ++kk) {
            int ch = sh.first_child[j] + kk;
            if (sh.kind[ch] != 1 && sh.kind[ch] != 2) continue;
            o << "  const double qo_" << ch << " = qo_" << j << (kd == opp + 1 ? " * " + S_("s", ch) : std::string()) << ";\n";
        }
    }
    for (int j = n - 1; j >= 0; --j) {
        int kd = sh.kind[j];
        if (kd != 1 && kd != 2) continue;
        int fc = sh.first_child[j], nc = sh.nchild[j];
        for (int kk = 0; kk < nc; ++kk) {
            int ch = fc + kk, ck = sh.kind[ch];
            if (ck == -1)
                o << "  const double v" << ch << " = ((pc" << j << " * qo_" << j << ")" << (kd == opp + 1 ? " * " + S_("s", ch) : std::string()) << ") * "
                  << (pl == 0 ? "" : "-") << "x" << ch << ";\n";
            else if (ck == 4) o << "  const double v" << ch << " = sm_fan[m * " << std::max(1, c.F) << " + " << sh.fan_index[ch] << "];\n";
            else if (ck == 3) o << "  const double v" << ch << " = __ldcg(pvl + " << sh.portal_rank[ch] * GB << " + m);\n";
            else o << "  const double v" << ch << " = V" << ch << ";\n";
        }
        if (kd == opp + 1) {
            o << "  double V" << j << " = 0.0;\n";
            for (int kk = 0; kk < nc; ++kk) o << "  V" << j << " = V" << j << " + v" << (fc + kk) << ";\n";
        } else {
            const int jj = own_jj[s.dec_idx[j]];
            o << "  if (act) {\n";
            for (int kk = 0; kk < nc; ++kk) o << "    sm_q[" << (jj * A + kk) * GB << " + m] = v" << (fc + kk) << ";\n";
            o << "  }\n  __syncthreads();\n  double V" << j << ";\n  {\n    int best = 0; double bq = 0.0;\n";
            for (int kk = 0; kk < nc; ++kk)
                o << "    { double q = 0.0; const double* col = sm_q + " << (jj * A + kk) * GB << ";\n"
                     "      for (int mm = 0; mm < " << GB << "; ++mm) q = q + col[mm];\n"
                     "      if (" << (kk == 0 ? "true" : "q > bq") << ") { bq = q; best = " << kk << "; } }\n";
            o << "    V" << j << " = v" << fc;
            for (int kk = 1; kk < nc; ++kk) o << ";\n    if (best == " << kk << ") V" << j << " = v" << (fc + kk);
            o << ";\n    if (tid == 0) P->br_astar[sm_ut[" << c.UT_OWN + 2 * c.MBS << " + " << jj << "]] = best;\n  }\n";
        }
    }
    o << "  if (act) P->br_top" << pl + 1 << "[P->tile_root[i]] = V0;\n"
         "  __syncthreads();\n  (void)qo_0;\n}\n";
    (void)MB;
}

void gen_br_kernel(std::ostringstream &o, const Shape &s, const Consts &c) {
    Terms u0 = upper_terms(s, 0), u1 = upper_terms(s, 1);
    const int NT = c.BLK / c.TEAM;
    const int q_team = std::max((int)s.dec[0].size(), (int)s.dec[1].size()) * c.MAXA * (c.TEAM + 1);
    const int q_up = std::max((int)s.udec[0].size() * c.GB[0], (int)s.udec[1].size() * c.GB[1]) * c.MAXA;
    const int us_n = std::max(1, std::max(u0.kr, u1.kr));
    const int fan_n = std::max(1, std::max(c.GB[0], c.GB[1]) * c.F);
    o << "extern \"C\" __global__ void __launch_bounds__(BLOCK, " << c.MINB << ") br_fused_k(const __grid_constant__ FusedParams PV) {\n"
         "  const FusedParams* __restrict__ P = &PV;\n"
         "  __shared__ double sm_q[" << std::max(NT * q_team, q_up) << "];\n"
         "  __shared__ double sm_fan[" << fan_n << "];\n"
         "  __shared__ double sm_s[" << us_n << "];\n"
         "  __shared__ int sm_ut[" << c.UT_N << "];\n"
         "  __shared__ int s_lead;\n"
         "  const unsigned int nb = gridDim.x;\n"
         "  unsigned int nbar = 0;\n"
         "  const int tid = threadIdx.x;\n"
         "  const int NU0 = " << c.NUG[0] << ", NU1 = " << c.NUG[1] << ";\n"
         "  if (tid == 0) {\n"
         "    unsigned int smid; asm volatile(\"mov.u32 %0, %%smid;\" : \"=r\"(smid));\n"
         "    const unsigned int rank = atomicAdd(P->bar + 32 + 32 * nb + (smid & 1023u), 1u);\n"
         "    s_lead = rank == 0 ? (int)atomicAdd(P->bar + 1, 1u) : -1;\n"
         "  }\n"
         "  __syncthreads();\n"
         "  const int lead = s_lead;\n"
         "  // phase 0: both players' reach to the portals under the evaluated strategy\n"
         "  for (int u = blockIdx.x; u < NU0 + NU1; u += nb) {\n"
         "    if (u < NU0) {\n"
         "      for (int x = tid; x < " << u0.kr << "; x += BLOCK) sm_s[x] = __ldcg(P->br_sig + (long long)UCA0[x] * " << c.NF << "LL + P->tg_fid0[u * " << s.udec[0].size() << " + UCJ0[x]]);\n"
         "      __syncthreads();\n"
         "      if (tid == 0) upper_reach0(P->br_reach1, u, sm_s);\n"
         "    } else {\n"
         "      const int uu = u - NU0;\n"
         "      for (int x = tid; x < " << u1.kr << "; x += BLOCK) sm_s[x] = __ldcg(P->br_sig + (long long)UCA1[x] * " << c.NF << "LL + P->tg_fid1[uu * " << s.udec[1].size() << " + UCJ1[x]]);\n"
         "      __syncthreads();\n"
         "      if (tid == 0) upper_reach1(P->br_reach2, uu, sm_s);\n"
         "    }\n"
         "    __syncthreads();\n"
         "  }\n"
         "  grid_barrier(P->bar, ++nbar, nb);\n"
         "  const int n_lead = (int)__ldcg(P->bar + 1);\n"
         "  // phase 1: the micro groups of both responders\n"
         "  {\n"
         "    const int team = tid / TEAM;\n"
         "    double* t_q = sm_q + team * " << q_team << ";\n"
         "    const long long nu0 = " << (c.NG[0] + c.GPC[0] - 1) / c.GPC[0] << "LL, nu1 = " << (c.NG[1] + c.GPC[1] - 1) / c.GPC[1] << "LL;\n"
         "    for (long long un = (long long)blockIdx.x * " << NT << " + team; un < nu0 + nu1; un += (long long)nb * " << NT << ") {\n"
         "      if (un < nu0) micro_br0(P, un, t_q); else micro_br1(P, un - nu0, t_q);\n"
         "    }\n"
         "  }\n"
         "  grid_barrier(P->bar, ++nbar, nb);\n"
         "  // phase 2: the upper groups of both responders, one leader CTA per SM\n"
         "  if (lead >= 0) {\n"
         "    for (int u = lead; u < NU0 + NU1; u += n_lead) {\n"
         "      if (u < NU0) upper_cache0(P, u, sm_ut); else upper_cache1(P, u - NU0, sm_ut);\n"
         "      __syncthreads();\n"
         "      if (u < NU0) upper_br0(P, u, sm_fan, sm_q, sm_ut); else upper_br1(P, u - NU0, sm_fan, sm_q, sm_ut);\n"
         "    }\n"
         "  }\n"
         "  grid_barrier(P->bar, ++nbar, nb);\n"
         "  // phase 3: the all-chance trunk above the tiles, children in order (one thread per responder)\n"
         "  if (blockIdx.x < 2) {\n"
         "    double* val = blockIdx.x == 0 ? P->br_top1 : P->br_top2;\n"
         "    const long long* lv = &P->tl0;\n"
         "    for (int l = (int)P->n_tl - 1; l >= 0; --l) {          // one thread per trunk node of the level\n"
         "      const long long hi = lv[l + 1] < P->n_top ? lv[l + 1] : P->n_top;\n"
         "      for (long long node = lv[l] + tid; node < hi; node += BLOCK) {\n"
         "        if (P->g_player[node] != 0) continue;\n"
         "        double acc = 0.0;\n"
         "        const int c0 = P->g_first_child[node], c1 = P->g_first_child[node + 1];\n"
         "        int ch = c0;\n"
         "        for (; ch + 8 <= c1; ch += 8) {\n"
         "          double vv[8];\n"
         "#pragma unroll\n"
         "          for (int q = 0; q < 8; ++q) vv[q] = __ldcg(val + ch + q);\n"
         "#pragma unroll\n"
         "          for (int q = 0; q < 8; ++q) acc = acc + vv[q];\n"
         "        }\n"
         "        for (; ch < c1; ++ch) acc = acc + __ldcg(val + ch);\n"
         "        val[node] = acc;\n"
         "      }\n"
         "      __syncthreads();\n"
         "    }\n"
         "    if (tid == 0) P->br_out[blockIdx.x] = __ldcg(val);\n"
         "  }\n"
         "}\n";
}


std::string generate_source(const Shape &s, const Consts &c) {
    std::ostringstream o;
    o.precision(17);
    o << "struct FusedParams {\n" << kFusedFieldsText << "};\n"
      << "#define BLOCK " << c.BLK << "\n"
      << (getenv("RLP_P2P_DEBUG") ? "#define XDEBUG 1\n" : "")
      // row stride of the per-thread shared arrays: odd, so the (set, action) threads of the ordered sums -- which read
      // different ROWS at the same column -- fall into different banks (a stride of BLOCK doubles is 11-way conflicted)
      << "#define TEAM " << c.TEAM << "\n"
      << "#define SMS " << (c.TEAM + 1) << "\n"
      << (c.TEAM == 32 ? "#define TEAM_SYNC() __syncwarp()\n" : "#define TEAM_SYNC() __syncthreads()\n")
      << R"(
__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#ifdef XDEBUG   // RLP_P2P_DEBUG: the last arriver of the cross-GPU barrier adds up where its time goes (bar[8..17], ns)
__device__ __forceinline__ void xdbg(unsigned int* bar, int i, unsigned long long& t0) {
  unsigned long long t = now_ns();
  atomicAdd((unsigned long long*)(bar + 8) + i, i ? t - t0 : 1ull);
  t0 = t;
}
#define XDBG(i) xdbg(bar, i, xdbg_t);
#else
#define XDBG(i)
#endif
// Grid barrier for co-resident CTAs (cooperative launch).  bar[0] counts arrivals (monotone over the launch); the LAST
// arriver of barrier number `k` writes k into a private flag word of every CTA (bar[32 + 32 * cta], 128 bytes apart,
// so spread over the L2 slices) and each waiting CTA polls only its own word: no address is polled by more than one
// thread.  (Polling bar[0] from ~300 CTAs saturates its L2 slice and slows every load that maps to that slice.)
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int k, unsigned int nb) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(bar, 1u);
    s_last = prev + 1u == k * nb;
    __threadfence();
  }
  __syncthreads();
  if (s_last) {
    for (unsigned int c = threadIdx.x; c < nb; c += BLOCK)
      asm volatile("st.relaxed.gpu.u32 [%0], %1;" :: "l"(bar + 32 + 32 * c), "r"(k) : "memory");
  } else if (threadIdx.x == 0) {
    const unsigned int* mine = bar + 32 + 32 * blockIdx.x;
    unsigned int v;
    for (;;) {
      asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v >= k) break;
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}
// The same barrier with the exchange step of the multi-GPU solve folded in: the subtree values were stored straight
// into every peer's buffer (NVLink peer stores) during the micro phase; the last CTA of this GPU to arrive publishes
// `token` in every peer's flag array and waits until every peer has published it here, then releases the local CTAs.
// System-scope fences order the peer stores of all local CTAs before the flag.  The wait is bounded: a lost peer
// sets bar[2] (the host reports it) instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier_x(const FusedParams* __restrict__ P, unsigned int k, unsigned int nb, unsigned long long token,
                                               bool pushed) {
  __shared__ int s_last;
  unsigned int* bar = P->bar;
  __syncthreads();
  if (threadIdx.x == 0) {
    // only the CTAs that stored into peer memory pay for a system-scope fence (hundreds of them at once are slow)
    if (pushed) __threadfence_system(); else __threadfence();
    const unsigned int prev = atomicAdd(bar, 1u);
    s_last = prev + 1u == k * nb;
    __threadfence();
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) {
      const int nr = (int)P->nranks, me = (int)P->rank;
      unsigned long long* const* peers = &P->xfp0;
      // ONE system-scope fence (it is what costs: it waits for this GPU's outstanding peer writes), then relaxed flag stores;
      // a st.release per peer would repeat that fence for every peer
      unsigned long long xdbg_t = 0; (void)xdbg_t;
      XDBG(0)
      __threadfence_system();
      XDBG(1)
      for (int r = 0; r < nr; ++r)
        if (r != me) asm volatile("st.relaxed.sys.u64 [%0], %1;" :: "l"(peers[r] + me), "l"(token) : "memory");
      XDBG(2)
      for (int r = 0; r < nr; ++r) {
        if (r == me) continue;
        unsigned long long v = 0;
        long long spins = 0;
        for (;;) {
          asm volatile("ld.relaxed.sys.u64 %0, [%1];" : "=l"(v) : "l"(P->xfl + r) : "memory");
          if (v >= token) break;
          if (++spins > (1ll << 25)) { bar[2] = 1u; break; }      // ~seconds: give up, flag the error
          __nanosleep(40);
        }
      }
      XDBG(3)
      __threadfence_system();
      XDBG(4)
    }
    __syncthreads();
    for (unsigned int c = threadIdx.x; c < nb; c += BLOCK)
      asm volatile("st.relaxed.gpu.u32 [%0], %1;" :: "l"(bar + 32 + 32 * c), "r"(k) : "memory");
  } else if (threadIdx.x == 0) {
    const unsigned int* mine = bar + 32 + 32 * blockIdx.x;
    unsigned int v;
    for (;;) {
      asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v >= k) break;
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}
// compute_regret_matching + normalise_probs (cfr_util.py:14-58) for action `a` of a set whose regrets are r[0..na)
__device__ __forceinline__ double rm_one(const double* r, int na, int a) {
  double mx = r[0];
  for (int b = 1; b < na; ++b) mx = r[b] > mx ? r[b] : mx;
  if (mx <= 0.0) return 1.0 / (double)na;
  double s = 0.0, c = 0.0, mine = 0.0;
  for (int b = 0; b < na; ++b) {
    double p = r[b] > 0.0 ? r[b] : 0.0;
    p = p > 1e-7 ? p : 1e-7;
    double t = s + p;
    if (fabs(s) >= fabs(p)) c += (s - t) + p; else c += (p - t) + s;
    s = t;
    if (b == a) mine = p;
  }
  return mine / (s + c);
}
)";
    if (!c.fan_cps.empty()) {
        o << "__constant__ double FCP[" << c.fan_cps.size() << "] = {";
        for (size_t f = 0; f < c.fan_cps.size(); ++f) o << (f ? "," : "") << c.fan_cps[f];
        o << "};\n";
    }
    o << "struct MIds { int pat, ri, ro, pi, acol, afid; double pcv; int f[" << (int)s.tm->n_nt << "]; };\n";
    for (int pl = 0; pl < 2; ++pl) {
        Terms mt = micro_terms(s, pl), ut = upper_terms(s, pl);
        emit_table(o, "MCJ", pl, mt.cj); emit_table(o, "MCA", pl, mt.ca); emit_table(o, "MCOFF", pl, mt.coff); emit_table(o, "MNA", pl, mt.na);
        emit_table(o, "UCJ", pl, ut.cj); emit_table(o, "UCA", pl, ut.ca); emit_table(o, "UCOFF", pl, ut.coff); emit_table(o, "UNA", pl, ut.na);
        gen_micro_eval(o, s, c, pl);
        gen_upper_eval(o, s, c, pl);
        gen_upper_reach(o, s, c, pl);
    }
    // ---- per-player batch / unit drivers (same text, different literals) ------------------------------------------
    for (int pl = 0; pl < 2; ++pl) {
        Terms mt = micro_terms(s, pl), ut = upper_terms(s, pl);
        const int G = c.G[pl], GPC = c.GPC[pl], KR = mt.kr, MO = (int)s.dec[pl].size();
        const int GPT = GPC;                            // groups per team
        const std::string gfirst = "P->first_group" + std::to_string(pl), gend = "P->end_group" + std::to_string(pl);
        const bool one_pass = GPT * KR <= c.TEAM;      // every (set, action) of the unit has its own thread
        // the static indices of a unit (fetched one unit ahead by the phase loop)
        o << "__device__ __forceinline__ MIds micro_unit_ids" << pl << "(const FusedParams* __restrict__ P, long long unit) {\n"
             "  const int tix = threadIdx.x % TEAM;\n"
             "  const int gl = tix / " << G << ", k = tix - gl * " << G << ";\n"
             "  const long long g0 = " << gfirst << " + unit * " << GPT << ", gend = " << gend << ";\n"
             "  const long long g = g0 + gl;\n";
        if (one_pass)
            o << "  const int agl = tix / " << KR << ", ci = tix - agl * " << KR << ";\n"
                 "  const bool acc = tix < " << GPT * KR << " && g0 + agl < gend;\n"
                 "  return micro_ids" << pl << "(P, g, k, gl < " << GPT << " && g < gend, g0 + agl, acc ? MCJ" << pl << "[ci] : 0, acc);\n}\n";
        else
            o << "  return micro_ids" << pl << "(P, g, k, gl < " << GPT << " && g < gend, 0, 0, false);\n}\n";
        o << "__device__ __forceinline__ void micro_unit" << pl << "(const FusedParams* __restrict__ P, long long unit, const MIds& id, double w, const double* sig_in,\n"
             "    double* sig_out, double* sm_r, double* sm_o, double* sm_rn, int par) {\n"
             "  const int tix = threadIdx.x % TEAM;\n"
             "  const int gl = tix / " << G << ";\n"
             "  const long long g0 = " << gfirst << " + unit * " << GPT << ", gend = " << gend << ";\n"
             "  const long long g = g0 + gl;\n";
        if (one_pass) {
            // the accumulating threads fetch regret, action count and strategy BEFORE the evaluation, so those loads
            // overlap it instead of sitting behind the team barrier
            o << "  const int agl = tix / " << KR << ", ci = tix - agl * " << KR << ";\n"
                 "  const long long ag = g0 + agl;\n"
                 "  const bool acc = tix < " << GPT * KR << " && ag < gend;\n"
                 "  int j = 0, a = 0, col = 0, fid = 0;\n"
                 "  double r = 0.0, sa = 0.0, sg = 0.0;\n"
                 "  if (acc) {\n"
                 "    j = MCJ" << pl << "[ci]; a = MCA" << pl << "[ci];\n"
                 "    col = id.acol + a;\n"
                 "    fid = id.afid;\n"
                 "    r = __ldcg(P->R + col); sa = __ldcg(P->S + col);\n"
                 "    sg = __ldcg(sig_in + (long long)a * " << c.NF << "LL + fid);\n"
                 "  }\n"
                 "  if (gl < " << GPT << " && g < gend) micro_eval" << pl << "(P, id, tix, w, sig_in, sm_r, sm_o, par);\n"
                 "  TEAM_SYNC();\n"
                 "  if (acc) {\n"
                 "    const double* cr = sm_r + ci * SMS + agl * " << G << ";\n"
                 "    const double* co = sm_o + j * SMS + agl * " << G << ";\n"
                 "#pragma unroll 8\n"
                 "    for (int m = 0; m < " << G << "; ++m) { r = r + cr[m]; sa = sa + co[m] * sg; }\n"
                 "    P->R[col] = r; P->S[col] = sa;\n"
                 "    sm_rn[tix] = r;\n"
                 "  }\n"
                 "  TEAM_SYNC();\n"
                 "  if (acc) {\n"
                 "    const double out = rm_one(sm_rn + agl * " << KR << " + MCOFF" << pl << "[j], MNA" << pl << "[j], a);\n"
                 "    P->sigma_row[col] = out;\n"
                 "    sig_out[(long long)a * " << c.NF << "LL + fid] = out;\n"
                 "    if (a == 0) { const int is = P->grp_iset" << pl << "[ag * " << MO << " + j]; const long long I = P->num_isets;\n"
                 "      P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }\n"
                 "  }\n"
                 "  TEAM_SYNC();\n"
                 "}\n";
        } else {
            o << "  if (gl < " << GPT << " && g < gend) micro_eval" << pl << "(P, id, tix, w, sig_in, sm_r, sm_o, par);\n"
                 "  TEAM_SYNC();\n"
                 "  for (int x = tix; x < " << GPT * KR << "; x += TEAM) {\n"
                 "    const int agl = x / " << KR << ", ci = x - agl * " << KR << ";\n"
                 "    const long long ag = g0 + agl;\n"
                 "    if (ag >= gend) continue;\n"
                 "    const int j = MCJ" << pl << "[ci], a = MCA" << pl << "[ci];\n"
                 "    const int col = P->grp_col" << pl << "[ag * " << MO << " + j] + a;\n"
                 "    const int fid = P->grp_fid" << pl << "[ag * " << MO << " + j];\n"
                 "    double r = __ldcg(P->R + col), sa = __ldcg(P->S + col);\n"
                 "    const double sg = __ldcg(sig_in + (long long)a * " << c.NF << "LL + fid);\n"
                 "    const double* cr = sm_r + ci * SMS + agl * " << G << ";\n"
                 "    const double* co = sm_o + j * SMS + agl * " << G << ";\n"
                 "#pragma unroll 8\n"
                 "    for (int m = 0; m < " << G << "; ++m) { r = r + cr[m]; sa = sa + co[m] * sg; }\n"
                 "    P->R[col] = r; P->S[col] = sa;\n"
                 "    sm_rn[x] = r;\n"
                 "  }\n"
                 "  TEAM_SYNC();\n"
                 "  for (int x = tix; x < " << GPT * KR << "; x += TEAM) {\n"
                 "    const int agl = x / " << KR << ", ci = x - agl * " << KR << ";\n"
                 "    const long long ag = g0 + agl;\n"
                 "    if (ag >= gend) continue;\n"
                 "    const int j = MCJ" << pl << "[ci], a = MCA" << pl << "[ci];\n"
                 "    const int col = P->grp_col" << pl << "[ag * " << MO << " + j] + a;\n"
                 "    const int fid = P->grp_fid" << pl << "[ag * " << MO << " + j];\n"
                 "    const double out = rm_one(sm_rn + agl * " << KR << " + MCOFF" << pl << "[j], MNA" << pl << "[j], a);\n"
                 "    P->sigma_row[col] = out;\n"
                 "    sig_out[(long long)a * " << c.NF << "LL + fid] = out;\n"
                 "    if (a == 0) { const int is = P->grp_iset" << pl << "[ag * " << MO << " + j]; const long long I = P->num_isets;\n"
                 "      P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }\n"
                 "  }\n"
                 "  TEAM_SYNC();\n"
                 "}\n";
        }
        const int GB = c.GB[pl], KRB = ut.kr, MB = (int)s.udec[pl].size();
        // mode 0: full unit (fan-in, evaluate, accumulate, regret matching, reach); mode 1: reach only (prologue)
        const bool acc_one = KRB <= c.BLK;
        // per-unit static tables cached in shared memory (sm_ut): [0, GBS) tiles | [GBS, GBS + n_udec * GBS) fused ids of the
        // tiles' decision nodes | then own-decision first pairs, fused ids, information sets (MBS each)
        o << "__device__ __noinline__ void upper_cache" << pl << "(const FusedParams* __restrict__ P, int u, int* sm_ut) {\n"
             "  const int tid = threadIdx.x;\n"
             "  for (int x = tid; x < " << GB << "; x += BLOCK) sm_ut[x] = P->tg_tile" << pl << "[(long long)u * " << GB << " + x];\n"
             "  for (int x = tid; x < " << s.n_udec * GB << "; x += BLOCK) {\n"
             "    const int d = x / " << GB << ", m = x - d * " << GB << ";\n"
             "    sm_ut[" << c.GBS << " + d * " << c.GBS << " + m] = P->u_fid[(long long)d * " << c.T << "LL + P->tg_tile" << pl << "[(long long)u * " << GB << " + m]];\n"
             "  }\n"
             "  for (int x = tid; x < " << MB << "; x += BLOCK) {\n"
             "    sm_ut[" << c.UT_OWN << " + x] = P->tg_col" << pl << "[u * " << MB << " + x];\n"
             "    sm_ut[" << c.UT_OWN + c.MBS << " + x] = P->tg_fid" << pl << "[u * " << MB << " + x];\n"
             "    sm_ut[" << c.UT_OWN + 2 * c.MBS << " + x] = P->tg_iset" << pl << "[u * " << MB << " + x];\n"
             "  }\n"
             "}\n";
        o << "__device__ __noinline__ void upper_unit" << pl << "(const FusedParams* __restrict__ P, int u, double w, const double* sig_in,\n"
             "    double* sig_out, int reach_only, double* sm_fan, double* sm_r, double* sm_o, double* sm_rn, double* sm_s, const int* sm_ut,\n"
             "    int par, unsigned long long* st) {\n"
             "  const int tid = threadIdx.x;\n"
             "  const int* tiles = sm_ut;\n"
             "  if (reach_only) {\n"
             "    for (int x = tid; x < " << KRB << "; x += BLOCK) {\n"
             "      const int j = UCJ" << pl << "[x], a = UCA" << pl << "[x];\n"
             "      sm_s[x] = __ldcg(sig_in + (long long)a * " << c.NF << "LL + sm_ut[" << c.UT_OWN + c.MBS << " + j]);\n"
             "    }\n"
             "  } else {\n";
        if (acc_one)        // regret / action count / strategy of the unit's (set, action) pairs: in flight during the fan-in
            o << "    int aj = 0, aa = 0, acol = 0, afid = 0;\n"
                 "    double ar = 0.0, asa = 0.0, asg = 0.0;\n"
                 "    if (tid < " << KRB << ") {\n"
                 "      aj = UCJ" << pl << "[tid]; aa = UCA" << pl << "[tid];\n"
                 "      acol = sm_ut[" << c.UT_OWN << " + aj] + aa;\n"
                 "      afid = sm_ut[" << c.UT_OWN + c.MBS << " + aj];\n"
                 "      ar = __ldcg(P->R + acol); asa = __ldcg(P->S + acol);\n"
                 "      asg = __ldcg(sig_in + (long long)aa * " << c.NF << "LL + afid);\n"
                 "    }\n";
        // this unit's subtree values: one contiguous block [portal][member] of this player's array
        o << "    const int* fd = P->fan_desc;\n"
             "    const double* blk = (par ? P->pvl1 : P->pvl0) + " << (long long)pl * c.PVLEN << "LL + (long long)u * " << (long long)c.KP * GB << "LL;\n"
             "    UIn" << pl << " uin;\n"
             "    if (tid < " << GB << ") upper_load" << pl << "(P, tiles[tid], tid, sm_ut, sig_in, blk, uin);\n"
             "    if (st && tid == 0) st[0] = now_ns();\n";
        {
            // Fan-in.  The block is copied through shared memory in its own layout (rows = portals, GB values each: the
            // threads of the sums below -- one per (tile, fan-in node) -- then read consecutive words), a few fan-in
            // nodes at a time when it does not fit at once.  When the whole block fits the registers of the CTA, every
            // load is issued before the first store (one memory round trip).
            const int stage = c.sm_r_n + c.sm_o_n;
            std::vector<int> k0s, nchs;
            for (const FanDesc &fd : c.fans) { k0s.push_back(fd.k0); nchs.push_back(fd.nchild); }
            std::vector<std::pair<int, int>> chunks;        // [first fan, end fan)
            for (int f0 = 0; f0 < c.F;) {
                int f1 = f0, rows = 0;
                while (f1 < c.F && (rows + nchs[f1]) * GB <= stage) { rows += nchs[f1]; ++f1; }
                if (f1 == f0) { chunks.clear(); break; }     // a single fan-in node does not fit
                chunks.push_back({f0, f1});
                f0 = f1;
            }
            // consecutive fan-in nodes must cover consecutive portal rows for the linear copy
            bool linear = !chunks.empty();
            for (int f = 0; f + 1 < c.F; ++f) if (k0s[f] + nchs[f] != k0s[f + 1]) linear = false;
            const long long total = (long long)c.KP * GB;
            const int NV = (int)((total + c.BLK - 1) / c.BLK);
            const bool in_regs = linear && NV <= 32 && c.F > 0 && k0s[0] == 0 && k0s.back() + nchs.back() == c.KP;
            const std::string ecp = c.fan_cps.empty() ? "" : "FCP[f]";
            auto emit_sums = [&](int f0, int f1, int row0) {
                o << "      for (int x = tid; x < " << (f1 - f0) * GB << "; x += BLOCK) {\n"
                     "        const int fl = x / " << GB << ", m = x - fl * " << GB << ", f = " << f0 << " + fl;\n"
                     "        const int k0 = fd[4 * f], nch = fd[4 * f + 1];\n";
                if (c.fan_cps.empty())
                    o << "        const double* cp = P->fan_cp + (long long)fd[4 * f + 2] * " << c.T << "LL + tiles[m];\n";
                else
                    o << "        const double e = FCP[f];\n";
                o << "        const double* pv = sm_r + (k0 - " << row0 << ") * " << GB << " + m;\n"
                     "        double acc = 0.0;\n"
                     "        for (int ch = 0; ch < nch; ++ch) acc = acc + " << (c.fan_cps.empty() ? "cp[(long long)ch * " + std::to_string(c.T) + "LL]" : std::string("e")) << " * pv[ch * " << GB << "];\n"
                     "        sm_fan[m * " << std::max(1, c.F) << " + f] = acc;\n"
                     "      }\n";
            };
            if (in_regs) {
                o << "    double fv[" << NV << "];\n"
                     "#pragma unroll\n"
                     "    for (int i = 0; i < " << NV << "; ++i) { const int L = tid + i * BLOCK; if (L < " << total << ") fv[i] = __ldcg(blk + L); }\n";
                for (auto &ch : chunks) {
                    const int lo = k0s[ch.first] * GB, hi = (k0s[ch.second - 1] + nchs[ch.second - 1]) * GB;
                    o << "    {\n#pragma unroll\n"
                         "      for (int i = 0; i < " << NV << "; ++i) { const int L = tid + i * BLOCK; if (L >= " << lo << " && L < " << hi << ") sm_r[L - " << lo << "] = fv[i]; }\n"
                         "      __syncthreads();\n";
                    emit_sums(ch.first, ch.second, k0s[ch.first]);
                    o << "      __syncthreads();\n    }\n";
                }
            } else if (linear) {
                // several passes through the staging area: the NEXT pass's values are fetched into registers while the sums
                // of the current one run, so a pass costs max(load latency, sums) instead of their sum
                int nvc = 0;
                for (auto &ch : chunks) {
                    const int lo = k0s[ch.first] * GB, hi = (k0s[ch.second - 1] + nchs[ch.second - 1]) * GB;
                    nvc = std::max(nvc, (hi - lo + c.BLK - 1) / c.BLK);
                }
                if (nvc <= 32) {
                    o << "    double fv[" << nvc << "];\n";
                    auto emit_load = [&](size_t ci) {
                        const int lo = k0s[chunks[ci].first] * GB, hi = (k0s[chunks[ci].second - 1] + nchs[chunks[ci].second - 1]) * GB;
                        o << "#pragma unroll\n"
                             "      for (int i = 0; i < " << nvc << "; ++i) { const int L = " << lo << " + tid + i * BLOCK; if (L < " << hi << ") fv[i] = __ldcg(blk + L); }\n";
                    };
                    o << "    {\n";
                    emit_load(0);
                    o << "    }\n";
                    for (size_t ci = 0; ci < chunks.size(); ++ci) {
                        const int lo = k0s[chunks[ci].first] * GB, hi = (k0s[chunks[ci].second - 1] + nchs[chunks[ci].second - 1]) * GB;
                        o << "    {\n#pragma unroll\n"
                             "      for (int i = 0; i < " << nvc << "; ++i) { const int L = tid + i * BLOCK; if (L < " << hi - lo << ") sm_r[L] = fv[i]; }\n"
                             "      __syncthreads();\n";
                        if (ci + 1 < chunks.size()) emit_load(ci + 1);
                        emit_sums(chunks[ci].first, chunks[ci].second, k0s[chunks[ci].first]);
                        o << "      __syncthreads();\n    }\n";
                    }
                } else {
                    for (auto &ch : chunks) {
                        const int lo = k0s[ch.first] * GB, hi = (k0s[ch.second - 1] + nchs[ch.second - 1]) * GB;
                        o << "    {\n"
                             "      for (int L = " << lo << " + tid; L < " << hi << "; L += BLOCK) sm_r[L - " << lo << "] = __ldcg(blk + L);\n"
                             "      __syncthreads();\n";
                        emit_sums(ch.first, ch.second, k0s[ch.first]);
                        o << "      __syncthreads();\n    }\n";
                    }
                }
            } else {        // no staging: every (tile, fan-in node) thread reads its values from the block directly
                o << "    for (int x = tid; x < " << GB * c.F << "; x += BLOCK) {\n"
                     "      const int f = x / " << GB << ", m = x - f * " << GB << ";\n"
                     "      const int k0 = fd[4 * f], nch = fd[4 * f + 1];\n"
                     "      const double* cp = P->fan_cp + (long long)fd[4 * f + 2] * " << c.T << "LL + tiles[m];\n"
                     "      const double* pv = blk + (long long)k0 * " << GB << " + m;\n"
                     "      double acc = 0.0;\n"
                     "      for (int ch = 0; ch < nch; ++ch) acc = acc + cp[(long long)ch * " << c.T << "LL] * __ldcg(pv + (long long)ch * " << GB << ");\n"
                     "      sm_fan[m * " << std::max(1, c.F) << " + f] = acc;\n"
                     "    }\n"
                     "    __syncthreads();\n";
            }
            (void)ecp;
        }
        o <<
             "    if (st && tid == 0) st[1] = now_ns();\n"
             "    if (tid < " << GB << ") upper_eval" << pl << "(tid, w, uin, sm_fan, sm_r, sm_o);\n"
             "    __syncthreads();\n"
             "    if (st && tid == 0) st[2] = now_ns();\n";
        if (acc_one) {
            o << "    if (tid < " << KRB << ") {\n"
                 "      const double* cr = sm_r + tid * " << GB << ";\n"
                 "      const double* co = sm_o + aj * " << GB << ";\n"
                 "      for (int m = 0; m < " << GB << "; ++m) { ar = ar + cr[m]; asa = asa + co[m] * asg; }\n"
                 "      P->R[acol] = ar; P->S[acol] = asa;\n"
                 "      sm_rn[tid] = ar;\n"
                 "    }\n"
                 "    __syncthreads();\n"
                 "    if (tid < " << KRB << ") {\n"
                 "      const double out = rm_one(sm_rn + UCOFF" << pl << "[aj], UNA" << pl << "[aj], aa);\n"
                 "      P->sigma_row[acol] = out;\n"
                 "      sig_out[(long long)aa * " << c.NF << "LL + afid] = out;\n"
                 "      sm_s[tid] = out;\n"
                 "      if (aa == 0) { const int is = sm_ut[" << c.UT_OWN + 2 * c.MBS << " + aj]; const long long I = P->num_isets;\n"
                 "        P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }\n"
                 "    }\n";
        } else {
            o << "    for (int x = tid; x < " << KRB << "; x += BLOCK) {\n"
                 "      const int j = UCJ" << pl << "[x], a = UCA" << pl << "[x];\n"
                 "      const int col = P->tg_col" << pl << "[u * " << MB << " + j] + a;\n"
                 "      const int fid = P->tg_fid" << pl << "[u * " << MB << " + j];\n"
                 "      double r = __ldcg(P->R + col), sa = __ldcg(P->S + col);\n"
                 "      const double sg = __ldcg(sig_in + (long long)a * " << c.NF << "LL + fid);\n"
                 "      const double* cr = sm_r + x * " << GB << ";\n"
                 "      const double* co = sm_o + j * " << GB << ";\n"
                 "      for (int m = 0; m < " << GB << "; ++m) { r = r + cr[m]; sa = sa + co[m] * sg; }\n"
                 "      P->R[col] = r; P->S[col] = sa;\n"
                 "      sm_rn[x] = r;\n"
                 "    }\n"
                 "    __syncthreads();\n"
                 "    for (int x = tid; x < " << KRB << "; x += BLOCK) {\n"
                 "      const int j = UCJ" << pl << "[x], a = UCA" << pl << "[x];\n"
                 "      const int col = P->tg_col" << pl << "[u * " << MB << " + j] + a;\n"
                 "      const int fid = P->tg_fid" << pl << "[u * " << MB << " + j];\n"
                 "      const double out = rm_one(sm_rn + UCOFF" << pl << "[j], UNA" << pl << "[j], a);\n"
                 "      P->sigma_row[col] = out;\n"
                 "      sig_out[(long long)a * " << c.NF << "LL + fid] = out;\n"
                 "      sm_s[x] = out;\n"
                 "      if (a == 0) { const int is = P->tg_iset" << pl << "[u * " << MB << " + j]; const long long I = P->num_isets;\n"
                 "        P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }\n"
                 "    }\n";
        }
        o << "  }\n"
             "  __syncthreads();\n"
             "  if (st && tid == 0) st[3] = now_ns();\n"
             "  if (tid == 0) upper_reach" << pl << "(P->portal_p" << pl + 1 << ", u, sm_s);\n"
             "  __syncthreads();\n"
             "  if (st && tid == 0) st[4] = now_ns();\n"
             "}\n";
    }
    Terms m0 = micro_terms(s, 0), m1 = micro_terms(s, 1), u0 = upper_terms(s, 0), u1 = upper_terms(s, 1);
    const int KRmax = std::max(m0.kr, m1.kr), MOmax = (int)std::max(s.dec[0].size(), s.dec[1].size());
    const int rn_a = std::max(c.GPC[0] * m0.kr, c.GPC[1] * m1.kr);
    const int fan_n = std::max(1, std::max(c.GB[0], c.GB[1]) * c.F);
    const int ur_n = std::max(1, std::max(u0.kr * c.GB[0], u1.kr * c.GB[1]));
    const int uo_n = std::max(1, std::max((int)s.udec[0].size() * c.GB[0], (int)s.udec[1].size() * c.GB[1]));
    const int us_n = std::max(1, std::max(u0.kr, u1.kr));
    const int NT = c.BLK / c.TEAM;                      // teams per CTA
    const int rn_team = std::max(1, std::max(c.GPC[0] * m0.kr, c.GPC[1] * m1.kr));
    const int sm_r_n = c.sm_r_n, sm_o_n = c.sm_o_n;
    (void)ur_n; (void)uo_n; (void)MOmax;
    const int sm_rn_n = std::max(NT * rn_team, us_n);
    (void)rn_a;
    o << "extern \"C\" __global__ void __launch_bounds__(BLOCK, " << c.MINB << ") cfr_fused_k(const __grid_constant__ FusedParams PV, long long t0, int iters,\n"
         "    int linear, int cur0) {\n"
         "  const FusedParams* __restrict__ P = &PV;      // parameter (constant) space: no global load behind a field read\n"
         "  __shared__ double sm_ro[" << sm_r_n + sm_o_n << "];\n"
         "  __shared__ int sm_ut[" << c.UT_N << "];\n"
         "  double* sm_r = sm_ro;\n"
         "  double* sm_o = sm_ro + " << sm_r_n << ";\n"
         "  int cached_u = -1;\n"
         "  __shared__ double sm_rn[" << sm_rn_n << "];\n"
         "  __shared__ double sm_fan[" << fan_n << "];\n"
         "  __shared__ double sm_s[" << us_n << "];\n"
         "  __shared__ int s_lead;\n"
         "  const unsigned int nb = gridDim.x;\n"
         "  unsigned int nbar = 0;\n"
         "  long long tlc = 0;\n"
         "  const int NU0 = " << c.NUG[0] << ", NU1 = " << c.NUG[1] << ";\n"
         "  // The upper phase has far fewer units than there are CTAs.  Consecutive CTAs share an SM, so the units go to\n"
         "  // one LEADER CTA per SM (the first CTA to register on its SM), spread over all SMs.\n"
         "  if (threadIdx.x == 0) {\n"
         "    unsigned int smid; asm volatile(\"mov.u32 %0, %%smid;\" : \"=r\"(smid));\n"
         "    const unsigned int rank = atomicAdd(P->bar + 32 + 32 * nb + (smid & 1023u), 1u);\n"
         "    s_lead = rank == 0 ? (int)atomicAdd(P->bar + 1, 1u) : -1;\n"
         "  }\n"
         "  __syncthreads();\n"
         "  const int lead = s_lead;\n"
         "  const bool logc = P->tl != 0 && lead == 0;            // the CTA whose view of the phases is recorded\n"
         "  const bool log = logc && threadIdx.x == 0;\n"
         "  // prologue: reach of both players to the portals from the current strategy\n"
         "  {\n"
         "    const double* sig_in = cur0 ? P->sigF1 : P->sigF0;\n"
         "    for (int u = blockIdx.x; u < NU0 + NU1; u += nb) {\n"
         "      if (u < NU0) upper_cache0(P, u, sm_ut); else upper_cache1(P, u - NU0, sm_ut);\n"
         "      __syncthreads();\n"
         "      if (u < NU0) upper_unit0(P, u, 1.0, sig_in, 0, 1, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, 0, 0);\n"
         "      else upper_unit1(P, u - NU0, 1.0, sig_in, 0, 1, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, 0, 0);\n"
         "    }\n"
         "    grid_barrier(P->bar, ++nbar, nb);\n"
         "  }\n"
         "  const int n_lead = (int)__ldcg(P->bar + 1);\n"
         "  const int team = threadIdx.x / TEAM;\n"
         "  double* t_r = sm_r + team * " << KRmax * (c.TEAM + 1) << ";\n"
         "  double* t_o = sm_o + team * " << MOmax * (c.TEAM + 1) << ";\n"
         "  double* t_rn = sm_rn + team * " << rn_team << ";\n"
         "  const long long nu0 = (P->end_group0 - P->first_group0 + " << c.GPC[0] - 1 << ") / " << c.GPC[0] << ";\n"
         "  const long long nu1 = (P->end_group1 - P->first_group1 + " << c.GPC[1] - 1 << ") / " << c.GPC[1] << ";\n"
         "  for (int it = 0; it < iters; ++it) {\n"
         "    const int cur = (cur0 + it) & 1;\n"
         "    const double* sig_in = cur ? P->sigF1 : P->sigF0;\n"
         "    double* sig_out = cur ? P->sigF0 : P->sigF1;\n"
         "    const double w = linear ? (double)(t0 + it) : 1.0;\n"
         "    const int par = (int)((P->xbase + it) & 1);        // portal-value buffer of this iteration\n"

         "    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();\n"
         "    // phase A: micro units of player 1, then of player 2 (one list), one unit per team at a time\n"
         "    {\n"
         "      const long long stride = (long long)nb * " << NT << ";\n"
         "      long long un = (long long)blockIdx.x * " << NT << " + team;\n";
    if (c.PREFETCH)
        o << "      MIds cur;\n"
             "      if (un < nu0 + nu1) cur = un < nu0 ? micro_unit_ids0(P, un) : micro_unit_ids1(P, un - nu0);\n"
             "      while (un < nu0 + nu1) {\n"
             "        const long long nx = un + stride;\n"
             "        MIds nxt;                    // the next unit's static indices: in flight while this unit runs\n"
             "        if (nx < nu0 + nu1) nxt = nx < nu0 ? micro_unit_ids0(P, nx) : micro_unit_ids1(P, nx - nu0);\n"
             "        if (un < nu0) micro_unit0(P, un, cur, w, sig_in, sig_out, t_r, t_o, t_rn, par);\n"
             "        else micro_unit1(P, un - nu0, cur, w, sig_in, sig_out, t_r, t_o, t_rn, par);\n"
             "        cur = nxt;\n"
             "        un = nx;\n"
             "      }\n";
    else
        o << "      for (; un < nu0 + nu1; un += stride) {\n"
             "        if (un < nu0) micro_unit0(P, un, micro_unit_ids0(P, un), w, sig_in, sig_out, t_r, t_o, t_rn, par);\n"
             "        else micro_unit1(P, un - nu0, micro_unit_ids1(P, un - nu0), w, sig_in, sig_out, t_r, t_o, t_rn, par);\n"
             "      }\n";
    o << "    }\n"
         "    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();\n"
         "    // Multi-GPU: the micro phase has stored the values of this rank's subtrees into every rank's buffer; the barrier\n"
         "    // that ends the phase also shakes hands with the peers (each CTA fences its peer stores at system scope first).\n"
         "    if (P->nranks > 1) grid_barrier_x(P, ++nbar, nb, (unsigned long long)(P->xbase + it + 1), " << (c.XFENCE_ALL ? "true" : "false") << ");\n"
         "    else grid_barrier(P->bar, ++nbar, nb);\n"
         "    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();\n"
         "    // phase B: upper groups (fan-in, tile, round-1 information sets, reach for the next iteration)\n"
         "    if (lead >= 0) {\n"
         "      for (int u = lead; u < NU0 + NU1; u += n_lead) {\n"
         "        unsigned long long* st = (logc && u == lead && tlc + 5 <= P->tl_cap) ? P->tl + tlc : 0;\n"
         "        if (u != cached_u) {          // a leader keeps its unit's static tables across iterations\n"
         "          if (u < NU0) upper_cache0(P, u, sm_ut); else upper_cache1(P, u - NU0, sm_ut);\n"
         "          cached_u = u;\n"
         "          __syncthreads();\n"
         "        }\n"
         "        if (u < NU0) upper_unit0(P, u, w, sig_in, sig_out, 0, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, par, st);\n"
         "        else upper_unit1(P, u - NU0, w, sig_in, sig_out, 0, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, par, st);\n"
         "      }\n"
         "    }\n"
         "    tlc += 5;\n"
         "    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();\n"
         "    grid_barrier(P->bar, ++nbar, nb);\n"
         "  }\n"
         "}\n";
    if (c.BR_OK) {
        for (int pl = 0; pl < 2; ++pl) { gen_micro_br(o, s, c, pl); gen_upper_br(o, s, c, pl); }
        gen_br_kernel(o, s, c);
    }
    return o.str();
}

__global__ void __launch_bounds__(256)
k_sigma_fused(int32_t I, long long NF, const int64_t *__restrict__ col_off, const int *__restrict__ fid,
              const double *__restrict__ sigma, double *sigF) {
    int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= I) return;
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    for (int a = 0; a < na; ++a) sigF[(long long)a * NF + fid[s]] = sigma[c0 + a];
}

}  // namespace

namespace rlp {

// Called at the end of build_tiles when all micro subtrees share one shape and all upper tiles share one shape.
// Leaves ts.fused null (the five-kernel pipeline stays in use) when the tree does not have the group structure.
int plan_fused(const FusedCtx &cx, cudaStream_t stream) {
    rlp_tree *t = cx.t;
    TileSet &ts = t->tiles;
    ts.fused.reset();
    const char *off = getenv("RLP_NO_FUSED");
    if (off && off[0] == '1') return RLP_OK;
    if (!t->zero_sum || t->max_act > 4 || ts.num_classes != 1 || !ts.upper_jit) return RLP_OK;
    const rlp_tree_desc *d = cx.d;
    const int32_t *fc = d->first_child;
    const MicroTmpl &tm = *cx.tm;
    const UpperShape &sh = *cx.sh;
    const int64_t P = (int64_t)cx.portal_node->size();
    const int64_t T = (int64_t)cx.desc->size();
    const int32_t I = t->I;
    if (P == 0 || T == 0) return RLP_OK;
    {   // every portal must hang off a tile (no portals directly under the trunk)
        const TileDesc &last = (*cx.desc)[T - 1];
        if (last.portal_off + last.n_portal != P) return RLP_OK;
    }
    // CTA shape: RLP_FUSED_BLOCK threads (default 256), RLP_FUSED_MINB resident CTAs per SM (default 2)
    int blk = 256, minb = 3;
    { const char *e = getenv("RLP_FUSED_BLOCK"); if (e && atoi(e) >= 32 && atoi(e) <= kMaxBlock && atoi(e) % 32 == 0) blk = atoi(e); }
    { const char *e = getenv("RLP_FUSED_MINB"); if (e && atoi(e) >= 1 && atoi(e) <= 8) minb = atoi(e); }
    const std::vector<int32_t> &srank = *cx.srank, &krank = *cx.krank;
    std::vector<int32_t> size(I, 0);
    for (int64_t g = 0; g < d->num_nodes; ++g) if (d->player[g] > 0) size[srank[g]]++;

    Shape s;
    s.tm = &tm; s.sh = &sh;
    for (int r = 0; r < tm.n_nt; ++r) {
        int p = tm.player[tm.nt_node[r]];
        if (p != 1 && p != 2) return RLP_OK;
        s.dec[p - 1].push_back(r);
    }
    if (s.dec[0].empty() || s.dec[1].empty()) return RLP_OK;
    const int n_up = (int)sh.kind.size();
    s.dec_idx.assign(n_up, -1); s.term_idx.assign(n_up, -1); s.edge_idx.assign(n_up, -1); s.parent.assign(n_up, -1); s.slot.assign(n_up, 0);
    for (int j = 0; j < n_up; ++j) {
        if (sh.kind[j] == 1 || sh.kind[j] == 2) { s.udec[sh.kind[j] - 1].push_back(s.n_udec); s.dec_idx[j] = s.n_udec++; }
        else if (sh.kind[j] == -1) s.term_idx[j] = s.n_uterm++;
        for (int k = 0; k < sh.nchild[j]; ++k) {
            int ch = sh.first_child[j] + k;
            s.parent[ch] = j; s.slot[ch] = k;
            if (sh.kind[j] == 0) s.edge_idx[ch] = s.n_uedge++;
        }
    }
    if (s.udec[0].empty() || s.udec[1].empty()) return RLP_OK;

    // ---- decision nodes of every micro instance (global node ids by non-terminal rank) --------------------------
    const int n_nt = tm.n_nt;
    std::vector<int32_t> inode((size_t)P * n_nt);
    {
        std::vector<int32_t> loc;
        for (int64_t p = 0; p < P; ++p) {
            loc.clear();
            loc.push_back((*cx.portal_node)[p]);
            for (size_t q = 0; q < loc.size(); ++q)
                for (int32_t c = fc[loc[q]]; c < fc[loc[q] + 1]; ++c) loc.push_back(c);
            for (int r = 0; r < n_nt; ++r) inode[(size_t)p * n_nt + r] = loc[tm.nt_node[r]];
        }
    }
    std::vector<uint8_t> covered(I, 0);
    // ---- micro groups per player ---------------------------------------------------------------------------------
    struct Groups { int G = 0; std::vector<int32_t> key; std::vector<int64_t> members; };   // members[g * G + k]
    Groups mg[2];
    std::vector<int32_t> group_of[2];
    for (int pl = 0; pl < 2; ++pl) {
        Groups &gr = mg[pl];
        const int r0 = s.dec[pl][0];
        gr.G = size[srank[inode[r0]]];
        if (gr.G < 1 || gr.G > blk) return RLP_OK;
        std::vector<int32_t> gid_of_iset(I, -1);
        group_of[pl].assign(P, -1);
        for (int64_t p = 0; p < P; ++p) {
            int32_t key = srank[inode[(size_t)p * n_nt + r0]];
            int32_t k = krank[inode[(size_t)p * n_nt + r0]];
            int32_t g = gid_of_iset[key];
            if (g < 0) {
                g = gid_of_iset[key] = (int32_t)gr.key.size();
                gr.key.push_back(key);
                gr.members.resize(gr.members.size() + gr.G, -1);
            }
            if (size[key] != gr.G || k >= gr.G || gr.members[(size_t)g * gr.G + k] >= 0) return RLP_OK;
            gr.members[(size_t)g * gr.G + k] = p;
            group_of[pl][p] = g;
        }
        for (int64_t m : gr.members) if (m < 0) return RLP_OK;           // incomplete group (e.g. a deal shard)
        const size_t NG = gr.key.size();
        for (size_t g = 0; g < NG; ++g) {
            const int64_t p0 = gr.members[g * gr.G];
            for (size_t jj = 0; jj < s.dec[pl].size(); ++jj) {
                const int r = s.dec[pl][jj];
                const int32_t is = srank[inode[(size_t)p0 * n_nt + r]];
                if (size[is] != gr.G || covered[is]) return RLP_OK;       // shared with another group / rank
                covered[is] = 1;
                for (int k = 0; k < gr.G; ++k) {
                    const int64_t p = gr.members[g * gr.G + k];
                    const int32_t node = inode[(size_t)p * n_nt + r];
                    if (srank[node] != is || krank[node] != k) return RLP_OK;
                }
            }
        }
    }
    // ---- upper groups per player ---------------------------------------------------------------------------------
    std::vector<int> dec_nodes;       // shape index of the d-th decision node
    for (int j = 0; j < n_up; ++j) if (s.dec_idx[j] >= 0) dec_nodes.push_back(j);
    auto tile_node = [&](int64_t tile, int d_idx) { return (*cx.locals)[tile][dec_nodes[d_idx]]; };
    Groups ug[2];
    for (int pl = 0; pl < 2; ++pl) {
        Groups &gr = ug[pl];
        const int d0 = s.udec[pl][0];
        gr.G = size[srank[tile_node(0, d0)]];
        if (gr.G < 1 || gr.G > blk) return RLP_OK;
        std::vector<int32_t> gid_of_iset(I, -1);
        for (int64_t k = 0; k < T; ++k) {
            int32_t node = tile_node(k, d0);
            int32_t key = srank[node], kr = krank[node];
            int32_t g = gid_of_iset[key];
            if (g < 0) {
                g = gid_of_iset[key] = (int32_t)gr.key.size();
                gr.key.push_back(key);
                gr.members.resize(gr.members.size() + gr.G, -1);
            }
            if (size[key] != gr.G || kr >= gr.G || gr.members[(size_t)g * gr.G + kr] >= 0) return RLP_OK;
            gr.members[(size_t)g * gr.G + kr] = k;
        }
        for (int64_t m : gr.members) if (m < 0) return RLP_OK;
        for (size_t g = 0; g < gr.key.size(); ++g)
            for (size_t jj = 0; jj < s.udec[pl].size(); ++jj) {
                const int dd = s.udec[pl][jj];
                const int32_t is = srank[tile_node(gr.members[g * gr.G], dd)];
                if (size[is] != gr.G || covered[is]) return RLP_OK;
                covered[is] = 1;
                for (int k = 0; k < gr.G; ++k) {
                    const int32_t node = tile_node(gr.members[g * gr.G + k], dd);
                    if (srank[node] != is || krank[node] != k) return RLP_OK;
                }
            }
    }
    for (int32_t is = 0; is < I; ++is) if (!covered[is] && size[is] > 0) return RLP_OK;

    // ---- components (instances linked through groups of either player); groups ordered by component -------------
    std::vector<int64_t> uf(P);
    std::iota(uf.begin(), uf.end(), 0);
    auto find = [&](int64_t x) { while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; } return x; };
    for (int pl = 0; pl < 2; ++pl)
        for (size_t g = 0; g < mg[pl].key.size(); ++g)
            for (int k = 1; k < mg[pl].G; ++k) {
                int64_t a = find(mg[pl].members[g * mg[pl].G]), b = find(mg[pl].members[g * mg[pl].G + k]);
                if (a != b) uf[std::max(a, b)] = std::min(a, b);
            }
    std::vector<int32_t> comp_of(P, -1);
    int ncomp = 0;
    for (int64_t p = 0; p < P; ++p) { int64_t r = find(p); if (comp_of[r] < 0) comp_of[r] = ncomp++; comp_of[p] = comp_of[r]; }
    std::vector<int32_t> order[2];
    for (int pl = 0; pl < 2; ++pl) {
        const size_t NG = mg[pl].key.size();
        order[pl].resize(NG);
        std::iota(order[pl].begin(), order[pl].end(), 0);
        std::stable_sort(order[pl].begin(), order[pl].end(), [&](int32_t a, int32_t b) {
            return comp_of[mg[pl].members[(size_t)a * mg[pl].G]] < comp_of[mg[pl].members[(size_t)b * mg[pl].G]];
        });
    }

    // ---- fused numbering of the information sets: opponent gathers of a group should hit consecutive words ---------
    std::vector<int> fid(I, -1);
    int next = 0;
    for (int pl = 0; pl < 2; ++pl) {
        const int opp = 1 - pl;
        for (int32_t g : order[pl])
            for (size_t jj = 0; jj < s.dec[opp].size(); ++jj)
                for (int k = 0; k < mg[pl].G; ++k) {
                    const int64_t p = mg[pl].members[(size_t)g * mg[pl].G + k];
                    const int32_t is = srank[inode[(size_t)p * n_nt + s.dec[opp][jj]]];
                    if (fid[is] < 0) fid[is] = next++;
                }
    }
    for (int pl = 0; pl < 2; ++pl) {
        const int opp = 1 - pl;
        for (size_t g = 0; g < ug[pl].key.size(); ++g)
            for (size_t jj = 0; jj < s.udec[opp].size(); ++jj)
                for (int k = 0; k < ug[pl].G; ++k) {
                    const int32_t is = srank[tile_node(ug[pl].members[g * ug[pl].G + k], s.udec[opp][jj])];
                    if (fid[is] < 0) fid[is] = next++;
                }
    }
    for (int32_t is = 0; is < I; ++is) if (fid[is] < 0) fid[is] = next++;
    const long long NF = I;

    // ---- payoff patterns --------------------------------------------------------------------------------------------
    const int NT = tm.n_term;
    std::vector<int> pat(P);
    std::vector<double> paytab;
    {
        std::map<std::string, int> seen;
        std::string key((size_t)NT * sizeof(double), '\0');
        for (int64_t p = 0; p < P; ++p) {
            for (int tr = 0; tr < NT; ++tr) memcpy(&key[(size_t)tr * sizeof(double)], &(*cx.m_pay1)[(size_t)tr * P + p], sizeof(double));
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, (int)seen.size()).first;
                for (int tr = 0; tr < NT; ++tr) paytab.push_back((*cx.m_pay1)[(size_t)tr * P + p]);
            }
            pat[p] = it->second;
            if (seen.size() > 65536) return RLP_OK;
        }
    }

    // ---- shared-memory budget of the generated kernel (static shared memory only: 48 KB) -----------------------------
    Consts c;
    c.BLK = blk; c.MINB = minb;
    c.NPUSH = 64;
    c.PREFETCH = 0;
    // Multi-GPU: by default only the LAST CTA to arrive at the exchange barrier issues a system-scope fence before it
    // publishes the flag.  The other CTAs' peer stores are ordered before it by causality (store -> bar.sync -> gpu-scope
    // fence + arrive -> the last arriver's gpu-scope acquire -> its system-scope fence; PTX fences are cumulative), and a
    // system fence in each of ~440 CTAs costs 3-4 us per iteration.  RLP_P2P_FENCE_ALL=1 fences in every CTA.
    c.XFENCE_ALL = 0;
    { const char *e = getenv("RLP_P2P_FENCE_ALL"); if (e && e[0] == '1') c.XFENCE_ALL = 1; }
    { const char *e = getenv("RLP_FUSED_PREFETCH"); if (e && e[0] == '1') c.PREFETCH = 1; }
    c.BR_OK = 0; c.MAXA = std::max(1, (int)t->max_act);
    { const char *e = getenv("RLP_P2P_NPUSH"); if (e && atoi(e) >= 1) c.NPUSH = atoi(e); }
    // The whole CTA works on one micro unit (BLK / G groups) by default.  RLP_FUSED_TEAM=warp gives every warp its own
    // unit instead (no CTA-wide barrier in the micro phase; needs G <= 32); measured slower at Leduc-13 because the
    // ordered sums then run on 11 lanes of every warp instead of on 4 full warps per CTA.
    c.TEAM = blk;
    { const char *e = getenv("RLP_FUSED_TEAM"); if (e && e[0] == 'w' && mg[0].G <= 32 && mg[1].G <= 32) c.TEAM = 32; }
    Terms mt[2] = {micro_terms(s, 0), micro_terms(s, 1)}, ut[2] = {upper_terms(s, 0), upper_terms(s, 1)};
    for (int pl = 0; pl < 2; ++pl) {
        c.G[pl] = mg[pl].G; c.GPC[pl] = std::max(1, c.TEAM / mg[pl].G); c.NG[pl] = (int)mg[pl].key.size();
        c.NS[pl] = (long long)c.NG[pl] * c.G[pl];
        c.GB[pl] = ug[pl].G; c.NUG[pl] = (int)ug[pl].key.size();
    }
    c.NF = NF; c.T = T; c.NT = NT; c.KP = cx.kp; c.F = (int)cx.fans->size();
    c.PVLEN = (long long)T * cx.kp; c.fans = *cx.fans;
    c.NCH = cx.fans->empty() ? 0 : (*cx.fans)[0].nchild;
    for (const FanDesc &fd : *cx.fans) if (fd.nchild != c.NCH) c.NCH = 0;
    // one chance probability per fan-in node when all its edges in all tiles agree (Leduc: 1 / #remaining cards)
    for (const FanDesc &fd : *cx.fans) {
        const double v0 = (*cx.fan_cp)[(size_t)fd.e0 * T];
        bool same = true;
        for (int ch = 0; ch < fd.nchild && same; ++ch)
            for (int64_t k = 0; k < T && same; ++k) same = (*cx.fan_cp)[(size_t)(fd.e0 + ch) * T + k] == v0;
        if (!same) { c.fan_cps.clear(); break; }
        c.fan_cps.push_back(v0);
    }
    if (c.fan_cps.size() != cx.fans->size()) c.fan_cps.clear();
    // small integer payoffs: a 1-byte pattern table (exact), rows padded to 16 bytes
    c.PAY8 = 1;
    for (double v : paytab) if (!(v == (double)(int)v && v >= -127.0 && v <= 127.0)) { c.PAY8 = 0; break; }
    c.PAYROW = ((NT + 15) / 16) * 16;
    if (c.PAYROW > 64) c.PAY8 = 0;
    {
        const int nt_ = blk / c.TEAM;
        c.sm_r_n = std::max(nt_ * std::max(mt[0].kr, mt[1].kr) * (c.TEAM + 1), std::max(1, std::max(ut[0].kr * c.GB[0], ut[1].kr * c.GB[1])));
        c.sm_o_n = std::max(nt_ * (int)std::max(s.dec[0].size(), s.dec[1].size()) * (c.TEAM + 1),
                            std::max(1, std::max((int)s.udec[0].size() * c.GB[0], (int)s.udec[1].size() * c.GB[1])));
        c.GBS = std::max(c.GB[0], c.GB[1]);
        c.MBS = (int)std::max(s.udec[0].size(), s.udec[1].size());
        c.UT_OWN = c.GBS + s.n_udec * c.GBS;
        c.UT_N = c.UT_OWN + 3 * c.MBS;
    }
    // ---- best-response kernel: needs an all-chance trunk whose children are trunk nodes or tile roots, and no chance
    // node inside a tile other than the fan-in nodes
    long long n_top = 0;
    {
        bool ok = true;
        for (size_t j = 0; j < sh.kind.size(); ++j) if (sh.kind[j] == 0) ok = false;
        std::vector<uint8_t> is_root((size_t)d->num_nodes, 0);
        if (ok) {
            for (int64_t k = 0; k < T; ++k) { n_top = std::max<long long>(n_top, (*cx.desc)[k].root_global + 1); is_root[(*cx.desc)[k].root_global] = 1; }
            for (long long v = 0; v < n_top && ok; ++v) {
                if (is_root[v]) continue;
                if (d->player[v] != 0 || fc[v + 1] > n_top) ok = false;          // a trunk node: chance, children on top
            }
        }
        int n_tl = 0;
        while (n_tl < t->L && t->level_off[n_tl] < n_top) ++n_tl;         // levels that hold top nodes
        if (n_tl > 8) ok = false;
        const size_t q_team = std::max(s.dec[0].size(), s.dec[1].size()) * (size_t)c.MAXA * (c.TEAM + 1);
        const size_t q_up = std::max(s.udec[0].size() * c.GB[0], s.udec[1].size() * c.GB[1]) * (size_t)c.MAXA;
        const size_t br_smem = 8 * (std::max((size_t)(blk / c.TEAM) * q_team, q_up) + (size_t)std::max(c.GB[0], c.GB[1]) * std::max(1, c.F) + 16) + 4 * (size_t)c.UT_N;
        if (br_smem > 46 * 1024) ok = false;
        { const char *e = getenv("RLP_NO_FUSED_BR"); if (e && e[0] == '1') ok = false; }
        c.BR_OK = ok ? 1 : 0;
    }
    {
        size_t kr = std::max(mt[0].kr, mt[1].kr), mo = std::max(s.dec[0].size(), s.dec[1].size());
        size_t ur = std::max(ut[0].kr * c.GB[0], ut[1].kr * c.GB[1]), uo = std::max(s.udec[0].size() * c.GB[0], s.udec[1].size() * c.GB[1]);
        size_t nt = (size_t)(blk / c.TEAM);
        size_t rn = nt * std::max(c.GPC[0] * mt[0].kr, c.GPC[1] * mt[1].kr);
        size_t bytes = 4 * (size_t)(std::max(c.GB[0], c.GB[1]) * (1 + s.n_udec) + 16) +
                       8 * (std::max(nt * kr * (c.TEAM + 1), ur) + std::max(nt * mo * (c.TEAM + 1), uo) + rn + (size_t)std::max(c.GB[0], c.GB[1]) * std::max(1, c.F) + 64);
        if (bytes > 46 * 1024) return RLP_OK;
    }

    // ---- tables -----------------------------------------------------------------------------------------------------
    std::unique_ptr<FusedPlan, FusedPlanDeleter> fp(new FusedPlan);
    fp->NF = NF; fp->max_act = t->max_act; fp->num_components = ncomp; fp->block = blk;
    const std::vector<int> &ridx = *cx.ridx;
    // upper group of every tile, per player (a player's reach of a portal depends only on that player's upper group)
    std::vector<int> tile_ug[2], tile_mi[2];
    for (int pl = 0; pl < 2; ++pl) {
        tile_ug[pl].assign(T, -1); tile_mi[pl].assign(T, -1);
        for (size_t g = 0; g < ug[pl].key.size(); ++g)
            for (int k = 0; k < ug[pl].G; ++k) {
                tile_ug[pl][ug[pl].members[g * ug[pl].G + k]] = (int)g;
                tile_mi[pl][ug[pl].members[g * ug[pl].G + k]] = k;
            }
    }
    if (2 * (long long)T * cx.kp >= (long long)INT32_MAX) return RLP_OK;
    std::vector<int> tile_of_portal(P), rank_of_portal(P);
    for (int64_t k = 0; k < T; ++k)
        for (int q = 0; q < (*cx.desc)[k].n_portal; ++q) {
            tile_of_portal[(*cx.desc)[k].portal_off + q] = (int)k;
            rank_of_portal[(*cx.desc)[k].portal_off + q] = q;
        }
    if ((long long)T * cx.kp >= (long long)INT32_MAX) return RLP_OK;
    for (int pl = 0; pl < 2; ++pl) {
        const int opp = 1 - pl, G = mg[pl].G, MO = (int)s.dec[pl].size(), MP = (int)s.dec[opp].size();
        const size_t NG = mg[pl].key.size(), NS = NG * G;
        std::vector<int> h_opp((size_t)MP * NS), h_pat(NS), h_ridx(NS), h_pidx(NS), h_fid(NG * MO), h_col(NG * MO), h_iset(NG * MO), h_gridx(NG);
        std::vector<double> h_pc(NS);
        fp->group_comp[pl].assign(NG, 0);
        for (size_t gi = 0; gi < NG; ++gi) {
            const int32_t g = order[pl][gi];
            for (int k = 0; k < G; ++k) {
                const int64_t p = mg[pl].members[(size_t)g * G + k];
                const size_t slot = gi * G + k;
                for (int jj = 0; jj < MP; ++jj) h_opp[(size_t)jj * NS + slot] = fid[srank[inode[(size_t)p * n_nt + s.dec[opp][jj]]]];
                h_pat[slot] = pat[p];
                {   // ridx[p] = reach slot * T + tile (tiles.cu); here: reach slot * #upper groups + the player's upper group
                    const int rslot = (int)(ridx[p] / T), tile = tile_of_portal[p];
                    h_ridx[slot] = rslot * c.NUG[opp] + tile_ug[opp][tile];
                    const int own = rslot * c.NUG[pl] + tile_ug[pl][tile];
                    if (k == 0) h_gridx[gi] = own;
                    else if (h_gridx[gi] != own) return RLP_OK;      // members of a group must share their own reach
                }
                {   // value slot: [pass][upper group of this player][portal][member of the upper group]
                    const int tile = tile_of_portal[p];
                    h_pidx[slot] = (int)((long long)pl * T * cx.kp +
                                         ((long long)tile_ug[pl][tile] * cx.kp + rank_of_portal[p]) * ug[pl].G + tile_mi[pl][tile]);
                }
                h_pc[slot] = (*cx.pcs)[(*cx.portal_node)[p]];
            }
            const int64_t p0 = mg[pl].members[(size_t)g * G];
            for (int jj = 0; jj < MO; ++jj) {
                const int32_t is = srank[inode[(size_t)p0 * n_nt + s.dec[pl][jj]]];
                h_fid[gi * MO + jj] = fid[is];
                h_col[gi * MO + jj] = (int)t->col_off_h[is];
                h_iset[gi * MO + jj] = is;
            }
            fp->group_comp[pl][gi] = comp_of[p0];
        }
        RLP_CUDA(fp->mem_opp[pl].upload(h_opp, stream)); RLP_CUDA(fp->mem_pat[pl].upload(h_pat, stream));
        RLP_CUDA(fp->mem_ridx[pl].upload(h_ridx, stream)); RLP_CUDA(fp->mem_pidx[pl].upload(h_pidx, stream));
        RLP_CUDA(fp->grp_ridx[pl].upload(h_gridx, stream));
        RLP_CUDA(fp->reach[pl].alloc((size_t)cx.kp * (size_t)c.NUG[pl]));
        RLP_CUDA(fp->mem_pc[pl].upload(h_pc, stream)); RLP_CUDA(fp->grp_fid[pl].upload(h_fid, stream));
        RLP_CUDA(fp->grp_col[pl].upload(h_col, stream)); RLP_CUDA(fp->grp_iset[pl].upload(h_iset, stream));
        fp->h_grp_iset[pl] = h_iset; fp->mo[pl] = MO;

        fp->n_batches[pl] = (int)((NG + c.GPC[pl] - 1) / c.GPC[pl]);
        // upper groups
        const int GB = ug[pl].G, MB = (int)s.udec[pl].size();
        const size_t NU = ug[pl].key.size();
        std::vector<int> h_tile(NU * GB), h_ufid(NU * MB), h_ucol(NU * MB), h_uiset(NU * MB);
        for (size_t g = 0; g < NU; ++g) {
            for (int k = 0; k < GB; ++k) h_tile[g * GB + k] = (int)ug[pl].members[g * GB + k];
            for (int jj = 0; jj < MB; ++jj) {
                const int32_t is = srank[tile_node(ug[pl].members[g * GB], s.udec[pl][jj])];
                h_ufid[g * MB + jj] = fid[is]; h_ucol[g * MB + jj] = (int)t->col_off_h[is]; h_uiset[g * MB + jj] = is;
            }
        }
        RLP_CUDA(fp->tg_tile[pl].upload(h_tile, stream)); RLP_CUDA(fp->tg_fid[pl].upload(h_ufid, stream));
        RLP_CUDA(fp->tg_col[pl].upload(h_ucol, stream)); RLP_CUDA(fp->tg_iset[pl].upload(h_uiset, stream));
        fp->h_tg_iset[pl] = h_uiset;
        fp->n_ugroups[pl] = (int)NU;
        RLP_CUDA(rlp::sync_stream(stream));
    }
    {
        std::vector<int> h_ufid((size_t)s.n_udec * T);
        for (int64_t k = 0; k < T; ++k)
            for (int dd = 0; dd < s.n_udec; ++dd) h_ufid[(size_t)dd * T + k] = fid[srank[tile_node(k, dd)]];
        RLP_CUDA(fp->u_fid.upload(h_ufid, stream));
        RLP_CUDA(fp->paytab.upload(paytab, stream));
        {
            std::vector<int> roots_h(T);
            for (int64_t k = 0; k < T; ++k) roots_h[k] = (*cx.desc)[k].root_global;
            RLP_CUDA(fp->tile_root.upload(roots_h, stream));
            fp->n_top = n_top; fp->kp = cx.kp; fp->nug[0] = c.NUG[0]; fp->nug[1] = c.NUG[1];
        }
        {
            std::vector<signed char> p8;
            if (c.PAY8) {
                const size_t npat = paytab.size() / (size_t)NT;
                p8.assign(npat * c.PAYROW, 0);
                for (size_t q = 0; q < npat; ++q)
                    for (int tr = 0; tr < NT; ++tr) p8[q * c.PAYROW + tr] = (signed char)(int)paytab[q * NT + tr];
            }
            RLP_CUDA(fp->paytab8.upload(p8, stream));
        }
        RLP_CUDA(fp->fid_of_iset.upload(fid, stream));
        RLP_CUDA(fp->bar.alloc(32 + 32 * 4096 + 1024));      // arrivals, leader count | one flag line per CTA | per-SM counters
        RLP_CUDA(rlp::sync_stream(stream));
    }
    // ---- generate, compile, load -----------------------------------------------------------------------------------------
    fp->source = generate_source(s, c);
    std::string log;
    int rc = compile_and_load_source(fp->source, &fp->lib, &log);
    if (rc) return rc;
    FusedParams &h = fp->host;
    UpperJit &uj = *ts.upper_jit;
    h.mem_opp0 = fp->mem_opp[0].p; h.mem_opp1 = fp->mem_opp[1].p; h.mem_pat0 = fp->mem_pat[0].p; h.mem_pat1 = fp->mem_pat[1].p;
    h.grp_ridx0 = fp->grp_ridx[0].p; h.grp_ridx1 = fp->grp_ridx[1].p; h.paytab8 = fp->paytab8.p;
    h.mem_ridx0 = fp->mem_ridx[0].p; h.mem_ridx1 = fp->mem_ridx[1].p; h.mem_pidx0 = fp->mem_pidx[0].p; h.mem_pidx1 = fp->mem_pidx[1].p;
    h.mem_pc0 = fp->mem_pc[0].p; h.mem_pc1 = fp->mem_pc[1].p; h.grp_fid0 = fp->grp_fid[0].p; h.grp_fid1 = fp->grp_fid[1].p;
    h.grp_col0 = fp->grp_col[0].p; h.grp_col1 = fp->grp_col[1].p; h.grp_iset0 = fp->grp_iset[0].p; h.grp_iset1 = fp->grp_iset[1].p;
    h.paytab = fp->paytab.p;
    h.tg_tile0 = fp->tg_tile[0].p; h.tg_tile1 = fp->tg_tile[1].p; h.tg_fid0 = fp->tg_fid[0].p; h.tg_fid1 = fp->tg_fid[1].p;
    h.tg_col0 = fp->tg_col[0].p; h.tg_col1 = fp->tg_col[1].p; h.tg_iset0 = fp->tg_iset[0].p; h.tg_iset1 = fp->tg_iset[1].p;
    h.u_fid = fp->u_fid.p; h.u_pcs = uj.pcs.p; h.u_cprob = uj.cprob.p; h.u_pay1 = uj.pay1.p;
    h.fan_desc = reinterpret_cast<const int *>(uj.fan.p); h.fan_cp = uj.fan_cp.p;
    h.portal_p1 = fp->reach[0].p; h.portal_p2 = fp->reach[1].p;
    RLP_CUDA(fp->pvbuf.alloc(2 * (size_t)T * (size_t)cx.kp));
    h.pvl0 = h.pvl1 = h.pvA0 = h.pvB0 = fp->pvbuf.p;        // single GPU: one buffer, this rank only
    h.nranks = 1; h.rank = 0; h.xbase = 0; h.xfl = nullptr;
    h.tile_root = fp->tile_root.p; h.g_first_child = t->first_child.p; h.g_player = reinterpret_cast<const signed char *>(t->player.p);
    h.n_top = n_top;
    {
        long long *lv = &h.tl0;
        int n_tl = 0;
        while (n_tl < t->L && n_tl < 8 && t->level_off[n_tl] < n_top) { lv[n_tl] = t->level_off[n_tl]; ++n_tl; }
        lv[n_tl] = n_tl < t->L ? t->level_off[n_tl] : n_top;
        h.n_tl = n_tl;
    }
    fp->pv_len = (size_t)T * (size_t)cx.kp;
    h.bar = fp->bar.p; h.tl = nullptr; h.tl_cap = 0;
    h.num_isets = I;
    h.first_group0 = 0; h.end_group0 = c.NG[0]; h.first_group1 = 0; h.end_group1 = c.NG[1];
    if (!rlp::g_dry_run) {
        if (!fp->lib) return RLP_OK;                       // RLP_NO_JIT=1
        if (cudaLibraryGetKernel(&fp->kernel, fp->lib, "cfr_fused_k") != cudaSuccess) {
            cudaGetLastError();
            return fail(RLP_ERR_CUDA, "generated library has no cfr_fused_k");
        }
        if (c.BR_OK && cudaLibraryGetKernel(&fp->br_kernel, fp->lib, "br_fused_k") != cudaSuccess) {
            cudaGetLastError();
            return fail(RLP_ERR_CUDA, "generated library has no br_fused_k");
        }
        int dev = 0, sms = 0, per_sm = 0;
        RLP_CUDA(cudaGetDevice(&dev));
        RLP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fp->kernel, blk, 0) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 1;
        }
        const char *e = getenv("RLP_FUSED_CTAS_PER_SM");
        if (e && atoi(e) >= 1) per_sm = std::min(per_sm, atoi(e));
        fp->grid = std::min(4096, sms * std::min(per_sm, minb));
    } else {
        fp->grid = 296;
    }
    ts.fused = std::move(fp);
    return RLP_OK;
}

bool fused_available(const rlp_cfr *c) {
    const TileSet &ts = c->tree->tiles;
    if (!ts.enabled || !ts.fused || !ts.fused->kernel) return false;
    const char *off = getenv("RLP_NO_FUSED");
    return !(off && off[0] == '1');
}

int fused_stat(const rlp_tree *t, int which) {
    const FusedPlan *f = t->tiles.fused.get();
    if (!f) return which == 0 ? 0 : -1;
    switch (which) {
        case 0: return f->kernel ? 1 : 0;
        case 1: return f->n_batches[0] + f->n_batches[1];
        case 2: return f->n_ugroups[0] + f->n_ugroups[1];
        case 3: return f->grid;
        case 4: return f->num_components;
        case 5: return f->br_kernel ? 1 : 0;
        default: return -1;
    }
}

const std::string *fused_source(const rlp_tree *t) { return t->tiles.fused ? &t->tiles.fused->source : nullptr; }

bool fused_br_available(const rlp_cfr *c) {
    const TileSet &ts = c->tree->tiles;
    if (!ts.enabled || !ts.fused || !ts.fused->br_kernel) return false;
    const char *off = getenv("RLP_NO_FUSED_BR");
    return !(off && off[0] == '1');
}

// Best responses of BOTH players against sigma_int ([Q], device, internal pair order) in one cooperative launch:
// values into out2_dev[0..1], chosen actions into c->astar (internal information-set ids).
int fused_br(rlp_cfr *c, const double *sigma_int, double *out2_dev, cudaStream_t s) {
    rlp_tree *t = c->tree;
    FusedPlan &f = *t->tiles.fused;
    const size_t plane = (size_t)f.NF * (size_t)std::max(1, f.max_act);
    const size_t reach_n = (size_t)f.kp * (size_t)std::max(f.nug[0], f.nug[1]);
    const size_t need = plane + 2 * reach_n + 2 * f.pv_len + 2 * (size_t)f.n_top + 8;
    if (c->br_buf.n < need) RLP_CUDA(c->br_buf.alloc(need));
    double *base = c->br_buf.p;
    if (t->I > 0) {
        k_sigma_fused<<<(unsigned)((t->I + 255) / 256), 256, 0, s>>>(t->I, f.NF, t->col_off.p, f.fid_of_iset.p, sigma_int, base);
        RLP_LAUNCH_CHECK();
    }
    FusedParams hp = f.host;
    hp.br_sig = base;
    hp.br_reach1 = base + plane; hp.br_reach2 = hp.br_reach1 + reach_n;
    hp.br_pv1 = hp.br_reach2 + reach_n; hp.br_pv2 = hp.br_pv1 + f.pv_len;     // one array: [pass][group][portal][member]
    hp.br_top1 = hp.br_pv2 + f.pv_len; hp.br_top2 = hp.br_top1 + f.n_top;
    hp.br_out = out2_dev; hp.br_astar = c->astar.p;
    for (;;) {
        RLP_CUDA(cudaMemsetAsync(f.bar.p, 0, sizeof(unsigned int) * (32 + 32 * (size_t)f.grid + 1024), s));
        void *args[] = {&hp};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)f.grid); cfg.blockDim = dim3((unsigned)f.block); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)f.br_kernel, args);
        if (e == cudaErrorCooperativeLaunchTooLarge && f.grid > 1) {
            cudaGetLastError();
            int dev = 0, sms = 0;
            RLP_CUDA(cudaGetDevice(&dev));
            RLP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            if (f.grid <= sms) return fail(RLP_ERR_CUDA, "cooperative launch of %d CTAs refused", f.grid);
            f.grid = sms;
            continue;
        }
        if (e != cudaSuccess) return fail(RLP_ERR_CUDA, "launch of the fused best-response kernel: %s", cudaGetErrorString(e));
        break;
    }
    rlp::g_launches.fetch_add(1, std::memory_order_relaxed);
    return RLP_OK;
}

// Multi-GPU state of one solver (sharding by public state, see rlp_cfr_p2p_* below).
struct P2P {
    void *base = nullptr;               // this rank's exchange buffer: [portal values, even | odd | 8 flag words], IPC-exported
    void *peer[8] = {};                 // the other ranks' buffers, opened through CUDA IPC
    int rank = 0, nranks = 1;
    bool active = false;
    long long g_lo[2] = {0, 0}, g_hi[2] = {0, 0};   // this rank's micro groups per player (whole components)
    int c_lo = 0, c_hi = 0;
    unsigned long long xcount = 0;      // iterations exchanged so far = the last flag token
};

void p2p_release(rlp_cfr *c) {
    P2P *x = static_cast<P2P *>(c->p2p);
    if (!x) return;
    for (int r = 0; r < 8; ++r) if (x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->base) cudaFree(x->base);
    delete x;
    c->p2p = nullptr;
}

// n iterations t0 .. t0+n-1 in ONE cooperative launch.  `tl` (device, nullable) receives globaltimer stamps of CTA 0:
// per iteration {start, phase A done, barrier passed, phase B done}.
int fused_run(rlp_cfr *c, int64_t t0, int64_t n_iters, int linear_weight, cudaStream_t s, unsigned long long *tl, long long tl_cap) {
    rlp_tree *t = c->tree;
    FusedPlan &f = *t->tiles.fused;
    const size_t plane = (size_t)f.NF * (size_t)std::max(1, f.max_act);
    if (!c->sigF.p) {
        RLP_CUDA(c->sigF.alloc(2 * plane));
        RLP_CUDA(cudaMemsetAsync(c->sigF.p, 0, c->sigF.bytes(), s));
        c->sigF_dirty = true;
    }
    if (c->sigF_dirty) {
        if (t->I > 0) {
            k_sigma_fused<<<(unsigned)((t->I + 255) / 256), 256, 0, s>>>(t->I, f.NF, t->col_off.p, f.fid_of_iset.p, c->sigma.p, c->sigF.p);
            RLP_LAUNCH_CHECK();
        }
        c->sigF_cur = 0;
        c->sigF_dirty = false;
    }
    FusedParams hp = f.host;                              // passed BY VALUE (kernel parameter space)
    hp.R = c->R.p; hp.S = c->S.p; hp.sigma_row = c->sigma.p; hp.sigF0 = c->sigF.p; hp.sigF1 = c->sigF.p + plane;
    hp.flags = c->flags.p; hp.tl = tl; hp.tl_cap = tl_cap;
    P2P *x = static_cast<P2P *>(c->p2p);
    if (x && x->active) {                                 // sharded by public state: this rank's components only
        double **pa = &hp.pvA0, **pb = &hp.pvB0;
        unsigned long long **xf = &hp.xfp0;
        for (int r = 0; r < x->nranks; ++r) {
            char *base = static_cast<char *>(r == x->rank ? x->base : x->peer[r]);
            pa[r] = reinterpret_cast<double *>(base);
            pb[r] = reinterpret_cast<double *>(base) + 2 * f.pv_len;
            xf[r] = reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(base) + 4 * f.pv_len);
        }
        hp.pvl0 = pa[x->rank]; hp.pvl1 = pb[x->rank]; hp.xfl = xf[x->rank];
        hp.nranks = x->nranks; hp.rank = x->rank;
        hp.first_group0 = x->g_lo[0]; hp.end_group0 = x->g_hi[0]; hp.first_group1 = x->g_lo[1]; hp.end_group1 = x->g_hi[1];
    }
    int64_t left = n_iters, tt = t0;
    while (left > 0) {
        const int chunk = (int)std::min<int64_t>(left, 1 << 20);
        RLP_CUDA(cudaMemsetAsync(f.bar.p, 0, sizeof(unsigned int) * (32 + 32 * (size_t)f.grid + 1024), s));
        long long t_arg = tt;
        if (x && x->active) hp.xbase = (long long)x->xcount;
        int iters = chunk, lin = linear_weight ? 1 : 0, cur0 = c->sigF_cur;
        void *args[] = {&hp, &t_arg, &iters, &lin, &cur0};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)f.grid); cfg.blockDim = dim3((unsigned)f.block); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)f.kernel, args);
        if (e == cudaErrorCooperativeLaunchTooLarge && f.grid > 1) {
            cudaGetLastError();
            int dev = 0, sms = 0;
            RLP_CUDA(cudaGetDevice(&dev));
            RLP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            if (f.grid <= sms) return fail(RLP_ERR_CUDA, "cooperative launch of %d CTAs refused", f.grid);
            f.grid = sms;
            continue;
        }
        if (e != cudaSuccess) return fail(RLP_ERR_CUDA, "launch of the fused iteration kernel: %s", cudaGetErrorString(e));
        rlp::g_launches.fetch_add(1, std::memory_order_relaxed);
        c->sigF_cur = (c->sigF_cur + chunk) & 1;
        if (x && x->active) x->xcount += (unsigned long long)chunk;
        left -= chunk; tt += chunk;
    }
    c->sigma_am_dirty = true;            // the five-kernel pipeline's action-major copy is stale now
    if (x && x->active && getenv("RLP_P2P_DEBUG")) {      // where the last arriver of the cross-GPU barrier spent its time
        unsigned long long h[5] = {0, 0, 0, 0, 0};
        cudaStreamSynchronize(s);
        cudaMemcpy(h, f.bar.p + 8, sizeof(h), cudaMemcpyDeviceToHost);
        if (h[0])
            fprintf(stderr, "[rlp p2p rank %d] %llu exchanges: fence %.2f us, flag stores %.2f us, wait for peers %.2f us, fence %.2f us\n",
                    x->rank, h[0], 1e-3 * h[1] / h[0], 1e-3 * h[2] / h[0], 1e-3 * h[3] / h[0], 1e-3 * h[4] / h[0]);
    }
    return RLP_OK;
}

}  // namespace rlp

// ---- multi-GPU: sharding by public state ------------------------------------------------------------------------
// The micro subtrees fall into COMPONENTS (in Leduc: one per board card and round-1 betting line) that share no
// information set; rank r sweeps the components [r * n / W, (r + 1) * n / W) and is the only one to hold current
// regrets / action counts / strategy for their information sets.  What the ranks exchange is the VALUE of every micro
// subtree (8 bytes each, 1.1 MB per iteration at Leduc-13 in total): the micro phase stores it straight into every
// peer's buffer over NVLink, and the grid barrier that ends the phase doubles as the cross-GPU handshake (fused.cu,
// grid_barrier_x).  The upper phase (round-1 information sets: 208 of 47 008 at Leduc-13) is replicated.  Every sum
// keeps the single-GPU order, so the result is bit-identical to one GPU and to the reference -- unlike the all-reduce of
// regret deltas (comm.cu), which changes the association.
extern "C" int rlp_cfr_p2p_export(rlp_cfr *c, unsigned char *handle64) {
    if (!c || !handle64) return fail(RLP_ERR_INVALID, "null argument");
    if (!rlp::fused_available(c)) return fail(RLP_ERR_UNSUPPORTED, "sharding by public state needs the group-fused kernel");
    rlp::p2p_release(c);
    FusedPlan &f = *c->tree->tiles.fused;
    std::unique_ptr<rlp::P2P> x(new rlp::P2P);
    const size_t bytes = sizeof(double) * 4 * f.pv_len + sizeof(unsigned long long) * 8;     // [parity][pass][T * KP] | flags
    RLP_CUDA(cudaMalloc(&x->base, bytes));
    RLP_CUDA(cudaMemset(x->base, 0, bytes));
    cudaIpcMemHandle_t h;
    RLP_CUDA(cudaIpcGetMemHandle(&h, x->base));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    c->p2p = x.release();
    return RLP_OK;
}

extern "C" int rlp_cfr_p2p_connect(rlp_cfr *c, int rank, int nranks, const unsigned char *handles) {
    if (!c || !handles || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) return fail(RLP_ERR_INVALID, "bad argument");
    rlp::P2P *x = static_cast<rlp::P2P *>(c->p2p);
    if (!x || !x->base) return fail(RLP_ERR_INVALID, "rlp_cfr_p2p_export has not been called");
    FusedPlan &f = *c->tree->tiles.fused;
    x->rank = rank; x->nranks = nranks;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        RLP_CUDA(cudaIpcOpenMemHandle(&x->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    const int n = f.num_components;
    x->c_lo = (int)((long long)rank * n / nranks); x->c_hi = (int)((long long)(rank + 1) * n / nranks);
    for (int pl = 0; pl < 2; ++pl) {
        const std::vector<int> &gc = f.group_comp[pl];
        x->g_lo[pl] = std::lower_bound(gc.begin(), gc.end(), x->c_lo) - gc.begin();
        x->g_hi[pl] = std::lower_bound(gc.begin(), gc.end(), x->c_hi) - gc.begin();
    }
    x->xcount = 0;
    x->active = true;
    return RLP_OK;
}

// owned[I] (canonical ids): 1 for the information sets whose regrets / action counts / strategy are current on this
// rank and that this rank contributes when the ranks' arrays are merged (round-1 sets are replicated: rank 0 owns them).
extern "C" int rlp_cfr_p2p_owned(rlp_cfr *c, unsigned char *owned) {
    if (!c || !owned) return fail(RLP_ERR_INVALID, "null argument");
    rlp::P2P *x = static_cast<rlp::P2P *>(c->p2p);
    if (!x || !x->active) return fail(RLP_ERR_INVALID, "not connected");
    rlp_tree *t = c->tree;
    FusedPlan &f = *t->tiles.fused;
    memset(owned, 0, (size_t)t->I);
    for (int pl = 0; pl < 2; ++pl) {
        for (long long g = x->g_lo[pl]; g < x->g_hi[pl]; ++g)
            for (int j = 0; j < f.mo[pl]; ++j) owned[t->iset_of_srank[f.h_grp_iset[pl][(size_t)g * f.mo[pl] + j]]] = 1;
        if (x->rank == 0)
            for (int is : f.h_tg_iset[pl]) owned[t->iset_of_srank[is]] = 1;
    }
    return RLP_OK;
}

// 1 when the last sharded launch gave up waiting for a peer (the solver state is then undefined).
extern "C" int rlp_cfr_p2p_failed(rlp_cfr *c) {
    if (!c || !c->tree->tiles.fused) return 0;
    unsigned int flag = 0;
    cudaMemcpy(&flag, c->tree->tiles.fused->bar.p + 2, sizeof(flag), cudaMemcpyDeviceToHost);
    return flag != 0;
}
