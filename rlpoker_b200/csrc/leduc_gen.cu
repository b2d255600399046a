// Native Leduc-n generator (host code): emits, directly in breadth-first structure-of-arrays form, the
// tree that rlpoker/games/leduc.py:44-278 `Leduc(cards)` builds out of Python objects, node for node:
// node i here is the i-th node that reference builder dequeues, children in the same order, the same
// chance probabilities (1 / #remaining cards as an IEEE double), pots and showdown payoffs.
// The Python builder needs ~1.6 kB and ~20 us per node (5 GB / 75 s at 13 ranks); this needs ~40 B.
#include <deque>
#include <unordered_map>

#include "common.cuh"

using rlp::fail;

namespace {

struct Spec {
    int32_t values, suits, max_raises, raise_amount;
    bool operator==(const Spec &o) const {
        return values == o.values && suits == o.suits && max_raises == o.max_raises && raise_amount == o.raise_amount;
    }
};

struct Out {
    std::vector<int32_t> parent, infoset, label;
    std::vector<int8_t> player;
    std::vector<uint8_t> hidden;
    std::vector<double> cprob, u1, u2;
    int32_t num_infosets = 0;
    int64_t size() const { return (int64_t)parent.size(); }
    int64_t add(int32_t par, int8_t pl, int32_t lab, double cp) {
        parent.push_back(par); player.push_back(pl); label.push_back(lab); cprob.push_back(cp);
        infoset.push_back(-1); hidden.push_back(0); u1.push_back(0.0); u2.push_back(0.0);
        return size() - 1;
    }
};

enum Stage : uint8_t { DEAL1, DEAL2, BET, BOARD };

// What the builder needs to know about a node waiting to be expanded.
struct Pending {
    int32_t id;
    Stage stage;
    uint8_t round;        // 1 or 2 (BET)
    uint8_t player;       // acting player (BET)
    uint8_t last;         // 0xff none, else last action of this round
    uint8_t nbets;        // bets so far in this round
    int16_t hole1, hole2, board;
    int32_t pot1, pot2;
    uint32_t code1, code2;   // betting sequences of round 1 / 2, base 4 digits (action + 1)
};

constexpr int kFold = 0, kCall = 1, kBet = 2, kCardLabel = 16;

struct KeyHash {
    size_t operator()(uint64_t k) const {
        k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
        return (size_t)k;
    }
};

int generate(const Spec &sp, Out &o) {
    const int C = sp.values * sp.suits;
    if (sp.values < 1 || sp.suits < 1 || C < 3 || C > 250) return fail(RLP_ERR_INVALID, "deck of %d cards", C);
    if (sp.max_raises < 1 || sp.max_raises > 6) return fail(RLP_ERR_UNSUPPORTED, "max_raises must be 1..6");
    const int stake[3] = {0, sp.raise_amount, 2 * sp.raise_amount};
    std::unordered_map<uint64_t, int32_t, KeyHash> isets;
    std::deque<Pending> queue;

    auto value_of = [&](int card) { return card / sp.suits; };
    auto showdown = [&](const Pending &p, double &a, double &b) {   // leduc.py:237-278
        int r1 = value_of(p.hole1), r2 = value_of(p.hole2), rb = value_of(p.board);
        bool pair1 = r1 == rb, pair2 = r2 == rb;
        int win = 0;
        if (pair1 && pair2) win = 0;
        else if (pair1) win = 1;
        else if (pair2) win = 2;
        else win = r1 > r2 ? 1 : (r2 > r1 ? 2 : 0);
        if (win == 1) { a = p.pot2; b = -(double)p.pot2; }
        else if (win == 2) { a = -(double)p.pot1; b = p.pot1; }
        else { a = 0.0; b = 0.0; }
    };
    auto infoset_of = [&](const Pending &p) {
        // visible to the actor: own hole card, both betting sequences, the board (extensive_game.py:286-335)
        uint64_t own = p.player == 1 ? (uint64_t)p.hole1 : (uint64_t)p.hole2;
        uint64_t board = p.round == 2 ? (uint64_t)p.board + 1 : 0;
        uint64_t key = (uint64_t)(p.player - 1) | own << 1 | board << 9 | (uint64_t)p.code1 << 17 | (uint64_t)p.code2 << 40;
        auto it = isets.find(key);
        if (it != isets.end()) return it->second;
        int32_t id = (int32_t)isets.size();
        isets.emplace(key, id);
        return id;
    };

    Pending root{};
    root.id = (int32_t)o.add(-1, 0, 0, 1.0);
    root.stage = DEAL1; root.hole1 = root.hole2 = root.board = -1; root.pot1 = root.pot2 = 1; root.last = 0xff;
    o.hidden[0] = 2;                                        // hidden_from = [2]  (leduc.py:69)
    queue.push_back(root);

    while (!queue.empty()) {
        Pending p = queue.front();
        queue.pop_front();
        if (o.size() > (int64_t)INT32_MAX - 512) return fail(RLP_ERR_UNSUPPORTED, "tree exceeds 2^31 nodes");
        if (p.stage == DEAL1 || p.stage == DEAL2 || p.stage == BOARD) {
            int remaining = C - (p.stage == DEAL1 ? 0 : p.stage == DEAL2 ? 1 : 2);
            double prob = 1.0 / (double)remaining;         // n / len(remaining_cards), n == 1
            for (int card = 0; card < C; ++card) {
                if (card == p.hole1 || card == p.hole2) continue;
                Pending c = p;
                if (p.stage == DEAL1) {
                    c.id = (int32_t)o.add(p.id, 0, kCardLabel + card, prob);
                    o.hidden[c.id] = 1;                    // hidden_from = [1]  (leduc.py:80)
                    c.stage = DEAL2; c.hole1 = (int16_t)card;
                } else if (p.stage == DEAL2) {
                    c.id = (int32_t)o.add(p.id, 1, kCardLabel + card, prob);
                    c.stage = BET; c.round = 1; c.player = 1; c.hole2 = (int16_t)card; c.last = 0xff; c.nbets = 0;
                } else {
                    c.id = (int32_t)o.add(p.id, 2, kCardLabel + card, prob);
                    c.stage = BET; c.round = 2; c.player = 2; c.board = (int16_t)card; c.last = 0xff; c.nbets = 0;
                }
                queue.push_back(c);
            }
            continue;
        }
        // betting node of p.player
        o.infoset[p.id] = infoset_of(p);
        const uint8_t other = (uint8_t)(3 - p.player);
        auto extend = [&](Pending &c, int action) {
            if (c.round == 1) c.code1 = c.code1 * 4 + (uint32_t)(action + 1); else c.code2 = c.code2 * 4 + (uint32_t)(action + 1);
        };
        auto continue_betting = [&](int action, bool pay) {
            Pending c = p;
            extend(c, action);
            if (pay) { if (p.player == 1) c.pot1 += stake[p.round]; else c.pot2 += stake[p.round]; }
            c.player = other; c.last = (uint8_t)action; c.nbets = (uint8_t)(p.nbets + (action == kBet));
            c.id = (int32_t)o.add(p.id, (int8_t)other, action, 1.0);
            queue.push_back(c);
        };
        auto end_round = [&](int action, bool pay) {
            Pending c = p;
            extend(c, action);
            if (pay) { if (p.player == 1) c.pot1 += stake[p.round]; else c.pot2 += stake[p.round]; }
            if (p.round == 1) {
                c.id = (int32_t)o.add(p.id, 0, action, 1.0);
                c.stage = BOARD;
                queue.push_back(c);
            } else {
                c.id = (int32_t)o.add(p.id, -1, action, 1.0);
                showdown(c, o.u1[c.id], o.u2[c.id]);
            }
        };
        if (p.last == 0xff) {                               // leduc.py:113-131
            continue_betting(kCall, false);
            continue_betting(kBet, true);
        } else if (p.last == kCall) {                       // leduc.py:132-157
            end_round(kCall, false);
            continue_betting(kBet, true);
        } else {                                            // facing a bet, leduc.py:158-214
            end_round(kCall, true);
            int64_t f = o.add(p.id, -1, kFold, 1.0);
            int lost = p.player == 1 ? p.pot1 : p.pot2;     // leduc.py:215-222
            o.u1[f] = p.player == 1 ? -(double)lost : (double)lost;
            o.u2[f] = -o.u1[f];
            if (p.nbets < sp.max_raises - 1) continue_betting(kBet, true);
            else end_round(kBet, true);
        }
    }
    o.num_infosets = (int32_t)isets.size();
    return RLP_OK;
}

struct Cache { bool valid = false; Spec spec{}; Out out; };
thread_local Cache g_cache;

}  // namespace

extern "C" int rlp_leduc_generate(int32_t num_values, int32_t num_suits, int32_t max_raises, int32_t raise_amount,
                                  int64_t *num_nodes, int32_t *num_infosets, int32_t *parent, int8_t *player,
                                  int32_t *infoset, double *cprob, double *u1, double *u2, int32_t *label,
                                  uint8_t *hidden) {
    if (!num_nodes || !num_infosets) return fail(RLP_ERR_INVALID, "null size pointer");
    Spec sp{num_values, num_suits, max_raises, raise_amount};
    if (!(g_cache.valid && g_cache.spec == sp)) {
        g_cache = Cache();
        int rc = generate(sp, g_cache.out);
        if (rc) { g_cache = Cache(); return rc; }
        g_cache.valid = true;
        g_cache.spec = sp;
    }
    Out &o = g_cache.out;
    if (!parent) {                                          // size query; keep the result for the fill call
        *num_nodes = o.size();
        *num_infosets = o.num_infosets;
        return RLP_OK;
    }
    if (*num_nodes != o.size()) return fail(RLP_ERR_INVALID, "buffer size %lld != %lld nodes", (long long)*num_nodes, (long long)o.size());
    if (!player || !infoset || !cprob || !u1 || !u2 || !label || !hidden) return fail(RLP_ERR_INVALID, "null output array");
    size_t n = (size_t)o.size();
    memcpy(parent, o.parent.data(), n * sizeof(int32_t));
    memcpy(player, o.player.data(), n);
    memcpy(infoset, o.infoset.data(), n * sizeof(int32_t));
    memcpy(cprob, o.cprob.data(), n * sizeof(double));
    memcpy(u1, o.u1.data(), n * sizeof(double));
    memcpy(u2, o.u2.data(), n * sizeof(double));
    memcpy(label, o.label.data(), n * sizeof(int32_t));
    memcpy(hidden, o.hidden.data(), n);
    *num_infosets = o.num_infosets;
    g_cache = Cache();
    return RLP_OK;
}
