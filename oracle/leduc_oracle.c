/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see cfr_oracle.c).  Plain-C restatement of the reference's Leduc builder,
 *   rlpoker/games/leduc.py:44-234   (create_tree: FIFO deque, children in insertion order)
 *   rlpoker/games/leduc.py:237-278  (compute_showdown)
 *   rlpoker/games/card.py:6-7       (get_deck: value-major, suits innermost)
 *   rlpoker/extensive_game.py:286-335 (information-set ids: the actor sees its own hole card, the board and all bets)
 * straight into the breadth-first arrays cfr_oracle.c works on, so that the CPU baseline / --impl reference arm of
 * bench.py can build Leduc-13 (3.2 M nodes) without the product package and without 5 GB of Python objects.
 * Pinned: tests/test_oracle_golden.py compares its output for n = 2, 3, 4 with the layouts the reference itself
 * produced (tests/golden/reference_arrays.npz), and tests/test_leduc_native.py with the product's generator.
 *
 * Node i is the i-th node the reference pops from its deque; the queue below is the output array itself (a node's
 * pending state lives next to it until it has been expanded).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ST_DEAL1, ST_DEAL2, ST_BOARD, ST_ROUND1, ST_ROUND2, ST_LEAF };
enum { A_FOLD = 0, A_CALL = 1, A_BET = 2, CARD_LABEL = 16 };

typedef struct {
    int8_t state, actor;
    int16_t hole1, hole2, board;
    int8_t last;          /* -1: no bet action yet in this round, else the last action */
    int8_t nbets;         /* bets so far in this round */
    int32_t pot1, pot2;
    uint32_t seq1, seq2;  /* betting sequences of the two rounds, base-4 digits (action + 1) */
} pend_t;

typedef struct {
    int64_t n, cap;
    int32_t *parent, *infoset, *label;
    int8_t *player;
    uint8_t *hidden;
    double *cprob, *u1, *u2;
    pend_t *pend;
} out_t;

static int grow(out_t *o) {
    int64_t cap = o->cap ? o->cap * 2 : 1 << 16;
#define GROW(f, T) { T *p = (T *)realloc(o->f, (size_t)cap * sizeof(T)); if (!p) return -1; o->f = p; }
    GROW(parent, int32_t) GROW(infoset, int32_t) GROW(label, int32_t) GROW(player, int8_t) GROW(hidden, uint8_t)
    GROW(cprob, double) GROW(u1, double) GROW(u2, double) GROW(pend, pend_t)
#undef GROW
    o->cap = cap;
    return 0;
}

static int64_t push(out_t *o, int32_t parent, int player, int label, double cp, const pend_t *p) {
    if (o->n == o->cap && grow(o)) return -1;
    int64_t i = o->n++;
    o->parent[i] = parent; o->player[i] = (int8_t)player; o->label[i] = label; o->cprob[i] = cp;
    o->infoset[i] = -1; o->hidden[i] = 0; o->u1[i] = 0.0; o->u2[i] = 0.0;
    o->pend[i] = *p;
    return i;
}

/* open-addressing map: information-set key -> id in order of first appearance */
typedef struct { uint64_t *key; int32_t *val; size_t cap, used; } map_t;

static int map_get(map_t *m, uint64_t key) {
    if (m->used * 2 >= m->cap) {
        size_t ncap = m->cap ? m->cap * 2 : 1 << 12;
        uint64_t *nk = (uint64_t *)malloc(ncap * sizeof(uint64_t));
        int32_t *nv = (int32_t *)malloc(ncap * sizeof(int32_t));
        if (!nk || !nv) return -1;
        memset(nk, 0xff, ncap * sizeof(uint64_t));
        for (size_t i = 0; i < m->cap; ++i)
            if (m->key[i] != UINT64_MAX) {
                size_t h = (size_t)((m->key[i] * 0x9E3779B97F4A7C15ull) >> 20) & (ncap - 1);
                while (nk[h] != UINT64_MAX) h = (h + 1) & (ncap - 1);
                nk[h] = m->key[i]; nv[h] = m->val[i];
            }
        free(m->key); free(m->val);
        m->key = nk; m->val = nv; m->cap = ncap;
    }
    size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (m->cap - 1);
    while (m->key[h] != UINT64_MAX) {
        if (m->key[h] == key) return m->val[h];
        h = (h + 1) & (m->cap - 1);
    }
    m->key[h] = key; m->val[h] = (int32_t)m->used;
    return (int32_t)m->used++;
}

/* leduc.py:237-278 on card values (card index / suits) */
static void showdown(const pend_t *p, int suits, double *a, double *b) {
    int v1 = p->hole1 / suits, v2 = p->hole2 / suits, vb = p->board / suits;
    int pair1 = v1 == vb, pair2 = v2 == vb, winner;
    if (pair1 && pair2) winner = 0;
    else if (pair1) winner = 1;
    else if (pair2) winner = 2;
    else winner = v1 > v2 ? 1 : (v2 > v1 ? 2 : 0);
    if (winner == 1) { *a = (double)p->pot2; *b = -(double)p->pot2; }
    else if (winner == 2) { *a = -(double)p->pot1; *b = (double)p->pot1; }
    else { *a = 0.0; *b = 0.0; }
}

static out_t g_out;
static int32_t g_isets;

static void release(void) {
    free(g_out.parent); free(g_out.infoset); free(g_out.label); free(g_out.player); free(g_out.hidden);
    free(g_out.cprob); free(g_out.u1); free(g_out.u2); free(g_out.pend);
    memset(&g_out, 0, sizeof(g_out));
}

/* Build the tree; returns the node count (or -1) and keeps it for orc_leduc_fetch. */
int64_t orc_leduc_build(int values, int suits, int max_raises, int raise_amount, int32_t *num_infosets) {
    release();
    out_t *o = &g_out;
    const int C = values * suits;
    if (C < 3 || C > 250 || max_raises < 1) return -1;
    map_t m = {0, 0, 0, 0};
    pend_t root;
    memset(&root, 0, sizeof(root));
    root.state = ST_DEAL1; root.hole1 = root.hole2 = root.board = -1; root.pot1 = root.pot2 = 1; root.last = -1;
    if (push(o, -1, 0, 0, 1.0, &root) < 0) return -1;
    o->hidden[0] = 2;                                   /* leduc.py:69: hidden_from = [2] */
    for (int64_t i = 0; i < o->n; ++i) {
        pend_t p = o->pend[i];                           /* copy: the arrays may be reallocated below */
        if (p.state == ST_LEAF) continue;
        if (p.state == ST_DEAL1 || p.state == ST_DEAL2 || p.state == ST_BOARD) {
            int left = C - (p.state == ST_DEAL1 ? 0 : p.state == ST_DEAL2 ? 1 : 2);
            double prob = 1.0 / (double)left;            /* n / len(remaining_cards), every card distinct */
            for (int card = 0; card < C; ++card) {
                if (card == p.hole1 || card == p.hole2) continue;
                pend_t c = p;
                int64_t id;
                if (p.state == ST_DEAL1) {
                    c.state = ST_DEAL2; c.hole1 = (int16_t)card;
                    id = push(o, (int32_t)i, 0, CARD_LABEL + card, prob, &c);
                    if (id >= 0) o->hidden[id] = 1;     /* leduc.py:80: hidden_from = [1] */
                } else if (p.state == ST_DEAL2) {
                    c.state = ST_ROUND1; c.hole2 = (int16_t)card; c.actor = 1; c.last = -1; c.nbets = 0;
                    id = push(o, (int32_t)i, 1, CARD_LABEL + card, prob, &c);
                } else {
                    c.state = ST_ROUND2; c.board = (int16_t)card; c.actor = 2; c.last = -1; c.nbets = 0;
                    id = push(o, (int32_t)i, 2, CARD_LABEL + card, prob, &c);
                }
                if (id < 0) return -1;
            }
            continue;
        }
        /* a betting node: its information set is what the actor sees */
        {
            uint64_t own = (uint64_t)(p.actor == 1 ? p.hole1 : p.hole2);
            uint64_t board = p.state == ST_ROUND2 ? (uint64_t)p.board + 1 : 0;
            uint64_t key = (uint64_t)(p.actor - 1) | own << 1 | board << 9 | (uint64_t)p.seq1 << 17 | (uint64_t)p.seq2 << 40;
            int id = map_get(&m, key);
            if (id < 0) return -1;
            o->infoset[i] = id;
        }
        const int round2 = p.state == ST_ROUND2;
        const int stake = round2 ? 2 * raise_amount : raise_amount;
        const int other = 3 - p.actor;
        /* the three kinds of child: betting goes on / the round ends / fold */
        for (int k = 0; k < 3; ++k) {
            int action, pays, ends;
            if (p.last < 0) {                            /* leduc.py:113-131: call or bet, both continue */
                if (k == 2) break;
                action = k == 0 ? A_CALL : A_BET; pays = k == 1; ends = 0;
            } else if (p.last == A_CALL) {               /* leduc.py:132-157: call ends the round, bet continues */
                if (k == 2) break;
                action = k == 0 ? A_CALL : A_BET; pays = k == 1; ends = k == 0;
            } else {                                     /* leduc.py:158-214: call, fold, bet */
                action = k == 0 ? A_CALL : (k == 1 ? A_FOLD : A_BET);
                pays = action != A_FOLD;
                ends = action == A_CALL || (action == A_BET && !(p.nbets < max_raises - 1));
            }
            pend_t c = p;
            if (round2) c.seq2 = c.seq2 * 4 + (uint32_t)(action + 1); else c.seq1 = c.seq1 * 4 + (uint32_t)(action + 1);
            if (pays) { if (p.actor == 1) c.pot1 += stake; else c.pot2 += stake; }
            int64_t id;
            if (action == A_FOLD && p.last == A_BET) {   /* leduc.py:215-222 */
                c.state = ST_LEAF;
                id = push(o, (int32_t)i, -1, action, 1.0, &c);
                if (id < 0) return -1;
                if (p.actor == 1) { o->u1[id] = -(double)p.pot1; o->u2[id] = (double)p.pot1; }
                else { o->u1[id] = (double)p.pot2; o->u2[id] = -(double)p.pot2; }
            } else if (!ends) {
                c.actor = (int8_t)other; c.last = (int8_t)action; c.nbets = (int8_t)(p.nbets + (action == A_BET));
                id = push(o, (int32_t)i, other, action, 1.0, &c);
                if (id < 0) return -1;
            } else if (!round2) {
                c.state = ST_BOARD;
                id = push(o, (int32_t)i, 0, action, 1.0, &c);
                if (id < 0) return -1;
            } else {
                c.state = ST_LEAF;
                id = push(o, (int32_t)i, -1, action, 1.0, &c);
                if (id < 0) return -1;
                showdown(&c, suits, &o->u1[id], &o->u2[id]);
            }
        }
    }
    g_isets = (int32_t)m.used;
    free(m.key); free(m.val);
    if (num_infosets) *num_infosets = g_isets;
    return o->n;
}

/* Copy the arrays of the last orc_leduc_build out (first_child is derived: children are contiguous) and free them. */
int orc_leduc_fetch(int32_t *parent, int32_t *first_child, int8_t *player, int32_t *infoset, int32_t *label,
                    uint8_t *hidden, double *cprob, double *u1, double *u2) {
    out_t *o = &g_out;
    if (!o->n) return -1;
    size_t n = (size_t)o->n;
    memcpy(parent, o->parent, n * sizeof(int32_t));
    memcpy(player, o->player, n);
    memcpy(infoset, o->infoset, n * sizeof(int32_t));
    memcpy(label, o->label, n * sizeof(int32_t));
    memcpy(hidden, o->hidden, n);
    memcpy(cprob, o->cprob, n * sizeof(double));
    memcpy(u1, o->u1, n * sizeof(double));
    memcpy(u2, o->u2, n * sizeof(double));
    /* first_child[i] = 1 + number of nodes whose parent is < i */
    memset(first_child, 0, (n + 1) * sizeof(int32_t));
    for (size_t i = 1; i < n; ++i) first_child[o->parent[i] + 1] += 1;
    first_child[0] = 1;
    for (size_t i = 1; i <= n; ++i) first_child[i] += first_child[i - 1];
    release();
    return 0;
}
