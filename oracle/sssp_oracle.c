/*
 * TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's route-generation algorithm.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product path (mezzo_b200/) never links or calls it.
 *
 * Parity pinning: the reference's own tests pin nothing for this path (calc_paths=0 in the
 * only fixture), so this restatement is pinned against the reference ITSELF: oracle/_ref/libref_sssp.so
 * (the unmodified B/Graph.h + B/Graph.cpp + B/linktimes.cpp compiled in place) on thousands of random
 * graphs (tests/test_oracle_sssp.py, when /root/reference is present) and against committed golden
 * vectors produced by that library (tests/golden/sssp_*.npz, generator tests/golden/make_sssp_golden.py).
 *
 * B/ = /root/reference/branches/busmezzo/mezzo_lib/src/
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SP_UNSET (-2) /* B/Graph.h:27 */
#define SP_START (-1)
#define ORC_INF 9999999.0 /* B/network.cpp:6598 */

typedef struct {
    int32_t n_nodes, n_links;
    int32_t *up, *dn;     /* GraphLink::upNode_/dnNode_ (SP_UNSET if never added)        B/Graph.h:96-97 */
    int8_t *legal;        /* GraphLink::dnLegal_ (char)                                   B/Graph.h:102 */
    int8_t *dnidx;        /* GraphLink::dnIndex_ (char), set_downlink_indices            B/Graph.cpp:689-706 */
    int32_t *dn_off, *dn_list; /* GraphNode::dnLinks_ in addLink order                   B/Graph.cpp:61 */
    int32_t *up_off, *up_list; /* GraphNode::upLinks_ in addLink order                   B/Graph.cpp:62 */
    int32_t P;
    double periodlength;
    const double *T;      /* borrowed: [n_links][P], NaN = period absent */
} orc_graph;

static void orc_free(orc_graph *g)
{
    if (!g) return;
    free(g->up); free(g->dn); free(g->legal); free(g->dnidx);
    free(g->dn_off); free(g->dn_list); free(g->up_off); free(g->up_list);
    free(g);
}

void orc_graph_destroy(orc_graph *g) { orc_free(g); }

/* Graph ctor + addLink in add_order + set_downlink_indices   (B/Graph.cpp:15-63, 689-706) */
orc_graph *orc_graph_create(int32_t n_node_slots, int32_t n_link_slots, int32_t n_added,
                            const int32_t *add_order, const int32_t *up_node, const int32_t *dn_node,
                            const int8_t *dn_legal)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(*g));
    int32_t i, k;
    g->n_nodes = n_node_slots;
    g->n_links = n_link_slots;
    g->up = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_link_slots + 1));
    g->dn = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_link_slots + 1));
    g->legal = (int8_t *)malloc((size_t)n_link_slots + 1);
    g->dnidx = (int8_t *)calloc((size_t)n_link_slots + 1, 1);
    g->dn_off = (int32_t *)calloc((size_t)n_node_slots + 2, sizeof(int32_t));
    g->up_off = (int32_t *)calloc((size_t)n_node_slots + 2, sizeof(int32_t));
    for (i = 0; i < n_link_slots; ++i) {
        g->up[i] = up_node[i];
        g->dn[i] = dn_node[i];
        g->legal[i] = dn_legal ? dn_legal[i] : (int8_t)0xFF; /* addLink default legal = (char)0xFF, B/Graph.h:203 */
    }
    /* count, then fill in add order so each node's list keeps addLink order */
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_link_slots + 1));
    if (add_order) {
        memcpy(order, add_order, sizeof(int32_t) * (size_t)n_added);
    } else { /* Network iterates linkmap = ascending link id, B/network.cpp:6602 */
        n_added = 0;
        for (i = 0; i < n_link_slots; ++i)
            if (up_node[i] != SP_UNSET) order[n_added++] = i;
    }
    for (k = 0; k < n_added; ++k) {
        g->dn_off[g->up[order[k]] + 1]++;
        g->up_off[g->dn[order[k]] + 1]++;
    }
    for (i = 0; i < n_node_slots; ++i) {
        g->dn_off[i + 1] += g->dn_off[i];
        g->up_off[i + 1] += g->up_off[i];
    }
    g->dn_list = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_added + 1));
    g->up_list = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_added + 1));
    int32_t *dc = (int32_t *)calloc((size_t)n_node_slots + 1, sizeof(int32_t));
    int32_t *uc = (int32_t *)calloc((size_t)n_node_slots + 1, sizeof(int32_t));
    for (k = 0; k < n_added; ++k) {
        int32_t l = order[k], u = g->up[l], d = g->dn[l];
        g->dnidx[l] = (int8_t)dc[u]; /* dnIndex_ = position among the up-node's dnLinks_, stored in a char */
        g->dn_list[g->dn_off[u] + dc[u]++] = l;
        g->up_list[g->up_off[d] + uc[d]++] = l;
    }
    free(dc); free(uc); free(order);
    return g;
}

void orc_linktimes_set(orc_graph *g, int32_t P, double periodlength, const double *T)
{
    g->P = P;
    g->periodlength = periodlength;
    g->T = T;
}

/* LinkTimeInfo::cost → LinkTime::cost   (B/linktimes.cpp:108-121, 36-82) */
static inline double orc_cost(const orc_graph *g, int32_t l, double t)
{
    /* unsigned int i = static_cast<int>(time/periodlength); times.count(i) ? times[i] : 0.1 */
    int32_t i = (int32_t)(t / g->periodlength);
    if (i < 0 || i >= g->P) return 0.1;
    double v = g->T[(size_t)l * (size_t)g->P + (size_t)i];
    return isnan(v) ? 0.1 : v;
}

/* GraphLink::dnLegal(short i): return type char truncates (dnLegal_ & (1<<i)) to 8 bits  (B/Graph.h:127-129) */
static inline int orc_dn_legal(const orc_graph *g, int32_t u, int k)
{
    int m = (int)g->legal[u] & (1 << k);
    return (int8_t)m != 0;
}

/*
 * Graph::labelCorrecting(s, entry, info) followed by calcCostToNodes   (B/Graph.cpp:313-454, 458-483).
 * Returns the number of cost() evaluations (the reference-counted relaxations incl. the root label).
 * canonical (optional) receives Σ_{reached u} |legal dnLinks(dnNode(u)) with dnNode(v) != rootnode|.
 */
int64_t orc_label_correcting(const orc_graph *g, int32_t s, double entry, int32_t *link_pred,
                             int32_t *node_pred, double *node_cost, int64_t *canonical)
{
    const int32_t nlinks = g->n_links;
    double *label = (double *)malloc(sizeof(double) * (size_t)nlinks);
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)nlinks);
    int32_t *pred = link_pred ? link_pred : (int32_t *)malloc(sizeof(int32_t) * (size_t)nlinks);
    int32_t i, u, v, front, rear;
    int64_t ncost = 0;
    const int32_t rootnode = g->up[s];

    for (i = 0; i < nlinks; ++i) { /* B/Graph.cpp:348-352 */
        pred[i] = SP_UNSET;
        label[i] = ORC_INF;
        queue[i] = SP_UNSET;
    }
    pred[s] = SP_START; /* B/Graph.cpp:356-363 */
    label[s] = orc_cost(g, s, entry);
    ++ncost;
    queue[s] = nlinks;
    u = rear = s;

    while (u != nlinks) { /* B/Graph.cpp:367-436 */
        front = queue[u];
        queue[u] = SP_START;
        const int32_t pivot = g->dn[u];
        const int32_t b = g->dn_off[pivot], e = g->dn_off[pivot + 1];
        for (i = b; i < e; ++i) {
            v = g->dn_list[i];
            if (g->dn[v] == rootnode) continue;            /* B/Graph.cpp:385 */
            if (!orc_dn_legal(g, u, g->dnidx[v])) continue; /* B/Graph.cpp:391-396 */
            double newlabel = label[u] + orc_cost(g, v, entry + label[u]); /* B/Graph.cpp:399 */
            ++ncost;
            /* penalties_ are never set by Network (graph->penalty call is commented out, B/network.cpp:6623);
               grade_ is 0 on every link (addLink default), so neither term applies. B/Graph.cpp:404-410 */
            if (newlabel < label[v]) { /* strict '<', B/Graph.cpp:412 */
                pred[v] = u;
                label[v] = newlabel;
                if (queue[v] == SP_UNSET) { /* never in list: append, B/Graph.cpp:418-425 */
                    queue[rear] = v; /* executed even if rear was already dequeued */
                    queue[v] = nlinks;
                    rear = v;
                    if (front == nlinks) front = v;
                } else if (queue[v] == SP_START) { /* was in list: push FRONT, B/Graph.cpp:427-430 */
                    queue[v] = front;
                    front = v;
                }
            }
        }
        u = front;
    }

    if (canonical) {
        int64_t c = 0;
        for (u = 0; u < nlinks; ++u) {
            if (pred[u] == SP_UNSET) continue;
            const int32_t pivot = g->dn[u];
            for (i = g->dn_off[pivot]; i < g->dn_off[pivot + 1]; ++i) {
                v = g->dn_list[i];
                if (g->dn[v] == rootnode) continue;
                if (!orc_dn_legal(g, u, g->dnidx[v])) continue;
                ++c;
            }
        }
        *canonical = c;
    }

    /* calcCostToNodes: first minimum over upLinks_ in stored order   (B/Graph.cpp:458-483) */
    if (node_pred || node_cost) {
        int32_t n;
        for (n = 0; n < g->n_nodes; ++n) {
            double best = ORC_INF;
            int32_t p = SP_UNSET;
            for (i = g->up_off[n]; i < g->up_off[n + 1]; ++i) {
                u = g->up_list[i];
                if (label[u] < best) {
                    best = label[u];
                    p = u;
                }
            }
            if (node_pred) node_pred[n] = p;
            if (node_cost) node_cost[n] = best;
        }
    }
    free(label);
    free(queue);
    if (!link_pred) free(pred);
    return ncost;
}

/*
 * reachable(dest) + get_path(dest) + "prepend root link if missing"
 *   (B/Graph.h:229-230, B/Graph.cpp:653-685, B/network.cpp:6726-6742).
 * Returns path length (0 = unreachable or circular), -1 if cap too small.
 * The O(len^2) duplicate search of the reference only detects cycles; a visited-stamp array is equivalent.
 */
int32_t orc_route(const orc_graph *g, int32_t root, const int32_t *link_pred, const int32_t *node_pred,
                  int32_t dest, int32_t *out, int32_t cap)
{
    int32_t p = node_pred[dest];
    int32_t len = 0, i;
    if (p == SP_UNSET) return 0; /* !reachable */
    uint8_t *seen = (uint8_t *)calloc((size_t)g->n_links, 1);
    while (p >= 0) {
        if (seen[p]) { free(seen); return 0; } /* "circular path": linkids.clear() */
        seen[p] = 1;
        if (len >= cap) { free(seen); return -1; }
        out[len++] = p; /* reversed; fixed below */
        p = link_pred[p];
    }
    free(seen);
    if (len == 0) return 0; /* rlinks.size()==0 → no route */
    if (out[len - 1] != root) { /* frontid != root → insert root at begin */
        if (len >= cap) return -1;
        out[len++] = root;
    }
    for (i = 0; i < len / 2; ++i) {
        int32_t t = out[i];
        out[i] = out[len - 1 - i];
        out[len - 1 - i] = t;
    }
    return len;
}

/* batch helper used by bench.py's cpu_baseline: n trees, one after the other, single thread */
int64_t orc_sssp_batch(const orc_graph *g, int32_t n, const int32_t *root, const double *entry,
                       int32_t *link_pred, int32_t *node_pred, double *node_cost, int64_t *canonical_total)
{
    int64_t total = 0, ctot = 0;
    int32_t t;
    for (t = 0; t < n; ++t) {
        int64_t c = 0;
        total += orc_label_correcting(g, root[t], entry[t],
                                      link_pred ? link_pred + (size_t)t * g->n_links : NULL,
                                      node_pred ? node_pred + (size_t)t * g->n_nodes : NULL,
                                      node_cost ? node_cost + (size_t)t * g->n_nodes : NULL, &c);
        ctot += c;
    }
    if (canonical_total) *canonical_total = ctot;
    return total;
}
