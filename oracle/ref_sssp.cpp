// TEST INFRASTRUCTURE — not part of the product path.
// Thin C wrapper that instantiates the UNMODIFIED reference Graph<double,LinkTimeInfo>
// (B/Graph.h, B/Graph.cpp, B/linktimes.{h,cpp}; B/ = /root/reference/branches/busmezzo/mezzo_lib/src)
// from the sources where they lie (-I on the reference tree, see oracle/Makefile) and exposes it
// so the C restatement (sssp_oracle.c) and the CUDA path can be checked against the real thing.
// Output: oracle/_ref/libref_sssp.so (git-ignored, travels to the GPU box).
#include "Graph.h"
#include "linktimes.h"
#include <cmath>
#include <cstdint>

// With -DMZ_DROPIN (and mezzo_b200/host/Graph.h force-included, which shadows the reference's Graph.h through its include
// guards) the very same wrapper drives the GPU engine through the drop-in class: oracle/_ref/libdropin_sssp.so.
namespace {
#ifdef MZ_DROPIN
struct CountingInfo : LinkTimeInfo {
    LinkTimeInfo *inner;
    long long calls;
};
typedef Graph<double, LinkTimeInfo> RefGraph;
#else
// counts info->cost() calls = "reference-counted relaxations" (+1 for the root label)
struct CountingInfo {
    LinkTimeInfo *inner;
    long long calls;
    double cost(int i, double t) { ++calls; return inner->cost(i, t); }
};
typedef Graph<double, CountingInfo> RefGraph;
#endif
struct Ref {
    RefGraph *g;
    LinkTimeInfo info;
    CountingInfo counting;
    int n_nodes, n_links;
};
}

extern "C" {

void *ref_graph_create(int n_node_slots, int n_link_slots)
{
    Ref *r = new Ref;
    r->g = new RefGraph(n_node_slots, n_link_slots, 9999999.0); // B/network.cpp:6598
    r->counting.inner = &r->info;
    r->counting.calls = 0;
    r->n_nodes = n_node_slots;
    r->n_links = n_link_slots;
    return r;
}

void ref_graph_add_link(void *h, int id, int up, int dn, double cost)
{
    ((Ref *)h)->g->addLink(id, up, dn, cost); // B/network.cpp:6613
}

void ref_graph_set_downlink_indices(void *h) { ((Ref *)h)->g->set_downlink_indices(); }

void ref_graph_set_turning_prohibitor(void *h, int in, int out)
{
    ((Ref *)h)->g->set_turning_prohibitor(in, out); // B/network.cpp:6624
}

// raw dnLegal_ byte of a link after prohibitors (needs -fno-access-control)
#ifdef MZ_DROPIN
int ref_graph_dn_legal(void *h, int link) { return (int)((Ref *)h)->g->legal_[link]; }
#else
int ref_graph_dn_legal(void *h, int link) { return (int)(signed char)((Ref *)h)->g->link(link)->dnLegal_; }
#endif

// link-time table: row-major [n_link_slots][P]; NaN entries are left out of LinkTime::times;
// has_row[l]==0 ⇒ link l has no LinkTime at all in LinkTimeInfo::times.
void ref_linktimes_set(void *h, int P, double periodlength, const double *T, const signed char *has_row)
{
    Ref *r = (Ref *)h;
    for (auto &kv : r->info.times) delete kv.second;
    r->info.times.clear();
    for (int l = 0; l < r->n_links; ++l) {
        if (has_row && !has_row[l]) continue;
        LinkTime *lt = new LinkTime(l, P, periodlength);
        for (int p = 0; p < P; ++p) {
            double v = T[(size_t)l * P + p];
            if (!std::isnan(v)) lt->times[p] = v;
        }
        r->info.times[l] = lt;
    }
}

long long ref_label_correcting(void *h, int root, double entry, int32_t *link_pred, int32_t *node_pred,
                               double *node_cost)
{
    Ref *r = (Ref *)h;
    r->counting.calls = 0;
#ifdef MZ_DROPIN
    r->g->labelCorrecting(root, entry, &r->info);
#else
    r->g->labelCorrecting(root, entry, &r->counting); // B/network.cpp:6714
#endif
    if (link_pred)
        for (int i = 0; i < r->n_links; ++i) link_pred[i] = r->g->predecessorToLink(i);
    if (node_pred)
        for (int i = 0; i < r->n_nodes; ++i) node_pred[i] = r->g->predecessorToNode(i);
    if (node_cost)
        for (int i = 0; i < r->n_nodes; ++i) node_cost[i] = r->g->costToNode(i);
    return r->counting.calls;
}

int ref_reachable(void *h, int dest) { return ((Ref *)h)->g->reachable(dest) ? 1 : 0; }

// Graph::shortest_path_vector for the tree of the last ref_label_correcting call; returns length
// (0 for a circular path), -1 if cap is too small. Caller must check ref_reachable first (the reference asserts).
int ref_shortest_path_vector(void *h, int dest, int32_t *out, int cap)
{
    std::vector<int> v = ((Ref *)h)->g->shortest_path_vector(dest);
    if ((int)v.size() > cap) return -1;
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}

void ref_graph_destroy(void *h)
{
    Ref *r = (Ref *)h;
    for (auto &kv : r->info.times) delete kv.second;
    delete r->g;
    delete r;
}
}
