// mezzo_b200/host/Graph.h — drop-in replacement of mezzo_lib's Graph<T,I> (B/Graph.h + B/Graph.cpp,
// B/ = /root/reference/branches/busmezzo/mezzo_lib/src/) whose label-correcting search runs on the GPU through the
// C ABI of include/mezzo_b200.h. Host code stays C++ in mezzo_lib; this header is the thin layer in between.
//
// How a mezzo_lib build picks it up: force-include this file (g++ -include .../host/Graph.h) or put
// its directory first and rename — it defines the SAME include guards as the reference's Graph.h/Graph.cpp
// (GRAPH_HEADER, GRAPH_IMPLEMENTATION), so `#include "Graph.h"` inside network.h becomes a no-op and every use of
// Graph<double, LinkTimeInfo> in network.cpp (B/network.cpp:6576-6791) binds to this class. Link with -lmezzo_b200.
//
// Same names, argument meaning and error behaviour as the reference for everything Network uses:
//   Graph(n, m, infinity)            B/Graph.cpp:15-36        addLink(i,u,d,w)                 B/Graph.cpp:47-63
//   set_downlink_indices()           B/Graph.cpp:689-706      set_turning_prohibitor(in,out)   B/Graph.cpp:709-717
//   labelCorrecting(root, t, info)   B/Graph.cpp:313-454      reachable(des)                   B/Graph.h:229-230
//   shortest_path_vector(des)        B/Graph.cpp:653-685      predecessorToLink/Node, costToNode, nNodes, nLinks
// plus a batch entry point (labelCorrectingBatch) for callers that can name their trees up front.
//
// Batching WITHOUT patching mezzo_lib. Network::shortest_paths_all (B/network.cpp:6667-6774) asks for one tree per call:
// every out-link of every origin, at every entry time, in ascending origin order. A GPU wants those trees in one batch. So a
// call whose tree is not at hand builds, in the same mz_sssp_batch, the trees of the NEXT source links too — links whose up
// node has no incoming link, which is what an Origin's out-links are — at the same entry time, and keeps them; the calls
// that follow are answered from that cache (each tree is handed out once, then dropped; a different entry time or info
// object, linkCost(i, w), addLink and set_turning_prohibitor drop all of it). The link-time table is read from the info
// object once per batch instead of once per tree. The one assumption: the table is not edited between two calls of one
// such run of calls at the same entry time — true for every caller in mezzo_lib; MEZZO_B200_PREFETCH=0 switches the
// prefetch off (one tree, one table read per call, like the reference).
//
// Not provided (unused by Network; see DESIGN.md "out of scope"): labelSetting, turning penalties (graph->penalty is
// commented out at B/network.cpp:6623), link grades, the print* debugging helpers.
#ifndef GRAPH_HEADER
#define GRAPH_HEADER
#define GRAPH_IMPLEMENTATION

#include <assert.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <vector>

#include "../../include/mezzo_b200.h"

enum { SP_UNSET = -2, SP_START = -1 };

// B/Graph.h:37-44
template <class T>
class GraphNoInfo {
  public:
    GraphNoInfo() {}
    ~GraphNoInfo() {}
    T cost(int, double) { return 1; }
};

namespace mezzo_b200_detail {
// Extracts the dense [link slot][period] table from the reference's LinkTimeInfo (map<int,LinkTime*> of
// map<int,double>, B/linktimes.h:34-66) without naming the type here: anything with the same members works.
template <class I>
auto fill_table(I *info, int n_links, std::vector<double> &T, int &P, double &periodlength, int)
    -> decltype(info->times.begin()->second->times, void())
{
    P = 1;
    periodlength = 1.0;
    for (auto &kv : info->times) {
        if (kv.second->nrperiods > P) P = kv.second->nrperiods;
        if (kv.second->periodlength > 0) periodlength = kv.second->periodlength;
        for (auto &tv : kv.second->times)
            if (tv.first + 1 > P) P = tv.first + 1;
    }
    T.assign((size_t)n_links * P, std::numeric_limits<double>::quiet_NaN());  // NaN = period absent ⇒ cost 0.1
    for (auto &kv : info->times) {
        if (kv.first < 0 || kv.first >= n_links) continue;
        for (auto &tv : kv.second->times)
            if (tv.first >= 0 && tv.first < P) T[(size_t)kv.first * P + tv.first] = tv.second;
    }
}
template <class I>
void fill_table(I *, int, std::vector<double> &, int &, double &, long)
{
}  // info types without a table (GraphNoInfo): the default link costs are used
}  // namespace mezzo_b200_detail

template <class T, class I>
class Graph {
    static_assert(sizeof(T) == sizeof(double) && (T)0.5 == 0.5, "Network instantiates Graph<double, ...> (B/network.h)");

  public:
    Graph(const int n = 2, const int m = 1, const T u = 0x7FFF / 3, const T p = 0)
        : n_nodes_(n), n_links_(m), infinity_(u), root_(SP_UNSET), handle_(NULL), table_info_(NULL), device_(0)
    {
        (void)p;
        up_.assign(m, SP_UNSET);
        dn_.assign(m, SP_UNSET);
        cost_.assign(m, 1);
        legal_.assign(m, (int8_t)0xFF);  // addLink default legal = (char)0xFF, B/Graph.h:203
        dn_index_.assign(m, 0);
        link_pred_.assign(m, SP_UNSET);
        node_pred_.assign(n, SP_UNSET);
        node_cost_.assign(n, 0);
        const char *dev = std::getenv("MEZZO_B200_DEVICE");
        if (dev) device_ = std::atoi(dev);
        const char *pf = std::getenv("MEZZO_B200_PREFETCH");
        prefetch_ = !(pf && std::atoi(pf) == 0);
        cache_info_ = NULL;
        cache_t_ = 0;
        n_batches_ = n_served_ = 0;
    }
    virtual ~Graph() { release(); }

    int nNodes() const { return n_nodes_; }
    int nLinks() const { return n_links_; }

    void addLink(int i, int u, int d, const T w = 1, char grade = 0, char legal = (char)0xFF, char index = 0)
    {
        (void)grade;
        (void)index;
        up_[i] = u;
        dn_[i] = d;
        cost_[i] = w;
        legal_[i] = (int8_t)legal;
        order_.push_back(i);
        release();
    }
    void linkCost(int i, const T w)
    {
        cost_[i] = w;
        table_info_ = NULL;
        cache_rows_.clear();
    }
    inline T &linkCost(int i) { return cost_[i]; }

    void set_downlink_indices()
    {
        std::map<int, int> cnt;
        for (size_t k = 0; k < order_.size(); ++k) dn_index_[order_[k]] = cnt[up_[order_[k]]]++;
    }
    void set_turning_prohibitor(int inlink, int outlink)
    {
        const int index = dn_index_[outlink];
        // char arithmetic like the reference: 1<<index with index >= 8 leaves the byte unchanged
        legal_[inlink] = (int8_t)(legal_[inlink] ^ (1 << index));
        assert((int8_t)(legal_[inlink] & (1 << index)) == 0);  // same assert as B/Graph.cpp:716
        release();
    }

    // one tree (the reference's call); info == NULL uses the default link costs like the reference does
    void labelCorrecting(int root, double t = 0, I *info = NULL)
    {
        root_ = root;
        if (prefetch_ && !cache_rows_.empty() && t == cache_t_ && info == cache_info_) {
            typename std::map<int, size_t>::iterator it = cache_rows_.find(root);
            if (it != cache_rows_.end()) {  // built together with an earlier call's tree
                take_row(it->second);
                cache_rows_.erase(it);
                ++n_served_;
                return;
            }
        } else {
            cache_rows_.clear();
        }
        // this tree plus, while the cache is empty, the trees of the source links that follow `root` in link order
        std::vector<int> roots(1, root);
        if (prefetch_ && cache_rows_.empty()) {
            if (sources_.empty()) find_sources();
            const size_t row_bytes = ((size_t)n_links_ + (size_t)n_nodes_) * sizeof(int) + (size_t)n_nodes_ * sizeof(T);
            const size_t kmax = std::max<size_t>(1, ((size_t)768 << 20) / std::max<size_t>(1, row_bytes));
            size_t start = std::lower_bound(sources_.begin(), sources_.end(), root) - sources_.begin();
            for (size_t k = 0; k < sources_.size() && roots.size() < kmax; ++k) {
                const int r = sources_[(start + k) % sources_.size()];
                if (r != root) roots.push_back(r);
            }
        }
        if (roots.size() == 1) {
            double e = t;
            batch(1, &roots[0], &e, info, &link_pred_[0], &node_pred_[0], &node_cost_[0]);
            return;
        }
        const size_t n = roots.size();
        std::vector<double> entries(n, t);
        c_link_pred_.resize(n * (size_t)n_links_);
        c_node_pred_.resize(n * (size_t)n_nodes_);
        c_node_cost_.resize(n * (size_t)n_nodes_);
        batch((int)n, &roots[0], &entries[0], info, &c_link_pred_[0], &c_node_pred_[0], &c_node_cost_[0]);
        ++n_batches_;
        cache_t_ = t;
        cache_info_ = info;
        for (size_t k = 1; k < n; ++k) cache_rows_[roots[k]] = k;
        take_row(0);
    }
    // diagnostics: batches built by the prefetch and trees answered from them
    long prefetchBatches() const { return n_batches_; }
    long prefetchServed() const { return n_served_; }
    // n trees at once; outputs are row-major [n][nLinks] / [n][nNodes] (any of them may be NULL)
    void labelCorrectingBatch(int n, const int *roots, const double *entries, I *info, int *link_pred, int *node_pred,
                              double *node_cost)
    {
        batch(n, roots, entries, info, link_pred, node_pred, node_cost);
    }

    inline int predecessorToLink(int i) { return link_pred_[i]; }
    inline int predecessorToNode(int i) { return node_pred_[i]; }
    inline T &nodeCost(int i) { return node_cost_[i]; }
    inline T &costToNode(int i) { return node_cost_[i]; }
    inline bool reachable(int des) { return node_pred_[des] != SP_UNSET; }

    std::vector<int> shortest_path_vector(int des)
    {
        assert(root_ != SP_UNSET);
        int p = node_pred_[des];
        assert(p != SP_UNSET);
        std::vector<int> linkids;
        std::vector<char> seen(n_links_, 0);
        while (p >= 0) {
            if (seen[p]) {
                std::cout << "circular path" << std::endl;
                linkids.clear();
                return linkids;
            }
            seen[p] = 1;
            linkids.insert(linkids.begin(), p);
            p = link_pred_[p];
        }
        return linkids;
    }

  private:
    void release()
    {
        if (handle_) mz_graph_destroy(handle_);
        handle_ = NULL;
        table_info_ = NULL;
        cache_rows_.clear();
        sources_.clear();
    }
    // links whose up node has no incoming link, ascending link id
    void find_sources()
    {
        std::vector<int> indeg(n_nodes_, 0);
        for (size_t k = 0; k < order_.size(); ++k)
            if (dn_[order_[k]] >= 0 && dn_[order_[k]] < n_nodes_) indeg[dn_[order_[k]]]++;
        for (size_t k = 0; k < order_.size(); ++k)
            if (up_[order_[k]] >= 0 && up_[order_[k]] < n_nodes_ && indeg[up_[order_[k]]] == 0) sources_.push_back(order_[k]);
        std::sort(sources_.begin(), sources_.end());
    }
    void take_row(size_t k)
    {
        std::copy(c_link_pred_.begin() + k * n_links_, c_link_pred_.begin() + (k + 1) * n_links_, link_pred_.begin());
        std::copy(c_node_pred_.begin() + k * n_nodes_, c_node_pred_.begin() + (k + 1) * n_nodes_, node_pred_.begin());
        std::copy(c_node_cost_.begin() + k * n_nodes_, c_node_cost_.begin() + (k + 1) * n_nodes_, node_cost_.begin());
    }
    static void die(const char *what, int rc)
    {
        // the reference asserts/crashes on internal errors; a failed GPU call must not be silently ignored either
        std::fprintf(stderr, "mezzo_b200: %s failed (%d): %s\n", what, rc, mz_last_error());
        std::abort();
    }
    void batch(int n, const int *roots, const double *entries, I *info, int *lp, int *np, double *nc)
    {
        if (!handle_) {
            int rc = mz_graph_create(device_, n_nodes_, n_links_, (int)order_.size(), order_.empty() ? NULL : &order_[0],
                                     &up_[0], &dn_[0], &legal_[0], &handle_);
            if (rc != MZ_OK) die("mz_graph_create", rc);
        }
        // The link-time table may have changed between calls (Network updates histtimes between iterations), so it is
        // re-read on every call; use labelCorrectingBatch to amortise this over many trees.
        std::vector<double> table;
        int P = 0;
        double pl = 1.0;
        if (info) mezzo_b200_detail::fill_table(info, n_links_, table, P, pl, 0);
        if (table.empty()) {  // no info object or no table inside it: static default costs (B/Graph.cpp:361, 401)
            P = 1;
            pl = 1e300;
            table.resize(n_links_);
            for (int i = 0; i < n_links_; ++i) table[i] = cost_[i];
        }
        int rc = mz_linktimes_set(handle_, P, pl, &table[0]);
        if (rc != MZ_OK) die("mz_linktimes_set", rc);
        rc = mz_sssp_batch(handle_, n, roots, entries, MZ_SSSP_AUTO, lp, np, nc);
        if (rc != MZ_OK) die("mz_sssp_batch", rc);
    }

    int n_nodes_, n_links_;
    T infinity_;
    int root_;
    std::vector<int> up_, dn_, order_, dn_index_;
    std::vector<T> cost_;
    std::vector<int8_t> legal_;
    std::vector<int> link_pred_, node_pred_;
    std::vector<T> node_cost_;
    mz_graph *handle_;
    I *table_info_;
    int device_;
    // prefetched trees (see the header comment)
    bool prefetch_;
    std::vector<int> sources_;
    std::map<int, size_t> cache_rows_;  // root link -> row of the c_* arrays, trees not handed out yet
    double cache_t_;
    I *cache_info_;
    std::vector<int> c_link_pred_, c_node_pred_;
    std::vector<T> c_node_cost_;
    long n_batches_, n_served_;
};

#endif  // GRAPH_HEADER
