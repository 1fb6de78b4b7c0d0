"""Node partition for the multi-GPU link update ("METIS-style cut").

METIS itself is not available, so this is the classic graph-growing recursive bisection METIS uses for its initial
partitions: start a breadth-first search at a peripheral node, cut the BFS order in two halves of (nearly) equal event
weight, recurse on both halves. On road networks the parts come out as contiguous regions with short boundaries, which is
what matters here — only nodes next to a cut link ever wait for another GPU. A node owns its turnings and the queues of
its IN-links (a link lives on the rank of its downstream node, mezzo_b200/csrc/sim_core.h)."""
import numpy as np


def _bfs_order(adj_off, adj, nodes, start):
    """BFS order of the induced subgraph on `nodes` (boolean mask), all components, starting at `start`."""
    n = len(adj_off) - 1
    seen = np.zeros(n, bool)
    order = []
    todo = [start] + [int(v) for v in np.nonzero(nodes)[0]]
    for s in todo:
        if seen[s] or not nodes[s]:
            continue
        seen[s] = True
        queue = [s]
        while queue:
            nxt = []
            for u in queue:
                order.append(u)
                for v in adj[adj_off[u]:adj_off[u + 1]]:
                    if nodes[v] and not seen[v]:
                        seen[v] = True
                        nxt.append(int(v))
            queue = nxt
    return order


def partition_nodes(n_nodes, link_in_node, link_out_node, world, weight=None):
    """node_rank[int32, n_nodes] in [0, world). `weight` (per node, default = 1 + degree) balances the expected
    number of events; `world` need not be a power of two."""
    world = int(world)
    rank = np.zeros(n_nodes, np.int32)
    if world <= 1:
        return rank
    a = np.concatenate([link_in_node, link_out_node]).astype(np.int64)
    b = np.concatenate([link_out_node, link_in_node]).astype(np.int64)
    keep = a != b
    a, b = a[keep], b[keep]
    o = np.argsort(a, kind="stable")
    a, b = a[o], b[o]
    adj_off = np.zeros(n_nodes + 1, np.int64)
    np.add.at(adj_off, a + 1, 1)
    adj_off = np.cumsum(adj_off)
    adj = b
    if weight is None:
        weight = 1.0 + np.diff(adj_off)
    weight = np.asarray(weight, np.float64)

    def split(mask, r0, parts):
        if parts == 1:
            rank[mask] = r0
            return
        ids = np.nonzero(mask)[0]
        # peripheral start: the last node of a BFS from an arbitrary node
        far = _bfs_order(adj_off, adj, mask, int(ids[0]))[-1]
        order = np.array(_bfs_order(adj_off, adj, mask, far), np.int64)
        left_parts = parts // 2
        target = weight[order].sum() * left_parts / parts
        cut = int(np.searchsorted(np.cumsum(weight[order]), target)) + 1
        cut = min(max(cut, left_parts), len(order) - (parts - left_parts))
        lm = np.zeros(n_nodes, bool)
        lm[order[:cut]] = True
        rm = mask & ~lm
        split(lm, r0, left_parts)
        split(rm, r0 + left_parts, parts - left_parts)

    if n_nodes < world:
        rank[:] = np.arange(n_nodes) % world
        return rank
    split(np.ones(n_nodes, bool), 0, world)
    return rank


def cut_links(node_rank, link_in_node, link_out_node):
    """Boolean mask of the links whose end nodes sit on different ranks (their state is reached through NVLink by the
    upstream node's rank)."""
    return node_rank[link_in_node] != node_rank[link_out_node]
