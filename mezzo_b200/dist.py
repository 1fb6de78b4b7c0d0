"""Multi-GPU plumbing for hot path 2: the unit of work is one (root link, entry time) tree, trees are
independent, so the tree list is cut into contiguous blocks — one per rank — with NO collective during compute; the graph
and the link-time table are replicated. Afterwards the predecessor slices / route sets are concatenated in rank order,
which is the reference's loop order (reruns → origins → root links), so route ids and exists_same_route de-duplication
(B/network.cpp:6743-6762) come out as in a single-process run. torch.distributed is only the transport (nccl on GPUs,
gloo in the CPU tests)."""
import numpy as np


def shard(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one item and cover [0, n_items) in rank order."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local, group=None):
    """Concatenates per-rank row blocks (numpy, first dimension may differ per rank) in rank order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    t = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def sharded_trees(compute, roots, entries, rank, world, gather=True, group=None):
    """Runs `compute(roots_block, entries_block) -> (link_pred, node_pred, node_cost)` on this rank's block and (optionally)
    reassembles the full arrays on every rank. `compute` is Graph.sssp_batch on a GPU rank."""
    lo, hi = shard(len(roots), rank, world)
    lp, npd, nc = compute(np.ascontiguousarray(roots[lo:hi]), np.ascontiguousarray(entries[lo:hi]))
    if not gather or world == 1:
        return lp, npd, nc
    return all_gather_rows(lp, group), all_gather_rows(npd, group), all_gather_rows(nc, group)


# ---------------------------------------------------------------------------------------------------------------------
# hot path 1: the link update partitioned over ranks
# ---------------------------------------------------------------------------------------------------------------------
def _allreduce_np(a, op, group=None):
    """All-reduce of a numpy array through torch.distributed (nccl: staged through the current CUDA device)."""
    import torch
    import torch.distributed as dist
    if dist.get_world_size(group) == 1:
        return a
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dist.all_reduce(t, op=op, group=group)
    return t.cpu().numpy()


class PartitionedSimulation:
    """Network::step over several GPUs: the nodes are partitioned (mezzo_b200/partition.py), every rank runs the
    persistent node-visit kernel on its own nodes, and neighbouring nodes on different ranks synchronise pairwise through
    the published keys and link records in each other's arenas (NVLink peer memory; mz_sim_arena_handle /
    mz_sim_attach_peer). There is no collective on the data path: torch.distributed only exchanges the 64-byte IPC
    handles at start-up, separates reset / run / read-out with barriers, picks the globally first event beyond the
    horizon and merges the per-rank outputs.

    `backend` is the per-rank engine: mezzo_b200.sim.CudaPartBackend on GPUs; the CPU test-suite injects the host
    emulation (tests/orc.py) to exercise exactly this orchestration under gloo."""

    def __init__(self, flat, rank, world, backend, group=None, node_rank=None):
        import torch.distributed as dist
        from . import partition
        self.flat, self.rank, self.world, self.be, self.group = flat, int(rank), int(world), backend, group
        st = flat.struct
        if node_rank is None:
            node_rank = partition.partition_nodes(st.n_nodes, flat.link_in_node, flat.link_out_node, world)
        self.node_rank = np.ascontiguousarray(node_rank, np.int32)
        self.link_home = self.node_rank[flat.link_out_node]
        self.be.create(flat, self.world, self.rank, self.node_rank)
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.be.export_handle(), group=group)
            for r in range(self.world):
                if r != self.rank:
                    self.be.attach(r, handles[r])
            dist.barrier(group=group)
        self._stats = None

    def reset(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody is still reading this rank's arena
        self.be.reset()

    def step(self):
        """Runs the horizon on all ranks; returns Network::time after the loop (the time of the last executed event)."""
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(group=self.group)  # every arena is (re)initialised before any rank starts polling it
        err = None
        try:
            self.be.run()
        except Exception as e:  # (a failed rank must not leave the others waiting in the barrier below)
            if self.world == 1:
                raise
            err = e
        if self.world == 1:
            self._stats = self.be.stats()
            return self._stats["end_time"]
        failed = _allreduce_np(np.array([1 if err is not None else 0], np.int64), dist.ReduceOp.SUM, self.group)
        if int(failed[0]):
            raise RuntimeError(f"mz_sim_run failed on {int(failed[0])} of {self.world} ranks"
                               + (f" (this rank: {err})" if err is not None else " (another rank)"))
        keys = [None] * self.world
        dist.all_gather_object(keys, self.be.min_key(), group=self.group)
        cand = [k for k in keys if k[3] >= 0 and np.isfinite(k[0])]
        node, end_time = -1, 0.0
        if cand:
            t, tp, _, node = min(cand)  # (t, t_parent, sub, node): the reference's event order
            end_time = t
        # a MatrixAction / the poll of an inactive ODaction may be the event that ends the loop (MzNet::phantom_final_t)
        pt, ptp = getattr(self.flat, "phantom_final", (0.0, 0.0))
        if pt > 0.0 and (not cand or (pt, ptp) < (t, tp)):
            node, end_time = -1, pt
        self.be.final_event(node, end_time)
        dist.barrier(group=self.group)
        st = self.be.stats()
        sums = _allreduce_np(np.array([st[k] for k in ("n_traversals", "n_exits", "n_events", "n_arrived", "n_supersteps",
                                                         "n_launches")], np.int64), dist.ReduceOp.SUM, self.group)
        mx = _allreduce_np(np.array([st["kernel_ms"], st["total_ms"]], np.float64), dist.ReduceOp.MAX, self.group)
        self._stats = dict(n_traversals=int(sums[0]), n_exits=int(sums[1]), n_events=int(sums[2]), n_arrived=int(sums[3]),
                           n_supersteps=int(sums[4]), n_launches=int(sums[5]), kernel_ms=float(mx[0]),
                           total_ms=float(mx[1]), end_time=end_time)
        return end_time

    def stats(self):
        return self._stats

    def arrivals(self):
        """arrival[n_trips]: a vehicle is reported by the rank that owns its destination node (-1 elsewhere)."""
        import torch.distributed as dist
        a = self.be.arrivals()
        return a if self.world == 1 else _allreduce_np(a, dist.ReduceOp.MAX, self.group)

    def linktimes(self, final=False):
        """Link::avgtimes: a link's period averages are kept by the rank of its downstream node. final=True applies
        Link::end_of_simulation (what linktimes.dat holds; mz_sim_get_linktimes_final reports zeros for foreign links)."""
        import torch.distributed as dist
        lt = self.be.linktimes(final) if final else self.be.linktimes()
        if self.world == 1:
            return lt
        lt = np.where((self.link_home == self.rank)[:, None], lt, 0.0)
        return _allreduce_np(lt, dist.ReduceOp.SUM, self.group)

    def exit_log(self):
        """All ranks' exit records merged and ordered by (vehicle, entry time) = position on the route."""
        import torch.distributed as dist
        ex = self.be.exit_log()
        if self.world > 1:
            parts = [None] * self.world
            dist.all_gather_object(parts, ex, group=self.group)
            ex = np.concatenate(parts)
        return np.sort(ex, order=["vid", "entry"])

    def close(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(group=self.group)  # peers may still be reading this rank's arena
        self.be.destroy()
