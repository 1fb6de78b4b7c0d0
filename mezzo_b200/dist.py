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
