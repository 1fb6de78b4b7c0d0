"""Host-side mirror of the reference's route-generation interface on top of the C ABI.

Names and argument meaning follow Graph<double,LinkTimeInfo> (B/Graph.h:137-235) and LinkTimeInfo
(B/linktimes.h:56-66), B/ = /root/reference/branches/busmezzo/mezzo_lib/src/. The C++ twin that drops into
mezzo_lib is mezzo_b200/host/Graph.h. All computation happens in libmezzo_b200.so on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MZ_SSSP_AUTO, MZ_SSSP_EXACT, SP_UNSET, check, lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor (host or device)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class LinkTimeInfo:
    """Dense stand-in for LinkTimeInfo/LinkTime (map<int,LinkTime*> of map<int,double>, B/linktimes.h:34-66):
    times[l, p] = travel time of link l in period p, NaN where the period is absent (cost 0.1)."""

    def __init__(self, n_link_slots, nrperiods, periodlength):
        self.nrperiods = int(nrperiods)
        self.periodlength = float(periodlength)
        self.times = np.full((int(n_link_slots), self.nrperiods), np.nan, np.float64)

    def set_row(self, link_id, values):
        self.times[link_id, :len(values)] = values


class Graph:
    """Graph(n, m, infinity) / addLink / set_downlink_indices / set_turning_prohibitor / labelCorrecting /
    reachable / shortest_path_vector — same call sequence as Network::init_shortest_path (B/network.cpp:6576-6630)
    and Network::shortest_paths_all (B/network.cpp:6667-6774)."""

    def __init__(self, n_nodes, n_links, device=0):
        self.n_nodes, self.n_links, self.device = int(n_nodes), int(n_links), int(device)
        self.up = np.full(self.n_links, SP_UNSET, np.int32)
        self.dn = np.full(self.n_links, SP_UNSET, np.int32)
        self.legal = np.full(self.n_links, 0xFF, np.uint8)
        self._order = []
        self._dn_index = None
        self._h = None
        self._info = None
        self._root = None
        self._link_pred = self._node_pred = self._node_cost = None

    # --- construction (B/Graph.cpp:47-63, 689-717) ---
    def addLink(self, i, u, d, w=1.0):
        if self._h is not None:
            raise _lib.MzError(_lib.MZ_ESTATE, "graph already finalised")
        if not (0 <= i < self.n_links and 0 <= u < self.n_nodes and 0 <= d < self.n_nodes):
            raise _lib.MzError(_lib.MZ_EINVAL, f"addLink({i},{u},{d}) outside the graph's slots")
        self.up[i], self.dn[i] = u, d
        self._order.append(i)

    def set_downlink_indices(self):
        cnt = {}
        self._dn_index = np.zeros(self.n_links, np.int64)
        for l in self._order:
            u = int(self.up[l])
            self._dn_index[l] = cnt.get(u, 0)
            cnt[u] = self._dn_index[l] + 1

    def set_turning_prohibitor(self, inlink, outlink):
        if self._dn_index is None:
            raise _lib.MzError(_lib.MZ_ESTATE, "set_downlink_indices() must be called first (B/network.cpp:6617)")
        k = int(self._dn_index[outlink])
        new = (int(self.legal[inlink]) ^ (1 << k)) & 0xFF  # char arithmetic: index >= 8 leaves the byte unchanged
        if k < 8 and (new >> k) & 1:
            # the reference asserts dnLegal(index)==0 here (B/Graph.cpp:716): a repeated prohibitor aborts the program
            raise _lib.MzError(_lib.MZ_EINVAL, f"turning prohibitor ({inlink},{outlink}) given twice")
        self.legal[inlink] = new

    def finalize(self):
        if self._h is None:
            if self._dn_index is None:
                self.set_downlink_indices()
            h = C.c_void_p()
            order = np.ascontiguousarray(self._order, np.int32)
            check(lib().mz_graph_create(self.device, self.n_nodes, self.n_links, len(order), _ptr(order),
                                        _ptr(self.up), _ptr(self.dn), _ptr(self.legal.view(np.int8)), C.byref(h)))
            self._h = h
        return self

    @classmethod
    def from_arrays(cls, n_nodes, n_links, up, dn, prohibitors=(), device=0):
        """Build the way Network::init_shortest_path does: links in ascending id order, then prohibitors."""
        g = cls(n_nodes, n_links, device)
        up = np.asarray(up)
        dn = np.asarray(dn)
        ids = np.nonzero(up != SP_UNSET)[0]
        g.up[ids] = up[ids]
        g.dn[ids] = dn[ids]
        g._order = ids.tolist()
        g.set_downlink_indices()
        for a, b in prohibitors:
            g.set_turning_prohibitor(a, b)
        return g.finalize()

    # --- link times ---
    def set_linktimes(self, info_or_T, nrperiods=None, periodlength=None):
        self.finalize()
        if isinstance(info_or_T, LinkTimeInfo):
            T, P, pl = info_or_T.times, info_or_T.nrperiods, info_or_T.periodlength
        else:
            T, P, pl = info_or_T, int(nrperiods), float(periodlength)
        if isinstance(T, np.ndarray):
            T = np.ascontiguousarray(T, np.float64)
            assert T.size == self.n_links * P
        check(lib().mz_linktimes_set(self._h, P, pl, _ptr(T)))
        self._info = (P, pl)

    def set_option(self, option, value):
        self.finalize()
        check(lib().mz_graph_set_option(self._h, option, int(value)))

    # --- batch API (what the replaced loop of shortest_paths_all calls) ---
    def sssp_batch(self, roots, entries, mode=MZ_SSSP_EXACT, want_node_cost=True, out=None):
        """n trees → (link_pred [n,L] int32, node_pred [n,N] int32, node_cost [n,N] f64). `out` may hold
        preallocated numpy arrays or torch tensors (host-pinned or device) for the three outputs."""
        self.finalize()
        roots = np.ascontiguousarray(roots, np.int32)
        entries = np.ascontiguousarray(entries, np.float64)
        n = len(roots)
        if out is None:
            lp = np.empty((n, self.n_links), np.int32)
            npd = np.empty((n, self.n_nodes), np.int32)
            nc = np.empty((n, self.n_nodes), np.float64) if want_node_cost else None
        else:
            lp, npd, nc = out
        check(lib().mz_sssp_batch(self._h, n, _ptr(roots), _ptr(entries), mode, _ptr(lp), _ptr(npd), _ptr(nc)))
        return lp, npd, nc

    def sssp_routes(self, roots, entries, dest_off, dests, mode=MZ_SSSP_EXACT, path_cap=None):
        """Trees + reachable/get_path for the listed destination nodes → (path_off, path_links)."""
        self.finalize()
        roots = np.ascontiguousarray(roots, np.int32)
        entries = np.ascontiguousarray(entries, np.float64)
        dest_off = np.ascontiguousarray(dest_off, np.int64)
        dests = np.ascontiguousarray(dests, np.int32)
        n = len(roots)
        n_pairs = int(dest_off[-1]) if n else 0
        path_off = np.zeros(n_pairs + 1, np.int64)
        # The buffer is only virtual memory until the engine writes into it, so size it generously: a too small buffer
        # makes mz_sssp_routes report the needed size and the whole batch has to be computed again.
        cap = int(path_cap) if path_cap is not None else max(1024, min(n_pairs * min(self.n_links, 4096), 1 << 29))
        while True:
            links = np.empty(cap, np.int32)
            total = C.c_int64(0)
            rc = lib().mz_sssp_routes(self._h, n, _ptr(roots), _ptr(entries), mode, _ptr(dest_off), _ptr(dests),
                                      _ptr(path_off), _ptr(links), cap, C.byref(total))
            if rc == _lib.MZ_ELIMIT and total.value > cap and path_cap is None:
                cap = int(total.value)
                continue
            check(rc)
            return path_off, links[:total.value]

    def set_stream(self, cuda_stream):
        """Run on a caller-owned stream (int handle, e.g. torch.cuda.current_stream().cuda_stream); None resets."""
        self.finalize()
        check(lib().mz_graph_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def last_stats(self):
        st = _lib.SsspStats()
        check(lib().mz_sssp_last_stats(self._h, C.byref(st)))
        return st.asdict()

    # --- single-tree API with the reference's names (B/Graph.h:212-231) ---
    def labelCorrecting(self, root, t=0.0, info=None):
        if info is not None:
            self.set_linktimes(info)
        lp, npd, nc = self.sssp_batch([root], [t])
        self._root, self._link_pred, self._node_pred, self._node_cost = int(root), lp[0], npd[0], nc[0]

    def predecessorToLink(self, i):
        return int(self._link_pred[i])

    def predecessorToNode(self, i):
        return int(self._node_pred[i])

    def costToNode(self, i):
        return float(self._node_cost[i])

    def reachable(self, des):
        return self._node_pred[des] != SP_UNSET

    def shortest_path_vector(self, des):
        """B/Graph.cpp:653-685 on the downloaded predecessor arrays (pure index walk, no arithmetic)."""
        p = int(self._node_pred[des])
        if p == SP_UNSET:
            raise _lib.MzError(_lib.MZ_EINVAL, "destination not reachable (the reference asserts here)")
        links = []
        seen = set()
        while p >= 0:
            if p in seen:
                return []  # "circular path"
            seen.add(p)
            links.insert(0, p)
            p = int(self._link_pred[p])
        return links

    def close(self):
        if self._h is not None:
            lib().mz_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
