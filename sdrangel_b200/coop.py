"""Cooperative multi-GPU DownChannelizer bank (SURVEY.md 8e "cooperative top levels").

Plain channel sharding makes every rank redo the top log2(N) tree levels and needs the whole baseband on every GPU.  Here:
  1. the baseband of a step is cut into N time slices; rank r receives slice r (plus a short halo) -- a scatter, 1/N of the bytes;
  2. every rank runs the TOP of the tree (levels 1..k, k = ceil(log2 N)) on its slice only, for all depth-k nodes;
     FIR => finite memory => a HALO of earlier samples makes the slice's outputs exact (outputs of the halo are discarded);
  3. an all-to-all hands every depth-k node stream to the rank(s) whose channels live below that node (time slices from all
     ranks concatenate to the node's whole stream, at 1/2^k of the baseband rate);
  4. each rank feeds its node streams to ordinary banks holding its channels by path suffix (state carried as usual).
The top levels are computed once per node instead of once per rank, and a GPU ingests ~2/N of the baseband instead of all.

CoopPlan is pure host logic (CPU-testable); CoopRank owns one rank's GPU objects; the exchange itself is done by the caller
(torch.distributed in bench.py, plain array copies in the single-GPU emulation test).
"""
import math

from .sharding import shard_channels, filter_chain

HALO = 1536          # baseband samples of warm-up per slice: >= 46 * (2^3 - 1) history of three order-48 stages, multiple of 768


class CoopPlan:
    def __init__(self, input_rate, offsets, requested_rate, world, samples_per_step):
        self.fs, self.world, self.n = int(input_rate), int(world), int(samples_per_step)
        self.chains = [filter_chain(input_rate, requested_rate, fc) for fc in offsets]      # (out_rate, residual, path)
        min_len = min(len(p) for _, _, p in self.chains)
        self.k = min(int(math.ceil(math.log2(world))) if world > 1 else 0, min_len)
        if 46 * ((1 << self.k) - 1) > HALO:
            raise ValueError("CoopPlan: %d cooperative levels need %d samples of history per slice, the halo is %d (world > 32 is not supported)"
                             % (self.k, 46 * ((1 << self.k) - 1), HALO))
        self.nodes = sorted({p[:self.k] for _, _, p in self.chains})                         # depth-k nodes (path prefixes)
        self.ranges = [shard_channels(len(offsets), world, r) for r in range(world)]
        self.rank_nodes = [sorted({self.chains[i][2][:self.k] for i in range(lo, hi)}) for lo, hi in self.ranges]
        if self.n % (world * 768 * (1 << self.k)):
            raise ValueError("samples per step must be a multiple of world * 768 * 2^k")
        self.m = self.n // world                      # baseband samples per time slice
        self.mk = self.m >> self.k                    # node samples per time slice
        self.skip = HALO >> self.k                    # node outputs of the halo (discarded)

    def channels_of(self, rank, node):
        lo, hi = self.ranges[rank]
        return [i for i in range(lo, hi) if self.chains[i][2][:self.k] == node]

    def stage_inputs(self, rank):
        """Tree work of one rank per baseband sample: its share of the top levels plus its subtrees."""
        top = {p[:d] for p in self.nodes for d in range(1, self.k + 1)}
        t = sum(2.0 ** -(len(s) - 1) for s in top) * (self.m + HALO) / self.n
        lo, hi = self.ranges[rank]
        sub = {self.chains[i][2][:d] for i in range(lo, hi) for d in range(self.k + 1, len(self.chains[i][2]) + 1)}
        return t + sum(2.0 ** -(len(s) - 1) for s in sub)


class CoopRank:
    """One rank's banks: `top` (levels 1..k on a time slice, node outputs raw) and one ordinary bank per needed node."""

    def __init__(self, plan, rank, frontend=None, device=None):
        from .dsp import DownChannelizerBank
        self.plan, self.rank = plan, rank
        p = plan
        self.top = DownChannelizerBank(p.fs, device)
        self.top.set_chunk(max(p.m + HALO, 768))
        self.top_ids = {node: self.top.add_channel_path(node, 0) for node in p.nodes}
        self.subs, self.chan = {}, {}                  # node -> bank ; channel index -> (node, chan_id)
        for node in p.rank_nodes[rank]:
            b = DownChannelizerBank(p.fs >> p.k, device)
            b.set_chunk(max(p.n >> p.k, 768))
            for i in p.channels_of(rank, node):
                rate, ofs, path = p.chains[i]
                cid = b.add_channel_path(path[p.k:], len(path))
                if frontend is not None:
                    cutoff, out_rate = frontend
                    b.set_frontend(cid, -ofs, cutoff, out_rate)
                self.chan[i] = (node, cid)
            self.subs[node] = b

    def close(self):
        self.top.close()
        for b in self.subs.values():
            b.close()
