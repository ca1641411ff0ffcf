"""Channel sharding of a DownChannelizer bank over the ranks of one node: contiguous-in-frequency blocks,
so each rank's half-band tree is one narrow path down to depth log2(world) and then a full subtree.  Pure host logic."""
import ctypes as C

from . import capi


def shard_channels(n_channels, world, rank):
    """[lo, hi) of the channels (sorted by centre frequency) rank `rank` owns; blocks differ by at most one channel."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_channels, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def filter_chain(input_rate, requested_rate, center_offset):
    """(out_rate, residual_offset, path) exactly as DownChannelizer::applyConfiguration would choose it; no device needed."""
    rate, ofs = C.c_int32(), C.c_int32()
    modes = (C.c_int32 * 32)()
    n = capi.lib().b200dsp_filter_chain(int(input_rate), int(requested_rate), int(center_offset), C.byref(rate), C.byref(ofs), modes, 32)
    if n < 0:
        capi.check(n)
    return rate.value, ofs.value, "".join("CLU"[modes[i]] for i in range(n))


def tree_stage_inputs(paths):
    """Stage-input samples per baseband sample of the shared-prefix tree over `paths` (node at depth d costs 2^-(d-1))."""
    nodes = {p[:k] for p in paths for k in range(1, len(p) + 1)}
    return sum(2.0 ** -(len(s) - 1) for s in nodes), len(nodes)
