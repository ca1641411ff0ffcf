"""CPU tests of the N>1 host logic with world_size-2 gloo process groups: channel sharding is a partition, every rank
derives identical plans for its block from the pure-host filter-chain code, the union of the ranks' work is the whole
bank, and `bench.py --impl reference` under 2 ranks prints exactly one line."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, fs, fcs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from sdrangel_b200.sharding import shard_channels, filter_chain, tree_stage_inputs
    lo, hi = shard_channels(len(fcs), world, rank)
    plans = [filter_chain(fs, 48000, fc) for fc in fcs[lo:hi]]
    si, nodes = tree_stage_inputs([p for _, _, p in plans])
    # the baseband "broadcast": rank 0's buffer reaches every rank unchanged (gloo stands in for NCCL on CPU)
    x = torch.arange(1024, dtype=torch.int32) * (7 if rank == 0 else 0)
    dist.broadcast(x, src=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "range": (lo, hi), "plans": plans, "stage_inputs": si, "nodes": nodes, "bcast": int(x.sum())})
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_bank_sharding_partition_and_plans(golden_meta, world):
    plan = golden_meta["chan_plans"]["bank1024"]
    fs, rows = plan["input_rate"], plan["channels"]
    fcs = [r[0] for r in rows]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, fs, fcs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = sorted(res, key=lambda d: d["rank"])
    # partition: contiguous, disjoint, complete
    assert res[0]["range"][0] == 0 and res[-1]["range"][1] == len(fcs)
    for a, b in zip(res[:-1], res[1:]):
        assert a["range"][1] == b["range"][0]
    # plans derived independently on each rank equal the reference's (golden), channel by channel
    got = [tuple(p) for d in res for p in d["plans"]]
    assert got == [(rate, ofs, path) for _, rate, ofs, path in rows]
    # per-rank tree work at 2 ranks is half of the single-GPU tree (21.0 -> 10.5 stage inputs)
    assert all(abs(d["stage_inputs"] - 10.5) < 1e-9 for d in res)
    assert all(d["bcast"] == 7 * sum(range(1024)) for d in res)


def test_shard_channels_edge_cases():
    sys.path.insert(0, ROOT)
    from sdrangel_b200.sharding import shard_channels, filter_chain, tree_stage_inputs
    for n in (0, 1, 7, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_channels(n, world, r)
                cover += list(range(lo, hi))
                assert 0 <= hi - lo <= n // world + 1
            assert cover == list(range(n))
    assert filter_chain(10_000_000, 48000, 1234567) == (156250, -15433, "ULCCCC")
    assert filter_chain(10_000_000, 10_000_000, 0) == (10_000_000, 0, "")
    si, nodes = tree_stage_inputs(["ULC", "ULL", "C"])
    assert nodes == 5 and abs(si - (1 + 0.5 + 0.25 + 0.25 + 1)) < 1e-12


def test_filter_chain_host_matches_all_golden_plans(golden_meta):
    sys.path.insert(0, ROOT)
    from sdrangel_b200.sharding import filter_chain
    plans = golden_meta["chan_plans"]
    for name in ("bank64", "bank1024"):
        for fc, rate, ofs, path in plans[name]["channels"]:
            assert filter_chain(plans[name]["input_rate"], 48000, fc) == (rate, ofs, path)
    for fs, req, fc, rate, ofs, path in plans["random"]:
        assert filter_chain(fs, req, fc) == (rate, ofs, path)


def test_reference_arm_under_two_ranks_prints_one_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29870 + os.getpid() % 100), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--workload", "decimateii"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")


def test_coop_plan_host_logic(golden_meta):
    """Cooperative plan: every channel is owned by exactly one rank below one depth-k node, and the per-rank tree work is
    close to 1/N of the single-GPU tree (21.0 stage inputs for the 1024-channel plan)."""
    sys.path.insert(0, ROOT)
    from sdrangel_b200.coop import CoopPlan, HALO
    plan = golden_meta["chan_plans"]["bank1024"]
    fcs = [r[0] for r in plan["channels"]]
    n = 3 << 24
    for world in (1, 2, 4, 8):
        p = CoopPlan(plan["input_rate"], fcs, 48000, world, n)
        assert p.k == {1: 0, 2: 1, 4: 2, 8: 3}[world]
        owned = sorted(i for r in range(world) for v in p.rank_nodes[r] for i in p.channels_of(r, v))
        assert owned == list(range(len(fcs)))
        assert p.m * world == n and p.mk << p.k == p.m and p.skip << p.k == HALO
        work = [p.stage_inputs(r) for r in range(world)]
        assert max(work) <= 21.0 / world * 1.25 + 0.01, (world, work)
    with pytest.raises(ValueError):
        CoopPlan(plan["input_rate"], fcs, 48000, 8, 12345)
