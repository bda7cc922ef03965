"""N>1 host logic on CPU: world_size-2 gloo run of the read-id sharding and of the one
collective of the path (the counter sum).  Each rank's per-read counters come from the oracle chain here
(this is a test: on the GPU box the same host code wraps the CUDA path, see bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from elector_b200 import TALLY_FIELDS  # noqa: E402
from elector_b200 import shard  # noqa: E402


def _oracle_counters(wl):
    """oracle POA -> Donatello merge -> tally for every read of wl: int64[n_reads, K]"""
    from oracle import oracle, tally_oracle as to
    rf = wl["read_first"]
    n_reads = len(rf) - 1
    o = oracle.batch(wl["ref"], wl["ref_off"], wl["cor"], wl["cor_off"], wl["unc"], wl["unc_off"], nthreads=2)
    out = np.zeros((n_reads, len(TALLY_FIELDS)), np.int64)
    for r in range(n_reads):
        rows = [oracle.window_rows(o, w) for w in range(rf[r], rf[r + 1])]
        R = "".join(x[0] for x in rows); C = "".join(x[1] for x in rows); U = "".join(x[2] for x in rows)
        keep = [i for i, ch in enumerate(C) if ch != "n"]
        R, C, U = ("".join(s[i] for i in keep) for s in (R, C, U))
        t = to.tally_read(R, C, U)
        out[r] = [t[k] for k in TALLY_FIELDS]
    return out


def _sums(counters):
    s = counters.sum(axis=0)
    ext = TALLY_FIELDS.index("extended")
    s[ext] = counters[:, ext][counters[:, ext] >= 0].sum()   # elector_tally_sum_device: -1 = "not assessed"
    return s


def _synthetic_workload(n_reads, seed):
    """a few short reads cut into ~50-letter windows (no splitter binary needed on the CPU box)"""
    from oracle import synth
    rng = np.random.default_rng(seed)
    wins, first = [], [0]
    for _ in range(n_reads):
        k = int(rng.integers(2, 7))
        wins += synth.hard_windows(k, seed=int(rng.integers(1 << 30)))
        first.append(len(wins))
    import elector_b200
    r, ro = elector_b200.windows_to_csr([w[1] for w in wins])
    c, co = elector_b200.windows_to_csr([w[2] for w in wins])
    u, uo = elector_b200.windows_to_csr([w[3] for w in wins])
    return dict(ref=r, ref_off=ro, cor=c, cor_off=co, unc=u, unc_off=uo, read_first=np.asarray(first, np.int64))


def _worker(rank, world, port, n_reads, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = _synthetic_workload(n_reads, seed=11)
        lo, hi = shard.shard_reads(n_reads, rank, world)
        mine = shard.slice_windows(wl, lo, hi)
        counters = _oracle_counters(mine)
        sums = torch.from_numpy(_sums(counters).copy())
        shard.reduce_counters(sums)                       # the one collective of the path
        allc = shard.gather_counters(torch.from_numpy(counters))
        q.put((rank, lo, hi, sums.numpy().tolist(), allc.numpy().tolist()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 10000, 1000003):
        for world in (1, 2, 4, 8):
            cuts = [shard.shard_reads(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_reads(10, 2, 2)
    rng = np.random.default_rng(0)
    cost = rng.integers(1, 1000, 5000)
    cuts = [shard.shard_reads_balanced(cost, r, 8) for r in range(8)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 5000 and all(cuts[r][1] == cuts[r + 1][0] for r in range(7))
    loads = [cost[a:b].sum() for a, b in cuts]
    assert max(loads) - min(loads) <= 2 * cost.max()


def test_slices_partition_the_workload():
    wl = _synthetic_workload(9, seed=3)
    parts = [shard.slice_windows(wl, *shard.shard_reads(9, r, 4)) for r in range(4)]
    for k in ("ref", "cor", "unc"):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), wl[k])
        assert sum(len(p[k + "_off"]) - 1 for p in parts) == len(wl[k + "_off"]) - 1
        for p in parts:
            assert p[k + "_off"][0] == 0 and p[k + "_off"][-1] == len(p[k])
    assert sum(len(p["read_first"]) - 1 for p in parts) == 9


@pytest.mark.timeout(300)
def test_two_ranks_gloo_equal_one_rank():
    import torch.multiprocessing as mp
    n_reads, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_reads, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = _oracle_counters(_synthetic_workload(n_reads, seed=11))
    assert [g[1:3] for g in got] == [(0, 3), (3, 7)]
    for g in got:                                          # every rank ends with the global result
        assert g[3] == _sums(whole).tolist()
        assert g[4] == whole.tolist()
