"""world_size-2 gloo test of the multi-GPU host logic (shard ranges, all-gather of partials, min-reduced
first-error key) with the CPU oracle plugged in as the compute backend."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """partial = encoded affine partial sum (128 B) computed by the C oracle; combine = big-int point sums."""

    def __init__(self):
        import c_oracle
        import py_oracle
        self.c, self.po = c_oracle, py_oracle

    def partial(self, local_pairs, n, index_base):
        from blst_eip2537_b200.sharded import STATUS_OK
        data = bytes(local_pairs.numpy())
        key = STATUS_OK
        part = bytes(128)
        if n:
            # first failing pair index: scan pair by pair like the reference does
            for i in range(n):
                err, _ = self.c.call("g1mul", data[160 * i:160 * (i + 1)])
                if err:
                    key = ((index_base + i) << 8) | err
                    break
            else:
                part = self.c.call("g1multiexp", data)[1]
        return torch.frombuffer(bytearray(part), dtype=torch.uint8), torch.tensor([key], dtype=torch.int64)

    def combine(self, parts, count):
        acc = None
        raw = bytes(parts.numpy())
        for i in range(count):
            acc = self.po.ec_add(self.po.FP_OPS, acc, self.po.decode_g1(raw[128 * i:128 * (i + 1)])[1])
        return self.po.encode_g1(acc)


def _worker(rank, world, port, data, n, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blst_eip2537_b200.sharded import shard_range, sharded_multiexp
    lo, hi = shard_range(n, world, rank)
    local = torch.frombuffer(bytearray(data[160 * lo:160 * hi]), dtype=torch.uint8)
    code, out = sharded_multiexp(local, hi - lo, lo, OracleBackend())
    q.put((rank, code, out))
    dist.destroy_process_group()


def _run(data, n, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, data, n, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    return sorted(res)


def test_shard_ranges_cover_everything():
    from blst_eip2537_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_two_rank_multiexp_matches_single_call(oracle_c):
    import workloads as wl
    n = 37
    data, s = wl.g1_msm_input(n, 0xABC)
    want = oracle_c.call("g1multiexp", data)
    res = _run(data, n, 29531)
    assert [(r[1], r[2]) for r in res] == [want, want]


def test_two_rank_first_error_precedence(oracle_c):
    import py_oracle as po
    import workloads as wl
    n = 10
    data, _ = wl.g1_msm_input(n, 0xDEF)
    g = po.encode_g1(po.G1)
    off = g[:64] + po.fp_to_bytes(5)                       # not on curve -> 1
    bad = bytes(16) + po.P.to_bytes(48, "big") + g[64:]     # invalid element -> 3
    d = bytearray(data)
    d[160 * 7:160 * 7 + 128] = off     # shard 1
    d[160 * 2:160 * 2 + 128] = bad     # shard 0, earlier index: must win
    want = oracle_c.call("g1multiexp", bytes(d))
    assert want[0] == 3
    res = _run(bytes(d), n, 29532)
    assert [(r[1], r[2]) for r in res] == [(3, None), (3, None)]
