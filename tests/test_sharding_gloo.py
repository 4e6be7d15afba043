"""CPU-only, world_size 2 over gloo: the N>1 path of bench.py / shard.py — every rank converts its own
contiguous chunk range independently (here with the CPU oracle standing in for the kernels), the only
cross-rank datum is one byte total per string column, exclusive-scanned on the host."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from duckdb_mbt_b200 import chunks as ch
    from duckdb_mbt_b200 import shard
    b = ch.config_c3(20_000, pattern="ragged", seed=11)  # same seeded table on every rank
    c0, c1, row0, row1 = shard.shard_rows(b.counts, world, rank)
    sub = shard.slice_batch(b, c0, c1)
    o, d = oracle.OracleResult(sub).arrow_string(0, 0)
    assert o.shape[0] == row1 - row0 + 1
    # the one cross-rank datum: this rank's byte total
    totals = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(totals, torch.tensor([d.shape[0]], dtype=torch.int64))
    bases = shard.string_bases([int(t.item()) for t in totals])
    np.save(os.path.join(out_dir, f"off{rank}.npy"), o[:-1].astype(np.int64) + bases[rank])
    np.save(os.path.join(out_dir, f"dat{rank}.npy"), d)
    # bench.py timing rule: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_stitch_to_the_single_rank_result(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    import oracle
    from duckdb_mbt_b200 import chunks as ch
    b = ch.config_c3(20_000, pattern="ragged", seed=11)
    eo, ed = oracle.OracleResult(b).arrow_string(0, 1)
    offs = np.concatenate([np.load(tmp_path / f"off{r}.npy") for r in range(world)])
    data = np.concatenate([np.load(tmp_path / f"dat{r}.npy") for r in range(world)])
    assert np.array_equal(offs, eo[:-1]) and data.shape[0] == eo[-1]
    assert np.array_equal(data, ed)
