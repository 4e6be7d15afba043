"""The all-core columnar CPU line (oracle/columnar.c, bench.py's `cpu_columnar`) converts correctly: its Arrow buffers
equal the oracle's (oracle.c, the restatement of the reference's per-cell loops) on seeded batches, so the throughput
bench.py reports for it is the throughput of a correct conversion."""
import numpy as np
import pytest

import oracle
from oracle import columnar
from duckdb_mbt_b200 import chunks as ch


@pytest.mark.parametrize("pattern", ["full", "ragged"])
def test_columnar_cpu_matches_the_oracle(pattern):
    b = ch.config_c2(30_011, seed=5)
    if pattern == "ragged":
        rng = np.random.default_rng(4)
        b4 = ch.config_c4(20_000, ncols=6, pattern="ragged")
        b3 = ch.config_c3(20_000, pattern="ragged")
        batches = [b4, b3]
    else:
        batches = [b, ch.config_c4(10_000, ncols=3), ch.config_c3(40_000)]
    for batch in batches:
        conv = columnar.ColumnarConverter(batch, threads=4)
        conv.run()
        conv.run()  # a second pass over the same outputs, like the timed loop
        ora = oracle.OracleResult(batch)
        n = batch.nrows
        for j, (col, o) in enumerate(zip(batch.columns, conv.out)):
            if col.phys == ch.P_STRING:
                eo, ed = ora.arrow_string(j, 1 if o["large"] else 0)
                assert np.array_equal(o["offsets"], eo)
                assert o["total"] == ed.shape[0] and np.array_equal(o["data"][: ed.shape[0]], ed)
                _, bm, _, nc = ora.arrow_fixed(j, ch.D_SAME, 16, want_values=False)
            else:
                ev, bm, _, nc = ora.arrow_fixed(j, o["dst"], o["width"])
                assert np.array_equal(o["values"], ev), col.name
            assert np.array_equal(o["bitmap"][: (n + 7) // 8], bm[: (n + 7) // 8]), col.name
            assert o["nulls"] == nc
        ora.close()
