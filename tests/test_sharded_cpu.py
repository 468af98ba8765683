"""world_size-2 gloo test of the row-sharded exchange logic (recommendation-models_b200/sharded.py)
on CPU: the orchestration is the production code, the arithmetic is the oracle."""
import numpy as np
import pytest


def test_owner_mapping_is_a_bijection(pkg):
    from recommendation_models_b200.sharded import ShardSpec
    for rows, world in [(39 * 256, 2), (39 * 1024, 8), (1000, 4), (39 * (1 << 18), 8)]:
        specs = [ShardSpec(rows, world, r) for r in range(world)]
        ids = np.arange(min(rows, 200000))
        own, loc = specs[0].owner(ids), specs[0].local_row(ids)
        assert len(set(zip(own.tolist(), loc.tolist()))) == len(ids)
        for r in range(world):
            m = own == r
            assert np.array_equal(specs[r].global_id(r, loc[m]), ids[m])
        # balance: the hot first ids of the fields do not all land on one rank
        _, voc = pkg.synth.field_layout(rows, 39)
        hot = np.cumsum(np.concatenate([[0], voc[:-1]]))
        counts = np.bincount(specs[0].owner(hot), minlength=world)
        assert counts.max() <= int(np.ceil(39 / world)) + 2, counts


def test_range_partition_is_a_bijection(pkg):
    """mode="range": the reference's ColumnRangePartitioner layout (ParRecModel.scala:77,81,98,116)."""
    from recommendation_models_b200.sharded import ShardSpec
    for rows, world in [(39 * 256, 2), (1000, 4), (1001, 8)]:
        specs = [ShardSpec(rows, world, r, mode="range") for r in range(world)]
        assert specs[0].period == -specs[0].rows_local
        ids = np.arange(rows)
        own, loc = specs[0].owner(ids), specs[0].local_row(ids)
        assert np.all(np.diff(own) >= 0) and own.max() == world - 1      # contiguous, ascending ranges
        assert loc.max() < specs[0].rows_local
        for r in range(world):
            m = own == r
            assert np.array_equal(specs[r].global_id(r, loc[m]), ids[m])


@pytest.mark.parametrize("world", [2])
def test_sharded_step_gloo(pkg, world):
    import socket
    import torch.multiprocessing as mp
    import sharded_check
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=sharded_check._cpu_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, r in res:
        assert not (isinstance(r, str) and r.startswith("ERR")), r
    assert sum(r for _, r in res) > 0
