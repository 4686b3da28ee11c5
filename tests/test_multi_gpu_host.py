"""Host-side logic of the sharded candidate sweep on CPU: world_size-2 (and 3) gloo process groups, each rank
evaluating its shard of the counter-based candidate sequence with the plain-C oracle (the GPU sweep's checker) and
the winners combined by sharding.all_gather_winner. The result must equal the single-process arg-max."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import support as S

pkg = importlib.import_module("sequential-line-search_b200")
sharding = pkg.sharding


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_candidates_do_not_depend_on_the_split():
    whole = sharding.candidate_coords(7, 100, 50, 6)
    assert whole.shape == (6, 50) and whole.min() >= 0.0 and whole.max() < 1.0
    parts = [sharding.candidate_coords(7, 100 + f, c, 6) for f, c in (sharding.shard_range(50, 3, r) for r in range(3))]
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), whole)
    assert not np.array_equal(whole, sharding.candidate_coords(8, 100, 50, 6))


def test_select_winner_rules():
    assert sharding.select_winner([[1.0, 5], [2.0, 9]]) == (2.0, 9)
    assert sharding.select_winner([[2.0, 9], [2.0, 3]]) == (2.0, 3)             # tie -> lowest index
    assert sharding.select_winner([[float("nan"), 1], [0.5, 7]]) == (0.5, 7)    # NaN never wins
    assert sharding.select_winner([[9.0, -1], [0.5, 7]]) == (0.5, 7)            # empty shard
    with pytest.raises(ValueError):
        sharding.select_winner([[float("nan"), 1], [0.0, -1]])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    D, N = 5, 40
    X, theta = S.make_X(N, D, "uniform"), S.make_theta(D, "default")
    return D, X, theta, S.make_y(X)


def _local_best(oracle, m, f_best, D, seed, first, count):
    if count == 0:
        return 0.0, -1
    Q = sharding.candidate_coords(seed, first, count, D)
    val = oracle.acq_batch(m, S.EI, 1.0, f_best, Q)["val"]
    i = int(np.argmax(val))  # first maximum = lowest index
    return float(val[i]), first + i


def _worker(rank, world, port, total, seed, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D, X, theta, y = _model()
    oracle = S.Oracle()
    m = oracle.model(S.SE, X, theta, 0.005, y)
    _, f_best = oracle.f_best(m)
    first, count = sharding.shard_range(total, world, rank)
    v, i = _local_best(oracle, m, f_best, D, seed, first, count)
    out[rank] = sharding.all_gather_winner(v, i)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 3001), (3, 2000), (2, 1)])
def test_sharded_argmax_equals_single_process(world, total):
    seed = 11
    D, X, theta, y = _model()
    oracle = S.Oracle()
    m = oracle.model(S.SE, X, theta, 0.005, y)
    _, f_best = oracle.f_best(m)
    want = _local_best(oracle, m, f_best, D, seed, 0, total)
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), total, seed, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert out[r] == want, (r, out[r], want)


def test_select_best_point_rules():
    v, x, r = sharding.select_best_point([0.1, 0.7, 0.7], [[0, 0], [1, 1], [2, 2]])
    assert (v, r) == (0.7, 1) and np.array_equal(x, [1.0, 1.0])                  # tie -> lowest rank
    v, x, r = sharding.select_best_point([float("nan"), 0.2], [[0, 0], [3, 4]])
    assert (v, r) == (0.2, 1) and np.array_equal(x, [3.0, 4.0])                  # NaN never wins
    with pytest.raises(ValueError):
        sharding.select_best_point([float("nan")], [[0.0]])


def _point_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank "maximised" its own range: rank r found value r / 10 at the point (r, r + 0.5, r + 1)
    out[rank] = sharding.all_gather_best_point(rank / 10.0, np.array([rank, rank + 0.5, rank + 1.0]))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_best_point_world_2():
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_point_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        v, x, w = out[r]
        assert v == 0.1 and w == 1 and np.array_equal(x, [1.0, 1.5, 2.0])
