"""world_size-2 gloo test of the multi-GPU host logic (sample split + one sum-reduce of the accumulator).
The per-rank renderer is stood in for by the CPU oracle: the split / reduce code is the same."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cuda_pt_b200  # noqa: F401  (registers the package)
from cuda_pt_b200 import distributed as cd
from conftest import GOLDEN, ROOT


def test_sample_range_partitions_exactly():
    for total in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                base, n = cd.sample_range(total, world, r)
                seen += list(range(base + 1, base + n + 1))
            assert seen == list(range(1, total + 1))
            counts = [cd.sample_range(total, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        cd.sample_range(4, 2, 2)
    assert cd.reference_seed_offset(3, 10) == 3 * 4201 + 10


def test_row_range_partitions_exactly():
    for height in (4, 256, 1080, 1081, 30):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                lo, hi = cd.row_range(height, world, r)
                assert lo % 4 == 0 and lo <= hi <= height
                rows += list(range(lo, hi))
            assert rows == list(range(height))
            sizes = [cd.row_range(height, world, r)[1] - cd.row_range(height, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 4 + 3
    with pytest.raises(ValueError):
        cd.row_range(100, 2, 2)


def test_balanced_rows_partition_and_follow_the_cost():
    """Dynamic tile split (DistributedRenderer, split="tiles"): the new bands always partition the frame into multiples of 4
    rows with at least one unit per rank, and the cumulative cost (time per row of the previous bands) is cut into equal parts."""
    rng = np.random.default_rng(0)
    for height in (8, 256, 1080):
        for world in (1, 2, 4, 8):
            if height // 4 < world:
                continue
            rows = [cd.row_range(height, world, r) for r in range(world)]
            for _ in range(5):
                times = list(rng.uniform(1.0, 10.0, size=world))
                new = cd.balanced_rows(height, rows, times)
                assert new[0][0] == 0 and new[-1][1] == height
                for (a, b), (c, d) in zip(new[:-1], new[1:]):
                    assert b == c and a % 4 == 0 and b > a
                assert new[-1][1] > new[-1][0]
                rows = new
    # equal densities: bands stay (nearly) equal; a rank that took 3x as long per row gets about a third of the rows
    rows = [cd.row_range(1080, 4, r) for r in range(4)]
    same = cd.balanced_rows(1080, rows, [5.0] * 4)
    assert all(abs((b - a) - 270) <= 8 for a, b in same)
    skew = cd.balanced_rows(1080, rows, [10.0, 30.0, 30.0, 10.0])
    cost = lambda lo, hi: sum((10.0 if (y < 272 or y >= 812) else 30.0) / 270.0 for y in range(lo, hi))
    costs = [cost(a, b) for a, b in skew]
    assert max(costs) / min(costs) < 1.1, costs


def _tile_worker(rank, world, port, total_spp, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = oracle.OracleScene(os.path.join(GOLDEN, "cbox_wfpt_256.cpts"))
    lo, hi = cd.row_range(64, world, rank)            # split the band [96,160) of the frame into per-rank tiles
    s = np.zeros((256, 256, 4), np.float32)
    sc.render(total_spp, rows=(96 + lo, 96 + hi), sum_image=s)
    t = torch.from_numpy(s.copy())
    cd.reduce_sum_(t)                                  # zero-filled SUM of disjoint tiles
    if rank == 0:
        np.save(os.path.join(out_dir, "tiles.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_rank_tile_split_is_bit_identical_to_single_process(tmp_path):
    import oracle
    mp.spawn(_tile_worker, args=(2, 29519, 3, str(tmp_path)), nprocs=2, join=True)
    sc = oracle.OracleScene(os.path.join(GOLDEN, "cbox_wfpt_256.cpts"))
    one, _ = sc.render(3, rows=(96, 160))
    assert np.array_equal(np.load(tmp_path / "tiles.npy"), one)  # disjoint pixels: x + 0 == x exactly


def _worker(rank, world, port, total_spp, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = oracle.OracleScene(os.path.join(GOLDEN, "cbox_wfpt_256.cpts"))
    base, n = cd.sample_range(total_spp, world, rank)
    s = np.zeros((256, 256, 4), np.float32)
    if n:
        sc.render(n, first_accum=base + 1, rows=(96, 128), sum_image=s)
    local_mean = torch.from_numpy(s / max(n, 1))  # what PythonRenderer.render() returns on this rank
    t = torch.from_numpy(s.copy())
    cd.reduce_sum_(t)
    mean, spp = cd.reduce_rendered_image(local_mean, n)
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), t.numpy())
        np.save(os.path.join(out_dir, "mean.npy"), mean.numpy())
        assert spp == total_spp
    dist.destroy_process_group()


def test_two_rank_sample_split_reproduces_single_process(tmp_path):
    import oracle
    total = 5
    mp.spawn(_worker, args=(2, 29517, total, str(tmp_path)), nprocs=2, join=True)
    sc = oracle.OracleScene(os.path.join(GOLDEN, "cbox_wfpt_256.cpts"))
    one, _ = sc.render(total, rows=(96, 128))
    two = np.load(tmp_path / "sum.npy")
    assert np.allclose(two, one, rtol=1e-6, atol=1e-7)          # same paths, fp32 summation order differs
    assert np.allclose(np.load(tmp_path / "mean.npy"), one / total, rtol=1e-5, atol=1e-6)
