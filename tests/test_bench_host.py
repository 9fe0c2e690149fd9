"""bench.py's host-side logic on CPU: the configs[3] frame list and its i % N shards, the synthetic frame preparation of a
rank, the algorithmic-byte formulas of SURVEY.md 8(d) and the NUMA binding helper (a no-op without NVML)."""
import os

import numpy as np

from cases import ROOT  # noqa: F401  (puts the repo root on sys.path)

import bench


def test_configs3_list_is_the_1000_val_shapes_and_shards_partition_it():
    shapes = bench.ms_shape_list()
    assert len(shapes) == bench.MS_TOTAL == 1000
    assert shapes == bench.ms_shape_list()                       # fixed order (seeded shuffle)
    assert (480, 640) in shapes and min(h for h, w in shapes) >= 100
    for world in (1, 2, 4, 8):
        shards = [[i for i in range(len(shapes)) if i % world == r] for r in range(world)]
        assert sorted(i for s in shards for i in s) == list(range(1000))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        # shuffled list: every shard sees about the same pixels
        px = [sum(shapes[i][0] * shapes[i][1] for i in s) for s in shards]
        assert max(px) / min(px) < 1.1


def test_prepare_frames_of_a_rank(monkeypatch):
    monkeypatch.setattr(bench, "MS_TOTAL", 24)
    bench.prepare_frames(rank=1, world=4, want_secondary=True, workers=2)
    ss3, ss20, ms = bench._FRAMES["ss3"], bench._FRAMES["ss20"], bench._FRAMES["ms"]
    assert len(ss3) == 64 and len(ss20) == 32 and len(ms) == 6      # frames 1, 5, 9, ... of 24
    f = ss3[0]
    assert (f["H"], f["W"]) == bench.DEC_HW and len(f["scales"]) == 1
    paf, heat, pd, pr = f["scales"][0]
    assert paf.shape == (84, 89, 38) and heat.shape == (84, 89, 19) and paf.dtype == np.float32
    shapes = bench.ms_shape_list()
    for k, fr in enumerate(ms):
        assert (fr["H"], fr["W"]) == shapes[1 + 4 * k] and len(fr["scales"]) == 4
    # another rank gets other frames (different seeds)
    bench.prepare_frames(rank=0, world=4, want_secondary=False, workers=1)
    assert len(bench._FRAMES["ss3"]) == bench.DEC_FRAMES and not bench._FRAMES["ms"]
    assert not np.array_equal(bench._FRAMES["ss3"][0]["scales"][0][1], heat)


def test_algorithmic_bytes_follow_the_survey():
    per = bench.gt_bytes_per_sample(3)
    assert per["k_warp_fused"] == 406272 + 135424 + 406272 + 8464 + 48 == 956480
    assert per["step"] == 1438928 + 864 * 3
    # single scale: 4 (57 h w + 74 H W); ski.jpg-shaped frame = 143.75 MB
    assert bench.decode_bytes_per_frame(674, 712, [(84, 89)]) == 4 * (57 * 84 * 89 + 74 * 674 * 712) == 143751376
    # multi scale: 4*57*sum(h w) + 8*74*H*W
    grids = [(23, 31), (46, 62), (69, 92), (92, 123)]
    assert bench.decode_bytes_per_frame(480, 640, grids) == 4 * 57 * sum(h * w for h, w in grids) + 592 * 480 * 640
    assert bench.blob_bytes_per_frame([(84, 89)]) == 4 * 57 * 84 * 89


def test_numa_binding_is_harmless_without_nvml():
    before = os.sched_getaffinity(0)
    got = bench.bind_to_gpu_numa_node(0)
    assert got is None or isinstance(got, int)
    if got is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)
