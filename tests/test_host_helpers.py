"""Host-side helpers of dml_b200.ops that are plain torch (no kernels): they run on the CPU too."""
import torch

from dml_b200 import ops


def test_loss_scale_is_a_power_of_two_in_range():
    for amax in (3e-7, 0.02, 1.0, 900.0):
        bits = torch.tensor([amax], dtype=torch.float32).view(torch.int32)
        s, inv = ops.loss_scale_from_amax(bits).tolist()
        assert 4.0 < s * amax <= 8.0 and abs(s * inv - 1.0) < 1e-6
        assert abs(torch.log2(torch.tensor(s)).item() - round(torch.log2(torch.tensor(s)).item())) < 1e-6
    assert ops.loss_scale_from_amax(torch.zeros(1, dtype=torch.int32)).tolist() == [1.0, 1.0]


def test_centre_taps_and_key_count():
    assert ops.centre_taps(16385) == (8192, 8193, 1.0, 0.0)
    assert ops.centre_taps(128)[2:] == (0.5, 0.5)
    assert ops.kv_length(16385, 6, 4) == 4096 and ops.kv_length(2049, 6, 4) == 512


def test_length_bucketed_sampler_partitions_like_a_distributed_sampler():
    import random
    from dml_b200.parallel import LengthBucketedSampler
    rnd = random.Random(3)
    lengths = [rnd.randrange(4000, 16385, 2) for _ in range(37)]
    world = 4
    per_rank = []
    for rank in range(world):
        s = LengthBucketedSampler(lengths, world, rank, seed=11)
        s.set_epoch(2)
        per_rank.append(list(s))
        assert len(per_rank[-1]) == len(s) == 37 // world
    flat = [i for r in per_rank for i in r]
    assert len(set(flat)) == len(flat) == (37 // world) * world                  # each bag at most once, every rank the same count
    order = sorted(lengths, reverse=True)
    for step in zip(*per_rank):                                                  # the bags of one step are neighbours in length
        ls = sorted((lengths[i] for i in step), reverse=True)
        k = order.index(ls[0])
        assert ls == order[k:k + world] or len(set(ls)) < world
    s0 = LengthBucketedSampler(lengths, world, 0, seed=11)
    a = list(s0); s0.set_epoch(1); b = list(s0)
    assert a != b and sorted(a) != [] 


def test_forward_cta_plan_only_splits_when_the_spill_fits_one_wave():
    """ops.plan_half_blocks: the north-star shape on two towers and 148 SMs cuts 9 blocks per (bag, group); a single launch, a
    shape without a partial wave or a spill too large for one wave of short CTAs is left alone."""
    assert ops.plan_half_blocks(16385, 1, 4, 2, 148) == 9
    assert ops.plan_half_blocks(16385, 1, 4, 1, 148) == 0          # 256 CTAs = 1 wave + 108: 216 short CTAs do not fit one wave
    assert ops.plan_half_blocks(300, 1, 4, 2, 148) == 0            # less than one wave in all
    for n in (1, 127, 128, 129, 4097, 6001, 16385, 100001):
        for c in (1, 2, 3):
            for nsm in (132, 148):
                t = ops.plan_half_blocks(n, 1, 4, c, nsm)
                assert 0 <= t <= (-(-n // 128)) // 2
