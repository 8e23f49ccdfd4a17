"""Host-side logic of the sharded optimizer (no GPU): shard bounds of the flat parameter buffer, and the no-CPU-fallback rule."""
import pytest
import torch


def test_shard_range_partitions_the_padded_buffer():
    from ipoke_b200 import shard_range
    for n in (1, 7, 64, 1237326840):
        for world in (1, 2, 3, 4, 8):
            bounds = [shard_range(n, world, r) for r in range(world)]
            padded = bounds[0][0]
            assert padded % world == 0 and 0 <= padded - n < world
            assert bounds[0][1] == 0 and bounds[-1][2] == padded
            for (p0, lo0, hi0), (p1, lo1, hi1) in zip(bounds, bounds[1:]):
                assert p0 == p1 and hi0 == lo1 and hi0 - lo0 == hi1 - lo1


def test_trainer_refuses_cpu():
    import ipoke_b200 as ipk
    from oracle import ipoke_oracle as O
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=16, h_channels=8, num_steps=[1], factor=2)
    m = ipk.SupervisedMacowTransformer(cfg)
    m.load_state_dict(O.synth_flow_state_dict(cfg, seed=1), strict=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ipk.FlowTrainer(m)
