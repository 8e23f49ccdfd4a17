"""Edge cases of the native path on the GPU: ragged batches (half-filled 128-row tiles), single samples, single frames,
ragged decoder chunks, plan re-packing after an in-place parameter update, and argument validation."""
import pytest
import torch

from util import O, make_first_stage, make_flow, maxabs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [1, 3, 5])
@pytest.mark.parametrize("prec,tol", [("fp32_simt", 2e-4), ("fp32", 3e-4)])
def test_flow_ragged_batches(B, prec, tol):
    """B*64 pixel rows are not a multiple of the 128-row MMA tile for odd B; results must not depend on batch composition."""
    cfg = O.flow_config(flow_in_channels=32, flow_mid_channels=128, h_channels=128, num_steps=[2, 1, 1, 1, 1], factor=16)
    sd = O.synth_flow_state_dict(cfg, seed=11)
    z, cond, _ = O.synth_inputs(5, 32, 128, 8, seed=5)
    with torch.no_grad():
        x_ref = O.flow_reverse(sd, cfg, z[:B], cond[:B])
        zf_ref, ld_ref = O.flow_forward(sd, cfg, x_ref, cond[:B])
    m = make_flow(cfg, sd, prec, max_batch=8)
    x = m(z[:B].cuda(), cond[:B].cuda(), reverse=True)
    zf, ld = m(x_ref.cuda(), cond[:B].cuda())
    assert maxabs(x, x_ref) < tol and maxabs(zf, zf_ref) < tol
    assert maxabs(ld, ld_ref) < 50 * tol
    # sample i of a batch equals the same sample run alone (no cross-sample leakage through shared tiles / workspaces)
    x1 = m(z[B - 1:B].cuda(), cond[B - 1:B].cuda(), reverse=True)
    assert maxabs(x1, x[B - 1:B]) < 1e-6


def test_flow_rejects_bad_arguments():
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[1, 1], factor=4)
    m = make_flow(cfg, O.synth_flow_state_dict(cfg, seed=1), "fp32", max_batch=2)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 16, 4, 4, device="cuda"), torch.zeros(1, 16, 8, 8, device="cuda"))
    with pytest.raises(ValueError):
        m(torch.zeros(2, 16, 8, 8, device="cuda"), torch.zeros(1, 16, 8, 8, device="cuda"))
    # batch larger than the plan: the plan is rebuilt for the new size
    out = m(torch.zeros(3, 16, 8, 8, device="cuda"), torch.zeros(3, 16, 8, 8, device="cuda"), reverse=True)
    assert out.shape == (3, 16, 8, 8) and torch.isfinite(out).all()


def test_flow_repacks_after_inplace_update():
    """An optimizer-style in-place update of a parameter must invalidate the packed plan (version counters)."""
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[1, 1], factor=4)
    sd = O.synth_flow_state_dict(cfg, seed=2)
    z, cond, _ = O.synth_inputs(2, 16, 16, 8, seed=3)
    m = make_flow(cfg, sd, "fp32_simt")
    x0 = m(z.cuda(), cond.cuda(), reverse=True)
    key = "flow.layers.0.0.actnorm1.bias"
    with torch.no_grad():
        dict(m.named_parameters())[key].add_(0.25)
    sd2 = dict(sd)
    sd2[key] = sd[key] + 0.25
    with torch.no_grad():
        x_ref = O.flow_reverse(sd2, cfg, z, cond)
    x1 = m(z.cuda(), cond.cuda(), reverse=True)
    assert maxabs(x1, x_ref) < 2e-4
    assert maxabs(x1, x0) > 1e-3


@pytest.mark.parametrize("B,T,chunk", [(1, 1, 0), (3, 2, 2), (2, 5, 1)])
def test_decode_ragged_chunks_and_lengths(B, T, chunk):
    cfg = O.first_stage_config(z_dim=32, spatial=64)
    sd = O.synth_first_stage_state_dict(cfg, seed=8)
    g = torch.Generator().manual_seed(4)
    motion = torch.randn((B, 32, 8, 8), generator=g) * 1.3
    x0 = torch.rand((B, 3, 64, 64), generator=g) * 2 - 1
    with torch.no_grad():
        ref = O.decode_first_stage(sd, cfg, motion, x0, T)
    for prec, tol in (("fp32_simt", 1e-4), ("fp32", 1e-3)):
        m = make_first_stage(cfg, sd, prec, max_batch=B, max_frames=T, chunk_videos=chunk)
        out = m.decode(motion.cuda(), x0.cuda(), T)
        assert out.shape == ref.shape
        assert maxabs(out, ref) < tol, (prec, maxabs(out, ref))


def test_decode_rejects_bad_shapes():
    cfg = O.first_stage_config(z_dim=32, spatial=64)
    m = make_first_stage(cfg, O.synth_first_stage_state_dict(cfg, seed=8), "fp32", max_batch=1, max_frames=1)
    with pytest.raises(ValueError):
        m.decode(torch.zeros(1, 32, 8, 8, device="cuda"), torch.zeros(1, 3, 128, 128, device="cuda"), 1)
