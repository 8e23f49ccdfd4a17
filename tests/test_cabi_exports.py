"""CPU checks of the drop-in boundary: libipoke_b200.so loads, exports every symbol include/ipoke_b200.h declares, reports
errors through status codes + ipk_last_error (no compute calls: there is no GPU here), and the ctypes structs match the
header's layout."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ipoke_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ipk_[A-Za-z0-9_]+)\s*\(", src)))


def _lib():
    from ipoke_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_symbols_are_exported():
    L = _lib()
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/ipoke_b200.h but not exported: {missing}"
    assert sorted(L.EXPORTS) == names, "ipoke_b200/_lib.py EXPORTS must list exactly the header's entry points"


def test_version_and_error_conventions():
    L = _lib()
    lib = L.lib()
    assert lib.ipk_version() == 100
    h = ctypes.c_void_p()
    # null config -> IPK_ERR_INVALID, message available, nothing thrown across the boundary
    assert lib.ipk_flow_create(None, ctypes.byref(h)) == -1
    assert b"null" in lib.ipk_last_error()
    c = L.FlowConfig()
    c.flow_in_channels, c.flow_mid_channels, c.h_channels, c.n_levels, c.factor = 32, 2048, 128, 20, 16
    c.kernel_h, c.kernel_w, c.precision, c.max_batch = 2, 3, 1, 4
    assert lib.ipk_flow_create(ctypes.byref(c), ctypes.byref(h)) == -1          # num_layers < factor (macow2.py:834)
    assert b"factor" in lib.ipk_last_error()
    c.n_levels = 15
    for i, s in enumerate([10, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1, 1]):
        c.num_steps[i] = s
    c.kernel_h, c.kernel_w = 3, 3
    assert lib.ipk_flow_create(ctypes.byref(c), ctypes.byref(h)) == -6          # IPK_ERR_UNSUPPORTED
    c.kernel_h, c.kernel_w = 2, 3
    assert lib.ipk_flow_create(ctypes.byref(c), ctypes.byref(h)) == 0 and h.value
    # run before finalize -> IPK_ERR_STATE
    assert lib.ipk_flow_reverse(h, None, None, None, 1, None) == -5
    assert lib.ipk_flow_destroy(h) == 0
    d = L.FsConfig()
    d.z_dim, d.spatial, d.n_gru_layers, d.n_dec, d.precision, d.max_batch, d.max_frames = 32, 128, 4, 4, 1, 2, 2
    for i, ch in enumerate([256, 256, 128, 64]):
        d.dec_channels[i] = ch
    assert lib.ipk_fs_create(ctypes.byref(d), ctypes.byref(h)) == -1            # 4 dec channels -> 64x64, not 128
    with pytest.raises(RuntimeError, match="status -1"):
        L.check(-1, "x")


def test_struct_layout_matches_header():
    L = _lib()
    assert ctypes.sizeof(L.FlowConfig) == 4 * (4 + 32 + 5)
    assert ctypes.sizeof(L.FsConfig) == 4 * (4 + 8 + 4)
    assert ctypes.sizeof(L.EncConfig) == 4 * (5 + 8 + 3)
    assert ctypes.sizeof(L.CencConfig) == 4 * 6
    src = open(HEADER).read()
    assert "#define IPK_MAX_LEVELS 32" in src and "#define IPK_MAX_DEC 8" in src


def test_modules_refuse_cpu():
    """No CPU fallback: the drop-in modules raise when asked to run on CPU tensors."""
    import torch
    import ipoke_b200 as ipk
    from oracle import ipoke_oracle as O
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=16, h_channels=4, num_steps=[1], factor=2)
    m = ipk.SupervisedMacowTransformer(cfg)
    m.load_state_dict(O.synth_flow_state_dict(cfg, seed=1), strict=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 16, 8, 8), torch.zeros(1, 4, 8, 8), reverse=True)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 15, 8, 8), torch.zeros(1, 4, 8, 8))
    # uninitialised checkpoint (data-dependent init pending) is rejected
    m2 = ipk.SupervisedMacowTransformer(cfg)
    assert m2.flow.reshape == "none" and m2.flow.z_channels == 8


def test_state_dict_layout_matches_oracle_checkpoint():
    """The drop-in modules expose exactly the reference's state-dict keys/shapes/dtypes."""
    import ipoke_b200 as ipk
    from oracle import ipoke_oracle as O
    cfg = O.flow_config(flow_in_channels=32, flow_mid_channels=64, h_channels=128)
    sd = O.synth_flow_state_dict(cfg, seed=0)
    m = ipk.SupervisedMacowTransformer(cfg)
    own = m.state_dict()
    assert list(own.keys()) == list(sd.keys()) or sorted(own.keys()) == sorted(sd.keys())
    for k in sd:
        assert own[k].shape == sd[k].shape and own[k].dtype == sd[k].dtype, k
    assert len(own) == 6995
    dcfg = O.first_stage_config(z_dim=32, spatial=128)
    dsd = O.synth_first_stage_state_dict(dcfg, seed=0)
    d = ipk.SpadeCondMotionDecoder(dcfg)
    own = d.state_dict()
    assert sorted(own.keys()) == sorted(dsd.keys())
    for k in dsd:
        assert own[k].shape == dsd[k].shape, k
    for nf_in, size in ((2, 128), (3, 64)):
        ccfg = O.cond_encoder_config(nf_in=nf_in, spatial=size)
        csd = O.synth_cond_encoder_state_dict(ccfg, seed=0)
        ce = ipk.ConvEncoder(nf_in, ccfg["nf_max"], ccfg["n_stages"])
        own = ce.state_dict()
        assert sorted(own.keys()) == sorted(csd.keys())
        for k in csd:
            assert own[k].shape == csd[k].shape, k
    for size in (64, 128):
        ecfg = O.encoder_config(z_dim=32, img_size=size, max_frames=10)
        esd = O.synth_encoder_state_dict(ecfg, seed=0)
        e = ipk.ResNetMotionEncoder(dict(ecfg))
        own = e.state_dict()
        assert sorted(own.keys()) == sorted(esd.keys())
        for k in esd:
            assert own[k].shape == esd[k].shape, k


def _tc_plan(Npad, tiles_m, iters, nsub=1, nsplit=1, fused=0, sms=148):
    import ctypes
    from ipoke_b200 import _lib
    out = (ctypes.c_int32 * 40)()
    _lib.check(_lib.lib().ipk_test_tc_plan(Npad, tiles_m, iters, nsub, nsplit, fused, sms, out), "ipk_test_tc_plan")
    BN, bn, CG, tiles_n, nv = out[0], out[1], out[2], out[3], out[4]
    return BN, bn, CG, tiles_n, [(out[5 + 2 * i], out[6 + 2 * i]) for i in range(nv)]


def test_conv_engine_tile_planning_host_logic():
    """N tiling of the tcgen05 conv engine (conv_tc.cu: tc_plan_tiles), pure host logic, for the shapes of the sampling path on 148 SMs."""
    # invariants over a sweep of layer widths / batch sizes
    for Npad in (16, 32, 48, 64, 80, 96, 128, 144, 160, 192, 224, 256, 288, 512, 1024, 2048):
        for tiles_m in (1, 2, 7, 32, 2048):
            for iters in (1, 3, 9, 32, 288):
                BN, bn, CG, tiles_n, nv = _tc_plan(Npad, tiles_m, iters)
                assert BN in (32, 64, 128, 256) and bn % 16 == 0 and 32 <= bn <= BN and CG in (1, 2)
                if CG == 2:
                    assert BN == 256 and tiles_m >= 2 and iters >= 8
                if nv:
                    widths = [w for _, w in nv]
                    assert tiles_n == len(nv) and sum(widths) == Npad and max(widths) == bn == widths[0]
                    assert widths == sorted(widths, reverse=True) and all(w % 16 == 0 for w in widths)
                    assert [n0 for n0, _ in nv] == [sum(widths[:i]) for i in range(len(nv))]
                else:
                    assert tiles_n * bn >= Npad > (tiles_n - 1) * bn
    # NICE conv2 at B = 64 (M = 4096, N = K = 2048): CTA pairs, nine uneven tiles per M-tile group, no pair runs two wide tiles
    BN, bn, CG, tiles_n, nv = _tc_plan(2048, 32, 32)
    assert (BN, CG, tiles_n) == (256, 2, 9) and [w for _, w in nv] == [240, 240] + [224] * 7
    load = [0] * 74
    for u in range(16 * 9):
        load[u % 74] += nv[u // 16][1]
    assert max(load) == 240 + 224 < 2 * 256
    # NICE conv1 (K <= 192: three k-blocks): single CTAs, uniform 256-column tiles; conv3 (288 columns): two tiles of 144
    assert _tc_plan(2048, 32, 3)[:4] == (256, 256, 1, 8)
    assert _tc_plan(288, 32, 32)[1:4:2] == (144, 2)
    # the GUI's B = 1 call (one M tile): narrow tiles put the weight stream on at least half of the machine
    BN, bn, CG, tiles_n, nv = _tc_plan(2048, 1, 32)
    assert CG == 1 and not nv and tiles_n * 2 > 148 // 2 and bn <= 64
    # fused-epilogue launches keep the cost model's tile
    assert _tc_plan(256, 1, 36, fused=1)[1] == 256
