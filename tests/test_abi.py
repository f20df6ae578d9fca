"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the
header declares, the host parameter API matches the reference, and the engine refuses to
run without a GPU (no CPU fallback).  No GPU compute here."""
import os
import re

import numpy as np
import pytest

from dspi_b200 import api, layouts as L, workloads as W
from tests.util import same_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(api.LIB_PATH):
        from dspi_b200.build import build
        build()
    return api.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dspi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dspi_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dspi_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liborc" not in src and "tests.orc" not in src and "dspi_oracle" not in src, f


@pytest.mark.parametrize("variant", ["A", "B", "mixed"])
@pytest.mark.parametrize("fs", [44100.0, 48000.0, 96000.0])
def test_host_coefficients_match_oracle_and_reference(lib, oracle, variant, fs):
    params = W.eq_params(variant, 16, fs=fs, seed=21)
    for q28 in (False, True):
        p1, p2 = params.copy(), params.copy()
        mine = api.compute_coefficients(p1, q28=q28, fs=fs)
        theirs = np.zeros(params.shape, L.BIQUAD_Q28 if q28 else L.BIQUAD_F32)
        oracle.eq_coeffs(q28, p2, theirs, fs)
        assert same_bits(mine, theirs) and same_bits(p1, p2)


def test_host_coefficients_keep_state_unless_topology_flips(lib):
    p = np.zeros(1, L.EQ_PARAM)
    p[0] = (0, 0, L.PEAKING, 0, 1000.0, 1.0, 3.0)
    bq = api.compute_coefficients(p.copy(), fs=48000.0)
    assert bq[0]["use_svf"] == 1 and bq[0]["bypass"] == 0
    bq[0]["svic1eq"], bq[0]["svic2eq"] = 0.25, -0.5
    p[0]["gain_db"] = 4.0                                   # same topology: state survives (dsp_pipeline.c:87-92)
    api.compute_coefficients(p.copy(), fs=48000.0, biquads=bq)
    assert bq[0]["svic1eq"] == np.float32(0.25)
    p[0]["freq"] = 12000.0                                  # >= fs/7.5 -> biquad path, state cleared
    api.compute_coefficients(p.copy(), fs=48000.0, biquads=bq)
    assert bq[0]["use_svf"] == 0 and bq[0]["svic1eq"] == 0 and bq[0]["s1"] == 0
    p[0]["gain_db"] = 0.0                                   # flat -> bypass, b0 = 1 (KAT, dsp_pipeline.c:62-73)
    api.compute_coefficients(p.copy(), fs=48000.0, biquads=bq)
    assert bq[0]["bypass"] == 1 and bq[0]["b0"] == 1.0


def test_engine_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.dspi_device_count() == 0
    with pytest.raises(api.DspiError, match="no CUDA device|fallback"):
        api.EqEngine("f32f", 64)


def test_bad_arguments_are_rejected(lib):
    import ctypes as C
    h = C.c_void_p()
    desc = api._EqDesc(7, 64, 10, 0, 0)
    assert lib.dspi_eq_create(C.byref(h), C.byref(desc)) == -22
    desc = api._EqDesc(0, 64, 13, 0, 0)
    assert lib.dspi_eq_create(C.byref(h), C.byref(desc)) == -22
    assert lib.dspi_eq_create(None, None) == -22
    assert b"" != lib.dspi_last_error()


def test_chain_parameter_functions_match_the_reference(lib, oracle, refs):
    """crossfeed / leveller / loudness / volume parameter functions, bit-exact vs the strict reference build"""
    import ctypes as C
    from tests.util import field_bits
    for fs in (44100.0, 48000.0, 96000.0):
        for preset in range(4):
            for itd in (0, 1):
                mine = api.crossfeed_coefficients(fs, True, itd, preset, 1234.0, 7.0)
                st = np.zeros(1, L.XFEED_F32)
                refs["f32s"].lib.ref_xfeed_coeffs(st.ctypes.data, 1, itd, preset, 1234.0, 7.0, fs)
                assert np.array_equal(field_bits(np.array([mine])), field_bits(st))
        assert not np.any(field_bits(np.array([api.crossfeed_coefficients(fs, enabled=False)])))
        for speed in (0, 1, 2, 9):
            mine = api.leveller_coefficients(fs, 70.0, speed, 40.0, -120.0)
            want = np.zeros(1, L.LEV_COEFFS)
            refs["f32s"].lib.ref_lev_coeffs(want.ctypes.data, 70.0, speed, 40.0, -120.0, fs)
            assert np.array_equal(field_bits(np.array([mine])), field_bits(want))
        for ref_spl, inten in ((83.0, 100.0), (60.0, 35.0), (130.0, 200.0)):
            want = np.zeros((L.LOUD_STEPS, 2), L.LOUD_F32)
            refs["f32s"].lib.ref_loud_table(want.ctypes.data, ref_spl, inten, fs)
            assert np.array_equal(field_bits(api.loudness_table(fs, ref_spl, inten)), field_bits(want))
    # KATs the reference states: 4.5 dB feed -> G = 0.373 (crossfeed.c:65); 0 dB host volume -> int16 -32768
    g = api.crossfeed_coefficients(48000.0, preset=0)
    assert abs(float(g["lp_a0"]) / (1.0 - float(g["lp_b1"])) - 0.373) < 1e-3
    assert api.host_volume(0) == (-32768, 60) and api.host_volume(-20 * 256) == (0x0CCD, 40) and api.host_volume(-32768)[0] == 0
    for v in range(-70 * 256, 1 * 256, 97):
        assert api.host_volume(v)[0] == oracle.lib.orc_host_vol_mul(v, None)
    for ms in (0.0, 3.3, 42.0, 85.4, 500.0, -1.0):
        for fs in (48000.0, 96000.0):
            for last in (0, 1):
                assert api.delay_samples(ms, fs, last) == refs["f32s"].lib.ref_delay_samples(ms, fs, last)


def test_q28_chain_parameter_functions_match_the_reference(lib, refs):
    from tests.util import field_bits
    for fs in (48000.0, 96000.0):
        for preset in range(4):
            mine = api.crossfeed_coefficients_q28(fs, True, 1, preset, 1234.0, 7.0)
            st = np.zeros(1, L.XFEED_Q28)
            refs["q28"].lib.ref_xfeed_coeffs(st.ctypes.data, 1, 1, preset, 1234.0, 7.0, fs)
            assert np.array_equal(field_bits(np.array([mine])), field_bits(st))
        for ref_spl, inten in ((83.0, 100.0), (60.0, 35.0)):
            want = np.zeros((L.LOUD_STEPS, 2), L.LOUD_Q28)
            refs["q28"].lib.ref_loud_table(want.ctypes.data, ref_spl, inten, fs)
            assert np.array_equal(field_bits(api.loudness_table_q28(fs, ref_spl, inten)), field_bits(want))


def test_multi_gpu_entry_points_reject_bad_arguments(lib):
    """dspi_eqx_* / dspi_sg_* validate before they touch a device (runs without a GPU)."""
    import ctypes as C
    h = C.c_void_p()
    dev = (C.c_int32 * 8)(0, 1, 2, 3, 4, 5, 6, 7)
    assert lib.dspi_eqx_create(None, None) == -22
    assert lib.dspi_eqx_create(C.byref(h), C.byref(api._EqxDesc(0, 1024, 10, 0, dev, 0))) == -22          # no device
    assert lib.dspi_eqx_create(C.byref(h), C.byref(api._EqxDesc(0, 1024, 10, 9, dev, 0))) == -22          # more than DSPI_MAX_DEVICES
    assert lib.dspi_eqx_create(C.byref(h), C.byref(api._EqxDesc(0, 0, 10, 2, dev, 0))) == -22             # no channels
    twice = (C.c_int32 * 8)(0, 1, 0, 3, 4, 5, 6, 7)
    assert lib.dspi_eqx_create(C.byref(h), C.byref(api._EqxDesc(0, 1024, 10, 3, twice, 0))) == -22
    assert b"twice" in lib.dspi_last_error()
    assert not h.value
    lo, hi = C.c_uint32(), C.c_uint32()
    assert lib.dspi_eqx_shard_range(1024, 0, 0, C.byref(lo), C.byref(hi)) == -22
    assert lib.dspi_eqx_shard_range(1024, 2, 2, C.byref(lo), C.byref(hi)) == -22
    assert lib.dspi_eqx_process_host(None, None, 16) == -22
    assert lib.dspi_eqx_process_root(None, None, 16, 16) == -22
    assert lib.dspi_sg_create(None, None, 0, None, 0, 1, 0) == -22
    assert lib.dspi_sg_process(None, None, 64, 16, 0) == -22
    assert lib.dspi_nccl_unique_id(None) == -22


@pytest.mark.parametrize("total,world", [(64, 1), (64, 2), (65, 2), (1000, 3), (65536, 8), (524288, 8), (130, 8), (1, 4)])
def test_shard_ranges_tile_the_channels_on_64_row_units(lib, total, world):
    """dspi_eqx_shard_range (the one sharding rule of dspi_eqx_*, dspi_sg_* and dspi_b200/sharding.py): contiguous, complete,
    every boundary but the last on a multiple of 64 channels, sizes differing by at most one unit."""
    import ctypes as C
    from dspi_b200 import sharding
    edges = []
    for k in range(world):
        lo, hi = C.c_uint32(), C.c_uint32()
        assert lib.dspi_eqx_shard_range(total, world, k, C.byref(lo), C.byref(hi)) == 0
        edges.append((lo.value, hi.value))
        assert (lo.value, hi.value) == tuple(sharding.shard_range(total, k, world))
    assert edges[0][0] == 0 and edges[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert all(lo % 64 == 0 for lo, _ in edges if lo < total)
    units = [-(-(hi - lo) // 64) for lo, hi in edges]
    assert max(units) - min(units) <= 1


def test_chain_entry_points_reject_bad_arguments(lib):
    """dspi_chain_* / dspi_chainq_* argument checks come before any device work (runs without a GPU)."""
    import ctypes as C
    h = C.c_void_p()
    for create, good_arith in ((lib.dspi_chain_create, 0), (lib.dspi_chainq_create, 2)):
        assert create(None, None) == -22
        assert create(C.byref(h), C.byref(api._ChainDesc(2 - good_arith, 16, 10, 0, 96))) == -22            # the other engine's arithmetic
        assert create(C.byref(h), C.byref(api._ChainDesc(good_arith, 0, 10, 0, 96))) == -22                 # no instances
        assert create(C.byref(h), C.byref(api._ChainDesc(good_arith, 16, 10, 0, 0))) == -22                 # no frames
        assert create(C.byref(h), C.byref(api._ChainDesc(good_arith, 16, 13, 0, 96))) == -22                # more bands than MAX_BANDS
        assert not h.value
    assert lib.dspi_chain_create(C.byref(h), C.byref(api._ChainDesc(0, 16, 7, 0, 96))) == -22                # the float chain runs channel_band_counts = 10
    assert lib.dspi_chain_set_params(None, 0, 1, None) == -22
    assert lib.dspi_chainq_set_params(None, 0, 1, None) == -22
    assert lib.dspi_chain_set_dynamics_device(None, 0, 1, None, C.c_float(48000.0)) == -22
    assert lib.dspi_chainq_set_dynamics_device(None, 0, 1, None, C.c_float(48000.0)) == -22
    assert lib.dspi_chain_set_preset_mute(None, 0, 1, None, 48000) == -22
    assert lib.dspi_chainq_set_preset_mute(None, 0, 1, None, 48000) == -22
    assert lib.dspi_chain_state_export(None, None, 0) == -22
    assert lib.dspi_chainq_state_export(None, None, 0) == -22
