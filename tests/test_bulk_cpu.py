"""Bulk parameter ingest (SURVEY.md §8 b / f-1, host side): the product's dspi_bulk_params_* against the
reference's own bulk_params.c compiled for both platforms (oracle/_ref/libdspi_ref_bulk_*.so), against
the committed fixture made from it, and the derived chain records against the host parameter API."""
import ctypes as C
import os

import numpy as np
import pytest

from dspi_b200 import api, layouts as L
from tests.bulk_cases import dims, wire_packet
from tests.orc import ORACLE_DIR
from tests.util import load_golden

PLATFORMS = [L.PLATFORM_RP2350, L.PLATFORM_RP2040]
REF_SO = {L.PLATFORM_RP2350: "libdspi_ref_bulk_rp2350.so", L.PLATFORM_RP2040: "libdspi_ref_bulk_rp2040.so"}


def _ref(platform):
    path = os.path.join(ORACLE_DIR, "_ref", REF_SO[platform])
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    lib = C.CDLL(path)
    lib.ref_bulk_wire_size.restype = C.c_size_t
    assert lib.ref_bulk_wire_size() == L.WIRE_BULK.itemsize and lib.ref_bulk_platform() == platform
    return lib


def _same_state(a, b):
    """`a` from the product, `b` from the x86 build of the reference.  One documented difference: a float -> int32
    conversion that leaves the range (preamp above +18 dB in Q28) saturates on the firmware's ARM cores and in the
    product, while the x86 reference object returns INT_MIN."""
    b = b.copy()
    over = b["preamp_linear"] * np.float32(2.0 ** 28) >= np.float32(2.0 ** 31)
    b["preamp_q28"][over & (b["preamp_q28"] == np.iinfo(np.int32).min)] = np.iinfo(np.int32).max
    for name in L.BULK_STATE.names:
        x, y = np.ascontiguousarray(a[name]), np.ascontiguousarray(b[name])
        if x.dtype.names:                                   # records with padding: compare field by field
            for f in x.dtype.names:
                assert np.array_equal(np.ascontiguousarray(x[f]).view(np.uint8), np.ascontiguousarray(y[f]).view(np.uint8)), (name, f)
        else:
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), name


@pytest.mark.parametrize("platform", PLATFORMS)
@pytest.mark.parametrize("version", [2, 3, 4, 5, 6])
def test_apply_matches_reference(platform, version):
    ref = _ref(platform)
    for seed in range(12):
        w = wire_packet(platform, seed, version)
        a = api.bulk_state_defaults(platform)
        if seed % 3 == 1:                                   # a state that already carries an older configuration
            assert api.bulk_params_apply(wire_packet(platform, 1000 + seed, 6), a) == 0
        b = a.copy()
        rc = api.bulk_params_apply(w, a)
        assert rc == ref.ref_bulk_apply(w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)) == 0
        _same_state(a, b)


@pytest.mark.parametrize("platform", PLATFORMS)
def test_error_codes_match_reference(platform):
    ref = _ref(platform)
    nc, no = dims(platform)
    cases = [("format_version", 1), ("format_version", 7), ("platform_id", 1 - platform), ("num_channels", nc + 1),
             ("num_output_channels", no - 1), ("payload_length", 100), ("payload_length", 2897)]
    for field, value in cases:
        w = wire_packet(platform, 3)
        w["header"][0][field] = value
        a = api.bulk_state_defaults(platform)
        b = a.copy()
        rc = api.bulk_params_apply(w, a)
        assert rc == ref.ref_bulk_apply(w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
        assert rc in (-1, -2, -3, -4), (field, value, rc)
        _same_state(a, api.bulk_state_defaults(platform))   # a rejected packet changes nothing


@pytest.mark.parametrize("platform", PLATFORMS)
def test_collect_matches_reference_and_round_trips(platform):
    ref = _ref(platform)
    for seed in range(8):
        st = api.bulk_state_defaults(platform)
        assert api.bulk_params_apply(wire_packet(platform, 50 + seed), st) == 0
        got = api.bulk_params_collect(st)
        want = np.zeros(1, L.WIRE_BULK)
        ref.ref_bulk_collect(st.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
        assert got.tobytes() == want.tobytes()
        again = api.bulk_state_defaults(platform)
        assert api.bulk_params_apply(got, again) == 0
        _same_state(again, st)


@pytest.mark.parametrize("platform", PLATFORMS)
def test_fixture(platform):
    """Runs anywhere (the GPU box has no reference): packets and the states the reference build left."""
    g = load_golden("bulk.npz")
    key = "rp2350" if platform == L.PLATFORM_RP2350 else "rp2040"
    wires = np.frombuffer(np.ascontiguousarray(g[f"{key}_wire"]).tobytes(), L.WIRE_BULK).copy()
    states = np.frombuffer(np.ascontiguousarray(g[f"{key}_state"]).tobytes(), L.BULK_STATE).copy()
    rcs = g[f"{key}_rc"]
    for i in range(len(wires)):
        st = api.bulk_state_defaults(platform)
        assert api.bulk_params_apply(wires[i:i + 1], st) == int(rcs[i])
        _same_state(st, states[i:i + 1])


def test_exact_db_option():
    w = wire_packet(L.PLATFORM_RP2350, 5)
    w["outputs"][0]["gain_db"][0] = -30.0                    # the Taylor series is far off here
    a, b = api.bulk_state_defaults(), api.bulk_state_defaults()
    assert api.bulk_params_apply(w, a) == 0 and api.bulk_params_apply(w, b, exact_db=True) == 0
    assert abs(float(b["outputs"][0, 0]["gain_linear"]) - 10 ** (-30 / 20)) < 1e-7
    assert abs(float(a["outputs"][0, 0]["gain_linear"]) - float(b["outputs"][0, 0]["gain_linear"])) > 0.1


@pytest.mark.parametrize("platform", PLATFORMS)
def test_derived_chain_records(oracle, platform):
    """dspi_bulk_state_to_chain_* == the individual host parameter functions (each pinned against the reference
    elsewhere) applied the way the main loop does after an apply."""
    fs, vol = 96000.0, -17 * 256
    q28 = platform == L.PLATFORM_RP2040
    nc, no = dims(platform)
    for seed in range(4):
        st = api.bulk_state_defaults(platform)
        assert api.bulk_params_apply(wire_packet(platform, 80 + seed), st) == 0
        P, bq = api.bulk_state_to_chain(st, fs, vol, host_mute=seed == 2)
        s, p = st[0], P[0]
        vol_mul, row = api.host_volume(vol)
        assert p["host_vol_mul"] == vol_mul and p["host_mute"] == (seed == 2)
        assert p["bypass_master_eq"] == s["bypass_master_eq"] and p["loudness_enabled"] == s["loudness_enabled"]
        assert p["crossfeed_enabled"] == s["crossfeed"]["enabled"] and p["leveller_enabled"] == s["leveller"]["enabled"]
        assert p["leveller_lookahead"] == s["leveller"]["lookahead"]
        xf, lv = s["crossfeed"], s["leveller"]
        if q28:
            assert p["master_volume_q15"] == s["master_volume_q15"] and np.array_equal(p["preamp_q28"], s["preamp_q28"])
            table = api.loudness_table_q28(fs, float(s["loudness_ref_spl"]), float(s["loudness_intensity_pct"]))
            want_xf = api.crossfeed_coefficients_q28(fs, bool(xf["enabled"]), bool(xf["itd_enabled"]), int(xf["preset"]), float(xf["custom_fc"]),
                                                     float(xf["custom_feed_db"]))
        else:
            assert p["master_volume_linear"] == s["master_volume_linear"] and np.array_equal(p["preamp_linear"], s["preamp_linear"])
            table = api.loudness_table(fs, float(s["loudness_ref_spl"]), float(s["loudness_intensity_pct"]))
            want_xf = api.crossfeed_coefficients(fs, bool(xf["enabled"]), bool(xf["itd_enabled"]), int(xf["preset"]), float(xf["custom_fc"]),
                                                 float(xf["custom_feed_db"]))
        assert p["loudness"].tobytes() == table[row].tobytes()
        assert p["crossfeed"].tobytes() == np.asarray(want_xf).tobytes()
        want_lv = api.leveller_coefficients(fs, float(lv["amount"]), int(lv["speed"]), float(lv["max_gain_db"]), float(lv["gate_threshold_db"]))
        assert p["leveller"].tobytes() == np.asarray(want_lv).tobytes()
        max_delay = L.CHAINQ_MAX_DELAY if q28 else L.CHAIN_MAX_DELAY
        for o in range(no):
            oc = p["matrix"]["outputs"][o]
            assert oc["delay_samples"] == oracle.lib.orc_delay_samples(float(s["channel_delays_ms"][2 + o]), fs, int(o == no - 1), max_delay)
            assert oc["gain_linear"] == s["outputs"][o]["gain_linear"] and oc["mute"] == s["outputs"][o]["mute"]
            for side in range(2):
                assert p["matrix"]["crosspoints"][side, o].tobytes() == s["crosspoints"][side, o].tobytes()
        want_bq = api.compute_coefficients(s["recipes"][:nc].copy(), q28=q28, fs=fs)
        assert bq[0].tobytes() == want_bq.tobytes()
