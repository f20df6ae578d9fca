"""Preset slot images: dspi_preset_slot_collect / _apply against the reference's own
preset_save() / preset_load() (flash_storage.c compiled unmodified for both platforms against a RAM image of
the flash, oracle/ref_preset_shim.c) and against the committed fixture made from it."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from dspi_b200 import api, layouts as L
from tests.bulk_cases import wire_packet
from tests.orc import ORACLE_DIR
from tests.test_bulk_cpu import _same_state
from tests.util import load_golden

PLATFORMS = [L.PLATFORM_RP2350, L.PLATFORM_RP2040]
VERSION_OFFSET, INDEX_OFFSET, CRC_OFFSET, DATA_OFFSET = 4, 6, 8, 12


def _ref(platform):
    path = os.path.join(ORACLE_DIR, "_ref", "libdspi_ref_preset_%s.so" % ("rp2350" if platform == L.PLATFORM_RP2350 else "rp2040"))
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    lib = C.CDLL(path)
    assert lib.ref_preset_platform() == platform
    lib.ref_preset_load.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint8, C.c_float, C.c_void_p]
    return lib


def _state(platform, seed):
    st = api.bulk_state_defaults(platform)
    assert api.bulk_params_apply(wire_packet(platform, seed), st, exact_db=True) == 0
    return st


def _reseal(img):
    img[CRC_OFFSET:CRC_OFFSET + 4] = np.frombuffer(np.uint32(api.crc32(img[DATA_OFFSET:].tobytes())).tobytes(), np.uint8)
    return img


def test_crc32_is_the_zlib_polynomial():
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 2896):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert api.crc32(b) == zlib.crc32(b)


@pytest.mark.parametrize("platform", PLATFORMS)
def test_collect_matches_reference_preset_save(platform):
    ref = _ref(platform)
    n = api.preset_slot_size(platform)
    for seed in range(6):
        st = _state(platform, 400 + seed)
        slot = seed % 10
        sector = np.zeros(4096, np.uint8)
        assert ref.ref_preset_save(st.ctypes.data_as(C.c_void_p), slot, sector.ctypes.data_as(C.c_void_p)) == 0
        got = api.preset_slot_collect(st, slot)
        assert got.size == n and np.array_equal(got, sector[:n])
        assert len(set(sector[n:].tolist())) <= 2          # the rest of the sector is page padding / erased flash


@pytest.mark.parametrize("platform", PLATFORMS)
@pytest.mark.parametrize("version", [7, 9, 10, 11, 12])
@pytest.mark.parametrize("mode", [0, 1])
def test_apply_matches_reference_preset_load(platform, version, mode):
    ref = _ref(platform)
    for seed in range(5):
        src = _state(platform, 500 + seed)
        slot = (seed * 3) % 10
        img = api.preset_slot_collect(src, slot)
        img[VERSION_OFFSET:VERSION_OFFSET + 2] = np.frombuffer(np.uint16(version).tobytes(), np.uint8)   # an image written by older firmware
        _reseal(img)
        a = _state(platform, 600 + seed)                   # the device state before the load
        b = a.copy()
        rc = api.preset_slot_apply(img, slot, a, mode, -7.5)
        assert rc == ref.ref_preset_load(img.ctypes.data, img.size, slot, mode, -7.5, b.ctypes.data) == 0
        _same_state(a, b)


@pytest.mark.parametrize("platform", PLATFORMS)
def test_rejected_images(platform):
    ref = _ref(platform)
    src = _state(platform, 700)
    good = api.preset_slot_collect(src, 4)
    for what in ("crc", "magic", "index", "payload"):
        img = good.copy()
        slot = 4
        if what == "crc":
            img[CRC_OFFSET] ^= 1
        elif what == "magic":
            img[0] ^= 0x40
        elif what == "index":
            slot = 5
        else:
            img[200] ^= 0x10
        a = _state(platform, 701)
        b, before = a.copy(), a.copy()
        rc = api.preset_slot_apply(img, slot, a)
        assert rc == ref.ref_preset_load(img.ctypes.data, img.size, slot, 0, 0.0, b.ctypes.data) == 3, what
        _same_state(a, before)
        _same_state(b, before)


@pytest.mark.parametrize("platform", PLATFORMS)
def test_round_trip_and_fixture(platform):
    st = _state(platform, 800)
    img = api.preset_slot_collect(st, 2)
    again = api.bulk_state_defaults(platform)
    assert api.preset_slot_apply(img, 2, again, 1, 0.0) == 0
    for name in ("recipes", "crossfeed", "leveller", "channel_delays_ms", "preamp_db", "master_volume_db", "bypass_master_eq"):
        assert np.ascontiguousarray(again[name]).tobytes() == np.ascontiguousarray(st[name]).tobytes(), name
    g = load_golden("preset.npz")
    key = "rp2350" if platform == L.PLATFORM_RP2350 else "rp2040"
    states = np.frombuffer(np.ascontiguousarray(g[f"{key}_state"]).tobytes(), L.BULK_STATE).copy()
    for i in range(len(states)):
        assert np.array_equal(api.preset_slot_collect(states[i:i + 1], int(g[f"{key}_slot"][i])), g[f"{key}_image"][i])
        out = api.bulk_state_defaults(platform)
        assert api.preset_slot_apply(g[f"{key}_image"][i], int(g[f"{key}_slot"][i]), out, 1, 0.0) == 0
        after = np.frombuffer(np.ascontiguousarray(g[f"{key}_loaded"]).tobytes(), L.BULK_STATE).copy()
        _same_state(out, after[i:i + 1])
