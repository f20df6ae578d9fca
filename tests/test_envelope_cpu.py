"""Host side of the preset-mute envelope (usb_audio.c:456-498): dspi_preset_mute_arm / dspi_preset_mute_step of the product's
plain-C host library against the oracle's restatement, which tests/test_chain_vs_ref_cpu.py pins to the compiled usb_audio.c.
No GPU involved: these are the functions a host uses to predict or mirror the per-packet gain the chain engines apply."""
import ctypes as C

import numpy as np
import pytest

from dspi_b200 import api, layouts as L


@pytest.mark.parametrize("fs,fpp", [(96000, 96), (48000, 48), (44100, 45), (44100, 44), (48000, 1), (96000, 192), (8000, 8)])
def test_envelope_step_matches_oracle(oracle, fs, fpp):
    m = np.zeros(1, L.PRESET_MUTE)
    m["smooth_gain"] = 1.0
    api.lib().dspi_preset_mute_arm(m.ctypes.data_as(C.c_void_p), fs)
    assert int(m["loading"][0]) == 1 and int(m["counter"][0]) > 0           # PRESET_MUTE_SAMPLES at this rate (usb_audio.c:456-464)
    ld, cnt, g = C.c_uint8(1), C.c_uint32(int(m["counter"][0])), C.c_float(1.0)
    seen_zero = seen_back_to_one = False
    for _ in range(400):
        a = api.lib().dspi_preset_mute_step(m.ctypes.data_as(C.c_void_p), fpp, fs)
        b = oracle.lib.orc_mute_envelope(C.byref(ld), C.byref(cnt), C.byref(g), fpp, fs)
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
        assert int(m["counter"][0]) == cnt.value and int(m["loading"][0]) == ld.value
        assert np.float32(m["smooth_gain"][0]).view(np.uint32) == np.float32(g.value).view(np.uint32)
        seen_zero |= a == 0.0
        seen_back_to_one |= seen_zero and a == 1.0
    if fpp * 400 > 2 * int(fs * 0.3):                                       # long enough to fade out, hold and fade back in
        assert seen_zero and seen_back_to_one


def test_envelope_idle_state_is_a_fixed_point(oracle):
    m = np.zeros(1, L.PRESET_MUTE)
    m["smooth_gain"] = 1.0
    for _ in range(10):
        assert api.lib().dspi_preset_mute_step(m.ctypes.data_as(C.c_void_p), 96, 96000) == 1.0
    assert int(m["loading"][0]) == 0 and int(m["counter"][0]) == 0
