"""Pins the restated orchestrator glue and the restated delta-sigma core (oracle/liborc.so) against the
reference's OWN code: usb_audio.c's process_audio_packet() and pdm_generator.c's modulator loop, compiled
unmodified on the host (oracle/ref_chain_shim.c, oracle/ref_pdm_shim.c -> oracle/_ref/).

Every case runs the same instance record through orc_*_chain_packet and through the reference and compares
every byte of state (filters, loudness, crossfeed, leveller incl. look-ahead, delay rings + write index,
meters, clip flags, envelope) plus the S/PDIF words and the PDM bit stream."""
import ctypes as C

import numpy as np
import pytest

from tests.chain_cases import chain_params, chain_params_q28, pcm_bytes, quirk_cases
from tests.orc import (OrcChainF32, OrcChainQ28, RefChain, RefPdm, arm_mute_envelope, make_orc_chain, make_orc_chain_q28)

pytestmark = pytest.mark.skipif(not (RefChain.available() and RefPdm.available()),
                                reason="oracle/_ref chain / PDM builds missing (need /root/reference)")

FLAVOURS = ["f32s", "f32f", "q28"]


@pytest.fixture(scope="module")
def ref_chains():
    return {k: RefChain(k) for k in FLAVOURS}


@pytest.fixture(scope="module")
def ref_pdm():
    return RefPdm()


def _orc_run(oracle, flavour, chain, pcm, bit_depth, n_packets, fpp):
    F = n_packets * fpp
    bpf = 6 if bit_depth == 24 else 4
    pairs = 2 if flavour == "q28" else 4
    spdif = np.zeros((pairs, F, 2), np.int32)
    pdm = np.zeros((F, 8), np.uint32)
    data = np.ascontiguousarray(pcm)
    fn = getattr(oracle.lib, f"orc_{flavour}_chain_packet")
    for p in range(n_packets):
        n = fn(C.addressof(chain), data.ctypes.data + p * fpp * bpf, fpp * bpf, bit_depth,
               spdif.ctypes.data + p * fpp * 8, F * 2, pdm.ctypes.data + p * fpp * 32)
        assert n == fpp
    return spdif, pdm


def _first_field(struct_type, offset):
    name = None
    for f in struct_type._fields_:
        if getattr(struct_type, f[0]).offset <= offset:
            name = f[0]
    return name


def compare_instance(oracle, ref_chain, ref_pdm, flavour, params, biquads, fs, pcm, bit_depth, n_packets, fpp, prepare=None):
    q = flavour == "q28"
    make, T = (make_orc_chain_q28, OrcChainQ28) if q else (make_orc_chain, OrcChainF32)
    a, b = make(oracle, params, biquads), make(oracle, params, biquads)
    if prepare:
        prepare(a)
        prepare(b)
    sa, pa = _orc_run(oracle, flavour, a, pcm, bit_depth, n_packets, fpp)
    sb, sub = ref_chain.run(b, fs, pcm, bit_depth, n_packets, fpp)
    assert np.array_equal(sa, sb), "S/PDIF words differ"
    sub_on = bool(params["matrix"]["outputs"][4 if q else 8]["enabled"])
    assert len(sub) == (n_packets * fpp if sub_on else 0)
    if sub_on:
        words, rng_after = ref_pdm.run(sub)
        assert np.array_equal(pa, words), "PDM bit stream differs"
        assert rng_after == a.pdm.rng
    b.pdm = a.pdm                       # the reference's modulator state lives in locals of its loop
    ba, bb = np.frombuffer(bytes(a), np.uint8), np.frombuffer(bytes(b), np.uint8)
    d = np.nonzero(ba != bb)[0]
    assert len(d) == 0, f"state differs from byte {d[0]} ({_first_field(T, d[0])}), {len(d)} bytes"
    return a


@pytest.mark.parametrize("fpp", [96, 95, 192, 1])
@pytest.mark.parametrize("bit_depth", [16, 24])
@pytest.mark.parametrize("flavour", FLAVOURS)
def test_chain_restatement_matches_reference(oracle, ref_chains, ref_pdm, flavour, bit_depth, fpp):
    fs, N = 96000.0, 10
    n_packets = 6 if fpp > 1 else 40
    oracle.set_x86_cvt(1)               # the reference objects are x86 code: cvttss2si on overflow
    try:
        if flavour == "q28":
            P, bq = chain_params_q28(oracle, N, fs, 6)
        else:
            P, bq = chain_params(oracle, N, fs, 5)
        pcm = pcm_bytes(N, n_packets * fpp, bit_depth, 3)
        for i in range(N):
            if float(P[i]["preset_mute_gain"]) not in (0.0, 1.0):
                P[i]["preset_mute_gain"] = 1.0
            compare_instance(oracle, ref_chains[flavour], ref_pdm, flavour, P[i], bq[i], fs, pcm[i], bit_depth, n_packets, fpp)
    finally:
        oracle.set_x86_cvt(0)


@pytest.mark.parametrize("fs", [44100.0, 48000.0])
@pytest.mark.parametrize("flavour", FLAVOURS)
def test_chain_other_sample_rates(oracle, ref_chains, ref_pdm, flavour, fs):
    N, n_packets, fpp = 4, 8, 48 if fs == 48000.0 else 45
    oracle.set_x86_cvt(1)
    try:
        P, bq = chain_params_q28(oracle, N, fs, 11) if flavour == "q28" else chain_params(oracle, N, fs, 12)
        pcm = pcm_bytes(N, n_packets * fpp, 16, 4)
        for i in range(N):
            P[i]["preset_mute_gain"] = 1.0
            compare_instance(oracle, ref_chains[flavour], ref_pdm, flavour, P[i], bq[i], fs, pcm[i], 16, n_packets, fpp)
    finally:
        oracle.set_x86_cvt(0)


@pytest.mark.parametrize("case", ["full_volume_polarity", "delay_equals_max", "delay_max_minus_one", "clipping_hot_input",
                                  "pdm_saturation", "host_muted", "everything_off", "sub_only"])
@pytest.mark.parametrize("flavour", FLAVOURS)
def test_chain_quirks_match_reference(oracle, ref_chains, ref_pdm, flavour, case):
    """SURVEY §8 quirks 1, 4, 7 and the delay alias, against the reference's compiled code."""
    fs, n_packets, fpp = 96000.0, 50, 96            # 4800 frames: longer than both delay rings
    oracle.set_x86_cvt(1)
    try:
        P, bq, pcm, bit_depth = quirk_cases(oracle, flavour, case, fs, n_packets * fpp)
        got = compare_instance(oracle, ref_chains[flavour], ref_pdm, flavour, P, bq, fs, pcm, bit_depth, n_packets, fpp)
        if case == "clipping_hot_input":
            assert got.clip_flags != 0, "the case must actually set clip flags"
        if case == "delay_equals_max":
            assert got.delay_samples[0] == got.max_delay
    finally:
        oracle.set_x86_cvt(0)


@pytest.mark.parametrize("fs,fpp", [(96000.0, 96), (48000.0, 48), (44100.0, 45)])
@pytest.mark.parametrize("flavour", FLAVOURS)
def test_preset_mute_envelope_matches_reference(oracle, ref_chains, ref_pdm, flavour, fs, fpp):
    """update_preset_mute_envelope() (usb_audio.c:466-498): fade out, hold while the counter runs, fade in."""
    n_packets = 40
    P, bq = chain_params_q28(oracle, 1, fs, 21) if flavour == "q28" else chain_params(oracle, 1, fs, 22, uniform=True)
    P[0]["host_mute"] = 0
    pcm = pcm_bytes(1, n_packets * fpp, 24, 5)
    gains = []

    def prepare(c):
        arm_mute_envelope(c, fs)

    a = compare_instance(oracle, ref_chains[flavour], ref_pdm, flavour, P[0], bq[0], fs, pcm[0], 24, n_packets, fpp, prepare)
    assert a.preset_loading == 0 and a.preset_mute_smooth_gain == 1.0      # the fade completed inside the run
    # the trajectory itself: down to 0 in <= 8 ms, back up after the 10 ms hold
    ld, cnt, g = C.c_uint8(1), C.c_uint32(max(512, (int(fs) * 10 + 999) // 1000)), C.c_float(1.0)
    for _ in range(n_packets):
        gains.append(oracle.lib.orc_mute_envelope(C.byref(ld), C.byref(cnt), C.byref(g), fpp, int(fs)))
    assert min(gains) == 0.0 and gains[-1] == 1.0 and 0.0 < gains[0] < 1.0


def test_pdm_restatement_matches_reference_loop(oracle, ref_pdm):
    """orc_pdm_modulate vs the reference's own loop: fade-in, clipping at +-29500, dither, leaky integrators."""
    rng = np.random.default_rng(1)
    for trial, scale in enumerate([1 << 24, 1 << 27, 1 << 29, 1 << 30]):
        n = 2600
        x = (rng.standard_normal(n) * scale).clip(-2**31, 2**31 - 1).astype(np.int32)
        if trial == 3:
            x[::7] = np.int32(-2**31)
            x[3::11] = np.int32(2**31 - 1)
            x[5::13] = 0
        st = np.zeros(1, dtype=np.dtype([(k, "<i4") for k in ("err1", "err2", "x1", "x2", "y1", "y2", "err_acc")] + [("rng", "<u4"), ("fade", "<u4")]))
        st["rng"] = 123456789
        want = oracle.pdm(st, x)
        got, rng_after = ref_pdm.run(x)
        assert np.array_equal(got, want)
        assert rng_after == int(st["rng"][0]) and int(st["fade"][0]) == 1024


def test_silence_is_the_idle_pattern_after_fade(ref_pdm):
    """pdm_generator.c:128-130: 50 % duty cycle; the modulator's own output for digital silence has equal ones and zeros on average."""
    words, _ = ref_pdm.run(np.zeros(512, np.int32))
    ones = np.unpackbits(words.view(np.uint8)).sum()
    assert abs(int(ones) - 512 * 128) < 512 * 128 * 0.01


def test_host_volume_and_gain_helpers_match_reference(oracle, ref_chains):
    """audio_set_volume(), update_preamp(), update_master_volume() (usb_audio.c:244-269, 428-440)."""
    from dspi_b200 import api
    r = ref_chains["f32s"].lib
    for v in list(range(-32768, 32768, 37)) + [0, -1, -256, -255, -15360, -15361, 255, 256]:
        i1, i2 = C.c_uint8(), C.c_uint8()
        assert oracle.lib.orc_host_vol_mul(v, C.byref(i1)) == r.ref_host_vol_mul(v, C.byref(i2)) and i1.value == i2.value
        vm, row = api.host_volume(v)
        assert vm == r.ref_host_vol_mul(v, C.byref(i2)) and row == i2.value
    assert r.ref_host_vol_mul(0, None) == -32768               # quirk 1: 0 dB -> int16 wraps to -32768
    for db in [-60.0, -20.0, -6.0, -0.5, 0.0, 0.1, 3.0, 12.0, 17.9]:
        lin, q = C.c_float(), C.c_int32()
        r.ref_preamp(db, C.byref(lin), C.byref(q))
        assert api.preamp(db) == (lin.value, q.value)
    for db in [-128.0, -127.0, -60.0, -20.0, -3.0, 0.0, 5.0, -200.0]:
        lin, q = C.c_float(), C.c_int32()
        r.ref_master_volume(db, C.byref(lin), C.byref(q))
        assert api.master_volume(db) == (lin.value, q.value)
