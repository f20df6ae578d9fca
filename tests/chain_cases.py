"""Seeded configurations for the whole-chain parity tests (BASELINE config 3 shape, shrunk)."""
import ctypes as C

import numpy as np

from dspi_b200 import api, layouts as L, workloads as W


def chain_params(oracle, N, fs, seed, leveller=True, uniform=False):
    """CHAIN_PARAMS_F32 [N] + biquads [N, 11, 12]: every stage exercised, per-instance variety."""
    rng = np.random.default_rng(seed)
    P = np.zeros(N, L.CHAIN_PARAMS_F32)
    loud_tab = np.zeros((L.LOUD_STEPS, 2), L.LOUD_F32)
    oracle.lib.orc_loud_table_f32(loud_tab.ctypes.data, 83.0, 100.0, fs)
    for i in range(N):
        p = P[i]
        u = rng.random(16)
        vol_8_8 = int(-40 * 256 * u[0])                       # host volume 0 .. -40 dB
        idx = C.c_uint8()
        p["host_vol_mul"] = oracle.lib.orc_host_vol_mul(vol_8_8, C.byref(idx))
        p["host_mute"] = 1 if (not uniform and u[1] < 0.05) else 0
        p["preset_mute_gain"] = 1.0
        p["master_volume_linear"] = np.float32(10.0 ** (-6.0 * u[2] / 20.0))
        p["preamp_linear"] = [np.float32(10.0 ** ((-3 + 6 * u[3]) / 20.0)), np.float32(10.0 ** ((-3 + 6 * u[4]) / 20.0))]
        p["bypass_master_eq"] = 1 if (not uniform and u[5] < 0.15) else 0
        p["loudness_enabled"] = 1 if (uniform or u[6] < 0.7) else 0
        p["loudness"] = loud_tab[idx.value]
        xcfg = (C.c_uint8 * 12)()
        xn = np.frombuffer(xcfg, np.uint8)
        xn[0], xn[1], xn[2] = 1, (1 if u[7] < 0.8 else 0), int(u[8] * 4) % 4
        xn[4:8] = np.frombuffer(np.float32(500 + 1500 * u[9]).tobytes(), np.uint8)
        xn[8:12] = np.frombuffer(np.float32(15 * u[10]).tobytes(), np.uint8)
        xst = np.zeros(1, L.XFEED_F32)
        oracle.lib.orc_xfeed_coeffs_f32(xst.ctypes.data, C.addressof(xcfg), fs)
        p["crossfeed"] = xst[0]
        p["crossfeed_enabled"] = 1 if (uniform or u[11] < 0.7) else 0
        lcfg = np.zeros(24, np.uint8)
        lcfg[0] = 1
        lcfg[4:8] = np.frombuffer(np.float32(100 * u[12]).tobytes(), np.uint8)
        lcfg[8] = int(u[13] * 3) % 3
        lcfg[12:16] = np.frombuffer(np.float32(15.0).tobytes(), np.uint8)
        lcfg[16] = 1
        lcfg[20:24] = np.frombuffer(np.float32(-96.0).tobytes(), np.uint8)
        lc = np.zeros(1, L.LEV_COEFFS)
        oracle.lib.orc_lev_coeffs_compute(lc.ctypes.data, lcfg.ctypes.data, fs)
        p["leveller"] = lc[0]
        p["leveller_enabled"] = 1 if (leveller and (uniform or u[14] < 0.6)) else 0
        p["leveller_lookahead"] = 1 if (uniform or u[15] < 0.5) else 0
        m = p["matrix"]
        v = rng.random((9, 8))
        for o in range(9):
            oc = m["outputs"][o]
            oc["enabled"] = 1 if (uniform or v[o, 0] < 0.85) else 0
            oc["mute"] = 1 if (not uniform and v[o, 1] < 0.1) else 0
            oc["gain_db"] = np.float32(-6 * v[o, 2])
            oc["gain_linear"] = np.float32(10.0 ** (float(oc["gain_db"]) / 20.0))
            oc["delay_ms"] = np.float32(40.0 * v[o, 3]) if v[o, 4] < 0.7 else np.float32(0.0)
            oc["delay_samples"] = api.delay_samples(float(oc["delay_ms"]), fs, o == 8)
            # L -> odd outputs, R -> even, (L+R)/2 -> sub, plus some random routes and inversions
            for side in range(2):
                x = m["crosspoints"][side, o]
                on = (o == 8) or (o % 2 == side) or (not uniform and v[o, 5 + side] < 0.2)
                x["enabled"] = 1 if on else 0
                x["phase_invert"] = 1 if (not uniform and v[o, 7] < 0.2 and side == 1) else 0
                x["gain_db"] = np.float32(-6.0 if o == 8 else 0.0)
                x["gain_linear"] = np.float32(0.5 if o == 8 else 1.0)
    # EQ recipes: per role; per instance variety unless `uniform`
    bq = np.zeros((N, L.CHAIN_EQ_CHANNELS, L.MAX_BANDS), L.BIQUAD_F32)
    for i in range(N):
        variant = "B" if uniform else ("mixed" if i % 3 == 0 else ("A" if i % 3 == 1 else "B"))
        params = W.eq_params(variant, L.CHAIN_EQ_CHANNELS, fs=fs, seed=seed + (0 if uniform else i))
        bq[i] = api.compute_coefficients(params, q28=False, fs=fs)
        if not uniform and i % 7 == 3:
            bq[i, 4]["bypass"] = 1            # a fully flat output EQ (channel_bypassed)
    return P, bq


def pcm_bytes(N, F, bit_depth, seed):
    rng = np.random.default_rng(seed)
    if bit_depth == 16:
        s = (rng.integers(-20000, 20000, (N, F, 2))).astype("<i2")
        return s.view(np.uint8).reshape(N, F * 4)
    s = rng.integers(-(1 << 22), 1 << 22, (N, F, 2)).astype(np.int32)
    b = np.zeros((N, F, 2, 3), np.uint8)
    b[..., 0] = s & 0xFF
    b[..., 1] = (s >> 8) & 0xFF
    b[..., 2] = (s >> 16) & 0xFF
    return b.reshape(N, F * 6)


def chain_params_q28(oracle, N, fs, seed, leveller=True):
    """CHAIN_PARAMS_Q28 [N] + BIQUAD_Q28 [N, 7, 12] (RP2040 shape), every stage exercised."""
    rng = np.random.default_rng(seed)
    P = np.zeros(N, L.CHAIN_PARAMS_Q28)
    loud_tab = api.loudness_table_q28(fs, 83.0, 100.0)
    for i in range(N):
        p = P[i]
        u = rng.random(16)
        vol_mul, row = api.host_volume(int(-40 * 256 * u[0]))
        p["host_vol_mul"] = vol_mul
        p["host_mute"] = 1 if u[1] < 0.05 else 0
        p["preset_mute_gain"] = 1.0 if u[1] > 0.3 else np.float32(u[1] * 3)
        p["master_volume_q15"] = int(10.0 ** (-6.0 * u[2] / 20.0) * 32768.0)
        p["preamp_q28"] = [int(10.0 ** ((-3 + 6 * u[3]) / 20.0) * (1 << 28)), int(10.0 ** ((-3 + 6 * u[4]) / 20.0) * (1 << 28))]
        p["bypass_master_eq"] = 1 if u[5] < 0.15 else 0
        p["loudness_enabled"] = 1 if u[6] < 0.7 else 0
        p["loudness"] = loud_tab[row]
        p["crossfeed"] = api.crossfeed_coefficients_q28(fs, True, u[7] < 0.8, int(u[8] * 4) % 4, 500 + 1500 * u[9], 15 * u[10])
        p["crossfeed_enabled"] = 1 if u[11] < 0.7 else 0
        p["leveller"] = api.leveller_coefficients(fs, 100 * u[12], int(u[13] * 3) % 3, 15.0, -96.0)
        p["leveller_enabled"] = 1 if (leveller and u[14] < 0.6) else 0
        p["leveller_lookahead"] = 1 if u[15] < 0.5 else 0
        m = p["matrix"]
        v = rng.random((5, 8))
        for o in range(5):
            oc = m["outputs"][o]
            oc["enabled"] = 1 if v[o, 0] < 0.85 else 0
            oc["mute"] = 1 if v[o, 1] < 0.1 else 0
            oc["gain_db"] = np.float32(-6 * v[o, 2])
            oc["gain_linear"] = np.float32(10.0 ** (float(oc["gain_db"]) / 20.0))
            oc["delay_ms"] = np.float32(20.0 * v[o, 3]) if v[o, 4] < 0.7 else np.float32(0.0)
            ds = oracle.lib.orc_delay_samples(float(oc["delay_ms"]), fs, 1 if o == 4 else 0, 2048)
            oc["delay_samples"] = ds
            for side in range(2):
                x = m["crosspoints"][side, o]
                x["enabled"] = 1 if ((o == 4) or (o % 2 == side) or v[o, 5 + side] < 0.2) else 0
                x["phase_invert"] = 1 if (v[o, 7] < 0.2 and side == 1) else 0
                x["gain_db"] = np.float32(-6.0 if o == 4 else 0.0)
                x["gain_linear"] = np.float32(0.5 if o == 4 else 1.0)
    bq = np.zeros((N, L.CHAINQ_EQ_CHANNELS, L.MAX_BANDS), L.BIQUAD_Q28)
    for i in range(N):
        variant = "mixed" if i % 3 == 0 else ("A" if i % 3 == 1 else "B")
        bq[i] = api.compute_coefficients(W.eq_params(variant, L.CHAINQ_EQ_CHANNELS, fs=fs, seed=seed + i), q28=True, fs=fs)
        if i % 7 == 3:
            bq[i, 4]["bypass"] = 1
    return P, bq


def pcm_bytes_full_scale(F, bit_depth, seed):
    """One instance, every code of the format reachable (incl. the most negative one)."""
    rng = np.random.default_rng(seed)
    if bit_depth == 16:
        s = rng.integers(-32768, 32768, (F, 2)).astype("<i2")
        s[::17, 0] = -32768
        s[5::19, 1] = 32767
        return s.view(np.uint8).reshape(F * 4)
    s = rng.integers(-(1 << 23), 1 << 23, (F, 2)).astype(np.int32)
    s[::17, 0] = -(1 << 23)
    s[5::19, 1] = (1 << 23) - 1
    b = np.zeros((F, 2, 3), np.uint8)
    b[..., 0] = s & 0xFF
    b[..., 1] = (s >> 8) & 0xFF
    b[..., 2] = (s >> 16) & 0xFF
    return b.reshape(F * 6)


QUIRK_CASES = ["full_volume_polarity", "delay_equals_max", "delay_max_minus_one", "clipping_hot_input", "pdm_saturation",
               "host_muted", "everything_off", "sub_only"]


def quirk_cases(oracle, flavour, case, fs, F):
    """(params record, biquads, pcm bytes, bit depth) of one instance that exercises one documented quirk
    (SURVEY §8 'quirks', VERDICT r1 'quirks not exercised')."""
    q = flavour == "q28"
    if q:
        P, bq = chain_params_q28(oracle, 1, fs, 31)
        n_out, max_delay = 5, 2048
    else:
        P, bq = chain_params(oracle, 1, fs, 32, uniform=True)
        n_out, max_delay = 9, 4096
    p = P[0]
    m = p["matrix"]
    p["host_mute"] = 0
    p["preset_mute_gain"] = 1.0
    for o in range(n_out):
        m["outputs"][o]["enabled"] = 1
        m["outputs"][o]["mute"] = 0
    bit_depth = 24
    pcm = pcm_bytes(1, F, bit_depth, 9)[0]

    def set_master(db):
        lin, q15 = api.master_volume(db)
        if q:
            p["master_volume_q15"] = q15
        else:
            p["master_volume_linear"] = lin

    def set_preamp(db):
        lin, q28 = api.preamp(db)
        if q:
            p["preamp_q28"] = [q28, q28]
        else:
            p["preamp_linear"] = [lin, lin]

    if case == "full_volume_polarity":
        # quirk 1: 0 dB host volume -> vol_mul = (int16)0x8000 = -32768 -> gain -1.0 / Q15 -32768
        vm, row = api.host_volume(0)
        assert vm == -32768
        p["host_vol_mul"] = vm
        tab = api.loudness_table_q28(fs, 83.0, 100.0) if q else api.loudness_table(fs, 83.0, 100.0)
        p["loudness"] = tab[row]
        set_master(0.0)
    elif case in ("delay_equals_max", "delay_max_minus_one"):
        # dly == MAX_DELAY_SAMPLES aliases to zero delay through the ring mask (dsp_pipeline.c:232-236 clamps to MAX)
        d = max_delay if case == "delay_equals_max" else max_delay - 1
        for o in range(n_out):
            m["outputs"][o]["delay_samples"] = d if o % 2 == 0 else (d - 1 if o % 3 else 0)
    elif case == "clipping_hot_input":
        pcm = pcm_bytes_full_scale(F, 16, 13)
        bit_depth = 16
        set_preamp(6.0)
        set_master(0.0)
        p["host_vol_mul"] = api.host_volume(-256)[0]      # -1 dB: 0x7215
        p["leveller_enabled"] = 0
        for o in range(n_out):
            m["outputs"][o]["gain_db"] = np.float32(3.0)
            m["outputs"][o]["gain_linear"] = np.float32(10.0 ** (3.0 / 20.0))
    elif case == "pdm_saturation":
        # quirk 7: (int32)(x * 2^28) for |x| >= 8 (float path) / wrapping Q15 gain (Q28 path)
        pcm = pcm_bytes_full_scale(F, 24, 14)
        set_preamp(12.0)
        set_master(0.0)
        p["host_vol_mul"] = api.host_volume(-256)[0]
        p["leveller_enabled"] = 0
        m["outputs"][n_out - 1]["gain_db"] = np.float32(30.0)
        m["outputs"][n_out - 1]["gain_linear"] = np.float32(10.0 ** (30.0 / 20.0))
        for side in range(2):
            m["crosspoints"][side, n_out - 1]["gain_linear"] = np.float32(1.0)
    elif case == "host_muted":
        p["host_mute"] = 1
    elif case == "everything_off":
        for k in ("loudness_enabled", "crossfeed_enabled", "leveller_enabled"):
            p[k] = 0
        p["bypass_master_eq"] = 1
        for o in range(n_out):
            m["outputs"][o]["enabled"] = 1 if o == 1 else 0
            m["outputs"][o]["delay_samples"] = 0
        bq[0]["bypass"] = 1
    elif case == "sub_only":
        for o in range(n_out - 1):
            m["outputs"][o]["enabled"] = 0
    else:
        raise ValueError(case)
    return p, bq[0], pcm, bit_depth
