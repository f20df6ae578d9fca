"""Synthetic workloads of BASELINE.json's configs.

Parameter recipes (``EqParamPacket`` rows) and PCM inputs are generated from
fixed seeds so tests, fixtures and the bench agree on the same bits.
"""
import numpy as np

from . import layouts as L

XORSHIFT_SEED = 123456789          # pdm_generator.c:62 — the reference's own PRNG seed


def xorshift32(state):
    """One xorshift32 step on a uint32 array (pdm_generator.c:63-68)."""
    state = state ^ (state << np.uint32(13))
    state = state ^ (state >> np.uint32(17))
    state = state ^ (state << np.uint32(5))
    return state


def xorshift_s16(C, T, ch0=0):
    """int16 [C, T]: per-channel xorshift32 stream seeded ``123456789 ^ ch``."""
    st = (np.uint32(XORSHIFT_SEED) ^ (np.arange(C, dtype=np.uint32) + np.uint32(ch0))).astype(np.uint32)
    st[st == 0] = 1
    out = np.empty((C, T), np.int16)
    with np.errstate(over="ignore"):
        for t in range(T):
            st = xorshift32(st)
            out[:, t] = (st >> np.uint32(16)).astype(np.uint16).view(np.int16)
    return out


def inputs_device(C, T, n_blocks, q28, device, ch0=0):
    """The same per-channel xorshift32 streams generated on the GPU (torch, int64 lanes emulating uint32): returns
    ``n_blocks`` tensors [C, T], block b holding frames [b*T, (b+1)*T) of every channel's stream - float32
    ``s16 / 65536`` or int32 ``s16 << 13`` like :func:`inputs_f32` / :func:`inputs_q28` (bit-identical to them)."""
    import torch
    mask = 0xFFFFFFFF
    st = (torch.arange(C, dtype=torch.int64, device=device) + int(ch0)) ^ int(XORSHIFT_SEED)
    st = torch.where(st == 0, torch.ones_like(st), st)
    blocks = []
    for _ in range(n_blocks):
        raw = torch.empty((T, C), dtype=torch.int16, device=device)
        for t in range(T):
            st = st ^ ((st << 13) & mask)
            st = st ^ (st >> 17)
            st = st ^ ((st << 5) & mask)
            raw[t] = (st >> 16).to(torch.int16)             # wraps like the uint16 -> int16 view on the host
        s16 = raw.t().contiguous()
        if q28:
            blocks.append(s16.to(torch.int32) << 13)
        else:
            blocks.append(s16.to(torch.float32) / 65536.0)
    return blocks


def inputs_f32(C, T, ch0=0):
    """float32 [C, T] uniform in [-0.5, 0.5): s16 / 65536 (exact in float)."""
    return (xorshift_s16(C, T, ch0).astype(np.float32) / np.float32(65536.0)).astype(np.float32)


def inputs_q28(C, T, ch0=0):
    """int32 [C, T]: s16 << 14 (usb_audio.c:1010), halved to keep ±0.5 full scale."""
    return (xorshift_s16(C, T, ch0).astype(np.int32) << 13).astype(np.int32)


ISO_OCTAVES = [31.5, 63.0, 125.0, 250.0, 500.0, 1000.0, 2000.0, 4000.0, 8000.0, 16000.0]


def eq_params(variant, C, fs=96000.0, nbands=L.NUM_BANDS, seed=1, ch0=0):
    """EQ_PARAM [C, MAX_BANDS] recipes.

    ``A``      all-TDF2: 10 peaking bands at 13-42 kHz (>= fs/7.5, <= 0.45 fs)
    ``B``      realistic: ISO octave centres; low shelf, 8 peaking, high shelf
               (at 96 kHz: 9 SVF bands + 1 TDF2 band, dsp_pipeline.c:88)
    ``mixed``  per-channel random types incl. LP/HP/flat (non-uniform structure)
    Rows beyond ``nbands`` stay FLAT.  Per-channel streams are keyed by the
    absolute channel index so shards of a larger job generate identical rows.
    """
    p = np.zeros((C, L.MAX_BANDS), L.EQ_PARAM)
    p["freq"] = 1000.0
    p["Q"] = 0.707
    for c in range(C):
        rng = np.random.default_rng([seed, ch0 + c])
        u = rng.random((nbands, 4))
        for b in range(nbands):
            r = p[c, b]
            r["channel"] = 0
            r["band"] = b
            gain = np.float32(-6.0 + 12.0 * u[b, 0])
            if abs(gain) < 0.05:
                gain = np.float32(0.5)
            q = np.float32(0.7 + 3.3 * u[b, 1])
            if variant == "A":
                lo, hi = fs / 7.5 * 1.02, fs * 0.44
                r["type"], r["freq"], r["Q"], r["gain_db"] = L.PEAKING, np.float32(lo + (hi - lo) * u[b, 2]), q, gain
            elif variant == "B":
                t = L.LOWSHELF if b == 0 else (L.HIGHSHELF if b == nbands - 1 else L.PEAKING)
                qq = np.float32(0.707) if t != L.PEAKING else q
                r["type"], r["freq"], r["Q"], r["gain_db"] = t, np.float32(ISO_OCTAVES[b % 10]), qq, gain
            elif variant == "mixed":
                t = int(u[b, 3] * 6) % 6
                f = np.float32(20.0 * (fs * 0.45 / 20.0) ** u[b, 2])
                r["type"], r["freq"], r["Q"], r["gain_db"] = t, f, q, gain
            else:
                raise ValueError(variant)
    return p


def eq_params_fast(variant, C, fs=96000.0, nbands=L.NUM_BANDS, seed=1, ch0=0):
    """Vectorised generator for the big bench shapes (variants A and B only).

    Not bit-compatible with :func:`eq_params` (different stream layout); used
    where only the shape of the work matters, never for fixtures.
    """
    rng = np.random.default_rng([seed, ch0, C])
    p = np.zeros((C, L.MAX_BANDS), L.EQ_PARAM)
    p["freq"] = 1000.0
    p["Q"] = 0.707
    u = rng.random((C, nbands, 3)).astype(np.float32)
    gain = (-6.0 + 12.0 * u[..., 0]).astype(np.float32)
    gain[np.abs(gain) < 0.05] = 0.5
    q = (0.7 + 3.3 * u[..., 1]).astype(np.float32)
    p["band"][:, :nbands] = np.arange(nbands, dtype=np.uint8)[None, :]
    p["gain_db"][:, :nbands] = gain
    if variant == "A":
        lo, hi = fs / 7.5 * 1.02, fs * 0.44
        p["type"][:, :nbands] = L.PEAKING
        p["freq"][:, :nbands] = (lo + (hi - lo) * u[..., 2]).astype(np.float32)
        p["Q"][:, :nbands] = q
    elif variant == "B":
        types = np.full(nbands, L.PEAKING, np.uint8)
        types[0] = L.LOWSHELF
        types[nbands - 1] = L.HIGHSHELF
        p["type"][:, :nbands] = types[None, :]
        p["freq"][:, :nbands] = np.array([ISO_OCTAVES[b % 10] for b in range(nbands)], np.float32)[None, :]
        qq = q.copy()
        qq[:, 0] = 0.707
        qq[:, nbands - 1] = 0.707
        p["Q"][:, :nbands] = qq
    else:
        raise ValueError(variant)
    return p


def chain_config3(N, fs=96000.0, seed=1, distinct=64):
    """BASELINE config 3: N RP2350-shape instances, every stage on.

    Matrix L -> Out1/3/5/7, R -> Out2/4/6/8, (L+R)*0.5 -> sub; crossfeed preset 0 with ITD; loudness on
    at -20 dB host volume; leveller on (slow, look-ahead); output delays uniform in 0-40 ms; all 11 EQ
    rows with 10 active bands (variant B).  `distinct` different instance configurations are
    generated with the product's own host parameter API and tiled over the N instances.
    Returns (CHAIN_PARAMS_F32 [N], BIQUAD_F32 [N, 11, 12]).
    """
    from . import api
    rng = np.random.default_rng(seed)
    D = min(distinct, N)
    P = np.zeros(D, L.CHAIN_PARAMS_F32)
    vol_mul, row = api.host_volume(-20 * 256)
    table = api.loudness_table(fs, 83.0, 100.0)
    xf = api.crossfeed_coefficients(fs, True, True, 0)
    lev = api.leveller_coefficients(fs, 50.0, 0, 15.0, -96.0)
    bq = np.zeros((D, L.CHAIN_EQ_CHANNELS, L.MAX_BANDS), L.BIQUAD_F32)
    for i in range(D):
        p = P[i]
        p["host_vol_mul"], p["preset_mute_gain"], p["master_volume_linear"] = vol_mul, 1.0, 1.0
        p["preamp_linear"] = [1.0, 1.0]
        p["loudness_enabled"], p["crossfeed_enabled"], p["leveller_enabled"], p["leveller_lookahead"] = 1, 1, 1, 1
        p["loudness"], p["crossfeed"], p["leveller"] = table[row], xf, lev
        m = p["matrix"]
        for o in range(L.CHAIN_OUTPUTS):
            oc = m["outputs"][o]
            oc["enabled"], oc["mute"], oc["gain_db"], oc["gain_linear"] = 1, 0, 0.0, 1.0
            oc["delay_ms"] = np.float32(40.0 * rng.random())
            oc["delay_samples"] = api.delay_samples(float(oc["delay_ms"]), fs, o == L.CHAIN_OUTPUTS - 1)
            for side in range(2):
                x = m["crosspoints"][side, o]
                sub = o == L.CHAIN_OUTPUTS - 1
                x["enabled"] = 1 if (sub or o % 2 == side) else 0
                x["gain_db"], x["gain_linear"] = (-6.0206, 0.5) if sub else (0.0, 1.0)
        bq[i] = api.compute_coefficients(eq_params("B", L.CHAIN_EQ_CHANNELS, fs=fs, seed=seed + i), q28=False, fs=fs)
    reps = (N + D - 1) // D
    return np.tile(P, reps)[:N].copy(), np.tile(bq, (reps, 1, 1))[:N].copy()


def chain_config3_q28(N, fs=96000.0, seed=9):
    """Config 3 on the RP2040 shape (Q28 arithmetic, 2 in -> 4 S/PDIF channels + PDM sub): every stage on,
    delays 0-1900 samples, one EQ recipe set (variant B) shared by all instances.
    Returns (CHAIN_PARAMS_Q28 [N], BIQUAD_Q28 [N, 7, 12])."""
    from . import api
    P = np.zeros(N, L.CHAIN_PARAMS_Q28)
    vol_mul, row = api.host_volume(-20 * 256)
    tq = api.loudness_table_q28(fs, 83.0, 100.0)
    lev = api.leveller_coefficients(fs, 50.0, 0, 15.0, -96.0)
    xfq = api.crossfeed_coefficients_q28(fs, True, True, 0)
    for i in range(min(N, 64)):
        p = P[i]
        p["host_vol_mul"], p["preset_mute_gain"], p["master_volume_q15"] = vol_mul, 1.0, 32768
        p["preamp_q28"] = [1 << 28, 1 << 28]
        p["loudness_enabled"], p["crossfeed_enabled"], p["leveller_enabled"], p["leveller_lookahead"] = 1, 1, 1, 1
        p["loudness"], p["crossfeed"], p["leveller"] = tq[row], xfq, lev
        m = p["matrix"]
        for o in range(L.CHAINQ_OUTPUTS):
            oc = m["outputs"][o]
            oc["enabled"], oc["gain_linear"] = 1, 1.0
            oc["delay_samples"] = (97 * (o + 1) + 13 * i) % 1900
            for side in range(2):
                x = m["crosspoints"][side, o]
                sub = o == L.CHAINQ_OUTPUTS - 1
                x["enabled"] = 1 if (sub or o % 2 == side) else 0
                x["gain_linear"] = 0.5 if sub else 1.0
    P = np.tile(P[:min(N, 64)], (N + 63) // 64)[:N].copy()
    bq = api.compute_coefficients(eq_params("B", L.CHAINQ_EQ_CHANNELS, fs=fs, seed=seed), q28=True, fs=fs)
    return P, np.tile(bq[None], (N, 1, 1))
