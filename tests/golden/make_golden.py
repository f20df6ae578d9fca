"""Generates tests/golden/*.npz from the UNMODIFIED reference sources compiled on the host
(oracle/_ref/*.so, built by `make -C oracle ref` from /root/reference).  Run in the build
container only (the GPU box has no /root/reference); the outputs are committed.

    python tests/golden/make_golden.py

Fixture A  "cfg1": BASELINE config 1 - 2 channels x 3 bands @48 kHz:
           low shelf 100 Hz +4 dB Q.707 (SVF), peaking 1 kHz -3 dB Q1.4 (SVF), high shelf
           10 kHz +2 dB Q.707 (TDF2), bands 3-9 flat; inputs: impulse, 1 kHz sine -6 dBFS,
           log sweep 20 Hz-20 kHz, xorshift32 s16 noise; 100 packets x 48 samples.
Fixture B  "mix":  24 channels, random per-band types (all six), 96 kHz, 960 samples.
Fixture C  "spdif": 3 stereo streams x 421 frames of 24-bit words (edge values + random) through the
           reference's own spdif_update_subframe (compiled from its header, oracle/ref_spdif_shim.c) over
           buffers stamped as init_spdif_buffer does, block position starting at 187.
Each of A and B holds coefficients (reference dsp_compute_coefficients), inputs, and outputs + final
filter state of dsp_process_channel_block for the strict, fused and Q28 builds.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dspi_b200 import layouts as L          # noqa: E402
from dspi_b200 import workloads as W        # noqa: E402
from tests.orc import Ref, RefSpdif, Oracle, build_oracle     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cfg1_params():
    p = np.zeros((2, L.MAX_BANDS), L.EQ_PARAM)
    p["freq"], p["Q"] = 1000.0, 0.707
    for c in range(2):
        p[c, 0] = (c, 0, L.LOWSHELF, 0, 100.0, 0.707, 4.0)
        p[c, 1] = (c, 1, L.PEAKING, 0, 1000.0, 1.4, -3.0)
        p[c, 2] = (c, 2, L.HIGHSHELF, 0, 10000.0, 0.707, 2.0)
    return p


def cfg1_inputs(T):
    t = np.arange(T, dtype=np.float64)
    fs = 48000.0
    imp = np.zeros(T, np.float32); imp[0] = 1.0
    sine = (0.5 * np.sin(2 * np.pi * 1000.0 * t / fs)).astype(np.float32)
    k = np.log(20000.0 / 20.0) / (T / fs)
    sweep = (0.5 * np.sin(2 * np.pi * 20.0 * (np.exp(k * t / fs) - 1.0) / k)).astype(np.float32)
    noise = W.inputs_f32(1, T)[0]
    return {"impulse": imp, "sine": sine, "sweep": sweep, "noise": noise}


def run(refs, params, x_f32, fs, packet):
    out = {}
    for fl in ("f32s", "f32f"):
        bq = np.zeros(params.shape, L.BIQUAD_F32)
        refs["f32s"].eq_coeffs(params.copy(), bq, fs)       # coefficients are data: always the strict build's
        y = x_f32.copy()
        b = bq.copy()
        refs[fl].eq_many(b, y, 10, packet)
        out[fl] = (bq, y, b)
    bq = np.zeros(params.shape, L.BIQUAD_Q28)
    refs["q28"].eq_coeffs(params.copy(), bq, fs)
    xq = np.round(x_f32.astype(np.float64) * (1 << 28)).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32)
    y = xq.copy()
    b = bq.copy()
    refs["q28"].eq_many(b, y, 10, packet)
    out["q28"] = (bq, y, b, xq)
    return out


def main():
    build_oracle(with_ref=True)
    refs = {k: Ref(k) for k in ("f32s", "f32f", "q28")}
    # ---- fixture A
    T = 4800
    params = cfg1_params()
    ins = cfg1_inputs(T)
    save = {"params": params}
    for name, sig in ins.items():
        x = np.stack([sig, -0.5 * sig]).astype(np.float32)          # L and a scaled/inverted R
        r = run(refs, params, x, 48000.0, 48)
        save[f"{name}_x"] = x
        for fl in ("f32s", "f32f"):
            save[f"{name}_{fl}_y"] = r[fl][1]
            save[f"{name}_{fl}_state"] = r[fl][2]
        save[f"{name}_q28_x"] = r["q28"][3]
        save[f"{name}_q28_y"] = r["q28"][1]
        save[f"{name}_q28_state"] = r["q28"][2]
    save["bq_f32"] = r["f32s"][0]
    save["bq_q28"] = r["q28"][0]
    np.savez_compressed(os.path.join(HERE, "cfg1.npz"), **save)
    # ---- fixture B
    fs, Cn, T = 96000.0, 24, 960
    params = W.eq_params("mixed", Cn, fs=fs, seed=2024)
    x = W.inputs_f32(Cn, T)
    r = run(refs, params, x, fs, 96)
    np.savez_compressed(os.path.join(HERE, "mix.npz"), params=params, x=x, bq_f32=r["f32s"][0], bq_q28=r["q28"][0],
                        f32s_y=r["f32s"][1], f32s_state=r["f32s"][2], f32f_y=r["f32f"][1], f32f_state=r["f32f"][2],
                        q28_x=r["q28"][3], q28_y=r["q28"][1], q28_state=r["q28"][2])
    # ---- fixture C: S/PDIF subframes
    cs = bytes([0x04, 0x00, 0x00, 0x02, 0x0B])
    rng = np.random.default_rng(60958)
    n_streams, frames, pos0 = 3, 421, 187
    words = rng.integers(-2**31, 2**31, (n_streams, frames, 2), dtype=np.int64).astype(np.int32)
    words[0, :10, 0] = [0, 1, -1, 0x7FFFFF, -0x800000, 0x800000, 0xFFFFFF, 0x1000000, 0x55AA55, -0x55AA56]
    rs = RefSpdif(Oracle().spdif_table())                        # the table itself is restated (audio_spdif.c:141-153)
    sub = np.zeros((n_streams, frames, 2, 2), np.uint32)
    for s_ in range(n_streams):
        for n in range(frames):
            pos = (pos0 + n) % 192
            c = (cs[pos // 8] >> (pos % 8)) & 1 if pos < 40 else 0
            sub[s_, n, 0] = (0x39 if pos == 0 else 0xC9, 0x55000000 | (c << 29))
            sub[s_, n, 1] = (0x69, 0x55000000 | (c << 29))
        rs.copy_s32(sub[s_], words[s_])
    np.savez_compressed(os.path.join(HERE, "spdif.npz"), words=words, subframes=sub, pos0=np.uint32(pos0), cs=np.frombuffer(cs, np.uint8))
    for f in ("cfg1.npz", "mix.npz", "spdif.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")




def make_bulk():
    """Fixture D "bulk": WireBulkParams packets and the states the reference's bulk_params_apply() leaves (both platforms)."""
    import ctypes as C
    from dspi_b200 import api
    from tests.bulk_cases import wire_packet
    from tests.orc import ORACLE_DIR
    save = {}
    for platform, key in ((L.PLATFORM_RP2350, "rp2350"), (L.PLATFORM_RP2040, "rp2040")):
        ref = C.CDLL(os.path.join(ORACLE_DIR, "_ref", f"libdspi_ref_bulk_{key}.so"))
        wires, states, rcs = [], [], []
        for version in (2, 4, 5, 6, 6, 6):
            w = wire_packet(platform, 700 + len(wires), version)
            if len(wires) == 4:
                w["header"][0]["payload_length"] = 64        # rejected: -4
            st = api.bulk_state_defaults(platform)           # layout only; values come from the reference below
            rcs.append(ref.ref_bulk_apply(w.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p)))
            # the x86 object returns INT_MIN where the ARM firmware saturates: store the ARM value
            over = st["preamp_linear"] * np.float32(2.0 ** 28) >= np.float32(2.0 ** 31)
            st["preamp_q28"][over & (st["preamp_q28"] == np.iinfo(np.int32).min)] = np.iinfo(np.int32).max
            wires.append(w.copy())
            states.append(st.copy())
        # raw bytes: .npy headers cannot describe the padded C layouts
        save[f"{key}_wire"] = np.frombuffer(b"".join(x.tobytes() for x in wires), np.uint8).reshape(len(wires), -1)
        save[f"{key}_state"] = np.frombuffer(b"".join(x.tobytes() for x in states), np.uint8).reshape(len(states), -1)
        save[f"{key}_rc"] = np.array(rcs, np.int32)
    np.savez_compressed(os.path.join(HERE, "bulk.npz"), **save)
    print("bulk.npz", os.path.getsize(os.path.join(HERE, "bulk.npz")), "bytes")


def make_preset():
    """Fixture E "preset": device states, the slot images the reference's preset_save() writes for them, and the states
    its preset_load() leaves when those images are loaded into a factory-default device (both platforms)."""
    import ctypes as C
    from dspi_b200 import api
    from tests.bulk_cases import wire_packet
    from tests.orc import ORACLE_DIR
    save = {}
    for platform, key in ((L.PLATFORM_RP2350, "rp2350"), (L.PLATFORM_RP2040, "rp2040")):
        ref = C.CDLL(os.path.join(ORACLE_DIR, "_ref", f"libdspi_ref_preset_{key}.so"))
        ref.ref_preset_load.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint8, C.c_float, C.c_void_p]
        n = api.preset_slot_size(platform)
        states, images, loaded, slots = [], [], [], []
        for i in range(4):
            st = api.bulk_state_defaults(platform)
            assert api.bulk_params_apply(wire_packet(platform, 900 + i), st, exact_db=True) == 0
            sector = np.zeros(4096, np.uint8)
            assert ref.ref_preset_save(st.ctypes.data_as(C.c_void_p), i + 3, sector.ctypes.data_as(C.c_void_p)) == 0
            out = api.bulk_state_defaults(platform)
            assert ref.ref_preset_load(sector.ctypes.data, n, i + 3, 1, 0.0, out.ctypes.data) == 0
            over = out["preamp_linear"] * np.float32(2.0 ** 28) >= np.float32(2.0 ** 31)          # ARM saturation
            out["preamp_q28"][over & (out["preamp_q28"] == np.iinfo(np.int32).min)] = np.iinfo(np.int32).max
            states.append(st.copy()); images.append(sector[:n].copy()); loaded.append(out.copy()); slots.append(i + 3)
        save[f"{key}_state"] = np.frombuffer(b"".join(x.tobytes() for x in states), np.uint8).reshape(len(states), -1)
        save[f"{key}_loaded"] = np.frombuffer(b"".join(x.tobytes() for x in loaded), np.uint8).reshape(len(loaded), -1)
        save[f"{key}_image"] = np.stack(images)
        save[f"{key}_slot"] = np.array(slots, np.int32)
    np.savez_compressed(os.path.join(HERE, "preset.npz"), **save)
    print("preset.npz", os.path.getsize(os.path.join(HERE, "preset.npz")), "bytes")


if __name__ == "__main__":
    if "--bulk" in sys.argv:
        make_bulk()
    elif "--preset" in sys.argv:
        make_preset()
    else:
        main()
        make_bulk()
        make_preset()
