"""Random but well-formed WireBulkParams packets (bulk_params.h) for the bulk-ingest tests."""
import numpy as np

from dspi_b200 import layouts as L


def dims(platform):
    return (11, 9) if platform == L.PLATFORM_RP2350 else (7, 5)


def wire_packet(platform, seed, version=6):
    rng = np.random.default_rng([seed, platform, version])
    nc, no = dims(platform)
    w = np.zeros(1, L.WIRE_BULK)
    h = w["header"][0]
    h["format_version"], h["platform_id"], h["num_channels"], h["num_output_channels"] = version, platform, nc, no
    h["num_input_channels"], h["max_bands"] = 2, 12
    v5 = L.WIRE_BULK.itemsize - 32
    h["payload_length"] = {2: v5 - 32, 3: v5, 4: v5, 5: v5, 6: L.WIRE_BULK.itemsize}[version]
    h["fw_version_major"], h["fw_version_minor"] = 1, 1

    def db(lo=-70.0, hi=25.0, size=None):                      # beyond the Taylor clamp on both sides, plus exact zeros
        x = np.asarray(rng.uniform(lo, hi, size), np.float32)
        return np.where(np.asarray(rng.random(size)) < 0.15, np.float32(0.0), x).astype(np.float32)

    g = w["global"][0]
    g["preamp_gain_db"], g["bypass"], g["loudness_enabled"] = db(), rng.integers(0, 2), rng.integers(0, 2)
    g["loudness_ref_spl"], g["loudness_intensity_pct"] = rng.uniform(70, 95), rng.uniform(0, 150)
    x = w["crossfeed"][0]
    x["enabled"], x["preset"], x["itd_enabled"] = rng.integers(0, 2), rng.integers(0, 4), rng.integers(0, 2)
    x["custom_fc"], x["custom_feed_db"] = rng.uniform(400, 2500), rng.uniform(0, 16)
    w["legacy"][0]["gain_db"] = db(size=3)
    w["legacy"][0]["mute"] = rng.integers(0, 2, 3)
    w["delays"][0]["delay_ms"][:nc] = rng.uniform(0, 30, nc)
    cp = w["crosspoints"][0]
    cp["enabled"][:, :no] = rng.integers(0, 2, (2, no))
    cp["phase_invert"][:, :no] = rng.integers(0, 2, (2, no))
    cp["gain_db"][:, :no] = db(size=(2, no))
    o = w["outputs"][0]
    o["enabled"][:no], o["mute"][:no] = rng.integers(0, 2, no), rng.integers(0, 4, no) == 0
    o["gain_db"][:no] = db(size=no)
    o["delay_ms"][:no] = rng.uniform(-1, 45 if platform == L.PLATFORM_RP2350 else 22, no)
    w["pins"][0]["num_pin_outputs"] = 5 if platform == L.PLATFORM_RP2350 else 3
    e = w["eq"][0]
    e["type"][:nc] = rng.integers(0, 6, (nc, 12))
    e["freq"][:nc] = (20.0 * (1200.0 ** rng.random((nc, 12)))).astype(np.float32)
    e["q"][:nc] = rng.uniform(0.05, 25, (nc, 12))
    e["gain_db"][:nc] = np.where(rng.random((nc, 12)) < 0.3, 0.0, rng.uniform(-12, 12, (nc, 12)))
    lv = w["leveller"][0]
    lv["enabled"], lv["speed"], lv["lookahead"] = rng.integers(0, 2), rng.integers(0, 3), rng.integers(0, 2)
    lv["amount"], lv["max_gain_db"], lv["gate_threshold_db"] = rng.uniform(0, 100), rng.uniform(0, 35), rng.uniform(-96, 0)
    w["preamp"][0]["preamp_db"] = db(size=2)
    mv = rng.choice([0.0, -6.0, -127.0, -128.0, -200.0, 3.0, np.nan, np.inf, float(rng.uniform(-100, 0))])
    w["master_volume"][0]["master_volume_db"] = np.float32(mv)
    return w
