"""S/PDIF subframe encoder: the CPU restatement against the reference's own
``spdif_update_subframe`` compiled from its header, and against IEC 60958 properties for the parts
that only exist as restatement (table fill, preamble / channel-status stamping)."""
import numpy as np
import pytest

from tests.orc import Oracle, RefSpdif

CS = bytes([0x04, 0x00, 0x00, 0x02, 0x0B])         # byte 3: 48 kHz code, so that a 1 appears beyond bit 24


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _decode_cells(word, n_cells):
    """Biphase-mark cells (2 bits each, LSB first) -> data bits; asserts every cell starts with a transition."""
    bits = []
    for j in range(n_cells):
        cell = (word >> (2 * j)) & 3
        assert cell & 1, f"cell {j} lacks its leading transition: {cell:02b}"
        bits.append(cell >> 1)
    return bits


def test_lookup_table_properties(orc):
    t = orc.spdif_table()
    for i in range(256):
        v, p = int(t[i]) & 0xFFFF, int(t[i]) >> 16
        assert _decode_cells(v, 8) == [(i >> j) & 1 for j in range(8)]
        assert p == bin(i).count("1") & 1
        assert int(t[i]) >> 17 == 0


@pytest.mark.skipif(not RefSpdif.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_update_subframe_matches_reference_header(orc):
    table = orc.spdif_table()
    ref = RefSpdif(table)
    rng = np.random.default_rng(7)
    samples = np.concatenate([rng.integers(-2**31, 2**31, 4000, dtype=np.int64),
                              np.array([0, 1, -1, 0x7FFFFF, -0x800000, 0x800000, 0xFFFFFF, 0x1000000, 0x55AA55, -0x55AA56])])
    for k, smp in enumerate(samples):
        l0 = int(rng.integers(0, 2**32))                   # arbitrary previous contents: only l[7:0] and h[30:24] survive
        h0 = int(rng.integers(0, 2**32))
        assert orc.spdif_update(table, l0, h0, smp) == ref.update(l0, h0, smp), (k, hex(int(smp) & 0xFFFFFFFF))


@pytest.mark.skipif(not RefSpdif.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_encode_equals_reference_copy_over_prestamped_buffer(orc):
    """orc_spdif_encode == the reference's converting copy run over a buffer stamped like init_spdif_buffer does."""
    table = orc.spdif_table()
    ref = RefSpdif(table)
    rng = np.random.default_rng(11)
    frames, pos0 = 500, 150
    words = rng.integers(-2**23, 2**23, (1, frames, 2), dtype=np.int64).astype(np.int32)
    got = orc.spdif_encode(words, pos0, CS)[0]
    cs_bits = [(CS[p // 8] >> (p % 8)) & 1 if p < 40 else 0 for p in range(192)]
    buf = np.zeros((frames, 2, 2), np.uint32)
    for n in range(frames):
        pos = (pos0 + n) % 192
        buf[n, 0] = (0x39 if pos == 0 else 0xC9, 0x55000000 | (cs_bits[pos] << 29))     # audio_spdif.c:104-106
        buf[n, 1] = (0x69, 0x55000000 | (cs_bits[pos] << 29))                            # :108-109
    ref.copy_s32(buf, words[0])
    assert np.array_equal(got, buf)


def test_stream_properties(orc):
    """Preambles, channel status, data recovery and even parity over a whole stream (IEC 60958-1)."""
    rng = np.random.default_rng(3)
    frames, pos0 = 192 * 3 + 17, 5
    words = rng.integers(-2**31, 2**31, (2, frames, 2), dtype=np.int64).astype(np.int32)
    out = orc.spdif_encode(words, pos0, CS)
    for s in range(2):
        for n in range(frames):
            pos = (pos0 + n) % 192
            c = (CS[pos // 8] >> (pos % 8)) & 1 if pos < 40 else 0
            for ch in range(2):
                l, h = int(out[s, n, ch, 0]), int(out[s, n, ch, 1])
                assert l & 0xFF == (0x69 if ch else (0x39 if pos == 0 else 0xC9))
                data = _decode_cells(l >> 8, 12) + _decode_cells(h & 0xFFFFFF, 12)
                assert data == [(int(words[s, n, ch]) >> j) & 1 for j in range(24)]
                v, u, cc, p = _decode_cells(h >> 24, 4)
                assert (v, u, cc) == (0, 0, c)
                assert (sum(data) + v + u + cc + p) % 2 == 0


def test_product_table_equals_restatement(orc):
    """dspi_spdif_lookup_table (host C in the product library) == the oracle's table."""
    from dspi_b200 import api
    assert np.array_equal(api.spdif_lookup_table(), orc.spdif_table())
