"""Generates tests/golden/chain.npz from the reference's OWN process_audio_packet() and delta-sigma loop, compiled
unmodified on the host (oracle/_ref/libdspi_ref_chain_*.so, libdspi_ref_pdm.so; needs /root/reference).  Run in the build
container only; the output is committed so that machines without the reference still hold reference-made vectors.

    python tests/golden/make_golden_chain.py

Per flavour (f32s, f32f: RP2350 shape; q28: RP2040 shape): 6 instances x 4 packets of 96 frames of packed 24-bit PCM @96 kHz,
leveller off (its libm is the one policy item), every other stage as tests/chain_cases.py draws it:
parameter records, biquads, PCM bytes, the S/PDIF words of every pair, the Q28 sub samples handed to pdm_push_sample()
and the PDM words the reference's modulator makes of them, meters and clip flags after the last packet.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.chain_cases import chain_params, chain_params_q28, pcm_bytes                      # noqa: E402
from tests.orc import Oracle, RefChain, RefPdm, build_oracle, make_orc_chain, make_orc_chain_q28   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FS, N, NPK, FPP = 96000.0, 6, 4, 96


def main():
    build_oracle(with_ref=True)
    orc = Oracle()
    pdm = RefPdm()
    out = {"fs": FS, "n_packets": NPK, "fpp": FPP}
    for fl in ("f32s", "f32f", "q28"):
        q = fl == "q28"
        P, bq = chain_params_q28(orc, N, FS, 61, leveller=False) if q else chain_params(orc, N, FS, 62, leveller=False)
        for i in range(N):
            P[i]["preset_mute_gain"] = 1.0
            P[i]["matrix"]["outputs"][4 if q else 8]["enabled"] = 1
        pcm = pcm_bytes(N, NPK * FPP, 24, 63)
        ref = RefChain(fl)
        sp, sub, words, peaks, clip = [], [], [], [], []
        for i in range(N):
            ch = (make_orc_chain_q28 if q else make_orc_chain)(orc, P[i], bq[i])
            s, u = ref.run(ch, FS, pcm[i], 24, NPK, FPP)
            w, _ = pdm.run(u)
            sp.append(s); sub.append(u); words.append(w)
            peaks.append(list(ch.peaks)); clip.append(int(ch.clip_flags))
        out.update({f"{fl}_params": P, f"{fl}_biquads": bq, f"{fl}_pcm": pcm, f"{fl}_spdif": np.stack(sp), f"{fl}_sub_q28": np.stack(sub),
                    f"{fl}_pdm": np.stack(words), f"{fl}_peaks": np.array(peaks, np.uint16), f"{fl}_clip": np.array(clip, np.uint16)})
    np.savez_compressed(os.path.join(HERE, "chain.npz"), **out)
    print("wrote chain.npz", os.path.getsize(os.path.join(HERE, "chain.npz")), "bytes")


if __name__ == "__main__":
    main()
