"""Generates the golden fixtures of tests/golden/.

The reference (theAdamColton/spiht) cannot run in the build container: its coder is Rust/pyo3 (no
cargo/rustc/maturin) and its Python front end needs PyWavelets and colour-science (neither installed nor in
the wheelhouse).  So the fixtures come from two sources, recorded per entry in `source`:

  * "reference-test":  values asserted by the reference's own tests (src/encoder_decoder.rs:845-1025);
  * "hand-traced":     known-answer vectors traced by hand from src/encoder_decoder.rs:155-454 and
                       src/lib.rs:15-29 during the survey;
  * "oracle":          outputs of oracle/spiht_ref.c (the C restatement, itself pinned by the two
                       sources above) on seeded inputs -- regression anchors for the CUDA path that do
                       not depend on the oracle being built where the GPU tests run.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import spiht_oracle as o  # noqa: E402


def main():
    kat = {
        "kat1": {
            "source": "hand-traced",
            "array": [[[5, -3, 1, 0], [2, -7, 0, 1], [0, 1, -1, 0], [3, 0, 0, -2]]],
            "ll": [2, 2], "max_bits": 10 ** 9, "max_n": 2, "nbits": 49, "hex": "135a166971be00",
        },
        "kat2": {
            "source": "reference-test (encoder_decoder.rs:864-875: max_n == 5) + hand-traced bit count",
            "fill": 32, "shape": [1, 16, 16], "ll": [2, 2], "max_bits": 10000, "max_n": 5, "nbits": 1870,
            "ones_prefix": 590,
        },
        "helpers": {
            "source": "reference-test (encoder_decoder.rs:851-862, 988-1009)",
            "is_bit_set": [[32, 5, True], [32, 0, False], [32, 4, False], [-69, 6, True], [3590854, 8, False]],
            "set_bit": [[-96, 5, False, -64], [-96, 5, True, -96], [-64, 5, True, -96], [96, 5, True, 96],
                        [96, 5, False, 64]],
            "is_element_sig": [[-21, 6, False], [-64, 6, True], [64, 6, True], [55, 6, False]],
        },
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # oracle-generated streams on seeded coefficient arrays (several shapes, budgets, odd / even LL)
    cases = [
        (3, 40, 56, 4, 6, 3000, 16.0), (1, 32, 32, 2, 2, 0, 16.0), (4, 32, 32, 2, 2, 10 ** 7, 16.0),
        (3, 37, 53, 6, 8, 5000, 60.0), (1, 24, 24, 3, 3, 0, 30.0), (2, 64, 48, 8, 6, 1234, 200.0),
        (3, 70, 70, 13, 13, 20000, 80.0), (1, 17, 17, 2, 2, 10 ** 9, 50.0),
    ]
    out = {}
    for idx, (c, h, w, llh, llw, mb, sigma) in enumerate(cases):
        rng = np.random.default_rng(1000 + idx)
        arr = rng.normal(0.0, sigma, (c, h, w)).astype(np.int32)
        data, max_n, nbits = o.encode_nbits(arr, llh, llw, mb)
        rec = o.decode(data, max_n, c, h, w, llh, llw)
        # a byte prefix (progressive decode, make_gif.py:46-61) decodes too
        cut = max(1, len(data) // 3)
        rec_cut = o.decode(data[:cut], max_n, c, h, w, llh, llw)
        out[f"c{idx}_arr"] = arr
        out[f"c{idx}_meta"] = np.array([llh, llw, mb, max_n, nbits, cut], dtype=np.int64)
        out[f"c{idx}_bytes"] = np.frombuffer(data, dtype=np.uint8)
        out[f"c{idx}_rec"] = rec
        out[f"c{idx}_rec_cut"] = rec_cut
    np.savez_compressed(os.path.join(HERE, "spiht_streams.npz"), **out)
    print("wrote kat.json and spiht_streams.npz (%d cases)" % len(cases))


if __name__ == "__main__":
    main()
