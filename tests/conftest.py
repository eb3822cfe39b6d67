import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def synth_image(c, h, w, seed):
    """1/f-amplitude field per channel, min-max normalised to [0,1] (float64)."""
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    f = np.sqrt(fy * fy + fx * fx)
    f[0, 0] = 1.0
    base = np.fft.ifft2(np.fft.fft2(rng.normal(size=(h, w))) / f).real
    out = np.empty((c, h, w))
    for k in range(c):
        ind = np.fft.ifft2(np.fft.fft2(rng.normal(size=(h, w))) / f).real
        ch = 0.8 * base + 0.2 * ind
        out[k] = (ch - ch.min()) / (ch.max() - ch.min())
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import spiht_oracle
    spiht_oracle.lib()
    return spiht_oracle


def record_count(name, **values):
    """append a measured parity figure (e.g. quantised-coefficient mismatches against the float64 oracle) to
    gpurun_out/parity_counts.jsonl, so that a GPU run leaves the numbers behind (copied to profiles/)"""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_counts.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **values)) + "\n")
    except OSError:
        pass
