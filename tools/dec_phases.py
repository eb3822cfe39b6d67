"""Debug: per-phase cycle counts of the decoder for image 0 of a batch (spihtb_debug_dec_prof)."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spiht_b200 as spiht
from spiht_b200 import _lib, batch
from spiht_b200.utils import synthetic_images

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--bpp", type=float, default=0.5)
a = ap.parse_args()
st = spiht.SpihtSettings()
g = _lib.plan(a.size, a.size, "bior2.2", "reflect")
px = synthetic_images(a.batch, 3, a.size, a.size, seed=2000)
mb = int(a.size * a.size * a.bpp)
s, nbits, max_n, status, coeffs = batch.encode_images(px, g, st, mb)
nbytes = (nbits + 7) // 8
out = (ctypes.c_ulonglong * 16)()
lib = _lib.lib()
lib.spihtb_debug_enc_prof(out)
v = list(out)
print("encoder (image 0): cycles", {"lip": v[0], "lis": v[1], "refine": v[2], "image_total": v[3]},
      "chunks", {"lip": v[8], "lis": v[9], "refine": v[10]},
      "lis laps (SPIHTB_ENC_LAPS builds)", {"load+classify": v[4], "scan1+worklists": v[5], "denseA_tail": v[6], "denseB": v[7],
                                            "flush": v[11], "A_load+record": v[12], "A_scan2": v[13], "A_emit+append": v[14]})
for it in range(3):
    batch.decode_images(s, nbytes, max_n, 3, g, st, dtype=torch.float32)
    torch.cuda.synchronize()
    lib.spihtb_debug_dec_prof(out)
    v = list(out)
    names = ["lip_rounds", "lis_tmask", "chain", "lis_batches", "refine", "image_total"]
    print({n: v[i] for i, n in enumerate(names)},
          {"lip_rounds_n": v[8], "chain_iters": v[9], "chain_events": v[10], "lis_rounds": v[11], "lis_batches_n": v[12],
           "batch_cycles_not_last_round": v[13], "overlappable": v[14], "generations": v[15]})

# per image: total / walk cycles and the SM it ran on (prof builds)
if hasattr(lib, "spihtb_debug_dec_img"):
    import numpy as np
    n = min(a.batch, 1024)
    buf = (ctypes.c_ulonglong * (4 * n))()
    lib.spihtb_debug_dec_img(buf, n)
    arr = np.array(list(buf), dtype=np.int64).reshape(n, 4)
    sm = arr[:, 2]
    cnt = np.bincount(sm, minlength=160)
    shared = cnt[sm] > 1
    for name, m in (("alone on its SM", ~shared), ("sharing its SM", shared)):
        if m.any():
            print(name, int(m.sum()), "images: total cycles min/median/max", int(arr[m, 0].min()), int(np.median(arr[m, 0])),
                  int(arr[m, 0].max()), "walk median", int(np.median(arr[m, 1])))
    # hardware warp slot of warp 0 of the CTAs that share an SM
    pairs = {}
    for i in range(n):
        pairs.setdefault(int(arr[i, 2]), []).append(int(arr[i, 3]))
    print("hardware warp id of warp 0, CTAs sharing an SM (first 12 SMs):", [v for v in list(pairs.values()) if len(v) > 1][:12])
