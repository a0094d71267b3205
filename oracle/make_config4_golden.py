"""Oracle calls for the WHOLE BASELINE config[3] stream (SURVEY.md 8(d) "Config 4"): 500 000 HiFi-profile
rows grouped into regions, seed 20261019.  TEST INFRASTRUCTURE: run once in the build container
(~1 h of CPU), the result is committed as tests/golden/config4_oracle_calls.npz:

    python oracle/make_config4_golden.py [--rows 500000] [--threads 6]

Per row: the fp32 oracle's label (argmax of oracle/alexnet.py on the images of oracle/encoder_c.c, both
pinned elsewhere), the softmax of that label and the margin between the two largest logits (rows whose
margin is below the GPU path's logit error are near-ties: the tests report them instead of demanding
equality there).  The GPU tests compare all 500 000 rows against this and the VCF text the calls produce."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import alexnet, encoder_c          # noqa: E402
from svision_b200 import sites, weights       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=500_000)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "config4_oracle_calls.npz"))
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    encoder_c.build()
    encoder_c.set_threads(a.threads)
    w = weights.synthetic_weights()
    rows = sites.make_region_table(a.rows, seed=sites.SEED_CONFIG4, profile="hifi").rows
    labels = np.empty(a.rows, np.int8)
    score = np.empty(a.rows, np.float32)
    margin = np.empty(a.rows, np.float32)
    t0 = time.time()
    step = 256
    for s in range(0, a.rows, step):
        logits = alexnet.forward(encoder_c.encode_f32(rows[s:s + step]), w, torch.float32)
        p = torch.softmax(logits, 1)
        l = torch.argmax(logits, 1)
        top2 = torch.topk(logits, 2, dim=1).values
        labels[s:s + step] = l.numpy()
        score[s:s + step] = p.gather(1, l[:, None])[:, 0].numpy()
        margin[s:s + step] = (top2[:, 0] - top2[:, 1]).numpy()
        if (s // step) % 100 == 0:
            print(f"{s} rows, {time.time() - t0:.0f} s", flush=True)
    np.savez_compressed(a.out, labels=labels, score=score, margin=margin,
                        meta=np.array([a.rows, sites.SEED_CONFIG4], dtype=np.int64))
    print("wrote", a.out, os.path.getsize(a.out), "bytes")


if __name__ == "__main__":
    main()
