"""Soak test (development): for a few minutes classify random subsets of a fixed pool of rows in random
batch sizes and orders through the host entry and compare every result, bit for bit, with the pool's
canonical result (a site's result must not depend on the batch it travels in).  Catches races in the
shared-memory stages of the kernels (pooled epilogue tiles, staging) that a single pass would miss."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svision_b200 import classifier as C, sites, weights

seconds = float(os.environ.get("SOAK_SECONDS", 120))
rng = np.random.default_rng(int(os.environ.get("SOAK_SEED", 1)))
pool = np.concatenate([sites.make_sites_p1(40_000, seed=3), sites.make_sites_p2(10_000, seed=4),
                       sites.make_sites_p1(10_000, seed=5, profile="ont"), sites.edge_case_sites()])
with C.Classifier(weights.synthetic_weights(), device=0, max_batch=8192) as clf:
    ref_l, ref_p = clf.classify(pool)
    t0, it, checked, bad = time.time(), 0, 0, 0
    while time.time() - t0 < seconds:
        n = int(rng.choice([1, 7, 129, 841, 2048, 5000, 8192, 8193, 12345, 20000]))
        idx = rng.integers(0, pool.shape[0], size=n)
        l, p = clf.classify(pool[idx])
        if not (np.array_equal(l, ref_l[idx]) and np.array_equal(p, ref_p[idx])):
            bad += 1
            print(f"MISMATCH at iteration {it}: n={n}, {int((l != ref_l[idx]).sum())} labels, "
                  f"{int((p != ref_p[idx]).any(1).sum())} prob rows differ", flush=True)
        it += 1
        checked += n
    print(f"soak: {it} calls, {checked} sites in {time.time() - t0:.0f} s, {bad} calls with a mismatch", flush=True)
sys.exit(1 if bad else 0)
