"""VERDICT r1 #3, measured: what would a valid-only M tiling of conv3 cost if its A operand had to be
fetched per tap (im2col-mode TMA / a materialised im2col), which is what gives up the slab?

Two launches of the SAME layer kernel on the same arithmetic (13x13x256 -> 384 channels, 3x3, 3 passes):
  slab   : M = n * 196 rows of the padded 14x14 grid, 9 taps as row offsets of ONE slab per channel block
           (what the library runs: 16 % of the rows are padding)
  im2col : M = n * 169 valid rows only, K = 9 * 256 as one long reduction whose A tile is fetched for
           every k-block (no slab reuse) -- the traffic pattern of any scheme whose M index skips the pad
           positions, since then a tap is no longer a row offset of one matrix
Prints milliseconds of the layer kernel alone and the per-k-block cycle cost at the measured SM clock."""
import os, sys, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svision_b200 import classifier as C, _lib

n = int(os.environ.get("PAD_N", 4096))
lib = _lib.load()
torch.manual_seed(0)
dev = torch.device("cuda", 0)
w = torch.randn(384, 9 * 256, device=dev) * 0.02


def sm_clock():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip()
    return float(out.splitlines()[0])


res = {}
for name, m, taps, offs in (("slab", n * 196, 9, [(kh - 1) * 14 + (kw - 1) for kh in range(3) for kw in range(3)]),
                            ("im2col", n * 169, 1, [0])):
    k = 256 if taps == 9 else 9 * 256
    a = torch.randn(m, k, device=dev)
    ms, clk = [], []
    for it in range(4):
        if taps == 9:
            C.conv_selftest(a, w, offs, block_n=192)
        else:
            C.gemm_selftest(a, w, block_n=192)
        torch.cuda.synchronize()
        clk.append(sm_clock())
        ms.append(float(lib.svx_selftest_last_ms()))
    del a
    torch.cuda.empty_cache()
    t = float(np.median(ms[1:]))
    kblocks_per_cta = (m / 256) * 2 * 36 / 74                   # pair tiles x 2 column tiles x 36 k-blocks / 74 pairs
    res[name] = t
    print(f"{name:7s} M = {m:8d} rows  {t:7.3f} ms   ({ms})", flush=True)
print(f"valid-only rows / padded rows = {169/196:.3f}; im2col time / slab time = {res['im2col']/res['slab']:.3f} "
      f"(< 1 would mean the valid-only tiling pays; the MMA work alone predicts {169/196:.3f})")
