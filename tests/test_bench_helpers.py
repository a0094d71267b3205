"""CPU: the parts of bench.py that do not need a GPU -- the known-answer bookkeeping, the traffic figure
parsed from the committed ncu summary, and the reference arm (the reference's own BatchGenerator from
baseline/_ref when installed, the C port otherwise) on a tiny sample."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_known_answer_slots_and_check():
    rows, labels, scores = bench.load_known_answers()
    assert rows.shape == (bench.KNOWN, 12) and labels.shape == (bench.KNOWN,) and scores.shape == (bench.KNOWN,)
    assert ((scores > 0.2) & (scores <= 1.0)).all() and len(np.unique(labels)) >= 3
    per, world = 1000, 3
    slots = bench.known_slots(per)
    assert slots.size == bench.KNOWN and slots[0] == 0 and slots[-1] == per - 1 and len(set(slots)) == bench.KNOWN
    gl = np.zeros(world * per, np.int32)
    gs = np.zeros(world * per, np.float32)
    for r in range(world):
        gl[r * per + slots] = labels
        gs[r * per + slots] = scores
    ok, err = bench.check_known(gl, gs, world, per, labels, scores)
    assert ok and err < 1e-6
    gl2 = gl.copy()
    gl2[2 * per + slots[-1]] = (gl2[2 * per + slots[-1]] + 1) % 5        # one wrong label in the LAST rank's slice
    assert not bench.check_known(gl2, gs, world, per, labels, scores)[0]
    gs2 = gs.copy()
    gs2[per + slots[3]] += 2e-3                                           # a score off by more than 1e-3
    assert not bench.check_known(gl, gs2, world, per, labels, scores)[0]


def test_roofline_traffic_comes_from_the_committed_ncu_summary():
    got = bench.conv2_traffic_from_profiles()
    assert got is not None, "no profiles/*ncu_full_summary*.txt with a 'sites per launch' line"
    per_site, name = got
    assert name.endswith(".txt") and os.path.exists(os.path.join(ROOT, "profiles", name))
    # conv2 + pool2 at 10 000 sites: x2 read once (~3.2 GB), pooled maxima written (~2.5 GB)
    assert 3.0e5 < per_site < 9.0e5


def test_reference_arm_runs_on_a_tiny_sample():
    env = dict(os.environ, SVX_REF_SAMPLE="128")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "sites/s"
    assert line["cpu_baseline"]["kind"] in ("reference+proxy", "port")
    if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "src", "network", "create_batch.py")):
        assert line["cpu_baseline"]["kind"] == "reference+proxy"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("configs[1]")
