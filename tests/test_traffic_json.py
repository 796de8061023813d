"""profiles/traffic.json -- what bench.py reports as roofline.traffic / dram_frac -- must be (a) the sum that
tools/make_traffic.py makes of the committed per-launch ncu lists (profiles/r02_traffic_<kind>.csv) and
(b) keyed to the kernel sources that are shipped: a change to walt_core.cuh / walt_engine.cu without a new
ncu pass makes bench.py drop the figures (it prints traffic: null), and this test says so first."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_traffic_json_is_the_sum_of_the_committed_launch_lists():
    import bench
    import make_traffic as mt
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for kind in mt.KINDS:
        assert kind in t, kind
        path = os.path.join(ROOT, t[kind]["source"])
        assert os.path.exists(path), path
        ls = mt.launches(path)
        paired = kind in ("pe", "pe_stress")
        steps = mt.steps_of(ls, paired)
        if paired:
            assert len(steps) % mt.N_STEPS == 0
            per = len(steps) // mt.N_STEPS
            steps = [sum(steps[i * per:(i + 1) * per], []) for i in range(mt.N_STEPS)]
        step = steps[3]
        if kind == "verify":
            step = [l for l in step if "verify_kernel" in l["name"]]
        assert int(sum(l["read"] + l["write"] for l in step)) == t[kind]["dram_bytes_per_step"], kind
        assert len(step) == t[kind]["launches_per_step"]
        gmb, n, _ = bench.FULL_SIZE[kind]
        assert (t[kind]["reads_per_step"], t[kind]["genome_mb"]) == (n, gmb)
        # every step of the capture did the same work: the launch lists of steps 2..4 have the same kernels
        names = [[l["name"] for l in s] for s in steps[1:4]]
        assert names[0] == names[1] == names[2], kind


def test_traffic_json_belongs_to_the_shipped_kernel_sources():
    import bench
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    sha = bench.kernel_source_hash()
    stale = [k for k, v in t.items() if isinstance(v, dict) and v.get("source_sha16") != sha]
    assert not stale, f"kernel sources changed since the ncu pass ({sha}): re-run tools/gpu_round.sh for {stale}"
    assert bench.committed_traffic("se", *bench.FULL_SIZE["se"][1::-1]) is not None
