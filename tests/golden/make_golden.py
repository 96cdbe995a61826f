"""Generates tests/golden/reference_hashes.json by running the compiled, unmodified reference (oracle/_ref, built
from /root/reference by oracle/Makefile) on the synthetic scenes of forge2d_b200/scenes.py. Run in the build
container, where /root/reference exists:   python tests/golden/make_golden.py
The hashes pin (a) the oracle library wherever it is loaded later (GPU box) and (b) the CUDA path at full sizes."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness as H  # noqa: E402
from forge2d_b200 import scenes  # noqa: E402

# scene, kwargs, frames at which a hash is recorded
CASES = [
    ("bench2d", {}, [1, 64, 160, 256, 512]),
    ("large_pyramid", {}, [1, 32, 128, 160, 256, 400]),  # 160-400: the impact and the collapse (bottom row lands at ~147)
    ("many_pyramids", {}, [1, 16, 48]),
    ("joint_grid", {}, [1, 16, 64]),
    ("falling_shapes", {"count": 24}, [1, 100, 240]),
    ("many_pyramids", {"grid": 3, "base": 6}, [1, 30, 120]),
    ("joint_grid", {"n": 12}, [1, 30, 120]),
    ("polygon_soup", {"count": 40}, [1, 60, 200]),
    ("joint_zoo", {"sets": 2}, [1, 60, 200]),          # every joint type
    ("chain_terrain", {"count": 36}, [1, 80, 240]),    # chain-segment manifolds
    ("sensor_field", {"count": 30}, [1, 80, 200]),     # sensor overlaps
]


def case_key(name, kw):
    return name + "".join("|%s=%s" % (k, kw[k]) for k in sorted(kw))


def main():
    ref = H.load("reference")
    out = {}
    for name, kw, frames in CASES:
        s = scenes.SCENES[name](ref, **kw)
        rec = {}
        for f in range(1, max(frames) + 1):
            s.step()
            if f in frames:
                snap = H.snapshot(ref, s.world, trees=False)
                rec[str(f)] = {"hash": H.state_hash(snap), "bodies": int(len(snap["bodies"])),
                               "contacts": int(len(snap["contacts"])), "awake": int(len(snap["awake_order"])),
                               "colors": [int(c) for c in snap["color_contact_counts"]]}
        s.destroy()
        out[case_key(name, kw)] = rec
        print(case_key(name, kw), {k: (v["contacts"], v["awake"]) for k, v in rec.items()})
    with open(os.path.join(HERE, "reference_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
