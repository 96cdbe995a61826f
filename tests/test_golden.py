"""Golden-vector tests. tests/golden/reference_hashes.json holds sha256 state hashes produced by the unmodified
reference in the build container (tests/golden/make_golden.py). CPU: the oracle library and the emulated step must
reproduce them (pins the oracle wherever it runs). GPU: the CUDA path must reproduce them at BASELINE.json's full
sizes (all five configurations), bit for bit."""
import json
import os

import pytest

import harness as H
from forge2d_b200 import scenes

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_hashes.json")))


def _parse(key):
    parts = key.split("|")
    kw = {}
    for p in parts[1:]:
        k, v = p.split("=")
        kw[k] = int(v)
    return parts[0], kw


def _run(lib, key, max_frame=None, mode=None):
    name, kw = _parse(key)
    rec = GOLDEN[key]
    frames = sorted(int(f) for f in rec if max_frame is None or int(f) <= max_frame)
    s = scenes.SCENES[name](lib, **kw)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(s.world, mode)
    for f in range(1, frames[-1] + 1):
        s.step()
        if f in frames:
            snap = H.snapshot(lib, s.world, trees=False)
            want = rec[str(f)]
            assert len(snap["contacts"]) == want["contacts"], (key, f)
            assert len(snap["awake_order"]) == want["awake"], (key, f)
            assert [int(c) for c in snap["color_contact_counts"]] == want["colors"], (key, f)
            assert H.state_hash(snap) == want["hash"], "%s frame %d: state hash differs from the reference" % (key, f)
    s.destroy()


SMALL = [k for k in GOLDEN if "|" in k]
FULL = [k for k in GOLDEN if "|" not in k]


@pytest.mark.parametrize("key", SMALL + ["bench2d"])
def test_oracle_reproduces_golden(ref, key):
    _run(ref, key, max_frame=256)


@pytest.mark.parametrize("key", SMALL + ["bench2d"])
def test_emulated_step_reproduces_golden(emu, key):
    _run(emu, key, max_frame=160)


@pytest.mark.gpu
@pytest.mark.parametrize("key", SMALL)
@pytest.mark.parametrize("mode", [0, 1])
def test_cuda_step_reproduces_golden_small(gpu, key, mode):
    _run(gpu, key, mode=mode)
    assert gpu.f2dGetLastError() == b""


@pytest.mark.gpu
@pytest.mark.parametrize("key", FULL)
def test_cuda_step_reproduces_golden_full_size(gpu, key):
    _run(gpu, key)
    assert gpu.f2dGetLastError() == b""
