"""CPU parity tests of the HOST LOGIC: the team-parallel step templates of forge2d_b200/csrc (the code the CUDA
kernels instantiate) compiled for one host thread (tests/emu, test infrastructure) against the compiled, unmodified
reference (oracle/_ref). Every internal record must be bit-identical: bodies, contacts + manifolds (point counts,
feature ids, impulses), colour assignment, islands, move array, tree leaves, events. Tolerance: 0 ulp."""
import pytest

import harness as H
from forge2d_b200 import scenes

CASES = [
    # name, kwargs, frames, compare-every
    ("bench2d", dict(rows=12), 260, 13),          # free fall, impact (continuous/TOI), collapse
    ("bench2d", dict(rows=40), 40, 8),            # C1 at full size, first frames (pair creation for 820 proxies)
    ("falling_shapes", dict(count=24), 240, 6),   # every manifold function, restitution, rounded polygons
    ("many_pyramids", dict(grid=3, base=6), 120, 6),  # islands fall asleep (~frame 35) and stay asleep
    ("joint_grid", dict(n=12, rain_every=4), 120, 6),  # revolute joints + rain of mixed shapes
    ("large_pyramid", dict(rows=30), 60, 10),
    ("polygon_soup", dict(count=40), 220, 5),
    ("jointed_piles", dict(chains=6), 300, 5),      # islands with joints: merge, split (joint edges), sleep       # 3..8-gons, rounded polygons, capsules, circles on sloped segments
]


@pytest.mark.parametrize("name,kw,frames,every", CASES, ids=["%s-%s" % (c[0], "-".join(str(v) for v in c[1].values())) for c in CASES])
def test_emulated_step_is_bit_identical_to_reference(ref, emu, name, kw, frames, every):
    a = scenes.SCENES[name](ref, **kw)
    b = scenes.SCENES[name](emu, **kw)
    assert H.diff(H.snapshot(ref, a.world), H.snapshot(emu, b.world)) == []
    for f in range(frames):
        a.step()
        b.step()
        ea, eb = H.events(ref, a.world), H.events(emu, b.world)
        assert ea["begin"] == eb["begin"] and ea["end"] == eb["end"] and ea["hit"] == eb["hit"], "events, frame %d" % f
        assert (ea["moves"] == eb["moves"]).all(), "move events, frame %d" % f
        if f % every == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(emu, b.world))
            assert d == [], "frame %d: %s" % (f, d[:6])
    a.destroy()
    b.destroy()


def test_island_parallel_solve_path_is_used_and_exact(ref, emu):
    """25 small piles: the solver takes the island-by-island path (one group per island) and stays bit-identical."""
    import ctypes as C
    a = scenes.many_pyramids(ref, grid=5, base=5)
    b = scenes.many_pyramids(emu, grid=5, base=5)
    used = 0
    for f in range(50):
        a.step()
        b.step()
        info = (C.c_int * 8)()
        emu.f2dWorld_GetStepInfo(b.world, info, 8)
        used += info[0]
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(emu, b.world))
        assert d == [], "frame %d: %s" % (f, d[:6])
    assert used >= 20


def test_sleep_and_wake_sequence(ref, emu):
    """Everything falls asleep, then a velocity kick through the API wakes one island (b2WakeSolverSet order)."""
    import ctypes as C
    from forge2d_b200 import _abi as A
    a = scenes.many_pyramids(ref, grid=2, base=5)
    b = scenes.many_pyramids(emu, grid=2, base=5)
    for f in range(90):
        a.step()
        b.step()
    assert ref.b2World_GetAwakeBodyCount(a.world) == 0 and emu.b2World_GetAwakeBodyCount(b.world) == 0
    assert H.diff(H.snapshot(ref, a.world), H.snapshot(emu, b.world)) == []
    ref.b2Body_SetLinearVelocity(a.bodies[7], A.Vec2(3.0, 4.0))
    emu.b2Body_SetLinearVelocity(b.bodies[7], A.Vec2(3.0, 4.0))
    assert ref.b2World_GetAwakeBodyCount(a.world) == emu.b2World_GetAwakeBodyCount(b.world) > 0
    for f in range(40):
        a.step()
        b.step()
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(emu, b.world))
        assert d == [], "frame %d after wake: %s" % (f, d[:6])


def test_zero_dt_and_substep_variants(ref, emu):
    a = scenes.bench2d(ref, rows=8)
    b = scenes.bench2d(emu, rows=8)
    for f in range(60):
        dt, sub = ((0.0, 4) if f % 7 == 3 else (1.0 / 60.0, 1 + f % 5))
        a.step(dt, sub)
        b.step(dt, sub)
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(emu, b.world))
        assert d == [], "frame %d: %s" % (f, d[:6])


def test_batch_emulation_replicates_template(ref, emu):
    a = scenes.bench2d(ref, rows=6)
    b = scenes.bench2d(emu, rows=6)
    batch = emu.f2dBatch_Create(b.world, 3)
    assert emu.f2dBatch_GetWorldCount(batch) == 3
    for f in range(25):
        a.step()
    emu.f2dBatch_StepN(batch, scenes.TIME_STEP, scenes.SUB_STEPS, 25)
    assert emu.f2dBatch_GetErrorFlags(batch) == 0
    for index in (0, 2):
        emu.f2dBatch_DownloadWorld(batch, index, b.world)
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(emu, b.world))
        assert d == [], "batch world %d: %s" % (index, d[:6])
    emu.f2dBatch_Destroy(batch)
