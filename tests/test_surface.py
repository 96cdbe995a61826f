"""The wider b2* surface (body / shape / joint accessors and mutators, destruction, all joint types) against the
compiled reference: identical scripted sessions (tests/surface_script.py) and the joint_zoo scene, every observation
and every internal record bit-identical. CPU: the host emulation of the step templates; `-m gpu`: the CUDA product."""
import pytest

import harness as H
import surface_script
from forge2d_b200 import scenes


def _compare_sessions(ref, lib):
    obs_a, snaps_a = surface_script.run(ref)
    obs_b, snaps_b = surface_script.run(lib)
    assert len(obs_a) == len(obs_b)
    for (la, va), (lb, vb) in zip(obs_a, obs_b):
        assert la == lb
        assert va == vb, "observation %s: reference %r, ours %r" % (la, va, vb)
    for (fa, sa), (fb, sb) in zip(snaps_a, snaps_b):
        d = H.diff(sa, sb)
        assert d == [], "frame %d: %s" % (fa, d[:6])


def _joint_zoo(ref, lib, frames, every, mode=None):
    a = scenes.joint_zoo(ref, sets=3)
    b = scenes.joint_zoo(lib, sets=3)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    for f in range(frames):
        a.step()
        b.step()
        if f % every == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(lib, b.world))
            assert d == [], "frame %d: %s" % (f, d[:6])
            for ja, jb in zip(a.joints, b.joints):
                fa, fb = ref.b2Joint_GetConstraintForce(ja), lib.b2Joint_GetConstraintForce(jb)
                assert (fa.x, fa.y) == (fb.x, fb.y)
                assert ref.b2Joint_GetConstraintTorque(ja) == lib.b2Joint_GetConstraintTorque(jb)
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


def test_scripted_session_matches_reference_emu(ref, emu):
    _compare_sessions(ref, emu)


def test_all_joint_types_bit_identical_emu(ref, emu):
    _joint_zoo(ref, emu, 300, 5)


def test_unsupported_mutators_fail_loudly(emu):
    s = scenes.bench2d(emu, rows=2)
    emu.f2dClearLastError()
    emu.b2Body_SetType(s.bodies[1], 0)
    assert b"not supported" in emu.f2dGetLastError()
    s.destroy()


@pytest.mark.gpu
def test_scripted_session_matches_reference_gpu(ref, gpu):
    _compare_sessions(ref, gpu)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_all_joint_types_bit_identical_gpu(ref, gpu, mode):
    _joint_zoo(ref, gpu, 300, 5, mode)
