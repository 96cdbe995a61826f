"""CPU tests of the drop-in boundary: the product library loads, exports every symbol include/forge2d_b200.h
declares, keeps the reference's by-value struct layouts, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

from forge2d_b200 import _abi as A
from forge2d_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "forge2d_b200.h")).read()
    return sorted(set(re.findall(r"F2D_API\s+[^;(]*?\b((?:b2|f2d)[A-Za-z_0-9]+)\s*\(", text)))


def test_header_declares_step_path():
    names = _declared_symbols()
    for required in ("b2World_Step", "b2CreateWorld", "b2CreateBody", "b2CreatePolygonShape", "b2World_GetContactEvents",
                     "b2World_GetBodyEvents", "b2World_GetSensorEvents", "f2dBatch_Step"):
        assert required in names


def test_library_exports_every_declared_symbol(product):
    dll = C.CDLL(product.path)
    missing = [n for n in _declared_symbols() if not hasattr(dll, n)]
    assert missing == []
    assert product.missing == []


def test_ffi_symbols_resolve_by_name_from_c(product, tmp_path):
    """The binding forge2d generates is `@Native` lookups by symbol name in the loaded asset (box2d.g.dart): a C
    program dlopens the product and dlsyms each of the 257 names raw_box2d_ffi.dart calls (tests/golden/ffi_symbols.txt)."""
    import subprocess
    exe = str(tmp_path / "dlsym_check")
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "native", "dlsym_check.c"), "-ldl"])
    out = subprocess.run([exe, product.path, os.path.join(ROOT, "tests", "golden", "ffi_symbols.txt")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "resolved 257 of 257" in out.stdout


def test_profile_is_exported_with_the_reference_layout(product):
    assert hasattr(C.CDLL(product.path), "b2World_GetProfile")
    assert C.sizeof(A.Profile) == 22 * 4  # types.h:466-490: 22 floats


def test_struct_layouts_match_reference_abi():
    # sizes from the reference headers compiled with gcc x86-64 (SURVEY §2.2 probe, B2/include/box2d/*.h)
    assert C.sizeof(A.WorldId) == 4 and C.sizeof(A.BodyId) == 8 and C.sizeof(A.ShapeId) == 8
    assert C.sizeof(A.Polygon) == 144
    assert C.sizeof(A.ManifoldPoint) == 48 and C.sizeof(A.Manifold) == 112
    assert C.sizeof(A.BodyMoveEvent) == 40
    assert C.sizeof(A.ContactBeginTouchEvent) == 128
    assert C.sizeof(A.ContactHitEvent) == 36
    assert C.sizeof(A.Filter) == 24


def test_defaults_match_reference(product, ref):
    for name in ("b2DefaultWorldDef", "b2DefaultBodyDef", "b2DefaultShapeDef", "b2DefaultRevoluteJointDef"):
        a, b = getattr(product, name)(), getattr(ref, name)()
        for field, _ in a._fields_:
            va, vb = getattr(a, field), getattr(b, field)
            if isinstance(va, C.Structure):
                assert bytes(va) == bytes(vb), (name, field)
            else:
                assert va == vb, (name, field)


def test_geometry_helpers_bit_exact(product, ref):
    for lib_args in ((0.5, 0.5, A.Vec2(0.0, 0.0), A.Rot(1.0, 0.0), 0.0), (0.3, 0.2, A.Vec2(0.25, -1.5), A.Rot(0.8, 0.6), 0.1)):
        assert bytes(product.b2MakeOffsetRoundedBox(*lib_args)) == bytes(ref.b2MakeOffsetRoundedBox(*lib_args))
    pts = (A.Vec2 * 6)(A.Vec2(0, 0), A.Vec2(2, 0), A.Vec2(2.5, 1.5), A.Vec2(1, 3), A.Vec2(-0.5, 1.25), A.Vec2(1, 1))
    ha, hb = product.b2ComputeHull(pts, 6), ref.b2ComputeHull(pts, 6)
    assert ha.count == hb.count and bytes(ha)[: 8 * ha.count] == bytes(hb)[: 8 * hb.count]
    assert bytes(product.b2MakePolygon(C.byref(ha), 0.05)) == bytes(ref.b2MakePolygon(C.byref(hb), 0.05))


def test_world_construction_matches_reference_without_gpu(product, ref):
    """Host-side world building (ids, islands, proxies, move buffer, mass data) needs no device and must agree with
    the reference before the first step."""
    import harness as H
    a = scenes.falling_shapes(ref, count=12)
    b = scenes.falling_shapes(product, count=12)
    assert H.diff(H.snapshot(ref, a.world), H.snapshot(product, b.world)) == []
    a.destroy()
    b.destroy()


def test_step_fails_loudly_without_device(product):
    if product.f2dHasDevice():
        pytest.skip("a CUDA device is present")
    product.f2dClearLastError()
    s = scenes.bench2d(product, rows=2)
    s.step()
    err = product.f2dGetLastError().decode()
    assert "no CUDA device" in err and "no CPU fallback" in err
    # nothing was simulated: the boxes have not moved
    assert product.b2Body_GetPosition(s.bodies[1]).y == pytest.approx(0.75)
    assert product.f2dBatch_Create(s.world, 4) is None
    s.destroy()


def test_world_id_generations(product):
    wd = product.b2DefaultWorldDef()
    w1 = product.b2CreateWorld(C.byref(wd))
    assert product.b2World_IsValid(w1)
    product.b2DestroyWorld(w1)
    assert not product.b2World_IsValid(w1)
    w2 = product.b2CreateWorld(C.byref(wd))
    assert w2.index1 == w1.index1 and w2.generation == w1.generation + 1
    g = A.Vec2(1.5, -3.25)
    product.b2World_SetGravity(w2, g)
    out = product.b2World_GetGravity(w2)
    assert (out.x, out.y) == (1.5, -3.25)
    product.b2DestroyWorld(w2)
