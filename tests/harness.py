"""Parity-test helpers: load the three libraries behind one binding, snapshot internal state as numpy record arrays
and diff two snapshots field by field (ints bit-exact, floats bit-exact or within a stated tolerance)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from forge2d_b200 import _abi as A  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbox2d_ref.so")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libf2d_emu.so")
PRODUCT_SO = os.path.join(ROOT, "forge2d_b200", "csrc", "libforge2d_b200.so")

_cache = {}


def build_emu():
    src = os.path.join(ROOT, "tests", "emu", "f2d_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "forge2d_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "forge2d_b200", "csrc"))
                    if f.endswith((".h", ".inl"))]
    if os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
                           src, "-o", EMU_SO])


def load(kind):
    if kind in _cache:
        return _cache[kind]
    if kind == "reference":
        if not os.path.exists(REF_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        lib = A.Library(REF_SO, "reference")
    elif kind == "emu":
        build_emu()
        lib = A.Library(EMU_SO, "emu")
    elif kind == "product":
        lib = A.Library(PRODUCT_SO, "product")
    else:
        raise ValueError(kind)
    _cache[kind] = lib
    return lib


def have_reference():
    return os.path.exists(REF_SO) or os.path.isdir("/root/reference")


def _records(fn, world, rec_type, cap, *extra):
    buf = (rec_type * cap)()
    n = fn(world, *extra, buf, cap)
    if n > cap:
        return _records(fn, world, rec_type, n + 16, *extra)
    return np.frombuffer(buf, dtype=np.dtype(rec_type), count=n).copy() if n > 0 else np.zeros(0, dtype=np.dtype(rec_type))


def _ints(fn, world, cap, *extra):
    buf = (C.c_int * cap)()
    n = fn(world, *extra, buf, cap)
    if n > cap:
        return _ints(fn, world, n + 16, *extra)
    return np.array(buf[:n], dtype=np.int32)


def snapshot(lib, world, trees=True, cap=1 << 17):
    s = {}
    s["bodies"] = _records(lib.debug_bodies, world, A.BodyRecord, cap)
    s["contacts"] = _records(lib.debug_contacts, world, A.ContactRecord, cap)
    s["islands"] = _records(lib.debug_islands, world, A.IslandRecord, cap)
    s["shapes"] = _records(lib.debug_shapes, world, A.ShapeRecord, cap)
    s["joints"] = _records(lib.debug_joints, world, A.JointRecord, cap)
    s["awake_order"] = _ints(lib.debug_awake_order, world, cap)
    s["move_array"] = _ints(lib.debug_move_array, world, cap)
    s["awake_contacts"] = _ints(lib.debug_awake_contacts, world, cap)
    s["awake_islands"] = _ints(lib.debug_awake_islands, world, cap)
    cc = (C.c_int * 12)()
    jc = (C.c_int * 12)()
    lib.debug_color_counts(world, cc, jc)
    s["color_contact_counts"] = np.array(cc[:], dtype=np.int32)
    s["color_joint_counts"] = np.array(jc[:], dtype=np.int32)
    s["overflow_contacts"] = _ints(lib.debug_color_contacts, world, cap, 11)
    if trees:
        for t in range(3):
            s["tree%d" % t] = _records(lib.debug_tree, world, A.TreeLeafRecord, cap, t)
    return s


def events(lib, world):
    """Contact / body events of the last step as plain python tuples."""
    ce = lib.b2World_GetContactEvents(world)
    begin = [(ce.beginEvents[i].shapeIdA.index1, ce.beginEvents[i].shapeIdB.index1, ce.beginEvents[i].manifold.pointCount)
             for i in range(ce.beginCount)]
    end = [(ce.endEvents[i].shapeIdA.index1, ce.endEvents[i].shapeIdB.index1) for i in range(ce.endCount)]
    hit = [(ce.hitEvents[i].shapeIdA.index1, ce.hitEvents[i].shapeIdB.index1,
            np.float32(ce.hitEvents[i].approachSpeed).view(np.uint32)) for i in range(ce.hitCount)]
    be = lib.b2World_GetBodyEvents(world)
    moves = np.zeros((be.moveCount, 6), dtype=np.float32)
    for i in range(be.moveCount):
        m = be.moveEvents[i]
        moves[i] = (m.transform.p.x, m.transform.p.y, m.transform.q.c, m.transform.q.s, m.bodyId.index1, m.fellAsleep)
    se = lib.b2World_GetSensorEvents(world)
    sbegin = [(se.beginEvents[i].sensorShapeId.index1, se.beginEvents[i].visitorShapeId.index1,
               se.beginEvents[i].visitorShapeId.generation) for i in range(se.beginCount)]
    send = [(se.endEvents[i].sensorShapeId.index1, se.endEvents[i].visitorShapeId.index1,
             se.endEvents[i].visitorShapeId.generation) for i in range(se.endCount)]
    return {"begin": begin, "end": end, "hit": hit, "moves": moves, "sensor_begin": sbegin, "sensor_end": send}


FLOAT_KINDS = "f"


def diff(a, b, float_tol=0.0, skip=()):
    """Returns a list of human-readable mismatches between two snapshots. Ints and record counts must be equal;
    floats must be bit-identical when float_tol == 0 (treating -0.0 == +0.0), else within relative tolerance."""
    out = []
    for key in a:
        if key in skip:
            continue
        x, y = a[key], b[key]
        if x.shape != y.shape:
            out.append("%s: count %s vs %s" % (key, x.shape, y.shape))
            continue
        if x.dtype.names is None:
            if not np.array_equal(x, y):
                idx = np.nonzero(x != y)[0][:5]
                out.append("%s: differs at %s: %s vs %s" % (key, idx, x[idx], y[idx]))
            continue
        for name in x.dtype.names:
            xa, ya = x[name], y[name]
            if xa.dtype.kind in FLOAT_KINDS:
                if float_tol == 0.0:
                    bad = ~((xa == ya) | (np.isnan(xa) & np.isnan(ya)))
                else:
                    scale = np.maximum(np.maximum(np.abs(xa), np.abs(ya)), 1.0)
                    bad = np.abs(xa - ya) > float_tol * scale
            else:
                bad = xa != ya
            if bad.any():
                rows = np.nonzero(bad.reshape(len(x), -1).any(axis=1))[0]
                r = rows[0]
                out.append("%s.%s: %d rows differ, first row %d (id %s): %s vs %s" % (
                    key, name, len(rows), r, x[x.dtype.names[0]][r], xa[r], ya[r]))
    return out


def state_hash(snap, keys=("bodies", "contacts", "islands", "shapes", "joints", "awake_order", "move_array",
                           "awake_contacts", "awake_islands", "color_contact_counts", "color_joint_counts",
                           "overflow_contacts")):
    """sha256 over the records of a snapshot (floats canonicalised so that -0.0 hashes like +0.0)."""
    import hashlib
    h = hashlib.sha256()
    for key in keys:
        arr = snap[key]
        if arr.dtype.names is None:
            h.update(np.ascontiguousarray(arr).tobytes())
            continue
        for name in arr.dtype.names:
            col = np.ascontiguousarray(arr[name])
            if col.dtype.kind == "f":
                col = col + np.float32(0.0)
            h.update(col.tobytes())
    return h.hexdigest()


def record_draw(lib, world, bounds=None, **options):
    """Calls b2World_Draw with recording callbacks; returns the list of (primitive, arguments...) in emission order."""
    calls = []
    v = lambda p: (p.x, p.y)
    t = lambda x: (x.p.x, x.p.y, x.q.c, x.q.s)
    keep = [
        A.DrawPolygonFcn(lambda vs, n, color, ctx: calls.append(("polygon", tuple(v(vs[i]) for i in range(n)), color))),
        A.DrawSolidPolygonFcn(lambda xf, vs, n, r, color, ctx: calls.append(("solid_polygon", t(xf), tuple(v(vs[i]) for i in range(n)), r, color))),
        A.DrawCircleFcn(lambda c, r, color, ctx: calls.append(("circle", v(c), r, color))),
        A.DrawSolidCircleFcn(lambda xf, r, color, ctx: calls.append(("solid_circle", t(xf), r, color))),
        A.DrawSolidCapsuleFcn(lambda p1, p2, r, color, ctx: calls.append(("solid_capsule", v(p1), v(p2), r, color))),
        A.DrawSegmentFcn(lambda p1, p2, color, ctx: calls.append(("segment", v(p1), v(p2), color))),
        A.DrawTransformFcn(lambda xf, ctx: calls.append(("transform", t(xf)))),
        A.DrawPointFcn(lambda p, size, color, ctx: calls.append(("point", v(p), size, color))),
        A.DrawStringFcn(lambda p, s, color, ctx: calls.append(("string", v(p), bytes(s), color))),
    ]
    draw = lib.b2DefaultDebugDraw()
    (draw.DrawPolygonFcn, draw.DrawSolidPolygonFcn, draw.DrawCircleFcn, draw.DrawSolidCircleFcn, draw.DrawSolidCapsuleFcn,
     draw.DrawSegmentFcn, draw.DrawTransformFcn, draw.DrawPointFcn, draw.DrawStringFcn) = keep
    for name, value in options.items():
        assert name in A.DEBUG_DRAW_OPTIONS, name
        setattr(draw, name, value)
    if bounds is not None:
        draw.useDrawingBounds = True
        draw.drawingBounds = A.AABB(A.Vec2(bounds[0], bounds[1]), A.Vec2(bounds[2], bounds[3]))
    lib.b2World_Draw(world, C.byref(draw))
    return calls
