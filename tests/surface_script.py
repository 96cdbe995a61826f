"""TEST INFRASTRUCTURE: one scripted session against the b2* surface — builds a world, steps it, and between steps
calls mutators (teleport, impulses, destroy body / shape / joint, filters, mass data, sleep control, joint options) and
getters. The same script is run against the compiled reference and against the library under test; every observation
and every internal record must agree bit for bit (tests/test_surface.py)."""
import ctypes as C

import numpy as np

import harness as H
from forge2d_b200 import _abi as A
from forge2d_b200 import scenes


def _bits(x):
    return int(np.float32(x).view(np.uint32))


def _v(v):
    return (_bits(v.x), _bits(v.y))


def run(lib, frames=200, on_frame=None):
    """Returns (observations, snapshots): observations is a list of (label, value) with floats as bit patterns."""
    obs, snaps = [], []
    s = scenes.bench2d(lib, rows=8, ground_half_width=12.0)
    world = s.world
    bodies = s.bodies
    sd = lib.b2DefaultShapeDef()
    box = lib.b2MakeOffsetRoundedBox(0.4, 0.2, A.Vec2(0.0, 0.0), A.Rot(1.0, 0.0), 0.0)
    circle = A.Circle(A.Vec2(0.3, 0.0), 0.25)

    def body(x, y, btype=2, shapes=("box",)):
        bd = lib.b2DefaultBodyDef()
        bd.type = btype
        bd.position = A.Vec2(x, y)
        b = lib.b2CreateBody(world, C.byref(bd))
        ids = []
        for kind in shapes:
            if kind == "box":
                ids.append(lib.b2CreatePolygonShape(b, C.byref(sd), C.byref(box)))
            else:
                ids.append(lib.b2CreateCircleShape(b, C.byref(sd), C.byref(circle)))
        return b, ids

    # a pendulum chain (revolute), a slider (prismatic), a rod (distance) hanging from a static bar
    bar, _ = body(0.0, -20.0, btype=0, shapes=())
    chain = []
    prev = bar
    rj = []
    for k in range(4):
        b, _ = body(-6.0 + 0.9 * k, -20.0)
        d = lib.b2DefaultRevoluteJointDef()
        d.bodyIdA, d.bodyIdB = prev, b
        d.localAnchorA = A.Vec2(-6.45 if k == 0 else 0.45, 0.0)
        d.localAnchorB = A.Vec2(-0.45, 0.0)
        rj.append(lib.b2CreateRevoluteJoint(world, C.byref(d)))
        chain.append(b)
        prev = b
    slider, _ = body(3.0, -21.0)
    d = lib.b2DefaultPrismaticJointDef()
    d.bodyIdA, d.bodyIdB = bar, slider
    d.localAnchorA = A.Vec2(3.0, -1.0)
    d.localAxisA = A.Vec2(1.0, 0.0)
    d.enableLimit, d.lowerTranslation, d.upperTranslation = True, -2.0, 2.0
    pj = lib.b2CreatePrismaticJoint(world, C.byref(d))
    bob, bob_shapes = body(6.0, -22.0, shapes=("box", "circle"))
    d = lib.b2DefaultDistanceJointDef()
    d.bodyIdA, d.bodyIdB = bar, bob
    d.localAnchorA = A.Vec2(6.0, 0.0)
    d.length = 2.0
    dj = lib.b2CreateDistanceJoint(world, C.byref(d))
    pile = bodies[1:]

    def observe(tag):
        for name, b in (("pile0", pile[0]), ("pile5", pile[5]), ("chain1", chain[1]), ("slider", slider), ("bob", bob)):
            if not lib.b2Body_IsValid(b):
                obs.append((tag + name, "invalid"))
                continue
            md = lib.b2Body_GetMassData(b)
            obs.append((tag + name + ".mass", (_bits(md.mass), _v(md.center), _bits(md.rotationalInertia))))
            obs.append((tag + name + ".pos", _v(lib.b2Body_GetPosition(b))))
            obs.append((tag + name + ".vel", (_v(lib.b2Body_GetLinearVelocity(b)), _bits(lib.b2Body_GetAngularVelocity(b)))))
            obs.append((tag + name + ".wp", _v(lib.b2Body_GetWorldPoint(b, A.Vec2(0.25, -0.5)))))
            obs.append((tag + name + ".lp", _v(lib.b2Body_GetLocalPoint(b, A.Vec2(1.0, 2.0)))))
            obs.append((tag + name + ".flags", (lib.b2Body_IsAwake(b), lib.b2Body_IsBullet(b), lib.b2Body_IsFixedRotation(b),
                                                lib.b2Body_IsSleepEnabled(b), lib.b2Body_IsEnabled(b), lib.b2Body_GetJointCount(b),
                                                lib.b2Body_GetShapeCount(b))))
            obs.append((tag + name + ".damp", (_bits(lib.b2Body_GetLinearDamping(b)), _bits(lib.b2Body_GetAngularDamping(b)),
                                               _bits(lib.b2Body_GetGravityScale(b)), _bits(lib.b2Body_GetSleepThreshold(b)))))
        jarr = (A.JointId * 8)()
        n = lib.b2Body_GetJoints(bar, jarr, 8)
        obs.append((tag + "bar.joints", [jarr[i].index1 for i in range(n)]))
        sarr = (A.ShapeId * 4)()
        if lib.b2Body_IsValid(bob):
            n = lib.b2Body_GetShapes(bob, sarr, 4)
            obs.append((tag + "bob.shapes", [sarr[i].index1 for i in range(n)]))
            for i in range(n):
                obs.append((tag + "bob.shape%d" % i, (lib.b2Shape_GetType(sarr[i]), _bits(lib.b2Shape_GetDensity(sarr[i])),
                                                      _bits(lib.b2Shape_GetFriction(sarr[i])), lib.b2Shape_GetFilter(sarr[i]).maskBits,
                                                      lib.b2Shape_TestPoint(sarr[i], lib.b2Body_GetPosition(bob)))))
        for name, j in (("rj0", rj[0]), ("rj2", rj[2]), ("pj", pj), ("dj", dj)):
            if not lib.b2Joint_IsValid(j):
                obs.append((tag + name, "invalid"))
                continue
            obs.append((tag + name + ".force", (_v(lib.b2Joint_GetConstraintForce(j)), _bits(lib.b2Joint_GetConstraintTorque(j)))))
            obs.append((tag + name + ".meta", (lib.b2Joint_GetType(j), lib.b2Joint_GetBodyA(j).index1, lib.b2Joint_GetBodyB(j).index1,
                                               lib.b2Joint_GetCollideConnected(j), _v(lib.b2Joint_GetLocalAnchorA(j)))))
        obs.append((tag + "rj0.angle", (_bits(lib.b2RevoluteJoint_GetAngle(rj[0])), _bits(lib.b2RevoluteJoint_GetMotorTorque(rj[0])),
                                        lib.b2RevoluteJoint_IsLimitEnabled(rj[0]))))
        if lib.b2Joint_IsValid(pj):
            obs.append((tag + "pj.state", (_bits(lib.b2PrismaticJoint_GetTranslation(pj)), _bits(lib.b2PrismaticJoint_GetSpeed(pj)),
                                           _bits(lib.b2PrismaticJoint_GetMotorForce(pj)))))
        if lib.b2Joint_IsValid(dj):
            obs.append((tag + "dj.state", (_bits(lib.b2DistanceJoint_GetCurrentLength(dj)), _bits(lib.b2DistanceJoint_GetLength(dj)))))

    for f in range(frames):
        if f == 5:
            lib.b2Body_SetName(pile[0], b"first")
            obs.append(("name", lib.b2Body_GetName(pile[0])))
            lib.b2Body_ApplyLinearImpulse(pile[0], A.Vec2(3.0, 1.0), lib.b2Body_GetWorldPoint(pile[0], A.Vec2(0.1, 0.2)), True)
            lib.b2Body_ApplyAngularImpulse(pile[1], 0.5, True)
            lib.b2Body_ApplyForce(pile[2], A.Vec2(-50.0, 20.0), lib.b2Body_GetPosition(pile[2]), True)
            lib.b2Body_ApplyForceToCenter(pile[3], A.Vec2(0.0, 80.0), True)
            lib.b2Body_ApplyTorque(pile[4], 15.0, True)
            lib.b2Body_ApplyLinearImpulseToCenter(chain[3], A.Vec2(2.0, 0.0), True)
        if f == 12:
            lib.b2Body_SetTransform(pile[5], A.Vec2(4.0, -25.0), A.Rot(np.float32(0.8), np.float32(0.6)))
            lib.b2Body_SetLinearDamping(pile[6], 0.4)
            lib.b2Body_SetAngularDamping(pile[6], 0.7)
            lib.b2Body_SetGravityScale(pile[7], 0.25)
            lib.b2Body_SetBullet(pile[8], True)
            lib.b2Body_SetFixedRotation(chain[2], True)
            lib.b2RevoluteJoint_EnableLimit(rj[0], True)
            lib.b2RevoluteJoint_SetLimits(rj[0], -0.3, 0.4)
            lib.b2RevoluteJoint_EnableMotor(rj[1], True)
            lib.b2RevoluteJoint_SetMotorSpeed(rj[1], 1.5)
            lib.b2RevoluteJoint_SetMaxMotorTorque(rj[1], 8.0)
            lib.b2PrismaticJoint_EnableMotor(pj, True)
            lib.b2PrismaticJoint_SetMotorSpeed(pj, -1.0)
            lib.b2PrismaticJoint_SetMaxMotorForce(pj, 25.0)
            lib.b2DistanceJoint_EnableSpring(dj, True)
            lib.b2DistanceJoint_SetSpringHertz(dj, 2.0)
            lib.b2DistanceJoint_SetSpringDampingRatio(dj, 0.2)
        if f == 60:
            lib.b2DestroyBody(pile[9])                # a body inside the pile: touching contacts, island links
            lib.b2DestroyShape(bob_shapes[1], True)   # one of two shapes, mass recomputed
            lib.b2Shape_SetDensity(bob_shapes[0], 3.0, True)
            lib.b2Shape_SetFriction(bob_shapes[0], 0.1)
            lib.b2Shape_SetRestitution(bob_shapes[0], 0.5)
        if f == 75:
            lib.b2DestroyJoint(rj[2])                 # cuts the chain in two
            flt = lib.b2Shape_GetFilter(bob_shapes[0])
            flt.maskBits = 0xFFFE
            lib.b2Shape_SetFilter(bob_shapes[0], flt)   # mask change: proxy moved, contacts destroyed
            sarr = (A.ShapeId * 2)()
            lib.b2Body_GetShapes(pile[10], sarr, 2)
            flt = lib.b2Shape_GetFilter(sarr[0])
            flt.categoryBits = 4
            lib.b2Shape_SetFilter(sarr[0], flt)         # category change: proxy re-created
            md = lib.b2Body_GetMassData(pile[11])
            md.mass = np.float32(md.mass * 2.0)
            md.center = A.Vec2(0.1, 0.0)
            lib.b2Body_SetMassData(pile[11], md)
            lib.b2Joint_SetCollideConnected(rj[0], True)
            lib.b2Body_SetSleepThreshold(pile[12], 0.2)
        if f == 90:
            lib.b2Joint_SetCollideConnected(rj[0], False)
            lib.b2Body_ApplyMassFromShapes(pile[11])
            lib.b2Body_EnableSleep(pile[13], False)
        if f == 100:
            lib.b2Body_Disable(pile[15])              # contacts destroyed, proxies removed, joints (none) re-homed
            lib.b2Body_Disable(chain[0])              # a jointed body: its two joints move to the disabled list
            lib.b2Body_SetType(pile[16], 0)           # dynamic -> static
            lib.b2Body_SetType(slider, 1)             # dynamic -> kinematic, joint stays in the graph
        if f == 120:
            lib.b2Body_Enable(pile[15])
            lib.b2Body_Enable(chain[0])
            lib.b2Body_SetType(pile[16], 2)           # static -> dynamic
            lib.b2Body_SetType(slider, 2)
            lib.b2Body_SetType(bar, 2)                # the static anchor bar becomes dynamic: every joint re-coloured
        if f == 150:
            lib.b2Body_SetAwake(pile[14], False)      # force an island to sleep (split first if it lost constraints)
        if f == 160:
            lib.b2Body_SetAwake(pile[14], True)
            lib.b2Joint_WakeBodies(pj)
            lib.b2Body_SetLinearVelocity(slider, A.Vec2(1.0, 0.0))
        if on_frame is not None:
            on_frame(f, s)
        s.step()
        if f % 10 == 4 or f in (60, 61, 75, 76, 90, 100, 101, 120, 121, 150, 151, 160, 161):
            observe("f%d." % f)
            snaps.append((f, H.snapshot(lib, world)))
    ev = H.events(lib, world)
    obs.append(("events.end", ev["end"]))
    s.destroy()
    return obs, snaps
