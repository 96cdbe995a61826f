"""Generates the per-joint-type accessor surface of the C ABI from one table:

  forge2d_b200/csrc/f2d_capi_joints.inl   C definitions (included by f2d_capi.inl inside extern "C")
  include/forge2d_b200_joints.h           declarations (included by include/forge2d_b200.h)
  forge2d_b200/_abi_joints.py             ctypes signatures for the Python mirror / tests

Every accessor of B2/src/{distance,motor,mouse,prismatic,revolute,weld,wheel}_joint.c is a field read, a field write
(sometimes clamped) or a flag change that clears the impulses the flag governs; the table states which, citing the
reference function it mirrors. The handful of computed getters are written out by hand below.
Run from the repo root:  python tools/gen_joint_api.py
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (kind, Name, field, extra)
#   getf/getb/getv : return J.field                     setf/setv : J.field = extra.format(v=value) or value
#   enable         : J.field = flag; extra = (impulse fields cleared, only-when-changed?)
#   invh           : return world.inv_h * J.field
T = {
    "Distance": ("distance", "kDistanceJoint", "distance_joint.c:20-243", [
        ("special", "SetLength", None, None), ("getf", "GetLength", "length", None),
        ("enable", "EnableLimit", "enableLimit", ([], False)), ("getb", "IsLimitEnabled", "enableLimit", None),
        ("special", "SetLengthRange", None, None), ("getf", "GetMinLength", "minLength", None),
        ("getf", "GetMaxLength", "maxLength", None), ("special", "GetCurrentLength", None, None),
        ("enable", "EnableSpring", "enableSpring", ([], False)), ("getb", "IsSpringEnabled", "enableSpring", None),
        ("setf", "SetSpringHertz", "hertz", None), ("setf", "SetSpringDampingRatio", "dampingRatio", None),
        ("getf", "GetSpringHertz", "hertz", None), ("getf", "GetSpringDampingRatio", "dampingRatio", None),
        ("enable", "EnableMotor", "enableMotor", (["motorImpulse"], True)), ("getb", "IsMotorEnabled", "enableMotor", None),
        ("setf", "SetMotorSpeed", "motorSpeed", None), ("getf", "GetMotorSpeed", "motorSpeed", None),
        ("invh", "GetMotorForce", "motorImpulse", None),
        ("setf", "SetMaxMotorForce", "maxMotorForce", None), ("getf", "GetMaxMotorForce", "maxMotorForce", None),
    ]),
    "Motor": ("motor", "kMotorJoint", "motor_joint.c:14-72", [
        ("setv", "SetLinearOffset", "linearOffset", None), ("getv", "GetLinearOffset", "linearOffset", None),
        ("setf", "SetAngularOffset", "angularOffset", None), ("getf", "GetAngularOffset", "angularOffset", None),
        ("setf", "SetMaxForce", "maxForce", "maxf( 0.0f, {v} )"), ("getf", "GetMaxForce", "maxForce", None),
        ("setf", "SetMaxTorque", "maxTorque", "maxf( 0.0f, {v} )"), ("getf", "GetMaxTorque", "maxTorque", None),
        ("setf", "SetCorrectionFactor", "correctionFactor", "clampf( {v}, 0.0f, 1.0f )"),
        ("getf", "GetCorrectionFactor", "correctionFactor", None),
    ]),
    "Mouse": ("mouse", "kMouseJoint", "mouse_joint.c:14-72", [
        ("setv", "SetTarget", "targetA", None), ("getv", "GetTarget", "targetA", None),
        ("setf", "SetSpringHertz", "hertz", None), ("getf", "GetSpringHertz", "hertz", None),
        ("setf", "SetSpringDampingRatio", "dampingRatio", None), ("getf", "GetSpringDampingRatio", "dampingRatio", None),
        ("setf", "SetMaxForce", "maxForce", None), ("getf", "GetMaxForce", "maxForce", None),
    ]),
    "Prismatic": ("prismatic", "kPrismaticJoint", "prismatic_joint.c:17-251", [
        ("enable", "EnableSpring", "enableSpring", (["springImpulse"], True)), ("getb", "IsSpringEnabled", "enableSpring", None),
        ("setf", "SetSpringHertz", "hertz", None), ("getf", "GetSpringHertz", "hertz", None),
        ("setf", "SetSpringDampingRatio", "dampingRatio", None), ("getf", "GetSpringDampingRatio", "dampingRatio", None),
        ("setf", "SetTargetTranslation", "targetTranslation", None), ("getf", "GetTargetTranslation", "targetTranslation", None),
        ("enable", "EnableLimit", "enableLimit", (["lowerImpulse", "upperImpulse"], True)),
        ("getb", "IsLimitEnabled", "enableLimit", None),
        ("getf", "GetLowerLimit", "lowerTranslation", None), ("getf", "GetUpperLimit", "upperTranslation", None),
        ("limits", "SetLimits", ("lowerTranslation", "upperTranslation"), None),
        ("enable", "EnableMotor", "enableMotor", (["motorImpulse"], True)), ("getb", "IsMotorEnabled", "enableMotor", None),
        ("setf", "SetMotorSpeed", "motorSpeed", None), ("getf", "GetMotorSpeed", "motorSpeed", None),
        ("invh", "GetMotorForce", "motorImpulse", None),
        ("setf", "SetMaxMotorForce", "maxMotorForce", None), ("getf", "GetMaxMotorForce", "maxMotorForce", None),
        ("special", "GetTranslation", None, None), ("special", "GetSpeed", None, None),
    ]),
    "Revolute": ("revolute", "kRevoluteJoint", "revolute_joint.c:17-186", [
        ("enable", "EnableSpring", "enableSpring", (["springImpulse"], True)), ("getb", "IsSpringEnabled", "enableSpring", None),
        ("setf", "SetSpringHertz", "hertz", None), ("getf", "GetSpringHertz", "hertz", None),
        ("setf", "SetSpringDampingRatio", "dampingRatio", None), ("getf", "GetSpringDampingRatio", "dampingRatio", None),
        ("setf", "SetTargetAngle", "targetAngle", None), ("getf", "GetTargetAngle", "targetAngle", None),
        ("special", "GetAngle", None, None),
        ("enable", "EnableLimit", "enableLimit", (["lowerImpulse", "upperImpulse"], True)),
        ("getb", "IsLimitEnabled", "enableLimit", None),
        ("getf", "GetLowerLimit", "lowerAngle", None), ("getf", "GetUpperLimit", "upperAngle", None),
        ("limits", "SetLimits", ("lowerAngle", "upperAngle"), None),
        ("enable", "EnableMotor", "enableMotor", (["motorImpulse"], True)), ("getb", "IsMotorEnabled", "enableMotor", None),
        ("setf", "SetMotorSpeed", "motorSpeed", None), ("getf", "GetMotorSpeed", "motorSpeed", None),
        ("invh", "GetMotorTorque", "motorImpulse", None),
        ("setf", "SetMaxMotorTorque", "maxMotorTorque", None), ("getf", "GetMaxMotorTorque", "maxMotorTorque", None),
    ]),
    "Weld": ("weld", "kWeldJoint", "weld_joint.c:14-60", [
        ("setf", "SetLinearHertz", "linearHertz", None), ("getf", "GetLinearHertz", "linearHertz", None),
        ("setf", "SetLinearDampingRatio", "linearDampingRatio", None), ("getf", "GetLinearDampingRatio", "linearDampingRatio", None),
        ("setf", "SetAngularHertz", "angularHertz", None), ("getf", "GetAngularHertz", "angularHertz", None),
        ("setf", "SetAngularDampingRatio", "angularDampingRatio", None),
        ("getf", "GetAngularDampingRatio", "angularDampingRatio", None),
    ]),
    "Wheel": ("wheel", "kWheelJoint", "wheel_joint.c:17-215", [
        ("enable", "EnableSpring", "enableSpring", (["springImpulse"], True)), ("getb", "IsSpringEnabled", "enableSpring", None),
        ("setf", "SetSpringHertz", "hertz", None), ("getf", "GetSpringHertz", "hertz", None),
        ("setf", "SetSpringDampingRatio", "dampingRatio", None), ("getf", "GetSpringDampingRatio", "dampingRatio", None),
        ("enable", "EnableLimit", "enableLimit", (["lowerImpulse", "upperImpulse"], True)),
        ("getb", "IsLimitEnabled", "enableLimit", None),
        ("getf", "GetLowerLimit", "lowerTranslation", None), ("getf", "GetUpperLimit", "upperTranslation", None),
        ("limits", "SetLimits", ("lowerTranslation", "upperTranslation"), None),
        ("enable", "EnableMotor", "enableMotor", (["motorImpulse"], True)), ("getb", "IsMotorEnabled", "enableMotor", None),
        ("setf", "SetMotorSpeed", "motorSpeed", None), ("getf", "GetMotorSpeed", "motorSpeed", None),
        ("invh", "GetMotorTorque", "motorImpulse", None),
        ("setf", "SetMaxMotorTorque", "maxMotorTorque", None), ("getf", "GetMaxMotorTorque", "maxMotorTorque", None),
    ]),
}

SPECIAL = {
    "b2DistanceJoint_SetLength": ("void", "float length", """	JointSim* s = jointSimOfType( jointId, kDistanceJoint, nullptr, true );
	if ( s == nullptr )
		return;
	DistanceJointData& j = s->distance;
	j.length = clampf( length, kLinearSlop, kHuge );
	j.impulse = 0.0f;
	j.lowerImpulse = 0.0f;
	j.upperImpulse = 0.0f;
"""),
    "b2DistanceJoint_SetLengthRange": ("void", "float minLength, float maxLength", """	JointSim* s = jointSimOfType( jointId, kDistanceJoint, nullptr, true );
	if ( s == nullptr )
		return;
	DistanceJointData& j = s->distance;
	minLength = clampf( minLength, kLinearSlop, kHuge );
	maxLength = clampf( maxLength, kLinearSlop, kHuge );
	j.minLength = minf( minLength, maxLength );
	j.maxLength = maxf( minLength, maxLength );
	j.impulse = 0.0f;
	j.lowerImpulse = 0.0f;
	j.upperImpulse = 0.0f;
"""),
    "b2DistanceJoint_GetCurrentLength": ("float", "", """	HostWorld* hw = nullptr;
	JointSim* s = jointSimOfType( jointId, kDistanceJoint, &hw, false );
	if ( s == nullptr || hw->img->locked )
		return 0.0f;
	const BodySim* sims = ptr( hw->img, hw->img->sims );
	V2 pA = xfPoint( sims[s->bodyIdA].transform, s->localOriginAnchorA );
	V2 pB = xfPoint( sims[s->bodyIdB].transform, s->localOriginAnchorB );
	return length( sub( pB, pA ) );
"""),
    "b2PrismaticJoint_GetTranslation": ("float", "", """	HostWorld* hw = nullptr;
	JointSim* s = jointSimOfType( jointId, kPrismaticJoint, &hw, false );
	if ( s == nullptr )
		return 0.0f;
	const BodySim* sims = ptr( hw->img, hw->img->sims );
	Xf xfA = sims[s->bodyIdA].transform, xfB = sims[s->bodyIdB].transform;
	V2 axisA = rotate( xfA.q, s->prismatic.localAxisA );
	V2 pA = xfPoint( xfA, s->localOriginAnchorA );
	V2 pB = xfPoint( xfB, s->localOriginAnchorB );
	return dot( sub( pB, pA ), axisA );
"""),
    "b2PrismaticJoint_GetSpeed": ("float", "", """	HostWorld* hw = nullptr;
	JointSim* s = jointSimOfType( jointId, kPrismaticJoint, &hw, false );
	if ( s == nullptr )
		return 0.0f;
	World* w = hw->img;
	const Body& bodyA = ptr( w, w->bodies )[s->bodyIdA];
	const Body& bodyB = ptr( w, w->bodies )[s->bodyIdB];
	const BodySim& simA = ptr( w, w->sims )[s->bodyIdA];
	const BodySim& simB = ptr( w, w->sims )[s->bodyIdB];
	const BodyState* stA = bodyA.setIndex == kAwakeSet ? ptr( w, w->states ) + bodyA.localIndex : nullptr;
	const BodyState* stB = bodyB.setIndex == kAwakeSet ? ptr( w, w->states ) + bodyB.localIndex : nullptr;
	V2 axisA = rotate( simA.transform.q, s->prismatic.localAxisA );
	V2 rA = rotate( simA.transform.q, sub( s->localOriginAnchorA, simA.localCenter ) );
	V2 rB = rotate( simB.transform.q, sub( s->localOriginAnchorB, simB.localCenter ) );
	V2 d = add( sub( simB.center, simA.center ), sub( rB, rA ) );
	V2 vA = stA ? stA->v : V2{ 0.0f, 0.0f };
	V2 vB = stB ? stB->v : V2{ 0.0f, 0.0f };
	float wA = stA ? stA->w : 0.0f;
	float wB = stB ? stB->w : 0.0f;
	V2 vRel = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );
	return dot( d, crossSV( wA, axisA ) ) + dot( axisA, vRel );
"""),
    "b2RevoluteJoint_GetAngle": ("float", "", """	HostWorld* hw = nullptr;
	JointSim* s = jointSimOfType( jointId, kRevoluteJoint, &hw, false );
	if ( s == nullptr )
		return 0.0f;
	const BodySim* sims = ptr( hw->img, hw->img->sims );
	float angle = relativeAngle( sims[s->bodyIdB].transform.q, sims[s->bodyIdA].transform.q ) - s->revolute.referenceAngle;
	return unwindAngle( angle );
"""),
}


def main():
    c, h, py = [], [], []
    c.append("// GENERATED by tools/gen_joint_api.py - per-joint-type accessors of the C ABI. Do not edit by hand.\n")
    h.append("// GENERATED by tools/gen_joint_api.py - per-joint-type accessor declarations (B2/include/box2d/box2d.h:770-1250).\n")
    py.append('"""GENERATED by tools/gen_joint_api.py - ctypes signatures of the per-joint-type accessors."""\n'
              "# (restype, argtypes) by NAME; forge2d_b200/_abi.py resolves the names to its ctypes classes\n"
              "JointId, Vec2, c_float, c_bool = \"JointId\", \"Vec2\", \"c_float\", \"c_bool\"\n\nJOINT_ACCESSORS = {\n")
    for tname, (member, kconst, cite, rows) in T.items():
        c.append("\n// ---- b2%sJoint_*: %s\n" % (tname, cite))
        h.append("\n// b2%sJoint: %s\n" % (tname, cite))
        for kind, name, field, extra in rows:
            sym = "b2%sJoint_%s" % (tname, name)
            get = "	JointSim* s = jointSimOfType( jointId, %s, nullptr, false );\n" % kconst
            put = "	JointSim* s = jointSimOfType( jointId, %s, nullptr, true );\n	if ( s == nullptr )\n		return;\n" % kconst
            J = "s->%s" % member
            if kind == "getf":
                c.append("float %s( b2JointId jointId )\n{\n%s	return s ? %s.%s : 0.0f;\n}\n" % (sym, get, J, field))
                h.append("F2D_API float %s( b2JointId jointId );\n" % sym)
                py.append('    "%s": (c_float, [JointId]),\n' % sym)
            elif kind == "getb":
                c.append("bool %s( b2JointId jointId )\n{\n%s	return s ? %s.%s : false;\n}\n" % (sym, get, J, field))
                h.append("F2D_API bool %s( b2JointId jointId );\n" % sym)
                py.append('    "%s": (c_bool, [JointId]),\n' % sym)
            elif kind == "getv":
                c.append("b2Vec2 %s( b2JointId jointId )\n{\n%s	return s ? b2Vec2{ %s.%s.x, %s.%s.y } : b2Vec2{ 0.0f, 0.0f };\n}\n" % (
                    sym, get, J, field, J, field))
                h.append("F2D_API b2Vec2 %s( b2JointId jointId );\n" % sym)
                py.append('    "%s": (Vec2, [JointId]),\n' % sym)
            elif kind == "setf":
                expr = (extra or "{v}").format(v="value")
                c.append("void %s( b2JointId jointId, float value )\n{\n%s	%s.%s = %s;\n}\n" % (sym, put, J, field, expr))
                h.append("F2D_API void %s( b2JointId jointId, float value );\n" % sym)
                py.append('    "%s": (None, [JointId, c_float]),\n' % sym)
            elif kind == "setv":
                c.append("void %s( b2JointId jointId, b2Vec2 value )\n{\n%s	%s.%s = V2{ value.x, value.y };\n}\n" % (sym, put, J, field))
                h.append("F2D_API void %s( b2JointId jointId, b2Vec2 value );\n" % sym)
                py.append('    "%s": (None, [JointId, Vec2]),\n' % sym)
            elif kind == "invh":
                c.append("float %s( b2JointId jointId )\n{\n	HostWorld* hw = nullptr;\n	JointSim* s = jointSimOfType( jointId, %s, &hw, false );\n"
                         "	return s ? hw->img->inv_h * %s.%s : 0.0f;\n}\n" % (sym, kconst, J, field))
                h.append("F2D_API float %s( b2JointId jointId );\n" % sym)
                py.append('    "%s": (c_float, [JointId]),\n' % sym)
            elif kind == "enable":
                resets, conditional = extra
                body = "	%s.%s = flag;\n" % (J, field) + "".join("	%s.%s = 0.0f;\n" % (J, r) for r in resets)
                if conditional:
                    body = "	if ( flag != %s.%s )\n	{\n%s	}\n" % (J, field, "".join("	" + ln + "\n" for ln in body.rstrip("\n").split("\n")))
                c.append("void %s( b2JointId jointId, bool flag )\n{\n%s%s}\n" % (sym, put, body))
                h.append("F2D_API void %s( b2JointId jointId, bool flag );\n" % sym)
                py.append('    "%s": (None, [JointId, c_bool]),\n' % sym)
            elif kind == "limits":
                lo, hi = field
                c.append("void %s( b2JointId jointId, float lower, float upper )\n{\n%s	if ( lower != %s.%s || upper != %s.%s )\n	{\n"
                         "		%s.%s = minf( lower, upper );\n		%s.%s = maxf( lower, upper );\n		%s.lowerImpulse = 0.0f;\n		%s.upperImpulse = 0.0f;\n	}\n}\n"
                         % (sym, put, J, lo, J, hi, J, lo, J, hi, J, J))
                h.append("F2D_API void %s( b2JointId jointId, float lower, float upper );\n" % sym)
                py.append('    "%s": (None, [JointId, c_float, c_float]),\n' % sym)
            elif kind == "special":
                ret, args, body = SPECIAL[sym]
                sig = "b2JointId jointId" + (", " + args if args else "")
                c.append("%s %s( %s )\n{\n%s}\n" % (ret, sym, sig, body))
                h.append("F2D_API %s %s( %s );\n" % (ret, sym, sig))
                pyargs = ["JointId"] + ["c_float" for _ in args.split(",") if _.strip()]
                py.append('    "%s": (%s, [%s]),\n' % (sym, "c_float" if ret == "float" else "None", ", ".join(pyargs)))
    py.append("}\n")
    open(os.path.join(ROOT, "forge2d_b200", "csrc", "f2d_capi_joints.inl"), "w").write("".join(c))
    open(os.path.join(ROOT, "include", "forge2d_b200_joints.h"), "w").write("".join(h))
    open(os.path.join(ROOT, "forge2d_b200", "_abi_joints.py"), "w").write("".join(py))
    n = sum(len(r[3]) for r in T.values())
    print("generated %d joint accessors" % n)


if __name__ == "__main__":
    main()
