// forge2d_b200 — b2World_Draw / b2DefaultDebugDraw: debug geometry of the device-resident state through the host
// callbacks of b2DebugDraw (types.h:1383-1460). Runs on the host image of the world (downloaded if the device copy is
// newer), like the queries: the protocol is one synchronous host callback per primitive.
//
// Reference: B2/src/world.c:814-1489 (b2DrawShape, DrawQueryCallback, b2DrawWithBounds, b2World_Draw), B2/src/joint.c:1523-1597
// (b2DrawJoint) and the per-type b2Draw*Joint functions (distance_joint.c:512-556, prismatic_joint.c:633-668,
// revolute_joint.c:490-550, wheel_joint.c:517-551). What is reproduced is the SEQUENCE of callback invocations and
// their arguments: bodies are visited set by set in set order (static, disabled, awake, sleeping sets by id) and in
// each set's array order, contacts colour by colour in colour-array order, joints / islands / names by id.
// Included by f2d_capi_ext.inl (inside extern "C", with `using namespace f2d`).

extern "C++" {
namespace
{

b2Vec2 pub( V2 v ) { return b2Vec2{ v.x, v.y }; }
b2Transform pubXf( Xf t ) { return b2Transform{ { t.p.x, t.p.y }, { t.q.c, t.q.s } }; }

const b2HexColor kGraphColors[kColorCount] = { b2_colorRed,		  b2_colorOrange,	 b2_colorYellow, b2_colorGreen,
												b2_colorCyan,	  b2_colorBlue,		 b2_colorViolet, b2_colorPink,
												b2_colorChocolate, b2_colorGoldenRod, b2_colorCoral,  b2_colorBlack };

// every body of every solver set, in the reference's set-array order (world.c:1179-1186)
template <class F> void forEachBodyBySet( World* w, F&& visit )
{
	const Arr<int32_t>* fixedSets[3] = { &w->staticBodies, &w->disabledBodies, &w->awakeBodies };
	for ( int setIndex = 0; setIndex < 3; ++setIndex )
	{
		const int32_t* list = ptr( w, *fixedSets[setIndex] );
		for ( int i = 0; i < fixedSets[setIndex]->count; ++i )
			visit( list[i] );
	}
	SolverSet* sets = ptr( w, w->sets );
	for ( int setIndex = kFirstSleepingSet; setIndex < w->sets.count; ++setIndex )
	{
		SolverSet& set = sets[setIndex];
		if ( set.setIndex == kNull )
			continue;
		const int32_t* list = setBodyList( w, set );
		for ( int i = 0; i < set.bodyCount; ++i )
			visit( list[i] );
	}
}

// world.c:1196-1241 and :886-934 (same chain in both)
b2HexColor shapeDrawColor( const Shape& shape, const Body& body, const BodySim& sim )
{
	if ( shape.customColor != 0 )
		return (b2HexColor)shape.customColor;
	if ( body.type == kDynamicBody && body.mass == 0.0f )
		return b2_colorRed;
	if ( body.setIndex == kDisabledSet )
		return b2_colorSlateGray;
	if ( shape.sensorIndex != kNull )
		return b2_colorWheat;
	if ( sim.isBullet && body.setIndex == kAwakeSet )
		return b2_colorTurquoise;
	if ( body.isSpeedCapped )
		return b2_colorYellow;
	if ( sim.isFast )
		return b2_colorSalmon;
	if ( body.type == kStaticBody )
		return b2_colorPaleGreen;
	if ( body.type == kKinematicBody )
		return b2_colorRoyalBlue;
	if ( body.setIndex == kAwakeSet )
		return b2_colorPink;
	return b2_colorGray;
}

// world.c:814-866
void drawShape( b2DebugDraw* draw, const Shape& shape, Xf xf, b2HexColor color )
{
	switch ( shape.type )
	{
		case kCapsule:
			draw->DrawSolidCapsuleFcn( pub( xfPoint( xf, shape.capsule.c1 ) ), pub( xfPoint( xf, shape.capsule.c2 ) ), shape.capsule.radius, color,
									   draw->context );
			break;
		case kCircle:
			xf.p = xfPoint( xf, shape.circle.center );
			draw->DrawSolidCircleFcn( pubXf( xf ), shape.circle.radius, color, draw->context );
			break;
		case kPolygon:
			draw->DrawSolidPolygonFcn( pubXf( xf ), reinterpret_cast<const b2Vec2*>( shape.polygon.v ), shape.polygon.count, shape.polygon.radius,
									   color, draw->context );
			break;
		case kSegment:
			draw->DrawSegmentFcn( pub( xfPoint( xf, shape.segment.p1 ) ), pub( xfPoint( xf, shape.segment.p2 ) ), color, draw->context );
			break;
		case kChainSegment:
		{
			V2 p1 = xfPoint( xf, shape.chainSegment.segment.p1 );
			V2 p2 = xfPoint( xf, shape.chainSegment.segment.p2 );
			draw->DrawSegmentFcn( pub( p1 ), pub( p2 ), color, draw->context );
			draw->DrawPointFcn( pub( p2 ), 4.0f, color, draw->context );
			draw->DrawSegmentFcn( pub( p1 ), pub( lerp( p1, p2, 0.1f ) ), b2_colorPaleGreen, draw->context );
		}
		break;
		default:
			break;
	}
}

void drawBox( b2DebugDraw* draw, Box aabb, b2HexColor color )
{
	b2Vec2 vs[4] = { { aabb.lo.x, aabb.lo.y }, { aabb.hi.x, aabb.lo.y }, { aabb.hi.x, aabb.hi.y }, { aabb.lo.x, aabb.hi.y } };
	draw->DrawPolygonFcn( vs, 4, color, draw->context );
}

void drawSegment( b2DebugDraw* draw, V2 a, V2 b, b2HexColor color ) { draw->DrawSegmentFcn( pub( a ), pub( b ), color, draw->context ); }
void drawPoint( b2DebugDraw* draw, V2 p, float size, b2HexColor color ) { draw->DrawPointFcn( pub( p ), size, color, draw->context ); }

// distance_joint.c:512-556
void drawDistanceJoint( b2DebugDraw* draw, const JointSim& base, Xf transformA, Xf transformB )
{
	const DistanceJointData& joint = base.distance;
	V2 pA = xfPoint( transformA, base.localOriginAnchorA );
	V2 pB = xfPoint( transformB, base.localOriginAnchorB );
	V2 axis = normalize( sub( pB, pA ) );
	if ( joint.minLength < joint.maxLength && joint.enableLimit )
	{
		V2 pMin = mulAdd( pA, joint.minLength, axis );
		V2 pMax = mulAdd( pA, joint.maxLength, axis );
		V2 offset = mulSV( 0.05f * 1.0f, rightPerp( axis ) ); // b2_lengthUnitsPerMeter == 1
		if ( joint.minLength > kLinearSlop )
			drawSegment( draw, sub( pMin, offset ), add( pMin, offset ), b2_colorLightGreen );
		if ( joint.maxLength < kHuge )
			drawSegment( draw, sub( pMax, offset ), add( pMax, offset ), b2_colorRed );
		if ( joint.minLength > kLinearSlop && joint.maxLength < kHuge )
			drawSegment( draw, pMin, pMax, b2_colorGray );
	}
	drawSegment( draw, pA, pB, b2_colorWhite );
	drawPoint( draw, pA, 4.0f, b2_colorWhite );
	drawPoint( draw, pB, 4.0f, b2_colorWhite );
	if ( joint.hertz > 0.0f && joint.enableSpring )
		drawPoint( draw, mulAdd( pA, joint.length, axis ), 4.0f, b2_colorBlue );
}

// prismatic_joint.c:633-668 and wheel_joint.c:517-551 share one shape; only the colour of pB and of the pA-pB segment differ
void drawSliderJoint( b2DebugDraw* draw, const JointSim& base, Xf transformA, Xf transformB, V2 localAxisA, bool enableLimit,
					  float lowerTranslation, float upperTranslation, b2HexColor linkColor, b2HexColor pointBColor )
{
	V2 pA = xfPoint( transformA, base.localOriginAnchorA );
	V2 pB = xfPoint( transformB, base.localOriginAnchorB );
	V2 axis = rotate( transformA.q, localAxisA );
	drawSegment( draw, pA, pB, linkColor );
	if ( enableLimit )
	{
		V2 lower = mulAdd( pA, lowerTranslation, axis );
		V2 upper = mulAdd( pA, upperTranslation, axis );
		V2 perp = leftPerp( axis );
		drawSegment( draw, lower, upper, b2_colorGray );
		drawSegment( draw, mulSub( lower, 0.1f, perp ), mulAdd( lower, 0.1f, perp ), b2_colorGreen );
		drawSegment( draw, mulSub( upper, 0.1f, perp ), mulAdd( upper, 0.1f, perp ), b2_colorRed );
	}
	else
	{
		drawSegment( draw, mulSub( pA, 1.0f, axis ), mulAdd( pA, 1.0f, axis ), b2_colorGray );
	}
	drawPoint( draw, pA, 5.0f, b2_colorGray );
	drawPoint( draw, pB, 5.0f, pointBColor );
}

// revolute_joint.c:490-550
void drawRevoluteJoint( b2DebugDraw* draw, const JointSim& base, Xf transformA, Xf transformB, float drawSize )
{
	const RevoluteJointData& joint = base.revolute;
	V2 pA = xfPoint( transformA, base.localOriginAnchorA );
	V2 pB = xfPoint( transformB, base.localOriginAnchorB );
	const float L = drawSize;
	draw->DrawCircleFcn( pub( pB ), L, b2_colorGray, draw->context );
	float angle = relativeAngle( transformB.q, transformA.q );
	Rot rot = makeRotAngle( angle );
	V2 r = { L * rot.c, L * rot.s };
	V2 pC = add( pB, r );
	drawSegment( draw, pB, pC, b2_colorGray );
	if ( draw->drawJointExtras )
	{
		float jointAngle = unwindAngle( angle - joint.referenceAngle );
		char buffer[32];
		snprintf( buffer, 32, " %.1f deg", 180.0f * jointAngle / kPi );
		draw->DrawStringFcn( pub( pC ), buffer, b2_colorWhite, draw->context );
	}
	float lowerAngle = joint.lowerAngle + joint.referenceAngle;
	float upperAngle = joint.upperAngle + joint.referenceAngle;
	if ( joint.enableLimit )
	{
		Rot rotLo = makeRotAngle( lowerAngle );
		V2 rlo = { L * rotLo.c, L * rotLo.s };
		Rot rotHi = makeRotAngle( upperAngle );
		V2 rhi = { L * rotHi.c, L * rotHi.s };
		drawSegment( draw, pB, add( pB, rlo ), b2_colorGreen );
		drawSegment( draw, pB, add( pB, rhi ), b2_colorRed );
		Rot rotRef = makeRotAngle( joint.referenceAngle );
		V2 ref = { L * rotRef.c, L * rotRef.s };
		drawSegment( draw, pB, add( pB, ref ), b2_colorBlue );
	}
	drawSegment( draw, transformA.p, pA, b2_colorGold );
	drawSegment( draw, pA, pB, b2_colorGold );
	drawSegment( draw, transformB.p, pB, b2_colorGold );
}

// joint.c:1523-1597
void drawJoint( b2DebugDraw* draw, World* w, const Joint& joint )
{
	const Body* bodies = ptr( w, w->bodies );
	const Body& bodyA = bodies[joint.edges[0].bodyId];
	const Body& bodyB = bodies[joint.edges[1].bodyId];
	if ( bodyA.setIndex == kDisabledSet || bodyB.setIndex == kDisabledSet )
		return;
	const JointSim& jointSim = ptr( w, w->jointSims )[joint.jointId];
	Xf transformA = ptr( w, w->sims )[bodyA.id].transform;
	Xf transformB = ptr( w, w->sims )[bodyB.id].transform;
	V2 pA = xfPoint( transformA, jointSim.localOriginAnchorA );
	V2 pB = xfPoint( transformB, jointSim.localOriginAnchorB );
	switch ( joint.type )
	{
		case kDistanceJoint:
			drawDistanceJoint( draw, jointSim, transformA, transformB );
			break;
		case kMouseJoint:
		{
			V2 target = jointSim.mouse.targetA;
			drawPoint( draw, target, 4.0f, b2_colorGreen );
			drawPoint( draw, pB, 4.0f, b2_colorGreen );
			drawSegment( draw, target, pB, b2_colorLightGray );
		}
		break;
		case kFilterJoint:
			drawSegment( draw, pA, pB, b2_colorGold );
			break;
		case kPrismaticJoint:
			drawSliderJoint( draw, jointSim, transformA, transformB, jointSim.prismatic.localAxisA, jointSim.prismatic.enableLimit,
							 jointSim.prismatic.lowerTranslation, jointSim.prismatic.upperTranslation, b2_colorDimGray, b2_colorBlue );
			break;
		case kRevoluteJoint:
			drawRevoluteJoint( draw, jointSim, transformA, transformB, joint.drawSize );
			break;
		case kWheelJoint:
			drawSliderJoint( draw, jointSim, transformA, transformB, jointSim.wheel.localAxisA, jointSim.wheel.enableLimit,
							 jointSim.wheel.lowerTranslation, jointSim.wheel.upperTranslation, b2_colorBlue, b2_colorDimGray );
			break;
		default:
			drawSegment( draw, transformA.p, pA, b2_colorDarkSeaGreen );
			drawSegment( draw, pA, pB, b2_colorDarkSeaGreen );
			drawSegment( draw, transformB.p, pB, b2_colorDarkSeaGreen );
	}
	if ( draw->drawGraphColors )
	{
		int colorIndex = joint.colorIndex;
		if ( colorIndex != kNull )
			drawPoint( draw, lerp( pA, pB, 0.5f ), 5.0f, kGraphColors[colorIndex] );
	}
}

// One manifold, as both drawing paths emit it. The two paths differ in three constants (world.c:1029-1137 vs
// :1336-1430): the speculative colour, which impulse the "contact impulses" option shows and the number formats.
struct ContactDrawStyle
{
	b2HexColor speculativeColor;
	bool totalImpulse;			// true: totalNormalImpulse "%.2f", false: normalImpulse "%.1f"
	bool frictionScaledBy1000; // bounded path prints 1000 x tangentImpulse "%.1f", the full path tangentImpulse "%.2f"
};
void drawManifold( b2DebugDraw* draw, const Manifold& m, int colorIndex, const ContactDrawStyle& style )
{
	const float k_impulseScale = 1.0f;
	const float k_axisScale = 0.3f;
	V2 normal = m.normal;
	char buffer[32];
	for ( int j = 0; j < m.pointCount; ++j )
	{
		const ManifoldPoint& point = m.points[j];
		if ( draw->drawGraphColors )
			drawPoint( draw, point.point, colorIndex == kOverflow ? 7.5f : 5.0f, kGraphColors[colorIndex] );
		else if ( point.separation > kLinearSlop )
			drawPoint( draw, point.point, 5.0f, style.speculativeColor );
		else if ( point.persisted == false )
			drawPoint( draw, point.point, 10.0f, b2_colorGreen );
		else if ( point.persisted == true )
			drawPoint( draw, point.point, 5.0f, b2_colorBlue );

		if ( draw->drawContactNormals )
		{
			drawSegment( draw, point.point, mulAdd( point.point, k_axisScale, normal ), b2_colorDimGray );
		}
		else if ( draw->drawContactImpulses )
		{
			float impulse = style.totalImpulse ? point.totalNormalImpulse : point.normalImpulse;
			drawSegment( draw, point.point, mulAdd( point.point, k_impulseScale * impulse, normal ), b2_colorMagenta );
			snprintf( buffer, sizeof( buffer ), style.totalImpulse ? "%.2f" : "%.1f", 1000.0f * impulse );
			draw->DrawStringFcn( pub( point.point ), buffer, b2_colorWhite, draw->context );
		}
		if ( draw->drawContactFeatures )
		{
			snprintf( buffer, sizeof( buffer ), "%d", point.id );
			draw->DrawStringFcn( pub( point.point ), buffer, b2_colorOrange, draw->context );
		}
		if ( draw->drawFrictionImpulses )
		{
			V2 tangent = rightPerp( normal );
			drawSegment( draw, point.point, mulAdd( point.point, k_impulseScale * point.tangentImpulse, tangent ), b2_colorYellow );
			if ( style.frictionScaledBy1000 )
				snprintf( buffer, sizeof( buffer ), "%.1f", 1000.0f * point.tangentImpulse );
			else
				snprintf( buffer, sizeof( buffer ), "%.2f", point.tangentImpulse );
			draw->DrawStringFcn( pub( point.point ), buffer, b2_colorWhite, draw->context );
		}
	}
}

// world.c:966-1159: only what the drawing bounds touch, found through the broadphase trees
void drawWithBounds( World* w, b2DebugDraw* draw )
{
	const ContactDrawStyle style = { b2_colorGainsboro, false, true };
	std::vector<uint64_t> bodyBits( ( w->bodyIds.next + 63 ) / 64 + 1, 0 ), jointBits( ( w->jointIds.next + 63 ) / 64 + 1, 0 ),
		contactBits( ( w->contactIds.next + 63 ) / 64 + 1, 0 );
	Shape* shapes = ptr( w, w->shapes );
	Body* bodies = ptr( w, w->bodies );
	BodySim* sims = ptr( w, w->sims );
	Box bounds = { { draw->drawingBounds.lowerBound.x, draw->drawingBounds.lowerBound.y },
				   { draw->drawingBounds.upperBound.x, draw->drawingBounds.upperBound.y } };
	for ( int i = 0; i < 3; ++i )
	{
		treeQueryStats( w, w->trees[i], bounds, UINT64_MAX, [&]( int, uint64_t userData ) -> bool {
			const Shape& shape = shapes[(int)userData];
			bodyBits[shape.bodyId >> 6] |= 1ull << ( shape.bodyId & 63 );
			if ( draw->drawShapes )
			{
				const Body& body = bodies[shape.bodyId];
				const BodySim& sim = sims[shape.bodyId];
				drawShape( draw, shape, sim.transform, shapeDrawColor( shape, body, sim ) );
			}
			if ( draw->drawBounds )
				drawBox( draw, shape.fatAABB, b2_colorGold );
			return true;
		} );
	}
	const Joint* joints = ptr( w, w->joints );
	const Contact* contacts = ptr( w, w->contacts );
	const ContactSim* contactSims = ptr( w, w->contactSims );
	for ( size_t k = 0; k < bodyBits.size(); ++k )
	{
		uint64_t word = bodyBits[k];
		while ( word != 0 )
		{
			int bodyId = 64 * (int)k + __builtin_ctzll( word );
			const Body& body = bodies[bodyId];
			const BodySim& sim = sims[bodyId];
			if ( draw->drawBodyNames && body.name[0] != 0 )
			{
				Xf transform = { sim.center, sim.transform.q };
				draw->DrawStringFcn( pub( xfPoint( transform, V2{ 0.1f, 0.1f } ) ), body.name, b2_colorBlueViolet, draw->context );
			}
			if ( draw->drawMass && body.type == kDynamicBody )
			{
				Xf transform = { sim.center, sim.transform.q };
				draw->DrawTransformFcn( pubXf( transform ), draw->context );
				char buffer[32];
				snprintf( buffer, 32, "  %.2f", body.mass );
				draw->DrawStringFcn( pub( xfPoint( transform, V2{ 0.1f, 0.1f } ) ), buffer, b2_colorWhite, draw->context );
			}
			if ( draw->drawJoints )
			{
				int jointKey = body.headJointKey;
				while ( jointKey != kNull )
				{
					int jointId = jointKey >> 1;
					int edgeIndex = jointKey & 1;
					const Joint& joint = joints[jointId];
					if ( ( jointBits[jointId >> 6] & ( 1ull << ( jointId & 63 ) ) ) == 0 ) // avoid double draw
					{
						drawJoint( draw, w, joint );
						jointBits[jointId >> 6] |= 1ull << ( jointId & 63 );
					}
					jointKey = joint.edges[edgeIndex].nextKey;
				}
			}
			if ( draw->drawContacts && body.type == kDynamicBody && body.setIndex == kAwakeSet )
			{
				int contactKey = body.headContactKey;
				while ( contactKey != kNull )
				{
					int contactId = contactKey >> 1;
					int edgeIndex = contactKey & 1;
					const Contact& contact = contacts[contactId];
					contactKey = contact.edges[edgeIndex].nextKey;
					if ( contact.setIndex != kAwakeSet || contact.colorIndex == kNull )
						continue;
					if ( ( contactBits[contactId >> 6] & ( 1ull << ( contactId & 63 ) ) ) == 0 ) // avoid double draw
					{
						drawManifold( draw, unpackManifold( contactSims[contactId].manifold ), contact.colorIndex, style );
						contactBits[contactId >> 6] |= 1ull << ( contactId & 63 );
					}
				}
			}
			word = word & ( word - 1 );
		}
	}
}

} // namespace
} // extern "C++"

// types.c:86-151: every callback defaults to a no-op so that users may leave some unset
static void emptyDrawPolygon( const b2Vec2*, int, b2HexColor, void* ) {}
static void emptyDrawSolidPolygon( b2Transform, const b2Vec2*, int, float, b2HexColor, void* ) {}
static void emptyDrawCircle( b2Vec2, float, b2HexColor, void* ) {}
static void emptyDrawSolidCircle( b2Transform, float, b2HexColor, void* ) {}
static void emptyDrawSolidCapsule( b2Vec2, b2Vec2, float, b2HexColor, void* ) {}
static void emptyDrawSegment( b2Vec2, b2Vec2, b2HexColor, void* ) {}
static void emptyDrawTransform( b2Transform, void* ) {}
static void emptyDrawPoint( b2Vec2, float, b2HexColor, void* ) {}
static void emptyDrawString( b2Vec2, const char*, b2HexColor, void* ) {}

b2DebugDraw b2DefaultDebugDraw( void )
{
	b2DebugDraw draw;
	memset( &draw, 0, sizeof( draw ) );
	draw.DrawPolygonFcn = emptyDrawPolygon;
	draw.DrawSolidPolygonFcn = emptyDrawSolidPolygon;
	draw.DrawCircleFcn = emptyDrawCircle;
	draw.DrawSolidCircleFcn = emptyDrawSolidCircle;
	draw.DrawSolidCapsuleFcn = emptyDrawSolidCapsule;
	draw.DrawSegmentFcn = emptyDrawSegment;
	draw.DrawTransformFcn = emptyDrawTransform;
	draw.DrawPointFcn = emptyDrawPoint;
	draw.DrawStringFcn = emptyDrawString;
	return draw;
}

void b2World_Draw( b2WorldId worldId, b2DebugDraw* draw ) // world.c:1161-1489
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr || draw == nullptr )
		return;
	World* w = hostImage( *hw );
	if ( w->locked )
		return;
	if ( draw->useDrawingBounds )
	{
		drawWithBounds( w, draw );
		return;
	}
	Shape* shapes = ptr( w, w->shapes );
	Body* bodies = ptr( w, w->bodies );
	BodySim* sims = ptr( w, w->sims );
	if ( draw->drawShapes )
	{
		forEachBodyBySet( w, [&]( int bodyId ) {
			const Body& body = bodies[bodyId];
			const BodySim& sim = sims[bodyId];
			for ( int shapeId = body.headShapeId; shapeId != kNull; shapeId = shapes[shapeId].nextShapeId )
				drawShape( draw, shapes[shapeId], sim.transform, shapeDrawColor( shapes[shapeId], body, sim ) );
		} );
	}
	if ( draw->drawJoints )
	{
		const Joint* joints = ptr( w, w->joints );
		for ( int i = 0; i < w->joints.count; ++i )
		{
			if ( joints[i].setIndex == kNull )
				continue;
			drawJoint( draw, w, joints[i] );
		}
	}
	if ( draw->drawBounds )
	{
		forEachBodyBySet( w, [&]( int bodyId ) {
			const BodySim& sim = sims[bodyId];
			char buffer[32];
			snprintf( buffer, 32, "%d", bodyId );
			draw->DrawStringFcn( pub( sim.center ), buffer, b2_colorWhite, draw->context );
			for ( int shapeId = bodies[bodyId].headShapeId; shapeId != kNull; shapeId = shapes[shapeId].nextShapeId )
				drawBox( draw, shapes[shapeId].fatAABB, b2_colorGold );
		} );
	}
	if ( draw->drawBodyNames )
	{
		for ( int i = 0; i < w->bodies.count; ++i )
		{
			const Body& body = bodies[i];
			if ( body.setIndex == kNull || body.name[0] == 0 )
				continue;
			const BodySim& sim = sims[i];
			Xf transform = { sim.center, sim.transform.q };
			draw->DrawStringFcn( pub( xfPoint( transform, V2{ 0.05f, 0.05f } ) ), body.name, b2_colorBlueViolet, draw->context );
		}
	}
	if ( draw->drawMass )
	{
		forEachBodyBySet( w, [&]( int bodyId ) {
			const BodySim& sim = sims[bodyId];
			Xf transform = { sim.center, sim.transform.q };
			draw->DrawTransformFcn( pubXf( transform ), draw->context );
			char buffer[32];
			float mass = sim.invMass > 0.0f ? 1.0f / sim.invMass : 0.0f;
			snprintf( buffer, 32, "  %.2f", mass );
			draw->DrawStringFcn( pub( xfPoint( transform, V2{ 0.1f, 0.1f } ) ), buffer, b2_colorWhite, draw->context );
		} );
	}
	if ( draw->drawContacts )
	{
		const ContactDrawStyle style = { b2_colorLightGray, true, false };
		const ContactSim* contactSims = ptr( w, w->contactSims );
		for ( int colorIndex = 0; colorIndex < kColorCount; ++colorIndex )
		{
			const int32_t* list = ptr( w, w->colorContacts[colorIndex] );
			for ( int i = 0; i < w->colorContacts[colorIndex].count; ++i )
				drawManifold( draw, unpackManifold( contactSims[list[i]].manifold ), colorIndex, style );
		}
	}
	if ( draw->drawIslands )
	{
		const Island* islands = ptr( w, w->islands );
		for ( int i = 0; i < w->islands.count; ++i )
		{
			const Island& island = islands[i];
			if ( island.setIndex == kNull )
				continue;
			int shapeCount = 0;
			Box aabb = { { FLT_MAX, FLT_MAX }, { -FLT_MAX, -FLT_MAX } };
			for ( int bodyId = island.headBody; bodyId != kNull; bodyId = bodies[bodyId].islandNext )
			{
				for ( int shapeId = bodies[bodyId].headShapeId; shapeId != kNull; shapeId = shapes[shapeId].nextShapeId )
				{
					aabb = boxUnion( aabb, shapes[shapeId].fatAABB );
					shapeCount += 1;
				}
			}
			if ( shapeCount > 0 )
				drawBox( draw, aabb, b2_colorOrangeRed );
		}
	}
}
