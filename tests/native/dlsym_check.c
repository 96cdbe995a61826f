/* Resolves, by name and at run time, every b2* function forge2d's Dart FFI backend binds - the way dart:ffi's @Native
 * lookup does (dlopen the asset, dlsym the symbol) - and makes a first call through resolved pointers (a by-value
 * struct return, a world created and destroyed). Test infrastructure. usage: dlsym_check <library.so> <symbol list> */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

typedef struct
{
	float x, y;
} Vec2;
typedef struct
{
	float c, s;
} Rot;
typedef struct
{
	Vec2 v[8], n[8], centroid;
	float radius;
	int count;
} Polygon;
typedef struct
{
	unsigned short index1, generation;
} WorldId;

int main( int argc, char** argv )
{
	if ( argc < 3 )
		return 2;
	void* lib = dlopen( argv[1], RTLD_LAZY | RTLD_LOCAL );
	if ( lib == NULL )
	{
		fprintf( stderr, "dlopen failed: %s\n", dlerror() );
		return 3;
	}
	FILE* f = fopen( argv[2], "r" );
	if ( f == NULL )
		return 4;
	char line[256];
	int total = 0, missing = 0;
	while ( fgets( line, sizeof( line ), f ) )
	{
		line[strcspn( line, "\r\n" )] = 0;
		if ( line[0] == 0 || line[0] == '#' )
			continue;
		total += 1;
		if ( dlsym( lib, line ) == NULL )
		{
			printf( "missing %s\n", line );
			missing += 1;
		}
	}
	fclose( f );
	if ( missing == 0 )
	{
		/* b2MakeOffsetRoundedBox (collision.h): struct by value out, structs by value in */
		Polygon ( *makeBox )( float, float, Vec2, Rot, float ) = (Polygon( * )( float, float, Vec2, Rot, float ))dlsym( lib, "b2MakeOffsetRoundedBox" );
		Vec2 zero = { 0.0f, 0.0f };
		Rot identity = { 1.0f, 0.0f };
		Polygon box = makeBox( 0.5f, 0.25f, zero, identity, 0.0f );
		if ( box.count != 4 || box.v[2].x != 0.5f || box.v[2].y != 0.25f )
		{
			printf( "b2MakeOffsetRoundedBox returned count %d, v[2] = (%g, %g)\n", box.count, box.v[2].x, box.v[2].y );
			missing += 1;
		}
		/* b2DefaultWorldDef / b2CreateWorld / b2World_IsValid / b2DestroyWorld: host-side only, no device needed */
		typedef struct
		{
			char bytes[256];
		} WorldDefBlob; /* larger than b2WorldDef (types.h:62-134) */
		WorldDefBlob ( *defaultDef )( void ) = (WorldDefBlob( * )( void ))dlsym( lib, "b2DefaultWorldDef" );
		WorldId ( *create )( const void* ) = (WorldId( * )( const void* ))dlsym( lib, "b2CreateWorld" );
		_Bool ( *valid )( WorldId ) = (_Bool( * )( WorldId ))dlsym( lib, "b2World_IsValid" );
		void ( *destroy )( WorldId ) = (void ( * )( WorldId ))dlsym( lib, "b2DestroyWorld" );
		WorldDefBlob def = defaultDef();
		WorldId world = create( &def );
		if ( valid( world ) == 0 )
		{
			printf( "b2CreateWorld through dlsym did not give a valid world\n" );
			missing += 1;
		}
		destroy( world );
		if ( valid( world ) != 0 )
		{
			printf( "b2DestroyWorld through dlsym left the world valid\n" );
			missing += 1;
		}
	}
	printf( "resolved %d of %d\n", total - missing, total );
	dlclose( lib );
	return missing == 0 ? 0 : 1;
}
