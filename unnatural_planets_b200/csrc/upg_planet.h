// upg_planet.h -- the immutable per-planet parameter block handed to every kernel (by value, as a
// __grid_constant__ kernel argument: it lands in the constant bank, is re-entrant across the three
// concurrent field pipelines of sources/generator.cpp:432-439 and needs no global mutable state).
//
// Mode ids follow the ORDER of the reference's name tables (sources/modesConfiguration.cpp:78-113,
// 176-183, 207-211) so that `--shape random` (randomRange(0u, 34), :121) picks the same shape.
#pragma once

#include <cstdint>

#include "upg_noise.cuh"

namespace upg
{
	enum Shape : int32_t
	{
		SH_ASTEROID = 0,
		SH_BELT,
		SH_BOWL,
		SH_BOX,
		SH_BUNNY,
		SH_CAPSULE,
		SH_CUBE,
		SH_DISK,
		SH_DOUBLETORUS,
		SH_FIBERS,
		SH_GEAR,
		SH_H2O,
		SH_H3O,
		SH_H4O,
		SH_HEMISPHERES,
		SH_HEXAGON,
		SH_HEXAGONALPRISM,
		SH_INSIDECUBE,
		SH_KNOT,
		SH_MANDELBULB,
		SH_MOBIUSSTRIP,
		SH_MONKEYHEAD,
		SH_OCTAHEDRON,
		SH_PIPE,
		SH_SPHERE,
		SH_SQUARE,
		SH_TETRAHEDRON,
		SH_TORUS,
		SH_TORUSCROSS,
		SH_TRIANGULARPRISM,
		SH_TUBE,
		SH_TWISTEDHEXAGONALPRISM,
		SH_TWISTEDPLANE,
		SH_WORMHOLE,
		SH_COUNT
	};

	enum Elevation : int32_t
	{
		EL_NONE = 0,
		EL_SIMPLE,
		EL_LEGACY,
		EL_LAKES,
		EL_ISLANDS,
		EL_CRATERS,
		EL_COUNT
	};

	enum Coloring : int32_t
	{
		CL_DEFAULT = 0,
		CL_BARREN,
		CL_DEBUG,
		CL_COUNT
	};

	// terrainSdf* of sources/modesConfiguration.cpp:230-279
	enum Field : int32_t
	{
		FD_LAND = 0, // shape - elev / 10
		FD_WATER = 1, // shape
		FD_NAVIGATION = 2, // shape - max(elev, 0) / 10
		FD_ELEVATION_RAW = 3, // elev                (terrainSdfElevationRaw)
		FD_ELEVATION = 4, // shape * 10          (terrainSdfElevation)
		FD_COUNT,
		FD_ALL = FD_COUNT // internal to K1: land + water + navigation from one evaluation (upg_fields_samples_device)
	};

	// a triangle with everything closestPoint(Triangle) derives from it (geometry.cpp:644-673), prepared on the host
	struct TriangleConsts
	{
		V3 v[3];
		V3 normal; // Triangle::normal()
		float d; // Plane(point, normal).d
		V3 segOrigin[3]; // makeSegment(v[i], v[(i+1)%3])
		V3 segDir[3];
		float segMax[3];
	};

	struct VoronoiInst
	{
		int32_t seedHashes[3];
		uint32_t pointsPerCell;
		float cellSize;
		float frequency; // 1 / cellSize
	};

	struct ShapeConsts
	{
		V3 hexagonNormal; // normalize(Vec3(1, 1, -1))              sources/sdf.cpp:12
		V3 fibersOffset; // randomRange3(-100, 100)                 sources/sdf.cpp:133
		V3 h3oCenters[3]; // Quat({}, {}, Degs(k*120)) * (680,0,0)  sources/sdf.cpp:157-159
		float knotInvLen; // length(Vec3(1, 1, 0.3))                sources/sdf.cpp:86
		NoiseInst asteroid; // sources/sdf.cpp:335-345
		TriangleConsts prism; // sources/sdf.cpp:63
		TriangleConsts tetra[4]; // signedDistanceFunctions.cpp:44-50
	};

	struct ElevationConsts
	{
		NoiseInst n[4]; // in static-initialisation order of the mode
		VoronoiInst voronoi; // craters
	};

	struct GridConsts
	{
		int32_t res; // samples per axis
		float boxA; // box.a (same on the three axes): -boxSize/2
		float boxSpan; // box.b - box.a
		float resMinus5; // float(res) - 5
	};

	struct PlanetParams
	{
		int32_t shape;
		int32_t elevation;
		int32_t coloring;
		int32_t poles;
		int32_t flowers;
		GridConsts grid;
		ShapeConsts sh;
		ElevationConsts el;
	};
}
