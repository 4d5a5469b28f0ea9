// upg_chunking_host.h -- the DECISIONS of cage's meshChunking (externals/cage/sources/libcore/mesh/meshAlgorithms.cpp:1498-1571) as plain
// host functions in single precision (cage's Real), shared by the device path (upg_meshfix.cu: the kernels deliver a bounding box and eight
// serial sums, these functions turn them into "leaf" or "split here, clip against these two boxes") and by the CPU check against the
// reference (tests/hostlib/upg_host_chunking.cpp, tests/test_oracle_chunks.py).  Compile without FMA contraction (-ffp-contract=off).
#pragma once

#include <cmath>

namespace upg
{
	// interpolate(Real, Real, Real) (cage-core/math.h:1173-1177)
	inline float chunk_interpolate(float a, float b, float f)
	{
		return a + (b - a) * f;
	}

	// Positions of meshChunkingImpl (:1536): Real constructed from double literals
	inline const float *chunk_split_fractions()
	{
		static const float fractions[7] = { (float)0.3, (float)0.4, (float)0.45, (float)0.5, (float)0.55, (float)0.6, (float)0.7 };
		return fractions;
	}

	// boxLongestAxis (:1498-1507) over abs(dominantAxis(box.size())) (math/vectors.cpp:305-312); sign(0) = 0, so a degenerate winner
	// falls through to the next test like `a[0] == 1` does
	inline int chunk_longest_axis(const float a[3], const float b[3])
	{
		const float sx = b[0] - a[0], sy = b[1] - a[1], sz = b[2] - a[2];
		const float ax = std::fabs(sx), ay = std::fabs(sy), az = std::fabs(sz);
		float d[3] = { 0, 0, 0 }; // dominantAxis(size)
		auto sign = [](float x) { return x < 0 ? -1.0f : x > 0 ? 1.0f : 0.0f; };
		if (ax > ay && ax > az)
			d[0] = sign(sx);
		else if (ay > az)
			d[1] = sign(sy);
		else
			d[2] = sign(sz);
		if (std::fabs(d[0]) == 1)
			return 0;
		if (std::fabs(d[1]) == 1)
			return 1;
		return 2;
	}

	// the seven candidate planes on `axis` (:1539-1541)
	inline void chunk_candidates(const float a[3], const float b[3], int axis, float out[7])
	{
		const float *f = chunk_split_fractions();
		for (int i = 0; i < 7; i++)
			out[i] = chunk_interpolate(a[axis], b[axis], f[i]);
	}

	// sums[0] = surface area, sums[1 + i] = area of the triangles whose centre lies below candidate i.  Returns false when the mesh is a
	// leaf (`myArea < config.maxSurfaceArea`, :1524); otherwise *fraction = the candidate whose lower part is closest to half the area
	// (first one wins a tie: strict less, :1561-1571).  NaN when no candidate compares (the reference would carry a NaN position on).
	inline bool chunk_wants_split(const float sums[8], float maxSurfaceArea, float *fraction)
	{
		const float area = sums[0];
		if (area < maxSurfaceArea)
			return false;
		const float *f = chunk_split_fractions();
		float best = NAN, bestDiff = INFINITY;
		for (int i = 0; i < 7; i++)
		{
			const float s = std::fabs((float)0.5 - sums[1 + i] / area);
			if (s < bestDiff)
			{
				bestDiff = s;
				best = f[i];
			}
		}
		*fraction = best;
		return true;
	}

	// clippingBox (:1509-1520): the bounding box grown by 20 % around its centre, cut at `fraction` of the GROWN extent on `axis`;
	// box6 = { a.x, a.y, a.z, b.x, b.y, b.z }
	inline void chunk_clipping_box(const float a[3], const float b[3], int axis, float fraction, bool second, float box6[6])
	{
		for (int k = 0; k < 3; k++)
		{
			const float c = (a[k] + b[k]) * (float)0.5; // Aabb::center() (geometry.h:192)
			const float hs = (b[k] - a[k]) * (float)0.6; // box.size() * 0.6
			// Aabb(c - hs, c + hs) orders its corners (geometry.h:166); hs >= 0 for a valid box, so they already are
			const float lo = c - hs, hi = c + hs;
			box6[k] = lo < hi ? lo : hi;
			box6[3 + k] = lo < hi ? hi : lo;
		}
		const float s = chunk_interpolate(box6[axis], box6[3 + axis], fraction);
		if (second)
			box6[axis] = s;
		else
			box6[3 + axis] = s;
	}
}
