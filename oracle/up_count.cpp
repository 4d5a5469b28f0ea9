/*
 * up_count.cpp -- TEST INFRASTRUCTURE (oracle): the COUNTED build of the stage-1 restatement (up_oracle.cpp + up_noise.h).
 *
 * The same source, compiled with `float` replaced by the instrumented type of up_counted.h: every scalar operation the reference's
 * sampling loop performs per sample (MarchingCubes::updateByPosition x terrainSdfLand / Water / Navigation) is
 * executed natively (results bit-identical to the plain build; tests/test_oracle_counts.py checks it) and counted by class.
 * bench.py's roofline numerator and BASELINE.md's work table come from here (tools/count_ops.py).
 *
 * Build: oracle/Makefile target `count` -> oracle/libup_count.so
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <math.h>
#include <stdint.h>
#include <thread>
#include <vector>

#include "up_counted.h"

#define float upc::cfloat
#include "up_oracle.cpp"
#undef float

extern "C"
{
	int upc_op_count(void) { return upc::OP_COUNT; }
	const char *upc_op_name(int i)
	{
		static const char *const names[upc::OP_COUNT] = { "fadd", "fmul", "fdiv", "fsqrt", "fcmp", "fneg", "i2f", "f2i", "special", "dadd", "dmul", "ddiv", "dsqrt", "dcmp", "dspecial", "f2d" };
		return i >= 0 && i < upc::OP_COUNT ? names[i] : "";
	}
	// samples field `which` at every `stride`-th node (per axis) of slices [z0, z1) of the res^3 grid on `threads` threads; adds the
	// operation counts of all evaluations to counts[OP_COUNT], the number of evaluated nodes to *nodes, and (optionally) stores the
	// values of a full slice range like upo_grid when stride == 1 and out != nullptr
	void upc_grid_counts(void *p, int which, float boxSize, int res, int z0, int z1, int stride, int threads, uint64_t *counts, uint64_t *nodes, float *out)
	{
		const Planet *P = (const Planet *)p;
		std::atomic<int> next{ 0 };
		std::atomic<uint64_t> total[upc::OP_COUNT];
		for (auto &t : total)
			t = 0;
		std::atomic<uint64_t> evaluated{ 0 };
		const int slices = (z1 - z0 + stride - 1) / stride;
		// grid coordinates once; what MarchingCubesCreateConfig::position (marchingCubes.cpp:247-257) costs PER SAMPLE in the reference --
		// f = (Vec3(x, y, z) - 2) / (Vec3(res) - 5); return (box.b - box.a) * f + box.a -- is added analytically: per axis 4 fadd, 1 fmul,
		// 1 fdiv and 2 integer -> float conversions
		std::vector<real> coord(res);
		for (int i = 0; i < res; i++)
			coord[i] = gridCoord(boxSize, res, (uint32_t)i);
		upc::Counters perCoord = {};
		perCoord.n[upc::OP_FADD] = 4;
		perCoord.n[upc::OP_FMUL] = 1;
		perCoord.n[upc::OP_FDIV] = 1;
		perCoord.n[upc::OP_I2F] = 2;
		auto work = [&]()
		{
			upc::Counters &c = upc::counters();
			const upc::Counters start = c;
			uint64_t mine = 0;
			for (;;)
			{
				const int s = next++;
				if (s >= slices)
					break;
				const int z = z0 + s * stride;
				for (int y = 0; y < res; y += stride)
					for (int x = 0; x < res; x += stride)
					{
						const real v = field(*P, which, V3(coord[x], coord[y], coord[z]));
						if (out && stride == 1)
							out[((size_t)(z - z0) * res + y) * res + x] = (float)v;
						mine++;
					}
			}
			for (int k = 0; k < upc::OP_COUNT; k++)
				total[k] += c.n[k] - start.n[k] + 3 * perCoord.n[k] * mine; // position(x, y, z): three coordinates per sample (marchingCubes.cpp:247-257)
			evaluated += mine;
		};
		std::vector<std::thread> pool;
		for (int t = 1; t < threads; t++)
			pool.emplace_back(work);
		work();
		for (auto &t : pool)
			t.join();
		for (int k = 0; k < upc::OP_COUNT; k++)
			counts[k] += total[k];
		*nodes += evaluated;
	}
}
