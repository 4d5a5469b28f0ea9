// upg_host_count.cpp -- TEST-ONLY: the COUNTED host build of the product's device headers (terrain + colouring pipeline).
//
// tests/hostlib/upg_host_eval.cpp compiled with `float` / `double` replaced by the instrumented types of oracle/up_counted.h: every
// scalar operation of terrainTile (sources/modesConfiguration.cpp:281-286 -> coloringDefault & co, sources/coloringModes.cpp:905-966,
// incl. the nine shape probes of generateElevation / generateSlope and the FP64 HSLuv conversions) is executed natively and counted by
// class.  The plain host build is bit-identical to the reference on all 13 tile outputs (tests/test_host_coloring.py); this build is
// checked bit-identical to the plain one (tests/test_oracle_counts.py), so the counts are those of the reference's expression -- minus
// the two exact eliminations the headers carry (snow / dirt test their blend factor before the noise behind it; cellular pairs that
// share a seed share one 27-cell scan), i.e. a LOWER bound of the reference's work per texel.  SURVEY.md 8(d); tools/count_ops.py.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <math.h>
#include <stdint.h>

#include "../../oracle/up_counted.h"

#define float upc::cfloat
#define double upc::cdouble
#include "upg_host_eval.cpp"
#undef float
#undef double

extern "C"
{
	int upghc_op_count(void) { return upc::OP_COUNT; }
	// operation counts of the calling thread since the last reset
	void upghc_reset(void) { upc::counters() = upc::Counters{}; }
	void upghc_get(uint64_t *out)
	{
		for (int k = 0; k < upc::OP_COUNT; k++)
			out[k] = upc::counters().n[k];
	}
}
