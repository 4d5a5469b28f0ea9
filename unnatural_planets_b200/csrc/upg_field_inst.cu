// upg_field_inst.cu -- instantiates K1 for ONE elevation mode and a range of shapes.
// Compiled once per (UPG_ELEV, UPG_PART) by the Makefile so that the 34 x 6 specialisations build in parallel, and a second time with
// -fmad=true -DUPG_CONTRACT for the elevation modes that have noise: the tolerance build (UPG_ARITH_CONTRACT, include/upg.h).
#include <cuda_runtime.h>

#include "upg_field.cuh"

#ifndef UPG_ELEV
#error "compile with -DUPG_ELEV=<0..5> -DUPG_PART=<0..3>"
#endif

#define UPG_CAT2(a, b, c, d) a##b##c##d
#define UPG_CAT(a, b, c, d) UPG_CAT2(a, b, c, d)

namespace upg
{
	namespace
	{
		constexpr int PARTS = 4;
		constexpr int LO = (SH_COUNT * UPG_PART) / PARTS;
		constexpr int HI = (SH_COUNT * (UPG_PART + 1)) / PARTS;

		// ELEV and PART are part of the type on purpose: nvcc names this anonymous namespace after the SOURCE FILE, which the 24 translation
		// units share, and emits the members as weak symbols -- without the two arguments the linker would merge the 24 different run()
		// bodies into whichever came first (it did, as soon as run() grew too big to be inlined: every planet sampled with elevation "none").
		template<int S, int ELEV = UPG_ELEV, int PART = UPG_PART, int ARITH = UPG_ARITH_BUILD>
		struct Dispatch
		{
			static bool run(int shape, const PlanetParams &P, const FieldLaunch &L, cudaStream_t stream)
			{
				if (shape == S)
				{
#if UPG_ELEV == 0 && !UPG_ARITH_BUILD
					// whole grid rows of the water field: the row-wise kernel (launch_field routes every planet's water launches here)
					if (L.field == FD_WATER && !L.positions && L.count % (uint64_t)P.grid.res == 0)
					{
						const uint64_t rows = L.count / (uint64_t)P.grid.res;
						shape_rows_kernel<S><<<(unsigned)((rows + SHAPE_ROWS - 1) / SHAPE_ROWS), FIELD_BLOCK, 0, stream>>>(P, L);
						return true;
					}
#endif
					// whole slices of land / navigation / raw elevation under a mode that starts with a Value fBm: the column kernel
					const uint64_t slice = (uint64_t)P.grid.res * (uint64_t)P.grid.res;
#ifndef UPG_NO_COLUMN
					if constexpr (ElevationStartsWithValueFbm<UPG_ELEV>::value)
					{
						if (!L.positions && L.count % slice == 0 && (L.field == FD_ALL || L.field == FD_LAND || L.field == FD_NAVIGATION || L.field == FD_ELEVATION_RAW))
						{
							const dim3 grid((unsigned)((P.grid.res + COLUMN_TILE_X - 1) / COLUMN_TILE_X), (unsigned)((P.grid.res + COLUMN_TILE_Y - 1) / COLUMN_TILE_Y), (unsigned)((L.count / slice + COLUMN_Z - 1) / COLUMN_Z));
							sdf_column_kernel<S, UPG_ELEV><<<grid, FIELD_BLOCK, 0, stream>>>(P, L);
							return true;
						}
					}
#endif
					const uint64_t blocks = (L.count + FIELD_BLOCK - 1) / FIELD_BLOCK;
					sdf_sample_kernel<S, UPG_ELEV><<<(unsigned)blocks, FIELD_BLOCK, 0, stream>>>(P, L);
					return true;
				}
				return Dispatch<S + 1, ELEV, PART, ARITH>::run(shape, P, L, stream);
			}
		};
		template<>
		struct Dispatch<HI, UPG_ELEV, UPG_PART, UPG_ARITH_BUILD>
		{
			static bool run(int, const PlanetParams &, const FieldLaunch &, cudaStream_t) { return false; }
		};
	}

#if UPG_ARITH_BUILD
#define UPG_LAUNCHER_PREFIX field_launchc_e
#else
#define UPG_LAUNCHER_PREFIX field_launch_e
#endif
	bool UPG_CAT(UPG_LAUNCHER_PREFIX, UPG_ELEV, _p, UPG_PART)(int shape, const PlanetParams &P, const FieldLaunch &L, void *stream)
	{
		if (shape < LO || shape >= HI)
			return false;
		return Dispatch<LO, UPG_ELEV, UPG_PART>::run(shape, P, L, (cudaStream_t)stream);
	}
}
