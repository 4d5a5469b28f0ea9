// tools/noise_bounds.cpp -- numeric suprema of the Perlin and OpenSimplex2 generators as oracle/up_noise.h (and upg_noise.cuh) state them.
//
//   g++ -O2 -DUPO_WORST_CASE_GRADIENTS -o /tmp/noise_bounds tools/noise_bounds.cpp && /tmp/noise_bounds
//
// With UPO_WORST_CASE_GRADIENTS every lattice point contributes the largest value any of its possible gradients could give at the
// sample's offset, so the generator evaluates an upper bound of |noise| that is ATTAINED for a suitable hash assignment.  Both
// generators are periodic in the unit cell of their (already transformed) input, so a dense scan of [0,1)^3 finds the supremum up to
// the grid's Lipschitz error; the narrow-band sampling of the isosurface path (upg_terrain.cuh: elevation_bounds) uses these numbers
// with a further safety factor.
#include <cstdio>

#include "../oracle/up_noise.h"

int main()
{
	const int N = 320;
	double perlin = 0, simplex = 0;
	for (int k = 0; k < N; k++)
		for (int j = 0; j < N; j++)
			for (int i = 0; i < N; i++)
			{
				const float x = (i + 0.5f) / N, y = (j + 0.5f) / N, z = (k + 0.5f) / N;
				const double p = upo_single_perlin3(1, x, y, z), s = upo_single_simplex3(1, x, y, z);
				perlin = p > perlin ? p : perlin;
				simplex = s > simplex ? s : simplex;
			}
	std::printf("sup |Perlin3|       <= %.6f  (grid %d^3)\n", perlin, N);
	std::printf("sup |OpenSimplex2 3| <= %.6f  (grid %d^3)\n", simplex, N);
	return 0;
}
