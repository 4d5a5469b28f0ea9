// TEST-ONLY: the host side of the product's meshChunking (csrc/upg_chunking_host.h: the decisions the device path takes from a bounding
// box and eight serial sums) behind a C interface, for tests/test_oracle_chunks.py.
#include "upg_chunking_host.h"

extern "C"
{
	int upgh_chunk_longest_axis(const float *a, const float *b) { return upg::chunk_longest_axis(a, b); }
	void upgh_chunk_candidates(const float *a, const float *b, int axis, float *out7) { upg::chunk_candidates(a, b, axis, out7); }
	int upgh_chunk_wants_split(const float *sums8, float maxSurfaceArea, float *fraction) { return upg::chunk_wants_split(sums8, maxSurfaceArea, fraction) ? 1 : 0; }
	void upgh_chunk_clipping_box(const float *a, const float *b, int axis, float fraction, int second, float *box6) { upg::chunk_clipping_box(a, b, axis, fraction, second != 0, box6); }
}
