// upg_dmc.cu -- K2..K4: isosurface extraction (Dual Marching Cubes with Wenger's manifold fix and shared
// vertices) over the sampled field, entirely on the device, with the reference's vertex / quad / triangle ORDER.
//
// Replaces dualmc::DualMC<float>::build(iso 0, manifold, shared)  (externals/cage/externals/dualmc/dualmc/include/
// dualmc.tpp:12-32 getCellCode, :37-86 getDualPointCode, :91-188 calculateDualPoint, :193-217
// getSharedDualPointIndex, :384-461 buildSharedVerticesQuads) and the triangle tail of MarchingCubes::makeMesh
// (externals/cage/sources/libcore/mesh/marchingCubes.cpp:186-227: grid->world, quad split by the shorter diagonal,
// degenerate filter, area-weighted normals).
//
// The serial reference numbers a dual vertex when the z/y/x/axis/slot scan first touches it (an unordered_map).
// Here every quad slot has the ordinal  ord = (((z*(ry-2) + y)*(rx-2) + x)*3 + axis)*4 + slot  of that scan; a dual
// point (cell, pointCode) is created by the smallest ordinal among the slots that reference it, which every slot can
// compute locally from the point code.  Vertex id = rank of the creating slot = exclusive scan of "I am the creator"
// flags in scan order: identical indices to the hash-map version, no atomics, no hash table.
//
// Passes (HBM roofline: the field is read once, 4 B/sample; everything after works on the ~0.2 % surface cells):
//   cell_codes   dense   8-corner sign code per cell (1 B/cell) + per-block count of emitted quads
//   scan #1      CUB     block offsets -> ordered quad compaction
//   emit_quads   dense   over the 1 B codes: compact (cell, axis) list in scan order, first-quad index per grid point
//   slot_flags   sparse  per quad slot: manifold point code, creator ordinal -> creator flag
//   scan #2      CUB     vertex ids
//   build        sparse  creators compute the dual vertex (edge crossings), every slot writes its quad index
//   triangles    sparse  split / degenerate filter / face normals, scan #3, ordered index emission
//   normals      sparse  per vertex: gather its faces in reference order, sum, normalise
#include <cub/device/device_scan.cuh>

#include <cuda_runtime.h>

#include "../../include/upg.h"
#include "upg_internal.h"

namespace upg
{
#define UPG_TABLE_QUALIFIER __constant__
#define UPG_TABLE_NAME(n) c_##n
#include "upg_dmc_tables.inc"
#undef UPG_TABLE_QUALIFIER
#undef UPG_TABLE_NAME

	namespace
	{
		constexpr int DMC_BLOCK = 256;

		struct Dims
		{
			int32_t rx, ry, rz; // samples per axis
			int32_t z0; // global z index of local slice 0 (slab extraction; 0 for a whole volume)
			__host__ __device__ int32_t cx() const { return rx - 1; } // cells per axis
			__host__ __device__ int32_t cy() const { return ry - 1; }
			__host__ __device__ int32_t cz() const { return rz - 1; }
			__host__ __device__ uint64_t cells() const { return (uint64_t)cx() * cy() * cz(); }
			// the 1-byte cell codes are stored with rows padded to a multiple of 4 so that four codes move as one aligned word
			__host__ __device__ int32_t codePitch() const { return (cx() + 3) & ~3; }
			__host__ __device__ uint64_t rows() const { return (uint64_t)cy() * cz(); }
		};

		__device__ __forceinline__ uint64_t sample_index(const Dims &d, int32_t x, int32_t y, int32_t z)
		{
			return ((uint64_t)z * d.ry + y) * d.rx + x; // gA(), dualmc.h:229-231
		}
		__device__ __forceinline__ uint32_t cell_index(const Dims &d, int32_t x, int32_t y, int32_t z)
		{
			return (uint32_t)(((uint64_t)z * d.cy() + y) * d.cx() + x);
		}
		__device__ __forceinline__ uint64_t code_index(const Dims &d, int32_t x, int32_t y, int32_t z)
		{
			return ((uint64_t)z * d.cy() + y) * (uint32_t)d.codePitch() + x;
		}
		__device__ __forceinline__ void cell_coords(const Dims &d, uint32_t L, int32_t &x, int32_t &y, int32_t &z)
		{
			const uint32_t row = L / (uint32_t)d.cx();
			x = (int32_t)(L - row * (uint32_t)d.cx());
			z = (int32_t)(row / (uint32_t)d.cy());
			y = (int32_t)(row - (uint32_t)z * (uint32_t)d.cy());
		}

		// which of the three quads anchored at grid point (x, y, z) the reference loop emits (dualmc.tpp:401-460):
		// loop ranges x,y,z < dim-2, guards (z>0 && y>0) / (z>0 && x>0) / (x>0 && y>0), edge crossed
		__device__ __forceinline__ uint32_t quad_mask(const Dims &d, int32_t x, int32_t y, int32_t z, uint32_t code)
		{
			if (x >= d.rx - 2 || y >= d.ry - 2 || z >= d.rz - 2)
				return 0;
			const uint32_t b0 = code & 1u;
			uint32_t m = 0;
			if (z > 0 && y > 0 && (b0 ^ ((code >> 1) & 1u)))
				m |= 1u; // x edge
			if (z > 0 && x > 0 && (b0 ^ ((code >> 2) & 1u)))
				m |= 2u; // y edge
			if (x > 0 && y > 0 && (b0 ^ ((code >> 4) & 1u)))
				m |= 4u; // z edge
			return m;
		}

		// ---- pass 1: cell codes ---------------------------------------------------------------------------
		// Both dense passes are bandwidth work (4 B read per cell here, 1 B per cell in pass 2) and were instruction bound as
		// one-thread-per-cell kernels (profiles/r01_b: 80-90 % issue active at < 10 % of HBM).  Now one block walks one ROW of
		// cells (fixed y, z) and a thread handles four consecutive cells: five samples from each of the four sample rows that
		// bound the cell row (one 16-byte load + one scalar each), five sign tests per row instead of eight per cell, four codes
		// stored as one word.  Rows are visited in scan order, so per-row quad counts scan into ordered offsets as before.
		constexpr int DMC_ROW_MAX_THREADS = 256;

		// sign bits of samples x0 .. x0+4 of one sample row; `near` = how many of the first four (those that are cell origins,
		// x < cx) lie within `tol` of zero
		__device__ __forceinline__ uint32_t sign_bits5(const float *__restrict__ row, int32_t x0, int32_t rx, bool vec, float tol, uint32_t &near)
		{
			float v[5];
			if (vec)
			{
				const float4 q = __ldg(reinterpret_cast<const float4 *>(row + x0));
				v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
				v[4] = x0 + 4 < rx ? __ldg(row + x0 + 4) : -1.0f;
			}
			else
			{
#pragma unroll
				for (int k = 0; k < 5; k++)
					v[k] = x0 + k < rx ? __ldg(row + x0 + k) : -1.0f;
			}
			uint32_t bits = 0;
			near = 0;
#pragma unroll
			for (int k = 0; k < 5; k++)
			{
				bits |= (v[k] >= 0.0f ? 1u : 0u) << k; // getCellCode: bit set when the sample is >= iso (iso = 0)
				if (k < 4 && x0 + k < rx - 1)
					near += fabsf(v[k]) < tol ? 1u : 0u;
			}
			return bits;
		}

		// block-wide sum of two counters; valid in thread 0
		__device__ __forceinline__ void block_sum2(uint32_t &a, uint32_t &b)
		{
			__shared__ uint32_t sa[DMC_ROW_MAX_THREADS / 32], sb[DMC_ROW_MAX_THREADS / 32];
			for (int o = 16; o > 0; o >>= 1)
			{
				a += __shfl_down_sync(0xffffffffu, a, o);
				b += __shfl_down_sync(0xffffffffu, b, o);
			}
			if ((threadIdx.x & 31) == 0)
			{
				sa[threadIdx.x >> 5] = a;
				sb[threadIdx.x >> 5] = b;
			}
			__syncthreads();
			if (threadIdx.x == 0)
			{
				a = b = 0;
				for (uint32_t i = 0; i < (blockDim.x + 31) / 32; i++)
				{
					a += sa[i];
					b += sb[i];
				}
			}
		}

		__global__ void __launch_bounds__(DMC_ROW_MAX_THREADS) dmc_cell_codes(const float *__restrict__ dens, Dims d, uint8_t *__restrict__ codes, uint32_t *__restrict__ rowQuads, unsigned long long *__restrict__ nearZero, float nearTol)
		{
			const uint32_t row = blockIdx.x;
			const int32_t z = (int32_t)(row / (uint32_t)d.cy()), y = (int32_t)(row - (uint32_t)z * (uint32_t)d.cy());
			const int32_t cx = d.cx();
			const float *r00 = dens + sample_index(d, 0, y, z);
			const float *r01 = r00 + d.rx; // y + 1
			const float *r10 = r00 + (uint64_t)d.rx * d.ry; // z + 1
			const float *r11 = r10 + d.rx;
			// 16-byte loads need rows that start on 16-byte boundaries
			const bool vec = (d.rx & 3) == 0 && (reinterpret_cast<uintptr_t>(dens) & 15) == 0;
			uint32_t *codeWords = reinterpret_cast<uint32_t *>(codes + (uint64_t)row * (uint32_t)d.codePitch());
			uint32_t nq = 0, nz = 0;
			for (int32_t x0 = (int32_t)threadIdx.x * 4; x0 < cx; x0 += (int32_t)blockDim.x * 4)
			{
				uint32_t near, unused;
				const uint32_t s00 = sign_bits5(r00, x0, d.rx, vec, nearTol, near);
				const uint32_t s01 = sign_bits5(r01, x0, d.rx, vec, nearTol, unused);
				const uint32_t s10 = sign_bits5(r10, x0, d.rx, vec, nearTol, unused);
				const uint32_t s11 = sign_bits5(r11, x0, d.rx, vec, nearTol, unused);
				nz += near; // samples within tolerance of zero: where sign / topology parity with a CPU run is not promised
				uint32_t word = 0;
#pragma unroll
				for (int k = 0; k < 4; k++)
				{
					if (x0 + k >= cx)
						break;
					// corner id = x + 2y + 4z
					const uint32_t code = ((s00 >> k) & 3u) | (((s01 >> k) & 3u) << 2) | (((s10 >> k) & 3u) << 4) | (((s11 >> k) & 3u) << 6);
					word |= code << (8 * k);
					nq += __popc(quad_mask(d, x0 + k, y, z, code));
				}
				codeWords[x0 >> 2] = word;
			}
			block_sum2(nq, nz);
			if (threadIdx.x == 0)
			{
				rowQuads[row] = nq;
				if (nz)
					atomicAdd(nearZero, (unsigned long long)nz);
			}
		}

		// The same pass from the samplers' bit fields (SignBits, upg_internal.h): 1 bit instead of 4 bytes read per node.  One WARP walks
		// one row of cells, a lane 32 consecutive cells: bits x0 .. x0 + 32 of the four sample rows that bound the cell row (a word and
		// the first bit of the next; rows carry an all-zero padding word), 32 codes = 8 code words.  Nearly every lane sees four uniform
		// bit strings (far from the surface) and stores eight constant words; per-row quad counts come from a warp reduction.
		constexpr int DMC_BITS_WARPS = 8;
		__global__ void __launch_bounds__(DMC_BITS_WARPS * 32) dmc_cell_codes_bits(const uint32_t *__restrict__ signs, const uint32_t *__restrict__ nearBits, uint32_t pitch, Dims d, uint8_t *__restrict__ codes, uint32_t *__restrict__ rowQuads, unsigned long long *__restrict__ nearZero)
		{
			const uint32_t row = blockIdx.x * DMC_BITS_WARPS + (threadIdx.x >> 5);
			if (row >= (uint32_t)d.rows())
				return;
			const uint32_t lane = threadIdx.x & 31u;
			const int32_t z = (int32_t)(row / (uint32_t)d.cy()), y = (int32_t)(row - (uint32_t)z * (uint32_t)d.cy());
			const int32_t cx = d.cx();
			const uint64_t at = ((uint64_t)z * d.ry + y) * pitch;
			const uint32_t *b00 = signs + at;
			const uint32_t *b01 = b00 + pitch; // y + 1
			const uint32_t *b10 = b00 + (uint64_t)d.ry * pitch; // z + 1
			const uint32_t *b11 = b10 + pitch;
			uint32_t *codeWords = reinterpret_cast<uint32_t *>(codes + (uint64_t)row * (uint32_t)d.codePitch());
			uint32_t nq = 0, nz = 0;
			for (int32_t x0 = (int32_t)lane * 32; x0 < cx; x0 += 32 * 32)
			{
				const uint32_t w = (uint32_t)x0 >> 5;
				const uint32_t cellsHere = (uint32_t)min(32, cx - x0);
				// samples within tolerance of zero among the cell origins of this row (the statistic of the float pass)
				nz += __popc(__ldg(nearBits + at + w) & (cellsHere == 32u ? 0xffffffffu : (1u << cellsHere) - 1u));
				const uint32_t l00 = __ldg(b00 + w), l01 = __ldg(b01 + w), l10 = __ldg(b10 + w), l11 = __ldg(b11 + w);
				const uint32_t h00 = __ldg(b00 + w + 1), h01 = __ldg(b01 + w + 1), h10 = __ldg(b10 + w + 1), h11 = __ldg(b11 + w + 1);
				const uint32_t words = (cellsHere + 3u) >> 2;
				const uint32_t all = l00 & l01 & l10 & l11, any = l00 | l01 | l10 | l11;
				const uint32_t next = (h00 & h01 & h10 & h11 & 1u) | ((h00 | h01 | h10 | h11) & 1u) << 1; // 0: all clear, 3: all set
				if (cellsHere == 32u && ((any == 0u && next == 0u) || (all == 0xffffffffu && next == 3u)))
				{
					const uint32_t fillWord = any ? 0xffffffffu : 0u; // every corner of every cell on the same side: no quad here
#pragma unroll
					for (uint32_t k = 0; k < 8; k++)
						codeWords[(x0 >> 2) + k] = fillWord;
					continue;
				}
				const uint64_t s00 = l00 | (uint64_t)(h00 & 1u) << 32, s01 = l01 | (uint64_t)(h01 & 1u) << 32;
				const uint64_t s10 = l10 | (uint64_t)(h10 & 1u) << 32, s11 = l11 | (uint64_t)(h11 & 1u) << 32;
				for (uint32_t k4 = 0; k4 < words; k4++)
				{
					uint32_t word = 0;
#pragma unroll
					for (uint32_t k = 0; k < 4; k++)
					{
						const uint32_t c = k4 * 4 + k;
						if (c >= cellsHere)
							break;
						// corner id = x + 2y + 4z
						const uint32_t code = (uint32_t)((s00 >> c) & 3u) | ((uint32_t)((s01 >> c) & 3u) << 2) | ((uint32_t)((s10 >> c) & 3u) << 4) | ((uint32_t)((s11 >> c) & 3u) << 6);
						word |= code << (8 * k);
						nq += __popc(quad_mask(d, x0 + (int32_t)c, y, z, code));
					}
					codeWords[(x0 >> 2) + k4] = word;
				}
			}
			for (int o = 16; o > 0; o >>= 1)
			{
				nq += __shfl_down_sync(0xffffffffu, nq, o);
				nz += __shfl_down_sync(0xffffffffu, nz, o);
			}
			if (lane == 0)
			{
				rowQuads[row] = nq;
				if (nz)
					atomicAdd(nearZero, (unsigned long long)nz);
			}
		}

		// ---- pass 2: ordered quad compaction ----------------------------------------------------------------
		// quadCell[q] = linear cell index of the anchoring grid point, quadInfo[q] = axis | firstOrder << 2
		// One WARP per row of cells (rows without quads -- three in four even on the rows that cross a planet, nearly all elsewhere -- cost
		// one 4-byte read per warp), a lane four consecutive cells per step, the ordered offsets from a warp prefix: no block barrier.
		constexpr int DMC_EMIT_WARPS = 8;
		__global__ void __launch_bounds__(DMC_EMIT_WARPS * 32) dmc_emit_quads(Dims d, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ rowOffsets, const uint32_t *__restrict__ rowQuads, uint32_t *__restrict__ quadCell, uint8_t *__restrict__ quadInfo, uint32_t *__restrict__ quadBase)
		{
			const uint32_t row = blockIdx.x * DMC_EMIT_WARPS + (threadIdx.x >> 5);
			if (row >= (uint32_t)d.rows() || rowQuads[row] == 0)
				return;
			const uint32_t lane = threadIdx.x & 31u;
			const int32_t z = (int32_t)(row / (uint32_t)d.cy()), y = (int32_t)(row - (uint32_t)z * (uint32_t)d.cy());
			const int32_t cx = d.cx();
			const uint32_t *codeWords = reinterpret_cast<const uint32_t *>(codes + (uint64_t)row * (uint32_t)d.codePitch());
			const uint32_t rowL = row * (uint32_t)cx; // linear index of the row's first cell
			uint32_t carry = rowOffsets[row];
			for (int32_t c0 = 0; c0 < cx; c0 += 32 * 4)
			{
				const int32_t x0 = c0 + (int32_t)lane * 4;
				uint32_t word = 0, masks = 0, nq = 0;
				if (x0 < cx)
				{
					word = __ldg(codeWords + (x0 >> 2));
					// four cells all inside or all outside emit nothing (their shared edges do not cross either)
					if (word != 0u && word != 0xffffffffu)
					{
#pragma unroll
						for (int k = 0; k < 4; k++)
						{
							if (x0 + k >= cx)
								break;
							const uint32_t m = quad_mask(d, x0 + k, y, z, (word >> (8 * k)) & 0xffu);
							masks |= m << (4 * k);
							nq += __popc(m);
						}
					}
				}
				// exclusive prefix inside the warp, in x order
				uint32_t incl = nq;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
					if (lane >= (uint32_t)o)
						incl += t;
				}
				uint32_t q = carry + incl - nq;
				carry += __shfl_sync(0xffffffffu, incl, 31);
				if (!masks)
					continue;
#pragma unroll
				for (int k = 0; k < 4; k++)
				{
					const uint32_t mask = (masks >> (4 * k)) & 7u;
					if (!mask)
						continue;
					const uint32_t L = rowL + (uint32_t)(x0 + k);
					const uint32_t b0 = (word >> (8 * k)) & 1u;
					quadBase[L] = q;
					// x edge: entering (first sample < iso, second >= iso) keeps (i0,i1,i2,i3); y and z edges keep it when
					// exiting (dualmc.tpp:415-419, 434-438, 453-457).  entering <=> b0 == 0
					if (mask & 1u)
					{
						quadCell[q] = L;
						quadInfo[q] = (uint8_t)(0u | ((b0 == 0u ? 1u : 0u) << 2));
						q++;
					}
					if (mask & 2u)
					{
						quadCell[q] = L;
						quadInfo[q] = (uint8_t)(1u | ((b0 == 1u ? 1u : 0u) << 2));
						q++;
					}
					if (mask & 4u)
					{
						quadCell[q] = L;
						quadInfo[q] = (uint8_t)(2u | ((b0 == 1u ? 1u : 0u) << 2));
						q++;
					}
				}
			}
		}

		// ---- dual point machinery ---------------------------------------------------------------------------
		// the four (cell offset, cell edge) a quad asks for, in call order i0..i3 (dualmc.tpp:410-413, 429-432, 448-451)
		__constant__ int8_t c_slotCell[3][4][3] = {
			{ { 0, 0, 0 }, { 0, 0, -1 }, { 0, -1, -1 }, { 0, -1, 0 } }, // x edge: E0, E2, E6, E4
			{ { 0, 0, 0 }, { 0, 0, -1 }, { -1, 0, -1 }, { -1, 0, 0 } }, // y edge: E8, E11, E10, E9
			{ { 0, 0, 0 }, { -1, 0, 0 }, { -1, -1, 0 }, { 0, -1, 0 } }, // z edge: E3, E1, E5, E7
		};
		__constant__ uint8_t c_slotEdge[3][4] = { { 0, 2, 6, 4 }, { 8, 11, 10, 9 }, { 3, 1, 5, 7 } };
		// inverse view: cell edge e is slot c_edgeSlot[e] of the quad on axis c_edgeAxis[e] anchored at cell + c_edgeBase[e]
		__constant__ uint8_t c_edgeAxis[12] = { 0, 2, 0, 2, 0, 2, 0, 2, 1, 1, 1, 1 };
		__constant__ uint8_t c_edgeSlot[12] = { 0, 1, 1, 0, 3, 2, 2, 3, 0, 3, 2, 1 };
		__constant__ int8_t c_edgeBase[12][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 0, 0, 1 }, { 0, 0, 0 }, { 0, 1, 0 }, { 1, 1, 0 }, { 0, 1, 1 }, { 0, 1, 0 }, { 0, 0, 0 }, { 1, 0, 0 }, { 1, 0, 1 }, { 0, 0, 1 } };

		// getDualPointCode (dualmc.tpp:37-86) with generateManifold = true
		__device__ __forceinline__ uint32_t dual_point_code(const Dims &d, const uint8_t *__restrict__ codes, int32_t cx, int32_t cy, int32_t cz, uint32_t edgeBit)
		{
			uint32_t code = codes[code_index(d, cx, cy, cz)];
			const uint32_t dir = c_problematic[code];
			if (dir != 255u)
			{
				int32_t n[3] = { cx, cy, cz };
				const uint32_t comp = dir >> 1;
				n[comp] += (dir & 1u) ? 1 : -1;
				const int32_t lim = (comp == 0 ? d.rx : comp == 1 ? d.ry : d.rz) - 1;
				if (n[comp] >= 0 && n[comp] < lim)
				{
					const uint32_t ncode = codes[code_index(d, n[0], n[1], n[2])];
					if (c_problematic[ncode] != 255u)
						code ^= 0xffu;
				}
			}
#pragma unroll
			for (int i = 0; i < 4; i++)
				if (c_dualPoints[code][i] & edgeBit)
					return c_dualPoints[code][i];
			return 0;
		}

		// scan ordinal of the slot through which cell edge e of cell (cx,cy,cz) is referenced, or ~0 if that quad is
		// never emitted (outside the loop ranges / guards)
		__device__ __forceinline__ uint64_t slot_ordinal(const Dims &d, int32_t cx, int32_t cy, int32_t cz, uint32_t e)
		{
			const int32_t bx = cx + c_edgeBase[e][0], by = cy + c_edgeBase[e][1], bz = cz + c_edgeBase[e][2];
			const uint32_t axis = c_edgeAxis[e];
			if (bx >= d.rx - 2 || by >= d.ry - 2 || bz >= d.rz - 2)
				return ~0ull;
			const bool ok = axis == 0 ? (bz > 0 && by > 0) : axis == 1 ? (bz > 0 && bx > 0) : (bx > 0 && by > 0);
			if (!ok)
				return ~0ull;
			const uint64_t p = ((uint64_t)bz * (d.ry - 2) + by) * (d.rx - 2) + bx;
			return (p * 3 + axis) * 4 + c_edgeSlot[e];
		}

		struct SlotRef
		{
			int32_t cx, cy, cz; // cell of the dual point
			uint32_t pointCode;
			uint64_t myOrd; // ordinal of this slot
			uint64_t firstOrd; // ordinal of the creating slot
			uint32_t firstEdge; // cell edge through which the creator references the point
		};

		__device__ __forceinline__ SlotRef resolve_slot(const Dims &d, const uint8_t *__restrict__ codes, uint32_t cellL, uint32_t axis, uint32_t slot)
		{
			SlotRef r;
			int32_t x, y, z;
			cell_coords(d, cellL, x, y, z);
			r.cx = x + c_slotCell[axis][slot][0];
			r.cy = y + c_slotCell[axis][slot][1];
			r.cz = z + c_slotCell[axis][slot][2];
			const uint32_t e = c_slotEdge[axis][slot];
			r.pointCode = dual_point_code(d, codes, r.cx, r.cy, r.cz, 1u << e);
			r.myOrd = ((((uint64_t)z * (d.ry - 2) + y) * (d.rx - 2) + x) * 3 + axis) * 4 + slot;
			r.firstOrd = ~0ull;
			r.firstEdge = e;
			uint32_t pc = r.pointCode;
			while (pc)
			{
				const uint32_t k = __ffs(pc) - 1;
				pc &= pc - 1;
				const uint64_t o = slot_ordinal(d, r.cx, r.cy, r.cz, k);
				if (o < r.firstOrd)
				{
					r.firstOrd = o;
					r.firstEdge = k;
				}
			}
			return r;
		}

		// ---- pass 3: creator flags ----------------------------------------------------------------------------
		__global__ void __launch_bounds__(DMC_BLOCK) dmc_slot_flags(Dims d, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ quadCell, const uint8_t *__restrict__ quadInfo, uint32_t nSlots, uint32_t *__restrict__ flags)
		{
			const uint32_t i = blockIdx.x * DMC_BLOCK + threadIdx.x;
			if (i >= nSlots)
				return;
			const uint32_t q = i >> 2, slot = i & 3u;
			const SlotRef r = resolve_slot(d, codes, quadCell[q], quadInfo[q] & 3u, slot);
			flags[i] = r.firstOrd == r.myOrd ? 1u : 0u;
		}

		// calculateDualPoint (dualmc.tpp:91-188): mean of the edge crossings of the patch, accumulated in edge order
		__device__ __forceinline__ V3 dual_point(const Dims &d, const float *__restrict__ dens, int32_t cx, int32_t cy, int32_t cz, uint32_t pc)
		{
			const uint64_t s = sample_index(d, cx, cy, cz);
			const uint64_t sy = (uint64_t)d.rx, sz = (uint64_t)d.rx * d.ry;
			const float iso = 0.0f;
#define UPG_S(dx, dy, dz) __ldg(dens + s + (dx) + (dy) * sy + (dz) * sz)
#define UPG_CROSS(a, b) ((iso - (a)) / ((b) - (a)))
			float px = 0, py = 0, pz = 0;
			int points = 0;
			if (pc & (1u << 0))
			{
				px += UPG_CROSS(UPG_S(0, 0, 0), UPG_S(1, 0, 0));
				points++;
			}
			if (pc & (1u << 1))
			{
				px += 1.0f;
				pz += UPG_CROSS(UPG_S(1, 0, 0), UPG_S(1, 0, 1));
				points++;
			}
			if (pc & (1u << 2))
			{
				px += UPG_CROSS(UPG_S(0, 0, 1), UPG_S(1, 0, 1));
				pz += 1.0f;
				points++;
			}
			if (pc & (1u << 3))
			{
				pz += UPG_CROSS(UPG_S(0, 0, 0), UPG_S(0, 0, 1));
				points++;
			}
			if (pc & (1u << 4))
			{
				px += UPG_CROSS(UPG_S(0, 1, 0), UPG_S(1, 1, 0));
				py += 1.0f;
				points++;
			}
			if (pc & (1u << 5))
			{
				px += 1.0f;
				pz += UPG_CROSS(UPG_S(1, 1, 0), UPG_S(1, 1, 1));
				py += 1.0f;
				points++;
			}
			if (pc & (1u << 6))
			{
				px += UPG_CROSS(UPG_S(0, 1, 1), UPG_S(1, 1, 1));
				pz += 1.0f;
				py += 1.0f;
				points++;
			}
			if (pc & (1u << 7))
			{
				pz += UPG_CROSS(UPG_S(0, 1, 0), UPG_S(0, 1, 1));
				py += 1.0f;
				points++;
			}
			if (pc & (1u << 8))
			{
				py += UPG_CROSS(UPG_S(0, 0, 0), UPG_S(0, 1, 0));
				points++;
			}
			if (pc & (1u << 9))
			{
				px += 1.0f;
				py += UPG_CROSS(UPG_S(1, 0, 0), UPG_S(1, 1, 0));
				points++;
			}
			if (pc & (1u << 10))
			{
				px += 1.0f;
				py += UPG_CROSS(UPG_S(1, 0, 1), UPG_S(1, 1, 1));
				pz += 1.0f;
				points++;
			}
			if (pc & (1u << 11))
			{
				pz += 1.0f;
				py += UPG_CROSS(UPG_S(0, 0, 1), UPG_S(0, 1, 1));
				points++;
			}
#undef UPG_S
#undef UPG_CROSS
			const float inv = 1.0f / (float)points;
			px *= inv;
			py *= inv;
			pz *= inv;
			return v3((float)cx + px, (float)cy + py, (float)(cz + d.z0) + pz); // the GLOBAL cell origin enters the sum (slabs)
		}

		struct WorldMap
		{
			V3 mult; // cfg.box.size() / (Vec3(resolution) - 5)   marchingCubes.cpp:194
			V3 add; // cfg.position(0, 0, 0)                      marchingCubes.cpp:193
		};

		// index of the quad (axis `axis`) anchored at grid point `cellL` in the compact list
		__device__ __forceinline__ uint32_t quad_lookup(const Dims &d, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ quadBase, int32_t bx, int32_t by, int32_t bz, uint32_t axis)
		{
			const uint32_t L = cell_index(d, bx, by, bz);
			const uint32_t mask = quad_mask(d, bx, by, bz, codes[code_index(d, bx, by, bz)]);
			return quadBase[L] + __popc(mask & ((1u << axis) - 1u));
		}

		// ---- pass 4: vertices + quad indices ------------------------------------------------------------------
		__global__ void __launch_bounds__(DMC_BLOCK) dmc_build(Dims d, const float *__restrict__ dens, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ quadCell, const uint8_t *__restrict__ quadInfo, const uint32_t *__restrict__ quadBase, const uint32_t *__restrict__ vids, uint32_t nSlots, WorldMap wm, float *__restrict__ rawVerts, float *__restrict__ worldVerts, uint32_t *__restrict__ vertCell, uint16_t *__restrict__ vertCode, uint32_t *__restrict__ quadFinal)
		{
			const uint32_t i = blockIdx.x * DMC_BLOCK + threadIdx.x;
			if (i >= nSlots)
				return;
			const uint32_t q = i >> 2, slot = i & 3u;
			const uint32_t info = quadInfo[q];
			const SlotRef r = resolve_slot(d, codes, quadCell[q], info & 3u, slot);
			uint32_t vid;
			if (r.firstOrd == r.myOrd)
			{
				vid = vids[i];
				const V3 g = dual_point(d, dens, r.cx, r.cy, r.cz, r.pointCode);
				rawVerts[vid * 3 + 0] = g.x;
				rawVerts[vid * 3 + 1] = g.y;
				rawVerts[vid * 3 + 2] = g.z;
				const V3 w = g * wm.mult + wm.add;
				worldVerts[vid * 3 + 0] = w.x;
				worldVerts[vid * 3 + 1] = w.y;
				worldVerts[vid * 3 + 2] = w.z;
				vertCell[vid] = cell_index(d, r.cx, r.cy, r.cz);
				vertCode[vid] = (uint16_t)r.pointCode;
			}
			else
			{
				const uint32_t e = r.firstEdge;
				const uint32_t qf = quad_lookup(d, codes, quadBase, r.cx + c_edgeBase[e][0], r.cy + c_edgeBase[e][1], r.cz + c_edgeBase[e][2], c_edgeAxis[e]);
				vid = vids[qf * 4 + c_edgeSlot[e]];
			}
			// emitted order: (i0,i1,i2,i3) or (i0,i3,i2,i1)
			const uint32_t pos = (info & 4u) ? slot : ((4u - slot) & 3u);
			quadFinal[q * 4 + pos] = vid;
		}

		// ---- pass 5: triangles (marchingCubes.cpp:198-221) ----------------------------------------------------
		__device__ __forceinline__ V3 load3(const float *p, uint32_t i) { return v3(p[i * 3 + 0], p[i * 3 + 1], p[i * 3 + 2]); }

		// triInfo[q]: bit0 = first triangle kept, bit1 = second kept, bit2 = `which`; faceN[q*6..] = the two cross products
		__global__ void __launch_bounds__(DMC_BLOCK) dmc_triangles(const float *__restrict__ worldVerts, const uint32_t *__restrict__ quadFinal, uint32_t nQuads, uint8_t *__restrict__ triInfo, float *__restrict__ faceN, uint32_t *__restrict__ triCount)
		{
			const uint32_t q = blockIdx.x * DMC_BLOCK + threadIdx.x;
			if (q >= nQuads)
				return;
			uint32_t is[4];
			V3 p[4];
#pragma unroll
			for (int k = 0; k < 4; k++)
			{
				is[k] = quadFinal[q * 4 + k];
				p[k] = load3(worldVerts, is[k]);
			}
			const bool which = distance_squared(p[0], p[2]) < distance_squared(p[1], p[3]); // shorter diagonal
			const int a0 = which ? 0 : 1, a1 = which ? 1 : 2, a2 = which ? 2 : 3;
			const int b0 = which ? 0 : 1, b1 = which ? 2 : 3, b2 = which ? 3 : 0;
			uint32_t info = which ? 4u : 0u;
			uint32_t n = 0;
			{
				const V3 c = cross(p[a1] - p[a0], p[a2] - p[a0]);
				const bool degenerated = length(c) * 0.5f < 1e-5f; // Triangle::degenerated, shapes.cpp:119-122, geometry.h:89
				if (!degenerated)
				{
					info |= 1u;
					n++;
				}
				faceN[q * 6 + 0] = c.x;
				faceN[q * 6 + 1] = c.y;
				faceN[q * 6 + 2] = c.z;
			}
			{
				const V3 c = cross(p[b1] - p[b0], p[b2] - p[b0]);
				const bool degenerated = length(c) * 0.5f < 1e-5f;
				if (!degenerated)
				{
					info |= 2u;
					n++;
				}
				faceN[q * 6 + 3] = c.x;
				faceN[q * 6 + 4] = c.y;
				faceN[q * 6 + 5] = c.z;
			}
			triInfo[q] = (uint8_t)info;
			triCount[q] = n;
		}

		__global__ void __launch_bounds__(DMC_BLOCK) dmc_emit_indices(const uint32_t *__restrict__ quadFinal, const uint8_t *__restrict__ triInfo, const uint32_t *__restrict__ triOffset, uint32_t nQuads, uint32_t *__restrict__ indices)
		{
			const uint32_t q = blockIdx.x * DMC_BLOCK + threadIdx.x;
			if (q >= nQuads)
				return;
			const uint32_t info = triInfo[q];
			const bool which = info & 4u;
			uint32_t is[4];
#pragma unroll
			for (int k = 0; k < 4; k++)
				is[k] = quadFinal[q * 4 + k];
			uint32_t t = triOffset[q];
			if (info & 1u)
			{
				indices[t * 3 + 0] = is[which ? 0 : 1];
				indices[t * 3 + 1] = is[which ? 1 : 2];
				indices[t * 3 + 2] = is[which ? 2 : 3];
				t++;
			}
			if (info & 2u)
			{
				indices[t * 3 + 0] = is[which ? 0 : 1];
				indices[t * 3 + 1] = is[which ? 2 : 3];
				indices[t * 3 + 2] = is[which ? 3 : 0];
			}
		}

		// ---- pass 6: vertex normals (marchingCubes.cpp:214-227), gathered in the reference's accumulation order ----
		__global__ void __launch_bounds__(DMC_BLOCK) dmc_normals(Dims d, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ quadBase, const uint32_t *__restrict__ vertCell, const uint16_t *__restrict__ vertCode, const uint32_t *__restrict__ quadFinal, const uint8_t *__restrict__ triInfo, const float *__restrict__ faceN, uint32_t nVerts, float *__restrict__ normals)
		{
			const uint32_t v = blockIdx.x * DMC_BLOCK + threadIdx.x;
			if (v >= nVerts)
				return;
			int32_t cx, cy, cz;
			cell_coords(d, vertCell[v], cx, cy, cz);
			// the quads around this dual point, sorted by scan ordinal (insertion sort, <= 12 entries)
			uint64_t ord[12];
			uint32_t edge[12];
			int n = 0;
			uint32_t pc = vertCode[v];
			while (pc)
			{
				const uint32_t k = __ffs(pc) - 1;
				pc &= pc - 1;
				const uint64_t o = slot_ordinal(d, cx, cy, cz, k);
				if (o == ~0ull)
					continue;
				int j = n++;
				while (j > 0 && ord[j - 1] > o)
				{
					ord[j] = ord[j - 1];
					edge[j] = edge[j - 1];
					j--;
				}
				ord[j] = o;
				edge[j] = k;
			}
			V3 acc = v3(0, 0, 0);
			for (int j = 0; j < n; j++)
			{
				const uint32_t e = edge[j];
				const uint32_t q = quad_lookup(d, codes, quadBase, cx + c_edgeBase[e][0], cy + c_edgeBase[e][1], cz + c_edgeBase[e][2], c_edgeAxis[e]);
				const uint32_t info = triInfo[q];
				const bool which = info & 4u;
				int k = 0; // my position in the emitted quad
#pragma unroll
				for (int m = 0; m < 4; m++)
					if (quadFinal[q * 4 + m] == v)
						k = m;
				// first triangle: which ? (0,1,2) : (1,2,3); second: which ? (0,2,3) : (1,3,0)
				const bool inA = which ? (k != 3) : (k != 0);
				const bool inB = which ? (k != 1) : (k != 2);
				if ((info & 1u) && inA)
					acc = acc + load3(faceN, q * 2 + 0);
				if ((info & 2u) && inB)
					acc = acc + load3(faceN, q * 2 + 1);
			}
			if (!(acc.x == 0 && acc.y == 0 && acc.z == 0))
				acc = normalize(acc);
			normals[v * 3 + 0] = acc.x;
			normals[v * 3 + 1] = acc.y;
			normals[v * 3 + 2] = acc.z;
		}

		// slab extraction: the quads anchored at local z in [lo, hi) form one contiguous block of the ordered quad list; vertices
		// are numbered in creation (= scan) order and triangles follow their quads, so the same holds for both.
		// out = { firstQuad, endQuad, firstVertex, endVertex, firstTriangle, endTriangle }
		__global__ void dmc_owned_ranges(Dims d, const uint32_t *__restrict__ quadCell, uint32_t nQuads, const uint32_t *__restrict__ vids, uint32_t nVerts, const uint32_t *__restrict__ triOffset, uint32_t nTris, int32_t lo, int32_t hi, uint32_t *__restrict__ out)
		{
			if (threadIdx.x != 0 || blockIdx.x != 0)
				return;
			const uint64_t perSlice = (uint64_t)d.cx() * d.cy();
			for (int k = 0; k < 2; k++)
			{
				const int32_t z = k == 0 ? lo : hi;
				const uint64_t key = z <= 0 ? 0 : (uint64_t)z * perSlice; // first cell index of slice z
				uint32_t a = 0, b = nQuads;
				while (a < b)
				{
					const uint32_t m = a + (b - a) / 2;
					if ((uint64_t)quadCell[m] < key)
						a = m + 1;
					else
						b = m;
				}
				out[k] = a;
				out[2 + k] = a < nQuads ? vids[a * 4] : nVerts;
				out[4 + k] = a < nQuads ? triOffset[a] : nTris;
			}
		}

		// vertex ids relative to the first owned vertex (two's complement: ids below it belong to the previous slab)
		__global__ void __launch_bounds__(DMC_BLOCK) dmc_rebase(uint32_t *__restrict__ ids, uint64_t n, uint32_t first)
		{
			const uint64_t i = (uint64_t)blockIdx.x * DMC_BLOCK + threadIdx.x;
			if (i < n)
				ids[i] -= first;
		}

		template<class T>
		cudaError_t dalloc(T **p, uint64_t count, cudaStream_t s)
		{
			*p = nullptr;
			return device_alloc(p, sizeof(T) * (count ? count : 1), s);
		}

		cudaError_t exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t s)
		{
			size_t bytes = 0;
			cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, s);
			if (e != cudaSuccess)
				return e;
			void *tmp = nullptr;
			e = device_alloc(&tmp, bytes ? bytes : 1, s);
			if (e != cudaSuccess)
				return e;
			e = cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, s);
			cudaFreeAsync(tmp, s);
			return e;
		}

		// last element of an exclusive scan + last input = total
		cudaError_t scan_total(const uint32_t *in, const uint32_t *scanned, uint64_t n, uint32_t *total, cudaStream_t s)
		{
			*total = 0;
			if (n == 0)
				return cudaSuccess;
			uint32_t a = 0, b = 0;
			cudaError_t e = cudaMemcpyAsync(&a, in + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(&b, scanned + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = stream_sync(s);
			*total = a + b;
			return e;
		}
	}

	// Everything on `s`; the caller owns `dens` (rx*ry*rz floats on the device).  Outputs are device arrays allocated
	// with cudaMallocAsync on `s` (released by dmc_release).
	int dmc_extract(const float *dens, int rx, int ry, int rz, float boxSize, int fullRes, cudaStream_t s, DmcResult &out, int zOffset, int ownZ0, int ownZ1, const SignBits *bits)
	{
		out = DmcResult();
		const Dims d{ rx, ry, rz, zOffset };
		if (rx <= 2 || ry <= 2 || rz <= 2)
			return fail(UPG_ERR_ARGUMENT, "volume too small for dual marching cubes");
		const uint64_t cells = d.cells();
		if (cells >= (1ull << 32))
			return fail(UPG_ERR_ARGUMENT, "volume too large (cell index must fit 32 bits)");
		// the dense passes run one block per row of cells, a thread per four cells
		const uint32_t blocks = (uint32_t)d.rows();
		int rowThreads = ((d.cx() + 3) / 4 + 31) / 32 * 32;
		rowThreads = rowThreads > DMC_ROW_MAX_THREADS ? DMC_ROW_MAX_THREADS : rowThreads;

		uint8_t *codes = nullptr;
		uint32_t *blockQuads = nullptr, *blockOffsets = nullptr, *quadBase = nullptr;
		unsigned long long *nearZero = nullptr;
		uint32_t *quadCell = nullptr, *flags = nullptr, *vids = nullptr, *vertCell = nullptr, *quadFinal = nullptr, *triCount = nullptr, *triOffset = nullptr;
		uint8_t *quadInfo = nullptr, *triInfo = nullptr;
		uint16_t *vertCode = nullptr;
		float *faceN = nullptr;
		cudaError_t e = cudaSuccess;
		int rc = UPG_OK;
		uint32_t nQuads = 0, nVerts = 0, nTris = 0;

#define UPG_TRY(x) \
	do \
	{ \
		if (e == cudaSuccess && rc == UPG_OK) \
			e = (x); \
	} while (0)
#define UPG_LAUNCHED() \
	do \
	{ \
		if (e == cudaSuccess && rc == UPG_OK) \
			e = cudaGetLastError(); \
	} while (0)

		UPG_TRY(dalloc(&codes, d.rows() * (uint64_t)d.codePitch(), s));
		UPG_TRY(dalloc(&blockQuads, blocks, s));
		UPG_TRY(dalloc(&blockOffsets, blocks, s));
		UPG_TRY(dalloc(&quadBase, cells, s));
		UPG_TRY(dalloc(&nearZero, 1, s));
		UPG_TRY(cudaMemsetAsync(nearZero, 0, sizeof(unsigned long long), s));
		if (e == cudaSuccess)
		{
			if (bits && bits->signs)
				dmc_cell_codes_bits<<<(blocks + DMC_BITS_WARPS - 1) / DMC_BITS_WARPS, DMC_BITS_WARPS * 32, 0, s>>>(bits->signs, bits->near, bits->pitch, d, codes, blockQuads, nearZero);
			else
				dmc_cell_codes<<<blocks, rowThreads, 0, s>>>(dens, d, codes, blockQuads, nearZero, NEAR_ZERO_TOLERANCE);
			UPG_LAUNCHED();
		}
		UPG_TRY(exclusive_scan_u32(blockQuads, blockOffsets, blocks, s));
		UPG_TRY(scan_total(blockQuads, blockOffsets, blocks, &nQuads, s));
		if (e == cudaSuccess)
		{
			unsigned long long nz = 0;
			e = cudaMemcpyAsync(&nz, nearZero, sizeof(nz), cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = stream_sync(s);
			out.nearZeroSamples = nz;
		}

		const uint32_t nSlots = nQuads * 4;
		if (e == cudaSuccess && nQuads > 0)
		{
			UPG_TRY(dalloc(&quadCell, nQuads, s));
			UPG_TRY(dalloc(&quadInfo, nQuads, s));
			UPG_TRY(dalloc(&flags, nSlots, s));
			UPG_TRY(dalloc(&vids, nSlots, s));
			if (e == cudaSuccess)
			{
				dmc_emit_quads<<<(blocks + DMC_EMIT_WARPS - 1) / DMC_EMIT_WARPS, DMC_EMIT_WARPS * 32, 0, s>>>(d, codes, blockOffsets, blockQuads, quadCell, quadInfo, quadBase);
				UPG_LAUNCHED();
			}
			const uint32_t slotBlocks = (nSlots + DMC_BLOCK - 1) / DMC_BLOCK;
			if (e == cudaSuccess)
			{
				dmc_slot_flags<<<slotBlocks, DMC_BLOCK, 0, s>>>(d, codes, quadCell, quadInfo, nSlots, flags);
				UPG_LAUNCHED();
			}
			UPG_TRY(exclusive_scan_u32(flags, vids, nSlots, s));
			UPG_TRY(scan_total(flags, vids, nSlots, &nVerts, s));

			WorldMap wm;
			{
				// posMult = box.size() / (Vec3(resolution) - 5); posAdd = position(0,0,0) = size * ((0 - 2) / (res - 5)) + a
				const float a = boxSize * -0.5f, b = boxSize * 0.5f;
				const float size = b - a;
				const float m = size / ((float)fullRes - 5.0f);
				const float f = (0.0f - 2.0f) / ((float)fullRes - 5.0f);
				const float add = size * f + a;
				wm.mult = v3(m, m, m);
				wm.add = v3(add, add, add);
			}
			UPG_TRY(dalloc(&out.rawVertices, (uint64_t)nVerts * 3, s));
			UPG_TRY(dalloc(&out.positions, (uint64_t)nVerts * 3, s));
			UPG_TRY(dalloc(&out.normals, (uint64_t)nVerts * 3, s));
			UPG_TRY(dalloc(&vertCell, nVerts, s));
			UPG_TRY(dalloc(&vertCode, nVerts, s));
			UPG_TRY(dalloc(&quadFinal, nSlots, s));
			if (e == cudaSuccess)
			{
				dmc_build<<<slotBlocks, DMC_BLOCK, 0, s>>>(d, dens, codes, quadCell, quadInfo, quadBase, vids, nSlots, wm, out.rawVertices, out.positions, vertCell, vertCode, quadFinal);
				UPG_LAUNCHED();
			}
			const uint32_t quadBlocks = (nQuads + DMC_BLOCK - 1) / DMC_BLOCK;
			UPG_TRY(dalloc(&triInfo, nQuads, s));
			UPG_TRY(dalloc(&faceN, (uint64_t)nQuads * 6, s));
			UPG_TRY(dalloc(&triCount, nQuads, s));
			UPG_TRY(dalloc(&triOffset, nQuads, s));
			if (e == cudaSuccess)
			{
				dmc_triangles<<<quadBlocks, DMC_BLOCK, 0, s>>>(out.positions, quadFinal, nQuads, triInfo, faceN, triCount);
				UPG_LAUNCHED();
			}
			UPG_TRY(exclusive_scan_u32(triCount, triOffset, nQuads, s));
			UPG_TRY(scan_total(triCount, triOffset, nQuads, &nTris, s));
			UPG_TRY(dalloc(&out.indices, (uint64_t)nTris * 3, s));
			if (e == cudaSuccess && nTris > 0)
			{
				dmc_emit_indices<<<quadBlocks, DMC_BLOCK, 0, s>>>(quadFinal, triInfo, triOffset, nQuads, out.indices);
				UPG_LAUNCHED();
			}
			if (e == cudaSuccess && nVerts > 0)
			{
				dmc_normals<<<(nVerts + DMC_BLOCK - 1) / DMC_BLOCK, DMC_BLOCK, 0, s>>>(d, codes, quadBase, vertCell, vertCode, quadFinal, triInfo, faceN, nVerts, out.normals);
				UPG_LAUNCHED();
			}
			if (e == cudaSuccess)
			{
				out.quads = quadFinal; // dualmc's quads, in emitted order
				quadFinal = nullptr;
			}
			// slab: keep the blocks owned by scan iterations [ownZ0, ownZ1) (global z), ids relative to the first owned vertex
			if (e == cudaSuccess && (ownZ0 > zOffset || ownZ1 < zOffset + rz))
			{
				uint32_t *ranges = nullptr;
				uint32_t h[6] = { 0, 0, 0, 0, 0, 0 };
				UPG_TRY(dalloc(&ranges, 6, s));
				if (e == cudaSuccess)
				{
					dmc_owned_ranges<<<1, 1, 0, s>>>(d, quadCell, nQuads, vids, nVerts, triOffset, nTris, ownZ0 - zOffset, ownZ1 - zOffset, ranges);
					UPG_LAUNCHED();
				}
				UPG_TRY(cudaMemcpyAsync(h, ranges, sizeof(h), cudaMemcpyDeviceToHost, s));
				UPG_TRY(stream_sync(s));
				if (ranges)
					cudaFreeAsync(ranges, s);
				if (e == cudaSuccess)
				{
					out.firstQuad = h[0];
					out.ownedQuads = h[1] - h[0];
					out.firstVertex = h[2];
					out.ownedVertices = h[3] - h[2];
					out.firstTriangle = h[4];
					out.ownedTriangles = h[5] - h[4];
					out.slab = true;
					if (out.ownedQuads)
					{
						dmc_rebase<<<(out.ownedQuads * 4 + DMC_BLOCK - 1) / DMC_BLOCK, DMC_BLOCK, 0, s>>>(out.quads + (uint64_t)out.firstQuad * 4, (uint64_t)out.ownedQuads * 4, out.firstVertex);
						UPG_LAUNCHED();
					}
					if (out.ownedTriangles)
					{
						dmc_rebase<<<(out.ownedTriangles * 3 + DMC_BLOCK - 1) / DMC_BLOCK, DMC_BLOCK, 0, s>>>(out.indices + (uint64_t)out.firstTriangle * 3, (uint64_t)out.ownedTriangles * 3, out.firstVertex);
						UPG_LAUNCHED();
					}
				}
			}
		}
#undef UPG_TRY
#undef UPG_LAUNCHED

		out.vertexCount = nVerts;
		out.quadCount = nQuads;
		out.triangleCount = nTris;
		if (!out.slab)
		{
			out.ownedVertices = nVerts;
			out.ownedQuads = nQuads;
			out.ownedTriangles = nTris;
		}

		void *scratch[] = { codes, blockQuads, blockOffsets, quadBase, nearZero, quadCell, quadInfo, flags, vids, vertCell, vertCode, quadFinal, triInfo, faceN, triCount, triOffset };
		for (void *p : scratch)
			if (p)
				cudaFreeAsync(p, s);
		if (rc == UPG_OK && e != cudaSuccess)
			rc = cuda_fail(e, "dmc_extract");
		if (rc != UPG_OK)
			dmc_release(out, s);
		return rc;
	}

	int rebase_ids(uint32_t *ids, uint64_t n, uint32_t add, cudaStream_t s)
	{
		if (n == 0 || add == 0)
			return UPG_OK;
		dmc_rebase<<<(unsigned)((n + DMC_BLOCK - 1) / DMC_BLOCK), DMC_BLOCK, 0, s>>>(ids, n, 0u - add); // the kernel subtracts
		const cudaError_t e = cudaGetLastError();
		return e == cudaSuccess ? UPG_OK : cuda_fail(e, "rebase_ids");
	}

	void dmc_release(DmcResult &r, cudaStream_t s)
	{
		void *p[] = { r.rawVertices, r.positions, r.normals, r.indices, r.quads };
		for (void *x : p)
			if (x)
				cudaFreeAsync(x, s);
		r = DmcResult();
	}
}
