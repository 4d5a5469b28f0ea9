// upg_bake.cu -- K5..K7: per-texel baking of albedo / special (roughness, metallic) / height textures.
//
// Replaces Generator<Water>::generate (sources/textures.cpp:45-82):
//   K5  raster_owner      meshGenerateTexture's scan-line coverage (externals/cage/sources/libcore/mesh/meshAlgorithms.cpp:
//                         1676-1729).  The serial reference lets a later triangle overwrite an earlier one; here every
//                         triangle walks exactly the reference's integer scan lines and atomicMax'es its index into an
//                         owner image, so each covered texel is evaluated ONCE, by the triangle that would have won.
//   --  compaction        covered texels -> dense list in raster order (CUB DeviceSelect), so that the expensive kernel
//                         runs without idle lanes and neighbouring lanes stay neighbours on the surface.
//   K6a tile_prepare      barycentric weights of the owner (barycoord, :29-50), Mesh::positionAt / normalAt
//                         (mesh/meshAttributes.cpp:7-11,59-63), and the nine positions generateSlope / generateElevation
//                         need (sources/coloringModes.cpp:51-57,122-151)
//   K1  sdf_sample        terrainSdfElevation over those positions (explicit-position mode of upg_field.cuh)
//   K6b tile_color        terrainTile -> coloringDefault / Barren / Debug (upg_coloring.cuh); compute bound
//   K7  dilate7_quantise  imageDilation(img, 7, true) + imageConvert(U8) fused: a 64x32 texel tile plus a 7 texel apron is
//                         staged in shared memory, the seven rounds run there, one 8-bit store per channel leaves the SM
//                         (libcore/image/imageAlgorithms.cpp:349-527, image.cpp:244-262).  HBM bound: 4*CH B read (x1.75
//                         apron, mostly L2 hits) + CH B written per texel instead of 7 x 8*CH B for seven global rounds.
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <cuda_runtime.h>

#include <cstring>

#include "../../include/upg.h"
#include "upg_coloring.cuh"
#include "upg_internal.h"

namespace upg
{
	namespace
	{
		constexpr int BAKE_BLOCK = 128;
		constexpr uint32_t NO_OWNER = 0; // owner image holds triangle index + 1

		struct RasterMesh
		{
			const float *uvs; // vertex_count x 2
			const uint32_t *indices; // 3 per triangle
			uint32_t triangles;
			uint32_t width, height;
			uint32_t bandBegin, bandEnd; // rows [bandBegin, bandEnd) are kept (row-tile sharding); the owner image is band sized
			// index validation happens here, where the indices are read anyway: a serial host loop over the 8.7 M indices of a 1000^3 base
			// mesh cost 6 ms per bake call.  0 = not checked (the caller did); else triangles with an index >= vertices are skipped and
			// the largest offending index + 1 lands in *badIndex
			uint32_t vertices = 0;
			uint32_t *badIndex = nullptr;
		};

		// Vec2i * Real (meshAlgorithms.cpp:52-55): per component sint32(int * float)
		__device__ __forceinline__ int2 scale_trunc(int2 a, float b) { return make_int2((int)((float)a.x * b), (int)((float)a.y * b)); }

		// the scan of ONE triangle exactly as the reference walks it; `visit(x, y)` for every texel in order
		template<class Visit>
		__device__ __forceinline__ void scan_triangle(const RasterMesh &M, uint32_t tri, Visit visit)
		{
			const uint32_t i0 = M.indices[tri * 3 + 0], i1 = M.indices[tri * 3 + 1], i2 = M.indices[tri * 3 + 2];
			const float sx = (float)M.width, sy = (float)M.height;
			const float2 u0 = make_float2(M.uvs[i0 * 2] * sx, M.uvs[i0 * 2 + 1] * sy);
			const float2 u1 = make_float2(M.uvs[i1 * 2] * sx, M.uvs[i1 * 2 + 1] * sy);
			const float2 u2 = make_float2(M.uvs[i2 * 2] * sx, M.uvs[i2 * 2 + 1] * sy);
			int2 t0 = make_int2((int)u0.x, (int)u0.y);
			int2 t1 = make_int2((int)u1.x, (int)u1.y);
			int2 t2 = make_int2((int)u2.x, (int)u2.y);
			int2 tmp;
			if (t0.y > t1.y)
			{
				tmp = t0;
				t0 = t1;
				t1 = tmp;
			}
			if (t0.y > t2.y)
			{
				tmp = t0;
				t0 = t2;
				t2 = tmp;
			}
			if (t1.y > t2.y)
			{
				tmp = t1;
				t1 = t2;
				t2 = tmp;
			}
			const int totalHeight = t2.y - t0.y;
			const float totalHeightInv = 1.f / (float)totalHeight;
			const int2 d20 = make_int2(t2.x - t0.x, t2.y - t0.y);
			const int2 d21 = make_int2(t2.x - t1.x, t2.y - t1.y);
			const int2 d10 = make_int2(t1.x - t0.x, t1.y - t0.y);
			// scan lines are independent of each other: rows outside the band are skipped, not walked
			if (t2.y < (int)M.bandBegin || t0.y >= (int)M.bandEnd)
				return;
			const int iBegin = t0.y < (int)M.bandBegin ? (int)M.bandBegin - t0.y : 0;
			const int iEnd = t2.y >= (int)M.bandEnd ? (int)M.bandEnd - 1 - t0.y : totalHeight;
			for (int i = iBegin; i <= iEnd; i++)
			{
				const bool secondHalf = i > t1.y - t0.y || t1.y == t0.y;
				const uint32_t segmentHeight = secondHalf ? (uint32_t)(t2.y - t1.y) : (uint32_t)(t1.y - t0.y);
				const float alpha = totalHeight ? (float)i * totalHeightInv : 0.0f;
				const float beta = totalHeight ? (float)(i - (secondHalf ? t1.y - t0.y : 0)) / (float)segmentHeight : 0.0f;
				const int2 sa = scale_trunc(d20, alpha);
				int2 A = make_int2(t0.x + sa.x, t0.y + sa.y);
				int2 B;
				if (secondHalf)
				{
					const int2 sb = scale_trunc(d21, beta);
					B = make_int2(t1.x + sb.x, t1.y + sb.y);
				}
				else
				{
					const int2 sb = scale_trunc(d10, beta);
					B = make_int2(t0.x + sb.x, t0.y + sb.y);
				}
				if (A.x > B.x)
				{
					tmp = A;
					A = B;
					B = tmp;
				}
				const int y = t0.y + i;
				for (int x = A.x; x <= B.x; x++)
					visit(x, y);
			}
		}

		// K5: one thread per triangle (a triangle covers ~10^1..10^2 texels at the generator's 1.35 texels per unit)
		__global__ void __launch_bounds__(BAKE_BLOCK) raster_owner_kernel(const RasterMesh M, uint32_t *__restrict__ owner, unsigned long long *__restrict__ visits)
		{
			const uint32_t tri = blockIdx.x * BAKE_BLOCK + threadIdx.x;
			unsigned long long mine = 0;
			bool valid = tri < M.triangles;
			if (valid && M.vertices)
			{
				const uint32_t i0 = M.indices[tri * 3 + 0], i1 = M.indices[tri * 3 + 1], i2 = M.indices[tri * 3 + 2];
				const uint32_t top = max(i0, max(i1, i2));
				if (top >= M.vertices)
				{
					atomicMax(M.badIndex, top == 0xffffffffu ? top : top + 1);
					valid = false;
				}
			}
			if (valid)
			{
				scan_triangle(M, tri,
					[&](int x, int y)
					{
						mine++;
						// the reference asserts the texel is inside the image and otherwise writes out of bounds; skip instead
						if (x >= 0 && (uint32_t)x < M.width && y >= (int)M.bandBegin && y < (int)M.bandEnd)
							atomicMax(&owner[(size_t)(y - (int)M.bandBegin) * M.width + x], tri + 1);
					});
			}
			// one atomic per warp
			for (int o = 16; o > 0; o >>= 1)
				mine += __shfl_down_sync(0xffffffffu, mine, o);
			if ((threadIdx.x & 31) == 0 && mine)
				atomicAdd(visits, mine);
		}

		// barycoord (meshAlgorithms.cpp:29-50) for Vec2
		__device__ __forceinline__ V2 barycoord2(V2 a, V2 b, V2 c, V2 p)
		{
			const V2 v0 = b - a;
			const V2 v1 = c - a;
			const V2 v2_ = p - a;
			const float d00 = dot(v0, v0);
			const float d01 = dot(v0, v1);
			const float d11 = dot(v1, v1);
			const float d20 = dot(v2_, v0);
			const float d21 = dot(v2_, v1);
			const float denom = d00 * d11 - d01 * d01;
			if (abs_(denom) < 1e-7f)
				return v2(0, 0);
			const float invDenom = 1.0f / denom;
			const float v = (d11 * d20 - d01 * d21) * invDenom;
			const float w = (d00 * d21 - d01 * d20) * invDenom;
			const float u = 1 - v - w;
			return v2(u, v);
		}

		__device__ __forceinline__ V3 texel_weights(const RasterMesh &M, uint32_t tri, uint32_t x, uint32_t y)
		{
			const uint32_t i0 = M.indices[tri * 3 + 0], i1 = M.indices[tri * 3 + 1], i2 = M.indices[tri * 3 + 2];
			const float sx = (float)M.width, sy = (float)M.height;
			const V2 u0 = v2(M.uvs[i0 * 2] * sx, M.uvs[i0 * 2 + 1] * sy);
			const V2 u1 = v2(M.uvs[i1 * 2] * sx, M.uvs[i1 * 2 + 1] * sy);
			const V2 u2 = v2(M.uvs[i2 * 2] * sx, M.uvs[i2 * 2 + 1] * sy);
			const V2 b = barycoord2(u0, u1, u2, v2((float)(int)x, (float)(int)y));
			return v3(b.x, b.y, 1 - b.x - b.y);
		}

		// parity helper of upg_raster: owner image -> 0xffffffff convention + weights of the winning visit
		__global__ void __launch_bounds__(BAKE_BLOCK) raster_weights_kernel(const RasterMesh M, const uint32_t *__restrict__ owner, uint32_t *__restrict__ ownerOut, float *__restrict__ weights)
		{
			const size_t i = (size_t)blockIdx.x * BAKE_BLOCK + threadIdx.x;
			if (i >= (size_t)M.width * M.height)
				return;
			const uint32_t o = owner[i];
			if (o == NO_OWNER)
			{
				ownerOut[i] = 0xffffffffu;
				weights[i * 3 + 0] = weights[i * 3 + 1] = weights[i * 3 + 2] = 0;
				return;
			}
			const V3 w = texel_weights(M, o - 1, (uint32_t)(i % M.width), (uint32_t)(i / M.width));
			ownerOut[i] = o - 1;
			weights[i * 3 + 0] = w.x;
			weights[i * 3 + 1] = w.y;
			weights[i * 3 + 2] = w.z;
		}

		// covered texels of the band that fall into the caller's rows (the band adds a 7-row apron for the dilation)
		__global__ void __launch_bounds__(256) count_in_range_kernel(const uint32_t *__restrict__ list, uint32_t n, uint32_t lo, uint32_t hi, unsigned long long *__restrict__ out)
		{
			// grid-stride, one atomic per block (one per warp made half a million same-address atomics of a 4096^2 bake: 0.36 ms)
			__shared__ uint32_t warpSum[8];
			uint32_t mine = 0;
			for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
				mine += (list[i] >= lo && list[i] < hi) ? 1u : 0u;
			for (int o = 16; o > 0; o >>= 1)
				mine += __shfl_down_sync(0xffffffffu, mine, o);
			if ((threadIdx.x & 31) == 0)
				warpSum[threadIdx.x >> 5] = mine;
			__syncthreads();
			if (threadIdx.x == 0)
			{
				uint32_t total = 0;
				for (int k = 0; k < 8; k++)
					total += warpSum[k];
				if (total)
					atomicAdd(out, (unsigned long long)total);
			}
		}

		struct IsCovered
		{
			const uint32_t *owner;
			__device__ __forceinline__ bool operator()(uint32_t i) const { return i != 0xffffffffu && owner[i] != NO_OWNER; }
		};

		// The compacted list keeps the order in which the texels are offered to the selection.  Offering them tile by tile
		// (TEXEL_TILE_W x TEXEL_TILE_H texels, row-major inside a tile) gives the 32 texels of a colouring warp a compact footprint
		// on the surface -- ~8 x 4 world units instead of a 32 x 1 sliver -- so they meet the biome / slope masks of the layers and
		// the lattice cells of the noises (single_cubic's warp-cooperative rows) together more often.  Texels are independent and
		// written by id: the images do not depend on this order.  Slot -> texel id, 0xffffffff for the padding past the band's edges.
#if !defined(UPG_TEXEL_TILE_W)
#define UPG_TEXEL_TILE_W 8
#define UPG_TEXEL_TILE_H 4
#endif
		constexpr uint32_t TEXEL_TILE_W = UPG_TEXEL_TILE_W, TEXEL_TILE_H = UPG_TEXEL_TILE_H;
		struct TileOrder
		{
			uint32_t w, h, tilesX;
			__host__ __device__ __forceinline__ uint32_t operator()(uint32_t slot) const
			{
				const uint32_t tile = slot / (TEXEL_TILE_W * TEXEL_TILE_H), in = slot % (TEXEL_TILE_W * TEXEL_TILE_H);
				const uint32_t x = (tile % tilesX) * TEXEL_TILE_W + in % TEXEL_TILE_W;
				const uint32_t y = (tile / tilesX) * TEXEL_TILE_H + in / TEXEL_TILE_W;
				return x < w && y < h ? y * w + x : 0xffffffffu;
			}
		};

		// K6a.  Probe layout: probes[(k * n + i) * 3 ..], k = 0..7 the slope ring, k = 8 the tile position.
		struct PrepareArgs
		{
			RasterMesh mesh;
			const float *positions; // mesh vertices
			const float *normals;
			const uint32_t *owner;
			const uint32_t *texels; // this chunk of the compacted list
			uint32_t n;
			float *probes; // 9 * n * 3
			float *tileNormals; // n * 3
		};

		__device__ __forceinline__ V3 load3(const float *p, uint32_t i) { return v3(p[(size_t)i * 3 + 0], p[(size_t)i * 3 + 1], p[(size_t)i * 3 + 2]); }
		__device__ __forceinline__ void store3(float *p, size_t i, V3 v)
		{
			p[i * 3 + 0] = v.x;
			p[i * 3 + 1] = v.y;
			p[i * 3 + 2] = v.z;
		}

		__device__ __forceinline__ void write_probes(float *probes, uint32_t n, uint32_t i, V3 position, V3 normal)
		{
#pragma unroll
			for (int k = 0; k < 9; k++)
				store3(probes, (size_t)k * n + i, slope_probe(position, normal, k));
		}

		__global__ void __launch_bounds__(BAKE_BLOCK) tile_prepare_kernel(const PrepareArgs A)
		{
			const uint32_t i = blockIdx.x * BAKE_BLOCK + threadIdx.x;
			if (i >= A.n)
				return;
			const uint32_t texel = A.texels[i];
			const uint32_t tri = A.owner[texel] - 1;
			const V3 w = texel_weights(A.mesh, tri, texel % A.mesh.width, texel / A.mesh.width + A.mesh.bandBegin);
			const uint32_t i0 = A.mesh.indices[tri * 3 + 0], i1 = A.mesh.indices[tri * 3 + 1], i2 = A.mesh.indices[tri * 3 + 2];
			// attrAt: at(ids[0]) * w[0] + at(ids[1]) * w[1] + at(ids[2]) * w[2]; normals additionally normalised
			const V3 position = (load3(A.positions, i0) * w.x + load3(A.positions, i1) * w.y) + load3(A.positions, i2) * w.z;
			const V3 normal = normalize((load3(A.normals, i0) * w.x + load3(A.normals, i1) * w.y) + load3(A.normals, i2) * w.z);
			store3(A.tileNormals, i, normal);
			write_probes(A.probes, A.n, i, position, normal);
		}

		// explicit tiles (upg_tiles): positions / normals are given
		__global__ void __launch_bounds__(BAKE_BLOCK) tile_probes_kernel(const float *__restrict__ positions, const float *__restrict__ normals, uint32_t n, float *__restrict__ probes)
		{
			const uint32_t i = blockIdx.x * BAKE_BLOCK + threadIdx.x;
			if (i >= n)
				return;
			write_probes(probes, n, i, load3(positions, i), load3(normals, i));
		}

		// K6b
		struct ColorArgs
		{
			const float *probes; // 9 * n * 3 (k = 8: positions)
			const float *normals; // n * 3
			const float *elev9; // 9 * n: terrainSdfElevation at the probes
			const float *elevRaw; // n: terrainSdfElevationRaw at the positions (water / navigation purpose), or nullptr
			const float *cellular; // 1024 floats
			uint32_t n;
			int32_t purpose;
			// image outputs (bake) ...
			const uint32_t *texels;
			float *albedo;
			float *special;
			float *height;
			int32_t albedoChannels;
			// ... or tile records (upg_tiles)
			upg_tile *tiles;
			uint32_t *status;
		};

		// 80 registers per thread leave room for 768 resident threads per SM.  Two blocks of 384 beat one of 768 by 3.5 % (A/B on B200,
		// profiles/r01_l_ab_color_block.txt: 15.19 vs 15.74 ms per 16.5 M-texel land bake; 3 x 256: 16.7, 1 x 512: 16.2): while one
		// block waits at a layer barrier the other issues, and both still share the instruction cache.
#ifndef UPG_COLOR_BLOCK
#define UPG_COLOR_BLOCK 384
#endif
#ifndef UPG_COLOR_BLOCKS_PER_SM
#define UPG_COLOR_BLOCKS_PER_SM 2
#endif
		constexpr int COLOR_BLOCK = UPG_COLOR_BLOCK;

		template<int COLORING>
		__global__ void __launch_bounds__(COLOR_BLOCK, UPG_COLOR_BLOCKS_PER_SM) tile_color_kernel(const __grid_constant__ ColoringConsts C, const __grid_constant__ ColorArgs A)
		{
			// the jitter table is looked up with divergent indices: keep it in shared memory
			__shared__ float rv[1024];
			for (int k = threadIdx.x; k < 1024; k += COLOR_BLOCK)
				rv[k] = A.cellular[k];
			grad_table_init<COLOR_BLOCK>(); // gradient table of the Perlin / OpenSimplex generators (upg_noise.cuh)
			__syncthreads();
			// threads past the end keep running (on the last tile) so that every thread reaches the barriers between layers
			const uint32_t self = blockIdx.x * COLOR_BLOCK + threadIdx.x;
			const bool live = self < A.n;
			const uint32_t i = live ? self : A.n - 1;
			const size_t n = A.n;
			Tile t = tile_init(load3(A.probes, 8 * n + i), load3(A.normals, i), A.purpose);
			// generateElevation (coloringModes.cpp:51-57)
			t.elevation = A.purpose == MP_LAND ? A.elev9[8 * n + i] : A.elevRaw[i];
			float elevs[8];
#pragma unroll
			for (int k = 0; k < 8; k++)
				elevs[k] = A.elev9[k * n + i];
			t.slope = slope_from_probes(elevs);
			if (isnan_(t.elevation) || isnan_(t.slope))
				atomicOr(A.status, 1u); // "invalid elevation sdf value" (sources/modesConfiguration.cpp:233-234)
			coloring_run<COLORING>(C, rv, t);
			if (!live)
				return;
			if (A.tiles)
			{
				upg_tile &o = A.tiles[i];
				o.albedo[0] = t.albedo.x;
				o.albedo[1] = t.albedo.y;
				o.albedo[2] = t.albedo.z;
				o.roughness = t.roughness;
				o.metallic = t.metallic;
				o.height = t.height;
				o.elevation = t.elevation;
				o.slope = t.slope;
				o.temperature = t.temperature;
				o.precipitation = t.precipitation;
				o.opacity = t.opacity;
				o.biome = (float)t.biome;
				o.type = (float)t.type;
			}
			else
			{
				// Generator::pixel (sources/textures.cpp:33-42)
				const size_t texel = A.texels[i];
				float *a = A.albedo + texel * A.albedoChannels;
				a[0] = t.albedo.x;
				a[1] = t.albedo.y;
				a[2] = t.albedo.z;
				if (A.albedoChannels == 4)
					a[3] = t.opacity;
				A.special[texel * 2 + 0] = t.roughness;
				A.special[texel * 2 + 1] = t.metallic;
				A.height[texel] = t.height;
			}
		}

		// ---- K7 -------------------------------------------------------------------------------------------
		constexpr int DIL_TW = 64, DIL_TH = 32, DIL_HALO = 7, DIL_THREADS = 256;
		constexpr int DIL_RW = DIL_TW + 2 * DIL_HALO, DIL_RH = DIL_TH + 2 * DIL_HALO;

		template<int CH>
		__device__ __forceinline__ bool px_valid(const float *p)
		{
			bool ok = true;
#pragma unroll
			for (int c = 0; c < CH; c++)
				ok = ok && !isnan_(p[c]);
			return ok;
		}

		// Image::value quantisation (image.cpp:244-262): numeric_cast<uint8>(saturate(v) * 255.f), truncating; NaN -> 0
		__device__ __forceinline__ uint8_t quantise_u8(float v) { return (uint8_t)(saturate_(v) * 255.f); }

		// one Dilation<T, true>::processPixel (imageAlgorithms.cpp:386-445) on a region held in shared memory
		template<int CH>
		__device__ __forceinline__ void dilate_px(const float *src, float *dst, int rx, int ry, int gx, int gy, int w, int h)
		{
			const float *c = src + (ry * DIL_RW + rx) * CH;
			float *o = dst + (ry * DIL_RW + rx) * CH;
			if (px_valid<CH>(c))
			{
#pragma unroll
				for (int k = 0; k < CH; k++)
					o[k] = c[k];
				return;
			}
			float m[CH];
#pragma unroll
			for (int k = 0; k < CH; k++)
				m[k] = 0;
			uint32_t cnt = 0;
			for (int dy = -1; dy <= 1; dy++)
			{
				if (gy + dy < 0 || gy + dy >= h)
					continue;
				for (int dx = -1; dx <= 1; dx++)
				{
					if (gx + dx < 0 || gx + dx >= w)
						continue;
					const float *a = src + ((ry + dy) * DIL_RW + (rx + dx)) * CH;
					if (px_valid<CH>(a))
					{
#pragma unroll
						for (int k = 0; k < CH; k++)
							m[k] += a[k];
						cnt++;
					}
				}
			}
			if (cnt > 0)
			{
				const float d = (float)cnt;
#pragma unroll
				for (int k = 0; k < CH; k++)
					o[k] = m[k] / d;
			}
			// else: the target keeps what it held two rounds ago, exactly like the reference's ping-pong pair
		}

		template<int CH>
		__global__ void __launch_bounds__(DIL_THREADS) dilate7_quantise_kernel(const float *__restrict__ img, int w, int h, uint8_t *__restrict__ out)
		{
			extern __shared__ float smem[];
			float *bufA = smem;
			float *bufB = smem + DIL_RW * DIL_RH * CH;
			const int x0 = blockIdx.x * DIL_TW - DIL_HALO;
			const int y0 = blockIdx.y * DIL_TH - DIL_HALO;
			const float nanv = __int_as_float(0x7fc00000);

			// stage the tile + apron; texels outside the image are NaN and never take part (see dilate_px)
			int innerInvalid = 0;
			for (int p = threadIdx.x; p < DIL_RW * DIL_RH; p += DIL_THREADS)
			{
				const int rx = p % DIL_RW, ry = p / DIL_RW;
				const int gx = x0 + rx, gy = y0 + ry;
				const bool inside = gx >= 0 && gy >= 0 && gx < w && gy < h;
				bool ok = inside;
#pragma unroll
				for (int k = 0; k < CH; k++)
				{
					const float v = inside ? img[((size_t)gy * w + gx) * CH + k] : nanv;
					bufA[p * CH + k] = v;
					bufB[p * CH + k] = nanv; // the reference's ping-pong target starts NaN-filled (imageAlgorithms.cpp:453-458)
					ok = ok && !isnan_(v);
				}
				const bool inner = rx >= DIL_HALO && rx < DIL_HALO + DIL_TW && ry >= DIL_HALO && ry < DIL_HALO + DIL_TH;
				if (inside && inner && !ok)
					innerInvalid = 1;
			}
			// valid texels never change: a tile without holes only needs the quantisation
			const int anyInvalid = __syncthreads_or(innerInvalid);

			float *src = bufA, *dst = bufB;
			if (anyInvalid)
			{
				for (int r = 0; r < 7; r++)
				{
					// after round r only texels within (6 - r) of the tile still matter
					const int m = DIL_HALO - 1 - r;
					const int rw = DIL_TW + 2 * m, rh = DIL_TH + 2 * m;
					for (int p = threadIdx.x; p < rw * rh; p += DIL_THREADS)
					{
						const int rx = DIL_HALO - m + p % rw, ry = DIL_HALO - m + p / rw;
						const int gx = x0 + rx, gy = y0 + ry;
						if (gx < 0 || gy < 0 || gx >= w || gy >= h)
							continue;
						dilate_px<CH>(src, dst, rx, ry, gx, gy, w, h);
					}
					__syncthreads();
					float *t = src;
					src = dst;
					dst = t;
				}
			}

			for (int p = threadIdx.x; p < DIL_TW * DIL_TH; p += DIL_THREADS)
			{
				const int tx = p % DIL_TW, ty = p / DIL_TW;
				const int gx = x0 + DIL_HALO + tx, gy = y0 + DIL_HALO + ty;
				if (gx >= w || gy >= h)
					continue;
				const float *v = src + ((ty + DIL_HALO) * DIL_RW + (tx + DIL_HALO)) * CH;
#pragma unroll
				for (int k = 0; k < CH; k++)
					out[((size_t)gy * w + gx) * CH + k] = quantise_u8(v[k]);
			}
		}

		// generic single round in global memory (any round count, both validity conventions): upg_dilate's path for rounds != 7
		template<int CH>
		__global__ void __launch_bounds__(256) dilate_round_kernel(const float *__restrict__ src, float *__restrict__ dst, int w, int h, int useNan)
		{
			const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
			if (i >= (size_t)w * h)
				return;
			const int x = (int)(i % w), y = (int)(i / w);
			auto valid = [&](const float *p)
			{
				if (useNan)
					return px_valid<CH>(p);
				bool nz = false;
#pragma unroll
				for (int k = 0; k < CH; k++)
					nz = nz || p[k] != 0;
				return nz;
			};
			const float *c = src + i * CH;
			float *o = dst + i * CH;
			if (valid(c))
			{
#pragma unroll
				for (int k = 0; k < CH; k++)
					o[k] = c[k];
				return;
			}
			float m[CH];
#pragma unroll
			for (int k = 0; k < CH; k++)
				m[k] = 0;
			uint32_t cnt = 0;
			for (int yy = y - 1; yy <= y + 1; yy++)
			{
				if (yy < 0 || yy >= h)
					continue;
				for (int xx = x - 1; xx <= x + 1; xx++)
				{
					if (xx < 0 || xx >= w)
						continue;
					const float *a = src + ((size_t)yy * w + xx) * CH;
					if (valid(a))
					{
#pragma unroll
						for (int k = 0; k < CH; k++)
							m[k] += a[k];
						cnt++;
					}
				}
			}
			if (cnt > 0)
			{
				const float d = (float)cnt;
#pragma unroll
				for (int k = 0; k < CH; k++)
					o[k] = m[k] / d;
			}
			// else: dst keeps what it held (the reference's ping-pong target is pre-filled with NaN / zero and holes only shrink)
		}

		inline unsigned blocks_for_n(size_t n) { return (unsigned)((n + 255) / 256); }

		__global__ void __launch_bounds__(256) quantise_kernel(const float *__restrict__ src, uint8_t *__restrict__ dst, size_t n)
		{
			const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
			if (i < n)
				dst[i] = quantise_u8(src[i]);
		}

		// K8 `texture_tail`: what the generator does to a chunk's U8 images between the bake and the PNG export
		// (sources/generator.cpp:135-139): imageConvertSpecialToGltfPbr = channels (0, roughness, metallic)
		// (imageAlgorithms.cpp:296-302) and imageConvertHeigthToNormal(height, strength) = 3x3 Sobel over the clamped U8 height,
		// normalised, mapped to [0,1] and re-quantised by truncation (:202-246, image.cpp:468-481).  `special` / `height` point at
		// image row `bandBegin`; rows [row0, row1) are produced.  Two texels per thread: one 4-byte-aligned pair of loads per row.
		struct TailArgs
		{
			const uint8_t *special; // 2 channels, may be null (no pbr wanted)
			const uint8_t *height; // 1 channel, may be null (no normal map wanted)
			uint8_t *pbr; // 3 channels, rows [row0, row1)
			uint8_t *normal; // 3 channels, rows [row0, row1)
			int w, h; // the whole image
			int bandBegin; // image row of the first row behind `special` / `height`
			int row0, row1;
			float strength; // 1 / strength_ of the reference
		};

		__global__ void __launch_bounds__(256) texture_tail_kernel(const TailArgs A)
		{
			const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
			const size_t n = (size_t)(A.row1 - A.row0) * A.w;
			if (i >= n)
				return;
			const int y = A.row0 + (int)(i / A.w);
			const int x = (int)(i % A.w);
			if (A.special)
			{
				const uint8_t *sp = A.special + ((size_t)(y - A.bandBegin) * A.w + x) * 2;
				uint8_t *o = A.pbr + i * 3;
				o[0] = 0;
				o[1] = sp[0];
				o[2] = sp[1];
			}
			if (A.height)
			{
				const int xl = x > 0 ? x - 1 : 0, xr = x + 1 < A.w ? x + 1 : A.w - 1;
				const int yt = y > 0 ? y - 1 : 0, yb = y + 1 < A.h ? y + 1 : A.h - 1;
				const uint8_t *rt = A.height + (size_t)(yt - A.bandBegin) * A.w;
				const uint8_t *rc_ = A.height + (size_t)(y - A.bandBegin) * A.w;
				const uint8_t *rb = A.height + (size_t)(yb - A.bandBegin) * A.w;
				// Image::value on U8 (image.cpp:220-236): v / 255.f; one channel: (0 + v) / 1 == v
				const float tl = rt[xl] / 255.f, tc = rt[x] / 255.f, tr = rt[xr] / 255.f;
				const float lc = rc_[xl] / 255.f, rc = rc_[xr] / 255.f;
				const float bl = rb[xl] / 255.f, bc = rb[x] / 255.f, br = rb[xr] / 255.f;
				const float dX = (tr + 2.f * rc + br) - (tl + 2.f * lc + bl);
				const float dY = (bl + 2.f * bc + br) - (tl + 2.f * tc + tr);
				V3 v = normalize(v3(-dX, -dY, A.strength));
				v = v3((v.x + 1) * 0.5f, (v.y + 1) * 0.5f, (v.z + 1) * 0.5f);
				uint8_t *o = A.normal + i * 3;
				o[0] = quantise_u8(v.x);
				o[1] = quantise_u8(v.y);
				o[2] = quantise_u8(v.z);
			}
		}

		cudaError_t texture_tail(const TailArgs &A, cudaStream_t s)
		{
			const size_t n = (size_t)(A.row1 - A.row0) * A.w;
			if (n == 0)
				return cudaSuccess;
			texture_tail_kernel<<<blocks_for_n(n), 256, 0, s>>>(A);
			return cudaGetLastError();
		}

		template<int CH>
		cudaError_t launch_dilate7(const float *img, uint32_t w, uint32_t h, uint8_t *out, cudaStream_t s)
		{
			const size_t smem = sizeof(float) * DIL_RW * DIL_RH * CH * 2;
			static bool configured[8] = {};
			int dev = 0;
			cudaGetDevice(&dev);
			if (dev < 8 && !configured[dev])
			{
				const cudaError_t e = cudaFuncSetAttribute(dilate7_quantise_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
				if (e != cudaSuccess)
					return e;
				configured[dev] = true;
			}
			else if (dev >= 8)
				cudaFuncSetAttribute(dilate7_quantise_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			const dim3 grid((w + DIL_TW - 1) / DIL_TW, (h + DIL_TH - 1) / DIL_TH);
			dilate7_quantise_kernel<CH><<<grid, DIL_THREADS, smem, s>>>(img, (int)w, (int)h, out);
			return cudaGetLastError();
		}

		cudaError_t dilate7(const float *img, uint32_t w, uint32_t h, uint32_t ch, uint8_t *out, cudaStream_t s)
		{
			switch (ch)
			{
				case 1:
					return launch_dilate7<1>(img, w, h, out, s);
				case 2:
					return launch_dilate7<2>(img, w, h, out, s);
				case 3:
					return launch_dilate7<3>(img, w, h, out, s);
				case 4:
					return launch_dilate7<4>(img, w, h, out, s);
			}
			return cudaErrorInvalidValue;
		}

		cudaError_t dilate_round(const float *src, float *dst, uint32_t w, uint32_t h, uint32_t ch, int useNan, cudaStream_t s)
		{
			const unsigned blocks = (unsigned)(((size_t)w * h + 255) / 256);
			switch (ch)
			{
				case 1:
					dilate_round_kernel<1><<<blocks, 256, 0, s>>>(src, dst, (int)w, (int)h, useNan);
					break;
				case 2:
					dilate_round_kernel<2><<<blocks, 256, 0, s>>>(src, dst, (int)w, (int)h, useNan);
					break;
				case 3:
					dilate_round_kernel<3><<<blocks, 256, 0, s>>>(src, dst, (int)w, (int)h, useNan);
					break;
				case 4:
					dilate_round_kernel<4><<<blocks, 256, 0, s>>>(src, dst, (int)w, (int)h, useNan);
					break;
				default:
					return cudaErrorInvalidValue;
			}
			return cudaGetLastError();
		}

		// stream-ordered scratch that frees itself
		struct Scratch
		{
			cudaStream_t s;
			void *ptrs[32];
			int count = 0;
			cudaError_t err = cudaSuccess;
			explicit Scratch(cudaStream_t s) : s(s) {}
			template<class T>
			T *alloc(size_t n)
			{
				void *p = nullptr;
				if (err != cudaSuccess || count >= 32)
				{
					if (err == cudaSuccess)
						err = cudaErrorMemoryAllocation;
					return nullptr;
				}
				err = device_alloc(&p, (n ? n : 1) * sizeof(T), s);
				if (err == cudaSuccess)
					ptrs[count++] = p;
				return (T *)p;
			}
			~Scratch()
			{
				for (int i = 0; i < count; i++)
					cudaFreeAsync(ptrs[i], s);
				stream_sync(s);
			}
		};

		unsigned blocks_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

		// K1 over the nine probes (+ raw elevation for non-land purposes), then K6b
		int color_tiles(Planet *p, cudaStream_t s, int purpose, uint32_t n, float *probes, float *elev9, float *elevRaw, ColorArgs A)
		{
			FieldLaunch L;
			L.out = elev9;
			L.positions = probes;
			L.count = (uint64_t)9 * n;
			L.z0 = 0;
			L.field = FD_ELEVATION;
			L.status = status_word(p);
			int rc = launch_field(p, L, s);
			if (rc)
				return rc;
			if (purpose != MP_LAND)
			{
				L.out = elevRaw;
				L.positions = probes + (size_t)8 * n * 3;
				L.count = n;
				L.field = FD_ELEVATION_RAW;
				rc = launch_field(p, L, s);
				if (rc)
					return rc;
			}
			A.probes = probes;
			A.elev9 = elev9;
			A.elevRaw = purpose != MP_LAND ? elevRaw : nullptr;
			A.cellular = p->cellularTable;
			A.n = n;
			A.purpose = purpose;
			A.status = status_word(p);
			const unsigned blocks = blocks_for(n, COLOR_BLOCK);
			switch (p->params.coloring)
			{
				case CL_DEFAULT:
					tile_color_kernel<CL_DEFAULT><<<blocks, COLOR_BLOCK, 0, s>>>(p->coloring, A);
					break;
				case CL_BARREN:
					tile_color_kernel<CL_BARREN><<<blocks, COLOR_BLOCK, 0, s>>>(p->coloring, A);
					break;
				default:
					tile_color_kernel<CL_DEBUG><<<blocks, COLOR_BLOCK, 0, s>>>(p->coloring, A);
					break;
			}
			const cudaError_t e = cudaGetLastError();
			if (e != cudaSuccess)
				return cuda_fail(e, "tile_color_kernel launch");
			return UPG_OK;
		}

		thread_local uint64_t t_lastCovered = 0;

		// `indices` = also walk the index array on the host (upg_raster, small parity inputs); the bake validates them on the device
		int check_mesh(const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, bool indices)
		{
			if (!uvs || !idx)
				return fail(UPG_ERR_ARGUMENT, "generating texture requires uvs and indices");
			if (nidx % 3 != 0)
				return fail(UPG_ERR_ARGUMENT, "generating texture requires triangles mesh");
			if (w == 0 || h == 0 || (uint64_t)w * h > 0x7fffffffull)
				return fail(UPG_ERR_ARGUMENT, "bad texture resolution %u x %u", w, h);
			for (uint32_t i = 0; indices && i < nidx; i++)
				if (idx[i] >= nverts)
					return fail(UPG_ERR_ARGUMENT, "index %u out of range (%u vertices)", idx[i], nverts);
			return UPG_OK;
		}
	}
}

using namespace upg;

extern "C"
{
	uint64_t upg_bake_last_covered_texels(void)
	{
		return t_lastCovered;
	}

	int upg_tiles(upg_planet *planet, int purpose, const float *host_positions, const float *host_normals, uint64_t n, upg_tile *host_out)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || (n && (!host_positions || !host_normals || !host_out)))
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (purpose < MP_LAND || purpose > MP_NAVIGATION)
			return fail(UPG_ERR_ARGUMENT, "bad mesh purpose %d", purpose);
		if (n == 0)
			return UPG_OK;
		if (n > 0x10000000ull)
			return fail(UPG_ERR_ARGUMENT, "too many tiles in one call");
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		StatusSlot slot; // this call's own NaN flag
		rc = slot.open(p, cs.compute);
		if (rc)
		{
			cs.close();
			return rc;
		}
		{
			Scratch mem(cs.compute);
			const uint32_t cnt = (uint32_t)n;
			float *dpos = mem.alloc<float>(n * 3);
			float *dnrm = mem.alloc<float>(n * 3);
			float *probes = mem.alloc<float>(n * 27);
			float *elev9 = mem.alloc<float>(n * 9);
			float *elevRaw = mem.alloc<float>(n);
			upg_tile *dout = mem.alloc<upg_tile>(n);
			e = mem.err;
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(dpos, host_positions, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, cs.compute);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(dnrm, host_normals, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, cs.compute);
			if (e == cudaSuccess)
			{
				tile_probes_kernel<<<blocks_for(n, BAKE_BLOCK), BAKE_BLOCK, 0, cs.compute>>>(dpos, dnrm, cnt, probes);
				e = cudaGetLastError();
			}
			if (e == cudaSuccess)
			{
				ColorArgs A;
				std::memset(&A, 0, sizeof(A));
				A.normals = dnrm;
				A.tiles = dout;
				rc = color_tiles(p, cs.compute, purpose, cnt, probes, elev9, elevRaw, A);
			}
			if (e == cudaSuccess && rc == UPG_OK)
				e = cudaMemcpyAsync(host_out, dout, sizeof(upg_tile) * n, cudaMemcpyDeviceToHost, cs.compute);
			if (e == cudaSuccess)
				e = stream_sync(cs.compute);
			if (rc == UPG_OK && e != cudaSuccess)
				rc = cuda_fail(e, "upg_tiles");
			if (rc == UPG_OK)
				rc = check_status(p, cs.compute, "upg_tiles");
		}
		cs.close();
		return rc;
	}

	int upg_raster(int device, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t *owner, float *weights, uint64_t *callbacks)
	{
		if (!owner || !weights)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		int rc = check_mesh(uvs, nverts, idx, nidx, w, h, true);
		if (rc)
			return rc;
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		rc = cs.open();
		if (rc)
			return rc;
		{
			Scratch mem(cs.compute);
			const size_t texels = (size_t)w * h;
			float *duv = mem.alloc<float>((size_t)nverts * 2);
			uint32_t *didx = mem.alloc<uint32_t>(nidx);
			uint32_t *downer = mem.alloc<uint32_t>(texels);
			uint32_t *downerOut = mem.alloc<uint32_t>(texels);
			float *dweights = mem.alloc<float>(texels * 3);
			unsigned long long *dvisits = mem.alloc<unsigned long long>(1);
			e = mem.err;
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(duv, uvs, sizeof(float) * 2 * nverts, cudaMemcpyHostToDevice, cs.compute);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(didx, idx, sizeof(uint32_t) * nidx, cudaMemcpyHostToDevice, cs.compute);
			if (e == cudaSuccess)
				e = cudaMemsetAsync(downer, 0, sizeof(uint32_t) * texels, cs.compute);
			if (e == cudaSuccess)
				e = cudaMemsetAsync(dvisits, 0, sizeof(unsigned long long), cs.compute);
			RasterMesh M{ duv, didx, nidx / 3, w, h, 0, h };
			if (e == cudaSuccess && M.triangles)
			{
				raster_owner_kernel<<<blocks_for(M.triangles, BAKE_BLOCK), BAKE_BLOCK, 0, cs.compute>>>(M, downer, dvisits);
				e = cudaGetLastError();
			}
			if (e == cudaSuccess)
			{
				raster_weights_kernel<<<blocks_for(texels, BAKE_BLOCK), BAKE_BLOCK, 0, cs.compute>>>(M, downer, downerOut, dweights);
				e = cudaGetLastError();
			}
			unsigned long long visits = 0;
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(owner, downerOut, sizeof(uint32_t) * texels, cudaMemcpyDeviceToHost, cs.compute);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(weights, dweights, sizeof(float) * 3 * texels, cudaMemcpyDeviceToHost, cs.compute);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(&visits, dvisits, sizeof(visits), cudaMemcpyDeviceToHost, cs.compute);
			if (e == cudaSuccess)
				e = stream_sync(cs.compute);
			if (callbacks)
				*callbacks = visits;
			if (e != cudaSuccess)
				rc = cuda_fail(e, "upg_raster");
		}
		cs.close();
		return rc;
	}

	int upg_dilate(int device, float *image, uint32_t w, uint32_t h, uint32_t ch, uint32_t rounds, int use_nan, uint8_t *u8_out)
	{
		if (!image)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (ch < 1 || ch > 4)
			return fail(UPG_ERR_ARGUMENT, "image dilation with more than 4 channels");
		if (w == 0 || h == 0)
			return fail(UPG_ERR_ARGUMENT, "empty image");
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		{
			Scratch mem(cs.compute);
			const size_t n = (size_t)w * h * ch;
			float *a = mem.alloc<float>(n);
			float *b = mem.alloc<float>(n);
			uint8_t *q = mem.alloc<uint8_t>(n);
			e = mem.err;
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(a, image, sizeof(float) * n, cudaMemcpyHostToDevice, cs.compute);
			// the reference's ping-pong target starts as NaN (useNan) or zero (imageAlgorithms.cpp:453-458)
			if (e == cudaSuccess)
				e = cudaMemsetAsync(b, use_nan ? 0xff : 0, sizeof(float) * n, cs.compute);
			// the fused path the bake uses reads the ORIGINAL image: run it before the round-by-round pair overwrites `a`
			const bool fused = u8_out && rounds == 7 && use_nan;
			if (e == cudaSuccess && fused)
				e = dilate7(a, w, h, ch, q, cs.compute);
			float *src = a, *dst = b;
			for (uint32_t r = 0; r < rounds && e == cudaSuccess; r++)
			{
				e = dilate_round(src, dst, w, h, ch, use_nan, cs.compute);
				float *t = src;
				src = dst;
				dst = t;
			}
			if (e == cudaSuccess && u8_out)
			{
				if (!fused)
				{
					quantise_kernel<<<blocks_for(n, 256), 256, 0, cs.compute>>>(src, q, n);
					e = cudaGetLastError();
				}
				if (e == cudaSuccess)
					e = cudaMemcpyAsync(u8_out, q, n, cudaMemcpyDeviceToHost, cs.compute);
			}
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(image, src, sizeof(float) * n, cudaMemcpyDeviceToHost, cs.compute);
			if (e == cudaSuccess)
				e = stream_sync(cs.compute);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "upg_dilate");
		}
		cs.close();
		return rc;
	}

	// the bake proper; outputs may be device or host memory (cudaMemcpyDefault).  dSpecial / dHeight are the images
	// generateTextures* returns, dPbr / dNormal what the generator exports after K8; any of the four may be null.
	}

namespace upg
{
	int bake_core(upg_planet *planet, int water, const float *dpos, const float *dnrm, const float *duv, uint32_t nverts, const uint32_t *didx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *dAlbedo, uint8_t *dSpecial, uint8_t *dHeight, uint8_t *dPbr, uint8_t *dNormal, float normalStrength, void *cuda_stream)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !dpos || !dnrm || !duv || !didx || !dAlbedo)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (dNormal && !(normalStrength > 0))
			return fail(UPG_ERR_ARGUMENT, "converting height to normal requires positive strength");
		if (nidx % 3 != 0)
			return fail(UPG_ERR_ARGUMENT, "generating texture requires triangles mesh");
		if (w == 0 || h == 0 || (uint64_t)w * h > 0x7fffffffull || row0 > row1 || row1 > h)
			return fail(UPG_ERR_ARGUMENT, "bad texture resolution %u x %u or rows [%u, %u)", w, h, row0, row1);
		// the compaction walks the image padded to whole texel tiles (TileOrder)
		if (((uint64_t)w + TEXEL_TILE_W - 1) / TEXEL_TILE_W * (((uint64_t)h + TEXEL_TILE_H - 1) / TEXEL_TILE_H) * (TEXEL_TILE_W * TEXEL_TILE_H) > 0x7fffffffull)
			return fail(UPG_ERR_ARGUMENT, "bad texture resolution %u x %u", w, h);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		cudaStream_t s = (cudaStream_t)cuda_stream;
		t_lastCovered = 0;
		if (row0 == row1)
			return UPG_OK;
		StatusSlot slot; // this call's own NaN flag
		int rc = slot.open(p, s);
		if (rc)
			return rc;
		const bool trace = trace_enabled();
		double tr[6] = { trace ? now_ms() : 0, 0, 0, 0, 0, 0 };
		{
			Scratch mem(s);
			const int ach = water ? 4 : 3;
			const int purpose = water ? MP_WATER : MP_LAND;
			// the band: the caller's rows plus what seven dilation rounds can reach (imageDilation(img, 7), textures.cpp:73-77),
			// plus the one extra row of finished heights the 3x3 Sobel of the normal map reads
			const uint32_t apron = dNormal ? 8 : 7;
			const uint32_t bandBegin = row0 > apron ? row0 - apron : 0;
			const uint32_t bandEnd = row1 + apron < h ? row1 + apron : h;
			const uint32_t bandH = bandEnd - bandBegin;
			const size_t texels = (size_t)w * bandH;
			uint32_t *owner = mem.alloc<uint32_t>(texels);
			uint32_t *list = mem.alloc<uint32_t>(texels);
			uint32_t *dcount = mem.alloc<uint32_t>(1);
			unsigned long long *dvisits = mem.alloc<unsigned long long>(2);
			uint32_t *dbad = mem.alloc<uint32_t>(1);
			float *imgAlbedo = mem.alloc<float>(texels * ach);
			float *imgSpecial = mem.alloc<float>(texels * 2);
			float *imgHeight = mem.alloc<float>(texels);
			uint8_t *q = mem.alloc<uint8_t>(texels * (ach + 3));
			uint8_t *qTail = (dPbr || dNormal) ? mem.alloc<uint8_t>((size_t)w * (row1 - row0) * 6) : nullptr;
			e = mem.err;
			if (e == cudaSuccess)
				e = cudaMemsetAsync(owner, 0, sizeof(uint32_t) * texels, s);
			if (e == cudaSuccess)
				e = cudaMemsetAsync(dvisits, 0, sizeof(unsigned long long) * 2, s);
			if (e == cudaSuccess)
				e = cudaMemsetAsync(dbad, 0, sizeof(uint32_t), s);
			// imageFill(NaN) (sources/textures.cpp:47-63): all-ones is a quiet NaN
			if (e == cudaSuccess)
				e = cudaMemsetAsync(imgAlbedo, 0xff, sizeof(float) * texels * ach, s);
			if (e == cudaSuccess)
				e = cudaMemsetAsync(imgSpecial, 0xff, sizeof(float) * texels * 2, s);
			if (e == cudaSuccess)
				e = cudaMemsetAsync(imgHeight, 0xff, sizeof(float) * texels, s);

			RasterMesh M{ duv, didx, nidx / 3, w, h, bandBegin, bandEnd, nverts, dbad };
			if (e == cudaSuccess && M.triangles)
			{
				raster_owner_kernel<<<blocks_for(M.triangles, BAKE_BLOCK), BAKE_BLOCK, 0, s>>>(M, owner, dvisits);
				e = cudaGetLastError();
			}
			// covered texels, tile by tile (TileOrder)
			uint32_t covered = 0;
			unsigned long long coveredRows = 0;
			if (e == cudaSuccess)
			{
				size_t tmpBytes = 0;
				const TileOrder order{ w, bandH, (w + TEXEL_TILE_W - 1) / TEXEL_TILE_W };
				const uint64_t slots64 = (uint64_t)order.tilesX * ((bandH + TEXEL_TILE_H - 1) / TEXEL_TILE_H) * (TEXEL_TILE_W * TEXEL_TILE_H);
				const int slots = (int)slots64; // < 2^31: checked on entry
				auto ids = thrust::make_transform_iterator(thrust::counting_iterator<uint32_t>(0), order);
				IsCovered pred{ owner };
				e = cub::DeviceSelect::If(nullptr, tmpBytes, ids, list, dcount, slots, pred, s);
				void *tmp = nullptr;
				if (e == cudaSuccess)
					tmp = mem.alloc<char>(tmpBytes);
				if (e == cudaSuccess)
					e = mem.err;
				if (e == cudaSuccess)
					e = cub::DeviceSelect::If(tmp, tmpBytes, ids, list, dcount, slots, pred, s);
				uint32_t bad = 0;
				if (e == cudaSuccess)
					e = cudaMemcpyAsync(&covered, dcount, sizeof(covered), cudaMemcpyDeviceToHost, s);
				if (e == cudaSuccess)
					e = cudaMemcpyAsync(&bad, dbad, sizeof(bad), cudaMemcpyDeviceToHost, s);
				if (e == cudaSuccess)
					e = stream_sync(s);
				if (e == cudaSuccess && bad)
				{
					// nothing downstream may touch the mesh through these indices
					rc = fail(UPG_ERR_ARGUMENT, "index %u out of range (%u vertices)", bad == 0xffffffffu ? bad : bad - 1, nverts);
					covered = 0;
				}
				if (e == cudaSuccess && covered && row0 == bandBegin && row1 == bandEnd)
					coveredRows = covered; // the band is the caller's rows (whole images): nothing to separate
				else if (e == cudaSuccess && covered)
				{
					const unsigned countBlocks = blocks_for(covered, 256) < 1184u ? blocks_for(covered, 256) : 1184u; // 148 SMs x 8 blocks
					count_in_range_kernel<<<countBlocks, 256, 0, s>>>(list, covered, (row0 - bandBegin) * w, (row1 - bandBegin) * w, dvisits + 1);
					e = cudaGetLastError();
					if (e == cudaSuccess)
						e = cudaMemcpyAsync(&coveredRows, dvisits + 1, sizeof(coveredRows), cudaMemcpyDeviceToHost, s);
				}
			}

			tr[1] = trace ? now_ms() : 0; // allocation, clears, raster, compaction (ends in the synchronisation that reads the count)
			// K6 in chunks of the list: bounds the scratch (9 probes x 16 B per texel) whatever the texture size
			const uint32_t chunk = 1u << 22;
			if (e == cudaSuccess && covered)
			{
				const uint32_t cap = covered < chunk ? covered : chunk;
				float *probes = mem.alloc<float>((size_t)cap * 27);
				float *tnormals = mem.alloc<float>((size_t)cap * 3);
				float *elev9 = mem.alloc<float>((size_t)cap * 9);
				float *elevRaw = mem.alloc<float>(cap);
				e = mem.err;
				for (uint32_t first = 0; first < covered && e == cudaSuccess && rc == UPG_OK; first += chunk)
				{
					const uint32_t n = covered - first < chunk ? covered - first : chunk;
					PrepareArgs P;
					P.mesh = M;
					P.positions = dpos;
					P.normals = dnrm;
					P.owner = owner;
					P.texels = list + first;
					P.n = n;
					P.probes = probes;
					P.tileNormals = tnormals;
					tile_prepare_kernel<<<blocks_for(n, BAKE_BLOCK), BAKE_BLOCK, 0, s>>>(P);
					e = cudaGetLastError();
					if (e != cudaSuccess)
						break;
					ColorArgs A;
					std::memset(&A, 0, sizeof(A));
					A.normals = tnormals;
					A.texels = list + first;
					A.albedo = imgAlbedo;
					A.special = imgSpecial;
					A.height = imgHeight;
					A.albedoChannels = ach;
					rc = color_tiles(p, s, purpose, n, probes, elev9, elevRaw, A);
				}
			}

			if (trace)
			{
				stream_sync(s);
				tr[2] = now_ms();
			}
			// K7 over the band (its outer 7 rows are apron unless they are the image border), then the caller's rows
			uint8_t *qAlbedo = q, *qSpecial = q + texels * ach, *qHeight = qSpecial + texels * 2;
			const size_t skip = (size_t)(row0 - bandBegin) * w, keep = (size_t)(row1 - row0) * w;
			if (e == cudaSuccess && rc == UPG_OK)
				e = dilate7(imgAlbedo, w, bandH, ach, qAlbedo, s);
			if (e == cudaSuccess && rc == UPG_OK)
				e = dilate7(imgSpecial, w, bandH, 2, qSpecial, s);
			if (e == cudaSuccess && rc == UPG_OK)
				e = dilate7(imgHeight, w, bandH, 1, qHeight, s);
			if (trace)
			{
				stream_sync(s);
				tr[3] = now_ms();
			}
			if (e == cudaSuccess && rc == UPG_OK)
				e = cudaMemcpyAsync(dAlbedo, qAlbedo + skip * ach, keep * ach, cudaMemcpyDefault, s);
			if (e == cudaSuccess && rc == UPG_OK && dSpecial)
				e = cudaMemcpyAsync(dSpecial, qSpecial + skip * 2, keep * 2, cudaMemcpyDefault, s);
			if (e == cudaSuccess && rc == UPG_OK && dHeight)
				e = cudaMemcpyAsync(dHeight, qHeight + skip, keep, cudaMemcpyDefault, s);
			if (e == cudaSuccess && rc == UPG_OK && qTail)
			{
				// K8 straight from the quantised band still in HBM
				TailArgs T;
				T.special = dPbr ? qSpecial : nullptr;
				T.height = dNormal ? qHeight : nullptr;
				T.pbr = qTail;
				T.normal = qTail + keep * 3;
				T.w = (int)w;
				T.h = (int)h;
				T.bandBegin = (int)bandBegin;
				T.row0 = (int)row0;
				T.row1 = (int)row1;
				T.strength = 1.f / normalStrength;
				e = texture_tail(T, s);
				if (e == cudaSuccess && dPbr)
					e = cudaMemcpyAsync(dPbr, T.pbr, keep * 3, cudaMemcpyDefault, s);
				if (e == cudaSuccess && dNormal)
					e = cudaMemcpyAsync(dNormal, T.normal, keep * 3, cudaMemcpyDefault, s);
			}
			if (e == cudaSuccess)
				e = stream_sync(s);
			if (trace)
				std::fprintf(stderr, "[upg]   bake core %s %ux%u rows [%u, %u): raster + compaction %.2f ms, colouring %.2f ms (%u texels), dilation %.2f ms, tail + copies out %.2f ms\n", water ? "water" : "land", w, h, row0, row1, tr[1] - tr[0], tr[2] - tr[1], covered, tr[3] - tr[2], now_ms() - tr[3]);
			t_lastCovered = coveredRows;
			if (rc == UPG_OK && e != cudaSuccess)
				rc = cuda_fail(e, "upg_bake");
			if (rc == UPG_OK)
				rc = check_status(p, s, "upg_bake");
		}
		return rc;
	}

}

extern "C"
{
	int upg_bake_device(upg_planet *planet, int water, const float *dpos, const float *dnrm, const float *duv, uint32_t nverts, const uint32_t *didx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *dAlbedo, uint8_t *dSpecial, uint8_t *dHeight, void *cuda_stream)
	{
		if (!dSpecial || !dHeight)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		return bake_core(planet, water, dpos, dnrm, duv, nverts, didx, nidx, w, h, row0, row1, dAlbedo, dSpecial, dHeight, nullptr, nullptr, 1.f, cuda_stream);
	}

	int upg_bake_export_device(upg_planet *planet, int water, const float *dpos, const float *dnrm, const float *duv, uint32_t nverts, const uint32_t *didx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *dAlbedo, uint8_t *dPbr, uint8_t *dNormal, void *cuda_stream)
	{
		if (!dPbr || !dNormal)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		return bake_core(planet, water, dpos, dnrm, duv, nverts, didx, nidx, w, h, row0, row1, dAlbedo, nullptr, nullptr, dPbr, dNormal, 1.f, cuda_stream);
	}

	int upg_texture_tail(int device, const uint8_t *special, const uint8_t *heightMap, uint32_t w, uint32_t h, float strength, uint8_t *pbr, uint8_t *normal)
	{
		if ((!special && !heightMap) || (special && !pbr) || (heightMap && !normal))
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (heightMap && !(strength > 0))
			return fail(UPG_ERR_ARGUMENT, "converting height to normal requires positive strength");
		if (w == 0 || h == 0 || (uint64_t)w * h > 0x7fffffffull)
			return fail(UPG_ERR_ARGUMENT, "bad texture resolution %u x %u", w, h);
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		{
			cudaStream_t s = cs.compute;
			Scratch mem(s);
			const size_t n = (size_t)w * h;
			uint8_t *dsp = special ? mem.alloc<uint8_t>(n * 2) : nullptr;
			uint8_t *dh = heightMap ? mem.alloc<uint8_t>(n) : nullptr;
			uint8_t *out = mem.alloc<uint8_t>(n * 6);
			e = mem.err;
			if (e == cudaSuccess && special)
				e = cudaMemcpyAsync(dsp, special, n * 2, cudaMemcpyHostToDevice, s);
			if (e == cudaSuccess && heightMap)
				e = cudaMemcpyAsync(dh, heightMap, n, cudaMemcpyHostToDevice, s);
			TailArgs T;
			T.special = dsp;
			T.height = dh;
			T.pbr = out;
			T.normal = out + n * 3;
			T.w = (int)w;
			T.h = (int)h;
			T.bandBegin = 0;
			T.row0 = 0;
			T.row1 = (int)h;
			T.strength = 1.f / strength;
			if (e == cudaSuccess)
				e = texture_tail(T, s);
			if (e == cudaSuccess && special)
				e = cudaMemcpyAsync(pbr, T.pbr, n * 3, cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess && heightMap)
				e = cudaMemcpyAsync(normal, T.normal, n * 3, cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = stream_sync(s);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "upg_texture_tail");
		}
		cs.close();
		return rc;
	}

	static int bake_rows_host(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *albedo, uint8_t *special, uint8_t *heightMap, uint8_t *pbr, uint8_t *normalMap)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !positions || !normals || !albedo)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		int rc = check_mesh(uvs, nverts, idx, nidx, w, h, false);
		if (rc)
			return rc;
		// a multi-device planet splits the rows over its devices (upg_multi.cu); bands of a few rows are not worth the fan-out
		if (p->deviceCount > 1 && row1 - row0 >= 64u * (uint32_t)p->deviceCount)
			return multi_bake(p, water, positions, normals, uvs, nverts, idx, nidx, w, h, row0, row1, albedo, special, heightMap, pbr, normalMap);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		rc = cs.open();
		if (rc)
			return rc;
		{
			cudaStream_t s = cs.compute;
			Scratch mem(s);
			const double tUp = trace_enabled() ? now_ms() : 0;
			float *dpos = mem.alloc<float>((size_t)nverts * 3);
			float *dnrm = mem.alloc<float>((size_t)nverts * 3);
			float *duv = mem.alloc<float>((size_t)nverts * 2);
			uint32_t *didx = mem.alloc<uint32_t>(nidx);
			e = mem.err;
			void *staging = nullptr;
			size_t stagingCapacity = 0;
			if (e != cudaSuccess)
				rc = cuda_fail(e, "device_alloc(mesh)");
			else
			{
				// the rasteriser needs uvs + indices first; positions / normals follow while it runs
				const HostArray arrays[4] = { { duv, uvs, sizeof(float) * 2 * nverts }, { didx, idx, sizeof(uint32_t) * nidx }, { dpos, positions, sizeof(float) * 3 * nverts }, { dnrm, normals, sizeof(float) * 3 * nverts } };
				rc = upload_host_arrays(arrays, 4, s, &staging, &stagingCapacity);
			}
			const bool trace = trace_enabled();
			const double t0 = trace ? now_ms() : 0;
			if (trace)
			{
				stream_sync(s);
				std::fprintf(stderr, "[upg] bake (host buffers): mesh upload %.2f ms (%.1f MB, waited %.2f ms after enqueueing)\n", now_ms() - tUp, (32.0 * nverts + 4.0 * nidx) * 1e-6, now_ms() - t0);
			}
			// the outputs are host memory: the core's final copies are cudaMemcpyDefault, so they land there directly
			if (rc == UPG_OK)
				rc = bake_core(planet, water, dpos, dnrm, duv, nverts, didx, nidx, w, h, row0, row1, albedo, special, heightMap, pbr, normalMap, 1.f, s);
			if (staging)
			{
				stream_sync(s);
				pinned_block_release(staging, stagingCapacity);
			}
			if (trace)
				std::fprintf(stderr, "[upg] bake (host buffers): call %.2f ms\n", now_ms() - tUp);
		}
		cs.close();
		return rc;
	}

	int upg_bake_rows(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *albedo, uint8_t *special, uint8_t *heightMap)
	{
		if (!special || !heightMap)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		return bake_rows_host(planet, water, positions, normals, uvs, nverts, idx, nidx, w, h, row0, row1, albedo, special, heightMap, nullptr, nullptr);
	}

	int upg_bake_export_rows(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *albedo, uint8_t *pbr, uint8_t *normalMap)
	{
		if (!pbr || !normalMap)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		return bake_rows_host(planet, water, positions, normals, uvs, nverts, idx, nidx, w, h, row0, row1, albedo, nullptr, nullptr, pbr, normalMap);
	}

	int upg_bake_export(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint8_t *albedo, uint8_t *pbr, uint8_t *normalMap)
	{
		return upg_bake_export_rows(planet, water, positions, normals, uvs, nverts, idx, nidx, w, h, 0, h, albedo, pbr, normalMap);
	}

	int upg_bake(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint8_t *albedo, uint8_t *special, uint8_t *heightMap)
	{
		return upg_bake_rows(planet, water, positions, normals, uvs, nverts, idx, nidx, w, h, 0, h, albedo, special, heightMap);
	}
}

namespace upg
{
	int check_bake_mesh(const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h)
	{
		return check_mesh(uvs, nverts, idx, nidx, w, h, false);
	}
	// the calling thread's upg_bake_last_covered_texels() value (the multi-device bake sums its workers' into the caller's)
	uint64_t bake_thread_covered() { return t_lastCovered; }
	void bake_set_thread_covered(uint64_t n) { t_lastCovered = n; }
}
