// upg_meshfix.cu -- K9: the mesh clean-up that follows the isosurface, on the device.
//
// Replaces, with the reference's vertex and triangle ORDER:
//   removeNonManifoldTriangles            externals/cage/sources/libcore/mesh/marchingCubes.cpp:28-96
//   meshClip(mesh, box)                   .../mesh/meshAlgorithms.cpp:926-1114  (clipAddPoint, clipTriangles, triBoxOverlap of
//                                         .../geometry/triangleBox.cpp) incl. its removeUnusedVertices / meshRemoveInvalid /
//                                         meshMergeCloseVertices({}) tail (:100-109, 1143-1248, 1291-1311, 447-499)
//   meshRemoveDisconnected                :1313-1333 (splitComponentsTriangles :690-725)
//   meshFlipNormals                       :313-326
//   the water-vertex culling              sources/meshGeneration.cpp:62-95
//
// Every one of those steps is an order-preserving filter: triangles are dropped, vertices are dropped when nothing uses them,
// indices are re-pointed at the SMALLEST id of a group (cage's UnionFind links the larger root under the smaller, :395-405).  So the
// whole chain runs in the id space of the extracted mesh with "dead" flags and one compaction at the end; only the box clip builds a
// new triangle list (a clipped triangle is replaced, in place, by its pieces; new vertices are appended in triangle order -- two
// scans).  The serial hash set of directed edges becomes a radix sort, the spatial structure of the close-vertex merge a sorted
// cell list, the two union-finds lock-free hooking with compare-and-swap.
#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/upg.h"
#include "upg_internal.h"
#include "upg_chunking_host.h"

namespace upg
{
	namespace
	{
		constexpr int MF_BLOCK = 256;
		constexpr uint32_t LOCAL_BIT = 0x80000000u; // vertex ids of points created while clipping one triangle
		constexpr int CLIP_MAX_TRIS = 24; // pieces of one triangle after six plane cuts (9-gon = 7; the fan order can add a few)
		constexpr int CLIP_MAX_NEW = 32;

		unsigned blocks_of(uint64_t n) { return (unsigned)((n + MF_BLOCK - 1) / MF_BLOCK); }

		// stream-ordered scratch, released in one go
		struct Pool
		{
			cudaStream_t s;
			std::vector<void *> ptrs;
			cudaError_t err = cudaSuccess;
			explicit Pool(cudaStream_t s) : s(s) {}
			template<class T>
			T *get(uint64_t n)
			{
				void *p = nullptr;
				if (err != cudaSuccess)
					return nullptr;
				err = device_alloc(&p, sizeof(T) * (n ? n : 1), s);
				if (err == cudaSuccess)
					ptrs.push_back(p);
				return (T *)p;
			}
			~Pool()
			{
				for (void *p : ptrs)
					if (p)
						cudaFreeAsync(p, s);
			}
		};

		__device__ __forceinline__ V3 ld3(const float *__restrict__ a, uint32_t i) { return v3(a[(size_t)i * 3 + 0], a[(size_t)i * 3 + 1], a[(size_t)i * 3 + 2]); }
		__device__ __forceinline__ void st3(float *__restrict__ a, uint32_t i, V3 v)
		{
			a[(size_t)i * 3 + 0] = v.x;
			a[(size_t)i * 3 + 1] = v.y;
			a[(size_t)i * 3 + 2] = v.z;
		}
		__device__ __forceinline__ float axis_of(V3 v, int axis) { return axis == 0 ? v.x : axis == 1 ? v.y : v.z; }
		__device__ __forceinline__ bool valid3(V3 v) { return !(isnan_(v.x) || isnan_(v.y) || isnan_(v.z)); }
		// Triangle::degenerated (shapes.cpp:119-122): area() < 1e-5, area = length(cross(b - a, c - a)) * 0.5 (geometry.h:89)
		__device__ __forceinline__ bool degenerated(V3 a, V3 b, V3 c) { return length(cross(b - a, c - a)) * 0.5f < 1e-5f; }

		// ---- removeNonManifoldTriangles -------------------------------------------------------------------------------------
		// keys only have to be EQUAL for equal directed edges: the two vertex ids are packed into 2 x `bits` bits (bits = width of the
		// largest id) so that the radix sort runs over 42 instead of 64 key bits for a 1.5 M-vertex mesh
		__global__ void __launch_bounds__(MF_BLOCK) edge_keys_kernel(const uint32_t *__restrict__ idx, uint32_t tris, uint32_t bits, uint64_t *__restrict__ keys, uint32_t *__restrict__ faces)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris)
				return;
			const uint32_t a = idx[t * 3 + 0], b = idx[t * 3 + 1], c = idx[t * 3 + 2];
			keys[t * 3 + 0] = ((uint64_t)a << bits) | b;
			keys[t * 3 + 1] = ((uint64_t)b << bits) | c;
			keys[t * 3 + 2] = ((uint64_t)c << bits) | a;
			faces[t * 3 + 0] = faces[t * 3 + 1] = faces[t * 3 + 2] = t;
		}

		// a directed edge met twice marks both faces singular (marchingCubes.cpp:66-76): after the sort, equal neighbours
		__global__ void __launch_bounds__(MF_BLOCK) duplicate_edges_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ faces, uint32_t n, uint8_t *__restrict__ dead, uint32_t *__restrict__ counter)
		{
			const uint32_t i = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (i >= n)
				return;
			const uint64_t k = keys[i];
			if ((i > 0 && keys[i - 1] == k) || (i + 1 < n && keys[i + 1] == k))
			{
				dead[faces[i]] = 1;
				atomicAdd(counter, 1u);
			}
		}

		// ---- meshClip(box) --------------------------------------------------------------------------------------------------
		// one separating-axis test of triBoxOverlap (triangleBox.cpp:21-150): p = sa * a * vA[i] + sb * b * vA[j] for two vertices
		__device__ __forceinline__ bool axis_test(float pA, float pB, float rad)
		{
			const float mn = pA < pB ? pA : pB;
			const float mx = pA < pB ? pB : pA;
			return !(mn > rad || mx < -rad);
		}
		__device__ __forceinline__ bool axis_x(float a, float b, float fa, float fb, V3 u, V3 w, V3 h) { return axis_test(a * u.y - b * u.z, a * w.y - b * w.z, fa * h.y + fb * h.z); }
		__device__ __forceinline__ bool axis_y(float a, float b, float fa, float fb, V3 u, V3 w, V3 h) { return axis_test(-a * u.x + b * u.z, -a * w.x + b * w.z, fa * h.x + fb * h.z); }
		__device__ __forceinline__ bool axis_z(float a, float b, float fa, float fb, V3 u, V3 w, V3 h) { return axis_test(a * u.x - b * u.y, a * w.x - b * w.y, fa * h.x + fb * h.y); }
		__device__ __forceinline__ bool span_outside(float x0, float x1, float x2, float h)
		{
			float mn = x0, mx = x0;
			if (x1 < mn)
				mn = x1;
			if (x1 > mx)
				mx = x1;
			if (x2 < mn)
				mn = x2;
			if (x2 > mx)
				mx = x2;
			return mn > h || mx < -h;
		}

		// intersects(Triangle, Aabb) (triangleBox.cpp:152-225): Akenine-Moller's 13-axis test in the reference's operation order
		__device__ bool triangle_box_overlap(V3 center, V3 h, V3 t0, V3 t1, V3 t2)
		{
			const V3 v0 = t0 - center, v1 = t1 - center, v2 = t2 - center;
			const V3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
			float fx = abs_(e0.x), fy = abs_(e0.y), fz = abs_(e0.z);
			if (!axis_x(e0.z, e0.y, fz, fy, v0, v2, h) || !axis_y(e0.z, e0.x, fz, fx, v0, v2, h) || !axis_z(e0.y, e0.x, fy, fx, v1, v2, h))
				return false;
			fx = abs_(e1.x), fy = abs_(e1.y), fz = abs_(e1.z);
			if (!axis_x(e1.z, e1.y, fz, fy, v0, v2, h) || !axis_y(e1.z, e1.x, fz, fx, v0, v2, h) || !axis_z(e1.y, e1.x, fy, fx, v0, v1, h))
				return false;
			fx = abs_(e2.x), fy = abs_(e2.y), fz = abs_(e2.z);
			if (!axis_x(e2.z, e2.y, fz, fy, v0, v1, h) || !axis_y(e2.z, e2.x, fz, fx, v0, v1, h) || !axis_z(e2.y, e2.x, fy, fx, v1, v2, h))
				return false;
			if (span_outside(v0.x, v1.x, v2.x, h.x) || span_outside(v0.y, v1.y, v2.y, h.y) || span_outside(v0.z, v1.z, v2.z, h.z))
				return false;
			// planeBoxOverlap(normal, v0, h)
			const V3 n = cross(e0, e1);
			V3 vmin, vmax;
			vmin.x = n.x > 0 ? -h.x - v0.x : h.x - v0.x;
			vmax.x = n.x > 0 ? h.x - v0.x : -h.x - v0.x;
			vmin.y = n.y > 0 ? -h.y - v0.y : h.y - v0.y;
			vmax.y = n.y > 0 ? h.y - v0.y : -h.y - v0.y;
			vmin.z = n.z > 0 ? -h.z - v0.z : h.z - v0.z;
			vmax.z = n.z > 0 ? h.z - v0.z : -h.z - v0.z;
			if (dot(n, vmin) > 0)
				return false;
			return dot(n, vmax) >= 0;
		}

		__device__ __forceinline__ bool point_in_box(V3 p, V3 lo, V3 hi)
		{
			// intersects(Vec3, Aabb) (geometry.cpp:307-315): inclusive
			return !(p.x < lo.x || p.x > hi.x || p.y < lo.y || p.y > hi.y || p.z < lo.z || p.z > hi.z);
		}

		// the serial clip of ONE triangle against the six box planes (meshAlgorithms.cpp:1081-1099), in a thread
		struct ClipWork
		{
			const float *positions, *normals; // the mesh before the clip
			uint32_t tris[2][CLIP_MAX_TRIS * 3];
			int count[2];
			V3 newPos[CLIP_MAX_NEW], newNrm[CLIP_MAX_NEW];
			int created;
			bool overflow;

			__device__ V3 pos(uint32_t id) const { return (id & LOCAL_BIT) ? newPos[id & ~LOCAL_BIT] : ld3(positions, id); }
			__device__ V3 nrm(uint32_t id) const { return (id & LOCAL_BIT) ? newNrm[id & ~LOCAL_BIT] : ld3(normals, id); }

			// clipAddPoint (meshAlgorithms.cpp:928-949)
			__device__ uint32_t add_point(uint32_t ai, uint32_t bi, int axis, float value)
			{
				const V3 a = pos(ai), b = pos(bi);
				const float aa = axis_of(a, axis), ba = axis_of(b, axis);
				if (abs_(ba - aa) < 1e-5f)
					return ai; // the edge is very short or degenerated
				const float pu = (value - aa) / (ba - aa);
				if (pu < 1e-5f)
					return ai;
				if (pu > (float)(1 - 1e-5))
					return bi;
				if (created >= CLIP_MAX_NEW)
				{
					overflow = true;
					return ai;
				}
				newPos[created] = a + (b - a) * pu; // interpolate(a, b, pu)
				const V3 na = nrm(ai), nb = nrm(bi);
				newNrm[created] = normalize(na + (nb - na) * pu);
				return LOCAL_BIT | (uint32_t)created++;
			}

			__device__ void push(int list, uint32_t a, uint32_t b, uint32_t c)
			{
				if (count[list] >= CLIP_MAX_TRIS)
				{
					overflow = true;
					return;
				}
				uint32_t *o = tris[list] + count[list] * 3;
				o[0] = a;
				o[1] = b;
				o[2] = c;
				count[list]++;
			}

			// clipTriangles (meshAlgorithms.cpp:951-1054) from list `in` into list `in ^ 1`
			__device__ void clip_plane(int in, int axis, float value, bool side)
			{
				const int out = in ^ 1;
				count[out] = 0;
				for (int t = 0; t < count[in]; t++)
				{
					uint32_t i0 = tris[in][t * 3 + 0], i1 = tris[in][t * 3 + 1], i2 = tris[in][t * 3 + 2];
					bool as = axis_of(pos(i0), axis) > value;
					bool bs = axis_of(pos(i1), axis) > value;
					bool cs = axis_of(pos(i2), axis) > value;
					const int m = (int)as + (int)bs + (int)cs;
					if ((m == 0 && side) || (m == 3 && !side))
					{
						push(out, i0, i1, i2);
						continue;
					}
					if ((m == 0 && !side) || (m == 3 && side))
						continue;
					while (as || !bs)
					{
						// turnLeft: (a, b, c) -> (b, c, a)
						const uint32_t ti = i0;
						i0 = i1;
						i1 = i2;
						i2 = ti;
						const bool ts = as;
						as = bs;
						bs = cs;
						cs = ts;
					}
					const uint32_t ab = add_point(i0, i1, axis, value);
					if (m == 1)
					{
						const uint32_t bc = add_point(i1, i2, axis, value);
						if (side)
						{
							push(out, i0, ab, bc);
							push(out, i0, bc, i2);
						}
						else
							push(out, ab, i1, bc);
					}
					else
					{
						const uint32_t ac = add_point(i0, i2, axis, value);
						if (side)
							push(out, i0, ab, ac);
						else
						{
							push(out, ab, i1, ac);
							push(out, ac, i1, i2);
						}
					}
				}
			}

			// returns the list holding the result
			__device__ int run(uint32_t a, uint32_t b, uint32_t c, V3 lo, V3 hi)
			{
				created = 0;
				overflow = false;
				count[0] = 0;
				push(0, a, b, c);
				int cur = 0;
				for (int axis = 0; axis < 3; axis++)
					for (int side = 0; side < 2; side++)
					{
						clip_plane(cur, axis, axis_of(side ? hi : lo, axis), side != 0);
						cur ^= 1;
					}
				return cur;
			}
		};

		enum ClipClass : uint8_t
		{
			CC_DROP = 0,
			CC_KEEP = 1,
			CC_CUT = 2
		};

		__device__ ClipClass clip_classify(V3 a, V3 b, V3 c, V3 lo, V3 hi)
		{
			if (point_in_box(a, lo, hi) && point_in_box(b, lo, hi) && point_in_box(c, lo, hi))
				return CC_KEEP;
			// Aabb::center() = (a + b) * 0.5, size() * 0.5 (geometry.h:192)
			const V3 ce = (lo + hi) * 0.5f, hs = (hi - lo) * 0.5f;
			if (!triangle_box_overlap(ce, hs, a, b, c))
				return CC_DROP;
			return CC_CUT;
		}

		__global__ void __launch_bounds__(128) clip_count_kernel(const float *__restrict__ positions, const float *__restrict__ normals, const uint32_t *__restrict__ idx, const uint8_t *__restrict__ dead, uint32_t tris, V3 lo, V3 hi, uint32_t *__restrict__ triOut, uint32_t *__restrict__ newOut, uint32_t *__restrict__ status)
		{
			const uint32_t t = blockIdx.x * 128 + threadIdx.x;
			if (t >= tris)
				return;
			uint32_t nt = 0, nv = 0;
			if (!dead[t])
			{
				const uint32_t a = idx[t * 3 + 0], b = idx[t * 3 + 1], c = idx[t * 3 + 2];
				const ClipClass k = clip_classify(ld3(positions, a), ld3(positions, b), ld3(positions, c), lo, hi);
				if (k == CC_KEEP)
					nt = 1;
				else if (k == CC_CUT)
				{
					ClipWork w;
					w.positions = positions;
					w.normals = normals;
					const int r = w.run(a, b, c, lo, hi);
					nt = (uint32_t)w.count[r];
					nv = (uint32_t)w.created;
					if (w.overflow)
						atomicOr(status, 1u);
				}
			}
			triOut[t] = nt;
			newOut[t] = nv;
		}

		__global__ void __launch_bounds__(128) clip_emit_kernel(const float *__restrict__ positions, const float *__restrict__ normals, const uint32_t *__restrict__ idx, const uint8_t *__restrict__ dead, uint32_t tris, V3 lo, V3 hi, const uint32_t *__restrict__ triOff, const uint32_t *__restrict__ newOff, uint32_t oldVerts, float *__restrict__ positionsOut, float *__restrict__ normalsOut, uint32_t *__restrict__ idxOut)
		{
			const uint32_t t = blockIdx.x * 128 + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			const uint32_t a = idx[t * 3 + 0], b = idx[t * 3 + 1], c = idx[t * 3 + 2];
			const ClipClass k = clip_classify(ld3(positions, a), ld3(positions, b), ld3(positions, c), lo, hi);
			uint32_t *o = idxOut + (size_t)triOff[t] * 3;
			if (k == CC_KEEP)
			{
				o[0] = a;
				o[1] = b;
				o[2] = c;
			}
			else if (k == CC_CUT)
			{
				ClipWork w;
				w.positions = positions;
				w.normals = normals;
				const int r = w.run(a, b, c, lo, hi);
				const uint32_t base = oldVerts + newOff[t];
				for (int i = 0; i < w.created; i++)
				{
					st3(positionsOut, base + i, w.newPos[i]);
					st3(normalsOut, base + i, w.newNrm[i]);
				}
				for (int i = 0; i < w.count[r] * 3; i++)
				{
					const uint32_t id = w.tris[r][i];
					o[i] = (id & LOCAL_BIT) ? base + (id & ~LOCAL_BIT) : id;
				}
			}
		}

		// ---- meshRemoveInvalid (meshAlgorithms.cpp:1143-1248): faces with an invalid vertex, then degenerated faces ---------------
		__global__ void __launch_bounds__(MF_BLOCK) validate_kernel(const float *__restrict__ positions, const float *__restrict__ normals, const uint32_t *__restrict__ idx, uint32_t tris, uint8_t *__restrict__ dead)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			bool bad = false;
			V3 p[3];
#pragma unroll
			for (int k = 0; k < 3; k++)
			{
				const uint32_t v = idx[t * 3 + k];
				p[k] = ld3(positions, v);
				const V3 n = ld3(normals, v);
				bad |= !valid3(p[k]) || !valid3(n) || abs_(length_squared(n) - 1) > 1e-3f;
			}
			if (bad || degenerated(p[0], p[1], p[2]))
				dead[t] = 1;
		}

		__global__ void __launch_bounds__(MF_BLOCK) degenerate_kernel(const float *__restrict__ positions, const uint32_t *__restrict__ idx, uint32_t tris, uint8_t *__restrict__ dead)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			if (degenerated(ld3(positions, idx[t * 3 + 0]), ld3(positions, idx[t * 3 + 1]), ld3(positions, idx[t * 3 + 2])))
				dead[t] = 1;
		}

		__global__ void __launch_bounds__(MF_BLOCK) mark_used_kernel(const uint32_t *__restrict__ idx, const uint8_t *__restrict__ dead, uint32_t tris, uint32_t *__restrict__ used)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			used[idx[t * 3 + 0]] = 1;
			used[idx[t * 3 + 1]] = 1;
			used[idx[t * 3 + 2]] = 1;
		}

		// ---- lock-free union-find; the root of a set is always its smallest member, like cage's UnionFind (:374-405) ----------------
		__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x)
		{
			volatile uint32_t *p = parent;
			while (true)
			{
				const uint32_t px = p[x];
				if (px == x)
					return x;
				const uint32_t ppx = p[px];
				if (ppx != px)
					p[x] = ppx; // path halving: x is not a root and never becomes one again
				x = px;
			}
		}
		__device__ __forceinline__ bool uf_unite(uint32_t *parent, uint32_t a, uint32_t b)
		{
			while (true)
			{
				a = uf_find(parent, a);
				b = uf_find(parent, b);
				if (a == b)
					return false;
				if (a < b)
				{
					const uint32_t t = a;
					a = b;
					b = t;
				}
				if (atomicCAS(&parent[a], a, b) == a)
					return true;
			}
		}
		__global__ void __launch_bounds__(MF_BLOCK) iota_kernel(uint32_t *__restrict__ a, uint32_t n)
		{
			const uint32_t i = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (i < n)
				a[i] = i;
		}
		// parent[i] = root of i, for every i.  The walk is READ-ONLY on purpose: uf_find()'s path halving stores an ancestor it read a
		// moment ago into p[x], and when that store (by a thread walking THROUGH i) lands after thread i has written its final root, i
		// is left pointing at an inner node.  The next kernels take parent[] for the representative, so the faces whose first vertex is i
		// were counted into the wrong component and dropped by meshRemoveDisconnected -- one vertex and its six triangles lost in about
		// one mesh of a hundred when three clean-ups ran concurrently (profiles/r02_flake.md).  Here location i is written by thread i
		// alone, and every value a walker can read (an old ancestor or the final root) leads to the same root.
		__global__ void __launch_bounds__(MF_BLOCK) flatten_kernel(uint32_t *parent, uint32_t n)
		{
			const uint32_t i = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (i >= n)
				return;
#if defined(UPG_FLATTEN_HALVING) // the round-1 form, kept for the A/B run that demonstrates the race (profiles/r02_flake.md)
			parent[i] = uf_find(parent, i);
			return;
#endif
			const volatile uint32_t *p = parent;
			uint32_t x = i;
			for (uint32_t px = p[x]; px != x; px = p[x])
				x = px;
			parent[i] = x;
		}

		// ---- meshMergeCloseVertices({}) (meshAlgorithms.cpp:447-499): threshold 1e-5 (meshAlgorithms.h:23) ----------------------------
		constexpr float MERGE_CELLS_PER_UNIT = 64.0f; // cell size 1/64 >> threshold + a float ulp at |p| < 16384
		constexpr int32_t MERGE_CELL_BIAS = 1 << 20;

		__device__ __forceinline__ uint32_t merge_cell(float v)
		{
			float c = floorf(v * MERGE_CELLS_PER_UNIT);
			c = c < -1048575.0f ? -1048575.0f : c > 1048575.0f ? 1048575.0f : c; // clamped consistently for items and queries
			return (uint32_t)((int32_t)c + MERGE_CELL_BIAS);
		}
		__device__ __forceinline__ uint64_t merge_key(uint32_t cx, uint32_t cy, uint32_t cz) { return ((uint64_t)cx << 42) | ((uint64_t)cy << 21) | cz; }

		__global__ void __launch_bounds__(MF_BLOCK) merge_keys_kernel(const float *__restrict__ positions, const uint32_t *__restrict__ used, uint32_t verts, uint64_t *__restrict__ keys, uint32_t *__restrict__ ids)
		{
			const uint32_t v = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (v >= verts)
				return;
			ids[v] = v;
			if (!used[v])
			{
				keys[v] = ~0ull; // sorts behind every real cell
				return;
			}
			const V3 p = ld3(positions, v);
			keys[v] = merge_key(merge_cell(p.x), merge_cell(p.y), merge_cell(p.z));
		}

		__global__ void __launch_bounds__(MF_BLOCK) merge_close_kernel(const float *__restrict__ positions, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ ids, uint32_t n, uint32_t *__restrict__ parent, uint32_t *__restrict__ merges)
		{
			const uint32_t s = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (s >= n || keys[s] == ~0ull)
				return;
			const uint32_t i = ids[s];
			const V3 p = ld3(positions, i);
			const float t = 1e-5f;
			// q->intersection(Aabb(ps[i] - threshold, ps[i] + threshold)): items whose point lies in the closed box
			const V3 lo = p - t, hi = p + t;
			const uint32_t x0 = merge_cell(lo.x), x1 = merge_cell(hi.x), y0 = merge_cell(lo.y), y1 = merge_cell(hi.y), z0 = merge_cell(lo.z), z1 = merge_cell(hi.z);
			for (uint32_t cx = x0; cx <= x1; cx++)
				for (uint32_t cy = y0; cy <= y1; cy++)
					for (uint32_t cz = z0; cz <= z1; cz++)
					{
						const uint64_t key = merge_key(cx, cy, cz);
						uint32_t a = 0, b = n;
						while (a < b)
						{
							const uint32_t m = a + (b - a) / 2;
							if (keys[m] < key)
								a = m + 1;
							else
								b = m;
						}
						for (; a < n && keys[a] == key; a++)
						{
							const uint32_t j = ids[a];
							if (j == i)
								continue;
							const V3 q = ld3(positions, j);
							if (q.x < lo.x || q.x > hi.x || q.y < lo.y || q.y > hi.y || q.z < lo.z || q.z > hi.z)
								continue;
							if (uf_unite(parent, i, j))
								atomicAdd(merges, 1u);
						}
					}
		}

		__global__ void __launch_bounds__(MF_BLOCK) remap_kernel(uint32_t *__restrict__ idx, const uint8_t *__restrict__ dead, uint32_t tris, const uint32_t *__restrict__ rep)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			idx[t * 3 + 0] = rep[idx[t * 3 + 0]];
			idx[t * 3 + 1] = rep[idx[t * 3 + 1]];
			idx[t * 3 + 2] = rep[idx[t * 3 + 2]];
		}

		// ---- meshRemoveDisconnected (meshAlgorithms.cpp:690-725, 1313-1333) ------------------------------------------------------
		__global__ void __launch_bounds__(MF_BLOCK) components_kernel(const uint32_t *__restrict__ idx, const uint8_t *__restrict__ dead, uint32_t tris, uint32_t *__restrict__ parent)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			const uint32_t a = idx[t * 3 + 0], b = idx[t * 3 + 1], c = idx[t * 3 + 2];
			uf_unite(parent, a, b);
			uf_unite(parent, b, c);
		}
		__global__ void __launch_bounds__(MF_BLOCK) component_faces_kernel(const uint32_t *__restrict__ idx, const uint8_t *__restrict__ dead, uint32_t tris, const uint32_t *__restrict__ rep, uint32_t *__restrict__ faces)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			const bool live = t < tris && !dead[t];
			const uint32_t r = live ? rep[idx[t * 3 + 0]] : 0xffffffffu;
			// a planet is one big component: nearly every lane of a warp counts into the same word, so lanes with equal
			// representatives elect one of them to add their number (one atomic per distinct component per warp)
			const uint32_t peers = __match_any_sync(0xffffffffu, r);
			if (live && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1))
				atomicAdd(&faces[r], (uint32_t)__popc(peers));
		}
		// the component with the most faces; the FIRST one on ties, and components are ordered by their smallest vertex (:705-706, 1321-1330)
		__global__ void __launch_bounds__(MF_BLOCK) largest_component_kernel(const uint32_t *__restrict__ faces, uint32_t verts, unsigned long long *__restrict__ best)
		{
			const uint32_t v = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (v >= verts || faces[v] == 0)
				return;
			atomicMax(best, ((unsigned long long)faces[v] << 32) | (0xffffffffu - v));
		}
		__global__ void __launch_bounds__(MF_BLOCK) keep_component_kernel(const uint32_t *__restrict__ idx, uint32_t tris, const uint32_t *__restrict__ rep, const unsigned long long *__restrict__ best, uint8_t *__restrict__ dead)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			const uint32_t keep = 0xffffffffu - (uint32_t)(*best & 0xffffffffu);
			if (rep[idx[t * 3 + 0]] != keep)
				dead[t] = 1;
		}

		// ---- water-vertex culling (sources/meshGeneration.cpp:62-95) -----------------------------------------------------------------
		__global__ void __launch_bounds__(MF_BLOCK) cull_seed_kernel(const float *__restrict__ elevation, uint32_t verts, uint32_t *__restrict__ valid)
		{
			const uint32_t v = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (v < verts)
				valid[v] = elevation[v] < 0.1f ? 1u : 0u;
		}
		__global__ void __launch_bounds__(MF_BLOCK) cull_expand_kernel(const uint32_t *__restrict__ idx, uint32_t tris, const uint32_t *__restrict__ valid, uint32_t *__restrict__ valid2)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris)
				return;
			const uint32_t a = idx[t * 3 + 0], b = idx[t * 3 + 1], c = idx[t * 3 + 2];
			if (valid[a] | valid[b] | valid[c])
				valid2[a] = valid2[b] = valid2[c] = 1;
		}
		__global__ void __launch_bounds__(MF_BLOCK) cull_faces_kernel(const uint32_t *__restrict__ idx, uint32_t tris, const uint32_t *__restrict__ valid, uint8_t *__restrict__ dead)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris)
				return;
			dead[t] = (valid[idx[t * 3 + 0]] & valid[idx[t * 3 + 1]] & valid[idx[t * 3 + 2]]) ? 0 : 1;
		}

		// ---- the one compaction ----------------------------------------------------------------------------------------------------
		__global__ void __launch_bounds__(MF_BLOCK) alive_kernel(const uint8_t *__restrict__ dead, uint32_t tris, uint32_t *__restrict__ alive)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t < tris)
				alive[t] = dead[t] ? 0u : 1u;
		}
		__global__ void __launch_bounds__(MF_BLOCK) compact_vertices_kernel(const float *__restrict__ positions, const float *__restrict__ normals, const uint32_t *__restrict__ used, const uint32_t *__restrict__ newId, uint32_t verts, float normalSign, float *__restrict__ positionsOut, float *__restrict__ normalsOut)
		{
			const uint32_t v = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (v >= verts || !used[v])
				return;
			const uint32_t o = newId[v];
			st3(positionsOut, o, ld3(positions, v));
			st3(normalsOut, o, ld3(normals, v) * normalSign); // meshFlipNormals: n *= -1
		}
		__global__ void __launch_bounds__(MF_BLOCK) compact_faces_kernel(const uint32_t *__restrict__ idx, const uint8_t *__restrict__ dead, const uint32_t *__restrict__ faceOff, uint32_t tris, const uint32_t *__restrict__ newId, int flip, uint32_t *__restrict__ idxOut)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris || dead[t])
				return;
			uint32_t *o = idxOut + (size_t)faceOff[t] * 3;
			o[0] = newId[idx[t * 3 + 0]];
			o[flip ? 2 : 1] = newId[idx[t * 3 + 1]]; // meshFlipNormals swaps the 2nd and 3rd index of every triangle
			o[flip ? 1 : 2] = newId[idx[t * 3 + 2]];
		}

		cudaError_t scan_u32(Pool &pool, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total)
		{
			*total = 0;
			if (n == 0)
				return cudaSuccess;
			size_t bytes = 0;
			cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, pool.s);
			if (e != cudaSuccess)
				return e;
			void *tmp = pool.get<char>(bytes);
			if (pool.err != cudaSuccess)
				return pool.err;
			e = cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, pool.s);
			uint32_t a = 0, b = 0;
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(&a, in + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, pool.s);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(&b, out + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, pool.s);
			if (e == cudaSuccess)
				e = stream_sync(pool.s);
			*total = a + b;
			return e;
		}

		template<class K, class V>
		cudaError_t sort_pairs(Pool &pool, const K *kin, K *kout, const V *vin, V *vout, uint32_t n, int endBit)
		{
			if (n == 0)
				return cudaSuccess;
			size_t bytes = 0;
			cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, endBit, pool.s);
			if (e != cudaSuccess)
				return e;
			void *tmp = pool.get<char>(bytes);
			if (pool.err != cudaSuccess)
				return pool.err;
			return cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, (int)n, 0, endBit, pool.s);
		}

		// the working mesh: arrays live in `pool` until the final compaction hands its outputs over
		struct Work
		{
			float *positions = nullptr, *normals = nullptr;
			uint32_t *idx = nullptr;
			uint8_t *dead = nullptr;
			uint32_t verts = 0, tris = 0;
		};

#define MF_TRY(x) \
	do \
	{ \
		if (e == cudaSuccess) \
			e = (x); \
	} while (0)
#define MF_LAUNCHED() MF_TRY(cudaGetLastError())

		// used flags + exclusive scan of them + compaction of vertices and faces into freshly allocated outputs
		cudaError_t compact(Pool &pool, const Work &w, bool flip, float **positionsOut, float **normalsOut, uint32_t **idxOut, uint32_t *vertsOut, uint32_t *trisOut)
		{
			cudaStream_t s = pool.s;
			cudaError_t e = cudaSuccess;
			*positionsOut = *normalsOut = nullptr;
			*idxOut = nullptr;
			*vertsOut = *trisOut = 0;
			if (w.tris == 0 || w.verts == 0)
				return e;
			uint32_t *used = pool.get<uint32_t>(w.verts), *newId = pool.get<uint32_t>(w.verts);
			uint32_t *alive = pool.get<uint32_t>(w.tris), *faceOff = pool.get<uint32_t>(w.tris);
			MF_TRY(pool.err);
			MF_TRY(cudaMemsetAsync(used, 0, sizeof(uint32_t) * w.verts, s));
			if (e == cudaSuccess)
			{
				mark_used_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.idx, w.dead, w.tris, used);
				alive_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.dead, w.tris, alive);
				MF_LAUNCHED();
			}
			uint32_t nv = 0, nt = 0;
			MF_TRY(scan_u32(pool, used, newId, w.verts, &nv));
			MF_TRY(scan_u32(pool, alive, faceOff, w.tris, &nt));
			if (e != cudaSuccess || nt == 0)
				return e; // "if all triangles were removed ... poly->clear()"
			float *po = nullptr, *no = nullptr;
			uint32_t *io = nullptr;
			MF_TRY(device_alloc(&po, sizeof(float) * 3 * nv, s));
			MF_TRY(device_alloc(&no, sizeof(float) * 3 * nv, s));
			MF_TRY(device_alloc(&io, sizeof(uint32_t) * 3 * nt, s));
			if (e == cudaSuccess)
			{
				compact_vertices_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(w.positions, w.normals, used, newId, w.verts, flip ? -1.0f : 1.0f, po, no);
				compact_faces_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.idx, w.dead, faceOff, w.tris, newId, flip ? 1 : 0, io);
				MF_LAUNCHED();
			}
			if (e != cudaSuccess)
			{
				if (po)
					cudaFreeAsync(po, s);
				if (no)
					cudaFreeAsync(no, s);
				if (io)
					cudaFreeAsync(io, s);
				return e;
			}
			*positionsOut = po;
			*normalsOut = no;
			*idxOut = io;
			*vertsOut = nv;
			*trisOut = nt;
			return e;
		}
	}

	// `positions`, `normals` (verts x 3) and `indices` (tris x 3) are device arrays and stay untouched; the result replaces `out`
	// (its arrays come from the stream-ordered pool: dmc_release frees them).
	int mesh_clean(Planet *planet, const float *positions, const float *normals, uint32_t verts, const uint32_t *indices, uint32_t tris, float boxSize, unsigned stages, cudaStream_t s, DmcResult &out, const float *clipBox6)
	{
		out = DmcResult();
		if (verts >= LOCAL_BIT || (uint64_t)tris * 3 >= 0xffffffffull)
			return fail(UPG_ERR_ARGUMENT, "mesh too large for the clean-up (%u vertices, %u triangles)", verts, tris);
		if ((stages & UPG_MESH_WATER_CULL) && !planet)
			return fail(UPG_ERR_ARGUMENT, "the water culling needs a planet");
		cudaError_t e = cudaSuccess;
		int rc = UPG_OK;
		Pool pool(s);
		Work w;
		w.verts = verts;
		w.tris = tris;
		if (tris == 0)
			return UPG_OK; // meshClip / meshRemoveDisconnected return an empty mesh untouched; nothing to flip
		uint32_t *counters = pool.get<uint32_t>(4); // 0 singular edges, 1 clip overflow, 2 merges
		unsigned long long *best = pool.get<unsigned long long>(1);
		w.dead = pool.get<uint8_t>(tris);
		MF_TRY(pool.err);
		MF_TRY(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 4, s));
		MF_TRY(cudaMemsetAsync(best, 0, sizeof(unsigned long long), s));
		MF_TRY(cudaMemsetAsync(w.dead, 0, tris, s));

		if (stages & (UPG_MESH_MAKEMESH_TAIL | MESH_STAGE_CLIP_BOX))
		{
			// removeNonManifoldTriangles (makeMesh only; meshClip on its own -- the chunking -- has no such step)
			if (stages & UPG_MESH_MAKEMESH_TAIL)
			{
				const uint32_t n = tris * 3;
				uint32_t idBits = 1;
				while (idBits < 32 && (verts - 1) >> idBits)
					idBits++;
				uint64_t *k0 = pool.get<uint64_t>(n), *k1 = pool.get<uint64_t>(n);
				uint32_t *f0 = pool.get<uint32_t>(n), *f1 = pool.get<uint32_t>(n);
				MF_TRY(pool.err);
				if (e == cudaSuccess)
				{
					edge_keys_kernel<<<blocks_of(tris), MF_BLOCK, 0, s>>>(indices, tris, idBits, k0, f0);
					MF_LAUNCHED();
				}
				MF_TRY(sort_pairs(pool, k0, k1, f0, f1, n, (int)(2 * idBits)));
				if (e == cudaSuccess)
				{
					duplicate_edges_kernel<<<blocks_of(n), MF_BLOCK, 0, s>>>(k1, f1, n, w.dead, counters);
					MF_LAUNCHED();
				}
			}
			// meshClip(box): a new triangle list and appended vertices
			const bool ownBox = (stages & MESH_STAGE_CLIP_BOX) && clipBox6;
			const V3 lo = ownBox ? v3(clipBox6[0], clipBox6[1], clipBox6[2]) : v3(boxSize * -0.5f, boxSize * -0.5f, boxSize * -0.5f);
			const V3 hi = ownBox ? v3(clipBox6[3], clipBox6[4], clipBox6[5]) : v3(boxSize * 0.5f, boxSize * 0.5f, boxSize * 0.5f);
			uint32_t *triOut = pool.get<uint32_t>(tris), *newOut = pool.get<uint32_t>(tris), *triOff = pool.get<uint32_t>(tris), *newOff = pool.get<uint32_t>(tris);
			MF_TRY(pool.err);
			if (e == cudaSuccess)
			{
				clip_count_kernel<<<(tris + 127) / 128, 128, 0, s>>>(positions, normals, indices, w.dead, tris, lo, hi, triOut, newOut, counters + 1);
				MF_LAUNCHED();
			}
			uint32_t tris2 = 0, created = 0;
			MF_TRY(scan_u32(pool, triOut, triOff, tris, &tris2));
			MF_TRY(scan_u32(pool, newOut, newOff, tris, &created));
			uint32_t hostCounters[4] = { 0, 0, 0, 0 };
			MF_TRY(cudaMemcpyAsync(hostCounters, counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, s));
			MF_TRY(stream_sync(s));
			if (e == cudaSuccess && hostCounters[1])
				rc = fail(UPG_ERR_INVALID_VALUE, "clipping one triangle against the box produced more than %d pieces", CLIP_MAX_TRIS);
			if (e == cudaSuccess && rc == UPG_OK && (uint64_t)verts + created >= LOCAL_BIT)
				rc = fail(UPG_ERR_ARGUMENT, "mesh too large for the clean-up");
			if (e == cudaSuccess && rc == UPG_OK)
			{
				w.verts = verts + created;
				w.tris = tris2;
				w.positions = pool.get<float>((uint64_t)w.verts * 3);
				w.normals = pool.get<float>((uint64_t)w.verts * 3);
				w.idx = pool.get<uint32_t>((uint64_t)tris2 * 3);
				uint8_t *dead2 = pool.get<uint8_t>(tris2);
				MF_TRY(pool.err);
				MF_TRY(cudaMemcpyAsync(w.positions, positions, sizeof(float) * 3 * verts, cudaMemcpyDeviceToDevice, s));
				MF_TRY(cudaMemcpyAsync(w.normals, normals, sizeof(float) * 3 * verts, cudaMemcpyDeviceToDevice, s));
				MF_TRY(cudaMemsetAsync(dead2, 0, tris2 ? tris2 : 1, s));
				if (e == cudaSuccess)
				{
					clip_emit_kernel<<<(tris + 127) / 128, 128, 0, s>>>(positions, normals, indices, w.dead, tris, lo, hi, triOff, newOff, verts, w.positions, w.normals, w.idx);
					MF_LAUNCHED();
				}
				w.dead = dead2;
			}
			if (e == cudaSuccess && rc == UPG_OK && w.tris)
			{
				// removeUnusedVertices + meshRemoveInvalid, then meshMergeCloseVertices({}) + meshRemoveInvalid
				validate_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.positions, w.normals, w.idx, w.tris, w.dead);
				MF_LAUNCHED();
				uint32_t *used = pool.get<uint32_t>(w.verts), *parent = pool.get<uint32_t>(w.verts);
				uint64_t *k0 = pool.get<uint64_t>(w.verts), *k1 = pool.get<uint64_t>(w.verts);
				uint32_t *i0 = pool.get<uint32_t>(w.verts), *i1 = pool.get<uint32_t>(w.verts);
				MF_TRY(pool.err);
				MF_TRY(cudaMemsetAsync(used, 0, sizeof(uint32_t) * w.verts, s));
				if (e == cudaSuccess)
				{
					mark_used_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.idx, w.dead, w.tris, used);
					merge_keys_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(w.positions, used, w.verts, k0, i0);
					iota_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(parent, w.verts);
					MF_LAUNCHED();
				}
				MF_TRY(sort_pairs(pool, k0, k1, i0, i1, w.verts, 64));
				if (e == cudaSuccess)
				{
					merge_close_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(w.positions, k1, i1, w.verts, parent, counters + 2);
					MF_LAUNCHED();
				}
				MF_TRY(cudaMemcpyAsync(hostCounters, counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, s));
				MF_TRY(stream_sync(s));
				if (e == cudaSuccess && hostCounters[2])
				{
					flatten_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(parent, w.verts);
					remap_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.idx, w.dead, w.tris, parent);
					degenerate_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.positions, w.idx, w.tris, w.dead);
					MF_LAUNCHED();
				}
			}
		}
		else
		{
			// later stages only: work on copies so that the inputs stay untouched
			w.positions = pool.get<float>((uint64_t)verts * 3);
			w.normals = pool.get<float>((uint64_t)verts * 3);
			w.idx = pool.get<uint32_t>((uint64_t)tris * 3);
			MF_TRY(pool.err);
			MF_TRY(cudaMemcpyAsync(w.positions, positions, sizeof(float) * 3 * verts, cudaMemcpyDeviceToDevice, s));
			MF_TRY(cudaMemcpyAsync(w.normals, normals, sizeof(float) * 3 * verts, cudaMemcpyDeviceToDevice, s));
			MF_TRY(cudaMemcpyAsync(w.idx, indices, sizeof(uint32_t) * 3 * tris, cudaMemcpyDeviceToDevice, s));
		}

		if (e == cudaSuccess && rc == UPG_OK && (stages & UPG_MESH_REMOVE_DISCONNECTED) && w.tris)
		{
			uint32_t *parent = pool.get<uint32_t>(w.verts), *faces = pool.get<uint32_t>(w.verts);
			MF_TRY(pool.err);
			MF_TRY(cudaMemsetAsync(faces, 0, sizeof(uint32_t) * w.verts, s));
			if (e == cudaSuccess)
			{
				iota_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(parent, w.verts);
				components_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.idx, w.dead, w.tris, parent);
				flatten_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(parent, w.verts);
				component_faces_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.idx, w.dead, w.tris, parent, faces);
				largest_component_kernel<<<blocks_of(w.verts), MF_BLOCK, 0, s>>>(faces, w.verts, best);
				keep_component_kernel<<<blocks_of(w.tris), MF_BLOCK, 0, s>>>(w.idx, w.tris, parent, best, w.dead);
				MF_LAUNCHED();
			}
		}

		const bool flip = (stages & UPG_MESH_FLIP_NORMALS) != 0;
		float *po = nullptr, *no = nullptr;
		uint32_t *io = nullptr;
		uint32_t nv = 0, nt = 0;
		if (e == cudaSuccess && rc == UPG_OK)
			e = compact(pool, w, flip, &po, &no, &io, &nv, &nt);

		if (e == cudaSuccess && rc == UPG_OK && (stages & UPG_MESH_WATER_CULL) && nt)
		{
			// valid = terrainSdfElevationRaw(p) < 0.1, grown twice over whole triangles; the rest is dropped
			float *elev = pool.get<float>(nv);
			uint32_t *va = pool.get<uint32_t>(nv), *vb = pool.get<uint32_t>(nv);
			uint8_t *dead = pool.get<uint8_t>(nt);
			MF_TRY(pool.err);
			if (e == cudaSuccess)
			{
				FieldLaunch L;
				L.out = elev;
				L.positions = po;
				L.count = nv;
				L.z0 = 0;
				L.field = FD_ELEVATION_RAW;
				L.status = status_word(planet);
				rc = launch_field(planet, L, s);
			}
			if (e == cudaSuccess && rc == UPG_OK)
			{
				cull_seed_kernel<<<blocks_of(nv), MF_BLOCK, 0, s>>>(elev, nv, va);
				for (int round = 0; round < 2 && e == cudaSuccess; round++)
				{
					MF_TRY(cudaMemcpyAsync(vb, va, sizeof(uint32_t) * nv, cudaMemcpyDeviceToDevice, s));
					cull_expand_kernel<<<blocks_of(nt), MF_BLOCK, 0, s>>>(io, nt, va, vb);
					uint32_t *t = va;
					va = vb;
					vb = t;
				}
				cull_faces_kernel<<<blocks_of(nt), MF_BLOCK, 0, s>>>(io, nt, va, dead);
				MF_LAUNCHED();
			}
			if (e == cudaSuccess && rc == UPG_OK)
			{
				Work c;
				c.positions = po;
				c.normals = no;
				c.idx = io;
				c.dead = dead;
				c.verts = nv;
				c.tris = nt;
				float *po2 = nullptr, *no2 = nullptr;
				uint32_t *io2 = nullptr;
				uint32_t nv2 = 0, nt2 = 0;
				e = compact(pool, c, false, &po2, &no2, &io2, &nv2, &nt2);
				cudaFreeAsync(po, s);
				cudaFreeAsync(no, s);
				cudaFreeAsync(io, s);
				po = po2;
				no = no2;
				io = io2;
				nv = nv2;
				nt = nt2;
			}
		}

		if (rc == UPG_OK && e != cudaSuccess)
			rc = cuda_fail(e, "mesh clean-up");
		if (rc != UPG_OK)
		{
			if (po)
				cudaFreeAsync(po, s);
			if (no)
				cudaFreeAsync(no, s);
			if (io)
				cudaFreeAsync(io, s);
			return rc;
		}
		out.positions = po;
		out.normals = no;
		out.indices = io;
		out.vertexCount = out.ownedVertices = nv;
		out.triangleCount = out.ownedTriangles = nt;
		return UPG_OK;
	}

	// ---- meshChunking (meshAlgorithms.cpp:1484-1619) as meshSplit calls it (sources/meshGeneration.cpp:173-182) --------------------------
	// A recursion: surface area -> leaf or split; split = longest axis of the bounding box, the best of seven candidate positions by the
	// area below each, meshClip of a copy against the two (enlarged) half boxes.  The clip is K9's; what is new is the DECISION, and it
	// hangs on two serial float sums over the triangles in index order (`result += t.area()`, `areas[i] += a`): their rounding depends on
	// the order, so they are accumulated in that order -- areas and the seven below-position bits are computed by all threads, then ONE
	// thread adds them up from shared memory while its warp prefetches (eight independent dependency chains per triangle).  The host
	// turns eight sums and a bounding box into the decision with the reference's float expressions; results are the reference's chunks
	// bit for bit (tests/test_gpu_chunks.py).
	namespace
	{
		__device__ __forceinline__ uint32_t order_key(float f)
		{
			const uint32_t u = __float_as_uint(f);
			return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
		}
		float key_to_float(uint32_t k)
		{
			const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
			float f;
			std::memcpy(&f, &u, 4);
			return f;
		}
		// Mesh::boundingBox (mesh.cpp:124-130): minimum / maximum over ALL positions; keys[0..2] = min, keys[3..5] = max
		__global__ void __launch_bounds__(MF_BLOCK) chunk_bbox_kernel(const float *__restrict__ positions, uint32_t verts, uint32_t *__restrict__ keys)
		{
			const uint32_t v = blockIdx.x * MF_BLOCK + threadIdx.x;
			const bool live = v < verts;
			const V3 p = live ? ld3(positions, v) : v3(0, 0, 0);
			const float c[3] = { p.x, p.y, p.z };
#pragma unroll
			for (int a = 0; a < 3; a++)
			{
				const uint32_t k = order_key(c[a]);
				const uint32_t mn = __reduce_min_sync(0xffffffffu, live ? k : 0xffffffffu), mx = __reduce_max_sync(0xffffffffu, live ? k : 0u);
				if ((threadIdx.x & 31u) == 0)
				{
					atomicMin(&keys[a], mn);
					atomicMax(&keys[3 + a], mx);
				}
			}
		}
		struct ChunkSplit
		{
			float positions[7]; // candidate planes on `axis` (world units)
			int axis;
		};
		// Triangle::area() = length(cross(b - a, c - a)) * 0.5 (geometry.h:89), center()[axis] = ((a + b) + c) / 3; bit i = center < positions[i]
		__global__ void __launch_bounds__(MF_BLOCK) chunk_areas_kernel(const float *__restrict__ positions, const uint32_t *__restrict__ idx, uint32_t tris, ChunkSplit split, float *__restrict__ area, uint8_t *__restrict__ below)
		{
			const uint32_t t = blockIdx.x * MF_BLOCK + threadIdx.x;
			if (t >= tris)
				return;
			const V3 a = ld3(positions, idx[t * 3 + 0]), b = ld3(positions, idx[t * 3 + 1]), c = ld3(positions, idx[t * 3 + 2]);
			area[t] = length(cross(b - a, c - a)) * 0.5f;
			const float p = axis_of(((a + b) + c) / 3.0f, split.axis);
			uint32_t m = 0;
#pragma unroll
			for (int i = 0; i < 7; i++)
				m |= (p < split.positions[i] ? 1u : 0u) << i;
			below[t] = (uint8_t)m;
		}
		// sums[0] = total area, sums[1 + i] = area below position i, both accumulated in triangle order by ONE thread
		__global__ void __launch_bounds__(32) chunk_serial_sums_kernel(const float *__restrict__ area, const uint8_t *__restrict__ below, uint32_t tris, float *__restrict__ sums)
		{
			constexpr uint32_t TILE = 1024;
			__shared__ float sa[TILE];
			__shared__ uint8_t sm[TILE];
			float total = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
			for (uint32_t base = 0; base < tris; base += TILE)
			{
				const uint32_t n = tris - base < TILE ? tris - base : TILE;
				for (uint32_t i = threadIdx.x; i < n; i += 32)
				{
					sa[i] = area[base + i];
					sm[i] = below[base + i];
				}
				__syncwarp();
				if (threadIdx.x == 0)
					for (uint32_t i = 0; i < n; i++)
					{
						const float a = sa[i];
						const uint32_t m = sm[i];
						total += a;
						if (m & 1u)
							s0 += a;
						if (m & 2u)
							s1 += a;
						if (m & 4u)
							s2 += a;
						if (m & 8u)
							s3 += a;
						if (m & 16u)
							s4 += a;
						if (m & 32u)
							s5 += a;
						if (m & 64u)
							s6 += a;
					}
				__syncwarp();
			}
			if (threadIdx.x == 0)
			{
				sums[0] = total;
				sums[1] = s0, sums[2] = s1, sums[3] = s2, sums[4] = s3, sums[5] = s4, sums[6] = s5, sums[7] = s6;
			}
		}

		int chunk_recurse(const float *positions, const float *normals, uint32_t verts, const uint32_t *indices, uint32_t tris, float maxArea, cudaStream_t s, std::vector<DmcResult> &leaves, DmcResult *own, int depth)
		{
			// own != nullptr: this node's arrays belong to *own (released here unless it becomes a leaf)
			auto leaf = [&]() -> int
			{
				if (own)
				{
					leaves.push_back(*own);
					*own = DmcResult();
					return UPG_OK;
				}
				// the root itself is the only chunk: a copy, like meshChunking's msh->copy()
				DmcResult c;
				cudaError_t e = device_alloc(&c.positions, sizeof(float) * 3 * (size_t)(verts ? verts : 1), s);
				if (e == cudaSuccess)
					e = device_alloc(&c.normals, sizeof(float) * 3 * (size_t)(verts ? verts : 1), s);
				if (e == cudaSuccess)
					e = device_alloc(&c.indices, sizeof(uint32_t) * 3 * (size_t)(tris ? tris : 1), s);
				if (e == cudaSuccess && verts)
					e = cudaMemcpyAsync(c.positions, positions, sizeof(float) * 3 * (size_t)verts, cudaMemcpyDeviceToDevice, s);
				if (e == cudaSuccess && verts)
					e = cudaMemcpyAsync(c.normals, normals, sizeof(float) * 3 * (size_t)verts, cudaMemcpyDeviceToDevice, s);
				if (e == cudaSuccess && tris)
					e = cudaMemcpyAsync(c.indices, indices, sizeof(uint32_t) * 3 * (size_t)tris, cudaMemcpyDeviceToDevice, s);
				if (e != cudaSuccess)
				{
					dmc_release(c, s);
					return cuda_fail(e, "mesh chunking (copy)");
				}
				c.vertexCount = c.ownedVertices = verts;
				c.triangleCount = c.ownedTriangles = tris;
				leaves.push_back(c);
				return UPG_OK;
			};
			if (tris == 0)
				return leaf(); // meshSurfaceArea = 0 < maxSurfaceArea: an empty chunk stays in the list (meshClip cleared it)
			if (depth > 48)
				return fail(UPG_ERR_ARGUMENT, "mesh chunking does not terminate (a split leaves the surface area unchanged)");
			Pool pool(s);
			uint32_t *keys = pool.get<uint32_t>(6);
			float *area = pool.get<float>(tris), *sums = pool.get<float>(8);
			uint8_t *below = pool.get<uint8_t>(tris);
			cudaError_t e = pool.err;
			uint32_t hostKeys[6] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0 };
			MF_TRY(cudaMemcpyAsync(keys, hostKeys, sizeof(hostKeys), cudaMemcpyHostToDevice, s));
			if (e == cudaSuccess)
			{
				chunk_bbox_kernel<<<blocks_of(verts), MF_BLOCK, 0, s>>>(positions, verts, keys);
				MF_LAUNCHED();
			}
			MF_TRY(cudaMemcpyAsync(hostKeys, keys, sizeof(hostKeys), cudaMemcpyDeviceToHost, s));
			MF_TRY(stream_sync(s));
			if (e != cudaSuccess)
				return cuda_fail(e, "mesh chunking (bounding box)");
			float a[3], b[3];
			for (int k = 0; k < 3; k++)
			{
				a[k] = key_to_float(hostKeys[k]);
				b[k] = key_to_float(hostKeys[3 + k]);
			}
			ChunkSplit split;
			split.axis = chunk_longest_axis(a, b);
			const int axis = split.axis;
			chunk_candidates(a, b, axis, split.positions);
			float hostSums[8];
			if (e == cudaSuccess)
			{
				chunk_areas_kernel<<<blocks_of(tris), MF_BLOCK, 0, s>>>(positions, indices, tris, split, area, below);
				chunk_serial_sums_kernel<<<1, 32, 0, s>>>(area, below, tris, sums);
				MF_LAUNCHED();
			}
			MF_TRY(cudaMemcpyAsync(hostSums, sums, sizeof(hostSums), cudaMemcpyDeviceToHost, s));
			MF_TRY(stream_sync(s));
			if (e != cudaSuccess)
				return cuda_fail(e, "mesh chunking (areas)");
			float fraction = 0;
			if (!chunk_wants_split(hostSums, maxArea, &fraction))
				return leaf();
			if (!(fraction == fraction))
				return fail(UPG_ERR_INVALID_VALUE, "mesh chunking: no split position (NaN surface area)");
			for (int half = 0; half < 2; half++)
			{
				float box[6];
				chunk_clipping_box(a, b, axis, fraction, half != 0, box);
				DmcResult child;
				int rc = mesh_clean(nullptr, positions, normals, verts, indices, tris, 0.0f, MESH_STAGE_CLIP_BOX, s, child, box);
				if (rc == UPG_OK)
					rc = chunk_recurse(child.positions, child.normals, child.vertexCount, child.indices, child.triangleCount, maxArea, s, leaves, &child, depth + 1);
				dmc_release(child, s);
				if (rc != UPG_OK)
					return rc;
			}
			return UPG_OK;
		}
	}

	int mesh_chunking(const float *positions, const float *normals, uint32_t verts, const uint32_t *indices, uint32_t tris, float maxSurfaceArea, cudaStream_t s, std::vector<DmcResult> &leaves)
	{
		if (verts >= LOCAL_BIT || (uint64_t)tris * 3 >= 0xffffffffull)
			return fail(UPG_ERR_ARGUMENT, "mesh too large for the chunking (%u vertices, %u triangles)", verts, tris);
		const int rc = chunk_recurse(positions, normals, verts, indices, tris, maxSurfaceArea, s, leaves, nullptr, 0);
		if (rc != UPG_OK)
		{
			for (auto &l : leaves)
				dmc_release(l, s);
			leaves.clear();
		}
		return rc;
	}
}

