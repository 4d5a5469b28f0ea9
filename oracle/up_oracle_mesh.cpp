/*
 * up_oracle_mesh.cpp -- TEST INFRASTRUCTURE: CPU oracle for stage 2, isosurface extraction.
 *
 * Serial restatement of
 *   dualmc::DualMC<float>::build(iso, manifold, shared vertices)
 *       externals/cage/externals/dualmc/dualmc/include/dualmc.tpp:12-32 (cell code), :37-86 (dual point code with
 *       Wenger's manifold inversion), :91-188 (dual point), :193-217 (first-touch vertex numbering), :384-461 (scan)
 *   the triangle tail of MarchingCubes::makeMesh
 *       externals/cage/sources/libcore/mesh/marchingCubes.cpp:186-227 (grid->world, shorter-diagonal split,
 *       degenerate filter, area-weighted normals)
 * using lookup tables DERIVED from the cube topology (tools/gen_dmc_tables.py, checked equal to the reference's).
 *
 * PINNING: bit-for-bit against the reference's own dualmc output and against the reference's final meshes in
 * tests/golden/grid_*.npz (tests/test_oracle_mesh.py); the reference's dualmc example counts (cube.raw: 1249
 * vertices / 1216 quads) are checked there too.
 */
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace
{
#define UPG_TABLE_QUALIFIER static const
#define UPG_TABLE_NAME(n) t_##n
#include "up_dmc_tables.inc"
#undef UPG_TABLE_QUALIFIER
#undef UPG_TABLE_NAME

	struct Volume
	{
		const float *data;
		int32_t dx, dy, dz;
		float at(int32_t x, int32_t y, int32_t z) const { return data[x + dx * (y + dy * z)]; } // gA, dualmc.h:229-231
	};

	int cellCode(const Volume &v, int32_t cx, int32_t cy, int32_t cz, float iso) // dualmc.tpp:12-32
	{
		int code = 0;
		for (int c = 0; c < 8; c++)
			if (v.at(cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)) >= iso)
				code |= 1 << c;
		return code;
	}

	int dualPointCode(const Volume &v, int32_t cx, int32_t cy, int32_t cz, float iso, bool manifold, int edgeBit) // :37-86
	{
		int code = cellCode(v, cx, cy, cz, iso);
		if (manifold)
		{
			const uint8_t dir = t_problematic[code];
			if (dir != 255)
			{
				int32_t n[3] = { cx, cy, cz };
				const int32_t dims[3] = { v.dx, v.dy, v.dz };
				const int comp = dir >> 1;
				n[comp] += (dir & 1) ? 1 : -1;
				if (n[comp] >= 0 && n[comp] < dims[comp] - 1)
				{
					const int ncode = cellCode(v, n[0], n[1], n[2], iso);
					if (t_problematic[ncode] != 255)
						code ^= 0xff;
				}
			}
		}
		for (int i = 0; i < 4; i++)
			if (t_dualPoints[code][i] & edgeBit)
				return t_dualPoints[code][i];
		return 0;
	}

	// edge e joins corner A and corner B (corner id = x + 2y + 4z); the crossing parameter runs from A to B
	const int edgeCorner[12][2] = { { 0, 1 }, { 1, 5 }, { 4, 5 }, { 0, 4 }, { 2, 3 }, { 3, 7 }, { 6, 7 }, { 2, 6 }, { 0, 2 }, { 1, 3 }, { 5, 7 }, { 4, 6 } };

	void dualPoint(const Volume &v, int32_t cx, int32_t cy, int32_t cz, float iso, int pointCode, float out[3]) // :91-188
	{
		float p[3] = { 0, 0, 0 };
		int points = 0;
		for (int e = 0; e < 12; e++)
		{
			if (!(pointCode & (1 << e)))
				continue;
			const int a = edgeCorner[e][0], b = edgeCorner[e][1];
			const int ax = a & 1, ay = (a >> 1) & 1, az = (a >> 2) & 1;
			const float va = v.at(cx + ax, cy + ay, cz + az);
			const float vb = v.at(cx + (b & 1), cy + ((b >> 1) & 1), cz + ((b >> 2) & 1));
			const float t = (iso - va) / (vb - va);
			const int axis = (a ^ b) == 1 ? 0 : (a ^ b) == 2 ? 1 : 2;
			// the reference adds, per edge, the fixed offsets of corner A (1.0f each) and the crossing on the edge's axis;
			// x, y, z accumulate independently so only the per-axis order of additions matters: edge order
			const int off[3] = { ax, ay, az };
			for (int k = 0; k < 3; k++)
			{
				if (k == axis)
					p[k] += t;
				else if (off[k])
					p[k] += 1.0f;
			}
			points++;
		}
		const float inv = 1.0f / (float)points;
		out[0] = (float)cx + p[0] * inv;
		out[1] = (float)cy + p[1] * inv;
		out[2] = (float)cz + p[2] * inv;
	}

	struct Key
	{
		int32_t cell;
		int code;
		bool operator==(const Key &o) const { return cell == o.cell && code == o.code; }
	};
	struct KeyHash
	{
		size_t operator()(const Key &k) const { return size_t((uint32_t)k.cell) | (size_t(k.code) << 32u); }
	};

	struct Builder
	{
		Volume v;
		float iso;
		bool manifold;
		std::vector<float> verts; // x y z
		std::vector<int32_t> quads; // 4 per quad
		std::unordered_map<Key, int32_t, KeyHash> seen;

		int32_t shared(int32_t cx, int32_t cy, int32_t cz, int edgeBit) // :193-217
		{
			Key k;
			k.cell = cx + v.dx * (cy + v.dy * cz);
			k.code = dualPointCode(v, cx, cy, cz, iso, manifold, edgeBit);
			auto it = seen.find(k);
			if (it != seen.end())
				return it->second;
			const int32_t id = (int32_t)(verts.size() / 3);
			float p[3];
			dualPoint(v, cx, cy, cz, iso, k.code, p);
			verts.insert(verts.end(), p, p + 3);
			seen[k] = id;
			return id;
		}

		void quad(bool firstOrder, int32_t i0, int32_t i1, int32_t i2, int32_t i3)
		{
			if (firstOrder)
				quads.insert(quads.end(), { i0, i1, i2, i3 });
			else
				quads.insert(quads.end(), { i0, i3, i2, i1 });
		}

		void run() // :384-461
		{
			enum
			{
				E0 = 1, E1 = 2, E2 = 4, E3 = 8, E4 = 16, E5 = 32, E6 = 64, E7 = 128, E8 = 256, E9 = 512, E10 = 1024, E11 = 2048
			};
			for (int32_t z = 0; z < v.dz - 2; z++)
				for (int32_t y = 0; y < v.dy - 2; y++)
					for (int32_t x = 0; x < v.dx - 2; x++)
					{
						const bool in0 = v.at(x, y, z) >= iso;
						if (z > 0 && y > 0)
						{
							const bool in1 = v.at(x + 1, y, z) >= iso;
							if (in0 != in1)
							{
								const int32_t i0 = shared(x, y, z, E0), i1 = shared(x, y, z - 1, E2), i2 = shared(x, y - 1, z - 1, E6), i3 = shared(x, y - 1, z, E4);
								quad(!in0, i0, i1, i2, i3); // entering keeps the order
							}
						}
						if (z > 0 && x > 0)
						{
							const bool in1 = v.at(x, y + 1, z) >= iso;
							if (in0 != in1)
							{
								const int32_t i0 = shared(x, y, z, E8), i1 = shared(x, y, z - 1, E11), i2 = shared(x - 1, y, z - 1, E10), i3 = shared(x - 1, y, z, E9);
								quad(in0, i0, i1, i2, i3); // exiting keeps the order
							}
						}
						if (x > 0 && y > 0)
						{
							const bool in1 = v.at(x, y, z + 1) >= iso;
							if (in0 != in1)
							{
								const int32_t i0 = shared(x, y, z, E3), i1 = shared(x - 1, y, z, E1), i2 = shared(x - 1, y - 1, z, E5), i3 = shared(x, y - 1, z, E7);
								quad(in0, i0, i1, i2, i3);
							}
						}
					}
		}
	};

	struct Result
	{
		std::vector<float> dmcVerts, positions, normals;
		std::vector<int32_t> quads;
		std::vector<uint32_t> indices;
	};

	inline float d2(const float *a, const float *b)
	{
		// distanceSquared = dot(d, d) with the sum started at zero (math.h:915-934)
		float s = 0;
		for (int i = 0; i < 3; i++)
		{
			const float d = a[i] - b[i];
			s += d * d;
		}
		return s;
	}
	inline void crossOf(const float *a, const float *b, const float *c, float out[3])
	{
		const float u[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] };
		const float w[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
		out[0] = u[1] * w[2] - u[2] * w[1];
		out[1] = u[2] * w[0] - u[0] * w[2];
		out[2] = u[0] * w[1] - u[1] * w[0];
	}
	inline float lengthOf(const float *v)
	{
		float s = 0;
		for (int i = 0; i < 3; i++)
			s += v[i] * v[i];
		return std::sqrt(s);
	}

	// marchingCubes.cpp:186-227
	void tail(Result &r, float boxSize, int res)
	{
		const float a = boxSize * -0.5f, b = boxSize * 0.5f;
		const float mult = (b - a) / ((float)res - 5);
		const float add = (b - a) * (((float)0 - 2) / ((float)res - 5)) + a; // cfg.position(0, 0, 0)
		const size_t nv = r.dmcVerts.size() / 3;
		r.positions.resize(nv * 3);
		for (size_t i = 0; i < nv * 3; i++)
			r.positions[i] = r.dmcVerts[i] * mult + add;
		r.normals.assign(nv * 3, 0.0f);
		static const int first[6] = { 0, 1, 2, 0, 2, 3 };
		static const int second[6] = { 1, 2, 3, 1, 3, 0 };
		for (size_t q = 0; q < r.quads.size() / 4; q++)
		{
			const uint32_t is[4] = { (uint32_t)r.quads[q * 4 + 0], (uint32_t)r.quads[q * 4 + 1], (uint32_t)r.quads[q * 4 + 2], (uint32_t)r.quads[q * 4 + 3] };
			const float *p[4] = { &r.positions[is[0] * 3], &r.positions[is[1] * 3], &r.positions[is[2] * 3], &r.positions[is[3] * 3] };
			const bool which = d2(p[0], p[2]) < d2(p[1], p[3]);
			const int *sel = which ? first : second;
			for (int t = 0; t < 2; t++)
			{
				const int *in = sel + t * 3;
				float n[3];
				crossOf(p[in[0]], p[in[1]], p[in[2]], n);
				const bool degenerated = lengthOf(n) * 0.5f < (float)1e-5; // Triangle::degenerated (shapes.cpp:119-122)
				if (degenerated)
					continue;
				for (int k = 0; k < 3; k++)
				{
					r.indices.push_back(is[in[k]]);
					for (int c = 0; c < 3; c++)
						r.normals[is[in[k]] * 3 + c] += n[c];
				}
			}
		}
		for (size_t i = 0; i < nv; i++)
		{
			float *n = &r.normals[i * 3];
			if (n[0] != 0 || n[1] != 0 || n[2] != 0)
			{
				const float l = lengthOf(n);
				for (int c = 0; c < 3; c++)
					n[c] = n[c] / l;
			}
		}
	}
}

extern "C"
{
	// returns an opaque result; query sizes, copy out, free
	void *upo_isosurface(const float *dens, int rx, int ry, int rz, float iso, int manifold, float boxSize, int fullRes)
	{
		Builder b;
		b.v = Volume{ dens, rx, ry, rz };
		b.iso = iso;
		b.manifold = !!manifold;
		b.run();
		Result *r = new Result();
		r->dmcVerts.swap(b.verts);
		r->quads.swap(b.quads);
		tail(*r, boxSize, fullRes);
		return r;
	}
	void upo_isosurface_sizes(void *h, uint64_t *nv, uint64_t *nq, uint64_t *ni)
	{
		const Result *r = (const Result *)h;
		*nv = r->dmcVerts.size() / 3;
		*nq = r->quads.size() / 4;
		*ni = r->indices.size();
	}
	void upo_isosurface_copy(void *h, float *dmcVerts, int32_t *quads, float *positions, float *normals, uint32_t *indices)
	{
		const Result *r = (const Result *)h;
		std::memcpy(dmcVerts, r->dmcVerts.data(), sizeof(float) * r->dmcVerts.size());
		std::memcpy(quads, r->quads.data(), sizeof(int32_t) * r->quads.size());
		std::memcpy(positions, r->positions.data(), sizeof(float) * r->positions.size());
		std::memcpy(normals, r->normals.data(), sizeof(float) * r->normals.size());
		std::memcpy(indices, r->indices.data(), sizeof(uint32_t) * r->indices.size());
	}
	void upo_isosurface_free(void *h)
	{
		delete (Result *)h;
	}
}
