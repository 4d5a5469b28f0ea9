// upg_terrain.cuh -- the planet model: 34 base-shape SDFs, 6 elevation modes, the integer-hash
// Voronoi lattice and the field compositions, as device functions.
//
// Replaces (per sample) the function-pointer "plugins" of the reference:
//   shapes     sources/sdf.cpp:10-404  + cage primitives externals/cage/sources/libcore/signedDistanceFunctions.cpp:8-110
//   elevation  sources/elevationModes.cpp:9-198
//   voronoi    sources/voronoi.cpp:10-91
//   helpers    sources/math.cpp:8-38 (rescale, terrace, smoothMin/Max)
//   fields     sources/modesConfiguration.cpp:230-279 (land / water / navigation / elevation)
// Shape and elevation are template parameters so that each planet runs a kernel specialised for its
// two modes (no function pointers, no register cost from the shapes it does not use).
#pragma once

#include "upg_planet.h"

namespace upg
{
	// ---- sources/math.cpp -----------------------------------------------------------------------
	UPG_HD float smooth_min(float a, float b, float k)
	{
		const float h = saturate_(div_tol(b - a, k) * 0.5f + 0.5f);
		return lerp_(b, a, h) - k * h * (1 - h);
	}
	UPG_HD float smooth_max(float a, float b, float k) { return -smooth_min(-a, -b, k); }
	UPG_HD float terrace(float x, float steepness)
	{
		const float f = floorf(x);
		const float t = smootherstep_(saturate_((x - f) * fmax_(steepness, 1.0f))) + f;
		return lerp_(x, t, saturate_(steepness));
	}

	// ---- cage primitives (signedDistanceFunctions.cpp) ----------------------------------------------
	UPG_HD float sd_plane_origin(V3 pos, V3 normal)
	{
		// Plane(Vec3(), normal): d = -dot(0, n); c = n * d is a vector of signed zeros
		const float d = -dot(v3(0, 0, 0), normal);
		const V3 c = normal * d;
		return dot(normal, pos - c);
	}
	UPG_HD float sd_sphere(V3 pos, float radius) { return sqrt_tol(dot(pos, pos)) - radius; }
	UPG_HD float sd_capsule(V3 pos, float prolong, float radius)
	{
		V3 p = pos;
		p.z += prolong * 0.5f;
		p.z -= clamp_(p.z, 0.0f, prolong);
		return length(p) - radius;
	}
	UPG_HD float sd_cylinder(V3 pos, float halfHeight, float radius)
	{
		const V2 d = abs2(v2(length(v2(pos.x, pos.y)), pos.z)) - v2(radius, halfHeight);
		return fmin_(fmax_(d.x, d.y), 0.0f) + length(max2(d, 0.0f));
	}
	UPG_HD float sd_box(V3 pos, V3 radius)
	{
		const V3 p = abs3(pos) - radius;
		return length(max3(p, 0.0f)) + fmin_(fmax_(p.x, fmax_(p.y, p.z)), 0.0f);
	}
	UPG_HD float sd_octahedron(V3 pos, float radius)
	{
		const V3 p = abs3(pos);
		const float m = p.x + p.y + p.z - radius;
		V3 q;
		if (3.0f * p.x < m)
			q = p;
		else if (3.0f * p.y < m)
			q = v3(p.y, p.z, p.x);
		else if (3.0f * p.z < m)
			q = v3(p.z, p.x, p.y);
		else
			return m * 0.57735027f;
		const float k = clamp_(0.5f * (q.z - q.y + radius), 0.0f, radius);
		return length(v3(q.x, q.y - radius + k, q.z - k));
	}
	UPG_HD float sd_hexagonal_prism(V3 pos, float halfHeight, float radius)
	{
		const float kx = -0.8660254f, ky = 0.5f, kz = 0.57735f;
		V3 p = abs3(pos);
		{
			const float m = 2.0f * fmin_(dot(v2(kx, ky), v2(p.x, p.y)), 0.0f);
			p = p - v3(m * kx, m * ky, 0.0f);
		}
		const V2 e = v2(p.x, p.y) - v2(clamp_(p.x, -kz * radius, kz * radius), radius);
		const V2 d = v2(length(e) * (float)sign_(p.y - radius), p.z - halfHeight);
		return fmin_(fmax_(d.x, d.y), 0.0f) + length(max2(d, 0.0f));
	}
	UPG_HD float sd_torus(V3 pos, float major, float minor)
	{
		const V2 q = v2(length(v2(pos.x, pos.y)) - major, pos.z);
		return length(q) - minor;
	}
	UPG_HD float sd_knot(V3 pos, float scale, float k)
	{
		const float TAU = 3.14159274101257324f * 2;
		const float norm = (float)(2 / 14.4);
		scale *= norm;
		V3 p = pos / scale;
		const float r = length(v2(p.x, p.y));
		float a = cage_atan2(p.x, p.y);
		const float oa = k * a;
		a = fmod_(a, 0.001f * TAU) - 0.001f * TAU / 2;
		p.x = r * cosf(a) - 5;
		p.y = r * sinf(a);
		const float co = cosf(oa), so = sinf(oa);
		const V2 p2 = co * v2(p.x, p.z) + so * v2(-p.z, p.x);
		p.x = abs_(p2.x) - 1.5f;
		p.z = p2.y;
		return (length(p) - 0.7f) * scale;
	}

	// closestPoint(Vec3, Plane) / closestPoint(Vec3, Triangle) / distance(Vec3, Triangle)
	// (externals/cage/sources/libcore/geometry/geometry.cpp:638-679, 67-70) over host-prepared triangle constants
	UPG_HD V3 closest_on_plane(V3 pt, V3 normal, float d)
	{
		const V3 origin = normal * -d;
		return pt - dot(normal, pt - origin) * normal;
	}
	UPG_HD V3 closest_on_segment(V3 p, V3 origin, V3 dir, float maximum)
	{
		float d = dot(p - origin, dir);
		d = clamp_(d, 0.0f, maximum);
		return origin + dir * d;
	}
	UPG_HD V3 closest_on_triangle(V3 source, const TriangleConsts &t)
	{
		const V3 p = closest_on_plane(source, t.normal, t.d);
		{
			const V3 a = t.v[0] - p;
			const V3 b = t.v[1] - p;
			const V3 c = t.v[2] - p;
			const V3 u = cross(a, b);
			const V3 v = cross(b, c);
			const V3 w = cross(c, a);
			if (dot(u, v) > 0 && dot(u, w) > 0)
				return p;
		}
		const V3 pab = closest_on_segment(p, t.segOrigin[0], t.segDir[0], t.segMax[0]);
		const V3 pbc = closest_on_segment(p, t.segOrigin[1], t.segDir[1], t.segMax[1]);
		const V3 pca = closest_on_segment(p, t.segOrigin[2], t.segDir[2], t.segMax[2]);
		const float dab = distance_squared(p, pab);
		const float dbc = distance_squared(p, pbc);
		const float dca = distance_squared(p, pca);
		const float dm = fmin_(fmin_(dab, dbc), dca);
		if (dm == dab)
			return pab;
		if (dm == dbc)
			return pbc;
		return pca;
	}
	UPG_HD float distance_to_triangle(V3 a, const TriangleConsts &t) { return distance(closest_on_triangle(a, t), a); }

	UPG_HD float sd_tetrahedron(V3 pos, float radius, const TriangleConsts *tris)
	{
		const V3 p = pos / radius;
		float mad = -INFINITY;
#pragma unroll 1
		for (int i = 0; i < 4; i++)
		{
			const TriangleConsts &t = tris[i];
			const V3 k = closest_on_plane(p, t.normal, t.d);
			const float d = distance_to_triangle(p, t) * (float)sign_(dot(t.normal, p - k));
			mad = fmax_(mad, d);
		}
		return mad * radius;
	}

	// Quat(pitch, yaw, roll) (externals/cage/sources/libcore/math/vectors.cpp:118-133)
	UPG_HD V4 quat_euler(float pitch, float yaw, float roll)
	{
		const float ex = pitch * 0.5f, ey = yaw * 0.5f, ez = roll * 0.5f;
		const float c0 = cosf(ex), c1 = cosf(ey), c2 = cosf(ez);
		const float s0 = sinf(ex), s1 = sinf(ey), s2 = sinf(ez);
		V4 q;
		q.x = s0 * c1 * c2 - c0 * s1 * s2;
		q.y = c0 * s1 * c2 + s0 * c1 * s2;
		q.z = c0 * c1 * s2 - s0 * s1 * c2;
		q.w = c0 * c1 * c2 + s0 * s1 * s2;
		const float l = sqrtf((((0.0f + q.x * q.x) + q.y * q.y) + q.z * q.z) + q.w * q.w);
		q.x /= l;
		q.y /= l;
		q.z /= l;
		q.w /= l;
		return q;
	}

	// ---- neural SDFs (sources/sdf.cpp:243-324): 3 -> 16 -> 16 -> 16 -> 1 sine MLPs; weights in upg_neural_weights.inc ----
	struct NeuralWeights
	{
		float first[4][4][4]; // [block][row: y, z, x, bias][lane]
		float hidden[2][4][4][16]; // [layer][out block][in block][column-major Mat4]
		float bias[2][4][4]; // [layer][out block][lane]
		float out[4][4];
		float outBias;
		float xSign; // bunny: p[0] enters with a minus sign
		float div[2]; // 1.0 and 1.4
	};
#if defined(__CUDACC__)
#define UPG_CONST __constant__
#else
#define UPG_CONST static const
#endif
#include "upg_neural_weights.inc"

	UPG_HD V4 mat4_mul(const float *m, V4 v)
	{
		// res[i] = sum_j v[j] * m[j * 4 + i], accumulated from 0 (matrices.cpp:156-165)
		V4 r;
		r.x = (((0.0f + v.x * m[0]) + v.y * m[4]) + v.z * m[8]) + v.w * m[12];
		r.y = (((0.0f + v.x * m[1]) + v.y * m[5]) + v.z * m[9]) + v.w * m[13];
		r.z = (((0.0f + v.x * m[2]) + v.y * m[6]) + v.z * m[10]) + v.w * m[14];
		r.w = (((0.0f + v.x * m[3]) + v.y * m[7]) + v.z * m[11]) + v.w * m[15];
		return r;
	}
	UPG_HD V4 add4(V4 a, V4 b) { return V4{ a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
	UPG_HD V4 sin4(V4 a) { return V4{ sinf(a.x), sinf(a.y), sinf(a.z), sinf(a.w) }; }

	UPG_HD float sd_neural(const NeuralWeights &w, V3 p)
	{
		if (length(p) > 1)
			return length(p) - 0.8f;
		V4 f[4], g[4];
#pragma unroll
		for (int b = 0; b < 4; b++)
		{
			const float *wy = w.first[b][0], *wz = w.first[b][1], *wx = w.first[b][2], *wb = w.first[b][3];
			V4 a;
			// p[1] * A + p[2] * B (-/+) p[0] * C + D, evaluated left to right
			a.x = p.y * wy[0] + p.z * wz[0];
			a.y = p.y * wy[1] + p.z * wz[1];
			a.z = p.y * wy[2] + p.z * wz[2];
			a.w = p.y * wy[3] + p.z * wz[3];
			if (w.xSign < 0)
			{
				a.x = a.x - p.x * wx[0];
				a.y = a.y - p.x * wx[1];
				a.z = a.z - p.x * wx[2];
				a.w = a.w - p.x * wx[3];
			}
			else
			{
				a.x = a.x + p.x * wx[0];
				a.y = a.y + p.x * wx[1];
				a.z = a.z + p.x * wx[2];
				a.w = a.w + p.x * wx[3];
			}
			a.x += wb[0];
			a.y += wb[1];
			a.z += wb[2];
			a.w += wb[3];
			f[b] = sin4(a);
		}
#pragma unroll
		for (int layer = 0; layer < 2; layer++)
		{
#pragma unroll
			for (int o = 0; o < 4; o++)
			{
				V4 acc = mat4_mul(w.hidden[layer][o][0], f[0]);
				acc = add4(acc, mat4_mul(w.hidden[layer][o][1], f[1]));
				acc = add4(acc, mat4_mul(w.hidden[layer][o][2], f[2]));
				acc = add4(acc, mat4_mul(w.hidden[layer][o][3], f[3]));
				acc = add4(acc, V4{ w.bias[layer][o][0], w.bias[layer][o][1], w.bias[layer][o][2], w.bias[layer][o][3] });
				V4 s = sin4(acc);
				const float dv = w.div[layer];
				s = V4{ s.x / dv, s.y / dv, s.z / dv, s.w / dv };
				g[o] = add4(s, f[o]);
			}
#pragma unroll
			for (int o = 0; o < 4; o++)
				f[o] = g[o];
		}
		float r = dot(f[0], V4{ w.out[0][0], w.out[0][1], w.out[0][2], w.out[0][3] });
		r = r + dot(f[1], V4{ w.out[1][0], w.out[1][1], w.out[1][2], w.out[1][3] });
		r = r + dot(f[2], V4{ w.out[2][0], w.out[2][1], w.out[2][2], w.out[2][3] });
		r = r + dot(f[3], V4{ w.out[3][0], w.out[3][1], w.out[3][2], w.out[3][3] });
		return r + w.outBias;
	}

	// ---- the 34 shapes (sources/sdf.cpp) -------------------------------------------------------------
	UPG_HD float sd_rotated_rect(V2 point, V2 halfSizes, float rotation)
	{
		float a = cage_atan2(point.x, point.y);
		a += rotation;
		const V2 p = v2(cosf(a), sinf(a)) * length(point);
		const V2 d = abs2(p) - halfSizes;
		return length(max2(d, 0.0f)) + fmin_(fmax_(d.x, d.y), 0.0f);
	}

	UPG_HD V3 hemi_rotate(V3 p)
	{
		const float s = 0.707106781186547524400844362104f;
		return v3((p.x - p.y) * s, (p.x + p.y) * s, p.z);
	}

	template<int SHAPE>
	UPG_HD float shape_sdf(const PlanetParams &P, V3 pos)
	{
		const ShapeConsts &S = P.sh;
		switch (SHAPE)
		{
			case SH_HEXAGON:
				return sd_plane_origin(pos, S.hexagonNormal);
			case SH_SQUARE:
				return sd_plane_origin(pos, v3(0, 0, -1));
			case SH_SPHERE:
				return sd_sphere(pos, 1100);
			case SH_CAPSULE:
				return sd_capsule(v3(pos.x, pos.z, pos.y), 2500, 600);
			case SH_TUBE:
				return sd_cylinder(v3(pos.x, pos.z, pos.y), 10000, 800);
			case SH_DISK:
				return sd_cylinder(pos, 200, 600) - 100;
			case SH_BOX:
				return sd_box(pos, v3(550, 1000, 550)) - 200;
			case SH_CUBE:
				return sd_box(pos, v3(450, 450, 450)) - 100;
			case SH_TETRAHEDRON:
				return sd_tetrahedron(pos, 600, S.tetra) - 100;
			case SH_OCTAHEDRON:
				return sd_octahedron(pos, 800) - 100;
			case SH_TRIANGULARPRISM:
			{
				V3 p = v3(pos.x, pos.z, pos.y);
				p.z = fmax_(abs_(p.z) - 300 * 0.5f, 0.0f);
				return distance_to_triangle(p, S.prism) - 100;
			}
			case SH_HEXAGONALPRISM:
				return sd_hexagonal_prism(v3(pos.x, pos.z, pos.y), 300, 400) - 100;
			case SH_TORUS:
				return sd_torus(pos, 950, 530);
			case SH_KNOT:
				return sd_knot(pos * v3(1, 1, 0.3f), 1600, 1.5f) / S.knotInvLen;
			case SH_MOBIUSSTRIP:
			{
				const V3 p = v3(pos.x, pos.z, pos.y);
				const float planeRotation = cage_atan2(p.x, p.z);
				const V2 xz = v2(p.x, p.z);
				const V2 proj = v2(dot(xz, normalize(xz)), p.y);
				return sd_rotated_rect(proj + v2(-1000, 0), v2(400, 40), planeRotation / 2) - 200;
			}
			case SH_FIBERS:
			{
				const float scale = 0.0007f;
				const V3 p = pos * scale + S.fibersOffset;
				float g;
				{
					const V3 q = p * 3.23f;
					const V3 a = v3(sinf(q.x), sinf(q.y), sinf(q.z));
					const V3 b = v3(cosf(q.z), cosf(q.x), cosf(q.y));
					g = abs_(dot(a, b) + 1.6f) / 3.23f - 0.2f;
				}
				const float v = g * 0.7f / scale;
				const float d = fmax_(length(pos) - 1500, 0.0f) * 0.5f;
				return v + d;
			}
			case SH_H2O:
			{
				const float scale = 0.75f;
				const V3 p = pos * scale;
				const float h1 = sd_sphere(p - v3(-550, 300, 0), 450);
				const float h2 = sd_sphere(p - v3(+550, 300, 0), 450);
				const float o = sd_sphere(p - v3(0, -100, 0), 650);
				return smooth_min(o, fmin_(h1, h2), 150) / scale;
			}
			case SH_H3O:
			{
				const float scale = 0.75f;
				const V3 p = pos * scale;
				const float h0 = sd_sphere(p - S.h3oCenters[0], 450);
				const float h1 = sd_sphere(p - S.h3oCenters[1], 450);
				const float h2 = sd_sphere(p - S.h3oCenters[2], 450);
				const float o = sd_sphere(p, 650);
				const float h = fmin_(fmin_(h0, h1), h2);
				return smooth_min(o, h, 150) / scale;
			}
			case SH_H4O:
			{
				const float scale = 0.75f;
				const V3 p = pos * scale;
				const float h0 = sd_sphere(p - v3(-550, +400, 0), 450);
				const float h1 = sd_sphere(p - v3(+550, +400, 0), 450);
				const float h2 = sd_sphere(p - v3(0, -400, -550), 450);
				const float h3 = sd_sphere(p - v3(0, -400, +550), 450);
				const float o = sd_sphere(p, 650);
				const float h = fmin_(fmin_(h0, h1), fmin_(h2, h3));
				return smooth_min(o, h, 150) / scale;
			}
			case SH_GEAR:
			{
				const float scale = 2;
				const V3 q3 = v3(pos.x, pos.z, pos.y) / scale;
				const V2 p = v2(q3.x, q3.z);
				const uint32_t teeths = 9;
				const float angle = (3.14159274101257324f * 2) / (float)teeths;
				const uint32_t sector = (uint32_t)(cage_atan2(p.x, p.y) / angle + (float)teeths + 0.5f) % teeths;
				V2 q;
				{
					const float a = (float)sector * -angle;
					const float s = sinf(a), c = cosf(a);
					q = v2(c * p.x - s * p.y, s * p.x + c * p.y);
				}
				const float b = length(max2(abs2(q - v2(700, 0)) - v2(150, 70), 0.0f));
				const float c = abs_(length(p) - 450) - 150;
				const float h = abs_(q3.y) - 30;
				return (smooth_max(fmin_(b, c), h, 50) - 50) * scale;
			}
			case SH_MANDELBULB:
			{
				if (length_squared(pos) < 1e-3f)
					return 0;
				const V3 c = pos * 0.0004f;
				const float Power = 3.0f;
				V3 z = c;
				float dr = 1.0f;
				float r = 0.0f;
#pragma unroll 1
				for (uint32_t i = 0; i < 20; i++)
				{
					r = length(z);
					if (r > 4.0f)
						break;
					float theta = acosf(z.z / r);
					float phi = cage_atan2(z.y, z.x);
					dr = powf(r, Power - 1.0f) * Power * dr + 1.0f;
					const float zr = powf(r, Power);
					theta = theta * Power;
					phi = phi * Power;
					z = zr * v3(sinf(theta) * cosf(phi), sinf(phi) * sinf(theta), cosf(theta));
					z = z + c;
				}
				const float value = 0.5f * logf(r) * r / dr;
				return (value + 0.1f) * 150;
			}
			case SH_TWISTEDHEXAGONALPRISM:
			{
				const float angle = pos.y * 0.001f;
				V3 p = v3(pos.x, pos.z, 0);
				p = quat_rotate(quat_euler(0.0f, 0.0f, angle), p);
				p = v3(p.x, pos.y, p.y);
				return sd_hexagonal_prism(v3(p.x, p.z, p.y), 900, 650) - 200;
			}
			case SH_BUNNY:
			{
				const float scale = 0.0005f;
				return sd_neural(g_bunnyWeights, v3(pos.x, pos.z, pos.y) * scale) / scale;
			}
			case SH_MONKEYHEAD:
			{
				const float scale = 0.0005f;
				return sd_neural(g_monkeyWeights, v3(pos.x, pos.z, pos.y) * scale) / scale;
			}
			case SH_DOUBLETORUS:
			{
				const V3 p1 = v3(pos.y, pos.z, pos.x);
				const V3 p2 = v3(pos.y, pos.x, pos.z);
				return smooth_min(sd_torus(p1, 900, 400), sd_torus(p2, 900, 400), 200);
			}
			case SH_TORUSCROSS:
			{
				const float ma = 750;
				const V3 p1 = v3(pos.y - ma, pos.z, pos.x);
				const V3 p2 = v3(pos.y + ma, pos.x, pos.z);
				return smooth_min(sd_torus(p1, ma, 400), sd_torus(p2, ma, 400), 200);
			}
			case SH_BOWL:
				return 2000 - distance(pos, v3(0, 0, -3000));
			case SH_INSIDECUBE:
				return -(sd_box(pos, v3(1300, 1300, 1300)) - 200);
			case SH_ASTEROID:
			{
				const float l = length_rn(pos); // exact in every build: d reaches the lattice of the noise (OpenSimplex2 has small steps at axis ties)
				if (l < 1)
					return -100;
				const V3 d = pos / l;
				const float n = noise3<NT_SIMPLEX, FT_FBM, 2>(S.asteroid, d);
				return l + n * 600 - 1300;
			}
			case SH_PIPE:
			{
				V3 p = pos;
				if (p.x < 0)
				{
					p.x *= -1;
					const float t = p.y;
					p.y = p.z;
					p.z = t;
				}
				p.y = abs_(p.y) - 800;
				{
					const float t = p.y;
					p.y = p.z;
					p.z = t;
				}
				return sd_torus(p, 800, 400);
			}
			case SH_TWISTEDPLANE:
			{
				const float rot = 1 - sqr_(1 - fmin_(1.0f, abs_(pos.z) / 1000));
				const float degs = (float)sign_(pos.z) * rot * 35;
				const float angle = degs * 3.14159274101257324f / 180;
				return pos.y * cosf(angle) + pos.x * sinf(angle);
			}
			case SH_BELT:
			{
				const float Radius = 1800;
				if (pos.x > 0)
					return sd_cylinder(v3(pos.x, pos.z, pos.y), 10000, Radius) * -1;
				return Radius - abs_(pos.z);
			}
			case SH_HEMISPHERES:
			{
				const V3 p = v3(abs_(pos.x), abs_(pos.y), pos.z);
				const V3 q = hemi_rotate(p + v3(-2000, 0, 0));
				const float c = sd_cylinder(v3(q.y, q.z, q.x), 10000, 200);
				const float s1 = sd_sphere(q, 1000);
				const float s2 = sd_sphere(hemi_rotate(p + v3(0, -2000, 0)), 1000);
				const float s = fmin_(s1, s2);
				return smooth_min(c, s, 100);
			}
			case SH_WORMHOLE:
			{
				const float x = fmin_(length(v2(pos.x, pos.z)) - 1400, 0.0f);
				const float y = pos.y;
				return 700 - length(v2(x, y));
			}
		}
		return 0;
	}

	// ---- Voronoi (sources/voronoi.cpp) -------------------------------------------------------------
	struct I3
	{
		int32_t x, y, z;
	};
	UPG_HD I3 voronoi_mix(const VoronoiInst &v, I3 s)
	{
		I3 r;
		r.x = iadd(iadd(iadd(v.seedHashes[0], s.x), s.y), (int32_t)hash32((uint32_t)s.x));
		r.y = iadd(iadd(iadd(v.seedHashes[1], s.y), s.z), (int32_t)hash32((uint32_t)s.y));
		r.z = iadd(iadd(iadd(v.seedHashes[2], s.z), s.x), (int32_t)hash32((uint32_t)s.z));
		return r;
	}
	UPG_HD V3 voronoi_point(const VoronoiInst &v, I3 s, I3 cell)
	{
		// Vec3(s % 65536) / 65535: C++ remainder keeps the sign of the dividend
		const V3 g = v3((float)(s.x % 65536), (float)(s.y % 65536), (float)(s.z % 65536)) / 65535.0f;
		return (g + v3((float)cell.x, (float)cell.y, (float)cell.z)) * v.cellSize;
	}

	// nearest N feature points (N <= 4), optional projection of the candidates into the tangent plane
	template<int N, bool PROJECT>
	UPG_HD void voronoi_nearest(const VoronoiInst &v, V3 position, V3 normal, V3 *outPoints)
	{
		float bestD[N];
#pragma unroll
		for (int i = 0; i < N; i++)
		{
			bestD[i] = INFINITY;
			outPoints[i] = v3(0, 0, 0);
		}
		const V3 pf = position * v.frequency;
		const I3 base = I3{ (int32_t)pf.x, (int32_t)pf.y, (int32_t)pf.z };
		float planeD = 0;
		if (PROJECT)
			planeD = -dot(position, normal); // Plane(position, normal)
#pragma unroll 1
		for (int32_t dz = -1; dz < 2; dz++)
#pragma unroll 1
			for (int32_t dy = -1; dy < 2; dy++)
#pragma unroll 1
				for (int32_t dx = -1; dx < 2; dx++)
				{
					const I3 cell = I3{ base.x + dx, base.y + dy, base.z + dz };
					I3 s = voronoi_mix(v, cell);
					for (uint32_t i = 0; i < v.pointsPerCell; i++)
					{
						s = voronoi_mix(v, s);
						V3 p = voronoi_point(v, s, cell);
						if (PROJECT)
							p = closest_on_plane(p, normal, planeD);
						const float d = distance_squared(p, position);
						// insertion into the sorted best-N list (strict <, so the earlier point wins a tie)
#pragma unroll
						for (int k = 0; k < N; k++)
						{
							if (d < bestD[k])
							{
#pragma unroll
								for (int m = N - 1; m > k; m--)
								{
									bestD[m] = bestD[m - 1];
									outPoints[m] = outPoints[m - 1];
								}
								bestD[k] = d;
								outPoints[k] = p;
								break;
							}
						}
					}
				}
	}

	// ---- elevation modes (sources/elevationModes.cpp) ------------------------------------------------
	// `maskColumn`: the column kernel's cache of the mask noise (PerlinColumn, upg_noise.cuh), or null = evaluate it from scratch
	UPG_HD float elevation_mountains(const ElevationConsts &E, V3 pos, float land, PerlinColumn *maskColumn = nullptr, PerlinFbmColumn<3> *terraceColumn = nullptr)
	{
		const float cover = 1 - saturate_(land * -0.1f);
		if (cover < 1e-7f)
			return land;

		// noise3<NT_PERLIN, FT_NONE>: frequency scale, then the single generator
		const float mask = maskColumn ? perlin_column(E.n[1].seed, mul_rn(pos.x, E.n[1].frequency), mul_rn(pos.y, E.n[1].frequency), mul_rn(pos.z, E.n[1].frequency), *maskColumn) : noise3<NT_PERLIN, FT_NONE, 3>(E.n[1], pos);
		const float rm = smoothstep_(saturate_(mask * +7 - 0.3f));
		const float tm = smoothstep_(saturate_(mask * -7 - 1.5f));

		// The two masks are never non-zero together (rm > 0 needs mask > 0.3/7, tm > 0 needs mask < -1.5/7), and a mask that is
		// exactly +0 makes its term exactly +0: pow(max(.,0), 1.6) and terrace(max(.,0) * 2.5, 4) are finite and >= +0, so
		// term * (+0 * cover) * scale == +0.0f bit for bit.  The noise behind a zero mask is therefore not evaluated: 4
		// OpenSimplex2 octaves and / or 3 Perlin octaves saved per sample with IDENTICAL results (tests/test_host_terrain.py
		// checks the host compilation of this header against the reference bit for bit).
		float ridge = 0.0f;
		if (rm != 0.0f)
		{
			ridge = noise3<NT_SIMPLEX, FT_RIDGED, 4>(E.n[2], pos);
			ridge = fmax_(ridge - 0.1f, 0.0f);
			ridge = pow_pos(ridge, 1.6f); // ridge >= 0 after the max above
			ridge *= rm * cover;
			ridge *= 1000;
		}

		float terraces = 0.0f;
		if (tm != 0.0f)
		{
			terraces = terraceColumn ? perlin_fbm_column<3>(E.n[3], pos.x, pos.y, pos.z, *terraceColumn) : noise3<NT_PERLIN, FT_FBM, 3>(E.n[3], pos);
			terraces = fmax_(terraces + 0.1f, 0.0f) * 2.5f;
			terraces = terrace(terraces, 4);
			terraces *= tm * cover;
			terraces *= 250;
		}

		return land + smooth_max(0, fmax_(ridge, terraces), 50);
	}

	// the modes whose FIRST noise is a four-octave Value fBm of the sample position itself (the column kernel caches it, upg_field.cuh)
	template<int ELEV>
	struct ElevationStartsWithValueFbm
	{
		static constexpr bool value = ELEV == EL_LEGACY || ELEV == EL_LAKES || ELEV == EL_ISLANDS;
	};

	// elevation_raw with that first noise value handed in (`first` is ignored by the other modes)
	// `cratersNearest`: the nearest Voronoi point of `pos` when the caller already has it (the dense kernel finds it from a table of
	// points its block shares, upg_field.cuh), or null = search the 27 cells here
	template<int ELEV>
	UPG_HD float elevation_raw_given(const PlanetParams &P, V3 pos, float first, PerlinColumn *maskColumn = nullptr, PerlinFbmColumn<3> *terraceColumn = nullptr, const V3 *cratersNearest = nullptr)
	{
		const ElevationConsts &E = P.el;
		switch (ELEV)
		{
			case EL_NONE:
				return 100;
			case EL_SIMPLE:
			{
				float a = noise3<NT_SIMPLEX, FT_RIDGED, 6>(E.n[0], pos);
				a = -a + 0.3f;
				a = powf(a * 1.3f - 0.35f, 3.0f) + 0.1f;
				return 100 - a * 1000;
			}
			case EL_LEGACY:
			{
				const float scale = first * 0.0005f + 0.0015f;
				float a = noise3<NT_VALUE, FT_FBM, 4>(E.n[1], pos * scale);
				a += 0.11f;
				if (a < 0)
					a = -powf(-a, 0.85f);
				else
					a = powf(a, 1.7f);
				return a * 2500;
			}
			case EL_LAKES:
			case EL_ISLANDS:
			{
				float land = first * 0.5f + 0.5f;
				land = saturate_(land);
				land = 1 - pow_pos(land, ELEV == EL_LAKES ? 1.24f : 0.83f); // land is saturated to [0, 1]
				land = land * 2 - 1;
				land = div_tol(land, abs_(land) + 0.17f) + 0.15f;
				land *= 150;
				return elevation_mountains(E, pos, land, maskColumn, terraceColumn);
			}
			case EL_CRATERS:
			{
				V3 c;
				if (cratersNearest)
					c = *cratersNearest;
				else
					voronoi_nearest<1, false>(E.voronoi, pos, v3(0, 0, 0), &c);
				const float s = 100 + noise3<NT_VALUE, FT_NONE, 3>(E.n[0], c) * 50;
				const float d = distance(pos, c);
				const float x = saturate_(d / s);
				const float h = (powf(x, 4.0f) - 0.9f * powf(x, 5.0f) - 0.1f) * 10;
				return 200 + h * 190;
			}
		}
		return 0;
	}

	template<int ELEV>
	UPG_HD float elevation_raw(const PlanetParams &P, V3 pos)
	{
		float first = 0;
		if (ElevationStartsWithValueFbm<ELEV>::value)
			first = noise3<NT_VALUE, FT_FBM, 4>(P.el.n[0], pos);
		return elevation_raw_given<ELEV>(P, pos, first);
	}

	// ---- field compositions (sources/modesConfiguration.cpp:227-279), meshElevationRatio = 10 ------
	template<int SHAPE, int ELEV, int FIELD>
	UPG_HD float terrain_field(const PlanetParams &P, V3 pos)
	{
		if (FIELD == FD_WATER)
			return shape_sdf<SHAPE>(P, pos);
		if (FIELD == FD_ELEVATION)
			return shape_sdf<SHAPE>(P, pos) * 10.0f;
		if (FIELD == FD_ELEVATION_RAW)
			return elevation_raw<ELEV>(P, pos);
		const float base = shape_sdf<SHAPE>(P, pos);
		const float elev = elevation_raw<ELEV>(P, pos) / 10.0f;
		if (FIELD == FD_NAVIGATION)
			return base - fmax_(elev, 0.0f);
		return base - elev;
	}

	// grid index -> world coordinate on one axis (externals/cage/sources/libcore/mesh/marchingCubes.cpp:247-257)
	UPG_HD float grid_coord(const GridConsts &g, uint32_t i)
	{
		const float f = ((float)i - 2.0f) / g.resMinus5;
		return add_rn(mul_rn(g.boxSpan, f), g.boxA); // never fused: the sample positions are the reference's in every build
	}
}
