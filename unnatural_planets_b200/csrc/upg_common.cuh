// upg_common.cuh -- scalar/vector helpers shared by every kernel of the generation core.
//
// Arithmetic contract: everything is IEEE binary32 evaluated in the order the
// reference writes it, WITHOUT fused multiply-add (the reference is built for baseline x86-64,
// which has no FMA: externals/cage/cmake/cage_build_configuration.cmake sets no -march).
// The library is compiled with nvcc -fmad=false; division and sqrt are the IEEE-rounded forms.
// Under that contract +,-,*,/,sqrt and every integer operation are bit-identical to the CPU
// reference; only libm transcendentals (powf, sinf, ...) differ (by a few ulp).
//
// The headers are also compilable by a plain host compiler (UPG_HD collapses to `inline`) so the
// host-side unit tests can exercise the very same source without a GPU.  That build is test-only;
// the shipped library has no CPU execution path.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define UPG_HD __host__ __device__ __forceinline__
#define UPG_D __device__ __forceinline__
#else
#define UPG_HD inline
#define UPG_D inline
#endif

// UPG_CONTRACT: the translation unit is compiled with -fmad=true (the tolerance build of the dense K1 kernels, UPG_ARITH_CONTRACT of
// include/upg.h).  The kernels carry the build in their type so that the two builds of one template never share a symbol.
#if defined(UPG_CONTRACT)
#define UPG_ARITH_BUILD 1
#else
#define UPG_ARITH_BUILD 0
#endif

namespace upg
{
	// Single IEEE operations that are NEVER fused with a neighbour, whatever the build: the grid coordinates and everything that feeds
	// a lattice index (frequency / lacunarity scaling, the simplex rotation) go through these, so that the integer lattice of the
	// contracted build is the exact build's bit for bit.  In the exact build (and on the host) they are the plain operators.
#if defined(__CUDA_ARCH__)
	UPG_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
	UPG_HD float add_rn(float a, float b) { return __fadd_rn(a, b); }
	UPG_HD float sub_rn(float a, float b) { return __fsub_rn(a, b); }
#else
	UPG_HD float mul_rn(float a, float b) { return a * b; }
	UPG_HD float add_rn(float a, float b) { return a + b; }
	UPG_HD float sub_rn(float a, float b) { return a - b; }
#endif

	// pow of a base that is known to be >= 0.  Exact build (and host): libm-grade powf, a few ulp from the CPU's (~45 instructions and a
	// call; 10 % of the dense kernel's issue slots).  Contracted build: ex2(y * lg2(x)) on the special-function unit, relative error
	// ~5e-7 for the arguments of the elevation curves -- the tolerance build states <= 1e-4 relative on the field (include/upg.h).
	UPG_HD float pow_pos(float x, float y)
	{
#if defined(__CUDA_ARCH__) && UPG_ARITH_BUILD
		return __powf(x, y);
#else
		return powf(x, y);
#endif
	}

	// a / b and sqrt(x) where the tolerance build may use the fast forms (no IEEE slow-path call inside the sample loop: a call makes the
	// compiler rebuild the shared-memory window base of the gradient table after it).  Exact build and host: the IEEE operations.
	UPG_HD float div_tol(float a, float b)
	{
#if defined(__CUDA_ARCH__) && UPG_ARITH_BUILD
		return __fdividef(a, b);
#else
		return a / b;
#endif
	}
	UPG_HD float sqrt_tol(float x)
	{
#if defined(__CUDA_ARCH__) && UPG_ARITH_BUILD
		return x > 0.0f ? x * rsqrtf(x) : sqrtf(x); // rsqrt(0) is +inf: zero, negative and NaN inputs take the IEEE form
#else
		return sqrtf(x);
#endif
	}

	// ---- cage scalar semantics (externals/cage/sources/include/cage-core/math.h:1021-1059, 798-809, 1173-1177) ----
	// min/max are the ternaries the reference uses (NaN-order sensitive), NOT fminf/fmaxf.
	UPG_HD float fmin_(float a, float b) { return a < b ? a : b; }
	UPG_HD float fmax_(float a, float b) { return a > b ? a : b; }
	UPG_HD float clamp_(float v, float lo, float hi) { return fmin_(fmax_(v, lo), hi); }
	UPG_HD float saturate_(float v) { return clamp_(v, 0.0f, 1.0f); }
	UPG_HD float abs_(float a) { return a < 0 ? -a : a; }
	UPG_HD float sqr_(float a) { return a * a; }
	UPG_HD float lerp_(float a, float b, float f) { return a + (b - a) * f; } // cage interpolate()
	UPG_HD float smoothstep_(float x) { return x * x * (3 - 2 * x); }
	UPG_HD float smootherstep_(float x) { return x * x * x * (x * (x * 6 - 15) + 10); }
	UPG_HD int sign_(float x) { return x < 0 ? -1 : x > 0 ? 1 : 0; }
	// Real % Real (math.h:442-446): l - r * (sint32)(l / r)
	UPG_HD float fmod_(float l, float r) { return l - (r * (float)(int32_t)(l / r)); }

	// murmur3 finaliser, cage hash() (math.h:1231-1239)
	UPG_HD uint32_t hash32(uint32_t key)
	{
		key ^= key >> 16;
		key *= 0x85ebca6bu;
		key ^= key >> 13;
		key *= 0xc2b2ae35u;
		key ^= key >> 16;
		return key;
	}

	struct V2
	{
		float x, y;
	};
	struct V3
	{
		float x, y, z;
	};
	struct V4
	{
		float x, y, z, w;
	};

	UPG_HD V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
	UPG_HD V2 v2(float x, float y) { return V2{ x, y }; }
	UPG_HD V3 operator+(V3 a, V3 b) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
	UPG_HD V3 operator-(V3 a, V3 b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
	UPG_HD V3 operator*(V3 a, V3 b) { return V3{ a.x * b.x, a.y * b.y, a.z * b.z }; }
	UPG_HD V3 operator*(V3 a, float s) { return V3{ a.x * s, a.y * s, a.z * s }; }
	UPG_HD V3 operator*(float s, V3 a) { return V3{ s * a.x, s * a.y, s * a.z }; }
	UPG_HD V3 operator/(V3 a, float s) { return V3{ a.x / s, a.y / s, a.z / s }; }
	UPG_HD V3 operator+(V3 a, float s) { return V3{ a.x + s, a.y + s, a.z + s }; }
	UPG_HD V3 operator-(V3 a, float s) { return V3{ a.x - s, a.y - s, a.z - s }; }
	UPG_HD V2 operator+(V2 a, V2 b) { return V2{ a.x + b.x, a.y + b.y }; }
	UPG_HD V2 operator-(V2 a, V2 b) { return V2{ a.x - b.x, a.y - b.y }; }
	UPG_HD V2 operator*(V2 a, float s) { return V2{ a.x * s, a.y * s }; }
	UPG_HD V2 operator*(float s, V2 a) { return V2{ s * a.x, s * a.y }; }
	UPG_HD V2 operator/(V2 a, float s) { return V2{ a.x / s, a.y / s }; }

	// cage dot(): sum starts at 0 and adds the products left to right (math.h:915-925)
	UPG_HD float dot(V3 a, V3 b) { return ((0.0f + a.x * b.x) + a.y * b.y) + a.z * b.z; }
	UPG_HD float dot(V2 a, V2 b) { return (0.0f + a.x * b.x) + a.y * b.y; }
	UPG_HD float dot(V4 a, V4 b) { return (((0.0f + a.x * b.x) + a.y * b.y) + a.z * b.z) + a.w * b.w; }
	UPG_HD float length_squared(V3 a) { return dot(a, a); }
	UPG_HD float length_squared(V2 a) { return dot(a, a); }
	UPG_HD float length(V3 a) { return sqrtf(dot(a, a)); }
	// length() out of never-fused operations: for a length that feeds a lattice index (the asteroid shape samples a noise at pos / |pos|)
	UPG_HD float length_rn(V3 a) { return sqrtf(add_rn(add_rn(add_rn(0.0f, mul_rn(a.x, a.x)), mul_rn(a.y, a.y)), mul_rn(a.z, a.z))); }
	UPG_HD float length(V2 a) { return sqrtf(dot(a, a)); }
	UPG_HD float distance_squared(V3 a, V3 b) { return length_squared(a - b); }
	UPG_HD float distance(V3 a, V3 b) { return sqrtf(distance_squared(a, b)); }
	UPG_HD V3 normalize(V3 a) { return a / length(a); }
	UPG_HD V2 normalize(V2 a) { return a / length(a); }
	UPG_HD V3 cross(V3 l, V3 r) { return V3{ l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x }; }
	UPG_HD V3 abs3(V3 a) { return V3{ abs_(a.x), abs_(a.y), abs_(a.z) }; }
	UPG_HD V2 abs2(V2 a) { return V2{ abs_(a.x), abs_(a.y) }; }
	UPG_HD V3 max3(V3 a, float s) { return V3{ fmax_(a.x, s), fmax_(a.y, s), fmax_(a.z, s) }; }
	UPG_HD V2 max2(V2 a, float s) { return V2{ fmax_(a.x, s), fmax_(a.y, s) }; }
	UPG_HD bool isnan_(float v) { return v != v; }

	// cage atan2(x, y) == std::atan2(y, x)  (libcore/math/scalars.cpp:84-87)
	UPG_HD float cage_atan2(float x, float y) { return atan2f(y, x); }

	// Quat * Vec3 (libcore/math/vectors.cpp:189-193); q = (x, y, z, w)
	UPG_HD V3 quat_rotate(V4 q, V3 r)
	{
		const V3 qv = V3{ q.x, q.y, q.z };
		const V3 t = cross(qv, r) * 2.0f;
		return (r + t * q.w) + cross(qv, t);
	}
}
