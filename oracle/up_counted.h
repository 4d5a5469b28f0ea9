/*
 * up_counted.h -- TEST INFRASTRUCTURE (oracle): an instrumented stand-in for `float` / `double`.
 *
 * SURVEY.md section 8(d) asks for the arithmetic work per unit (SDF sample, baked texel) as EXACT COUNTS obtained by instrumenting
 * the CPU restatement, not as estimates.  A translation unit that does
 *
 *     #include "up_counted.h"
 *     #define float upc::cfloat
 *     #define double upc::cdouble
 *     #include "<the restatement>.cpp"
 *
 * compiles the very same source with every scalar operation routed through the operators below, which compute in the native
 * type (bit-identical results: the counted build is checked against the plain one) and bump one per-thread counter per
 * operation class.  The reference is built for baseline x86-64 without FMA, so a multiply-add is one fmul and one fadd here;
 * the roofline numerator counts flops = fadd + fmul + fdiv + fsqrt (FMA = 2 under that convention), comparisons / negations /
 * conversions and libm calls are reported separately.
 */
#pragma once

#include <cmath>
#include <cstdint>
#include <type_traits>

namespace upc
{
	enum Op
	{
		OP_FADD = 0, // + and - (binary)
		OP_FMUL,
		OP_FDIV,
		OP_FSQRT,
		OP_FCMP, // < <= > >= == !=
		OP_FNEG, // unary minus, fabs
		OP_I2F, // integer -> float
		OP_F2I, // float -> integer (truncation), floor / round helpers
		OP_SPECIAL, // libm: pow, sin, cos, tan, atan2, acos, asin, atan, log, exp, fmod, cbrt
		OP_DADD, // the same classes in binary64 (HSLuv)
		OP_DMUL,
		OP_DDIV,
		OP_DSQRT,
		OP_DCMP,
		OP_DSPECIAL,
		OP_F2D, // float <-> double
		OP_COUNT
	};

	struct Counters
	{
		uint64_t n[OP_COUNT];
	};
	inline Counters &counters()
	{
		static thread_local Counters c = {};
		return c;
	}
	inline void bump(int op) { counters().n[op]++; }

	template<class T, int ADD, int MUL, int DIV, int SQRT, int CMP, int SPECIAL>
	struct counted
	{
		T v;
		counted() = default;
		// from any arithmetic type: integer sources are a conversion instruction, floating literals are free
		template<class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>
		constexpr __attribute__((always_inline)) counted(U u) : v((T)u)
		{
			// an integer LITERAL (`x * 2`) is folded by every compiler: only run-time integers cost a conversion
			if (std::is_integral<U>::value && !__builtin_is_constant_evaluated() && !__builtin_constant_p(u))
				bump(OP_I2F);
		}
		// float -> double promotes implicitly like the native types do (`lo + (hi - lo) * f` with a float f); narrowing needs a cast
		template<class U, int A, int M, int D, int S, int C, int X, typename std::enable_if<(sizeof(U) < sizeof(T)), int>::type = 0>
		counted(counted<U, A, M, D, S, C, X> o) : v((T)o.v)
		{
			bump(OP_F2D);
		}
		template<class U, int A, int M, int D, int S, int C, int X, typename std::enable_if<(sizeof(U) > sizeof(T)), int>::type = 0>
		explicit counted(counted<U, A, M, D, S, C, X> o) : v((T)o.v)
		{
			bump(OP_F2D);
		}
		template<class U, class = typename std::enable_if<std::is_integral<U>::value>::type>
		explicit operator U() const
		{
			bump(OP_F2I);
			return (U)v;
		}
		explicit operator float() const { return (float)v; }
		explicit operator double() const { return (double)v; }
		explicit operator bool() const { return v != 0; }

		counted operator-() const
		{
			bump(OP_FNEG);
			counted r;
			r.v = -v;
			return r;
		}
		counted operator+() const { return *this; }
		counted &operator+=(counted o)
		{
			bump(ADD);
			v += o.v;
			return *this;
		}
		counted &operator-=(counted o)
		{
			bump(ADD);
			v -= o.v;
			return *this;
		}
		counted &operator*=(counted o)
		{
			bump(MUL);
			v *= o.v;
			return *this;
		}
		counted &operator/=(counted o)
		{
			bump(DIV);
			v /= o.v;
			return *this;
		}
		friend counted operator+(counted a, counted b)
		{
			bump(ADD);
			counted r;
			r.v = a.v + b.v;
			return r;
		}
		friend counted operator-(counted a, counted b)
		{
			bump(ADD);
			counted r;
			r.v = a.v - b.v;
			return r;
		}
		friend counted operator*(counted a, counted b)
		{
			bump(MUL);
			counted r;
			r.v = a.v * b.v;
			return r;
		}
		friend counted operator/(counted a, counted b)
		{
			bump(DIV);
			counted r;
			r.v = a.v / b.v;
			return r;
		}
#define UPC_CMP(OPERATOR) \
	friend bool operator OPERATOR(counted a, counted b) \
	{ \
		bump(CMP); \
		return a.v OPERATOR b.v; \
	}
		UPC_CMP(<)
		UPC_CMP(<=)
		UPC_CMP(>)
		UPC_CMP(>=)
		UPC_CMP(==)
		UPC_CMP(!=)
#undef UPC_CMP
	};

	typedef counted<float, OP_FADD, OP_FMUL, OP_FDIV, OP_FSQRT, OP_FCMP, OP_SPECIAL> cfloat;
	typedef counted<double, OP_DADD, OP_DMUL, OP_DDIV, OP_DSQRT, OP_DCMP, OP_DSPECIAL> cdouble;
	static_assert(sizeof(cfloat) == sizeof(float) && sizeof(cdouble) == sizeof(double), "layout compatible with the native types");
	// trivially copyable: passed and returned in registers like the native type, so the counted libraries keep the C ABI of the plain ones
	static_assert(std::is_trivially_copyable<cfloat>::value && std::is_trivially_copyable<cdouble>::value, "same calling convention as the native types");

	template<class C>
	struct traits;
	template<>
	struct traits<cfloat>
	{
		enum { SQRT = OP_FSQRT, SPECIAL = OP_SPECIAL };
	};
	template<>
	struct traits<cdouble>
	{
		enum { SQRT = OP_DSQRT, SPECIAL = OP_DSPECIAL };
	};

#define UPC_UNARY(NAME, CLASS) \
	inline cfloat NAME(cfloat a) \
	{ \
		bump(CLASS == 0 ? (int)OP_FSQRT : CLASS == 1 ? (int)OP_SPECIAL : CLASS == 2 ? (int)OP_FNEG : (int)OP_F2I); \
		cfloat r; \
		r.v = std::NAME(a.v); \
		return r; \
	} \
	inline cdouble NAME(cdouble a) \
	{ \
		bump(CLASS == 0 ? (int)OP_DSQRT : CLASS == 1 ? (int)OP_DSPECIAL : CLASS == 2 ? (int)OP_FNEG : (int)OP_F2I); \
		cdouble r; \
		r.v = std::NAME(a.v); \
		return r; \
	}
	UPC_UNARY(sqrt, 0)
	UPC_UNARY(sin, 1)
	UPC_UNARY(cos, 1)
	UPC_UNARY(tan, 1)
	UPC_UNARY(asin, 1)
	UPC_UNARY(acos, 1)
	UPC_UNARY(atan, 1)
	UPC_UNARY(exp, 1)
	UPC_UNARY(log, 1)
	UPC_UNARY(cbrt, 1)
	UPC_UNARY(fabs, 2)
	UPC_UNARY(abs, 2)
	UPC_UNARY(floor, 3)
	UPC_UNARY(ceil, 3)
	UPC_UNARY(round, 3)
	UPC_UNARY(trunc, 3)
#undef UPC_UNARY
#define UPC_BINARY(NAME) \
	inline cfloat NAME(cfloat a, cfloat b) \
	{ \
		bump(OP_SPECIAL); \
		cfloat r; \
		r.v = std::NAME(a.v, b.v); \
		return r; \
	} \
	inline cdouble NAME(cdouble a, cdouble b) \
	{ \
		bump(OP_DSPECIAL); \
		cdouble r; \
		r.v = std::NAME(a.v, b.v); \
		return r; \
	}
	UPC_BINARY(pow)
	UPC_BINARY(atan2)
	UPC_BINARY(fmod)
#undef UPC_BINARY
	inline bool isnan(cfloat a) { return std::isnan(a.v); }
	inline bool isnan(cdouble a) { return std::isnan(a.v); }
	inline bool isfinite(cfloat a) { return std::isfinite(a.v); }
	inline bool isfinite(cdouble a) { return std::isfinite(a.v); }
	inline cfloat fmin(cfloat a, cfloat b)
	{
		bump(OP_FCMP);
		cfloat r;
		r.v = std::fmin(a.v, b.v);
		return r;
	}
	inline cfloat fmax(cfloat a, cfloat b)
	{
		bump(OP_FCMP);
		cfloat r;
		r.v = std::fmax(a.v, b.v);
		return r;
	}
	inline cdouble fmin(cdouble a, cdouble b)
	{
		bump(OP_DCMP);
		cdouble r;
		r.v = std::fmin(a.v, b.v);
		return r;
	}
	inline cdouble fmax(cdouble a, cdouble b)
	{
		bump(OP_DCMP);
		cdouble r;
		r.v = std::fmax(a.v, b.v);
		return r;
	}
}

// the restatements call libm both qualified (std::pow) and unqualified (sqrtf, powf, ...): give both spellings to the counted types
namespace std
{
	using upc::abs;
	using upc::acos;
	using upc::asin;
	using upc::atan;
	using upc::atan2;
	using upc::cbrt;
	using upc::ceil;
	using upc::cos;
	using upc::exp;
	using upc::fabs;
	using upc::floor;
	using upc::fmax;
	using upc::fmin;
	using upc::fmod;
	using upc::isfinite;
	using upc::isnan;
	using upc::log;
	using upc::pow;
	using upc::round;
	using upc::sin;
	using upc::sqrt;
	using upc::tan;
	using upc::trunc;
}
#define UPC_F(NAME, STD) \
	inline upc::cfloat NAME(upc::cfloat a) { return upc::STD(a); }
UPC_F(sqrtf, sqrt)
UPC_F(sinf, sin)
UPC_F(cosf, cos)
UPC_F(tanf, tan)
UPC_F(asinf, asin)
UPC_F(acosf, acos)
UPC_F(atanf, atan)
UPC_F(expf, exp)
UPC_F(logf, log)
UPC_F(cbrtf, cbrt)
UPC_F(fabsf, fabs)
UPC_F(floorf, floor)
UPC_F(ceilf, ceil)
UPC_F(roundf, round)
UPC_F(truncf, trunc)
#undef UPC_F
inline upc::cfloat powf(upc::cfloat a, upc::cfloat b) { return upc::pow(a, b); }
inline upc::cfloat atan2f(upc::cfloat a, upc::cfloat b) { return upc::atan2(a, b); }
inline upc::cfloat fmodf(upc::cfloat a, upc::cfloat b) { return upc::fmod(a, b); }
inline upc::cfloat fminf(upc::cfloat a, upc::cfloat b) { return upc::fmin(a, b); }
inline upc::cfloat fmaxf(upc::cfloat a, upc::cfloat b) { return upc::fmax(a, b); }
using upc::sqrt;
using upc::sin;
using upc::cos;
using upc::pow;
using upc::atan2;
using upc::acos;
using upc::cbrt;
using upc::exp;
using upc::log;
using upc::fabs;
using upc::floor;
using upc::fmin;
using upc::fmax;
