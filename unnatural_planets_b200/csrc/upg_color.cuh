// upg_color.cuh -- colour-space helpers of the texel-baking stage.
//
// Replaces, for the hot path, cage's colorRgbToHsv / colorHsvToRgb (float) and colorRgbToHsluv /
// colorHsluvToRgb / interpolateColor (externals/cage/sources/libcore/color.cpp:78-183, 263-283), the
// HSLuv reference implementation they call in DOUBLE precision (externals/cage/externals/hsluv/hsluv.c),
// and colorHueShift / colorDeviation of sources/math.cpp:40-52.
// Operation order follows the reference expression by expression; no FMA contraction (upg_common.cuh).
#pragma once

#include <cfloat>

#include "upg_common.cuh"

// the two double-precision HSLuv conversions are real functions on the device (four call sites, ~1 k instructions each)
#if defined(__CUDACC__)
#define UPG_COLOR_OUTLINED static __host__ __device__ __noinline__
#else
#define UPG_COLOR_OUTLINED inline
#endif

namespace upg
{
	// Real wrap(Real) (include/cage-core/math.h:810-816)
	UPG_HD float wrap_(float x)
	{
		if (x < 0)
			return 1 - fmod_(-x, 1.0f);
		return fmod_(x, 1.0f);
	}

	// interpolateWrap (libcore/math/scalars.cpp:189-200)
	UPG_HD float interpolate_wrap(float a, float b, float f)
	{
		a = wrap_(a);
		b = wrap_(b);
		if (abs_(a - b) <= 0.5f)
			return lerp_(a, b, f);
		if (a < b)
			a += 1;
		else
			b += 1;
		return wrap_(lerp_(a, b, f));
	}

	// colorRgbToHsv (color.cpp:78-104)
	UPG_HD V3 rgb_to_hsv(V3 in)
	{
		V3 out = v3(0, 0, 0);
		float mn = in.x < in.y ? in.x : in.y;
		mn = mn < in.z ? mn : in.z;
		float mx = in.x > in.y ? in.x : in.y;
		mx = mx > in.z ? mx : in.z;
		out.z = mx;
		const float delta = mx - mn;
		if (delta < 0.00001f)
			return out;
		if (mx > 0)
			out.y = delta / mx;
		else
			return out;
		if (in.x >= mx)
			out.x = (in.y - in.z) / delta;
		else if (in.y >= mx)
			out.x = 2 + (in.z - in.x) / delta;
		else
			out.x = 4 + (in.x - in.y) / delta;
		out.x *= 60.f / 360.f;
		if (out.x < 0)
			out.x += 1;
		return out;
	}

	// colorHsvToRgb (color.cpp:106-163)
	UPG_HD V3 hsv_to_rgb(V3 in)
	{
		if (in.y <= 0)
			return v3(in.z, in.z, in.z);
		float hh = in.x;
		if (hh >= 1)
			hh = 0;
		hh /= 60.f / 360.f;
		const uint32_t i = (uint32_t)hh;
		const float ff = hh - (float)i;
		const float p = in.z * (1 - in.y);
		const float q = in.z * (1 - (in.y * ff));
		const float t = in.z * (1 - (in.y * (1 - ff)));
		switch (i)
		{
			case 0:
				return v3(in.z, t, p);
			case 1:
				return v3(q, in.z, p);
			case 2:
				return v3(p, in.z, t);
			case 3:
				return v3(p, q, in.z);
			case 4:
				return v3(t, p, in.z);
			default:
				return v3(in.z, p, q);
		}
	}

	// colorHueShift (sources/math.cpp:47-52)
	UPG_HD V3 color_hue_shift(V3 rgb, float shift)
	{
		V3 hsv = rgb_to_hsv(rgb);
		hsv.x = fmod_(hsv.x + shift + 1, 1.0f);
		return hsv_to_rgb(hsv);
	}

	// ---- hsluv.c ------------------------------------------------------------------------------------
	// The reference evaluates HSLuv in DOUBLE (hsluv.c), and so does the default build: UPG_HSLUV_REAL = double.  -DUPG_HSLUV_REAL=float
	// is an experiment switch, NOT a supported build: the same expressions in single precision (sinf / cosf / powf / cbrtf / atan2f),
	// used by tools/hsluv_precision.py to count on the CPU how many texels of a bake an FP32 HSLuv would change.
#ifndef UPG_HSLUV_REAL
#define UPG_HSLUV_REAL double
#endif
	typedef UPG_HSLUV_REAL HR;
	UPG_HD double hr_sin(double x) { return sin(x); }
	UPG_HD double hr_cos(double x) { return cos(x); }
	UPG_HD double hr_pow(double x, double y) { return pow(x, y); }
	UPG_HD double hr_cbrt(double x) { return cbrt(x); }
	UPG_HD double hr_sqrt(double x) { return sqrt(x); }
	UPG_HD double hr_atan2(double y, double x) { return atan2(y, x); }
	UPG_HD float hr_sin(float x) { return sinf(x); }
	UPG_HD float hr_cos(float x) { return cosf(x); }
	UPG_HD float hr_pow(float x, float y) { return powf(x, y); }
	UPG_HD float hr_cbrt(float x) { return cbrtf(x); }
	UPG_HD float hr_sqrt(float x) { return sqrtf(x); }
	UPG_HD float hr_atan2(float y, float x) { return atan2f(y, x); }
	UPG_HD double hr_max(double) { return DBL_MAX; }
	UPG_HD float hr_max(float) { return FLT_MAX; }
	struct D3
	{
		HR a, b, c;
	};

	namespace hsluv
	{
		struct Bounds
		{
			HR a, b;
		};

		UPG_HD void get_bounds(HR l, Bounds *bounds)
		{
			const HR m[3][3] = {
				{ HR(3.24096994190452134377), HR(-1.53738317757009345794), HR(-0.49861076029300328366) },
				{ HR(-0.96924363628087982613), HR(1.87596750150772066772), HR(0.04155505740717561247) },
				{ HR(0.05563007969699360846), HR(-0.20397695888897656435), HR(1.05697151424287856072) },
			};
			const HR kappa = HR(903.29629629629629629630);
			const HR epsilon = HR(0.00885645167903563082);
			const HR tl = l + HR(16.0);
			const HR sub1 = (tl * tl * tl) / HR(1560896.0);
			const HR sub2 = (sub1 > epsilon ? sub1 : (l / kappa));
			for (int channel = 0; channel < 3; channel++)
			{
				const HR m1 = m[channel][0], m2 = m[channel][1], m3 = m[channel][2];
				for (int t = 0; t < 2; t++)
				{
					const HR top1 = (HR(284517.0) * m1 - HR(94839.0) * m3) * sub2;
					const HR top2 = (HR(838422.0) * m3 + HR(769860.0) * m2 + HR(731718.0) * m1) * l * sub2 - HR(769860.0) * t * l;
					const HR bottom = (HR(632260.0) * m3 - HR(126452.0) * m2) * sub2 + HR(126452.0) * t;
					bounds[channel * 2 + t].a = top1 / bottom;
					bounds[channel * 2 + t].b = top2 / bottom;
				}
			}
		}

		// `sn` / `cs` hand sin / cos of the hue back to hsluv2rgb, whose lch2luv step needs the same two values (hsluv.c computes them twice)
		UPG_HD HR max_chroma_for_lh(HR l, HR h, HR &sn, HR &cs)
		{
			HR minLen = hr_max(HR(0));
			const HR hrad = h * HR(0.01745329251994329577);
			Bounds bounds[6];
			get_bounds(l, bounds);
			sn = hr_sin(hrad);
			cs = hr_cos(hrad);
			for (int i = 0; i < 6; i++)
			{
				const HR len = bounds[i].b / (sn - bounds[i].a * cs);
				if (len >= 0 && len < minLen)
					minLen = len;
			}
			return minLen;
		}

		UPG_HD HR from_linear(HR c)
		{
			if (c <= HR(0.0031308))
				return HR(12.92) * c;
			return HR(1.055) * hr_pow(c, HR(1.0) / HR(2.4)) - HR(0.055);
		}

		UPG_HD HR to_linear(HR c)
		{
			if (c > HR(0.04045))
				return hr_pow((c + HR(0.055)) / HR(1.055), HR(2.4));
			return c / HR(12.92);
		}

		UPG_COLOR_OUTLINED D3 rgb2hsluv(HR r, HR g, HR b)
		{
			const HR kappa = HR(903.29629629629629629630);
			const HR epsilon = HR(0.00885645167903563082);
			const HR refU = HR(0.19783000664283680764), refV = HR(0.46831999493879100370);
			// rgb2xyz
			const HR lr = to_linear(r), lg = to_linear(g), lb = to_linear(b);
			const HR X = HR(0.41239079926595948129) * lr + HR(0.35758433938387796373) * lg + HR(0.18048078840183428751) * lb;
			const HR Y = HR(0.21263900587151035754) * lr + HR(0.71516867876775592746) * lg + HR(0.07219231536073371500) * lb;
			const HR Z = HR(0.01933081871559185069) * lr + HR(0.11919477979462598791) * lg + HR(0.95053215224966058086) * lb;
			// xyz2luv
			const HR varU = (HR(4.0) * X) / (X + (HR(15.0) * Y) + (HR(3.0) * Z));
			const HR varV = (HR(9.0) * Y) / (X + (HR(15.0) * Y) + (HR(3.0) * Z));
			const HR L = Y <= epsilon ? Y * kappa : HR(116.0) * hr_cbrt(Y) - HR(16.0);
			HR U = HR(13.0) * L * (varU - refU);
			HR V = HR(13.0) * L * (varV - refV);
			if (L < HR(0.00000001))
			{
				U = HR(0.0);
				V = HR(0.0);
			}
			// luv2lch
			const HR C = hr_sqrt(U * U + V * V);
			HR H;
			if (C < HR(0.00000001))
				H = 0;
			else
			{
				H = hr_atan2(V, U) * HR(57.29577951308232087680);
				if (H < HR(0.0))
					H += HR(360.0);
			}
			// lch2hsluv
			HR S;
			if (L > HR(99.9999999) || L < HR(0.00000001))
				S = HR(0.0);
			else
			{
				HR sn, cs;
				S = C / max_chroma_for_lh(L, H, sn, cs) * HR(100.0);
			}
			if (C < HR(0.00000001))
				H = HR(0.0);
			return D3{ H, S, L };
		}

		UPG_COLOR_OUTLINED D3 hsluv2rgb(HR h, HR s, HR l)
		{
			const HR kappa = HR(903.29629629629629629630);
			const HR refU = HR(0.19783000664283680764), refV = HR(0.46831999493879100370);
			// hsluv2lch
			HR c, sn = HR(0.0), cs = HR(1.0);
			bool trig = false; // sn / cs hold sin / cos of h * pi / 180
			if (l > HR(99.9999999) || l < HR(0.00000001))
				c = HR(0.0);
			else
			{
				c = max_chroma_for_lh(l, h, sn, cs) / HR(100.0) * s;
				trig = true;
			}
			if (s < HR(0.00000001))
			{
				h = HR(0.0);
				sn = HR(0.0); // hr_sin(HR(0.0)), hr_cos(HR(0.0)): exact in every libm
				cs = HR(1.0);
				trig = true;
			}
			// lch2luv
			if (!trig)
			{
				const HR hrad = h * HR(0.01745329251994329577);
				sn = hr_sin(hrad);
				cs = hr_cos(hrad);
			}
			const HR u = cs * c;
			const HR v = sn * c;
			// luv2xyz
			HR X, Y, Z;
			if (l <= HR(0.00000001))
			{
				X = Y = Z = HR(0.0);
			}
			else
			{
				const HR varU = u / (HR(13.0) * l) + refU;
				const HR varV = v / (HR(13.0) * l) + refV;
				if (l <= HR(8.0))
					Y = l / kappa;
				else
				{
					const HR x = (l + HR(16.0)) / HR(116.0);
					Y = (x * x * x);
				}
				X = -(HR(9.0) * Y * varU) / ((varU - HR(4.0)) * varV - varU * varV);
				Z = (HR(9.0) * Y - (HR(15.0) * varV * Y) - (varV * X)) / (HR(3.0) * varV);
			}
			// xyz2rgb
			const HR r = from_linear(HR(3.24096994190452134377) * X + HR(-1.53738317757009345794) * Y + HR(-0.49861076029300328366) * Z);
			const HR g = from_linear(HR(-0.96924363628087982613) * X + HR(1.87596750150772066772) * Y + HR(0.04155505740717561247) * Z);
			const HR b = from_linear(HR(0.05563007969699360846) * X + HR(-0.20397695888897656435) * Y + HR(1.05697151424287856072) * Z);
			return D3{ r, g, b };
		}
	}

	// colorRgbToHsluv (color.cpp:165-174): h / 360, s / 100, l / 100 in double, narrowed to Real
	UPG_HD V3 rgb_to_hsluv(V3 rgb)
	{
		const D3 r = hsluv::rgb2hsluv((HR)rgb.x, (HR)rgb.y, (HR)rgb.z);
		return v3((float)(r.a / 360), (float)(r.b / 100), (float)(r.c / 100));
	}

	// colorHsluvToRgb (color.cpp:176-183): the products are FLOAT products widened afterwards; result saturated
	UPG_HD V3 hsluv_to_rgb(V3 hsl)
	{
		const HR h = (HR)(hsl.x * 360.0f), s = (HR)(hsl.y * 100.0f), l = (HR)(hsl.z * 100.0f);
		const D3 r = hsluv::hsluv2rgb(h, s, l);
		return v3(saturate_((float)r.a), saturate_((float)r.b), saturate_((float)r.c));
	}

	// interpolateColor (color.cpp:263-283) with the HSLuv form of either end point possibly known beforehand
	UPG_HD V3 interpolate_hsluv(V3 hsluv1, V3 hsluv2, float factor)
	{
		const float h = interpolate_wrap(hsluv1.x, hsluv2.x, factor);
		const float s = lerp_(hsluv1.y, hsluv2.y, factor);
		const float l = lerp_(hsluv1.z, hsluv2.z, factor);
		return hsluv_to_rgb(v3(h, s, l));
	}
	UPG_HD V3 interpolate_color(V3 rgb1, V3 rgb2, float factor) { return interpolate_hsluv(rgb_to_hsluv(rgb1), rgb_to_hsluv(rgb2), factor); }

	// colorDeviation (sources/math.cpp:40-45) with the three randomChance() draws supplied by the caller
	UPG_HD V3 color_deviation(V3 color, float deviation, V3 chance)
	{
		V3 hsl = rgb_to_hsluv(color) + (chance - 0.5f) * deviation;
		hsl.x = fmod_(hsl.x + 1, 1.0f);
		return hsluv_to_rgb(v3(saturate_(hsl.x), saturate_(hsl.y), saturate_(hsl.z)));
	}
}
