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

	// ---- hsluv.c, double precision ---------------------------------------------------------------
	struct D3
	{
		double a, b, c;
	};

	namespace hsluv
	{
		struct Bounds
		{
			double a, b;
		};

		UPG_HD void get_bounds(double l, Bounds *bounds)
		{
			const double m[3][3] = {
				{ 3.24096994190452134377, -1.53738317757009345794, -0.49861076029300328366 },
				{ -0.96924363628087982613, 1.87596750150772066772, 0.04155505740717561247 },
				{ 0.05563007969699360846, -0.20397695888897656435, 1.05697151424287856072 },
			};
			const double kappa = 903.29629629629629629630;
			const double epsilon = 0.00885645167903563082;
			const double tl = l + 16.0;
			const double sub1 = (tl * tl * tl) / 1560896.0;
			const double sub2 = (sub1 > epsilon ? sub1 : (l / kappa));
			for (int channel = 0; channel < 3; channel++)
			{
				const double m1 = m[channel][0], m2 = m[channel][1], m3 = m[channel][2];
				for (int t = 0; t < 2; t++)
				{
					const double top1 = (284517.0 * m1 - 94839.0 * m3) * sub2;
					const double top2 = (838422.0 * m3 + 769860.0 * m2 + 731718.0 * m1) * l * sub2 - 769860.0 * t * l;
					const double bottom = (632260.0 * m3 - 126452.0 * m2) * sub2 + 126452.0 * t;
					bounds[channel * 2 + t].a = top1 / bottom;
					bounds[channel * 2 + t].b = top2 / bottom;
				}
			}
		}

		// `sn` / `cs` hand sin / cos of the hue back to hsluv2rgb, whose lch2luv step needs the same two values (hsluv.c computes them twice)
		UPG_HD double max_chroma_for_lh(double l, double h, double &sn, double &cs)
		{
			double minLen = DBL_MAX;
			const double hrad = h * 0.01745329251994329577;
			Bounds bounds[6];
			get_bounds(l, bounds);
			sn = sin(hrad);
			cs = cos(hrad);
			for (int i = 0; i < 6; i++)
			{
				const double len = bounds[i].b / (sn - bounds[i].a * cs);
				if (len >= 0 && len < minLen)
					minLen = len;
			}
			return minLen;
		}

		UPG_HD double from_linear(double c)
		{
			if (c <= 0.0031308)
				return 12.92 * c;
			return 1.055 * pow(c, 1.0 / 2.4) - 0.055;
		}

		UPG_HD double to_linear(double c)
		{
			if (c > 0.04045)
				return pow((c + 0.055) / 1.055, 2.4);
			return c / 12.92;
		}

		UPG_COLOR_OUTLINED D3 rgb2hsluv(double r, double g, double b)
		{
			const double kappa = 903.29629629629629629630;
			const double epsilon = 0.00885645167903563082;
			const double refU = 0.19783000664283680764, refV = 0.46831999493879100370;
			// rgb2xyz
			const double lr = to_linear(r), lg = to_linear(g), lb = to_linear(b);
			const double X = 0.41239079926595948129 * lr + 0.35758433938387796373 * lg + 0.18048078840183428751 * lb;
			const double Y = 0.21263900587151035754 * lr + 0.71516867876775592746 * lg + 0.07219231536073371500 * lb;
			const double Z = 0.01933081871559185069 * lr + 0.11919477979462598791 * lg + 0.95053215224966058086 * lb;
			// xyz2luv
			const double varU = (4.0 * X) / (X + (15.0 * Y) + (3.0 * Z));
			const double varV = (9.0 * Y) / (X + (15.0 * Y) + (3.0 * Z));
			const double L = Y <= epsilon ? Y * kappa : 116.0 * cbrt(Y) - 16.0;
			double U = 13.0 * L * (varU - refU);
			double V = 13.0 * L * (varV - refV);
			if (L < 0.00000001)
			{
				U = 0.0;
				V = 0.0;
			}
			// luv2lch
			const double C = sqrt(U * U + V * V);
			double H;
			if (C < 0.00000001)
				H = 0;
			else
			{
				H = atan2(V, U) * 57.29577951308232087680;
				if (H < 0.0)
					H += 360.0;
			}
			// lch2hsluv
			double S;
			if (L > 99.9999999 || L < 0.00000001)
				S = 0.0;
			else
			{
				double sn, cs;
				S = C / max_chroma_for_lh(L, H, sn, cs) * 100.0;
			}
			if (C < 0.00000001)
				H = 0.0;
			return D3{ H, S, L };
		}

		UPG_COLOR_OUTLINED D3 hsluv2rgb(double h, double s, double l)
		{
			const double kappa = 903.29629629629629629630;
			const double refU = 0.19783000664283680764, refV = 0.46831999493879100370;
			// hsluv2lch
			double c, sn = 0.0, cs = 1.0;
			bool trig = false; // sn / cs hold sin / cos of h * pi / 180
			if (l > 99.9999999 || l < 0.00000001)
				c = 0.0;
			else
			{
				c = max_chroma_for_lh(l, h, sn, cs) / 100.0 * s;
				trig = true;
			}
			if (s < 0.00000001)
			{
				h = 0.0;
				sn = 0.0; // sin(0.0), cos(0.0): exact in every libm
				cs = 1.0;
				trig = true;
			}
			// lch2luv
			if (!trig)
			{
				const double hrad = h * 0.01745329251994329577;
				sn = sin(hrad);
				cs = cos(hrad);
			}
			const double u = cs * c;
			const double v = sn * c;
			// luv2xyz
			double X, Y, Z;
			if (l <= 0.00000001)
			{
				X = Y = Z = 0.0;
			}
			else
			{
				const double varU = u / (13.0 * l) + refU;
				const double varV = v / (13.0 * l) + refV;
				if (l <= 8.0)
					Y = l / kappa;
				else
				{
					const double x = (l + 16.0) / 116.0;
					Y = (x * x * x);
				}
				X = -(9.0 * Y * varU) / ((varU - 4.0) * varV - varU * varV);
				Z = (9.0 * Y - (15.0 * varV * Y) - (varV * X)) / (3.0 * varV);
			}
			// xyz2rgb
			const double r = from_linear(3.24096994190452134377 * X + -1.53738317757009345794 * Y + -0.49861076029300328366 * Z);
			const double g = from_linear(-0.96924363628087982613 * X + 1.87596750150772066772 * Y + 0.04155505740717561247 * Z);
			const double b = from_linear(0.05563007969699360846 * X + -0.20397695888897656435 * Y + 1.05697151424287856072 * Z);
			return D3{ r, g, b };
		}
	}

	// colorRgbToHsluv (color.cpp:165-174): h / 360, s / 100, l / 100 in double, narrowed to Real
	UPG_HD V3 rgb_to_hsluv(V3 rgb)
	{
		const D3 r = hsluv::rgb2hsluv((double)rgb.x, (double)rgb.y, (double)rgb.z);
		return v3((float)(r.a / 360), (float)(r.b / 100), (float)(r.c / 100));
	}

	// colorHsluvToRgb (color.cpp:176-183): the products are FLOAT products widened afterwards; result saturated
	UPG_HD V3 hsluv_to_rgb(V3 hsl)
	{
		const double h = (double)(hsl.x * 360.0f), s = (double)(hsl.y * 100.0f), l = (double)(hsl.z * 100.0f);
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
