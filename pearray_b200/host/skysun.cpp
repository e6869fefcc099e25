// Sky and sun infinite lights: host-side precompute of everything the device samples.
//
//   sky   reference src/plugins/main/infinitelights/sky.cpp:27-252 + src/skysun/skysun/SkyModel.cpp:19-60:
//         the Hosek-Wilkie spectral sky-dome model (11 bands, 320..720 nm) baked into an [elevation][azimuth][band]
//         radiance table, plus the Distribution2D (core/sampler/Distribution2D.cpp) the light is importance sampled with.
//   sun   reference src/plugins/main/infinitelights/sun.cpp:20-311 + src/skysun/skysun/SunRadiance.cpp:
//         Preetham-style attenuated solar spectrum (64 samples, 360..760 nm) over a cone of SUN_VIS_RADIUS * radius.
//   sun position: src/skysun/skysun/SunLocation.cpp (Blanco-Muriel et al. 2001).
//
// The Hosek-Wilkie model ("An Analytic Model for Full Spectral Sky-Dome Radiance", SIGGRAPH 2012, and the authors' public
// reference implementation v1.4a, which the reference vendors as src/skysun/skysun/model/ArHosekSkyModel.cpp) is
// restated here from its published form: per band a 9-coefficient configuration and a mean radiance, each a quintic
// Bezier in cbrt(solar elevation / (pi/2)), bilinear in (albedo, turbidity); the coefficient tables are data imported by
// tools/import_reference_data.py into data/hosek_spectral.bin.
//
// Everything here runs once per scene on the host; device and oracle only read the resulting tables from the pool, so
// host libm rounding does not enter GPU-vs-oracle parity.
#include "prh.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace PR {
std::string dataDirectory(); // loader.cpp

namespace {
constexpr float ELEVATION_RANGE = PR_PI * 0.5f; // skysun/ElevationAzimuth.h:6-7
constexpr float AZIMUTH_RANGE	= PR_PI * 2;
constexpr int AR_BANDS			= PRB_SKY_BANDS;
constexpr float AR_START		= PRB_SKY_BAND_START;
constexpr float AR_DELTA		= PRB_SKY_BAND_DELTA;
constexpr int RES_AZ			= 512; // SkySunConfig.h:11-12
constexpr int RES_EL			= 256;

struct ElevationAzimuth { // skysun/ElevationAzimuth.h:9-45 (up is +Z)
	float Elevation, Azimuth;
	float theta() const { return 0.5f * PR_PI - Elevation; }
	float phi() const { return Azimuth; }
	static ElevationAzimuth fromThetaPhi(float theta, float phi)
	{
		ElevationAzimuth ea{ 0.5f * PR_PI - theta, phi };
		if (ea.Azimuth < 0)
			ea.Azimuth += 2 * PR_PI;
		return ea;
	}
	static ElevationAzimuth fromDirection(const Vector3f& D)
	{ // Spherical::from_direction, src/base/math/Spherical.h:8-15
		const float x = (D.x == 0 && D.y == 0) ? 1e-5f : D.x;
		float phi	  = std::atan2(D.y, x);
		phi			  = phi < 0 ? phi + 2 * PR_PI : phi;
		return fromThetaPhi(std::acos(D.z), phi);
	}
	Vector3f toDirection() const
	{ // Spherical::cartesian(theta, phi)
		const float th = theta(), ph = phi();
		return Vector3f(std::sin(th) * std::cos(ph), std::sin(th) * std::sin(ph), std::cos(th));
	}
};

// ------------------------------------------------------------------ sun position
struct TimePoint { // SunLocation.h:9-16 (Saarbruecken 2020-05-06 12:00)
	int Year = 2020, Month = 5, Day = 6, Hour = 12, Minute = 0;
	float Seconds = 0.0f;
};
struct MapLocation { // SunLocation.h:18-22
	float Longitude = 6.9965744f, Latitude = 49.235422f, Timezone = 2;
};
ElevationAzimuth computeSunEA(const TimePoint& tp, const MapLocation& loc)
{ // SunLocation.cpp:10-102
	constexpr double EARTH_MEAN_RADIUS = 6371.01, ASTRONOMICAL_UNIT = 149597890;
	const double decHours = tp.Hour - loc.Timezone + (tp.Minute + tp.Seconds / 60.0) / 60.0;
	const int aux1		  = (tp.Month - 14) / 12;
	const int aux2		  = (1461 * (tp.Year + 4800 + aux1)) / 4 + (367 * (tp.Month - 2 - 12 * aux1)) / 12 - (3 * ((tp.Year + 4900 + aux1) / 100)) / 4 + tp.Day - 32075;
	const double julian	  = (double)aux2 - 0.5 + decHours / 24.0;
	const double elapsed  = julian - 2451545.0;

	const double omega		   = 2.1429 - 0.0010394594 * elapsed;
	const double meanLongitude = 4.8950630 + 0.017202791698 * elapsed;
	const double anomaly	   = 6.2400600 + 0.0172019699 * elapsed;
	const double eclLongitude  = meanLongitude + 0.03341607 * std::sin(anomaly) + 0.00034894 * std::sin(2 * anomaly) - 0.0001134 - 0.0000203 * std::sin(omega);
	const double eclObliquity  = 0.4090928 - 6.2140e-9 * elapsed + 0.0000396 * std::cos(omega);

	const double sinEcl	  = std::sin(eclLongitude);
	double dY			  = std::cos(eclObliquity) * sinEcl;
	double dX			  = std::cos(eclLongitude);
	double rightAscension = std::atan2(dY, dX);
	if (rightAscension < 0.0)
		rightAscension += 2 * PR_PI;
	const double declination = std::asin(std::sin(eclObliquity) * sinEcl);

	const double greenwich = 6.6974243242 + 0.0657098283 * elapsed + decHours;
	const double localMean = PR_DEG2RAD * ((float)((greenwich * 15 + loc.Longitude)));
	const double latitude  = PR_DEG2RAD * loc.Latitude;
	const double cosLat = std::cos(latitude), sinLat = std::sin(latitude);
	const double hourAngle	  = localMean - rightAscension;
	const double cosHourAngle = std::cos(hourAngle);
	double elevation		  = std::acos(cosLat * cosHourAngle * std::cos(declination) + std::sin(declination) * sinLat);
	dY						  = -std::sin(hourAngle);
	dX						  = std::tan(declination) * cosLat - sinLat * cosHourAngle;
	double azimuth			  = std::atan2(dY, dX);
	if (azimuth < 0.0)
		azimuth += 2 * PR_PI;
	elevation += (EARTH_MEAN_RADIUS / ASTRONOMICAL_UNIT) * std::sin(elevation); // parallax
	return ElevationAzimuth{ PR_PI / 2 - (float)elevation, (float)azimuth };
}
ElevationAzimuth computeSunEA(const ParameterGroup& params)
{ // SunLocation.cpp:104-127
	if (params.hasParameter("direction"))
		return ElevationAzimuth::fromDirection(params.getVector3f("direction", Vector3f(0, 0, 1)));
	if (params.hasParameter("theta"))
		return ElevationAzimuth::fromThetaPhi(params.getNumber("theta", 0), params.getNumber("phi", 0));
	if (params.hasParameter("elevation"))
		return ElevationAzimuth{ params.getNumber("elevation", 0), params.getNumber("azimuth", 0) };
	TimePoint tp;
	MapLocation loc;
	tp.Year		  = (int)params.getInt("year", tp.Year);
	tp.Month	  = (int)params.getInt("month", tp.Month);
	tp.Day		  = (int)params.getInt("day", tp.Day);
	tp.Hour		  = (int)params.getInt("hour", tp.Hour);
	tp.Minute	  = (int)params.getInt("minute", tp.Minute);
	tp.Seconds	  = params.getNumber("seconds", tp.Seconds);
	loc.Latitude  = params.getNumber("latitude", loc.Latitude);
	loc.Longitude = params.getNumber("longitude", loc.Longitude);
	loc.Timezone  = params.getNumber("timezone", loc.Timezone);
	return computeSunEA(tp, loc);
}

// ------------------------------------------------------------------ Hosek-Wilkie sky-dome model (spectral variant)
struct HosekData {
	std::vector<double> conf; // [band][2 albedo][10 turbidity][6 control points][9]
	std::vector<double> rad;  // [band][2][10][6]
};
const HosekData& hosekData()
{
	static HosekData data = [] {
		HosekData d;
		const std::string file = dataDirectory() + "/hosek_spectral.bin";
		std::ifstream f(file, std::ios::binary);
		char magic[8];
		uint32 hdr[3];
		if (!f.read(magic, 8) || std::memcmp(magic, "PRBHOSEK", 8) != 0 || !f.read(reinterpret_cast<char*>(hdr), 12) || hdr[0] != (uint32)AR_BANDS
			|| hdr[1] != 1080 || hdr[2] != 120)
			throw std::runtime_error("sky: cannot read " + file);
		d.conf.resize((size_t)AR_BANDS * 1080);
		d.rad.resize((size_t)AR_BANDS * 120);
		if (!f.read(reinterpret_cast<char*>(d.conf.data()), d.conf.size() * 8) || !f.read(reinterpret_cast<char*>(d.rad.data()), d.rad.size() * 8))
			throw std::runtime_error("sky: truncated " + file);
		return d;
	}();
	return data;
}
// quintic Bezier over the six control points (stride apart) at x in [0,1]
double bezier5(const double* c, int stride, double x)
{
	const double y = 1.0 - x;
	return std::pow(y, 5.0) * c[0] + 5.0 * std::pow(y, 4.0) * x * c[stride] + 10.0 * std::pow(y, 3.0) * std::pow(x, 2.0) * c[2 * stride]
		   + 10.0 * std::pow(y, 2.0) * std::pow(x, 3.0) * c[3 * stride] + 5.0 * y * std::pow(x, 4.0) * c[4 * stride] + std::pow(x, 5.0) * c[5 * stride];
}
struct HosekState {
	double config[AR_BANDS][9];
	double radiance[AR_BANDS];
	HosekState(double solarElevation, double turbidity, double albedo)
	{
		const HosekData& d = hosekData();
		const int iT	   = (int)turbidity;
		const double tRem  = turbidity - (double)iT;
		const double x	   = std::pow(solarElevation / (M_PI / 2.0), 1.0 / 3.0);
		for (int b = 0; b < AR_BANDS; ++b) {
			const double* cs = d.conf.data() + (size_t)b * 1080;
			const double* rs = d.rad.data() + (size_t)b * 120;
			// the four (albedo, turbidity) corners, weights as in the published implementation (turbidity 10 has no upper neighbour)
			const double w[4]  = { (1.0 - albedo) * (1.0 - tRem), albedo * (1.0 - tRem), (1.0 - albedo) * tRem, albedo * tRem };
			const int alb[4]   = { 0, 1, 0, 1 };
			const int turb[4]  = { iT - 1, iT - 1, iT, iT };
			const int nCorners = (iT == 10) ? 2 : 4;
			for (int i = 0; i < 9; ++i)
				config[b][i] = 0;
			radiance[b] = 0;
			for (int k = 0; k < nCorners; ++k) {
				const double* ce = cs + 9 * 6 * 10 * alb[k] + 9 * 6 * turb[k];
				for (int i = 0; i < 9; ++i)
					config[b][i] += w[k] * bezier5(ce + i, 9, x);
				radiance[b] += w[k] * bezier5(rs + 6 * 10 * alb[k] + 6 * turb[k], 1, x);
			}
		}
	}
	double internal(int b, double theta, double gamma) const
	{ // the model's F(theta, gamma)
		const double* c	  = config[b];
		const double expM = std::exp(c[4] * gamma);
		const double cg	  = std::cos(gamma);
		const double rayM = cg * cg;
		const double mieM = (1.0 + cg * cg) / std::pow((1.0 + c[8] * c[8] - 2.0 * c[8] * cg), 1.5);
		const double zen  = std::sqrt(std::cos(theta));
		return (1.0 + c[0] * std::exp(c[1] / (std::cos(theta) + 0.01))) * (c[2] + c[3] * expM + c[5] * rayM + c[6] * mieM + c[7] * zen);
	}
	double skyRadiance(double theta, double gamma, double wavelength) const
	{ // linear interpolation between the two neighbouring bands
		const int low = (int)((wavelength - 320.0) / 40.0);
		if (low < 0 || low >= AR_BANDS)
			return 0.0;
		const double interp = std::fmod((wavelength - 320.0) / 40.0, 1.0);
		const double vLow	= internal(low, theta, gamma) * radiance[low];
		if (interp < 1e-6)
			return vLow;
		double r = (1.0 - interp) * vLow;
		if (low + 1 < AR_BANDS)
			r += interp * internal(low + 1, theta, gamma) * radiance[low + 1];
		return r;
	}
};

class SkyModel { // skysun/SkyModel.cpp:19-60, SkyModel.h:19-24
public:
	SkyModel(const std::shared_ptr<FloatSpectralNode>& groundAlbedo, const ElevationAzimuth& sunEA, const ParameterGroup& params)
	{
		mAz = (int)params.getInt("azimuth_resolution", RES_AZ);
		mEl = (int)params.getInt("elevation_resolution", RES_EL);
		const float solarElevation = PR_PI / 2 - sunEA.Elevation;
		const float turbidity	   = params.getNumber("turbidity", 3.0f);
		const float sunSe = std::sin(solarElevation), sunCe = std::cos(solarElevation);
		mData.resize((size_t)mEl * mAz * AR_BANDS);
		for (int k = 0; k < AR_BANDS; ++k) {
			const float wavelength = AR_START + k * AR_DELTA;
			ShadingContext ctx;
			ctx.WavelengthNM   = SpectralBlob(wavelength);
			const float albedo = groundAlbedo->eval(ctx)[0];
			const HosekState state(solarElevation, turbidity, albedo);
			for (int y = 0; y < mEl; ++y) {
				const float theta = PR_PI / 2 - std::max(0.001f, ELEVATION_RANGE * y / (float)mEl);
				const float st = std::sin(theta), ct = std::cos(theta);
				for (int x = 0; x < mAz; ++x) {
					const float azimuth	 = AZIMUTH_RANGE * x / (float)mAz;
					const float cosGamma = ct * sunCe + st * sunSe * std::cos(azimuth - sunEA.Azimuth);
					const float gamma	 = std::acos(std::min(1.0f, std::max(-1.0f, cosGamma)));
					const float radiance = (float)state.skyRadiance(theta, gamma, wavelength + 0.005f);
					mData[((size_t)y * mAz + x) * AR_BANDS + k] = std::max(0.0f, radiance);
				}
			}
		}
	}
	int azimuthCount() const { return mAz; }
	int elevationCount() const { return mEl; }
	float radiance(int band, const ElevationAzimuth& ea) const
	{
		const int az = std::max(0, std::min<int>(mAz - 1, int(ea.Azimuth / AZIMUTH_RANGE * mAz)));
		const int el = std::max(0, std::min<int>(mEl - 1, int(ea.Elevation / ELEVATION_RANGE * mEl)));
		return mData[((size_t)el * mAz + az) * AR_BANDS + band];
	}
	const std::vector<float>& data() const { return mData; }

private:
	std::vector<float> mData;
	int mAz, mEl;
};

constexpr float GROUND_PENALTY = 0.001f; // sky.cpp:23

class SkyLight : public IInfiniteLight { // sky.cpp:25-173
public:
	SkyLight(const std::string& name, const Transformf& t, const SkyModel& model, bool extend, bool allowCompensation)
		: IInfiniteLight(name, t)
		, mModel(model)
		, mExtend(extend)
	{
		// buildDistribution, sky.cpp:134-166
		mW = mModel.azimuthCount();
		mH = extend ? 2 * mModel.elevationCount() : mModel.elevationCount();
		const SpectralBlob WVLS(560.0f, 540.0f, 400.0f, 600.0f);
		std::vector<float> integrals(mH, 0.0f);
		mConditional.assign(mH, Distribution1D(mW));
		for (int y = 0; y < mH; ++y) {
			mConditional[y].generate(
				[&](size_t x) {
					const float azimuth = AZIMUTH_RANGE * x / (float)mModel.azimuthCount();
					float elevation;
					if (mExtend)
						elevation = (2 * ELEVATION_RANGE) * (y / (float)(2 * mModel.elevationCount()) - 0.5f);
					else
						elevation = ELEVATION_RANGE * y / (float)mModel.elevationCount();
					const float f	= std::cos(elevation);
					const SpectralBlob rb = radiance(WVLS, ElevationAzimuth{ elevation, azimuth });
					const float val		= std::max(0.0f, f * std::max(std::max(rb[0], rb[1]), std::max(rb[2], rb[3])));
					return (mExtend && elevation < 0.0f) ? val * GROUND_PENALTY : val;
				},
				&integrals[y]);
		}
		mMarginal = Distribution1D(mH);
		mMarginal.generate([&](size_t y) { return integrals[y]; });
		if (allowCompensation)
			throw std::runtime_error("sky: ':compensation true' (MIS compensation, disabled by default in the reference) is not supported");
	}
	SpectralBlob power(const SpectralBlob& wvl) const override { return radiance(wvl, ElevationAzimuth::fromDirection(Vector3f(0, 0, 1))); }
	SpectralRange spectralRange() const override { return SpectralRange(); }
	void describe(prb_light& out, NodeEmitter& e) const override
	{
		out.type = PRB_LIGHT_SKY;
		for (int i = 0; i < 9; ++i) {
			out.normal_matrix[i]	 = normalMatrix().m[i];
			out.inv_normal_matrix[i] = invNormalMatrix().m[i];
		}
		std::vector<float>& pool = *e.pool;
		out.table_offset		 = (uint32)pool.size();
		out.table_count			 = (uint32)mModel.data().size();
		out.table_start			 = AR_START;
		out.table_end			 = AR_START + AR_BANDS * AR_DELTA;
		pool.insert(pool.end(), mModel.data().begin(), mModel.data().end());
		out.az_count	= (uint32)mModel.azimuthCount();
		out.el_count	= (uint32)mModel.elevationCount();
		out.dist_offset = (uint32)pool.size();
		out.dist_w		= (uint32)mW;
		out.dist_h		= (uint32)mH;
		out.sky_extend	= mExtend ? 1u : 0u;
		pool.insert(pool.end(), mMarginal.cdf().begin(), mMarginal.cdf().end());
		for (int y = 0; y < mH; ++y)
			pool.insert(pool.end(), mConditional[y].cdf().begin(), mConditional[y].cdf().end());
	}

private:
	SpectralBlob radiance(const SpectralBlob& wvls, const ElevationAzimuth& ea) const
	{ // sky.cpp:168-184
		SpectralBlob blob;
		for (int i = 0; i < 4; ++i) {
			const float af	= std::max(0.0f, (wvls[i] - AR_START) / AR_DELTA);
			const int index = (int)std::min<float>(AR_BANDS - 2, af);
			const float t	= std::min<float>(AR_BANDS - 1, af) - index;
			blob[i]			= mModel.radiance(index, ea) * (1 - t) + mModel.radiance(index + 1, ea) * t;
		}
		return blob;
	}
	SkyModel mModel;
	bool mExtend;
	int mW = 0, mH = 0;
	std::vector<Distribution1D> mConditional;
	Distribution1D mMarginal;
};

// ------------------------------------------------------------------ sun radiance (SunRadiance.cpp; Preetham et al. / "MI" tables)
float orderedLookup(const float* data, const float* wavelengths, int n, float wavelength)
{ // OrderedSpectrumView::lookup, src/core/spectral/OrderedSpectrum.inl:11-20 (Interval::binary_search)
	int first = 0, len = n;
	while (len > 0) {
		const int half = len / 2, middle = first + half;
		if (wavelengths[middle] <= wavelength) {
			first = middle + 1;
			len -= half + 1;
		} else {
			len = half;
		}
	}
	const int index = std::max(0, std::min(first - 1, n - 2));
	const float t	= std::max(0.0f, std::min(1.0f, (wavelength - wavelengths[index]) / (wavelengths[index + 1] - wavelengths[index])));
	return data[index] * (1 - t) + data[index + 1] * t;
}
const float k_oWavelengths[64] = { 300, 305, 310, 315, 320, 325, 330, 335, 340, 345, 350, 355, 445, 450, 455, 460, 465, 470, 475, 480, 485, 490,
								   495, 500, 505, 510, 515, 520, 525, 530, 535, 540, 545, 550, 555, 560, 565, 570, 575, 580, 585, 590, 595, 600,
								   605, 610, 620, 630, 640, 650, 660, 670, 680, 690, 700, 710, 720, 730, 740, 750, 760, 770, 780, 790 };
const float k_oAmplitudes[64]  = { 10.0f, 4.8f,	 2.7f,	1.35f, .8f,	  .380f, .160f, .075f, .04f,  .019f, .007f, .0f,   .003f, .003f, .004f, .006f,
								   .008f, .009f, .012f, .014f, .017f, .021f, .025f, .03f,  .035f, .04f,	 .045f, .048f, .057f, .063f, .07f,	.075f,
								   .08f,  .085f, .095f, .103f, .110f, .12f,	 .122f, .12f,  .118f, .115f, .12f,	.125f, .130f, .12f,	 .105f, .09f,
								   .079f, .067f, .057f, .048f, .036f, .028f, .023f, .018f, .014f, .011f, .010f, .009f, .007f, .004f, .0f,	0.0f };
const float k_gWavelengths[4]  = { 759, 760, 770, 771 };
const float k_gAmplitudes[4]   = { 0, 3.0f, 0.210f, 0 };
const float k_waWavelengths[13] = { 689, 690, 700, 710, 720, 730, 740, 750, 760, 770, 780, 790, 800 };
const float k_waAmplitudes[13]	= { 0, 0.160e-1f, 0.240e-1f, 0.125e-1f, 0.100e+1f, 0.870f, 0.610e-1f, 0.100e-2f, 0.100e-4f, 0.100e-4f, 0.600e-3f, 0.175e-1f, 0.360e-1f };
const float solWavelengths[38]	= { 380, 390, 400, 410, 420, 430, 440, 450, 460, 470, 480, 490, 500, 510, 520, 530, 540, 550, 560,
									570, 580, 590, 600, 610, 620, 630, 640, 650, 660, 670, 680, 690, 700, 710, 720, 730, 740, 750 };
const float solAmplitudes[38]	= { 16559.0f, 16233.7f, 21127.5f, 25888.2f, 25829.1f, 24232.3f, 26760.5f, 29658.3f, 30545.4f, 30057.5f,
									30663.7f, 28830.4f, 28712.1f, 27825.0f, 27100.6f, 27233.6f, 26361.3f, 25503.8f, 25060.2f, 25311.6f,
									25355.9f, 25134.2f, 24631.5f, 24173.2f, 23685.3f, 23212.1f, 22827.7f, 22339.8f, 21970.2f, 21526.7f,
									21097.9f, 20728.3f, 20240.4f, 19870.8f, 19427.2f, 19072.4f, 18628.9f, 18259.2f };
float computeSunRadiance(float wavelength, float theta, float turbidity)
{ // SunRadiance.cpp:77-116
	const float beta  = 0.04608365822050f * turbidity - 0.04586025928522f;
	const float m	  = 1.0f / (std::cos(theta) + 0.15f * std::pow(93.885f - theta / PR_PI * 180.0f, -1.253f)); // relative optical mass
	const float tauR  = std::exp(-m * 0.008735f * std::pow(wavelength / 1000.0f, -4.08));						  // Rayleigh
	const float alpha = 1.3f;
	const float tauA  = std::exp(-m * beta * std::pow(wavelength / 1000.0f, -alpha)); // aerosol
	const float lOzone = 0.35f;
	const float tauO   = std::exp(-m * orderedLookup(k_oAmplitudes, k_oWavelengths, 64, wavelength) * lOzone);
	const float kg	   = orderedLookup(k_gAmplitudes, k_gWavelengths, 4, wavelength);
	const float tauG   = std::exp(-1.41f * kg * m / std::pow(1 + 118.93f * kg * m, 0.45f));
	const float w	   = 2.0f;
	const float kwa	   = orderedLookup(k_waAmplitudes, k_waWavelengths, 13, wavelength);
	const float tauWA  = std::exp(-0.2385f * kwa * w * m / std::pow(1 + 20.07f * kwa * w * m, 0.45f));
	return std::max(0.0f, orderedLookup(solAmplitudes, solWavelengths, 38, wavelength) * tauR * tauA * tauO * tauG * tauWA);
}

constexpr float SUN_WAVELENGTH_START = 360.0f; // sun.cpp:21-24
constexpr float SUN_WAVELENGTH_END	 = 760.0f;
constexpr int SUN_WAVELENGTH_SAMPLES = 64;
constexpr float SUN_VIS_RADIUS		 = PR_DEG2RAD * 0.5358f * 0.5f;

void frameDuff(const Vector3f& N, Vector3f& Nx, Vector3f& Ny)
{ // Tangent::frame -> frame_duff + normalise, src/base/math/Tangent.h:50-73
	const float sign = std::copysign(1.0f, N.z);
	const float a	 = -1.0f / (sign + N.z);
	const float b	 = N.x * N.y * a;
	Nx				 = Vector3f(1.0f + sign * N.x * N.x * a, sign * b, -sign * N.x).normalized();
	Ny				 = Vector3f(b, sign + N.y * N.y * a, -N.y).normalized();
}

class SunLight : public IInfiniteLight { // sun.cpp:27-140 (cone) and :142-240 (delta)
public:
	SunLight(const std::string& name, const Transformf& t, const ElevationAzimuth& ea, float turbidity, float radius, float scale, bool delta)
		: IInfiniteLight(name, t)
		, mSpectrum(SUN_WAVELENGTH_SAMPLES)
		, mDelta(delta)
	{
		mDirection = (normalMatrix() * ea.toDirection()).normalized();
		frameDuff(mDirection, mDx, mDy);
		mCosTheta = delta ? 1.0f : std::cos(SUN_VIS_RADIUS * radius);
		mPDF	  = delta ? 1.0f : PR_INV_PI * 0.5f / (1 - mCosTheta); // Sampling::uniform_cone_pdf
		float factor;
		if (delta)
			factor = 2 * PR_PI * (1 - std::cos(SUN_VIS_RADIUS)) * scale; // solid angle of the real sun
		else
			factor = scale / (radius * radius); // compensate for different radii
		const float dl = (SUN_WAVELENGTH_END - SUN_WAVELENGTH_START) / (SUN_WAVELENGTH_SAMPLES - 1);
		for (int i = 0; i < SUN_WAVELENGTH_SAMPLES; ++i)
			mSpectrum[i] = computeSunRadiance(SUN_WAVELENGTH_START + i * dl, ea.theta(), turbidity) * factor;
	}
	bool hasDeltaDistribution() const override { return mDelta; }
	SpectralBlob power(const SpectralBlob& wvl) const override
	{
		SpectralBlob r;
		for (int i = 0; i < 4; ++i)
			r[i] = equidistantLookup(mSpectrum.data(), mSpectrum.size(), SUN_WAVELENGTH_START, SUN_WAVELENGTH_END, wvl[i]);
		return r;
	}
	SpectralRange spectralRange() const override { return SpectralRange(SUN_WAVELENGTH_START, SUN_WAVELENGTH_END); }
	void describe(prb_light& out, NodeEmitter& e) const override
	{
		out.type = mDelta ? PRB_LIGHT_SUN_DELTA : PRB_LIGHT_SUN;
		for (int i = 0; i < 9; ++i) {
			out.normal_matrix[i]	 = normalMatrix().m[i];
			out.inv_normal_matrix[i] = invNormalMatrix().m[i];
		}
		std::vector<float>& pool = *e.pool;
		out.table_offset		 = (uint32)pool.size();
		out.table_count			 = (uint32)mSpectrum.size();
		out.table_start			 = SUN_WAVELENGTH_START;
		out.table_end			 = SUN_WAVELENGTH_END;
		pool.insert(pool.end(), mSpectrum.begin(), mSpectrum.end());
		const Vector3f* v[3] = { &mDirection, &mDx, &mDy };
		float* dst[3]		 = { out.sun_dir, out.sun_dx, out.sun_dy };
		for (int k = 0; k < 3; ++k) {
			dst[k][0] = v[k]->x;
			dst[k][1] = v[k]->y;
			dst[k][2] = v[k]->z;
		}
		out.sun_cos_theta = mCosTheta;
		out.sun_pdf		  = mPDF;
	}

private:
	std::vector<float> mSpectrum;
	Vector3f mDirection, mDx, mDy;
	float mCosTheta = 1, mPDF = 1;
	bool mDelta;
};

class SkyLightFactory : public IInfiniteLightPlugin { // sky.cpp:186-249
public:
	std::shared_ptr<IInfiniteLight> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const ParameterGroup& params = ctx.parameters();
		const auto groundAlbedo		 = ctx.lookupSpectralNode("albedo", 0.15f);
		const ElevationAzimuth sunEA = computeSunEA(params);
		return std::make_shared<SkyLight>(params.getString("name", "__unknown"), ctx.transform(), SkyModel(groundAlbedo, sunEA, params),
										  params.getBool("extend", true), params.getBool("compensation", false));
	}
	const std::vector<std::string>& getNames() const override
	{
		static std::vector<std::string> names({ "sky" });
		return names;
	}
	std::string specification(const std::string&) const override
	{
		return "Sky Light (Hosek-Wilkie): albedo (spectral, 0.15), turbidity (3), extend (true), azimuth_resolution (512), elevation_resolution (256); "
			   "sun location: direction | theta, phi | elevation, azimuth | year, month, day, hour, minute, seconds, latitude, longitude, timezone";
	}
};
class SunLightFactory : public IInfiniteLightPlugin { // sun.cpp:242-308
public:
	std::shared_ptr<IInfiniteLight> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const ParameterGroup& params = ctx.parameters();
		const ElevationAzimuth sunEA = computeSunEA(params);
		const float radius			 = params.getNumber("radius", 1.0f);
		const float turbidity		 = params.getNumber("turbidity", 3.0f);
		const float scale			 = params.getNumber("power_scale", 1.0f);
		return std::make_shared<SunLight>(params.getString("name", "__unknown"), ctx.transform(), sunEA, turbidity, radius, scale, radius <= PR_EPSILON);
	}
	const std::vector<std::string>& getNames() const override
	{
		static std::vector<std::string> names({ "sun" });
		return names;
	}
	std::string specification(const std::string&) const override
	{
		return "Sun: radius (1), turbidity (3), power_scale (1); location: direction | theta, phi | elevation, azimuth | date/time + latitude/longitude/timezone";
	}
};
} // namespace

// test hooks (c_api.cpp): the restated model and sun position, compared against the reference's own C sources
// (oracle/_ref/libarhosek.so) and its documented default (SunLocation.h:7-8)
double hosekSkyRadiance(double solarElevation, double turbidity, double albedo, double theta, double gamma, double wavelength)
{
	return HosekState(solarElevation, turbidity, albedo).skyRadiance(theta, gamma, wavelength);
}
void sunElevationAzimuth(int year, int month, int day, int hour, int minute, float seconds, float latitude, float longitude, float timezone, float* elevation,
						 float* azimuth)
{
	TimePoint tp;
	tp.Year = year, tp.Month = month, tp.Day = day, tp.Hour = hour, tp.Minute = minute, tp.Seconds = seconds;
	MapLocation loc;
	loc.Latitude = latitude, loc.Longitude = longitude, loc.Timezone = timezone;
	const ElevationAzimuth ea = computeSunEA(tp, loc);
	*elevation				  = ea.Elevation;
	*azimuth				  = ea.Azimuth;
}
float sunRadiance(float wavelength, float theta, float turbidity) { return computeSunRadiance(wavelength, theta, turbidity); }

std::vector<std::shared_ptr<IPlugin>> createSkySunPlugins()
{
	return { std::make_shared<SkyLightFactory>(), std::make_shared<SunLightFactory>() };
}
} // namespace PR
