// Parameters, spectral tables, distributions, RNG helpers, transformables, mesh container.
#include "prh.h"
#include "tables_generated.inc"

#include <cstdio>
#include <fstream>

namespace PR {
// ---------------------------------------------------------------- logging
int& logVerbosity()
{
	static int v = L_WARNING;
	return v;
}
namespace {
struct NullBuf : std::streambuf {
	int overflow(int c) override { return c; }
};
} // namespace
std::ostream& logStream(LogLevel lvl)
{
	static NullBuf nb;
	static std::ostream null(&nb);
	if ((int)lvl < logVerbosity())
		return null;
	static const char* names[] = { "DEBUG", "INFO", "WARNING", "ERROR", "FATAL" };
	std::cerr << "[prh " << names[lvl] << "] ";
	return std::cerr;
}

// ---------------------------------------------------------------- Parameter
Parameter Parameter::fromBool(bool v)
{
	Parameter p;
	p.mType = ParameterType::Bool;
	p.mInts = { v ? 1 : 0 };
	return p;
}
Parameter Parameter::fromInt(int64 v)
{
	Parameter p;
	p.mType = ParameterType::Int;
	p.mInts = { v };
	return p;
}
Parameter Parameter::fromUInt(uint64 v)
{
	Parameter p;
	p.mType = ParameterType::UInt;
	p.mInts = { (int64)v };
	return p;
}
Parameter Parameter::fromNumber(float v)
{
	Parameter p;
	p.mType	   = ParameterType::Number;
	p.mNumbers = { v };
	return p;
}
Parameter Parameter::fromString(const std::string& v)
{
	Parameter p;
	p.mType	   = ParameterType::String;
	p.mStrings = { v };
	return p;
}
Parameter Parameter::fromReference(uint32 id)
{
	Parameter p;
	p.mType = ParameterType::Reference;
	p.mInts = { (int64)id };
	return p;
}
Parameter Parameter::fromBoolArray(const std::vector<bool>& v)
{
	Parameter p;
	p.mType	   = ParameterType::Bool;
	p.mIsArray = true;
	for (bool b : v)
		p.mInts.push_back(b ? 1 : 0);
	return p;
}
Parameter Parameter::fromIntArray(const std::vector<int64>& v)
{
	Parameter p;
	p.mType	   = ParameterType::Int;
	p.mIsArray = true;
	p.mInts	   = v;
	return p;
}
Parameter Parameter::fromNumberArray(const std::vector<float>& v)
{
	Parameter p;
	p.mType	   = ParameterType::Number;
	p.mIsArray = true;
	p.mNumbers = v;
	return p;
}
Parameter Parameter::fromStringArray(const std::vector<std::string>& v)
{
	Parameter p;
	p.mType	   = ParameterType::String;
	p.mIsArray = true;
	p.mStrings = v;
	return p;
}
size_t Parameter::arraySize() const
{
	switch (mType) {
	case ParameterType::Number: return mNumbers.size();
	case ParameterType::String: return mStrings.size();
	case ParameterType::Invalid: return 0;
	default: return mInts.size();
	}
}
bool Parameter::getBool(bool def) const { return (mType == ParameterType::Bool && !mInts.empty()) ? mInts[0] != 0 : def; }
int64 Parameter::getInt(int64 def) const
{
	if ((mType == ParameterType::Int || mType == ParameterType::UInt) && !mInts.empty())
		return mInts[0];
	return def;
}
uint64 Parameter::getUInt(uint64 def) const
{
	if ((mType == ParameterType::Int || mType == ParameterType::UInt) && !mInts.empty() && mInts[0] >= 0)
		return (uint64)mInts[0];
	return def;
}
float Parameter::getNumber(float def) const { return getNumber(0, def); }
float Parameter::getNumber(size_t idx, float def) const
{
	if (mType == ParameterType::Number && idx < mNumbers.size())
		return mNumbers[idx];
	if ((mType == ParameterType::Int || mType == ParameterType::UInt) && idx < mInts.size())
		return (float)mInts[idx];
	return def;
}
std::string Parameter::getString(const std::string& def) const { return getString(0, def); }
std::string Parameter::getString(size_t idx, const std::string& def) const
{
	return (mType == ParameterType::String && idx < mStrings.size()) ? mStrings[idx] : def;
}
Vector3f ParameterGroup::getVector3f(const std::string& n, const Vector3f& def) const
{
	const Parameter p = getParameter(n);
	if (p.isArray() && p.arraySize() == 3 && (p.type() == ParameterType::Number || p.type() == ParameterType::Int))
		return Vector3f(p.getNumber(0, def.x), p.getNumber(1, def.y), p.getNumber(2, def.z));
	return def;
}

// ---------------------------------------------------------------- CIE / spectra
float equidistantLookup(const float* data, size_t count, float start, float end, float wavelength)
{
	const float delta = (end - start) / (count - 1);
	const float af	  = std::max(0.0f, (wavelength - start) / delta);
	const int index	  = (int)std::min<float>((float)(count - 2), af);
	const float t	  = std::min<float>((float)(count - 1), af) - index;
	return data[index] * (1 - t) + data[index + 1] * t;
}
namespace CIE {
const float* table(int c) { return c == 0 ? PRH_CIE2006_X : (c == 1 ? PRH_CIE2006_Y : PRH_CIE2006_Z); }
static float evalc(int c, float w)
{
	return equidistantLookup(table(c), PR_CIE_SAMPLE_COUNT, PR_CIE_WAVELENGTH_START, PR_CIE_WAVELENGTH_END, w) / PR_CIE_Y_NORM * PR_CIE_WAVELENGTH_RANGE;
}
float eval_x(float w) { return evalc(0, w); }
float eval_y(float w) { return evalc(1, w); }
float eval_z(float w) { return evalc(2, w); }
} // namespace CIE

const float* illuminantTable(const std::string& lname, size_t& count, float& start, float& end)
{
	struct E {
		const char* n;
		const float* d;
		bool f;
	};
	static const E tab[] = {
		{ "d65", PRH_ILLUM_D65, false }, { "d50", PRH_ILLUM_D50, false }, { "d55", PRH_ILLUM_D55, false }, { "d75", PRH_ILLUM_D75, false },
		{ "a", PRH_ILLUM_A, false }, { "c", PRH_ILLUM_C, false }, { "f1", PRH_ILLUM_F1, true }, { "f2", PRH_ILLUM_F2, true },
		{ "f3", PRH_ILLUM_F3, true }, { "f4", PRH_ILLUM_F4, true }, { "f5", PRH_ILLUM_F5, true }, { "f6", PRH_ILLUM_F6, true },
		{ "f7", PRH_ILLUM_F7, true }, { "f8", PRH_ILLUM_F8, true }, { "f9", PRH_ILLUM_F9, true }, { "f10", PRH_ILLUM_F10, true },
		{ "f11", PRH_ILLUM_F11, true }, { "f12", PRH_ILLUM_F12, true }
	};
	for (const E& e : tab) {
		if (lname == e.n) {
			count = e.f ? 81 : 107; // IlluminantData.inl:4-10
			start = e.f ? 380.0f : 300.0f;
			end	  = e.f ? 780.0f : 830.0f;
			return e.d;
		}
	}
	return nullptr;
}

// ---------------------------------------------------------------- Distribution1D
void Distribution1D::generate(const std::function<float(size_t)>& f, float* sum)
{
	mCDF[0]		   = 0.0f;
	const size_t n = numberOfValues();
	for (size_t i = 0; i < n; ++i)
		mCDF[i + 1] = mCDF[i] + f(i);
	const float intr = mCDF[n];
	if (sum)
		*sum = intr;
	if (intr <= PR_EPSILON) {
		for (size_t i = 1; i < n + 1; ++i)
			mCDF[i] = float(i) / float(n);
	} else {
		for (size_t i = 1; i < n + 1; ++i)
			mCDF[i] /= intr;
	}
	mCDF[n] = 1.0f;
}
static int interval_binary_search(int size, const std::function<bool(int)>& pred)
{ // reference src/base/container/Interval.h
	int first = 0, len = size;
	while (len > 0) {
		const int half = len / 2, middle = first + half;
		if (pred(middle)) {
			first = middle + 1;
			len -= half + 1;
		} else {
			len = half;
		}
	}
	return std::max(0, std::min(first - 1, size - 2));
}
size_t Distribution1D::sampleDiscrete(float u, float& pdf, float* rem) const
{
	const size_t off = interval_binary_search((int)mCDF.size(), [&](int i) { return mCDF[i] <= u; });
	if (rem) {
		*rem		  = u - mCDF[off];
		const float k = mCDF[off + 1] - mCDF[off];
		if (k > PR_EPSILON)
			*rem /= k;
	}
	pdf = discretePdf(off);
	return off;
}
float Distribution1D::sampleContinuous(float u, float& pdf) const
{
	float rem;
	const size_t off = sampleDiscrete(u, pdf, &rem);
	pdf *= (mCDF.size() - 1);
	return (off + rem) / (mCDF.size() - 1);
}

// ---------------------------------------------------------------- SpectralUpsampler
SpectralUpsampler::SpectralUpsampler(const std::string& file)
{
	std::ifstream f(file, std::ios::binary);
	if (!f)
		throw std::runtime_error("Could not open spectral coefficient file " + file);
	char magic[8];
	uint32 hdr[2];
	f.read(magic, 8);
	f.read(reinterpret_cast<char*>(hdr), 8);
	if (std::memcmp(magic, "PRB2SPEC", 8) != 0 || hdr[1] != 3)
		throw std::runtime_error("Given spectral coefficients file is invalid");
	mRes = hdr[0];
	mScale.resize(mRes);
	mData.resize((size_t)mRes * mRes * mRes * 3 * 3);
	f.read(reinterpret_cast<char*>(mScale.data()), mScale.size() * sizeof(float));
	f.read(reinterpret_cast<char*>(mData.data()), mData.size() * sizeof(float));
	if (!f)
		throw std::runtime_error("Spectral coefficient file truncated");
}
static int find_interval(const float* values, int size_, float x)
{
	int left = 0, last_interval = size_ - 2, size = last_interval;
	while (size > 0) {
		const int half = size >> 1, middle = left + half + 1;
		if (values[middle] < x) {
			left = middle;
			size -= half + 1;
		} else {
			size = half;
		}
	}
	return std::min(left, last_interval);
}
void SpectralUpsampler::prepare(const float* r, const float* g, const float* b, float* out_a, float* out_b, float* out_c, size_t elems) const
{
	constexpr float EPS = 0.0001f;
	for (size_t e = 0; e < elems; ++e) {
		const float rgb[3] = { r[e], g[e], b[e] };
		float coeffs[3];
		if (rgb[0] <= EPS && rgb[1] <= EPS && rgb[2] <= EPS) {
			coeffs[0] = 0;
			coeffs[1] = 0;
			coeffs[2] = -500.0f;
		} else if (1 - rgb[0] <= EPS && 1 - rgb[1] <= EPS && 1 - rgb[2] <= EPS) {
			coeffs[0] = 0;
			coeffs[1] = 0;
			coeffs[2] = 5000000.0f;
		} else {
			const uint32 res = mRes;
			const uint32 dx = 3, dy = 3 * res, dz = 3 * res * res;
			int largest = 0;
			for (int j = 1; j < 3; ++j)
				if (rgb[largest] <= rgb[j])
					largest = j;
			const float z	  = rgb[largest];
			const float scale = (res - 1) / z;
			const float x	  = rgb[(largest + 1) % 3] * scale;
			const float y	  = rgb[(largest + 2) % 3] * scale;
			const uint32 xi	  = std::min((uint32)x, res - 2);
			const uint32 yi	  = std::min((uint32)y, res - 2);
			const uint32 zi	  = find_interval(mScale.data(), res, z);
			uint32 off		  = (((largest * res + zi) * res + yi) * res + xi) * 3;
			const float x1 = x - xi, x0 = 1.0f - x1, y1 = y - yi, y0 = 1.0f - y1;
			const float z1 = (z - mScale[zi]) / (mScale[zi + 1] - mScale[zi]), z0 = 1.0f - z1;
			const float* d = mData.data();
			for (int j = 0; j < 3; ++j) {
				coeffs[j] = ((d[off] * x0 + d[off + dx] * x1) * y0 + (d[off + dy] * x0 + d[off + dy + dx] * x1) * y1) * z0
							+ ((d[off + dz] * x0 + d[off + dz + dx] * x1) * y0 + (d[off + dz + dy] * x0 + d[off + dz + dy + dx] * x1) * y1) * z1;
				++off;
			}
		}
		out_a[e] = coeffs[0];
		out_b[e] = coeffs[1];
		out_c[e] = coeffs[2];
	}
}
void SpectralUpsampler::computeSingle(float a, float b, float c, const float* wavelengths, float* out_weights, size_t elems)
{
	for (size_t i = 0; i < elems; ++i) {
		const float x  = std::fma(std::fma(a, wavelengths[i], b), wavelengths[i], c);
		const float y  = 1.0f / std::sqrt(std::fma(x, x, 1.0f));
		out_weights[i] = std::fma(0.5f * x, y, 0.5f);
	}
}

// ---------------------------------------------------------------- Random helpers
void Random::advance(uint64 delta)
{ // MCG jump-ahead: state *= MULT^delta (pcg_random.hpp advance(), increment 0)
	uint64 acc = 1, cur = MULT;
	while (delta > 0) {
		if (delta & 1)
			acc *= cur;
		cur *= cur;
		delta >>= 1;
	}
	mState *= acc;
}
uint32 Random::get32(uint32 start, uint32 end)
{ // std::uniform_int_distribution<uint32>(start, end-1)(pcg32_fast): libstdc++ _S_nd<uint64_t>
	const uint32 urange = (end - 1) - start;
	if (urange == 0xFFFFFFFFu)
		return get32() + start;
	const uint32 uerange = urange + 1;
	uint64 product		 = (uint64)get32() * (uint64)uerange;
	uint32 low			 = (uint32)product;
	if (low < uerange) {
		const uint32 threshold = (0u - uerange) % uerange;
		while (low < threshold) {
			product = (uint64)get32() * (uint64)uerange;
			low		= (uint32)product;
		}
	}
	return (uint32)(product >> 32) + start;
}
uint64 Random::get64(uint64 start, uint64 end)
{ // uniform_int_distribution<uint64>(start,end-1) over a 64-bit URBG: _S_nd<unsigned __int128>
	const uint64 urange = (end - 1) - start;
	if (urange == ~0ULL)
		return get64() + start;
	const uint64 uerange	   = urange + 1;
	unsigned __int128 product = (unsigned __int128)get64() * (unsigned __int128)uerange;
	uint64 low				   = (uint64)product;
	if (low < uerange) {
		const uint64 threshold = (0ULL - uerange) % uerange;
		while (low < threshold) {
			product = (unsigned __int128)get64() * (unsigned __int128)uerange;
			low		= (uint64)product;
		}
	}
	return (uint64)(product >> 64) + start;
}
void libstdcxxShuffle(std::vector<uint32>& v, Random& rnd)
{ // libstdc++ 13 std::shuffle, URBG range 2^64-1 (Random::min()=0, max()=2^64-1, operator() = get64()):
  // since urngrange / urange >= urange the "two swaps per draw" path is taken.
	const size_t n = v.size();
	if (n < 2)
		return;
	size_t i = 1;
	if ((n % 2) == 0) {
		const uint64 k = rnd.get64(0, 2); // distr(0,1)
		std::swap(v[i], v[k]);
		++i;
	}
	while (i < n) {
		const uint64 swap_range = i + 1;
		// __gen_two_uniform_ints(swap_range, swap_range + 1, g): x = distr(0, b0*b1-1)(g); return (x / b1, x % b1)
		const uint64 b0 = swap_range, b1 = swap_range + 1;
		const uint64 x	= rnd.get64(0, b0 * b1);
		const uint64 p0 = x / b1, p1 = x % b1;
		std::swap(v[i], v[p0]);
		++i;
		std::swap(v[i], v[p1]);
		++i;
	}
}

// ---------------------------------------------------------------- ITransformable / MeshBase / RenderSettings
ITransformable::ITransformable(const std::string& name, const Transformf& t)
	: mName(name)
	, mTransform(t)
	, mInvTransformCache(t.inverse())
	, mNormalMatrixCache(t.linear().inverse().transpose())
	, mInvNormalMatrixCache(mNormalMatrixCache.inverse())
	, mJacobianDeterminant(std::abs(t.linear().determinant()))
{
}

static float triArea(const Vector3f& a, const Vector3f& b, const Vector3f& c) { return 0.5f * (b - a).cross(c - a).norm(); }
float MeshBase::faceArea(size_t f) const
{
	const Vector3f v0 = vertex(indices[4 * f]), v1 = vertex(indices[4 * f + 1]), v2 = vertex(indices[4 * f + 2]);
	if (isQuad(f)) { // Quad::surfaceArea, src/core/geometry/Quad.h: 0.5 * |(p3-p1) x (p4-p2)|
		const Vector3f v3 = vertex(indices[4 * f + 3]);
		return 0.5f * (v2 - v0).cross(v3 - v1).norm();
	}
	return triArea(v0, v1, v2);
}
float MeshBase::surfaceArea(const Transformf& t) const
{
	float a = 0;
	for (size_t f = 0; f < faceCount(); ++f) {
		const Vector3f v0 = t * vertex(indices[4 * f]), v1 = t * vertex(indices[4 * f + 1]), v2 = t * vertex(indices[4 * f + 2]);
		if (isQuad(f))
			a += 0.5f * (v2 - v0).cross(t * vertex(indices[4 * f + 3]) - v1).norm();
		else
			a += triArea(v0, v1, v2);
	}
	return a;
}
BoundingBox MeshBase::constructBoundingBox() const
{
	BoundingBox b;
	for (size_t i = 0; i < vertexCount(); ++i)
		b.combine(vertex((uint32)i));
	return b;
}
bool MeshBase::isValid(std::string* err) const
{
	auto fail = [&](const char* m) {
		if (err)
			*err = m;
		return false;
	};
	if (faceCount() == 0)
		return fail("No faces given");
	if (vertices.empty() || vertices.size() % 3 != 0)
		return fail("Invalid vertex component");
	if (hasNormals() && normals.size() != vertices.size())
		return fail("Normal count does not match vertex count");
	if (hasUVs() && uvs.size() / 2 != vertexCount())
		return fail("UV count does not match vertex count");
	for (size_t f = 0; f < faceCount(); ++f)
		for (int k = 0; k < 4; ++k) {
			const uint32 i = indices[4 * f + k];
			if (k == 3 && i == PR_INVALID_ID)
				continue;
			if (i >= vertexCount())
				return fail("Face index out of range");
		}
	if (!materialSlots.empty() && materialSlots.size() != faceCount())
		return fail("Material slot count does not match face count");
	return true;
}

uint32 RenderSettings::maxSampleCount() const
{
	if (progressive)
		return 0;
	if (sampleCountOverride > 0)
		return sampleCountOverride;
	return aaSamplerFactory->requestedSampleCount() * lensSamplerFactory->requestedSampleCount()
		   * timeSamplerFactory->requestedSampleCount() * spectralSamplerFactory->requestedSampleCount();
}

std::vector<uint64> buildRenderRandomMap(uint64 seed, uint32 width, uint32 height, uint32 rngDelta)
{
	const size_t n = (size_t)width * height;
	std::vector<Random> r(n, Random(seed));
	// warm-up: pixel i = pixel i-1 advanced by rngDelta draws -> jump-ahead, bit-identical (SURVEY hard part 2)
	uint64 mulDelta = 1, cur = Random::MULT;
	for (uint64 d = rngDelta; d > 0; d >>= 1) {
		if (d & 1)
			mulDelta *= cur;
		cur *= cur;
	}
	for (size_t i = 1; i < n; ++i)
		r[i].setState(r[i - 1].state() * mulDelta);
	// permutation: std::swap(r[i], r[r[0].get32(1, n)])
	if (n > 1)
		for (size_t i = 1; i < n; ++i)
			std::swap(r[i], r[r[0].get32(1, (uint32)n)]);
	std::vector<uint64> out(n);
	for (size_t i = 0; i < n; ++i)
		out[i] = r[i].state();
	return out;
}

std::vector<RenderTile> buildTileMap(uint32 vx, uint32 vy, uint32 vw, uint32 vh, uint32 rtx, uint32 rty)
{ // RenderTileMap::init, ZOrder mode: tiles of ceil(w/rtx) x ceil(h/rty), enumerated along a Morton curve
	rtx				= std::max(1u, std::min(rtx, vw));
	rty				= std::max(1u, std::min(rty, vh));
	const uint32 tw = (vw + rtx - 1) / rtx, th = (vh + rty - 1) / rty;
	std::vector<RenderTile> tiles;
	auto compact = [](uint64 x) {
		x &= 0x5555555555555555ULL;
		x = (x ^ (x >> 1)) & 0x3333333333333333ULL;
		x = (x ^ (x >> 2)) & 0x0f0f0f0f0f0f0f0fULL;
		x = (x ^ (x >> 4)) & 0x00ff00ff00ff00ffULL;
		x = (x ^ (x >> 8)) & 0x0000ffff0000ffffULL;
		x = (x ^ (x >> 16)) & 0x00000000ffffffffULL;
		return (uint32)x;
	};
	uint32 side = 1;
	while (side < std::max(rtx, rty))
		side <<= 1;
	for (uint64 m = 0; m < (uint64)side * side; ++m) {
		const uint32 tx = compact(m), ty = compact(m >> 1);
		if (tx >= rtx || ty >= rty)
			continue;
		const uint32 sx = vx + tx * tw, sy = vy + ty * th;
		const uint32 ex = std::min(vx + vw, sx + tw), ey = std::min(vy + vh, sy + th);
		if (sx < ex && sy < ey)
			tiles.push_back(RenderTile{ sx, sy, ex, ey });
	}
	return tiles;
}
} // namespace PR
