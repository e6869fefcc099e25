// Shading-node plugins needed by the config scenes: constants, 'refl' / 'illum'
// (SpectralValueNode.cpp:16-47), 'illuminant' (IlluminantNode.cpp), 'lookup_index'
// (ReflectiveNode.cpp:225-245,387-398), 'smul' (SpectralMathNode.cpp:275-277), 'checkerboard'
// (CheckerboardNode.cpp:26-48), 'spectrum' (SpectralConstNode.cpp:12-33).
// Every node can be evaluated on the host (used for light power / SPD distribution set-up) and
// flattened into the device node table.
#include "prh.h"

namespace PR {
const float* illuminantTable(const std::string& lname, size_t& count, float& start, float& end);

uint32 NodeEmitter::add(const FloatSpectralNode* key, const prb_node& n)
{
	nodes.push_back(n);
	const uint32 id = (uint32)nodes.size() - 1;
	mCache[key]		= id;
	return id;
}
bool NodeEmitter::find(const FloatSpectralNode* key, uint32& id) const
{
	auto it = mCache.find(key);
	if (it == mCache.end())
		return false;
	id = it->second;
	return true;
}

static uint32 devFlags(uint32 nf)
{
	uint32 f = 0;
	if (nf & NF_SpectralVarying)
		f |= PRB_NODE_FLAG_SPECTRAL_VARYING;
	if (nf & NF_TextureVarying)
		f |= PRB_NODE_FLAG_TEXTURE_VARYING;
	return f;
}

namespace {
class ConstScalarNode : public FloatScalarNode { // loader/shader/ConstNode.cpp:7-22
public:
	explicit ConstScalarNode(float f)
		: FloatScalarNode(NF_Const)
		, mValue(f)
	{
	}
	float eval(const ShadingContext&) const override { return mValue; }
	bool isConst() const override { return true; }
	std::string dumpInformation() const override { return std::to_string(mValue); }

private:
	float mValue;
};

class ConstSpectralNode : public FloatSpectralNode { // ConstNode.cpp:24-44
public:
	explicit ConstSpectralNode(float f)
		: FloatSpectralNode(NF_Const)
		, mValue(f)
	{
	}
	SpectralBlob eval(const ShadingContext&) const override { return SpectralBlob(mValue); }
	uint32 emit(NodeEmitter& e) const override
	{
		uint32 id;
		if (e.find(this, id))
			return id;
		prb_node n{};
		n.type	= PRB_NODE_CONST;
		n.flags = 0;
		n.p[0]	= mValue;
		return e.add(this, n);
	}
	std::string dumpInformation() const override { return std::to_string(mValue); }

private:
	float mValue;
};

// SpectralUpsampler::compute (src/core/spectral/SpectralUpsampler.h:45-49)
inline SpectralBlob upsampleCompute(const float* p, const SpectralBlob& w)
{
	SpectralBlob r;
	for (int i = 0; i < 4; ++i) {
		const float x = (p[0] * w[i] + p[1]) * w[i] + p[2];
		r[i]		  = 0.5f * x * (1.0f / std::sqrt(x * x + 1.0f)) + 0.5f;
	}
	return r;
}

class ParametricSpectralNode : public FloatSpectralNode { // ConstNode.cpp:46-68 (+ scaled variant :70-92)
public:
	ParametricSpectralNode(float a, float b, float c, float power, bool scaled)
		: FloatSpectralNode(NF_SpectralVarying)
		, mPower(power)
		, mScaled(scaled)
	{
		mP[0] = a;
		mP[1] = b;
		mP[2] = c;
	}
	SpectralBlob eval(const ShadingContext& ctx) const override
	{
		const SpectralBlob v = upsampleCompute(mP, ctx.WavelengthNM);
		return mScaled ? v * mPower : v;
	}
	uint32 emit(NodeEmitter& e) const override
	{
		uint32 id;
		if (e.find(this, id))
			return id;
		prb_node n{};
		n.type	= mScaled ? PRB_NODE_PARAM_SCALED : PRB_NODE_PARAM;
		n.flags = devFlags(flags());
		n.p[0]	= mP[0];
		n.p[1]	= mP[1];
		n.p[2]	= mP[2];
		n.p[3]	= mPower;
		return e.add(this, n);
	}
	std::string dumpInformation() const override
	{
		return "[" + std::to_string(mP[0]) + "," + std::to_string(mP[1]) + "," + std::to_string(mP[2]) + "]x" + std::to_string(mPower);
	}

private:
	float mP[3];
	float mPower;
	bool mScaled;
};

class EquidistantSpectrumNode : public FloatSpectralNode { // src/core/shader/EquidistantSpectrumNode.h
public:
	EquidistantSpectrumNode(const std::vector<float>& data, float start, float end, const std::string& label)
		: FloatSpectralNode(NF_SpectralVarying)
		, mData(data)
		, mStart(start)
		, mEnd(end)
		, mLabel(label)
	{
	}
	SpectralBlob eval(const ShadingContext& ctx) const override
	{
		SpectralBlob r;
		for (int i = 0; i < 4; ++i)
			r[i] = equidistantLookup(mData.data(), mData.size(), mStart, mEnd, ctx.WavelengthNM[i]);
		return r;
	}
	SpectralRange spectralRange() const override { return SpectralRange(mStart, mEnd); }
	uint32 emit(NodeEmitter& e) const override
	{
		uint32 id;
		if (e.find(this, id))
			return id;
		prb_node n{};
		n.type	= PRB_NODE_TABLE;
		n.flags = devFlags(flags());
		n.a		= (uint32)e.pool->size();
		n.b		= (uint32)mData.size();
		n.p[0]	= mStart;
		n.p[1]	= mEnd;
		e.pool->insert(e.pool->end(), mData.begin(), mData.end());
		return e.add(this, n);
	}
	std::string dumpInformation() const override { return mLabel; }

private:
	std::vector<float> mData;
	float mStart, mEnd;
	std::string mLabel;
};

class SellmeierIndexNode : public FloatSpectralNode { // ReflectiveNode.cpp:104-140 + Scattering.h:219-242
public:
	SellmeierIndexNode(const std::vector<float>& bs, const std::vector<float>& cs)
		: FloatSpectralNode(NF_SpectralVarying)
		, mBs(bs)
		, mCs(cs)
	{
	}
	SpectralBlob eval(const ShadingContext& ctx) const override
	{
		SpectralBlob r;
		for (int k = 0; k < 4; ++k) {
			const float qm	= ctx.WavelengthNM[k] / 1000;
			const float qm2 = qm * qm;
			float value		= 1;
			for (size_t i = 0; i < mBs.size(); ++i)
				value += mBs[i] * qm2 / (qm2 - mCs[i]);
			r[k] = std::sqrt(value);
		}
		return r;
	}
	uint32 emit(NodeEmitter& e) const override
	{
		uint32 id;
		if (e.find(this, id))
			return id;
		prb_node n{};
		n.type	= PRB_NODE_SELLMEIER;
		n.flags = devFlags(flags());
		n.a		= (uint32)e.pool->size();
		n.b		= (uint32)mBs.size();
		e.pool->insert(e.pool->end(), mBs.begin(), mBs.end());
		e.pool->insert(e.pool->end(), mCs.begin(), mCs.end());
		return e.add(this, n);
	}
	std::string dumpInformation() const override { return "SellmeierIndex"; }

private:
	std::vector<float> mBs, mCs;
};

class MulSpectralMath : public FloatSpectralNode { // SpectralMathNode.cpp 'smul'
public:
	MulSpectralMath(const std::shared_ptr<FloatSpectralNode>& a, const std::shared_ptr<FloatSpectralNode>& b)
		: FloatSpectralNode((a->flags() | b->flags()) & ~(uint32)NF_Const)
		, mA(a)
		, mB(b)
	{
	}
	SpectralBlob eval(const ShadingContext& ctx) const override { return mA->eval(ctx) * mB->eval(ctx); }
	SpectralRange spectralRange() const override { return mA->spectralRange() + mB->spectralRange(); }
	uint32 emit(NodeEmitter& e) const override
	{
		uint32 id;
		if (e.find(this, id))
			return id;
		prb_node n{};
		n.type	= PRB_NODE_MUL;
		n.flags = devFlags(flags());
		n.a		= mA->emit(e);
		n.b		= mB->emit(e);
		return e.add(this, n);
	}
	std::string dumpInformation() const override { return "(" + mA->dumpInformation() + " * " + mB->dumpInformation() + ")"; }

private:
	std::shared_ptr<FloatSpectralNode> mA, mB;
};

class CheckerboardNode : public FloatSpectralNode { // CheckerboardNode.cpp:14-70
public:
	CheckerboardNode(const std::shared_ptr<FloatSpectralNode>& a, const std::shared_ptr<FloatSpectralNode>& b, float su, float sv, int mode)
		: FloatSpectralNode((a->flags() | b->flags() | NF_TextureVarying) & ~(uint32)NF_Const)
		, mA(a)
		, mB(b)
		, mSU(su)
		, mSV(sv)
		, mMode(mode)
	{
	}
	bool check(const Vector2f& uv) const
	{
		float u = uv.x, v = uv.y;
		if (mMode == 1) {
			u *= mSU;
			v *= mSU;
		} else if (mMode == 2) {
			u *= mSU;
			v *= mSV;
		}
		return ((int)std::floor(u) + (int)std::floor(v)) % 2 == 0;
	}
	SpectralBlob eval(const ShadingContext& ctx) const override { return check(ctx.UV) ? mB->eval(ctx) : mA->eval(ctx); }
	SpectralRange spectralRange() const override { return mA->spectralRange() + mB->spectralRange(); }
	uint32 emit(NodeEmitter& e) const override
	{
		uint32 id;
		if (e.find(this, id))
			return id;
		prb_node n{};
		n.type	= PRB_NODE_CHECKER;
		n.flags = devFlags(flags());
		n.a		= mA->emit(e);
		n.b		= mB->emit(e);
		n.p[0]	= mSU;
		n.p[1]	= mSV;
		n.p[2]	= (float)mMode;
		return e.add(this, n);
	}
	std::string dumpInformation() const override { return "CheckerboardNode (" + mA->dumpInformation() + ", " + mB->dumpInformation() + ")"; }

private:
	std::shared_ptr<FloatSpectralNode> mA, mB;
	float mSU, mSV;
	int mMode;
};
} // namespace

std::shared_ptr<FloatScalarNode> makeConstScalarNode(float f) { return std::make_shared<ConstScalarNode>(f); }
std::shared_ptr<FloatSpectralNode> makeConstSpectralNode(float f) { return std::make_shared<ConstSpectralNode>(f); }

namespace NodeUtils {
SpectralBlob average(const SpectralBlob& wvls, const FloatSpectralNode* node)
{
	ShadingContext sc;
	sc.WavelengthNM = wvls;
	if (!(node->flags() & NF_TextureVarying)) { // uniform over UV: the 32x32 average of equal values
		sc.UV = Vector2f(0, 0);
		return node->eval(sc);
	}
	constexpr int SX = 32, SY = 32;
	SpectralBlob sum = SpectralBlob::Zero();
	for (int y = 0; y < SY; ++y)
		for (int x = 0; x < SX; ++x) {
			sc.UV = Vector2f(x / float(SX), y / float(SY));
			sum += node->eval(sc);
		}
	return sum * (1.0f / (SX * SY));
}
} // namespace NodeUtils

namespace {
// ------------------------------------------------------------------ image textures
// NonParametricImageNode (reference src/loader/shader/ImageNode.cpp:93-182; the parametric variant is compiled out there by
// PARAMETRIC_FILE_WORKAROUND, TextureParser.cpp:17): an RGB lookup, RGBConverter::linearize for sRGB encoded files, then
// SpectralUpsampler::prepare + ::compute per lookup.  The lookup itself is OpenImageIO's TextureSystem::texture() in the
// reference; restated here (and on the device, csrc/dev_shade.cuh evalImageNode) without MIP levels and derivatives.
class ImageSpectralNode : public FloatSpectralNode {
public:
	ImageSpectralNode(ImageData img, int interp, int wrapS, int wrapT, const std::shared_ptr<SpectralUpsampler>& upsampler, const std::string& file)
		: FloatSpectralNode(NF_TextureVarying | NF_SpectralVarying)
		, mImg(std::move(img))
		, mInterp(interp)
		, mWrapS(wrapS)
		, mWrapT(wrapT)
		, mUpsampler(upsampler)
		, mFile(file)
	{
		if (mImg.channels == 1) { // grayscale: the same value in every channel
			std::vector<float> rgb(mImg.data.size() * 3);
			for (size_t i = 0; i < mImg.data.size(); ++i)
				rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = mImg.data[i];
			mImg.data.swap(rgb);
			mImg.channels = 3;
		}
	}
	static bool wrapTexel(int& i, int size, int mode)
	{
		if (i >= 0 && i < size)
			return true;
		switch (mode) {
		default: return false;
		case PRB_WRAP_CLAMP: i = i < 0 ? 0 : size - 1; return true;
		case PRB_WRAP_PERIODIC:
			i %= size;
			if (i < 0)
				i += size;
			return true;
		case PRB_WRAP_MIRROR: {
			const int period = 2 * size;
			i %= period;
			if (i < 0)
				i += period;
			if (i >= size)
				i = period - 1 - i;
			return true;
		}
		}
	}
	void fetch(int x, int y, float rgb[3]) const
	{
		rgb[0] = rgb[1] = rgb[2] = 0;
		if (!wrapTexel(x, (int)mImg.width, mWrapS) || !wrapTexel(y, (int)mImg.height, mWrapT))
			return;
		const float* t = mImg.data.data() + 3 * ((size_t)y * mImg.width + x);
		rgb[0] = t[0], rgb[1] = t[1], rgb[2] = t[2];
	}
	static void bspline(float f, float w[4])
	{
		const float one_f = 1.0f - f;
		w[0]			  = (one_f * one_f * one_f) / 6.0f;
		w[1]			  = 2.0f / 3.0f - 0.5f * f * f * (2.0f - f);
		w[2]			  = 2.0f / 3.0f - 0.5f * one_f * one_f * (2.0f - one_f);
		w[3]			  = (f * f * f) / 6.0f;
	}
	void lookup(float u, float v, float rgb[3]) const
	{
		const int w = (int)mImg.width, h = (int)mImg.height;
		const float x = u * (float)w - 0.5f, y = (1 - v) * (float)h - 0.5f; // texture(s = u, t = 1 - v), ImageNode.cpp:141-145
		const float flx = std::floor(x), fly = std::floor(y);
		int ix = (int)flx, iy = (int)fly;
		const float fx = x - flx, fy = y - fly;
		if (mInterp == PRB_TEX_CLOSEST) {
			if (fx > 0.5f)
				++ix;
			if (fy > 0.5f)
				++iy;
			fetch(ix, iy, rgb);
		} else if (mInterp == PRB_TEX_BILINEAR) {
			float c00[3], c10[3], c01[3], c11[3];
			fetch(ix, iy, c00);
			fetch(ix + 1, iy, c10);
			fetch(ix, iy + 1, c01);
			fetch(ix + 1, iy + 1, c11);
			for (int c = 0; c < 3; ++c)
				rgb[c] = (c00[c] * (1 - fx) + c10[c] * fx) * (1 - fy) + (c01[c] * (1 - fx) + c11[c] * fx) * fy;
		} else {
			float wx[4], wy[4];
			bspline(fx, wx);
			bspline(fy, wy);
			rgb[0] = rgb[1] = rgb[2] = 0;
			for (int j = 0; j < 4; ++j) {
				float row[3] = { 0, 0, 0 };
				for (int i = 0; i < 4; ++i) {
					float t[3];
					fetch(ix - 1 + i, iy - 1 + j, t);
					for (int c = 0; c < 3; ++c)
						row[c] += wx[i] * t[c];
				}
				for (int c = 0; c < 3; ++c)
					rgb[c] += wy[j] * row[c];
			}
		}
		if (!mImg.linear) // RGBConverter::linearize as written (RGBConverter.cpp:53-58)
			for (int c = 0; c < 3; ++c)
				rgb[c] = rgb[c] <= 0.04045f ? rgb[c] / 12.92f * rgb[c] : (float)std::pow((double)((rgb[c] + 0.055f) / 1.055f), (double)2.4f);
	}
	SpectralBlob eval(const ShadingContext& ctx) const override
	{
		float rgb[3], k[3];
		lookup(ctx.UV.x, ctx.UV.y, rgb);
		mUpsampler->prepare(&rgb[0], &rgb[1], &rgb[2], &k[0], &k[1], &k[2], 1);
		SpectralBlob r;
		for (int i = 0; i < 4; ++i) { // SpectralUpsampler::compute, SpectralUpsampler.h:45-49
			const float q = (k[0] * ctx.WavelengthNM[i] + k[1]) * ctx.WavelengthNM[i] + k[2];
			r[i]		  = 0.5f * q * (1.0f / std::sqrt(q * q + 1.0f)) + 0.5f;
		}
		return r;
	}
	void queryRecommendedSize(int& w, int& h) const override
	{
		w = (int)mImg.width;
		h = (int)mImg.height;
	}
	uint32 emit(NodeEmitter& e) const override
	{
		uint32 id;
		if (e.find(this, id))
			return id;
		prb_node n{};
		n.type	= PRB_NODE_IMAGE;
		n.flags = devFlags(flags());
		n.a		= (uint32)e.pool->size();
		n.b		= mImg.width | (mImg.height << 16);
		n.p[0]	= (float)mInterp;
		n.p[1]	= (float)mWrapS;
		n.p[2]	= (float)mWrapT;
		n.p[3]	= mImg.linear ? 0.0f : 1.0f;
		e.pool->insert(e.pool->end(), mImg.data.begin(), mImg.data.end());
		e.upsampler = mUpsampler.get();
		return e.add(this, n);
	}
	std::string dumpInformation() const override { return mFile + " [NonParam]"; }

private:
	ImageData mImg;
	int mInterp, mWrapS, mWrapT;
	std::shared_ptr<SpectralUpsampler> mUpsampler;
	std::string mFile;
};
} // namespace
std::shared_ptr<FloatSpectralNode> makeImageNode(const std::string& file, int interp, int wrapS, int wrapT, const std::shared_ptr<SpectralUpsampler>& upsampler)
{
	ImageData img;
	if (!loadImage(file, img)) {
		PR_LOG(L_ERROR) << "Could not read image " << file << " (supported: EXR scanline half/float with no / ZIPS / ZIP compression, PFM, binary PPM / PGM)" << std::endl;
		return nullptr;
	}
	if (img.width == 0 || img.height == 0 || img.width > 65535 || img.height > 65535) {
		PR_LOG(L_ERROR) << "Image " << file << ": unsupported size " << img.width << "x" << img.height << std::endl;
		return nullptr;
	}
	return std::make_shared<ImageSpectralNode>(std::move(img), interp, wrapS, wrapT, upsampler, file);
}
namespace {
} // namespace

// ------------------------------------------------------------------ plugins
namespace {
class SpectralValuePlugin : public INodePlugin { // SpectralValueNode.cpp:12-70
public:
	std::shared_ptr<INode> create(const std::string& type_name, const SceneLoadContext& ctx) override
	{
		const auto upsampler = ctx.environment()->defaultSpectralUpsampler();
		float in[3]			 = { ctx.parameters().getParameter(0).getNumber(0.0f), ctx.parameters().getParameter(1).getNumber(0.0f),
						 ctx.parameters().getParameter(2).getNumber(0.0f) };
		const float max		 = std::max(in[0], std::max(in[1], in[2]));
		float blob[3];
		if (type_name == "refl" || type_name == "reflection") {
			if (max > 1.0f)
				PR_LOG(L_WARNING) << "Given reflective rgb contains coefficients above 1" << std::endl;
			upsampler->prepare(&in[0], &in[1], &in[2], &blob[0], &blob[1], &blob[2], 1);
			return std::make_shared<ParametricSpectralNode>(blob[0], blob[1], blob[2], 1.0f, false);
		} else { // illum
			float power = 1;
			if (max <= 0.0f) {
				upsampler->prepare(&in[0], &in[1], &in[2], &blob[0], &blob[1], &blob[2], 1);
			} else {
				const float scale = 2 * max;
				float s[3]		  = { in[0] / scale, in[1] / scale, in[2] / scale };
				upsampler->prepare(&s[0], &s[1], &s[2], &blob[0], &blob[1], &blob[2], 1);
				power = scale;
			}
			return std::make_shared<ParametricSpectralNode>(blob[0], blob[1], blob[2], power, true);
		}
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "refl", "reflection", "illum", "illumination" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Spectral Value Node: value1 value2 value3 (rgb)"; }
};

class IlluminantPlugin : public INodePlugin { // IlluminantNode.cpp:92-150
public:
	std::shared_ptr<INode> create(const std::string&, const SceneLoadContext& ctx) override
	{
		std::string illum = ctx.parameters().getString("spectrum", "");
		if (illum.empty())
			illum = ctx.parameters().getString(0, "D65");
		std::transform(illum.begin(), illum.end(), illum.begin(), [](char c) { return (char)std::tolower(c); });
		if (illum == "e")
			return makeConstSpectralNode(1.0f);
		size_t count;
		float start, end;
		const float* d = illuminantTable(illum, count, start, end);
		if (!d) {
			PR_LOG(L_ERROR) << "Unknown illuminant spectrum " << illum << std::endl;
			return nullptr;
		}
		return std::make_shared<EquidistantSpectrumNode>(std::vector<float>(d, d + count), start, end, "Illuminant " + illum);
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "illuminant" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Illuminant Node: spectrum d65|d50|d55|d75|a|c|e|f1..f12"; }
};

class SpectrumPlugin : public INodePlugin { // SpectralConstNode.cpp:12-33 ('spectrum :start :end v...')
public:
	std::shared_ptr<INode> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const float start = ctx.parameters().getNumber("start", PR_CIE_WAVELENGTH_START);
		const float end	  = ctx.parameters().getNumber("end", PR_CIE_WAVELENGTH_END);
		std::vector<float> data;
		for (size_t i = 0; i < ctx.parameters().positionalParameterCount(); ++i)
			data.push_back(ctx.parameters().getParameter(i).getNumber(0.0f));
		if (data.size() < 2) {
			PR_LOG(L_ERROR) << "spectrum node needs at least two values" << std::endl;
			return nullptr;
		}
		return std::make_shared<EquidistantSpectrumNode>(data, start, end, "Spectrum");
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "spectrum" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Spectrum Node: :start nm :end nm values..."; }
};

class ReflectiveNodePlugin : public INodePlugin { // ReflectiveNode.cpp:225-245,387-398 (lookup_index only)
public:
	std::shared_ptr<INode> create(const std::string&, const SceneLoadContext& ctx) override
	{
		std::string name = ctx.parameters().getParameter(0).getString("bk7");
		std::transform(name.begin(), name.end(), name.begin(), [](char c) { return (char)std::tolower(c); });
		if (name == "bk7" || name == "glass")
			return std::make_shared<SellmeierIndexNode>(std::vector<float>{ 1.03961212f, 0.231792344f, 1.01046945f },
														std::vector<float>{ 0.00600069867f, 0.0200179144f, 103.560653f });
		if (name == "h2o" || name == "water")
			return std::make_shared<SellmeierIndexNode>(std::vector<float>{ 5.684027565e-1f, 1.726177391e-1f, 2.086189578e-2f, 1.130748688e-1f },
														std::vector<float>{ 5.101829712e-3f, 1.821153936e-2f, 2.620722293e-2f, 1.069792721e1f });
		if (name == "diamond")
			return std::make_shared<SellmeierIndexNode>(std::vector<float>{ 0.3306f, 4.3356f }, std::vector<float>{ 0.030625f, 0.011236f });
		if (name == "vacuum" || name == "none")
			return makeConstSpectralNode(1.0f);
		if (name == "air")
			return makeConstSpectralNode(1.000277f);
		PR_LOG(L_ERROR) << "Unknown lookup name " << name << std::endl;
		return nullptr;
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "lookup_index" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Refractive index lookup: bk7|glass|h2o|water|diamond|vacuum|none|air"; }
};

class SpectralMathPlugin : public INodePlugin { // SpectralMathNode.cpp ('smul' only)
public:
	std::shared_ptr<INode> create(const std::string&, const SceneLoadContext& ctx) override
	{
		return std::make_shared<MulSpectralMath>(ctx.lookupSpectralNode(ctx.parameters().getParameter(0)),
												 ctx.lookupSpectralNode(ctx.parameters().getParameter(1)));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "smul" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Spectral 'smul' Node: op1 op2"; }
};

class CheckerboardPlugin : public INodePlugin { // CheckerboardNode.cpp:72-95
public:
	std::shared_ptr<INode> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const auto op1 = ctx.lookupSpectralNode(ctx.parameters().getParameter(0), 0.8f);
		const auto op2 = ctx.lookupSpectralNode(ctx.parameters().getParameter(1), 0.2f);
		const auto su  = ctx.lookupScalarNode(ctx.parameters().getParameter(2), 5);
		const auto sv  = ctx.lookupScalarNode(ctx.parameters().getParameter(3), 5);
		if (!su->isConst() || !sv->isConst()) {
			PR_LOG(L_ERROR) << "checkerboard: only constant scales are supported on the device path" << std::endl;
			return nullptr;
		}
		const ShadingContext sc;
		const size_t pc = ctx.parameters().positionalParameterCount();
		const int mode	= pc == 2 ? 0 : (pc == 3 ? 1 : 2);
		return std::make_shared<CheckerboardNode>(op1, op2, su->eval(sc), mode == 2 ? sv->eval(sc) : su->eval(sc), mode);
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "grid", "checkerboard" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Checkerboard Node: color1 color2 [scale_u [scale_v]]"; }
};
} // namespace

void registerNodePlugins(std::vector<std::shared_ptr<IPlugin>>& out)
{
	out.push_back(std::make_shared<SpectralValuePlugin>());
	out.push_back(std::make_shared<IlluminantPlugin>());
	out.push_back(std::make_shared<SpectrumPlugin>());
	out.push_back(std::make_shared<ReflectiveNodePlugin>());
	out.push_back(std::make_shared<SpectralMathPlugin>());
	out.push_back(std::make_shared<CheckerboardPlugin>());
}
} // namespace PR
