// Output specification and image writer: the `(output :name '..' (channel :type '..' :color '..') ...)`
// blocks of a .prc scene and the EXR files the reference writes for them.
//
//   OutputSpecification::parse   reference src/loader/output/io/OutputSpecification.cpp:256-437 (channel type tables :108-180,
//                                names "R,G,B" / "<name>.R,.G,.B", "<var>.x,.y,.z", "<var>")
//   ImageWriter::save            reference src/loader/output/io/ImageWriter.cpp:41-262: one float channel per component,
//                                colour through the ToneMapper (src/core/spectral/ToneMapper.cpp:14-62: linear sRGB via
//                                RGBConverter::fromXYZ, XYZ, normalised XYZ, luminance), technical AOVs divided by the pixel's
//                                sample count, counters as floats; data window = view, display window = film.
//   saveOutputs                  reference OutputSpecification::save :439-464: <dir>/results/<name>.exr
//
// The reference writes through OpenImageIO; here the OpenEXR container is written directly (single-part scanline file,
// FLOAT channels, no compression -- the subset every EXR reader accepts).  LPE and custom channels are not on the device
// path: a channel that asks for them, or for an AOV the device does not accumulate, is written as zeros with a warning,
// as the reference does for a channel it cannot acquire (ImageWriter.cpp:148-156).
#include "prh.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sys/stat.h>

namespace PR {
namespace {
struct VarName {
	const char* str;	   // accepted spelling
	int var;			   // OV_* (OV_Unsupported: known to the reference, not accumulated by the device path)
	const char* canonical; // variableToString(): the first spelling of the variable, used as the channel name
};
const VarName kSpectral[] = { { "color", OV_Output, "color" }, { "spectral", OV_Output, "color" }, { "output", OV_Output, "color" }, { "rgb", OV_Output, "color" },
							  { "online_mean", OV_OnlineMean, "online_mean" }, { "variance", OV_OnlineVariance, "variance" },
							  { "online_variance", OV_OnlineVariance, "variance" }, { "var", OV_OnlineVariance, "variance" } };
const VarName k1D[]		  = { { "entity_id", OV_EntityID, "entity_id" }, { "entity", OV_EntityID, "entity_id" }, { "id", OV_EntityID, "entity_id" },
							  { "material_id", OV_MaterialID, "material_id" }, { "material", OV_MaterialID, "material_id" }, { "mat", OV_MaterialID, "material_id" },
							  { "emission_id", OV_EmissionID, "emission_id" }, { "emission", OV_EmissionID, "emission_id" },
							  { "displace_id", OV_DisplaceID, "displace_id" }, { "displace", OV_DisplaceID, "displace_id" },
							  { "depth", OV_Depth, "depth" }, { "d", OV_Depth, "depth" } };
const VarName kCounter[]  = { { "sample_count", OV_SampleCount, "sample_count" }, { "samples", OV_SampleCount, "sample_count" }, { "s", OV_SampleCount, "sample_count" },
							  { "feedback", OV_Feedback, "feedback" }, { "f", OV_Feedback, "feedback" }, { "error", OV_Feedback, "feedback" } };
const VarName k3D[]		  = { { "position", OV_Position, "position" }, { "pos", OV_Position, "position" }, { "p", OV_Position, "position" },
							  { "normal", OV_Normal, "normal" }, { "norm", OV_Normal, "normal" }, { "n", OV_Normal, "normal" },
							  { "normal_geometric", OV_NormalG, "normal_geometric" }, { "ng", OV_NormalG, "normal_geometric" },
							  { "tangent", OV_Tangent, "tangent" }, { "tan", OV_Tangent, "tangent" }, { "nx", OV_Tangent, "tangent" },
							  { "bitangent", OV_Bitangent, "bitangent" }, { "binormal", OV_Bitangent, "bitangent" }, { "bi", OV_Bitangent, "bitangent" },
							  { "ny", OV_Bitangent, "bitangent" }, { "view", OV_View, "view" }, { "v", OV_View, "view" },
							  { "texture", OV_UVW, "texture" }, { "uvw", OV_UVW, "texture" }, { "uv", OV_UVW, "texture" }, { "tex", OV_UVW, "texture" } };
template <size_t N>
bool lookup(const VarName (&table)[N], const std::string& type, int& var, std::string& name)
{
	for (size_t i = 0; i < N; ++i)
		if (type == table[i].str) {
			var	 = table[i].var;
			name = table[i].canonical;
			return true;
		}
	return false;
}
std::string lower(std::string s)
{
	std::transform(s.begin(), s.end(), s.begin(), ::tolower);
	return s;
}
} // namespace

void OutputSpecification::parse(const DL::DataGroup& entry)
{
	const DL::Data nameD = entry.getFromKey("name");
	if (nameD.type() != DL::DT_String) {
		PR_LOG(L_ERROR) << "No name given for output" << std::endl;
		return;
	}
	OutputFile file;
	file.name = nameD.getString();
	for (size_t i = 0; i < entry.anonymousCount(); ++i) {
		const DL::Data channelD = entry.at(i);
		if (channelD.type() != DL::DT_Group)
			continue;
		const DL::DataGroup& channel = channelD.getGroup();
		if (channel.id() == "custom_channel") {
			PR_LOG(L_WARNING) << "Output '" << file.name << "': custom channels are not produced by the device path; skipped" << std::endl;
			continue;
		}
		if (channel.id() != "channel")
			continue;
		const DL::Data typeD = channel.getFromKey("type"), colorD = channel.getFromKey("color"), lpeD = channel.getFromKey("lpe");
		if (typeD.type() != DL::DT_String)
			continue;
		const std::string type = lower(typeD.getString());
		OutputChannel ch;
		if (colorD.type() == DL::DT_String) {
			const std::string color = lower(colorD.getString());
			if (color == "xyz")
				ch.tcm = ToneColorMode::XYZ;
			else if (color == "norm_xyz")
				ch.tcm = ToneColorMode::XYZNorm;
			else if (color == "lum" || color == "luminance" || color == "gray")
				ch.tcm = ToneColorMode::Luminance;
		}
		const std::string lpe = lpeD.type() == DL::DT_String ? lpeD.getString() : "";
		std::string name;
		if (lookup(kSpectral, type, ch.variable, name)) {
			ch.kind = OutputChannel::Spectral;
			if (ch.variable != OV_Output)
				ch.name = name; // raw spectral AOVs carry their variable name; the colour channel stays unnamed ("R","G","B")
		} else if (lookup(k3D, type, ch.variable, name)) {
			ch.kind = OutputChannel::ThreeD;
			ch.name = name;
		} else if (lookup(k1D, type, ch.variable, name)) {
			ch.kind = OutputChannel::OneD;
			ch.name = name;
		} else if (lookup(kCounter, type, ch.variable, name)) {
			ch.kind = OutputChannel::Counter;
			ch.name = name;
		} else {
			PR_LOG(L_ERROR) << "Unknown channel type " << type << std::endl;
			continue;
		}
		if (!lpe.empty()) { // OutputSpecification.cpp:296-331: an invalid expression is dropped with an error, the channel stays
			if (!compileLPE(lpe).valid) {
				PR_LOG(L_ERROR) << "Invalid LPE '" << lpe << "'. Skipping entry" << std::endl;
			} else {
				ch.name += "[" + lpe + "]";
				ch.lpe = lpe;
				if (ch.kind == OutputChannel::Spectral && ch.variable == OV_Output) {
					size_t k = 0;
					while (k < mLPEs.size() && mLPEs[k] != lpe)
						++k;
					if (k == mLPEs.size() && k < PRB_MAX_LPE)
						mLPEs.push_back(lpe);
					if (k < PRB_MAX_LPE) {
						ch.lpeIndex = (int)k;
					} else {
						PR_LOG(L_WARNING) << "Output '" << file.name << "': more than " << PRB_MAX_LPE << " distinct light path expressions; '" << lpe
										  << "' is written as zeros" << std::endl;
						ch.variable = OV_Unsupported;
					}
				}
				// shading-point channels (position, normal, ...) with an expression: the only path a shading-point fragment
				// carries is the camera token (direct.cpp:86-87), matched when the file is written (saveImage)
			}
		}
		if (ch.variable == OV_Unsupported)
			PR_LOG(L_WARNING) << "Output '" << file.name << "': channel '" << type << "' is not accumulated by the device path; written as zeros" << std::endl;
		file.channels.push_back(ch);
	}
	mFiles.push_back(file);
}

// ------------------------------------------------------------------ OpenEXR container (scanline, FLOAT, uncompressed)
namespace {
void put32(std::string& b, int32_t v) { b.append(reinterpret_cast<const char*>(&v), 4); }
void putf(std::string& b, float v) { b.append(reinterpret_cast<const char*>(&v), 4); }
void putAttr(std::string& b, const char* name, const char* type, const std::string& value)
{
	b.append(name);
	b.push_back('\0');
	b.append(type);
	b.push_back('\0');
	put32(b, (int32_t)value.size());
	b.append(value);
}
} // namespace

bool writeEXR(const std::string& path, const std::vector<std::string>& channelNames, const std::vector<const float*>& planes, uint32 width, uint32 height,
			  int32_t offX, int32_t offY, uint32 fullWidth, uint32 fullHeight)
{
	const size_t nch = channelNames.size();
	if (nch == 0 || planes.size() != nch)
		return false;
	std::vector<size_t> order(nch); // channels are stored in alphabetical order
	for (size_t i = 0; i < nch; ++i)
		order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return channelNames[a] < channelNames[b]; });

	std::string h;
	put32(h, 20000630); // magic
	put32(h, 2);		// version 2, single-part scanline
	{
		std::string v;
		for (size_t k : order) {
			v.append(channelNames[k]);
			v.push_back('\0');
			put32(v, 2); // FLOAT
			v.push_back('\0'); // pLinear
			v.append(3, '\0');
			put32(v, 1);
			put32(v, 1);
		}
		v.push_back('\0');
		putAttr(h, "channels", "chlist", v);
	}
	putAttr(h, "compression", "compression", std::string(1, '\0'));
	{
		std::string v;
		put32(v, offX), put32(v, offY), put32(v, offX + (int32_t)width - 1), put32(v, offY + (int32_t)height - 1);
		putAttr(h, "dataWindow", "box2i", v);
	}
	{
		std::string v;
		put32(v, 0), put32(v, 0), put32(v, (int32_t)fullWidth - 1), put32(v, (int32_t)fullHeight - 1);
		putAttr(h, "displayWindow", "box2i", v);
	}
	putAttr(h, "lineOrder", "lineOrder", std::string(1, '\0'));
	{
		std::string v;
		putf(v, 1.0f);
		putAttr(h, "pixelAspectRatio", "float", v);
	}
	{
		std::string v;
		putf(v, 0.0f), putf(v, 0.0f);
		putAttr(h, "screenWindowCenter", "v2f", v);
	}
	{
		std::string v;
		putf(v, 1.0f);
		putAttr(h, "screenWindowWidth", "float", v);
	}
	putAttr(h, "Software", "string", "prb200 (PearRay direct integrator on B200)");
	h.push_back('\0');

	std::ofstream f(path, std::ios::binary);
	if (!f)
		return false;
	f.write(h.data(), (std::streamsize)h.size());
	const uint64 lineBytes = 8 + (uint64)nch * width * 4;
	uint64 off			   = h.size() + (uint64)height * 8;
	for (uint32 y = 0; y < height; ++y, off += lineBytes)
		f.write(reinterpret_cast<const char*>(&off), 8);
	for (uint32 y = 0; y < height; ++y) {
		const int32_t yy = offY + (int32_t)y, sz = (int32_t)(nch * width * 4);
		f.write(reinterpret_cast<const char*>(&yy), 4);
		f.write(reinterpret_cast<const char*>(&sz), 4);
		for (size_t k : order)
			f.write(reinterpret_cast<const char*>(planes[k] + (size_t)y * width), (std::streamsize)width * 4);
	}
	return (bool)f;
}

// ------------------------------------------------------------------ ImageWriter
void toneMap(ToneColorMode tcm, const float* xyz, float* rgb)
{ // ToneMapper::map for one pixel, ToneMapper.cpp:14-62 + RGBConverter::fromXYZ, RGBConverter.cpp:15-24
	const float X = xyz[0], Y = xyz[1], Z = xyz[2];
	switch (tcm) {
	case ToneColorMode::SRGB:
		rgb[0] = std::max(0.0f, 3.240970e+00f * X - 1.537383e+00f * Y - 4.986108e-01f * Z);
		rgb[1] = std::max(0.0f, -9.692436e-01f * X + 1.875968e+00f * Y + 4.155506e-02f * Z);
		rgb[2] = std::max(0.0f, 5.563008e-02f * X - 2.039770e-01f * Y + 1.056972e+00f * Z);
		break;
	case ToneColorMode::XYZ: rgb[0] = X, rgb[1] = Y, rgb[2] = Z; break;
	case ToneColorMode::XYZNorm: {
		// the reference scales the (uninitialised) output in place here (ToneMapper.cpp:36-46); the evident intent --
		// chromaticity x, y, z = XYZ / (X + Y + Z) -- is what is written
		const float N = X + Y + Z, F = N != 0 ? 1.0f / N : 0;
		rgb[0] = X * F, rgb[1] = Y * F, rgb[2] = Z * F;
		break;
	}
	case ToneColorMode::Luminance: rgb[0] = rgb[1] = rgb[2] = Y; break;
	}
}

bool saveImage(const std::string& path, const OutputFile& file, const FilmView& film)
{
	const size_t n = (size_t)film.width * film.height;
	std::vector<std::string> names;
	std::vector<std::vector<float>> data;
	auto addPlane = [&](const std::string& nm) -> std::vector<float>& {
		names.push_back(nm);
		data.emplace_back(n, 0.0f);
		return data.back();
	};
	// channel order as the reference lays it out: spectral, 3D, 1D, counters (ImageWriter.cpp:80-103)
	for (const OutputChannel& c : file.channels) {
		if (c.kind != OutputChannel::Spectral)
			continue;
		const size_t base = data.size();
		addPlane(c.name.empty() ? "R" : c.name + ".R");
		addPlane(c.name.empty() ? "G" : c.name + ".G");
		addPlane(c.name.empty() ? "B" : c.name + ".B");
		const float* src = c.variable == OV_Output ? film.xyz : c.variable == OV_OnlineMean ? film.onlineMean : c.variable == OV_OnlineVariance ? film.onlineVariance : nullptr;
		if (!c.lpe.empty()) // the expression's own film (prb_film_download_lpe); the variance estimators have no per-expression copy
			src = c.lpeIndex >= 0 && (size_t)c.lpeIndex < film.lpe.size() ? film.lpe[c.lpeIndex] : nullptr;
		if (!src)
			continue;
		for (size_t i = 0; i < n; ++i) {
			float rgb[3];
			toneMap(c.tcm, src + 3 * i, rgb);
			data[base][i] = rgb[0], data[base + 1][i] = rgb[1], data[base + 2][i] = rgb[2];
		}
	}
	// a shading-point fragment (position, normal, ..., sample count) carries the path of the first camera vertex, i.e. the
	// camera token alone (direct.cpp:86-87), so a channel with an expression holds the AOV when the expression accepts "C"
	// and stays empty otherwise (LocalFrameOutputDevice.cpp:177-187, 289-302)
	auto passesLPE = [](const OutputChannel& c) { return c.lpe.empty() || compileLPE(c.lpe).match({ { 0, 2 } }); };
	auto sampleFactor = [&](size_t i) { // technical AOVs are sums over the samples: scaled by 1 / sample count
		const uint32 s = film.sampleCount ? film.sampleCount[i] : 0;
		return s == 0 ? 1.0f : 1.0f / s;
	};
	for (const OutputChannel& c : file.channels) {
		if (c.kind != OutputChannel::ThreeD)
			continue;
		const size_t base = data.size();
		addPlane(c.name + ".x");
		addPlane(c.name + ".y");
		addPlane(c.name + ".z");
		if (!film.aov || c.variable == OV_Unsupported || !passesLPE(c))
			continue;
		for (size_t i = 0; i < n; ++i) {
			const float* a = film.aov + 10 * i; // prb_film_aov layout: N(3) P(3) u v depth entity
			const float sf = sampleFactor(i);
			float v[3]	   = { 0, 0, 0 };
			const float* b = film.aovExt ? film.aovExt + PRB_AOV_EXT * i : nullptr; // prb_film_download_aov_ext layout
			if (c.variable == OV_Normal || c.variable == OV_NormalG) // Surface.N is a copy of Geometry.N (IntersectionPoint::setForSurface)
				v[0] = a[0], v[1] = a[1], v[2] = a[2];
			else if (b && c.variable == OV_Tangent)
				v[0] = b[0], v[1] = b[1], v[2] = b[2];
			else if (b && c.variable == OV_Bitangent)
				v[0] = b[3], v[1] = b[4], v[2] = b[5];
			else if (b && c.variable == OV_View)
				v[0] = b[6], v[1] = b[7], v[2] = b[8];
			else if (c.variable == OV_Position)
				v[0] = a[3], v[1] = a[4], v[2] = a[5];
			else if (c.variable == OV_UVW)
				v[0] = a[6], v[1] = a[7];
			data[base][i] = sf * v[0], data[base + 1][i] = sf * v[1], data[base + 2][i] = sf * v[2];
		}
	}
	for (const OutputChannel& c : file.channels) {
		if (c.kind != OutputChannel::OneD)
			continue;
		std::vector<float>& p = addPlane(c.name);
		if (!film.aov || c.variable == OV_Unsupported || !passesLPE(c))
			continue;
		for (size_t i = 0; i < n; ++i) {
			float sum = 0;
			if (c.variable == OV_Depth)
				sum = film.aov[10 * i + 8];
			else if (c.variable == OV_EntityID)
				sum = film.aov[10 * i + 9];
			else if (c.variable == OV_MaterialID && film.aovExt)
				sum = film.aovExt[PRB_AOV_EXT * i + 9];
			else if (c.variable == OV_EmissionID && film.aovExt)
				sum = film.aovExt[PRB_AOV_EXT * i + 10];
			else if (c.variable == OV_DisplaceID) // GeometryPoint::DisplaceID stays PR_INVALID_ID for meshes, spheres and planes (GeometryPoint.h:24)
				sum = (film.sampleCount ? (float)film.sampleCount[i] : 0.0f) * 4294967296.0f;
			p[i] = sampleFactor(i) * sum;
		}
	}
	for (const OutputChannel& c : file.channels) {
		if (c.kind != OutputChannel::Counter)
			continue;
		std::vector<float>& p = addPlane(c.name);
		if (!passesLPE(c))
			continue;
		if (c.variable == OV_SampleCount && film.sampleCount)
			for (size_t i = 0; i < n; ++i)
				p[i] = static_cast<float>(film.sampleCount[i]);
		if (c.variable == OV_Feedback && film.feedback) // AOV_Feedback: OR of the OutputFeedback bits of rejected fragments
			for (size_t i = 0; i < n; ++i)
				p[i] = static_cast<float>(film.feedback[i]);
	}
	if (data.empty())
		return false;
	std::vector<const float*> planes;
	for (const auto& d : data)
		planes.push_back(d.data());
	return writeEXR(path, names, planes, film.width, film.height, (int32_t)film.offsetX, (int32_t)film.offsetY, film.fullWidth, film.fullHeight);
}

int OutputSpecification::save(const std::string& workingDir, const FilmView& film, uint32 contextIndex) const
{ // OutputSpecification::save, OutputSpecification.cpp:439-464: <workingDir>/results[_<index>]/<name>.exr
	std::string dir = workingDir.empty() ? std::string(".") : workingDir;
	dir += "/results";
	if (contextIndex > 0)
		dir += "_" + std::to_string(contextIndex);
	::mkdir(dir.c_str(), 0777); // does not matter if it exists
	int written = 0;
	for (const OutputFile& f : mFiles) {
		const std::string path = dir + "/" + f.name + ".exr";
		if (saveImage(path, f, film))
			++written;
		else
			PR_LOG(L_ERROR) << "Couldn't save image file " << path << std::endl;
	}
	return written;
}
} // namespace PR
