// Image files for texture nodes (reference src/loader/image/ImageIO.cpp reads through OpenImageIO, which is not in this image):
// a small reader for the formats a scene on this path needs -- OpenEXR scanline files (HALF / FLOAT, uncompressed, ZIPS, ZIP:
// what this library's own writer produces and what OpenEXR tools write by default), Portable Float Maps (PF / Pf) and binary
// PPM / PGM (P6 / P5, 8 or 16 bit).  Float formats are linear, integer formats are sRGB encoded (OpenImageIO's
// "oiio:ColorSpace" default, which NonParametricImageNode tests, ImageNode.cpp:118).  Rows are returned top to bottom.
#include "prh.h"

#include <cstring>
#include <fstream>
#include <zlib.h>

namespace PR {
namespace {
bool readFile(const std::string& path, std::string& out)
{
	std::ifstream f(path, std::ios::binary);
	if (!f)
		return false;
	out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
	return true;
}
float halfToFloat(uint16_t h)
{
	const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
	uint32_t exp		= (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
	if (exp == 0) {
		if (man == 0) {
			bits = sign;
		} else { // subnormal
			exp = 127 - 15 + 1;
			while (!(man & 0x400u)) {
				man <<= 1;
				--exp;
			}
			bits = sign | (exp << 23) | ((man & 0x3FFu) << 13);
		}
	} else if (exp == 31) {
		bits = sign | 0x7F800000u | (man << 13);
	} else {
		bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
	}
	float f;
	std::memcpy(&f, &bits, 4);
	return f;
}

// ---- PNM / PFM
bool nextToken(const std::string& b, size_t& p, std::string& tok)
{
	for (;;) {
		while (p < b.size() && std::isspace((unsigned char)b[p]))
			++p;
		if (p < b.size() && b[p] == '#') {
			while (p < b.size() && b[p] != '\n')
				++p;
			continue;
		}
		break;
	}
	const size_t s = p;
	while (p < b.size() && !std::isspace((unsigned char)b[p]))
		++p;
	tok = b.substr(s, p - s);
	return !tok.empty();
}
bool readPNM(const std::string& b, ImageData& img)
{
	size_t p = 0;
	std::string magic, ws, hs, ms;
	if (!nextToken(b, p, magic) || !nextToken(b, p, ws) || !nextToken(b, p, hs) || !nextToken(b, p, ms))
		return false;
	img.width  = (uint32)std::stoul(ws);
	img.height = (uint32)std::stoul(hs);
	++p; // the single whitespace after the header
	const size_t n = (size_t)img.width * img.height;
	if (magic == "PF" || magic == "Pf") {
		img.channels	  = magic == "PF" ? 3 : 1;
		img.linear		  = true;
		const float scale = std::stof(ms);
		const bool little = scale < 0;
		if (b.size() < p + n * img.channels * 4)
			return false;
		img.data.resize(n * img.channels);
		for (uint32 y = 0; y < img.height; ++y) // PFM rows run bottom to top
			for (size_t i = 0; i < (size_t)img.width * img.channels; ++i) {
				unsigned char c[4];
				std::memcpy(c, b.data() + p + (((size_t)(img.height - 1 - y) * img.width * img.channels) + i) * 4, 4);
				if (!little)
					std::swap(c[0], c[3]), std::swap(c[1], c[2]);
				float f;
				std::memcpy(&f, c, 4);
				img.data[(size_t)y * img.width * img.channels + i] = f;
			}
		return true;
	}
	if (magic == "P6" || magic == "P5") {
		img.channels	 = magic == "P6" ? 3 : 1;
		img.linear		 = false;
		const uint32 max = (uint32)std::stoul(ms);
		const int bytes	 = max > 255 ? 2 : 1;
		if (max == 0 || b.size() < p + n * img.channels * bytes)
			return false;
		img.data.resize(n * img.channels);
		for (size_t i = 0; i < n * img.channels; ++i) {
			const unsigned char* s = reinterpret_cast<const unsigned char*>(b.data()) + p + i * bytes;
			const uint32 v		   = bytes == 2 ? ((uint32)s[0] << 8 | s[1]) : s[0];
			img.data[i]			   = (float)v / (float)max;
		}
		return true;
	}
	return false;
}

// ---- OpenEXR, single-part scanline
struct ExrChannel {
	std::string name;
	int type; // 0 uint, 1 half, 2 float
};
bool readEXR(const std::string& b, ImageData& img)
{
	auto rd32 = [&](size_t p) {
		int32_t v;
		std::memcpy(&v, b.data() + p, 4);
		return v;
	};
	if (b.size() < 8 || rd32(0) != 20000630)
		return false;
	const int32_t version = rd32(4);
	if (version & 0x1A00) // tiled, multi-part or deep
		return false;
	size_t p = 8;
	std::vector<ExrChannel> chans;
	int compression = 0;
	int32_t dw[4]	= { 0, 0, -1, -1 };
	bool increasing = true;
	while (p < b.size() && b[p] != '\0') {
		const std::string name(b.data() + p);
		p += name.size() + 1;
		const std::string type(b.data() + p);
		p += type.size() + 1;
		const int32_t size = rd32(p);
		p += 4;
		if (name == "channels") {
			size_t q = p;
			while (b[q] != '\0') {
				ExrChannel c;
				c.name = std::string(b.data() + q);
				q += c.name.size() + 1;
				c.type = rd32(q);
				q += 16; // pixel type, pLinear + 3 reserved, x sampling, y sampling
				chans.push_back(c);
			}
		} else if (name == "compression") {
			compression = (unsigned char)b[p];
		} else if (name == "dataWindow") {
			for (int i = 0; i < 4; ++i)
				dw[i] = rd32(p + 4 * i);
		} else if (name == "lineOrder") {
			increasing = b[p] == 0;
		}
		p += size;
	}
	++p;
	(void)increasing; // every block carries its y coordinate
	if (chans.empty() || dw[2] < dw[0] || dw[3] < dw[1] || (compression != 0 && compression != 2 && compression != 3))
		return false; // NO_COMPRESSION, ZIPS, ZIP
	const uint32 W = (uint32)(dw[2] - dw[0] + 1), H = (uint32)(dw[3] - dw[1] + 1);
	const int linesPerBlock = compression == 3 ? 16 : 1;
	const size_t nBlocks	= (H + linesPerBlock - 1) / linesPerBlock;
	// channel -> output slot: R, G, B (or Y) by name; the file stores channels alphabetically
	int slot[64];
	img.channels = 0;
	bool hasRGB	 = false;
	for (const ExrChannel& c : chans)
		if (c.name == "R" || c.name == "G" || c.name == "B")
			hasRGB = true;
	for (size_t i = 0; i < chans.size() && i < 64; ++i) {
		const std::string& n = chans[i].name;
		slot[i]				 = hasRGB ? (n == "R" ? 0 : n == "G" ? 1 : n == "B" ? 2 : -1) : (n == "Y" ? 0 : -1);
	}
	img.channels = hasRGB ? 3 : 1;
	img.width	 = W;
	img.height	 = H;
	img.linear	 = true;
	img.data.assign((size_t)W * H * img.channels, 0.0f);
	size_t rowBytes = 0;
	for (const ExrChannel& c : chans)
		rowBytes += (size_t)W * (c.type == 1 ? 2 : 4);
	std::vector<unsigned char> raw, tmp;
	for (size_t blk = 0; blk < nBlocks; ++blk) {
		uint64_t off;
		std::memcpy(&off, b.data() + p + 8 * blk, 8);
		if (off + 8 > b.size())
			return false;
		const int32_t y0 = rd32(off), size = rd32(off + 4);
		const int lines	 = std::min<int>(linesPerBlock, (int)H - (y0 - dw[1]));
		if (lines <= 0 || off + 8 + (size_t)size > b.size())
			return false;
		const size_t want = rowBytes * lines;
		raw.resize(want);
		if (compression == 0 || (size_t)size == want) {
			std::memcpy(raw.data(), b.data() + off + 8, want);
		} else {
			tmp.resize(want);
			uLongf got = (uLongf)want;
			if (uncompress(tmp.data(), &got, reinterpret_cast<const Bytef*>(b.data() + off + 8), (uLong)size) != Z_OK || got != want)
				return false;
			for (size_t i = 1; i < want; ++i) // predictor
				tmp[i] = (unsigned char)(tmp[i - 1] + tmp[i] - 128);
			const size_t half = (want + 1) / 2; // de-interleave
			for (size_t i = 0; i < want; ++i)
				raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
		}
		const unsigned char* s = raw.data();
		for (int l = 0; l < lines; ++l) {
			const uint32 y = (uint32)(y0 - dw[1] + l);
			for (size_t ci = 0; ci < chans.size(); ++ci) {
				const int t = chans[ci].type;
				for (uint32 x = 0; x < W; ++x) {
					float v;
					if (t == 1) {
						uint16_t h;
						std::memcpy(&h, s, 2);
						s += 2;
						v = halfToFloat(h);
					} else if (t == 2) {
						std::memcpy(&v, s, 4);
						s += 4;
					} else {
						uint32_t u;
						std::memcpy(&u, s, 4);
						s += 4;
						v = (float)u;
					}
					if (ci < 64 && slot[ci] >= 0)
						img.data[((size_t)y * W + x) * img.channels + slot[ci]] = v;
				}
			}
		}
	}
	return true;
}
} // namespace

bool loadImage(const std::string& path, ImageData& img)
{
	std::string b;
	if (!readFile(path, b) || b.size() < 4)
		return false;
	try {
		if ((unsigned char)b[0] == 0x76 && (unsigned char)b[1] == 0x2F && (unsigned char)b[2] == 0x31 && (unsigned char)b[3] == 0x01)
			return readEXR(b, img);
		if (b[0] == 'P')
			return readPNM(b, img);
	} catch (const std::exception&) {
		return false;
	}
	return false;
}
} // namespace PR
