// Recursive-descent reader for the DataLisp subset used by .prc scene files.
// Grammar (reference external/DataLisp/src/internal/{Lexer,Parser}.cpp):
//   file   := group*
//   group  := '(' id { ':'key value | value } ')'
//   array  := '[' value { [','] value } ']'
//   value  := int | float | true | false | 'str' | "str" | array | group
// ';' starts a line comment; stray commas between entries are tolerated.
// The '$(expr)' expression VM is not supported (no example scene on the hot path uses it).
#pragma once
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace DL {
enum DataType { DT_None = 0, DT_Integer, DT_Float, DT_Bool, DT_String, DT_Group };

class DataGroup;
class Data {
public:
	Data() = default;
	explicit Data(const std::string& key)
		: mKey(key)
	{
	}
	DataType type() const { return mType; }
	bool isValid() const { return mType != DT_None; }
	bool isNumber() const { return mType == DT_Integer || mType == DT_Float; }
	const std::string& key() const { return mKey; }
	int64_t getInt() const { return mType == DT_Float ? (int64_t)mFloat : mInt; }
	float getNumber() const { return mType == DT_Integer ? (float)mInt : mFloat; }
	bool getBool() const { return mBool; }
	const std::string& getString() const { return mString; }
	const DataGroup& getGroup() const { return *mGroup; }
	DataGroup& getGroup() { return *mGroup; }

	void setInt(int64_t v)
	{
		mType = DT_Integer;
		mInt  = v;
	}
	void setFloat(float v)
	{
		mType  = DT_Float;
		mFloat = v;
	}
	void setBool(bool v)
	{
		mType = DT_Bool;
		mBool = v;
	}
	void setString(const std::string& s)
	{
		mType	= DT_String;
		mString = s;
	}
	void setGroup(const std::shared_ptr<DataGroup>& g)
	{
		mType  = DT_Group;
		mGroup = g;
	}

private:
	std::string mKey;
	DataType mType = DT_None;
	int64_t mInt   = 0;
	float mFloat   = 0;
	bool mBool	   = false;
	std::string mString;
	std::shared_ptr<DataGroup> mGroup;
};

class DataGroup {
public:
	explicit DataGroup(const std::string& id = "", bool array = false)
		: mID(id)
		, mArray(array)
	{
	}
	const std::string& id() const { return mID; }
	bool isArray() const { return mArray; }
	size_t anonymousCount() const { return mAnonymous.size(); }
	const Data& at(size_t i) const { return mAnonymous.at(i); }
	Data& at(size_t i) { return mAnonymous.at(i); }
	const std::vector<Data>& getAnonymousEntries() const { return mAnonymous; }
	const std::vector<Data>& getNamedEntries() const { return mNamed; }
	bool hasKey(const std::string& k) const
	{
		for (const auto& d : mNamed)
			if (d.key() == k)
				return true;
		return false;
	}
	Data getFromKey(const std::string& k) const
	{
		for (const auto& d : mNamed)
			if (d.key() == k)
				return d;
		return Data();
	}
	bool isAllAnonymousNumber() const
	{
		for (const auto& d : mAnonymous)
			if (!d.isNumber())
				return false;
		return true;
	}
	bool isAllAnonymousOfType(DataType t) const
	{
		for (const auto& d : mAnonymous)
			if (d.type() != t)
				return false;
		return true;
	}
	void add(const Data& d)
	{
		if (d.key().empty())
			mAnonymous.push_back(d);
		else
			mNamed.push_back(d);
	}
	void clear()
	{
		mAnonymous.clear();
		mNamed.clear();
	}

private:
	std::string mID;
	bool mArray;
	std::vector<Data> mAnonymous;
	std::vector<Data> mNamed;
};

class ParseError : public std::runtime_error {
public:
	using std::runtime_error::runtime_error;
};

class Parser {
public:
	explicit Parser(const std::string& src)
		: mSrc(src)
	{
	}
	std::vector<DataGroup> parse()
	{
		std::vector<DataGroup> groups;
		skip();
		while (mPos < mSrc.size()) {
			if (mSrc[mPos] != '(')
				fail("expected '(' at top level");
			groups.push_back(*parseGroup());
			skip();
		}
		return groups;
	}

private:
	[[noreturn]] void fail(const std::string& msg) const
	{
		std::stringstream s;
		s << "DataLisp parse error at line " << mLine << ": " << msg;
		throw ParseError(s.str());
	}
	void skip()
	{
		while (mPos < mSrc.size()) {
			const char c = mSrc[mPos];
			if (c == '\n') {
				++mLine;
				++mPos;
			} else if (c == ' ' || c == '\t' || c == '\r' || c == ',') {
				++mPos;
			} else if (c == ';') {
				while (mPos < mSrc.size() && mSrc[mPos] != '\n')
					++mPos;
			} else {
				break;
			}
		}
	}
	static bool isIdChar(char c)
	{
		return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' || c == '.';
	}
	std::string parseIdentifier()
	{
		const size_t s = mPos;
		while (mPos < mSrc.size() && isIdChar(mSrc[mPos]))
			++mPos;
		if (s == mPos)
			fail("expected identifier");
		return mSrc.substr(s, mPos - s);
	}
	std::shared_ptr<DataGroup> parseGroup()
	{
		++mPos; // '('
		skip();
		auto grp = std::make_shared<DataGroup>(parseIdentifier(), false);
		for (;;) {
			skip();
			if (mPos >= mSrc.size())
				fail("unterminated group '" + grp->id() + "'");
			const char c = mSrc[mPos];
			if (c == ')') {
				++mPos;
				return grp;
			}
			if (c == ':') {
				++mPos;
				const std::string key = parseIdentifier();
				skip();
				Data d(key);
				parseValue(d);
				grp->add(d);
			} else {
				Data d;
				parseValue(d);
				grp->add(d);
			}
		}
	}
	std::shared_ptr<DataGroup> parseArray()
	{
		++mPos; // '['
		auto grp = std::make_shared<DataGroup>("", true);
		for (;;) {
			skip();
			if (mPos >= mSrc.size())
				fail("unterminated array");
			if (mSrc[mPos] == ']') {
				++mPos;
				return grp;
			}
			Data d;
			parseValue(d);
			grp->add(d);
		}
	}
	void parseValue(Data& d)
	{
		if (mPos >= mSrc.size())
			fail("unexpected end of input");
		const char c = mSrc[mPos];
		if (c == '(') {
			d.setGroup(parseGroup());
		} else if (c == '[') {
			d.setGroup(parseArray());
		} else if (c == '\'' || c == '"') {
			++mPos;
			std::string s;
			while (mPos < mSrc.size() && mSrc[mPos] != c) {
				if (mSrc[mPos] == '\\' && mPos + 1 < mSrc.size()) {
					++mPos;
					switch (mSrc[mPos]) {
					case 'n': s += '\n'; break;
					case 't': s += '\t'; break;
					case 'r': s += '\r'; break;
					default: s += mSrc[mPos]; break;
					}
				} else {
					if (mSrc[mPos] == '\n')
						fail("string not closed");
					s += mSrc[mPos];
				}
				++mPos;
			}
			if (mPos >= mSrc.size())
				fail("string not closed");
			++mPos;
			d.setString(s);
		} else if (c == '$') {
			fail("DataLisp expressions '$(...)' are not supported");
		} else if ((c >= '0' && c <= '9') || c == '-' || c == '+' || c == '.') {
			const size_t s = mPos;
			bool isFloat   = false;
			if (c == '-' || c == '+')
				++mPos;
			while (mPos < mSrc.size()) {
				const char k = mSrc[mPos];
				if (k >= '0' && k <= '9') {
					++mPos;
				} else if (k == '.') {
					isFloat = true;
					++mPos;
				} else if (k == 'e' || k == 'E') {
					isFloat = true;
					++mPos;
					if (mPos < mSrc.size() && (mSrc[mPos] == '-' || mSrc[mPos] == '+'))
						++mPos;
				} else {
					break;
				}
			}
			const std::string tok = mSrc.substr(s, mPos - s);
			if (isFloat)
				d.setFloat(std::strtof(tok.c_str(), nullptr));
			else
				d.setInt(std::strtoll(tok.c_str(), nullptr, 10));
		} else {
			const std::string id = parseIdentifier();
			if (id == "true")
				d.setBool(true);
			else if (id == "false")
				d.setBool(false);
			else
				d.setString(id); // bare word
		}
	}

	const std::string& mSrc;
	size_t mPos = 0;
	int mLine	= 1;
};

inline std::vector<DataGroup> parseString(const std::string& s)
{
	Parser p(s);
	return p.parse();
}
inline std::vector<DataGroup> parseFile(const std::string& path)
{
	std::ifstream f(path, std::ios::binary);
	if (!f)
		throw ParseError("could not open '" + path + "'");
	std::stringstream ss;
	ss << f.rdbuf();
	const std::string src = ss.str();
	Parser p(src);
	return p.parse();
}
} // namespace DL
