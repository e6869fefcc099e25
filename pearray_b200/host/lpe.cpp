// Light path expressions: the `:lpe '...'` strings of (output (channel ...)) blocks compiled to a dense DFA
// table that the device walks one token per path vertex.
//
// Grammar and matching semantics follow the reference (src/core/path/LPE_Parser.cpp:65-298, LPE_RegState.h:39-78,
// LPE_Automaton.h:17-33):
//   full    := 'C' expr?                      every expression starts at the camera
//   expr    := term+
//   term    := (token | '(' expr ')' | '[' term+ ']') op?          [..] is the union of its terms; '[^' is rejected
//   token   := D | S | E | L | B | R | T | .  |  '<' type [','] event [[','] '"' label '"'] '>'
//   op      := '*' | '+' | '?' | '{' n '}' | '{' n ',' m '}'         {n,m} with m < n is an error; {0} = '*' (repeatLast(0, 0))
// A path token is (ScatteringType, ScatteringEvent) in {Camera, Emissive, Refraction, Reflection, Background} x {Diffuse,
// Specular, None} (LightPathToken.h:6-20).  Expression tokens match classes of them: D / S any scattering (R or T) with that
// event, '.' any scattering, E emissive, B background, L either, R / T one scattering type, any event.  Labelled tokens
// (<R,D,"name">) only match path tokens carrying that label; the `direct` integrator never labels its tokens, so here they
// match nothing.  The construction is this file's own (Thompson NFA over 15-bit symbol classes, subset construction), not the
// reference's RegExpr machinery; parity is on the match results, pinned by the reference's own cases (src/tests/lpe.cpp).
#include "prh.h"

#include <map>
#include <set>

namespace PR {
namespace {
constexpr int NTYPE = 5, NEVENT = 3, NSYM = NTYPE * NEVENT; // symbol = type * 3 + event
enum { T_CAMERA = 0, T_EMISSIVE = 1, T_REFRACTION = 2, T_REFLECTION = 3, T_BACKGROUND = 4 };
uint32 classMask(char type, char event)
{
	uint32 tm = 0;
	switch (type) { // Token::match(ScatteringType), LPE_RegState.h:44-64
	case 'C': tm = 1u << T_CAMERA; break;
	case 'E': tm = 1u << T_EMISSIVE; break;
	case 'B': tm = 1u << T_BACKGROUND; break;
	case 'L': tm = (1u << T_EMISSIVE) | (1u << T_BACKGROUND); break;
	case 'R': tm = 1u << T_REFLECTION; break;
	case 'T': tm = 1u << T_REFRACTION; break;
	case '.': tm = (1u << T_REFLECTION) | (1u << T_REFRACTION); break;
	}
	uint32 em = 0;
	switch (event) { // Token::match(ScatteringEvent), :66-78
	case 'D': em = 1; break;
	case 'S': em = 2; break;
	case '.': em = 7; break;
	}
	uint32 m = 0;
	for (int t = 0; t < NTYPE; ++t)
		for (int e = 0; e < NEVENT; ++e)
			if ((tm >> t & 1) && (em >> e & 1))
				m |= 1u << (t * NEVENT + e);
	return m;
}

struct NFA { // fragment with one entry and one exit state; edges carry a symbol mask, 0 = epsilon
	struct Edge {
		int to;
		uint32 mask;
	};
	std::vector<std::vector<Edge>> st;
	int add()
	{
		st.emplace_back();
		return (int)st.size() - 1;
	}
};
struct Frag {
	int in, out;
};

class Parser {
public:
	Parser(const std::string& s, NFA& n)
		: mS(s)
		, mN(n)
	{
	}
	bool error = false;
	Frag full()
	{
		if (cur() != 'C') {
			error = true;
			return {};
		}
		++mP;
		Frag f = literal(classMask('C', '.'));
		if (!eos()) {
			Frag e = expr();
			f	   = concat(f, e);
		}
		if (!eos())
			error = true; // a stray ')' or ']'
		return f;
	}

private:
	const std::string& mS;
	NFA& mN;
	size_t mP = 0;
	bool eos() const { return mP >= mS.size(); }
	char cur() const { return eos() ? '\0' : mS[mP]; }
	Frag literal(uint32 mask)
	{
		Frag f{ mN.add(), mN.add() };
		if (mask) // a class that matches nothing leaves the two states unconnected
			mN.st[f.in].push_back({ f.out, mask });
		return f;
	}
	Frag concat(Frag a, Frag b)
	{
		mN.st[a.out].push_back({ b.in, 0 });
		return { a.in, b.out };
	}
	Frag alternate(Frag a, Frag b)
	{
		Frag f{ mN.add(), mN.add() };
		mN.st[f.in].push_back({ a.in, 0 });
		mN.st[f.in].push_back({ b.in, 0 });
		mN.st[a.out].push_back({ f.out, 0 });
		mN.st[b.out].push_back({ f.out, 0 });
		return f;
	}
	Frag clone(Frag a)
	{ // copy of the sub-automaton reachable from a.in (fragments are closed: nothing leaves them except through a.out)
		std::map<int, int> m;
		std::vector<int> todo{ a.in };
		m[a.in] = mN.add();
		while (!todo.empty()) {
			const int s = todo.back();
			todo.pop_back();
			const std::vector<NFA::Edge> edges = mN.st[s];
			for (const NFA::Edge& e : edges) {
				if (!m.count(e.to)) {
					m[e.to] = mN.add();
					todo.push_back(e.to);
				}
				mN.st[m[s]].push_back({ m[e.to], e.mask });
			}
		}
		if (!m.count(a.out))
			m[a.out] = mN.add();
		return { m[a.in], m[a.out] };
	}
	Frag optional(Frag a)
	{
		Frag f{ mN.add(), mN.add() };
		mN.st[f.in].push_back({ a.in, 0 });
		mN.st[f.in].push_back({ f.out, 0 });
		mN.st[a.out].push_back({ f.out, 0 });
		return f;
	}
	Frag star(Frag a)
	{
		Frag f{ mN.add(), mN.add() };
		mN.st[f.in].push_back({ a.in, 0 });
		mN.st[f.in].push_back({ f.out, 0 });
		mN.st[a.out].push_back({ a.in, 0 });
		mN.st[a.out].push_back({ f.out, 0 });
		return f;
	}
	Frag repeat(Frag a, uint32 mn, uint32 mx)
	{ // RegExpr::repeatLast(min, max), LPE_RegExpr.cpp:96-146: min mandatory copies, then max == 0 ? X* : (max - min) optional ones
		Frag acc{ -1, -1 };
		auto append = [&](Frag f) { acc = acc.in < 0 ? f : concat(acc, f); };
		for (uint32 i = 0; i < mn; ++i)
			append(clone(a));
		if (mx == 0)
			append(star(clone(a)));
		else
			for (uint32 i = mn; i < mx; ++i)
				append(optional(clone(a)));
		if (acc.in < 0) { // {0,0} cannot happen (max == 0 is the star); keep a valid empty fragment anyway
			acc = Frag{ mN.add(), mN.add() };
			mN.st[acc.in].push_back({ acc.out, 0 });
		}
		return acc;
	}
	Frag expr()
	{
		Frag f = term();
		while (!eos() && !error && cur() != ')')
			f = concat(f, term());
		return f;
	}
	Frag term()
	{
		Frag f{};
		const char c = cur();
		if (c == 'D' || c == 'S' || c == 'E' || c == 'L' || c == 'B' || c == 'R' || c == 'T' || c == '.' || c == '<') {
			f = token();
		} else if (c == '(') {
			++mP;
			f = expr();
			if (cur() != ')')
				error = true;
			++mP;
		} else if (c == '[') {
			++mP;
			if (cur() == '^') { // "Negation in union groups currently not supported!", LPE_Parser.cpp:139-144
				error = true;
				return literal(0);
			}
			f = term();
			while (!eos() && !error && cur() != ']')
				f = alternate(f, term());
			if (cur() != ']')
				error = true;
			++mP;
		} else {
			error = true;
			++mP;
			return literal(0);
		}
		if (error)
			return f;
		return op(f);
	}
	Frag token()
	{
		const char c = cur();
		++mP;
		switch (c) {
		case 'D': return literal(classMask('.', 'D'));
		case 'S': return literal(classMask('.', 'S'));
		case 'E': return literal(classMask('E', '.'));
		case 'L': return literal(classMask('L', '.'));
		case 'B': return literal(classMask('B', '.'));
		case '.': return literal(classMask('.', '.'));
		case 'R': return literal(classMask('R', '.'));
		case 'T': return literal(classMask('T', '.'));
		default: break;
		}
		// '<' type [','] event [[','] '"' label '"'] '>'
		const char t = cur();
		if (!(t == 'E' || t == 'L' || t == 'B' || t == 'R' || t == 'T' || t == '.')) {
			error = true;
			return literal(0);
		}
		++mP;
		if (cur() == ',')
			++mP;
		const char e = cur();
		if (!(e == 'D' || e == 'S' || e == '.')) {
			error = true;
			return literal(0);
		}
		++mP;
		bool labelled = false;
		if (cur() == '"' || cur() == ',') {
			if (cur() == ',')
				++mP;
			if (cur() != '"')
				error = true;
			++mP;
			std::string lbl;
			while (!eos() && cur() != '"')
				lbl += mS[mP++];
			if (cur() != '"')
				error = true;
			++mP;
			labelled = !lbl.empty(); // Token::Label.empty() tokens are the unlabelled kind, LPE_Automaton.cpp:46-69
		}
		if (cur() != '>')
			error = true;
		++mP;
		return literal(labelled ? 0u : classMask(t, e)); // path tokens of the `direct` integrator carry no label
	}
	Frag op(Frag f)
	{
		const char c = cur();
		if (c == '*') {
			++mP;
			return repeat(f, 0, 0);
		}
		if (c == '+') {
			++mP;
			return repeat(f, 1, 0);
		}
		if (c == '?') {
			++mP;
			return repeat(f, 0, 1);
		}
		if (c == '{') {
			++mP;
			auto integer = [&]() -> uint32 {
				std::string n;
				while (!eos() && std::isdigit((unsigned char)cur()))
					n += mS[mP++];
				if (n.empty()) {
					error = true;
					return 0;
				}
				return (uint32)std::stoul(n);
			};
			const uint32 n = integer();
			uint32 l	   = n;
			if (cur() == ',') {
				++mP;
				l = integer();
			}
			if (cur() != '}')
				error = true;
			++mP;
			if (l < n)
				error = true;
			if (error)
				return f;
			return repeat(f, n, l);
		}
		return f;
	}
};
} // namespace

LPEAutomaton compileLPE(const std::string& expression)
{
	LPEAutomaton A;
	NFA nfa;
	Parser p(expression, nfa);
	const Frag f = p.full();
	if (p.error)
		return A;
	// subset construction
	auto closure = [&](std::set<int> s) {
		std::vector<int> todo(s.begin(), s.end());
		while (!todo.empty()) {
			const int x = todo.back();
			todo.pop_back();
			for (const NFA::Edge& e : nfa.st[x])
				if (e.mask == 0 && s.insert(e.to).second)
					todo.push_back(e.to);
		}
		return s;
	};
	std::map<std::set<int>, uint32> ids;
	std::vector<std::set<int>> sets;
	auto idOf = [&](const std::set<int>& s) -> int {
		auto it = ids.find(s);
		if (it != ids.end())
			return (int)it->second;
		if (sets.size() >= 254)
			return -1;
		ids[s] = (uint32)sets.size();
		sets.push_back(s);
		return (int)sets.size() - 1;
	};
	idOf(closure({ f.in }));
	for (size_t i = 0; i < sets.size(); ++i) {
		for (int sym = 0; sym < NSYM; ++sym) {
			std::set<int> to;
			for (int x : sets[i])
				for (const NFA::Edge& e : nfa.st[x])
					if (e.mask >> sym & 1)
						to.insert(e.to);
			uint8_t next = LPE_REJECT;
			if (!to.empty()) {
				const int id = idOf(closure(to));
				if (id < 0)
					return A; // more than 254 DFA states: refused (the device table stores states in bytes)
				next = (uint8_t)id;
			}
			if (A.next.size() < (i + 1) * NSYM)
				A.next.resize((i + 1) * NSYM, LPE_REJECT);
			A.next[i * NSYM + sym] = next;
		}
	}
	A.next.resize(sets.size() * NSYM, LPE_REJECT);
	A.final.resize(sets.size());
	for (size_t i = 0; i < sets.size(); ++i)
		A.final[i] = sets[i].count(f.out) ? 1 : 0;
	A.stateCount = (uint32)sets.size();
	A.valid		 = true;
	return A;
}

bool LPEAutomaton::match(const std::vector<std::pair<int, int>>& tokens) const
{ // Automaton::match, LPE_Automaton.h:17-33
	if (!valid)
		return false;
	uint8_t s = 0;
	for (const auto& t : tokens) {
		if (t.first < 0 || t.first >= NTYPE || t.second < 0 || t.second >= NEVENT)
			return false;
		s = next[(size_t)s * NSYM + t.first * NEVENT + t.second];
		if (s == LPE_REJECT)
			return false;
	}
	return final[s] != 0;
}
} // namespace PR
