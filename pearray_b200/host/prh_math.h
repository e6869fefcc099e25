// Small fp32 vector / matrix helpers for the host side (the reference uses Eigen; Eigen is not
// available here, so the handful of operations the loader needs are written out).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace PR {
using uint8	 = uint8_t;
using uint32 = uint32_t;
using uint64 = uint64_t;
using int32	 = int32_t;
using int64	 = int64_t;

constexpr float PR_EPSILON	   = std::numeric_limits<float>::epsilon();
constexpr float PR_INF		   = std::numeric_limits<float>::infinity();
constexpr uint32 PR_INVALID_ID = 0xFFFFFFFFu;
constexpr float PR_PI		   = 3.14159265358979323846f;
constexpr float PR_INV_PI	   = 0.31830988618379067154f;
constexpr float PR_INV_2_PI	   = 0.15915494309189533577f;
constexpr float PR_DEG2RAD	   = PR_PI / 180.0f;
constexpr float PR_RAD2DEG	   = 180.0f * PR_INV_PI;

struct Vector2f {
	float x = 0, y = 0;
	Vector2f() = default;
	Vector2f(float x_, float y_)
		: x(x_)
		, y(y_)
	{
	}
	float operator()(int i) const { return i == 0 ? x : y; }
};

struct Vector3f {
	float x = 0, y = 0, z = 0;
	Vector3f() = default;
	Vector3f(float x_, float y_, float z_)
		: x(x_)
		, y(y_)
		, z(z_)
	{
	}
	float operator()(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	Vector3f operator+(const Vector3f& o) const { return { x + o.x, y + o.y, z + o.z }; }
	Vector3f operator-(const Vector3f& o) const { return { x - o.x, y - o.y, z - o.z }; }
	Vector3f operator-() const { return { -x, -y, -z }; }
	Vector3f operator*(float f) const { return { x * f, y * f, z * f }; }
	Vector3f operator/(float f) const { return { x / f, y / f, z / f }; }
	float dot(const Vector3f& o) const { return x * o.x + y * o.y + z * o.z; }
	Vector3f cross(const Vector3f& o) const { return { y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x }; }
	float squaredNorm() const { return dot(*this); }
	float norm() const { return std::sqrt(squaredNorm()); }
	Vector3f normalized() const
	{
		const float n = norm();
		return n > 0 ? *this / n : *this;
	}
	void normalize() { *this = normalized(); }
	static Vector3f Zero() { return {}; }
};
inline Vector3f operator*(float f, const Vector3f& v) { return v * f; }

// SpectralBlob: four wavelengths/values; [0] is the hero (reference src/core/spectral/SpectralBlob.h)
struct SpectralBlob {
	float v[4] = { 0, 0, 0, 0 };
	SpectralBlob() = default;
	explicit SpectralBlob(float f) { v[0] = v[1] = v[2] = v[3] = f; }
	SpectralBlob(float a, float b, float c, float d)
	{
		v[0] = a;
		v[1] = b;
		v[2] = c;
		v[3] = d;
	}
	float& operator[](int i) { return v[i]; }
	float operator[](int i) const { return v[i]; }
	float& operator()(int i) { return v[i]; }
	float operator()(int i) const { return v[i]; }
	SpectralBlob operator*(const SpectralBlob& o) const { return { v[0] * o.v[0], v[1] * o.v[1], v[2] * o.v[2], v[3] * o.v[3] }; }
	SpectralBlob operator*(float f) const { return { v[0] * f, v[1] * f, v[2] * f, v[3] * f }; }
	SpectralBlob operator+(const SpectralBlob& o) const { return { v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2], v[3] + o.v[3] }; }
	SpectralBlob& operator+=(const SpectralBlob& o)
	{
		*this = *this + o;
		return *this;
	}
	float mean() const { return (v[0] + v[1] + v[2] + v[3]) / 4; }
	float sum() const { return v[0] + v[1] + v[2] + v[3]; }
	static SpectralBlob Zero() { return SpectralBlob(0.0f); }
	static SpectralBlob Ones() { return SpectralBlob(1.0f); }
};

struct Matrix3f { // row-major
	float m[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
	float operator()(int r, int c) const { return m[r * 3 + c]; }
	float& operator()(int r, int c) { return m[r * 3 + c]; }
	Vector3f operator*(const Vector3f& v) const
	{
		return { m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z };
	}
	Vector3f col(int c) const { return { m[c], m[3 + c], m[6 + c] }; }
	float determinant() const
	{
		return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
	}
	Matrix3f transpose() const
	{
		Matrix3f r;
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j)
				r(i, j) = (*this)(j, i);
		return r;
	}
	Matrix3f inverse() const // cofactor expansion (what Eigen does for fixed 3x3)
	{
		const float det = determinant();
		const float id	= 1.0f / det;
		Matrix3f r;
		r.m[0] = (m[4] * m[8] - m[5] * m[7]) * id;
		r.m[1] = (m[2] * m[7] - m[1] * m[8]) * id;
		r.m[2] = (m[1] * m[5] - m[2] * m[4]) * id;
		r.m[3] = (m[5] * m[6] - m[3] * m[8]) * id;
		r.m[4] = (m[0] * m[8] - m[2] * m[6]) * id;
		r.m[5] = (m[2] * m[3] - m[0] * m[5]) * id;
		r.m[6] = (m[3] * m[7] - m[4] * m[6]) * id;
		r.m[7] = (m[1] * m[6] - m[0] * m[7]) * id;
		r.m[8] = (m[0] * m[4] - m[1] * m[3]) * id;
		return r;
	}
	Matrix3f operator*(const Matrix3f& o) const
	{
		Matrix3f r;
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j) {
				float s = 0;
				for (int k = 0; k < 3; ++k)
					s += (*this)(i, k) * o(k, j);
				r(i, j) = s;
			}
		return r;
	}
};

// Affine transform (reference Transformf = Eigen::Affine3f after makeAffine())
struct Transformf {
	Matrix3f L;	  // linear part
	Vector3f T;	  // translation
	const Matrix3f& linear() const { return L; }
	const Vector3f& translation() const { return T; }
	Vector3f operator*(const Vector3f& p) const { return L * p + T; }
	Transformf operator*(const Transformf& o) const
	{
		Transformf r;
		r.L = L * o.L;
		r.T = L * o.T + T;
		return r;
	}
	Transformf inverse() const
	{
		Transformf r;
		r.L = L.inverse();
		r.T = -(r.L * T);
		return r;
	}
	static Transformf Identity() { return {}; }
	void to34(float* out) const
	{
		for (int r = 0; r < 3; ++r) {
			out[r * 4 + 0] = L(r, 0);
			out[r * 4 + 1] = L(r, 1);
			out[r * 4 + 2] = L(r, 2);
			out[r * 4 + 3] = T[r];
		}
	}
	// Scale part as Eigen computeRotationScaling would give for rotation*diag(scale): column norms
	Vector3f scaling() const { return { L.col(0).norm(), L.col(1).norm(), L.col(2).norm() }; }
};

inline Matrix3f quaternionToMatrix(float w, float x, float y, float z)
{
	const float n = std::sqrt(w * w + x * x + y * y + z * z);
	w /= n;
	x /= n;
	y /= n;
	z /= n;
	Matrix3f r;
	r(0, 0) = 1 - 2 * (y * y + z * z);
	r(0, 1) = 2 * (x * y - w * z);
	r(0, 2) = 2 * (x * z + w * y);
	r(1, 0) = 2 * (x * y + w * z);
	r(1, 1) = 1 - 2 * (x * x + z * z);
	r(1, 2) = 2 * (y * z - w * x);
	r(2, 0) = 2 * (x * z - w * y);
	r(2, 1) = 2 * (y * z + w * x);
	r(2, 2) = 1 - 2 * (x * x + y * y);
	return r;
}

struct BoundingBox {
	Vector3f lo{ PR_INF, PR_INF, PR_INF }, hi{ -PR_INF, -PR_INF, -PR_INF };
	void combine(const Vector3f& p)
	{
		for (int i = 0; i < 3; ++i) {
			lo[i] = std::min(lo[i], p[i]);
			hi[i] = std::max(hi[i], p[i]);
		}
	}
	void combine(const BoundingBox& b)
	{
		combine(b.lo);
		combine(b.hi);
	}
	bool valid() const { return lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z; }
	float halfArea() const
	{
		const Vector3f d = hi - lo;
		return d.x * d.y + d.y * d.z + d.z * d.x;
	}
	Vector3f center() const { return (lo + hi) * 0.5f; }
};
} // namespace PR
