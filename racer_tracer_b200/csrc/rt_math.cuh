// rt_math.cuh — small vector type templated on the scalar (float for the
// renderer, double for the f64 instantiation of the primary-visibility AOV),
// Philox4x32 and the uniform conversions of DESIGN.md "RNG streams".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__

template <typename T>
struct Vec3T {
    T x, y, z;
};
typedef Vec3T<float> vec3f;

template <typename T> RT_HD Vec3T<T> mk3(T x, T y, T z) { Vec3T<T> v; v.x = x; v.y = y; v.z = z; return v; }
template <typename T> RT_HD Vec3T<T> operator+(Vec3T<T> a, Vec3T<T> b) { return mk3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> RT_HD Vec3T<T> operator-(Vec3T<T> a, Vec3T<T> b) { return mk3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> RT_HD Vec3T<T> operator-(Vec3T<T> a) { return mk3<T>(-a.x, -a.y, -a.z); }
template <typename T> RT_HD Vec3T<T> operator*(Vec3T<T> a, Vec3T<T> b) { return mk3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> RT_HD Vec3T<T> operator*(T s, Vec3T<T> a) { return mk3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T> RT_HD Vec3T<T> operator*(Vec3T<T> a, T s) { return mk3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T> RT_HD T dot(Vec3T<T> a, Vec3T<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> RT_HD T length_squared(Vec3T<T> a) { return dot(a, a); }
template <typename T> RT_HD T axis_of(Vec3T<T> a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

RT_D float rt_sqrt(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
RT_D double rt_sqrt(double v) { return sqrt(v); }
RT_D float rt_rsqrt(float v) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
RT_D double rt_rsqrt(double v) { return 1.0 / sqrt(v); }
// single-instruction SFU approximations (<= 2 ulp); the render path does not
// need IEEE-rounded reciprocals / square roots
RT_D float fast_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
RT_D float fast_sqrt(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
RT_D float fast_rsqrt(float v) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
RT_D float rt_rcp(float v) { return fast_rcp(v); }
RT_D double rt_rcp(double v) { return 1.0 / v; }

template <typename T> RT_D Vec3T<T> unit_vector(Vec3T<T> a) { return a * rt_rsqrt(length_squared(a)); }

// vec3.rs:412-414
template <typename T> RT_D Vec3T<T> reflect(Vec3T<T> v, Vec3T<T> n) { return v - (T(2) * dot(v, n)) * n; }

// ---------------------------------------------------------------------------
// Philox4x32-R (Salmon et al., SC'11).  One call = four 32-bit words.
// ---------------------------------------------------------------------------
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

template <int ROUNDS>
RT_HD uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if (r > 0) { k0 += PHILOX_W0; k1 += PHILOX_W1; }
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;  // IMAD.WIDE.U32
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

RT_HD uint4 philox_rounds(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    if (rounds == 10) return philox4x32<10>(c0, c1, c2, c3, k0, k1);
    if (rounds == 7) return philox4x32<7>(c0, c1, c2, c3, k0, k1);
    for (int r = 0; r < rounds; ++r) {
        if (r > 0) { k0 += PHILOX_W0; k1 += PHILOX_W1; }
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

// stream tags (counter word 3, bits 24..31); see DESIGN.md "RNG streams"
#define RT_TAG_PIXEL  1u   // (pixel, 0, 0, tag): x -> per-pixel u jitter
#define RT_TAG_VJIT   3u   // (pixel, sample>>2, 0, tag): word sample&3 -> v jitter
#define RT_TAG_LENS   4u   // (pixel, sample, 0, tag|j): x,y -> lens disk; z -> time
#define RT_TAG_BOUNCE 5u   // (pixel, sample, (b+1)>>1, tag): b odd -> (x,y), b even -> (z,w)
#define RT_TAG_REJECT 6u   // (pixel, sample, b, tag|j): x,y,z of rejection iteration j

// 24-bit and 21-bit uniforms in [0,1): exactly representable in fp32 and f64
RT_HD float u24(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
RT_HD float u21(uint32_t w) { return (float)w * (1.0f / 2097152.0f); }  // w < 2^21
