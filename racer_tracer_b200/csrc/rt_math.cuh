// rt_math.cuh — small vector type templated on the scalar (float for the
// renderer, double for the f64 instantiation of the primary-visibility AOV),
// Philox4x32 and the uniform conversions of DESIGN.md "RNG streams".
#pragma once
#ifdef __CUDACC_RTC__
// NVRTC (scene-specialised kernels, rc_spec.cuh): no host headers
typedef unsigned char uint8_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

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
// Philox2x32-R (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
// SC'11): 64-bit counter, 32-bit key, 64 random bits per call.  One call costs
// R x (IMAD.WIDE.U32 + LOP3) with the key schedule in uniform registers, and
// 64 bits is exactly what one path event consumes, so nothing has to be cached
// between bounces.
// ---------------------------------------------------------------------------
#define PHILOX2_M 0xD256D193u
#define PHILOX2_W 0x9E3779B9u

template <int ROUNDS>
RT_HD uint2 philox2x32(uint32_t c0, uint32_t c1, uint32_t key) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p = (uint64_t)PHILOX2_M * c0;   // IMAD.WIDE.U32
        c0 = (uint32_t)(p >> 32) ^ (key + (uint32_t)r * PHILOX2_W) ^ c1;
        c1 = (uint32_t)p;
    }
    return make_uint2(c0, c1);
}

// Device form with the key schedule (key + r*W, r = 0..9) precomputed by the host into the
// kernel parameters: every round is IMAD.WIDE.U32 + one LOP3 with a constant-bank operand.
template <int ROUNDS>
RT_D uint2 philox2x32_ks(uint32_t c0, uint32_t c1, const uint32_t* ks) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p = (uint64_t)PHILOX2_M * c0;
        c0 = (uint32_t)(p >> 32) ^ ks[r] ^ c1;
        c1 = (uint32_t)p;
    }
    return make_uint2(c0, c1);
}

// Counter layout (DESIGN.md "RNG streams"):
//   c0 = pixel index (24 bits) | iteration j of a rejection loop << 24
//   c1 = sample index (24 bits) | bounce (6 bits) << 24 | stream tag << 30
//   key = low 32 bits of the seed XOR its high 32 bits
#define RT_TAG_PATH   0u   // bounce 0: x -> v jitter, y -> ray time; bounce b >= 1: the event's 64 bits
#define RT_TAG_PIXEL  1u   // sample = bounce = 0: x -> per-pixel u jitter (cpu.rs:35-36)
#define RT_TAG_LENS   2u   // j = 0: (x, y) -> direct lens sample; j >= 1: rejection iteration j
#define RT_TAG_REJECT 3u   // (bounce b, iteration j): three 21-bit uniforms of a rejection iteration

RT_HD uint32_t rt_ctr1(uint32_t sample, uint32_t bounce, uint32_t tag) { return sample | (bounce << 24) | (tag << 30); }

// 24-bit and 21-bit uniforms in [0,1): exactly representable in fp32 and f64
RT_HD float u24(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
RT_HD float u21(uint32_t w) { return (float)w * (1.0f / 2097152.0f); }  // w < 2^21
// three 21-bit uniforms out of 64 bits
RT_HD void u21x3(uint2 w, float& a, float& b, float& c) {
    a = u21(w.x >> 11); b = u21(w.y >> 11); c = u21(((w.x & 0x7FFu) << 10) | (w.y & 0x3FFu));
}
