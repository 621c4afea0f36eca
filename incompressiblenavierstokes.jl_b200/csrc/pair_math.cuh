// pair_math.cuh -- packed two-cell arithmetic shared by the marching stage kernels (stencil.cu,
// stage_fused.cu): every thread owns two x-adjacent cells and evaluates the face fluxes of
// src/operators.jl:378-393,556 on 2-wide values (packed FP32: FFMA2 / FADD2 / FMUL2 on sm_100).
#pragma once
#include "stencil_math.cuh"

namespace insb {

template <typename T, typename CT>
__device__ __forceinline__ T axpy_ct(T x, CT c, T k) {
    return (T)add_rn((CT)x, mul_rn(c, (CT)k));
}

template <typename T>
struct PairOf;
template <>
struct PairOf<float> {
    typedef float2 type;
};
template <>
struct PairOf<double> {
    typedef double2 type;
};
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 pbc(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 pmk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ double2 pfma(double2 a, double2 b, double2 c) {
    return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y));
}
__device__ __forceinline__ double2 padd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 pmul(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 psub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 pbc(double v) { return make_double2(v, v); }
__device__ __forceinline__ double2 pmk(double a, double b) { return make_double2(a, b); }

// separately rounded pair operations (never contracted into an FMA with a neighbouring operation)
__device__ __forceinline__ float2 psub_rn(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }  // exact a - b
__device__ __forceinline__ float2 pmul_rn(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ double2 psub_rn(double2 a, double2 b) { return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }
__device__ __forceinline__ double2 pmul_rn(double2 a, double2 b) { return make_double2(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)); }

// diagonal face flux: lo, hi = u_i on the two sides of the face, mnrd = -nu/Δ_i at the face
template <typename P>
__device__ __forceinline__ P flux_diag(P lo, P hi, P mnrd, P quarter) {
    P m = padd(lo, hi);
    P d = psub(hi, lo);
    return pfma(mnrd, d, pmul(pmul(m, m), quarter));
}
// off-diagonal face flux: a0, a1 = u_i at J, J+e_j; b0, b1 = u_j at J, J+e_i;
// hw2, hw1 = A2_i/2, A1_i/2; mnrd = -nu/Δu_j at the face
template <typename P>
__device__ __forceinline__ P flux_off(P a0, P a1, P b0, P b1, P hw2, P hw1, P mnrd) {
    P s = padd(a0, a1);
    P w = pfma(hw1, b1, pmul(hw2, b0));
    P d = psub(a1, a0);
    return pfma(s, w, pmul(mnrd, d));
}

// v (+)= c * k on pairs with separately rounded multiply and add, like the reference broadcast
// `u + Δt*A[i,j]*k` (step_explicit_runge_kutta.jl:28-32); CT is the coefficient/evaluation type
template <typename T, typename CT>
struct PairAxpy {
    typedef typename PairOf<T>::type P;
    static __device__ __forceinline__ P run(P x, CT c, P k) {
        return pmk(axpy_ct<T, CT>(x.x, c, k.x), axpy_ct<T, CT>(x.y, c, k.y));
    }
};
template <>
struct PairAxpy<float, float> {
    static __device__ __forceinline__ float2 run(float2 x, float c, float2 k) {
        return __fadd2_rn(x, __fmul2_rn(make_float2(c, c), k));
    }
};

}  // namespace insb
