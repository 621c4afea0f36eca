// cx.cuh -- minimal complex type shared by the Poisson kernels.
#pragma once

namespace insb {

template <typename T>
struct alignas(2 * sizeof(T)) Cx {  // 8-/16-byte aligned so that loads and stores are single 64-/128-bit accesses
    T x, y;
};
template <typename T>
__host__ __device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
    return Cx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
__host__ __device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) {
    return Cx<T>{a.x + b.x, a.y + b.y};
}
template <typename T>
__host__ __device__ __forceinline__ Cx<T> csub(Cx<T> a, Cx<T> b) {
    return Cx<T>{a.x - b.x, a.y - b.y};
}
template <typename T>
__host__ __device__ __forceinline__ Cx<T> cconj(Cx<T> a) {
    return Cx<T>{a.x, -a.y};
}


}  // namespace insb
