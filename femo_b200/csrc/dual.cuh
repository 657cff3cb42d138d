// dual.cuh -- forward-mode dual numbers for exact Gateaux derivatives of element residuals.
// Dual<N, T>: value of type T plus N directional derivatives of type T.  T = double gives first
// derivatives; T = Dual<M> nests (derivatives of derivatives), which the mesh-motion family needs
// because its residual already contains derivative(P, uhat, v) (motor_pde.py:174).
#pragma once
#include <cuda_runtime.h>

namespace femo {

template <int N, class T = double>
struct Dual {
    T v;
    T d[N];
    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(double a) : v(a) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = T(0.0);
    }
};

// Dual with the given value and zero derivatives
template <int N, class T>
__device__ __forceinline__ Dual<N, T> lift(const T &a) {
    Dual<N, T> r;
    r.v = a;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = T(0.0);
    return r;
}

#define FEMO_DUAL template <int N, class T> __device__ __forceinline__
FEMO_DUAL Dual<N, T> operator+(const Dual<N, T> &a, const Dual<N, T> &b) { Dual<N, T> r; r.v = a.v + b.v; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
FEMO_DUAL Dual<N, T> operator-(const Dual<N, T> &a, const Dual<N, T> &b) { Dual<N, T> r; r.v = a.v - b.v; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
FEMO_DUAL Dual<N, T> operator-(const Dual<N, T> &a) { Dual<N, T> r; r.v = -a.v; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
FEMO_DUAL Dual<N, T> operator*(const Dual<N, T> &a, const Dual<N, T> &b) { Dual<N, T> r; r.v = a.v * b.v; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
FEMO_DUAL Dual<N, T> operator/(const Dual<N, T> &a, const Dual<N, T> &b) { Dual<N, T> r; const T ib = 1.0 / b.v; r.v = a.v * ib; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
FEMO_DUAL Dual<N, T> operator*(double a, const Dual<N, T> &b) { Dual<N, T> r; r.v = a * b.v; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i]; return r; }
FEMO_DUAL Dual<N, T> operator*(const Dual<N, T> &b, double a) { return a * b; }
FEMO_DUAL Dual<N, T> operator+(const Dual<N, T> &a, double b) { Dual<N, T> r = a; r.v = r.v + b; return r; }
FEMO_DUAL Dual<N, T> operator+(double b, const Dual<N, T> &a) { return a + b; }
FEMO_DUAL Dual<N, T> operator-(const Dual<N, T> &a, double b) { Dual<N, T> r = a; r.v = r.v - b; return r; }
FEMO_DUAL Dual<N, T> operator/(double a, const Dual<N, T> &b) { Dual<N, T> r; const T ib = 1.0 / b.v; r.v = a * ib; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = -1.0 * (r.v * b.d[i]) * ib; return r; }
#undef FEMO_DUAL

// transcendental functions for first-order duals over double
template <int N> __device__ __forceinline__ Dual<N> dsqrt(const Dual<N> &a) { Dual<N> r; r.v = sqrt(a.v); const double s = (r.v > 0.0) ? 0.5 / r.v : 0.0; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
template <int N> __device__ __forceinline__ Dual<N> dexp(const Dual<N> &a) { Dual<N> r; r.v = exp(a.v); _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i]; return r; }
template <int N> __device__ __forceinline__ Dual<N> dpow(const Dual<N> &a, double p) { Dual<N> r; r.v = pow(a.v, p); const double s = (a.v > 0.0) ? p * pow(a.v, p - 1.0) : 0.0; _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
__device__ __forceinline__ double dsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ double dexp(double a) { return exp(a); }
__device__ __forceinline__ double dpow(double a, double p) { return pow(a, p); }
__device__ __forceinline__ double valof(double a) { return a; }
template <int N> __device__ __forceinline__ double valof(const Dual<N> &a) { return a.v; }

}  // namespace femo
