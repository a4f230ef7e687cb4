// ddsum.cuh — double-double accumulation of sums of squares: column norms that are (nearly) correctly rounded, like the
// extended-precision dnrm2 kernel of the LAPACK build the pivot rule is validated against (geqp3.cu, geqp3_blocked.cu).
#pragma once
namespace rsvd {
struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_add_sq(dd a, double x) {   // a += x*x, error-free product + two-sum
    double p = x * x;
    double e = fma(x, x, -p);
    double s = a.hi + p;
    double bb = s - a.hi;
    double err = (a.hi - (s - bb)) + (p - bb);
    a.hi = s; a.lo += err + e;
    return a;
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    double s = a.hi + b.hi;
    double bb = s - a.hi;
    double err = (a.hi - (s - bb)) + (b.hi - bb);
    dd r; r.hi = s; r.lo = a.lo + b.lo + err;
    return r;
}
__device__ __forceinline__ dd dd_warp_sum(dd a) {
    for (int o = 16; o > 0; o >>= 1) {
        dd b;
        b.hi = __shfl_xor_sync(0xffffffffu, a.hi, o);
        b.lo = __shfl_xor_sync(0xffffffffu, a.lo, o);
        a = dd_add(a, b);
    }
    return a;
}
__device__ __forceinline__ double dd_sqrt(dd a) { return sqrt(a.hi + a.lo); }

}  // namespace rsvd
