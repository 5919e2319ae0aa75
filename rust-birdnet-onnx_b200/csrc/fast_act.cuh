// SiLU for the issue-bound epilogues: five instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL).
//
// `__fdividef(v, 1 + __expf(-v))` compiles to ~11: without -ftz, ex2.approx and the division get range-handling code
// (FSETP / scale by 0.5 / square ...) for denormal results that SiLU does not need:
//   -v*log2e < -126  -> ex2.ftz = 0   -> 1/(1+0) = 1  -> v        (the correct limit for large v)
//   -v*log2e >  128  -> ex2 = +inf    -> rcp(inf) = 0 -> -0       (the correct limit for very negative v)
// 1 + e >= 1 is never denormal.  Accuracy is that of the two approximate units (ex2: 2 ulp, rcp: 1 ulp), as before.
#pragma once
namespace bn {
__device__ __forceinline__ float silu_approx(float v) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return v * r;
}
__device__ __forceinline__ float sigmoid_approx(float v) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
}  // namespace bn
