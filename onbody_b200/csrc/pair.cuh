/*
 * pair.cuh - the physics policies and the one-source-on-one-target pair functions shared by p2p.cu (list driven
 * block kernels) and pointwise.cu (fused traversal). See p2p.cu for the reference citations.
 */
#pragma once
#include "onb_internal.h"

namespace {

// ---------------------------------------------------------------------------------------------
// physics policies
// ---------------------------------------------------------------------------------------------
template <int PHYS> struct Phys;
template <> struct Phys<ONB_GRAV3D>     { static constexpr int PD = 3, SD = 1, OD = 3,  NF4 = 1; static constexpr bool F1 = true,  TR = false; };
template <> struct Phys<ONB_VORT3D>     { static constexpr int PD = 3, SD = 3, OD = 3,  NF4 = 2; static constexpr bool F1 = false, TR = false; };
template <> struct Phys<ONB_VORTGRAD3D> { static constexpr int PD = 3, SD = 3, OD = 12, NF4 = 2; static constexpr bool F1 = false, TR = false; };
template <> struct Phys<ONB_VORT2D>     { static constexpr int PD = 2, SD = 1, OD = 2,  NF4 = 1; static constexpr bool F1 = false, TR = false; };
template <> struct Phys<ONB_VORT2DTR>   { static constexpr int PD = 2, SD = 1, OD = 2,  NF4 = 1; static constexpr bool F1 = false, TR = true;  };

struct Tgt { float x, y, z, r2; };

// single-instruction SFU approximations (MUFU.RSQ / MUFU.RCP, ~1 ulp-level relative error 2^-22); the *_rn intrinsics
// are the correctly-rounded (multi-instruction, branchy) versions and are used only by the STRICT instantiations
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x)   { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
   // r2 = target radius squared (vort2dtr only)

// one source on one target. p0/p1/p2 are the packed planes' values for this source.
template <int PHYS, bool STRICT>
__device__ __forceinline__ void pair(const float4 p0, const float4 p1, const float p2, const Tgt& t, float* __restrict__ u) {
    if (PHYS == ONB_GRAV3D) {
        // p0 = (x,y,z,m)  p2 = r^2
        if (!STRICT) {
            const float dx = p0.x - t.x, dy = p0.y - t.y, dz = p0.z - t.z;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, p2)));
            const float ri = rsqrt_approx(r2);
            const float r3 = (p0.w * ri) * (ri * ri);
            u[0] = fmaf(r3, dx, u[0]); u[1] = fmaf(r3, dy, u[1]); u[2] = fmaf(r3, dz, u[2]);
        } else {
            const float dx = __fsub_rn(p0.x, t.x), dy = __fsub_rn(p0.y, t.y), dz = __fsub_rn(p0.z, t.z);
            float r3 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), p2);
            r3 = __fdiv_rn(p0.w, __fmul_rn(r3, __fsqrt_rn(r3)));
            u[0] = __fadd_rn(u[0], __fmul_rn(r3, dx)); u[1] = __fadd_rn(u[1], __fmul_rn(r3, dy)); u[2] = __fadd_rn(u[2], __fmul_rn(r3, dz));
        }
    } else if (PHYS == ONB_VORT3D) {
        // p0 = (x,y,z,r^2)  p1 = (wx,wy,wz,-)
        if (!STRICT) {
            const float dx = p0.x - t.x, dy = p0.y - t.y, dz = p0.z - t.z;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, p0.w)));
            const float ri = rsqrt_approx(r2);
            const float r3 = ri * (ri * ri);
            const float dxxw = fmaf(dz, p1.y, -(dy * p1.z));
            const float dyxw = fmaf(dx, p1.z, -(dz * p1.x));
            const float dzxw = fmaf(dy, p1.x, -(dx * p1.y));
            u[0] = fmaf(r3, dxxw, u[0]); u[1] = fmaf(r3, dyxw, u[1]); u[2] = fmaf(r3, dzxw, u[2]);
        } else {
            const float dx = __fsub_rn(p0.x, t.x), dy = __fsub_rn(p0.y, t.y), dz = __fsub_rn(p0.z, t.z);
            const float dsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float r2 = __fadd_rn(dsq, p0.w);
            const float r3 = __fdiv_rn(1.0f, __fmul_rn(r2, __fsqrt_rn(r2)));
            const float dxxw = __fsub_rn(__fmul_rn(dz, p1.y), __fmul_rn(dy, p1.z));
            const float dyxw = __fsub_rn(__fmul_rn(dx, p1.z), __fmul_rn(dz, p1.x));
            const float dzxw = __fsub_rn(__fmul_rn(dy, p1.x), __fmul_rn(dx, p1.y));
            u[0] = __fadd_rn(u[0], __fmul_rn(r3, dxxw)); u[1] = __fadd_rn(u[1], __fmul_rn(r3, dyxw)); u[2] = __fadd_rn(u[2], __fmul_rn(r3, dzxw));
        }
    } else if (PHYS == ONB_VORTGRAD3D) {
        // note the sign: d = target - source (onvortgrad3d.cpp:53-55)
        if (!STRICT) {
            const float dx = t.x - p0.x, dy = t.y - p0.y, dz = t.z - p0.z;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, p0.w)));
            const float ri = rsqrt_approx(r2);
            const float ri2 = ri * ri;
            const float r3 = ri * ri2;
            const float bbb = (-3.0f * r3) * ri2;
            float dxxw = fmaf(dz, p1.y, -(dy * p1.z));
            float dyxw = fmaf(dx, p1.z, -(dz * p1.x));
            float dzxw = fmaf(dy, p1.x, -(dx * p1.y));
            u[0] = fmaf(r3, dxxw, u[0]); u[1] = fmaf(r3, dyxw, u[1]); u[2] = fmaf(r3, dzxw, u[2]);
            dxxw *= bbb; dyxw *= bbb; dzxw *= bbb;
            u[3]  = fmaf(dx, dxxw, u[3]);
            u[4]  = fmaf(dx, dyxw, fmaf(p1.z, r3, u[4]));
            u[5]  = fmaf(dx, dzxw, fmaf(-p1.y, r3, u[5]));
            u[6]  = fmaf(dy, dxxw, fmaf(-p1.z, r3, u[6]));
            u[7]  = fmaf(dy, dyxw, u[7]);
            u[8]  = fmaf(dy, dzxw, fmaf(p1.x, r3, u[8]));
            u[9]  = fmaf(dz, dxxw, fmaf(p1.y, r3, u[9]));
            u[10] = fmaf(dz, dyxw, fmaf(-p1.x, r3, u[10]));
            u[11] = fmaf(dz, dzxw, u[11]);
        } else {
            const float dx = __fsub_rn(t.x, p0.x), dy = __fsub_rn(t.y, p0.y), dz = __fsub_rn(t.z, p0.z);
            const float dsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float r2 = __fadd_rn(dsq, p0.w);
            const float r3 = __fdiv_rn(1.0f, __fmul_rn(r2, __fsqrt_rn(r2)));
            const float bbb = __fmul_rn(__fmul_rn(-3.0f, r3), __fdiv_rn(1.0f, r2));
            float dxxw = __fsub_rn(__fmul_rn(dz, p1.y), __fmul_rn(dy, p1.z));
            float dyxw = __fsub_rn(__fmul_rn(dx, p1.z), __fmul_rn(dz, p1.x));
            float dzxw = __fsub_rn(__fmul_rn(dy, p1.x), __fmul_rn(dx, p1.y));
            u[0] = __fadd_rn(u[0], __fmul_rn(r3, dxxw)); u[1] = __fadd_rn(u[1], __fmul_rn(r3, dyxw)); u[2] = __fadd_rn(u[2], __fmul_rn(r3, dzxw));
            dxxw = __fmul_rn(dxxw, bbb); dyxw = __fmul_rn(dyxw, bbb); dzxw = __fmul_rn(dzxw, bbb);
            u[3]  = __fadd_rn(u[3],  __fmul_rn(dx, dxxw));
            u[4]  = __fadd_rn(u[4],  __fadd_rn(__fmul_rn(dx, dyxw), __fmul_rn(p1.z, r3)));
            u[5]  = __fadd_rn(u[5],  __fsub_rn(__fmul_rn(dx, dzxw), __fmul_rn(p1.y, r3)));
            u[6]  = __fadd_rn(u[6],  __fsub_rn(__fmul_rn(dy, dxxw), __fmul_rn(p1.z, r3)));
            u[7]  = __fadd_rn(u[7],  __fmul_rn(dy, dyxw));
            u[8]  = __fadd_rn(u[8],  __fadd_rn(__fmul_rn(dy, dzxw), __fmul_rn(p1.x, r3)));
            u[9]  = __fadd_rn(u[9],  __fadd_rn(__fmul_rn(dz, dxxw), __fmul_rn(p1.y, r3)));
            u[10] = __fadd_rn(u[10], __fsub_rn(__fmul_rn(dz, dyxw), __fmul_rn(p1.x, r3)));
            u[11] = __fadd_rn(u[11], __fmul_rn(dz, dzxw));
        }
    } else {
        // 2D: p0 = (x, y, r^2, strength); d = target - source; optional target radius
        if (!STRICT) {
            const float dx = t.x - p0.x, dy = t.y - p0.y;
            float r2c = fmaf(dx, dx, fmaf(dy, dy, p0.z));
            if (Phys<PHYS>::TR) r2c += t.r2;
            const float r2 = p0.w * rcp_approx(r2c);
            u[0] = fmaf(-r2, dy, u[0]); u[1] = fmaf(r2, dx, u[1]);
        } else {
            const float dx = __fsub_rn(t.x, p0.x), dy = __fsub_rn(t.y, p0.y);
            float r2c = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), p0.z);
            if (Phys<PHYS>::TR) r2c = __fadd_rn(r2c, t.r2);
            const float r2 = __fmul_rn(p0.w, __fdiv_rn(1.0f, r2c));
            u[0] = __fsub_rn(u[0], __fmul_rn(r2, dy)); u[1] = __fadd_rn(u[1], __fmul_rn(r2, dx));
        }
    }
}

// ACCUM = double: float contributions, fp64 accumulation (see p2p.cu k_p2p_lists_a64)
template <int PHYS, bool STRICT>
__device__ __forceinline__ void pair_a64(const float4 p0, const float4 p1, const float p2, const Tgt& t, double* __restrict__ acc) {
    constexpr int OD = Phys<PHYS>::OD;
    float cc[OD];
    #pragma unroll
    for (int d = 0; d < OD; ++d) cc[d] = 0.0f;
    pair<PHYS, STRICT>(p0, p1, p2, t, cc);             // 0 + x and fma(a, b, 0) are exact: cc holds the rounded float contributions
    #pragma unroll
    for (int d = 0; d < OD; ++d) acc[d] = __dadd_rn(acc[d], (double)cc[d]);
}

// one name for both accumulator types (pointwise.cu instantiates its traversal for float and for double outputs)
template <int PHYS, bool STRICT>
__device__ __forceinline__ void pair_acc(const float4 p0, const float4 p1, const float p2, const Tgt& t, float* __restrict__ acc) { pair<PHYS, STRICT>(p0, p1, p2, t, acc); }
template <int PHYS, bool STRICT>
__device__ __forceinline__ void pair_acc(const float4 p0, const float4 p1, const float p2, const Tgt& t, double* __restrict__ acc) { pair_a64<PHYS, STRICT>(p0, p1, p2, t, acc); }


}  // namespace

// ---------------------------------------------------------------------------------------------
// packed FP32x2 pair functions (fast arithmetic only).
// sm_100 adds add/mul/fma.f32x2 (SASS FADD2/FMUL2/FFMA2): one issue slot drives two FP32 lanes-ops per thread. The pair
// loop is issue-bound with scalar ops (ncu: issue active 90 %, FMA pipe 66 %), so each thread carries its targets two
// at a time in 64-bit register pairs; the source is broadcast into both halves. Per-lane arithmetic and its order are
// exactly those of the scalar fast path, so results are bit-identical to it.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 mk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 dup2(float a) { return mk2(a, a); }
__device__ __forceinline__ float lo2(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// two targets: positions (px) and negated positions (nx) packed, squared target radius packed (vort2dtr)
struct Tgt2 { f2 px, py, pz, nx, ny, nz, r2; };

template <int PHYS>
__device__ __forceinline__ void pair2(const float4 p0, const float4 p1, const float p2, const Tgt2& t, f2* __restrict__ u) {
    if (PHYS == ONB_GRAV3D) {
        const f2 dx = add2(dup2(p0.x), t.nx), dy = add2(dup2(p0.y), t.ny), dz = add2(dup2(p0.z), t.nz);
        const f2 r2 = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, dup2(p2))));
        const f2 ri = mk2(rsqrt_approx(lo2(r2)), rsqrt_approx(hi2(r2)));
        const f2 r3 = mul2(mul2(dup2(p0.w), ri), mul2(ri, ri));
        u[0] = fma2(r3, dx, u[0]); u[1] = fma2(r3, dy, u[1]); u[2] = fma2(r3, dz, u[2]);
    } else if (PHYS == ONB_VORT3D) {
        const f2 dx = add2(dup2(p0.x), t.nx), dy = add2(dup2(p0.y), t.ny), dz = add2(dup2(p0.z), t.nz);
        const f2 r2 = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, dup2(p0.w))));
        const f2 ri = mk2(rsqrt_approx(lo2(r2)), rsqrt_approx(hi2(r2)));
        const f2 r3 = mul2(ri, mul2(ri, ri));
        const f2 wx = dup2(p1.x), wy = dup2(p1.y), wz = dup2(p1.z), nwx = dup2(-p1.x), nwy = dup2(-p1.y), nwz = dup2(-p1.z);
        const f2 dxxw = fma2(dz, wy, mul2(dy, nwz));
        const f2 dyxw = fma2(dx, wz, mul2(dz, nwx));
        const f2 dzxw = fma2(dy, wx, mul2(dx, nwy));
        u[0] = fma2(r3, dxxw, u[0]); u[1] = fma2(r3, dyxw, u[1]); u[2] = fma2(r3, dzxw, u[2]);
    } else if (PHYS == ONB_VORTGRAD3D) {
        const f2 dx = add2(t.px, dup2(-p0.x)), dy = add2(t.py, dup2(-p0.y)), dz = add2(t.pz, dup2(-p0.z));
        const f2 r2 = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, dup2(p0.w))));
        const f2 ri = mk2(rsqrt_approx(lo2(r2)), rsqrt_approx(hi2(r2)));
        const f2 ri2 = mul2(ri, ri);
        const f2 r3 = mul2(ri, ri2);
        const f2 bbb = mul2(mul2(dup2(-3.0f), r3), ri2);
        const f2 wx = dup2(p1.x), wy = dup2(p1.y), wz = dup2(p1.z), nwx = dup2(-p1.x), nwy = dup2(-p1.y), nwz = dup2(-p1.z);
        f2 dxxw = fma2(dz, wy, mul2(dy, nwz));
        f2 dyxw = fma2(dx, wz, mul2(dz, nwx));
        f2 dzxw = fma2(dy, wx, mul2(dx, nwy));
        u[0] = fma2(r3, dxxw, u[0]); u[1] = fma2(r3, dyxw, u[1]); u[2] = fma2(r3, dzxw, u[2]);
        dxxw = mul2(dxxw, bbb); dyxw = mul2(dyxw, bbb); dzxw = mul2(dzxw, bbb);
        u[3]  = fma2(dx, dxxw, u[3]);
        u[4]  = fma2(dx, dyxw, fma2(wz, r3, u[4]));
        u[5]  = fma2(dx, dzxw, fma2(nwy, r3, u[5]));
        u[6]  = fma2(dy, dxxw, fma2(nwz, r3, u[6]));
        u[7]  = fma2(dy, dyxw, u[7]);
        u[8]  = fma2(dy, dzxw, fma2(wx, r3, u[8]));
        u[9]  = fma2(dz, dxxw, fma2(wy, r3, u[9]));
        u[10] = fma2(dz, dyxw, fma2(nwx, r3, u[10]));
        u[11] = fma2(dz, dzxw, u[11]);
    } else {
        const f2 dx = add2(t.px, dup2(-p0.x)), dy = add2(t.py, dup2(-p0.y));
        f2 r2c = fma2(dx, dx, fma2(dy, dy, dup2(p0.z)));
        if (Phys<PHYS>::TR) r2c = add2(r2c, t.r2);
        const f2 rc = mk2(rcp_approx(lo2(r2c)), rcp_approx(hi2(r2c)));
        const f2 r2 = mul2(dup2(p0.w), rc);
        const f2 nr2 = mul2(dup2(-p0.w), rc);
        u[0] = fma2(nr2, dy, u[0]); u[1] = fma2(r2, dx, u[1]);
    }
}
