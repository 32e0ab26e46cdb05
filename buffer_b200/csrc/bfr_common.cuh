// bfr_common.cuh — device helpers shared by the sm_100a kernels of the BUFFER correspondence-and-pose back end.
//
// Arithmetic contract (DESIGN.md §"bit-exactness"): every floating-point operation that takes part in a result that
// is compared bit-for-bit with the CPU oracle is spelled with an explicit round-to-nearest intrinsic
// (__fmaf_rn/__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn or the packed fma.rn.f32x2), so nvcc can neither
// contract nor re-associate it.  Packed FFMA2/FADD2 lanes are IEEE-identical to their scalar forms.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define BFR_DEVINL __device__ __forceinline__

namespace bfr {

// ---- packed f32x2 (sm_100a FFMA2 / FADD2) ----------------------------------------------------------------------
typedef unsigned long long f32x2;
BFR_DEVINL f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
BFR_DEVINL void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
BFR_DEVINL f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
BFR_DEVINL f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
BFR_DEVINL f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
BFR_DEVINL f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---- three-input max (FMNMX3) and warp-wide float max (CREDUX), both new on sm_100 -------------------------------
BFR_DEVINL float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
BFR_DEVINL float warp_max(float v) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }

// ---- order-preserving float -> uint key, packed (key << 32 | ~index) so that max() breaks ties to the lowest index
BFR_DEVINL uint32_t float_key(float v) { uint32_t b = __float_as_uint(v); return b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u); }
BFR_DEVINL float key_float(uint32_t k) { uint32_t b = k ^ ((k & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu); return __uint_as_float(b); }
BFR_DEVINL unsigned long long pack_best(uint32_t key, uint32_t idx) { return ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - idx); }
BFR_DEVINL void red_max_u64(unsigned long long* addr, unsigned long long v) { asm volatile("red.global.max.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory"); }
// predicated form: publishes only in lanes where a == b (no branch, so independent reductions can interleave)
BFR_DEVINL void red_max_u64_if_eq(unsigned long long* addr, unsigned long long v, float a, float b)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.f32 q, %2, %3;\n\t@q red.global.max.u64 [%0], %1;\n\t}" ::"l"(addr), "l"(v), "f"(a), "f"(b) : "memory");
}
BFR_DEVINL uint32_t packed_index(unsigned long long p) { return 0xFFFFFFFFu - (uint32_t)(p & 0xFFFFFFFFull); }

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk) --------------------------------------------------------------
BFR_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
BFR_DEVINL void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
BFR_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
BFR_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
BFR_DEVINL void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
BFR_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`
BFR_DEVINL void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
BFR_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- Philox4x32-10: the counter-based stream shared with the CPU oracle (north_star item 2) ----------------------
BFR_DEVINL void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// minimal set of hypothesis h of pair `pair_id`: counter (h, pair_id, 0, 0), index = mulhi(u, K) (with replacement)
BFR_DEVINL void sample3(uint64_t seed, uint32_t pair_id, uint32_t h, uint32_t K, uint32_t idx[3])
{
    uint32_t r[4];
    philox4x32_10(h, pair_id, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    idx[0] = __umulhi(r[0], K); idx[1] = __umulhi(r[1], K); idx[2] = __umulhi(r[2], K);
}

// ---- 3x3 Kabsch rotation (closed-form SVD with the reflection fix inline) ------------------------------------------
// Replaces torch.svd(H.cpu()) + det fix of rigid_transform_3d (reference models/BUFFER.py:455-460) and Eigen::umeyama
// inside Open3D's point-to-point estimator (models/BUFFER.py:320).  One-sided Jacobi, 4 cyclic sweeps, then
// R = v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T from the two largest singular pairs.
BFR_DEVINL float dot3(const float a[3], const float b[3]) { return __fmaf_rn(a[2], b[2], __fmaf_rn(a[1], b[1], __fmul_rn(a[0], b[0]))); }
BFR_DEVINL void cross3(const float a[3], const float b[3], float c[3])
{
    c[0] = __fmaf_rn(a[1], b[2], -__fmul_rn(a[2], b[1]));
    c[1] = __fmaf_rn(a[2], b[0], -__fmul_rn(a[0], b[2]));
    c[2] = __fmaf_rn(a[0], b[1], -__fmul_rn(a[1], b[0]));
}
BFR_DEVINL void gram_schmidt2(const float e1[3], float e2[3])
{
    const float d = dot3(e1, e2);
#pragma unroll
    for (int r = 0; r < 3; ++r) e2[r] = __fmaf_rn(-d, e1[r], e2[r]);
    const float n = __fsqrt_rn(dot3(e2, e2));
#pragma unroll
    for (int r = 0; r < 3; ++r) e2[r] = __fdiv_rn(e2[r], n);
}
// columns p,q of W (3x3, W[r][c]) and V; indices are compile-time after unrolling
template <int P, int Q>
BFR_DEVINL void jacobi_pair(float W[3][3], float V[3][3])
{
    const float alpha = __fmaf_rn(W[2][P], W[2][P], __fmaf_rn(W[1][P], W[1][P], __fmul_rn(W[0][P], W[0][P])));
    const float beta  = __fmaf_rn(W[2][Q], W[2][Q], __fmaf_rn(W[1][Q], W[1][Q], __fmul_rn(W[0][Q], W[0][Q])));
    const float gamma = __fmaf_rn(W[2][P], W[2][Q], __fmaf_rn(W[1][P], W[1][Q], __fmul_rn(W[0][P], W[0][Q])));
    if (gamma == 0.0f) return;
    const float zeta = __fdiv_rn(__fsub_rn(beta, alpha), __fmul_rn(2.0f, gamma));
    float t = __fdiv_rn(1.0f, __fadd_rn(fabsf(zeta), __fsqrt_rn(__fmaf_rn(zeta, zeta, 1.0f))));
    if (zeta < 0.0f) t = -t;
    const float c = __fdiv_rn(1.0f, __fsqrt_rn(__fmaf_rn(t, t, 1.0f)));
    const float s = __fmul_rn(c, t);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float x = W[r][P], y = W[r][Q];
        W[r][P] = __fmaf_rn(c, x, -__fmul_rn(s, y));
        W[r][Q] = __fmaf_rn(s, x, __fmul_rn(c, y));
        x = V[r][P]; y = V[r][Q];
        V[r][P] = __fmaf_rn(c, x, -__fmul_rn(s, y));
        V[r][Q] = __fmaf_rn(s, x, __fmul_rn(c, y));
    }
}
// H row-major 3x3 (H[r][c] = sum w a_r b_c). Returns false (R = I) when sigma_2 <= 1e-6 sigma_1.
BFR_DEVINL bool kabsch_rotation(const float H[9], float R[9])
{
    float W[3][3], V[3][3] = { { 1.f, 0.f, 0.f }, { 0.f, 1.f, 0.f }, { 0.f, 0.f, 1.f } };
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) W[r][c] = H[3 * r + c];
#pragma unroll 1
    for (int sweep = 0; sweep < 4; ++sweep) { jacobi_pair<0, 1>(W, V); jacobi_pair<0, 2>(W, V); jacobi_pair<1, 2>(W, V); }
    float n2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) n2[k] = __fmaf_rn(W[2][k], W[2][k], __fmaf_rn(W[1][k], W[1][k], __fmul_rn(W[0][k], W[0][k])));
    // largest and second-largest column norms, selected with compares only (run-time indexing of n2[] would put it in local memory)
    int i1 = 0; float m1 = n2[0];
    if (n2[1] > m1) { i1 = 1; m1 = n2[1]; }
    if (n2[2] > m1) { i1 = 2; m1 = n2[2]; }
    const int ia = (i1 == 0) ? 1 : 0, ib = (i1 == 2) ? 1 : 2;
    const float na = (i1 == 0) ? n2[1] : n2[0], nb = (i1 == 2) ? n2[1] : n2[2];
    const int i2 = (nb > na) ? ib : ia;
    const float m2 = (nb > na) ? nb : na;
    float w1[3], w2[3], v1[3], v2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        w1[r] = (i1 == 0) ? W[r][0] : (i1 == 1) ? W[r][1] : W[r][2];
        v1[r] = (i1 == 0) ? V[r][0] : (i1 == 1) ? V[r][1] : V[r][2];
        w2[r] = (i2 == 0) ? W[r][0] : (i2 == 1) ? W[r][1] : W[r][2];
        v2[r] = (i2 == 0) ? V[r][0] : (i2 == 1) ? V[r][1] : V[r][2];
    }
    const float s1 = __fsqrt_rn(m1), s2 = __fsqrt_rn(m2);
    if (!(s2 > __fmul_rn(1e-6f, s1)) || !(s1 > 0.0f)) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0f : 0.0f;
        return false;
    }
    float u1[3], u2[3], u3[3], v3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { u1[r] = __fdiv_rn(w1[r], s1); u2[r] = __fdiv_rn(w2[r], s2); }
    const float nv = __fsqrt_rn(dot3(v1, v1));
#pragma unroll
    for (int r = 0; r < 3; ++r) v1[r] = __fdiv_rn(v1[r], nv);
    gram_schmidt2(u1, u2);
    gram_schmidt2(v1, v2);
    cross3(u1, u2, u3);
    cross3(v1, v2, v3);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[3 * r + c] = __fmaf_rn(v3[r], u3[c], __fmaf_rn(v2[r], u2[c], __fmul_rn(v1[r], u1[c])));
    return true;
}

// squared residual of one correspondence under (R, t): utils/SE3.py:43-57 transform + squared norm
BFR_DEVINL float resid2(const float R[9], const float t[3], float sx, float sy, float sz, float qx, float qy, float qz)
{
    const float x = __fsub_rn(__fmaf_rn(R[0], sx, __fmaf_rn(R[1], sy, __fmaf_rn(R[2], sz, t[0]))), qx);
    const float y = __fsub_rn(__fmaf_rn(R[3], sx, __fmaf_rn(R[4], sy, __fmaf_rn(R[5], sz, t[1]))), qy);
    const float z = __fsub_rn(__fmaf_rn(R[6], sx, __fmaf_rn(R[7], sy, __fmaf_rn(R[8], sz, t[2]))), qz);
    return __fmaf_rn(x, x, __fmaf_rn(y, y, __fmul_rn(z, z)));
}

// One RANSAC hypothesis (Open3D 0.13 iteration as called at models/BUFFER.py:318-324; see oracle/bfr_oracle.c
// orc_hypothesis for the line-by-line statement), split in two stages so that the kernel can compact between them.
// corr: K records of 8 floats {sx sy sz 0 qx qy qz 0}.
BFR_DEVINL void load_sample(const float4* __restrict__ corr, const uint32_t id[3], float s[3][3], float q[3][3])
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float4 a = __ldg(&corr[2 * (size_t)id[i]]), b = __ldg(&corr[2 * (size_t)id[i] + 1]);
        s[i][0] = a.x; s[i][1] = a.y; s[i][2] = a.z; q[i][0] = b.x; q[i][1] = b.y; q[i][2] = b.z;
    }
}
// CorrespondenceCheckerBasedOnEdgeLength on squared lengths of the sample pairs (0,1),(0,2),(1,2)
BFR_DEVINL bool edge_lengths_ok(const float s[3][3], const float q[3][3], float sim_th2)
{
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int a = (e == 2) ? 1 : 0, b = (e == 0) ? 1 : 2;
        float dx = __fsub_rn(s[a][0], s[b][0]), dy = __fsub_rn(s[a][1], s[b][1]), dz = __fsub_rn(s[a][2], s[b][2]);
        const float ds2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
        dx = __fsub_rn(q[a][0], q[b][0]); dy = __fsub_rn(q[a][1], q[b][1]); dz = __fsub_rn(q[a][2], q[b][2]);
        const float dt2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
        if (ds2 < __fmul_rn(dt2, sim_th2) || dt2 < __fmul_rn(ds2, sim_th2)) return false;
    }
    return true;
}
// stage 1 (cheap, every hypothesis): Philox sample, repeated-index rejection, edge-length checker
BFR_DEVINL bool hypothesis_precheck(const float4* __restrict__ corr, uint32_t K, uint64_t seed, uint32_t pair_id, uint32_t h, float sim_th2,
                                    uint32_t id[3], float s[3][3], float q[3][3])
{
    sample3(seed, pair_id, h, K, id);
    if (id[0] == id[1] || id[0] == id[2] || id[1] == id[2]) return false;
    load_sample(corr, id, s, q);
    return edge_lengths_ok(s, q, sim_th2);
}
// stage 2 (survivors of stage 1): 3-point Kabsch + CorrespondenceCheckerBasedOnDistance on the three samples
BFR_DEVINL bool hypothesis_fit(const float s[3][3], const float q[3][3], float dist_th2, float R[9], float t[3])
{
    const float third = 0.33333334f;
    float cs[3], cq[3], Hm[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        cs[r] = __fmul_rn(__fadd_rn(__fadd_rn(s[0][r], s[1][r]), s[2][r]), third);
        cq[r] = __fmul_rn(__fadd_rn(__fadd_rn(q[0][r], q[1][r]), q[2][r]), third);
    }
    float a[3][3], b[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int r = 0; r < 3; ++r) { a[i][r] = __fsub_rn(s[i][r], cs[r]); b[i][r] = __fsub_rn(q[i][r], cq[r]); }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Hm[3 * r + c] = __fmaf_rn(a[2][r], b[2][c], __fmaf_rn(a[1][r], b[1][c], __fmul_rn(a[0][r], b[0][c])));
    if (!kabsch_rotation(Hm, R)) return false;
#pragma unroll
    for (int r = 0; r < 3; ++r)
        t[r] = __fsub_rn(cq[r], __fmaf_rn(R[3 * r + 2], cs[2], __fmaf_rn(R[3 * r + 1], cs[1], __fmul_rn(R[3 * r + 0], cs[0]))));
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (resid2(R, t, s[i][0], s[i][1], s[i][2], q[i][0], q[i][1], q[i][2]) > dist_th2) return false;
    return true;
}
BFR_DEVINL bool make_hypothesis(const float4* __restrict__ corr, uint32_t K, uint64_t seed, uint32_t pair_id, uint32_t h,
                                float dist_th2, float sim_th2, float R[9], float t[3])
{
    uint32_t id[3]; float s[3][3], q[3][3];
    if (!hypothesis_precheck(corr, K, seed, pair_id, h, sim_th2, id, s, q)) return false;
    return hypothesis_fit(s, q, dist_th2, R, t);
}

// ---- helpers with a bit-pinned CPU twin in oracle/bfr_oracle.c -----------------------------------------------------------
// sqrt_rn(d2) < thr  <=>  d2 < sqrt_threshold(thr): the smallest float whose correctly rounded square root reaches thr (sqrt_rn is
// monotone).  Lets the scoring loops keep a squared-distance compare while counting exactly what the reference's
// `torch.sqrt(d2) < thr` counts (models/BUFFER.py:305-308).  thr <= 0 or NaN: nothing is ever an inlier.
BFR_DEVINL float sqrt_threshold(float thr)
{
    if (!(thr > 0.0f)) return 0.0f;
    if (thr == __int_as_float(0x7f800000)) return thr;
    float x = __fmul_rn(thr, thr);
    if (x == __int_as_float(0x7f800000)) x = 3.402823466e38f;
    if (__fsqrt_rn(3.402823466e38f) < thr) return __int_as_float(0x7f800000);     // every finite d2 passes
    while (x > 0.0f && __fsqrt_rn(__uint_as_float(__float_as_uint(x) - 1u)) >= thr) x = __uint_as_float(__float_as_uint(x) - 1u);
    while (__fsqrt_rn(x) < thr) x = __uint_as_float(__float_as_uint(x) + 1u);
    return x;
}

// sin / cos of a non-negative angle with nothing but fma / mul / rint (Cody-Waite reduction by pi/2 in three parts, Cephes minimax
// polynomials on [-pi/4, pi/4]); max error ~1.2e-7 for angles up to a few hundred.  Used for the LRF angle of models/BUFFER.py:295.
BFR_DEVINL void det_sincos(float a, float& sn, float& cs)
{
    const float k = rintf(__fmul_rn(a, 0.63661977236758134f));
    float r = __fmaf_rn(-k, 1.5703125f, a);
    r = __fmaf_rn(-k, 4.837512969970703125e-4f, r);
    r = __fmaf_rn(-k, 7.54978995489188e-8f, r);
    const float z = __fmul_rn(r, r);
    float ps = __fmaf_rn(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = __fmaf_rn(z, ps, -1.6666654611e-1f);
    const float s = __fmaf_rn(__fmul_rn(z, r), ps, r);
    float pc = __fmaf_rn(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = __fmaf_rn(z, pc, 4.166664568298827e-2f);
    const float c = __fmaf_rn(__fmul_rn(z, z), pc, __fmaf_rn(-0.5f, z, 1.0f));
    const int q = ((int)k) & 3;
    sn = (q == 0) ? s : (q == 1) ? c : (q == 2) ? -s : -c;
    cs = (q == 0) ? c : (q == 1) ? -s : (q == 2) ? -c : s;
}

// natural logarithm of a positive finite double from + - * / fma only (exact exponent split, atanh series): every step is a
// correctly rounded IEEE operation, so the CPU oracle reproduces it bit for bit.  Used for Open3D's RANSAC iteration bound.
BFR_DEVINL double det_log(double v)
{
    long long bits = __double_as_longlong(v);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);     // [1, 2)
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    const double y = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0)), y2 = __dmul_rn(y, y);
    double p = 1.0 / 27.0;
#pragma unroll
    for (int k = 25; k >= 1; k -= 2) p = __fma_rn(p, y2, 1.0 / (double)k);
    return __fma_rn((double)e, 0.6931471805599453, __dmul_rn(__dmul_rn(2.0, y), p));
}
// Open3D RANSACConvergenceCriteria: iterations allowed once a hypothesis with `count` inliers of K is the best:
// min(max_iter, ceil(log(1 - confidence) / log(1 - (count / K)^3)))
BFR_DEVINL uint32_t ransac_exit_bound(uint32_t count, uint32_t K, double log_1m_conf, uint32_t max_iter)
{
    if (count == 0u || K == 0u) return max_iter;
    const double x = __ddiv_rn((double)count, (double)K), x3 = __dmul_rn(__dmul_rn(x, x), x);
    if (!(x3 < 1.0)) return 0u;
    const double den = det_log(__dsub_rn(1.0, x3));
    if (!(den < 0.0)) return max_iter;
    const double b = ceil(__ddiv_rn(log_1m_conf, den));
    if (!(b < (double)max_iter)) return max_iter;
    return b < 0.0 ? 0u : (uint32_t)b;
}

}  // namespace bfr
