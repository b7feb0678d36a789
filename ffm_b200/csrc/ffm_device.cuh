// Device-side building blocks shared by the FFM kernels (sm_100a).
//
//  * Philox4x32-10 counter-based draws (Salmon et al., SC'11) keyed (entity, step, episode, stream)
//    -- the replacement for the reference's process-global generators (model/ffm_core.py:84,95,96).
//  * IEEE helpers that pin NumPy's evaluation order: separate multiply and add (no FMA contraction).
//  * TMA bulk-copy / mbarrier wrappers (inline PTX).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ffm {

enum : uint32_t { STREAM_MOVE = 0, STREAM_CONFLICT = 1, STREAM_EPS = 2, STREAM_PLACE = 3 };

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;   // one IMAD.WIDE each
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// 53-bit double in [0,1) from two words, the construction of NumPy's random_sample().
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    const unsigned long long v = ((unsigned long long)(a >> 5) << 26) | (unsigned long long)(b >> 6);
    return (double)v * (1.0 / 9007199254740992.0);
}

struct Draw2 { double u0, u1; };

__device__ __forceinline__ Draw2 draw2(unsigned long long seed, uint32_t episode, uint32_t step,
                                       uint32_t stream, uint32_t entity) {
    const uint4 o = philox4x32_10(entity, step, episode, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
    Draw2 d;
    d.u0 = u53(o.x, o.y);
    d.u1 = u53(o.z, o.w);
    return d;
}

__device__ __forceinline__ double draw_u0(unsigned long long seed, uint32_t episode, uint32_t step,
                                          uint32_t stream, uint32_t entity) {
    const uint4 o = philox4x32_10(entity, step, episode, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
    return u53(o.x, o.y);
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion ---------------------------
// Used to stage an episode's fields into shared memory: one elected thread arms the barrier with the
// byte count and issues the copies; everybody waits on the barrier's phase.  Sizes and both addresses
// must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// shared -> global bulk store of data written by ordinary stores (needs the proxy fence + a barrier first)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- arithmetic with NumPy's rounding sequence -------------------------------------------------
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  div_rn(float a, float b)   { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float  exp_t(float x)  { return expf(x); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float  max_t(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
template <typename S> __device__ __forceinline__ S neg_inf();
template <> __device__ __forceinline__ float  neg_inf<float>()  { return -__int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double neg_inf<double>() { return -__longlong_as_double(0x7ff0000000000000LL); }

}  // namespace ffm
