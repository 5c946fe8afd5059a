// PTX wrappers shared by the inference tensor-core kernels (conv1d_tc.cu,
// conv_pair_tc.cu): mbarriers, 1-D bulk async copies, tcgen05 MMA / TMEM loads and
// the un-swizzled K-major shared-memory descriptors.  sm_100a only.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

namespace pmn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t address = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(address), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// e4m3 x e4m3 -> fp32 (K = 32 per instruction, twice the bf16 rate)
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tc_load16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
        "%12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_load32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
        "%12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, "
        "%30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor):
// [0,14) start >> 4 | [16,30) leading byte offset >> 4 (stride between the two
// 8-element K chunks) | [32,46) stride byte offset >> 4 (stride between 8-row
// groups) | [46,48) version = 1 | [61,64) layout = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t smem_desc(uint32_t address, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((address & 0x3FFFFu) >> 4) |
           ((uint64_t)(lbo >> 4) << 16) |
           ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major
__host__ __device__ constexpr uint32_t instr_desc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// The same with A / B format 0: fp16 operands under kind::f16, e4m3 operands under kind::f8f6f4
__host__ __device__ constexpr uint32_t instr_desc_format0(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}


// Make generic-proxy writes to shared memory (st.shared) visible to the async proxy
// (tcgen05.mma operand reads, bulk copies) before a barrier hands the buffer over
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// bf16 hi / lo halves of y (y ~ hi + lo to 16 mantissa bits), two values per word.  The packed
// convert (cvt.rn.bf16x2.f32 -> F2FP.BF16.F32.PACK_AB) rounds like two scalar converts but is one
// instruction on the ALU pipe instead of two quarter-rate F2F plus a PRMT
__device__ __forceinline__ void split_pair(float y0, float y1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(
        y0 - __uint_as_float(hi << 16), y1 - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// "fp16 + 2 x fp8" operand of two neighbouring channels (conv1d_tc.cuh): main = fp16x2 of
// x kF8ScaleMain (saturating), coarse = e4m3x2 of x kF8ScaleX (in the low 16 bits), low = e4m3x2 of
// (x - main / kF8ScaleMain) kF8ScaleXLow.  The coarse copy is taken from the fp16 value, not from x:
// the difference is 2^-11 of a term that is itself 2^-11 of the product.
constexpr float kF8Main = 128.f;        // = kF8ScaleMain
constexpr float kF8Coarse = 8.f;        // = kF8ScaleX
constexpr float kF8Low = 16384.f;       // = kF8ScaleXLow
__device__ __forceinline__ void split_pair_f8(float y0, float y1, uint32_t& main, uint32_t& coarse, uint32_t& low) {
    const float s0 = fminf(fmaxf(y0 * kF8Main, -65504.f), 65504.f);
    const float s1 = fminf(fmaxf(y1 * kF8Main, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(s0, s1);
    main = *reinterpret_cast<const uint32_t*>(&h);
    const float2 back = __half22float2(h);
    const __half2 scaled = __hmul2(h, __float2half2_rn(kF8Coarse / kF8Main));
    coarse = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&scaled), __NV_SATFINITE, __NV_E4M3);
    low = __nv_cvt_float2_to_fp8x2(
        make_float2((s0 - back.x) * (kF8Low / kF8Main), (s1 - back.y) * (kF8Low / kF8Main)),
        __NV_SATFINITE, __NV_E4M3);
}

}  // namespace tc
}  // namespace pmn
