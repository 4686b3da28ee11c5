// Blackwell (sm_100a) primitives used by the tensor-core sweep: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (TMEM allocation, UMMA issue / commit, TMEM loads) and the shared-memory / instruction descriptor encodings.
// Raw PTX only; nothing here is portable to other architectures by design.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace slsgp
{
    namespace tc
    {
        __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

        // ---- mbarrier ------------------------------------------------------------------------------------------
        __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
        }
        __device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __device__ __forceinline__ void mbar_arrive(uint32_t bar)
        {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
        __device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        }
        __device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
        {
            uint32_t ok;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}"
                : "=r"(ok)
                : "r"(bar), "r"(parity)
                : "memory");
            return ok != 0;
        }
        // Spin with a watchdog: a protocol bug must not hang the GPU box. On expiry the error word is set to `code`
        // and the kernel traps (the host sees a launch failure and reports the code).
        __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code)
        {
            uint32_t spins = 0;
            while (!mbar_try_wait(bar, parity))
            {
                if (++spins > (1u << 24))
                {
                    if (err) atomicCAS(err, 0, code);
                    __threadfence_system();
                    __trap();
                }
            }
        }

        // ---- TMA --------------------------------------------------------------------------------------------------
        __device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
        {
            asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
        }
        // 2-D tiled load, coordinates (c0 = innermost element index, c1 = row), completes `bytes` on the mbarrier.
        __device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
        {
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                : "memory");
        }

        // ---- tcgen05: TMEM management ----------------------------------------------------------------------------
        __device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) // whole warp, .sync.aligned
        {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) // whole warp
        {
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
        }
        __device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
        __device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

        // ---- tcgen05: UMMA ---------------------------------------------------------------------------------------------
        // Shared-memory matrix descriptor for a K-major operand tile whose rows are 128 bytes (64 fp16) long and are
        // stored with the 128-byte swizzle TMA produces: 8-row groups are 1024 bytes apart (SBO), LBO is unused for
        // swizzled K-major layouts, version = 1 (sm_100), layout type 2 = SWIZZLE_128B.
        __device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr)
        {
            uint64_t d = 0;
            d |= (uint64_t) ((smem_addr & 0x3FFFF) >> 4);  // bits [0,14)  start address >> 4
            d |= (uint64_t) 1 << 16;                       // bits [16,30) leading byte offset (ignored)
            d |= (uint64_t) (1024 >> 4) << 32;             // bits [32,46) stride byte offset
            d |= (uint64_t) 1 << 46;                       // bits [46,48) descriptor version
            d |= (uint64_t) 2 << 61;                       // bits [61,64) SWIZZLE_128B
            return d;
        }
        // Instruction descriptor, kind::f16: fp16 A and B (both K-major), fp32 accumulate, shape M x N.
        __host__ __device__ inline uint32_t instr_desc_f16(int M, int N)
        {
            uint32_t d = 0;
            d |= 1u << 4;                   // c_format = F32
            d |= 0u << 7;                   // a_format = F16
            d |= 0u << 10;                  // b_format = F16
            d |= 0u << 15;                  // a_major  = K
            d |= 0u << 16;                  // b_major  = K
            d |= (uint32_t) (N >> 3) << 17; // n_dim
            d |= (uint32_t) (M >> 4) << 24; // m_dim
            return d;
        }
        // D[tmem] (+)= A[smem] * B[smem]^T; issued by ONE thread.
        __device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate)
        {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}"
                ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                : "memory");
        }
        // Arrive on an mbarrier once every previously issued UMMA of this thread has completed (implies
        // tcgen05.fence::before_thread_sync).
        __device__ __forceinline__ void umma_commit(uint32_t bar)
        {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }

        // ---- CTA pair (cta_group::2) variants: the leader CTA (cluster rank 0) issues the MMAs for both SMs ----------
        __device__ __forceinline__ uint32_t cluster_ctarank()
        {
            uint32_t r;
            asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
            return r;
        }
        // shared::cluster address of `smem_addr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
        __device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank)
        {
            uint32_t r;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
            return r;
        }
        __device__ __forceinline__ void cluster_sync_all() // every thread of every CTA of the cluster
        {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        __device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) // bar given as a shared::cluster address
        {
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
        }
        // TMA load whose completion bytes are posted on a barrier that may live in the peer CTA of the pair
        __device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1)
        {
            asm volatile(
                "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
                : "memory");
        }
        __device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) // same warp id in both CTAs
        {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
        __device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols)
        {
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
        }
        // D (256 x N, rows split over the two CTAs' TMEM) (+)= A (128 rows per CTA) * B^T (N/2 rows per CTA)
        __device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                      uint32_t accumulate)
        {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}"
                ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                : "memory");
        }
        // arrive on the barrier at the same shared-memory offset in every CTA of `mask` once the issued UMMAs are done
        __device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask)
        {
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(bar), "h"(mask)
                         : "memory");
        }

        template <int NCTA>
        __device__ __forceinline__ void umma_f16_n(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
        {
            if (NCTA == 2)
                umma_f16_pair(tmem_d, desc_a, desc_b, idesc, accumulate);
            else
                umma_f16(tmem_d, desc_a, desc_b, idesc, accumulate);
        }
        template <int NCTA> __device__ __forceinline__ void umma_commit_n(uint32_t bar)
        {
            if (NCTA == 2)
                umma_commit_pair(bar, 3);
            else
                umma_commit(bar);
        }
        // One lane of a converged warp (elect.sync): true in exactly one lane.
        __device__ __forceinline__ bool elect_one()
        {
            uint32_t pred = 0;
            asm volatile(
                "{\n\t"
                ".reg .b32 rx;\n\t"
                ".reg .pred px;\n\t"
                "elect.sync rx|px, 0xFFFFFFFF;\n\t"
                "selp.u32 %0, 1, 0, px;\n\t"
                "}"
                : "=r"(pred));
            return pred != 0;
        }

        // ---- tcgen05: TMEM -> registers ----------------------------------------------------------------------------------
        // Each thread of the warp receives N consecutive 32-bit columns of its own lane (warp w may only touch lanes
        // 32 * (w % 4) .. + 31). Asynchronous: tmem_ld_wait() before the registers are read.
        __device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32])
        {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
        }
        __device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16])
        {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
        }
        __device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

        __device__ __forceinline__ void named_bar_sync(int id, int nthreads)
        {
            asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
        }

        // Packed FP32 FMA (Blackwell FFMA2): d = a * b + c on both halves of a register pair, one issue slot.
        __device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
        {
            unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                               rc = *reinterpret_cast<unsigned long long*>(&c), rd;
            asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
            return *reinterpret_cast<float2*>(&rd);
        }

        __device__ __forceinline__ float2 fadd2(float2 a, float2 b) // packed FP32 add (FADD2)
        {
            unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
            return *reinterpret_cast<float2*>(&rd);
        }

        __device__ __forceinline__ float ex2_approx(float x)
        {
            float y;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
            return y;
        }
    } // namespace tc
} // namespace slsgp
