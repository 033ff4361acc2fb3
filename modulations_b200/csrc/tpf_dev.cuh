// tpf_dev.cuh — device-side primitives shared by the thread-per-frame decoders (decode_tpf.cu: the parity mode,
// decode_nii.cu: the non-parity "nii" mode): tensor memory used as a lane-private scratchpad, cp.async (LDGSTS)
// rings, bulk copies completed on an mbarrier, L2 discards and cache-level stores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200dvb {
namespace {

// ---- tensor-memory access (lane-private scratchpad; no MMA anywhere in this kernel) ----
__device__ __forceinline__ void tm_ld8(unsigned taddr, float (&g)[8])
{   // thread t <- TMEM lane (quadrant + t), columns taddr.col .. +7
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(g[0]), "=f"(g[1]), "=f"(g[2]), "=f"(g[3]), "=f"(g[4]), "=f"(g[5]), "=f"(g[6]), "=f"(g[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_ld8_half(unsigned taddr, float (&g)[8])
{   // threads t and t + 16 <- TMEM lane (taddr.lane + t), same columns (t < 16)
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], 0;"
                 : "=f"(g[0]), "=f"(g[1]), "=f"(g[2]), "=f"(g[3]), "=f"(g[4]), "=f"(g[5]), "=f"(g[6]), "=f"(g[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld(float (&g)[8])
{   // the loaded registers become valid here: make every later use depend on this statement
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(g[0]), "+f"(g[1]), "+f"(g[2]), "+f"(g[3]), "+f"(g[4]), "+f"(g[5]), "+f"(g[6]), "+f"(g[7])
                 :: "memory");
}
__device__ __forceinline__ void tm_st8(unsigned taddr, const float (&g)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "f"(g[0]), "f"(g[1]), "f"(g[2]), "f"(g[3]), "f"(g[4]), "f"(g[5]), "f"(g[6]), "f"(g[7])
                 : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- cp.async (LDGSTS) into this warp's staging area: latency of the L2-resident workspace
//      is hidden by depth, not by registers or by other warps (there are none) --------------
__device__ __forceinline__ unsigned s_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// `one` is the value 1 read back from shared memory (Ctx::one).  ptxas 12.9 folds any warp-uniform
// part of the destination into a [R+UR+imm] operand, and LDGSTS with that operand form raises
// "illegal instruction" on sm_100a (found on the B200; modulations_b200/build.py checks the
// SASS).  A product with a value ptxas cannot see through keeps the address in one register.
__device__ __forceinline__ void cpa16(void *dst, const void *src, unsigned one)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_addr(dst) * one), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16_stream(void *dst, const void *src, unsigned long long pol, unsigned one)
{   // read-once data (channel LLRs): do not let it push the extrinsics out of L2
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;"
                 ::"r"(s_addr(dst) * one), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- bulk asynchronous copies (the TMA unit's non-tensor path: cp.async.bulk, SASS UBLKCP) completed on an
//      mbarrier.  Used where a transfer is one long contiguous run (whole LLR rows of the transposition): ONE
//      instruction of ONE lane moves a 5 KB row, against 10 LDGSTS warp-instructions with 12 shared-memory
//      wavefronts each --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{   // one arrival + the number of bytes the bulk copies of this phase will deliver
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" :: "r"(mbar), "r"(parity) : "memory");
}
// generic-proxy accesses to a shared-memory buffer (earlier reads / writes by ld/st) ordered before the async-proxy
// writes of the bulk copies that reuse it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s_stream(void *dst, const void *src, unsigned bytes, unsigned mbar, unsigned long long pol)
{   // bytes: a multiple of 16; dst / src 16-byte aligned; read-once data: L2 evict-first
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(s_addr(dst)), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
}

// The workspace is scratch: once a line of Y / checkpoints / old extrinsics has been consumed it
// is dead until it is rewritten.  discard.global.L2 drops it WITHOUT a write-back, so the 160 MB of
// scratch of the resident warps stops streaming through HBM (it was 150 KB of DRAM writes per frame).
__device__ __forceinline__ void l2_discard(const void *line128)
{
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(line128) : "memory");
}

template <typename T> __device__ __forceinline__ void st_ws(T *p, const T &v) { __stcg(p, v); }   // workspace store: L2 only
// ---- Y = Lc + La of the current window, parked in this lane's spare TMEM columns ------------
__device__ __forceinline__ void tm_ld4(unsigned taddr, float (&y)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(y[0]), "=f"(y[1]), "=f"(y[2]), "=f"(y[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld4(float (&y)[4])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(y[0]), "+f"(y[1]), "+f"(y[2]), "+f"(y[3]) :: "memory");
}
// cp.async with immediate offsets on both addresses: the four copies of a step pair share ONE destination
// register (and the two channel copies one source register): a handful of address instructions per pair instead
// of a multiply + LEA pair per copy.
template <int DOFF, int SOFF>
__device__ __forceinline__ void cpa16_off(unsigned d, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%3], 16;" ::"r"(d), "l"(src), "n"(DOFF), "n"(SOFF) : "memory");
}
template <int DOFF, int SOFF>
__device__ __forceinline__ void cpa16_stream_off(unsigned d, const void *src, unsigned long long pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0+%3], [%1+%4], 16, %2;"
                 ::"r"(d), "l"(src), "l"(pol), "n"(DOFF), "n"(SOFF) : "memory");
}

}  // namespace
}  // namespace b200dvb
