// Toeplitz complex GEMM on the 5th-generation tensor cores (tcgen05, TMEM accumulators), fp32-accurate through a
// 3xTF32 split AND fp32 promotion of partial sums.  Same contract as the CUDA-core kernel in thz_czt.cu
// (include/thzdoe.h: thz_toeplitz_gemm).
//
// Real-valued formulation.  For complex C = T . B' (B' = pro * B, folded into B by thz_k_cmul) two tile accumulators
// D1 = Tr . [Br | Bi]  and  D2 = Ti . [Br | Bi]  (128 TMEM columns each; column n real part, column 64 + n imaginary
// part of complex column n) are built by tcgen05.mma (kind::tf32, M = 128, N = 128, K = 8) and combined when they are
// drained:  Cr[n] = D1[n] - D2[64 + n],  Ci[n] = D1[64 + n] + D2[n].  B is therefore stored ONCE ([Br | Bi], K-major),
// not as the 2 x 2 real block matrix -- the producers' shared-memory stores compete with the tensor core's operand
// reads for the 128 B/clk of the SM, which is what bounds this kernel (profiles/README.md, "tcgen05 CZT timeline").
// Every operand is split x = hi + lo (hi = RN-to-TF32, lo = x - hi exactly) and the product accumulated as
// hi*hi + hi*lo + lo*hi (lo*lo < 2^-22 relative): 12 MMAs per 16 complex k.
//
// Toeplitz operand = sliding strip.  T(m0 + r, k0 + kk) = g[off + sm (m0 + rho) + sk kk] depends on r and k0 only
// through rho = r - k0 (sm = -sk), so the 128 x 16 A tile of k-block kb is rows [-16 kb, -16 kb + 128) of ONE strip
// whose row rho holds [Re g(rho, 0..15) | Im g(rho, 0..15)] (128 bytes), and each k-block adds just 16 new rows.  The
// strip lives in a circular buffer of TC_AW rows whose first 112 rows are mirrored behind its end, so that every
// 128-row window is contiguous for the UMMA descriptor.
//
// Promotion.  The tensor core adds into its fp32 accumulator with truncation, so a single long accumulation
// drifts linearly with K (measured: 7.8e-6 / 2.7e-5 / 5.3e-5 relative at K = 256 / 1024 / 2048 with one TMEM
// accumulator -- above the 1e-5 parity bound).  Therefore the K loop is cut into chunks of 128 complex k (96 MMAs;
// 1.9e-6 relative at every K, 1.0e-6 with 64-k chunks at 5 % more time)
// that each START FROM ZERO in one of two TMEM buffer pairs; 8 accumulate warps drain a finished pair with
// tcgen05.ld and add it to fp32 register accumulators with IEEE rounding while the MMA warp already works on the
// other pair (2 x 2 x 128 = all 512 TMEM columns).
//
// Nothing is loaded by TMA: both operands are *generated* -- the strip from the chirp filter g, the B tile from the
// prologue-scaled input -- by 8 producer warps that write the canonical K-major SWIZZLE_128B layout directly
// (16-byte chunk c of row r goes to chunk c ^ (r & 7) of its 128-byte row), then fence to the async proxy.
//
// Roles (17 warps): warps 0-7 accumulate/epilogue (warp w: TMEM lanes 32 (w%4).., complex columns 32 (w/4)..),
// warps 8-15 producers, warp 16 MMA issuer.  One 128 x 64 complex output tile per CTA; a stage is 32 complex k:
//   producers:  wait empty[s] -> 32 new strip rows + B_hi/B_lo[s] -> fence.proxy.async -> arrive full[s]
//   MMA warp :  per chunk: wait tmem_empty[b]; per stage: wait full[s] -> 24 x tcgen05.mma -> commit empty[s];
//               commit tmem_full[b]
//   accumulate: wait tmem_full[b] -> tcgen05.ld -> acc += -> arrive tmem_empty[b];  finally * epi -> global
#include "thz_common.cuh"
#include "thz_czt_args.h"
#include "thz_runtime.h"

#define TC_BM 128                     // complex rows per tile = TMEM lanes
#define TC_BN 64                      // complex columns per tile -> 128 real accumulator columns per D1 / D2
#define TC_KC 16                      // complex k per k-block = one strip row (16 re | 16 im = 128 bytes)
#define TC_KS 32                      // complex k per stage = one 128-byte row of the B tile = two k-blocks
#define TC_STAGES 3
#define TC_CHUNK 4                    // stages per promotion chunk (128 complex k)
#define TC_B_BYTES (2 * TC_BN * 128)  // 16 KB: 128 rows ([Br | Bi] of 64 columns) x 128 B (32 k)
#define TC_STAGE_BYTES (2 * TC_B_BYTES)                    // hi + lo of the B operand: 32 KB per stage
// Strip window: a row written for k-block kb replaces the row last read by k-block kb - (TC_AW - 128) / 16 - 1; the
// producers write the rows of k-blocks 2S and 2S + 1 after the MMAs of stage S - TC_STAGES (k-blocks <= 2S - 5) have
// completed, hence TC_AW >= 128 + 16 * 5.
#define TC_AW 208
#define TC_AROWS (TC_AW + 112)                             // + mirror of rows 0..111 (window start <= TC_AW - 16)
#define TC_AS_BYTES (TC_AROWS * 128)                       // one strip (hi or lo): 40 KB
#define TC_SMEM_BYTES (2 * TC_AS_BYTES + TC_STAGES * TC_STAGE_BYTES)   // 80 KB + 96 KB
#define TC_ACC_WARPS 8
#define TC_PROD_WARPS 8
#define TC_THREADS ((TC_ACC_WARPS + TC_PROD_WARPS + 1) * 32)
#define TC_TMEM_COLS 512              // two buffers of (D1, D2) = 2 x 128 columns each

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU -- trap instead (the launch then fails loudly).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc], kind::tf32, K = 8
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                   "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

// Two 8-column TMEM loads in flight, then ONE wait; the wait names the destination registers as read-write operands
// so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, uint32_t tb, float (&ra)[8], float (&rb)[8]) {
    uint32_t u[8], v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(ta));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tb));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]), "+r"(v[0]),
                   "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        ra[i] = __uint_as_float(u[i]);
        rb[i] = __uint_as_float(v[i]);
    }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B -> 64), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b_format TF32 (2) @7/@10,
// K-major A and B (0) @15/@16, N >> 3 @17, M >> 4 @24.
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a K-major SWIZZLE_128B tile (tile base 1024-aligned)
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// hi/lo split of four values and their 16-byte stores into the hi and lo tiles (32-bit shared addresses)
__device__ __forceinline__ void st_split4(uint32_t hi_tile, uint32_t lo_tile, uint32_t off, const float (&v)[4]) {
    float4 h, l;
    h.x = tf32_hi(v[0]);
    h.y = tf32_hi(v[1]);
    h.z = tf32_hi(v[2]);
    h.w = tf32_hi(v[3]);
    l.x = v[0] - h.x;
    l.y = v[1] - h.y;
    l.z = v[2] - h.z;
    l.w = v[3] - h.w;
    sts128(hi_tile + off, h);
    sts128(lo_tile + off, l);
}
__device__ __forceinline__ void split4(float4 v, float4& h, float4& l) {
    h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}
__device__ __forceinline__ float4 neg4(float4 v) { return make_float4(-v.x, -v.y, -v.z, -v.w); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, cpx v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ int wrap_mod(long long v, int L) {
    v %= L;
    return (int)(v < 0 ? v + L : v);
}

// ------------------------------------------------------------------------------- kernel
// Debug timeline (THZ_CZT_DEBUG=3): clock64 stamps of CTA (0,0,0): [role][stage or chunk][point]
#define TC_TL_N 256
__device__ long long g_tc_timeline[3][TC_TL_N][4];
#define TC_STAMP(role, idx, pt)                                                                          \
    do {                                                                                                 \
        if (tl && (idx) < TC_TL_N) g_tc_timeline[role][idx][pt] = clock64();                             \
    } while (0)
__global__ void __launch_bounds__(TC_THREADS, 1) thz_k_toeplitz_gemm_tc(const __grid_constant__ ToeplitzGemmArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[2 * TC_STAGES + 4];
    __shared__ uint32_t tmem_base_holder;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM, b = blockIdx.z;
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_STAGES]);
    const uint32_t tfull0 = smem_u32(&bars[2 * TC_STAGES]), tempty0 = smem_u32(&bars[2 * TC_STAGES + 2]);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full0 + 8 * s, TC_PROD_WARPS);   // one arrival per producer warp
            mbar_init(empty0 + 8 * s, 1);              // one tcgen05.commit
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(tfull0 + 8 * t, 1);              // one tcgen05.commit per chunk
            mbar_init(tempty0 + 8 * t, TC_ACC_WARPS);  // one arrival per accumulate warp
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_holder), TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_holder;
    const bool tl = (a.debug_mode == 3 || (a.debug_mode == 4 && a.epi == nullptr)) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (warp == 0 || warp == TC_ACC_WARPS || warp == TC_ACC_WARPS + TC_PROD_WARPS);
    const int NS = (a.K + TC_KS - 1) / TC_KS;               // stages
    const int NC = (NS + TC_CHUNK - 1) / TC_CHUNK;           // promotion chunks

    if (warp >= TC_ACC_WARPS && warp < TC_ACC_WARPS + TC_PROD_WARPS) {
        // ===================================================================== producers
        const int ptid = tid - TC_ACC_WARPS * 32;                  // 0..255
        const cpx* g = a.g + (size_t)b * a.L;
        const cpx* Bb = a.B + (size_t)b * a.sb_b;
        const uint32_t tiles_s = smem_u32(tiles);
        const uint32_t Ahi = tiles_s, Alo = tiles_s + TC_AS_BYTES, Bst0 = tiles_s + 2 * TC_AS_BYTES;
        // strip entry (rho, kk) = g[(off + sm (m0 + rho) + sk kk) mod L]; one item = 4 consecutive kk of one strip row,
        // written as a real chunk and an imaginary chunk, hi and lo (and mirrored if it is one of the first 112 rows)
        auto strip_load = [&](int idx, cpx (&gv)[4]) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                gv[e] = __ldg(g + idx);
                idx += a.sk;
                if (idx >= a.L) idx -= a.L;
                if (idx < 0) idx += a.L;
            }
        };
        auto strip_store = [&](int rho, int c, const cpx (&gv)[4]) {
            int row = rho % TC_AW;
            if (row < 0) row += TC_AW;
            float4 reh, rel, imh, iml;
            split4(make_float4(gv[0].x, gv[1].x, gv[2].x, gv[3].x), reh, rel);
            split4(make_float4(gv[0].y, gv[1].y, gv[2].y, gv[3].y), imh, iml);
            if (a.conj_g) {
                imh = neg4(imh);
                iml = neg4(iml);
            }
            const uint32_t o0 = sw128(row, c), o1 = sw128(row, 4 + c);
            sts128(Ahi + o0, reh);
            sts128(Alo + o0, rel);
            sts128(Ahi + o1, imh);
            sts128(Alo + o1, iml);
            if (row < TC_AROWS - TC_AW) {
                const uint32_t p0 = sw128(row + TC_AW, c), p1 = sw128(row + TC_AW, 4 + c);
                sts128(Ahi + p0, reh);
                sts128(Alo + p0, rel);
                sts128(Ahi + p1, imh);
                sts128(Alo + p1, iml);
            }
        };
        // initial strip: rows rho = -16 .. 127 (k-blocks 0 and 1 of stage 0): 144 rows x 4 chunks
        for (int item = ptid; item < 144 * 4; item += 256) {
            const int rho = (item >> 2) - 16, c = item & 3;
            cpx gv[4];
            strip_load(wrap_mod((long long)a.off + (long long)a.sm * (m0 + rho) + (long long)a.sk * (4 * c), a.L), gv);
            strip_store(rho, c, gv);
        }
        // stage S >= 1 adds the 32 rows rho = -32 S - 16 .. -32 S + 15: 128 items, threads 0..127; the g index of a
        // thread's item moves by -32 sm per stage
        const bool a_worker = ptid < 128;
        const int a_r = ptid & 31, a_c = (ptid >> 5) & 3;
        int a_idx = wrap_mod((long long)a.off + (long long)a.sm * (m0 - 48 + a_r) + (long long)a.sk * (4 * a_c), a.L);   // stage 1
        const int a_step = wrap_mod(-32LL * a.sm, a.L);
        // B items of a stage: 64 columns x 8 k-chunks of 4 complex -> two per thread.  Lanes run along the contiguous
        // axis of B in global memory: along n (column bn = ptid & 63, chunks ptid >> 6 and + 4), or, when k is the
        // contiguous axis (sb_k == 1: the second GEMM of a CZT reads its B transposed), along k (chunk ptid & 7,
        // columns ptid >> 3 and + 32) with 16-byte loads -- a warp-wide load then touches 4 rows instead of 32.
        const bool kmajor = a.sb_k == 1 && (a.sb_n & 1) == 0 && (a.sb_b & 1) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0;
        int bn_[2], ch_[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            bn_[it] = kmajor ? (ptid >> 3) + 32 * it : (ptid & 63);
            ch_[it] = kmajor ? (ptid & 7) : (ptid >> 6) + 4 * it;
        }
        auto load_b = [&](int S, cpx (&Bv)[2][4]) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int n = n0 + bn_[it], k = S * TC_KS + 4 * ch_[it];
                const cpx* src = Bb + (size_t)n * a.sb_n + (size_t)k * a.sb_k;
                if (kmajor && n < a.N && k + 3 < a.K) {
                    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 2);
                    Bv[it][0] = cmake(v0.x, v0.y);
                    Bv[it][1] = cmake(v0.z, v0.w);
                    Bv[it][2] = cmake(v1.x, v1.y);
                    Bv[it][3] = cmake(v1.z, v1.w);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        Bv[it][e] = (n < a.N && k + e < a.K) ? src[(size_t)e * a.sb_k] : cmake(0.f, 0.f);
                }
            }
        };
        // register prefetch, one stage ahead (a stage lasts longer than an L2 / HBM round trip)
        cpx gb[2][4], ga[4];
        load_b(0, gb);
        if (a_worker && NS > 1) strip_load(a_idx, ga);
        for (int S = 0; S < NS; ++S) {
            const int s = S % TC_STAGES, use = S / TC_STAGES;
            float4 brh[2], brl[2], bih[2], bil[2];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                split4(make_float4(gb[it][0].x, gb[it][1].x, gb[it][2].x, gb[it][3].x), brh[it], brl[it]);
                split4(make_float4(gb[it][0].y, gb[it][1].y, gb[it][2].y, gb[it][3].y), bih[it], bil[it]);
            }
            cpx gs[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) gs[e] = ga[e];
            if (S + 1 < NS) load_b(S + 1, gb);
            if (a_worker && S >= 1 && S + 1 < NS) {
                a_idx += a_step;
                if (a_idx >= a.L) a_idx -= a.L;
                strip_load(a_idx, ga);
            }
            TC_STAMP(0, S, 0);
            mbar_wait(empty0 + 8 * s, (use & 1) ^ 1);       // MMAs of stage S - TC_STAGES (and all older) are done
            TC_STAMP(0, S, 1);
            if (a.debug_mode != 1) {
                if (a_worker && S >= 1) strip_store(-TC_KS * S - 16 + a_r, a_c, gs);
                // ---- B: row n -> Br(k = 0..31), row 64 + n -> Bi(k = 0..31)
                const uint32_t Bhi = Bst0 + (uint32_t)(s * TC_STAGE_BYTES), Blo = Bhi + TC_B_BYTES;
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const uint32_t o0 = sw128(bn_[it], ch_[it]), o1 = sw128(TC_BN + bn_[it], ch_[it]);
                    sts128(Bhi + o0, brh[it]);
                    sts128(Blo + o0, brl[it]);
                    sts128(Bhi + o1, bih[it]);
                    sts128(Blo + o1, bil[it]);
                }
            }
            TC_STAMP(0, S, 2);
            fence_proxy_async();         // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
            TC_STAMP(0, S, 3);
        }
    } else if (warp == TC_ACC_WARPS + TC_PROD_WARPS) {
        // ===================================================================== MMA issuer
        const uint32_t idesc = make_idesc(TC_BM, 2 * TC_BN);
        const uint32_t tiles_s = smem_u32(tiles);
        int S = 0;
        for (int c = 0; c < NC; ++c) {
            const int buf = c & 1;
            TC_STAMP(1, S, 3);
            mbar_wait(tempty0 + 8 * buf, ((c >> 1) & 1) ^ 1);       // accumulate warps have drained this buffer pair
            tc_fence_after();
            const uint32_t tmem_d1 = tmem_base + (uint32_t)(buf * 4 * TC_BN), tmem_d2 = tmem_d1 + 2 * TC_BN;
            const int S_end = min(NS, S + TC_CHUNK);
            for (int first = 1; S < S_end; ++S) {
                const int s = S % TC_STAGES, use = S / TC_STAGES;
                TC_STAMP(1, S, 0);
                mbar_wait(full0 + 8 * s, use & 1);
                tc_fence_after();
                TC_STAMP(1, S, 1);
                if (lane == 0) {
                    const uint32_t Bhi = tiles_s + 2 * TC_AS_BYTES + (uint32_t)(s * TC_STAGE_BYTES), Blo = Bhi + TC_B_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // A window of k-block kb = 2S + h: strip rows [-16 kb, -16 kb + 128) = buffer rows [w0, w0 + 128)
                        const int kb = 2 * S + h;
                        const int w0 = (TC_AW - (kb * TC_KC) % TC_AW) % TC_AW;
                        const uint32_t Ahi = tiles_s + (uint32_t)(w0 * 128), Alo = Ahi + TC_AS_BYTES;
#pragma unroll
                        for (int combo = 0; combo < 3; ++combo) {
                            const uint32_t A = combo == 2 ? Alo : Ahi;
                            const uint32_t Bm = combo == 1 ? Blo : Bhi;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {   // K-steps of 8 tf32 = 32 bytes inside the 128-byte swizzle rows
                                const uint64_t db = make_desc(Bm + 32 * (2 * h + j));
                                if (a.debug_mode != 2 || first) {
                                    umma_tf32(tmem_d1, make_desc(A + 32 * j), db, idesc, first ? 0u : 1u);        // Tr . [Br|Bi]
                                    umma_tf32(tmem_d2, make_desc(A + 32 * (2 + j)), db, idesc, first ? 0u : 1u);  // Ti . [Br|Bi]
                                }
                                first = 0;
                            }
                        }
                    }
                    umma_commit(empty0 + 8 * s);                     // smem stage reusable once these MMAs have read it
                    if (S == S_end - 1) umma_commit(tfull0 + 8 * buf);   // chunk complete in TMEM
                }
                TC_STAMP(1, S, 2);
                __syncwarp();
                first = 0;
            }
        }
    } else {
        // ===================================================================== accumulate warps / epilogue
        const int q = warp & 3, half = warp >> 2;                  // TMEM lane quarter, complex column half
        float accr[32], acci[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) accr[i] = acci[i] = 0.f;
        for (int c = 0; c < NC; ++c) {
            const int buf = c & 1;
            TC_STAMP(2, c, 0);
            mbar_wait(tfull0 + 8 * buf, (c >> 1) & 1);
            tc_fence_after();
            TC_STAMP(2, c, 1);
            const uint32_t d1 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 4 * TC_BN + half * 32);
            const uint32_t d2 = d1 + 2 * TC_BN;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float p[8], m[8];
                tmem_ld8x2(d1 + 8 * j, d2 + TC_BN + 8 * j, p, m);           // Tr Br, Ti Bi
#pragma unroll
                for (int e = 0; e < 8; ++e) accr[8 * j + e] += p[e] - m[e];
                tmem_ld8x2(d1 + TC_BN + 8 * j, d2 + 8 * j, p, m);           // Tr Bi, Ti Br
#pragma unroll
                for (int e = 0; e < 8; ++e) acci[8 * j + e] += p[e] + m[e];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            TC_STAMP(2, c, 2);
        }
        const int m = m0 + q * 32 + lane;
        if (m < a.M) {
            cpx* Cb = a.C + (size_t)b * a.sc_b;
            const cpx* Eb = a.epi ? a.epi + (size_t)b * a.sc_b : nullptr;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int n = n0 + half * 32 + e;
                if (n >= a.N) continue;
                const size_t o = (size_t)m * a.sc_m + (size_t)n * a.sc_n;
                cpx v = cmake(accr[e], acci[e]);
                if (Eb) {
                    const cpx qv = __ldg(Eb + o);
                    v = a.conj_epi ? cmulc(v, qv) : cmul(v, qv);
                }
                Cb[o] = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// out[i] = x[i] * p[i]  (or x[i] * conj(p[i])): folds the prologue factor into the B operand once, so that the GEMM
// producers stream 8 instead of 16 bytes per element (they are latency/L2-bound, see profiles/README.md).
__global__ void __launch_bounds__(256) thz_k_cmul(const cpx* __restrict__ x, const cpx* __restrict__ p, cpx* __restrict__ out,
                                                  size_t n, int conj_p) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const cpx a = x[i], b = __ldg(p + i);
        out[i] = conj_p ? cmulc(a, b) : cmul(a, b);
    }
}

extern "C" int thz_debug_tc_timeline(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_tc_timeline, sizeof(long long) * 3 * TC_TL_N * 4);
}

// NULL if the tcgen05 kernel can serve the call, else the reason (static string).  No launch happens before this check.
const char* thz_toeplitz_gemm_tc_ineligible(const ToeplitzGemmArgs& a, const void* scratch) {
    if (a.pro) {
        const bool dense = a.sb_b == (long long)a.K * a.N && ((a.sb_k == a.N && a.sb_n == 1) || (a.sb_k == 1 && a.sb_n == a.K));
        if (!scratch) return "a prologue factor needs desc->scratch";
        if (!dense) return "a prologue factor needs a dense B operand";
    }
    if (a.sm != -a.sk) return "the sliding-strip A operand needs sm == -sk";   // T(m, k) = g[off + sm (m - k)]
    if (a.L < 64) return "chirp filter shorter than 64";
    return nullptr;
}

int thz_toeplitz_gemm_tc_launch(const ToeplitzGemmArgs& a_in, void* scratch, cudaStream_t stream) {
    ToeplitzGemmArgs a = a_in;
    if (thz_toeplitz_gemm_tc_ineligible(a, scratch)) return thz_set_error(THZ_E_UNSUPPORTED, "thz_toeplitz_gemm_tc_launch: not eligible");
    if (a.pro) {
        const size_t n = (size_t)a.batch * a.K * a.N;
        size_t blocks = (n + 255) / 256;
        const size_t cap = (size_t)thz_sm_count() * 16;
        thz_launch_begin(stream, THZ_KC_CZT_TC);
        thz_k_cmul<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, stream>>>(a.B, a.pro, (cpx*)scratch, n, a.conj_pro);
        thz_launch_end(stream, THZ_KC_CZT_TC);
        a.B = (const cpx*)scratch;
        a.pro = nullptr;
    }
    const size_t smem = (size_t)TC_SMEM_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(thz_k_toeplitz_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return thz_set_cuda_error("cudaFuncSetAttribute(thz_k_toeplitz_gemm_tc)", e);
    dim3 grid((a.N + TC_BN - 1) / TC_BN, (a.M + TC_BM - 1) / TC_BM, a.batch);
    thz_launch_begin(stream, THZ_KC_CZT_TC);
    thz_k_toeplitz_gemm_tc<<<grid, TC_THREADS, smem, stream>>>(a);
    thz_launch_end(stream, THZ_KC_CZT_TC);
    e = cudaGetLastError();
    if (e != cudaSuccess) return thz_set_cuda_error("thz_k_toeplitz_gemm_tc", e);
    return THZ_OK;
}
