// spmv.cu — y = A*x on the DASP layout, hand-written for sm_100a.
//
// Replaces the reference kernels dasp_spmv2<rowloop> + longPart_sum (src/dasp_f64.h:53-484,
// src/dasp_f16.h:106-590).  One fused launch (plus the two below where AUTO picks them); the block index selects the row category like the
// reference does (src/dasp_f64.h:90,145,281,296,357,424), but the geometry is this implementation's:
//
//   long    one warp per work unit (<= 32 reference warps = 2048/8192 slots of ONE row), 32-slot coalesced
//           loads, software pipelined; units are enumerated chunk-major over groups of 8 rows so the warps of
//           a CTA share x sectors in L1; a row that spans several units is merged deterministically by the
//           last unit to arrive (self-resetting counter) — no second launch (K1+K2 of SURVEY §8a).
//           Alternatives behind dasp_set_variant: DMMA tiles, TMA bulk-copy ring.
//   medium  one warp per 4 blocks of 8 rows; lane = (block, row).  CUDA-core variant: every lane walks
//           the 8x4 tiles of its row with one 256-bit value load + one 64-bit load of four 16-bit column offsets
//           per tile and keeps its own accumulator (no cross-lane traffic, same summation order as serial CSR),
//           groups walked in locality order.  Small, L2-resident matrices (KEEP) run the same loop at 4 CTAs per
//           SM - one wave - with programmatic dependent launch; a pipelined form of the loop (round 1) is kept.
//           Alternatives: the reference's DMMA m8n8k4 / HMMA m16n8k16 formulation on the same tiles (K3), 4 lanes
//           per row, x windows in shared memory (mb_kernel), SM-affine queues (smq_kernel).
//   Separate launches for matrices whose gathers do not coalesce: lcb_kernel (long rows, column-blocked copy, x
//   blocks staged in shared memory by TMA), sb_kernel (short rows by row band, x windows staged by TMA).
//   short   1 / 1&3 / 3&4 / 2&2 segments read as flat coalesced streams (alignment-free: the FP64
//           short segments start at slot short_row_1, which is not a multiple of 4) and folded with
//           shuffles inside each 4-slot tile row (K4-K7).
//   zero    rows without entries are written as 0 every call (the reference relies on a one-time
//           cudaMemset, src/dasp_f64.h:1242).
//
// y is produced in permuted order (K11); with `scatter` (= order_rid) it is written to original order.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "dasp_internal.h"

namespace dasp {
namespace {

constexpr int CTA = SPMV_CTA;
constexpr int WARPS = CTA / 32;
#ifndef MED_TB
#define MED_TB 4 // tiles per batch of the large-matrix medium kernel (compact-index path)
#endif
#ifndef MED_MINB
#define MED_MINB 6 // CTAs per SM the large-matrix kernels are compiled for
#endif
#ifndef KEEP_MINB
#define KEEP_MINB 3 // CTAs per SM the small-matrix (KEEP) kernels are compiled for (<= 80 registers)
#endif

struct SpmvArgs {
    const void *x;
    void *y;
    const int *scatter; // nullptr: permuted order
    double alpha, beta; // y = alpha*A*x + beta*y when axpby != 0
    int axpby;
    int out_f32; // FP16 matrices only: y (and the extra destinations) are float, not half
    // fused exchange (dasp_spmv_scatter_to): the result is also stored into n_extra more vectors (peer GPUs' copies of
    // the next x, or one NVSwitch multicast mapping), at element offset row_offset, scaled by 1/sqrt(*rs_ptr)
    void *y_extra[7];
    int n_extra;
    long row_offset;
    const double *rs_ptr;
    // long
    const void *long_val;
    const int *long_cid, *long_rpt_new, *unit_row, *unit_chunk, *unit_first;
    const int *long_cbase;             // compact indices of the long part: base per 32-slot group
    const unsigned short *long_cdelta; //   16-bit offsets, 0xFFFF = column 0
    const unsigned char *long_wide;    //   per unit (execution order): 1 = read long_cid; nullptr = compression off
    void *partial;
    unsigned *done;
    int n_units, longw, unit_warps;
    // medium
    const void *reg_val;
    const int *reg_cid, *blockPtr, *irreg_rpt;
    const void *irreg_val;
    const int *irreg_cid;
    const unsigned char *has_irreg;
    const int *reg_cbase;             // compact indices: per-tile base
    const unsigned short *reg_cdelta; //                  per-slot 16-bit offset, 0xFFFF = column 0
    const unsigned char *blk_wide;    // nullptr: compression off; else 1 = block reads reg_cid
    const unsigned short *blk_live;   // tiles of the block worth reading (trailing all-zero tiles dropped)
    const int *med_order;             // locality order of the 32-row groups (nullptr: identity)
    // SM-affine queues (small matrices): the groups, in locality order, are cut into smq_n contiguous ranges (one per SM) of
    // smq_k chunks of <= 8 groups; a medium CTA takes the next chunk of the SM it runs on (nullptr: CTA index = chunk)
    int *smq_cnt;
    int smq_n, smq_k;
    int keep_compact; // small-matrix kernels read the compact column indices (chosen per launch, see medium_rows)
    const int *short_map;             // locality order of the short CTAs: category << 28 | CTA inside it (nullptr: off)
    int row_long, row_block, blocknum;
    // short
    const void *short_val;
    const int *short_cid;
    int n1, c13, n34, n2;
    int s1, s13, s34, s22;     // slot bases
    int y1, y13, y34, y22, y0; // y bases (K11)
    int G;                     // 8 (f64) / 32 (f16)
    int row_zero;
    // column-blocked long rows (LCB, derive.cu)
    const void *lcb_val;
    const unsigned *lcb_idx; // long row << 16 | column inside the block
    const unsigned short *lcb_idx16; // FP64: column | row delta << 13 (nullptr: read lcb_idx)
    const int *lcb_chunk_row;        //   restart row of every 1024-entry chunk
    const unsigned char *lcb_chunk_wide; // 1: the chunk has a delta > 7 and is read through lcb_idx
    const int *lcb_blk_ptr, *lcb_cta_first;
    void *lcb_acc;
    unsigned *lcb_done;
    int lcb_bw_log2, lcb_nblk, lcb_nctas, ncols;
    // medium-band kernel (MB, derive.cu): first column of the x window of every CTA (8 groups in processing order)
    const int *mb_lo;
    int mb_wcap;
    // short-band kernel (SB, derive.cu): warp items of the four short segments sorted by row band
    const int *sb_item;     // segment << 28 | warp item inside the segment
    const int *sb_band_ptr; // [sb_nbands + 1] first item of every band
    const int *sb_lo;       // [sb_nbands] first column of the band's x window
    int sb_nbands, sb_wcap;
    int e[7];      // exclusive CTA-range end of category k (long, medium, singles, 1&3, 3/4, 2&2, zero)
    long items[7]; // warp-level work items of category k
};

template <typename T> struct Acc;
template <> struct Acc<double> { using type = double; };
template <> struct Acc<__half> { using type = float; };

__device__ __forceinline__ double to_acc(double v) { return v; }
__device__ __forceinline__ float to_acc(__half v) { return __half2float(v); }
__device__ __forceinline__ void from_acc(double *p, double v) { *p = v; }
__device__ __forceinline__ void from_acc(__half *p, float v) { *p = __float2half_rn(v); }

// ---- streaming loads of the packed matrix: read once, so keep them out of L1 and mark them evict-first in
// L2 (x must survive there).  Only the 256-bit form takes .L2::evict_first directly; narrower loads go
// through a createpolicy descriptor.
// keep != 0: the whole layout fits in L2 (small matrices iterated back to back), so it is left at normal
// priority and later launches hit in L2; otherwise evict-first.
// (compile-time: only the 256-bit load form takes the L2 priority as an instruction modifier)
struct StreamPol {
    uint64_t desc;
};
template <bool KEEP> __device__ __forceinline__ StreamPol make_stream_policy()
{
    StreamPol p;
    if (KEEP) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p.desc));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.desc));
    return p;
}
#define DASP_LD_HINT "ld.global.nc.L1::no_allocate.L2::cache_hint"
template <bool KEEP> __device__ __forceinline__ void ld_stream4d(const double *p, double (&v)[4])
{
    if (KEEP)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
template <bool KEEP> __device__ __forceinline__ void ld_stream4(const double *p, double (&v)[4], const StreamPol &) { ld_stream4d<KEEP>(p, v); }
template <bool KEEP> __device__ __forceinline__ void ld_stream4(const __half *p, __half (&v)[4], const StreamPol &pol)
{
    unsigned a, b;
    asm volatile(DASP_LD_HINT ".v2.u32 {%0,%1}, [%2], %3;" : "=r"(a), "=r"(b) : "l"(p), "l"(pol.desc));
    __half2 h0 = *reinterpret_cast<__half2 *>(&a), h1 = *reinterpret_cast<__half2 *>(&b);
    v[0] = __low2half(h0); v[1] = __high2half(h0); v[2] = __low2half(h1); v[3] = __high2half(h1);
}
template <bool KEEP> __device__ __forceinline__ void ld_stream4(const int *p, int (&v)[4], const StreamPol &pol)
{
    asm volatile(DASP_LD_HINT ".v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p), "l"(pol.desc));
}
__device__ __forceinline__ double ld_stream1(const double *p, const StreamPol &pol)
{
    double v;
    asm volatile(DASP_LD_HINT ".f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol.desc));
    return v;
}
__device__ __forceinline__ __half ld_stream1(const __half *p, const StreamPol &pol)
{
    unsigned short v;
    asm volatile(DASP_LD_HINT ".u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol.desc));
    return __ushort_as_half(v);
}
__device__ __forceinline__ int ld_stream1(const unsigned short *p, const StreamPol &pol)
{
    unsigned short v;
    asm volatile(DASP_LD_HINT ".u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol.desc));
    return (int)v;
}
__device__ __forceinline__ int ld_stream1(const int *p, const StreamPol &pol)
{
    int v;
    asm volatile(DASP_LD_HINT ".s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol.desc));
    return v;
}

// ---- TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP) and mbarrier helpers for the long-row stream ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
constexpr int TMA_STAGES = 3;          // stages of the per-warp ring
constexpr int TMA_STAGE_SLOTS = 128;   // slots per stage: 1 KB of FP64 values + 512 B of indices
constexpr int TMA_STAGE_BYTES = TMA_STAGE_SLOTS * 12;
constexpr int TMA_WARP_BYTES = TMA_STAGES * TMA_STAGE_BYTES;
constexpr int TMA_SMEM_BYTES = WARPS * TMA_WARP_BYTES + WARPS * TMA_STAGES * 8;

// Programmatic dependent launch (small matrices): a KEEP kernel may start while its predecessor on the stream is
// still draining.  Everything read before pdl_wait() is immutable matrix data; x (possibly the predecessor's y) and y
// are only touched after it.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// x gathers: read-only path, allocate in L1 (neighbouring rows reuse the same entries)
template <typename T> __device__ __forceinline__ typename Acc<T>::type gather(const T *x, int c) { return to_acc(__ldg(x + c)); }

// A window x[lo, lo + len) staged in shared memory (short-band kernel): columns inside it are read from there (a scattered
// 32-lane gather costs a few bank-conflict cycles instead of 32 L1 wavefronts), the others from global memory.
template <typename T> struct XWin {
    const T *xs;
    int lo;
    unsigned len;
    uint32_t bar = 0; // mbarrier (shared-memory address) the window's TMA copies complete on; 0: the window is already there
};
template <typename T> __device__ __forceinline__ typename Acc<T>::type gather(const T *x, int c, const XWin<T> *w)
{
    if (w) {
        const unsigned d = (unsigned)(c - w->lo);
        if (d < w->len) return to_acc(w->xs[d]);
    }
    return to_acc(__ldg(x + c));
}

// element `idx` of an output vector: the handle's value type, or float for an FP16 matrix asked for FP32 output
template <typename T> __device__ __forceinline__ void put(const SpmvArgs &a, void *base, long idx, typename Acc<T>::type v)
{
    if constexpr (sizeof(T) == 2) {
        if (a.out_f32) { static_cast<float *>(base)[idx] = v; return; }
    }
    from_acc(static_cast<T *>(base) + idx, v);
}
template <typename T> __device__ __forceinline__ typename Acc<T>::type get(const SpmvArgs &a, const void *base, long idx)
{
    if constexpr (sizeof(T) == 2) {
        if (a.out_f32) return static_cast<const float *>(base)[idx];
    }
    return to_acc(static_cast<const T *>(base)[idx]);
}

template <typename T>
__device__ __forceinline__ void store_y(const SpmvArgs &a, long idx, typename Acc<T>::type v)
{
    using A = typename Acc<T>::type;
    if (a.scatter) idx = a.scatter[idx];
    if (a.axpby) { // one flag for every non-default form: alpha/beta, 1/sqrt(norm^2) scaling, offset, extra destinations
        v = (A)a.alpha * v + (a.beta != 0.0 ? (A)a.beta * get<T>(a, a.y, idx) : A(0));
        if (a.rs_ptr) v *= (A)rsqrt(__ldg(a.rs_ptr));
        idx += a.row_offset;
#pragma unroll 1
        for (int p = 0; p < a.n_extra; p++) put<T>(a, a.y_extra[p], idx, v); // P2P / multicast stores
    }
    put<T>(a, a.y, idx, v);
}

template <typename A> __device__ __forceinline__ A warp_sum(A v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// FP16 tensor-core step: D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col)  -> SASS HMMA.16816.F32.
// The reference's FP16 kernels use mma.m8n8k4.f16 (src/dasp_f16.h:31-75), which sm_100a only emulates; m16n8k16 is the
// native shape.  Fragment map: lane (g = lane>>2, t = lane&3) holds A[g][2t,2t+1] (a0), A[g+8][..] (a1), A[g][2t+8,2t+9]
// (a2), A[g+8][..] (a3), B[2t,2t+1][g] (b0), B[2t+8,2t+9][g] (b1), D[g][2t,2t+1] (d0,d1), D[g+8][..] (d2,d3).
// DASP use: A row g = 16 consecutive entries of matrix row g (four 8x4 tiles), B column g = x gathered for exactly those
// entries by the SAME lane, so D[g][g] is the dot product; rows 8..15 of A are zero (a1 = a3 = 0).
__device__ __forceinline__ void hmma16816(float (&d)[4], unsigned a0, unsigned a2, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned pack_half2(__half lo, __half hi)
{
    return (unsigned)__half_as_ushort(lo) | ((unsigned)__half_as_ushort(hi) << 16);
}

// ------------------------------------------------------------------------------------------------
// long rows

// LONGV: 0 CUDA-core with register-staged loads, 1 DMMA, 2 CUDA-core fed by a per-warp TMA bulk-copy ring
template <typename T, int LONGV, bool KEEP>
__device__ __forceinline__ void long_rows(const SpmvArgs &a, long w, unsigned char *smem)
{
    constexpr bool MMA = LONGV == 1;
    const StreamPol pol = make_stream_policy<KEEP>();
    using A = typename Acc<T>::type;
    const int lane = threadIdx.x & 31;
    if (w >= a.n_units) return;
    const int u = (int)w;
    const T *x = static_cast<const T *>(a.x);
    const int row = __ldg(a.unit_row + u), chunk = __ldg(a.unit_chunk + u);
    const int first = __ldg(a.unit_first + row), nunits = __ldg(a.unit_first + row + 1) - first;
    const long row_beg = (long)__ldg(a.long_rpt_new + row) * a.longw, row_end = (long)__ldg(a.long_rpt_new + row + 1) * a.longw;
    const long beg = row_beg + (long)chunk * a.unit_warps * a.longw;
    const long end = min(beg + (long)a.unit_warps * a.longw, row_end);
    const T *val = static_cast<const T *>(a.long_val);
    A acc = 0;
    if constexpr (MMA && sizeof(T) == 8) {
        // the reference's formulation (src/dasp_f64.h:105-122): 32 slots per DMMA, the useful products sit on
        // the diagonal of C; everything in C is summed, off-diagonal terms are masked out by zeroing B.
        double c[2] = {0.0, 0.0};
        const int grp = lane >> 2;
        for (long p = beg + lane; p < end; p += 128) {
            double av[4], bv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                long q = p + 32 * j;
                bool ok = q < end;
                av[j] = ok ? (double)ld_stream1(reinterpret_cast<const double *>(val) + q, pol) : 0.0;
                int cid = ok ? ld_stream1(a.long_cid + q, pol) : 0;
                bv[j] = ok ? (double)gather(reinterpret_cast<const double *>(x), cid) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(c, av[j], bv[j]);
        }
        // C[g][n]: lane holds row g = lane>>2, columns 2*(lane&3)+{0,1}; keep only n == g
        const int n0 = 2 * (lane & 3);
        double d = (n0 == grp ? c[0] : 0.0) + (n0 + 1 == grp ? c[1] : 0.0);
        acc = (A)warp_sum(d);
    } else if constexpr (MMA && sizeof(T) == 2) {
        // FP16 tensor-core variant: 128 slots per HMMA step, lane (g, t) owns slots p + 16 g + 4 t .. +3 (one 64-bit value
        // load, one 128-bit index load, four gathers); the eight partial sums sit on the diagonal of D.
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        const int g = lane >> 2, t = lane & 3;
        const __half *hval = reinterpret_cast<const __half *>(val);
        const __half *hx = reinterpret_cast<const __half *>(x);
        for (long p = beg + 16 * g + 4 * t; p < end; p += 128) { // units are multiples of 256 slots: no partial step
            __half v[4];
            int c[4];
            ld_stream4<KEEP>(hval + p, v, pol);
            ld_stream4<KEEP>(a.long_cid + p, c, pol);
            __half xv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) xv[j] = __ldg(hx + c[j]);
            hmma16816(d, pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(xv[0], xv[1]), pack_half2(xv[2], xv[3]));
        }
        const int n0 = 2 * t; // D[g][n0], D[g][n0+1]: keep only the diagonal
        const float dd = (n0 == g ? d[0] : 0.f) + (n0 + 1 == g ? d[1] : 0.f);
        acc = (A)warp_sum(dd);
    } else if constexpr (LONGV == 2) {
        // The dense value/index streams of the unit are moved by TMA 1-D bulk copies into a 3-stage shared-memory
        // ring owned by this warp (lane 0 is the producer, one mbarrier per stage); the lanes read their slots from
        // shared memory (conflict-free 8-byte / 4-byte accesses) and only the x gathers go through L1.  Bytes in
        // flight per warp: 3 x 1.5 KB without holding a single register.
        const int warp = threadIdx.x >> 5;
        unsigned char *ring = smem + warp * TMA_WARP_BYTES;
        const uint32_t bar0 = smem_u32(smem + WARPS * TMA_WARP_BYTES + warp * TMA_STAGES * 8);
        if (lane == 0) {
#pragma unroll
            for (int st = 0; st < TMA_STAGES; st++) mbar_init(bar0 + 8 * st, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        const int nstage = (int)((end - beg + TMA_STAGE_SLOTS - 1) / TMA_STAGE_SLOTS);
        auto issue = [&](int it) { // lane 0 only
            const int st = it % TMA_STAGES;
            const long q = beg + (long)it * TMA_STAGE_SLOTS;
            const uint32_t n = (uint32_t)min((long)TMA_STAGE_SLOTS, end - q);
            const uint32_t dst = smem_u32(ring + st * TMA_STAGE_BYTES);
            mbar_expect_tx(bar0 + 8 * st, n * (uint32_t)(sizeof(T) + 4));
            bulk_g2s(dst, val + q, n * (uint32_t)sizeof(T), bar0 + 8 * st, pol.desc);
            bulk_g2s(dst + TMA_STAGE_SLOTS * 8, a.long_cid + q, n * 4u, bar0 + 8 * st, pol.desc);
        };
        if (lane == 0)
            for (int it = 0; it < TMA_STAGES && it < nstage; it++) issue(it);
        A s0 = 0, s1 = 0;
        for (int it = 0; it < nstage; it++) {
            const int st = it % TMA_STAGES;
            mbar_wait(bar0 + 8 * st, (uint32_t)((it / TMA_STAGES) & 1));
            const T *sv = reinterpret_cast<const T *>(ring + st * TMA_STAGE_BYTES);
            const int *sc = reinterpret_cast<const int *>(ring + st * TMA_STAGE_BYTES + TMA_STAGE_SLOTS * 8);
            const int n = (int)min((long)TMA_STAGE_SLOTS, end - (beg + (long)it * TMA_STAGE_SLOTS));
            T v[4];
            int c[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool ok = lane + 32 * j < n;
                v[j] = ok ? sv[lane + 32 * j] : T(0);
                c[j] = ok ? sc[lane + 32 * j] : 0;
            }
            __syncwarp(); // every lane has its slots in registers: the stage may be refilled
            if (lane == 0 && it + TMA_STAGES < nstage) issue(it + TMA_STAGES);
            A g[4];
#pragma unroll
            for (int j = 0; j < 4; j++) g[j] = gather(x, c[j]);
            s0 += to_acc(v[0]) * g[0] + to_acc(v[2]) * g[2];
            s1 += to_acc(v[1]) * g[1] + to_acc(v[3]) * g[3];
        }
        acc = warp_sum(s0 + s1);
    } else {
        // 32 slots per warp load (lane, lane+32, ...): fully coalesced value/index streams and, for ascending
        // columns, x gathers that share 128-byte lines inside one instruction.  Software pipelined by hand (the
        // asm loads keep program order): the streams of batch i+1 are in flight while the gathers of batch i
        // are pending.  Units are multiples of 64 slots, the row tail is zero padding; loads past `end` are
        // predicated off.
        constexpr int LB = 4; // slots per lane per batch = 128 slots per warp
        A s0 = 0, s1 = 0;
        // column indices: compact form (one 32-bit base per 32-slot group + 16-bit offsets, decoded only when the
        // gathers are issued so that nothing depends on a load while the next loads are being requested) unless
        // this unit holds a group spanning >= 65535 columns
        auto run = [&](auto tag) {
            constexpr bool CP = decltype(tag)::value;
            T v0[LB], v1[LB];
            int c0[LB], c1[LB], b0[LB], b1[LB];
            auto load = [&](T(&v)[LB], int(&c)[LB], int(&bs)[LB], long q) {
#pragma unroll
                for (int j = 0; j < LB; j++) {
                    const bool ok = q + 32 * j < end;
                    v[j] = ok ? ld_stream1(val + q + 32 * j, pol) : T(0);
                    if constexpr (CP) {
                        c[j] = ok ? ld_stream1(a.long_cdelta + q + 32 * j, pol) : 0xFFFF;
                        bs[j] = ok ? __ldg(a.long_cbase + ((q + 32 * j) >> 5)) : 0;
                    } else {
                        c[j] = ok ? ld_stream1(a.long_cid + q + 32 * j, pol) : 0;
                    }
                }
            };
            auto consume = [&](const T(&v)[LB], const int(&c)[LB], const int(&bs)[LB]) {
                A g[LB];
#pragma unroll
                for (int j = 0; j < LB; j++) {
                    if constexpr (CP) g[j] = gather(x, c[j] == 0xFFFF ? 0 : bs[j] + c[j]);
                    else g[j] = gather(x, c[j]);
                }
#pragma unroll
                for (int j = 0; j < LB; j += 2) { s0 += to_acc(v[j]) * g[j]; s1 += to_acc(v[j + 1]) * g[j + 1]; }
            };
            long p = beg + lane;
            load(v0, c0, b0, p);
            for (; p < end; p += 64 * LB) {
                load(v1, c1, b1, p + 32 * LB);
                consume(v0, c0, b0);
                load(v0, c0, b0, p + 64 * LB);
                consume(v1, c1, b1);
            }
        };
        if (a.long_wide != nullptr && __ldg(a.long_wide + u) == 0) run(std::true_type{});
        else run(std::false_type{});
        acc = warp_sum(s0 + s1);
    }
    if (nunits == 1) {
        if (lane == 0) store_y<T>(a, row, acc);
        return;
    }
    // multi-unit row: publish the partial, the last unit to arrive folds them in unit order (deterministic)
    A *partial = static_cast<A *>(a.partial);
    unsigned prev = 0;
    if (lane == 0) {
        __stcg(partial + first + chunk, acc);
        __threadfence();
        prev = atomicAdd(a.done + row, 1u);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev != (unsigned)(nunits - 1)) return;
    __threadfence();
    A t = 0;
    for (int i = lane; i < nunits; i += 32) t += __ldcg(partial + first + i);
    t = warp_sum(t);
    if (lane == 0) {
        store_y<T>(a, row, t);
        a.done[row] = 0; // ready for the next call
    }
}

// ------------------------------------------------------------------------------------------------
// medium rows (row blocks)

// LEAN (with KEEP): the register-lean loop of the large matrices with the L2 policy and the dependent-launch wait of the small ones
template <typename T, bool MMA, bool KEEP, bool LEAN = false, int PB = 0>
__device__ __forceinline__ void medium_rows(const SpmvArgs &a, long w, const XWin<T> *win = nullptr)
{
    const StreamPol pol = make_stream_policy<KEEP>();
    using A = typename Acc<T>::type;
    const int lane = threadIdx.x & 31;
    if (w * 4 >= a.blocknum) return;
    // 32 rows = 4 blocks; large matrices walk the groups in order of the original id of their first row
    const int group = (!MMA && a.med_order) ? __ldg(a.med_order + w) : (int)w; // (the launch passes med_order only where it is used)
    const T *x = static_cast<const T *>(a.x);
    const T *val = static_cast<const T *>(a.reg_val);
    const int g = group * 32 + lane;
    A acc = 0;
    if constexpr (MMA && sizeof(T) == 8) {
        // DMMA m8n8k4 per 8x4 tile, one block after the other (src/dasp_f64.h:240-269 restated)
        const double *dval = reinterpret_cast<const double *>(val);
        const double *dx = reinterpret_cast<const double *>(x);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = group * 4 + i;
            const int bp0 = __ldg(a.blockPtr + b), bp1 = __ldg(a.blockPtr + b + 1);
            double c[2] = {0.0, 0.0};
            int p = bp0 + lane;
            for (; p + 96 < bp1; p += 128) {
                double av[4], bv[4];
                int cid[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { av[j] = ld_stream1(dval + p + 32 * j, pol); cid[j] = ld_stream1(a.reg_cid + p + 32 * j, pol); }
#pragma unroll
                for (int j = 0; j < 4; j++) bv[j] = __ldg(dx + cid[j]);
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(c, av[j], bv[j]);
            }
            for (; p < bp1; p += 32) {
                double av = ld_stream1(dval + p, pol);
                double bv = __ldg(dx + ld_stream1(a.reg_cid + p, pol));
                dmma884(c, av, bv);
            }
            // C[r][r] sits in lane 4r + (r>>1), register r&1; route it to lane 8i + r
            const int r = lane & 7, src = 4 * r + (r >> 1);
            double v0 = __shfl_sync(0xffffffffu, c[0], src), v1 = __shfl_sync(0xffffffffu, c[1], src);
            if ((lane >> 3) == i) acc = (A)((r & 1) ? v1 : v0);
        }
    } else if constexpr (MMA && sizeof(T) == 2) {
        // FP16 tensor-core variant (HMMA.16816.F32): per step four 8x4 tiles of one block; lane (q, t) owns row q of tile
        // k + t: one 64-bit value load + one 128-bit index load + four gathers feed A[q][..] and B[..][q] of the same lane.
        // FP16 blocks are multiples of 4 tiles (src/dasp_f16.h:1356); all-zero trailing steps are skipped (blk_live).
        const __half *hval = reinterpret_cast<const __half *>(val);
        const __half *hx = reinterpret_cast<const __half *>(x);
        const int q = lane >> 2, t = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = group * 4 + i;
            const int bp0 = __ldg(a.blockPtr + b), bp1 = __ldg(a.blockPtr + b + 1);
            const int nt = min((bp1 - bp0) >> 5, ((int)__ldg(a.blk_live + b) + 3) & ~3);
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            const __half *pv = hval + bp0 + 32 * t + 4 * q;
            const int *pc = a.reg_cid + bp0 + 32 * t + 4 * q;
            for (int k = 0; k < nt; k += 8) { // two steps in flight
                __half v[2][4], xv[2][4];
                int c[2][4];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (k + 4 * j < nt) { ld_stream4<KEEP>(pv + 32 * (k + 4 * j), v[j], pol); ld_stream4<KEEP>(pc + 32 * (k + 4 * j), c[j], pol); }
                    else {
#pragma unroll
                        for (int e = 0; e < 4; e++) { v[j][e] = __ushort_as_half(0); c[j][e] = 0; }
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int e = 0; e < 4; e++) xv[j][e] = __ldg(hx + c[j][e]);
#pragma unroll
                for (int j = 0; j < 2; j++)
                    hmma16816(d, pack_half2(v[j][0], v[j][1]), pack_half2(v[j][2], v[j][3]), pack_half2(xv[j][0], xv[j][1]),
                              pack_half2(xv[j][2], xv[j][3]));
            }
            // D[r][r] sits in lane 4r + (r>>1), register r&1; route it to lane 8i + r
            const int r = lane & 7, src = 4 * r + (r >> 1);
            const float v0 = __shfl_sync(0xffffffffu, d[0], src), v1 = __shfl_sync(0xffffffffu, d[1], src);
            if ((lane >> 3) == i) acc = (A)((r & 1) ? v1 : v0);
        }
    } else {
        const int b = g >> 3, r = g & 7;
        const int bp0 = __ldg(a.blockPtr + b), bp1 = __ldg(a.blockPtr + b + 1);
        // Bounds of the irregular tail, requested together with the block bounds so that the dependent chain of a
        // row is {bounds} -> {tiles, tail entries} -> {gathers}.  Large matrices skip the read for the 32-row
        // groups that have no irregular entry at all (1 flag byte per group instead of 128 bytes of irreg_rpt).
        int lo = 0, hi = 0;
        if (g < a.row_block && (KEEP || a.has_irreg[group])) { lo = __ldg(a.irreg_rpt + g); hi = __ldg(a.irreg_rpt + g + 1); }
        const T *pv = val + bp0 + 4 * r;
        const int *pc = a.reg_cid + bp0 + 4 * r;
        const T *iv = static_cast<const T *>(a.irreg_val);
        bool arrived = false; // medium-band kernel: the x window is being copied into shared memory while the tiles load
        auto window_ready = [&]() {
            if (win && win->bar && !arrived) mbar_wait(win->bar, 0);
            arrived = true;
        };
        // FP16 blocks are padded to 4 tiles (src/dasp_f16.h:1356): stop at the last tile that holds a value; FP64 blocks are not
        const int nt = sizeof(T) == 2 ? min((bp1 - bp0) >> 5, (int)__ldg(a.blk_live + b)) : (bp1 - bp0) >> 5;
        // column indices of tile k: compact form (tile base + 16-bit offsets) unless a block of this warp is flagged
        // wide (warp-uniform choice: no divergence, one code path live at a time)
        const bool compact = __all_sync(0xffffffffu, a.blk_wide != nullptr && a.blk_wide[b] == 0);
        const unsigned short *pd = a.reg_cdelta + bp0 + 4 * r;
        const int *pb = a.reg_cbase + (bp0 >> 5);
        if constexpr (!KEEP || LEAN) {
            // Large matrices (bandwidth-bound): four tiles (4 x (256-bit values + 128-bit indices)) in flight per
            // lane, consumed as they arrive (40 registers, 6 CTAs per SM); the last batch is predicated.
            // Small matrices (KEEP && LEAN): the first two entries of the irregular tail are requested with the first tiles
            // (a tail of up to two entries then adds one round trip to the row's chain instead of four).
            constexpr int IRP = (KEEP && LEAN) ? 2 : 0;
            T wv[IRP > 0 ? IRP : 1];
            int wc[IRP > 0 ? IRP : 1];
            if constexpr (IRP > 0) {
#pragma unroll
                for (int j = 0; j < IRP; j++) {
                    const bool ok = lo + j < hi;
                    wv[j] = ok ? ld_stream1(iv + lo + j, pol) : T(0);
                    wc[j] = ok ? ld_stream1(a.irreg_cid + lo + j, pol) : 0;
                }
            }
            if (compact) {
                // compact indices: 8 bytes of 16-bit offsets + a broadcast 32-bit tile base per tile, decoded just
                // before the gathers so that only the packed form is live while the loads are in flight
                constexpr int TB = (LEAN && PB > 0) ? PB : MED_TB; // tiles per batch
                for (int k = 0; k < nt; k += TB) {
                    T v[TB][4];
                    unsigned d[TB][2];
                    int base[TB];
#pragma unroll
                    for (int j = 0; j < TB; j++) {
                        if (k + j < nt) {
                            ld_stream4<KEEP>(pv + 32 * (k + j), v[j], pol);
                            asm volatile(DASP_LD_HINT ".v2.u32 {%0,%1}, [%2], %3;"
                                         : "=r"(d[j][0]), "=r"(d[j][1]) : "l"(pd + 32 * (k + j)), "l"(pol.desc));
                            base[j] = __ldg(pb + k + j);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; e++) v[j][e] = T(0);
                            d[j][0] = d[j][1] = 0xFFFFFFFFu;
                            base[j] = 0;
                        }
                    }
                    if (k == 0) window_ready();
                    if constexpr (KEEP) { if (k == 0) pdl_wait(); } // x may be the previous product's y
                    A xv[TB][4];
#pragma unroll
                    for (int j = 0; j < TB; j++)
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const unsigned h16 = (e & 1) ? (d[j][e >> 1] >> 16) : (d[j][e >> 1] & 0xFFFFu);
                            xv[j][e] = gather(x, h16 == 0xFFFFu ? 0 : base[j] + (int)h16, win);
                        }
#pragma unroll
                    for (int j = 0; j < TB; j++)
#pragma unroll
                        for (int e = 0; e < 4; e++) acc += to_acc(v[j][e]) * xv[j][e];
                }
            } else {
                for (int k = 0; k < nt; k += 4) {
                    T v[4][4];
                    int c[4][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (k + j < nt) { ld_stream4<KEEP>(pv + 32 * (k + j), v[j], pol); ld_stream4<KEEP>(pc + 32 * (k + j), c[j], pol); }
                        else {
#pragma unroll
                            for (int e = 0; e < 4; e++) { v[j][e] = T(0); c[j][e] = 0; }
                        }
                    }
                    if (k == 0) window_ready();
                    if constexpr (KEEP) { if (k == 0) pdl_wait(); }
                    A xv[4][4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int e = 0; e < 4; e++) xv[j][e] = gather(x, c[j][e], win);
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int e = 0; e < 4; e++) acc += to_acc(v[j][e]) * xv[j][e];
                }
            }
            window_ready(); // rows without a regular tile reach their first gather here
            if constexpr (KEEP) pdl_wait();
            if constexpr (IRP > 0) {
                A xw[IRP];
#pragma unroll
                for (int j = 0; j < IRP; j++) xw[j] = gather(x, wc[j], win);
#pragma unroll
                for (int j = 0; j < IRP; j++) acc += to_acc(wv[j]) * xw[j];
            }
            for (int i = lo + IRP; i < hi; i++) acc += to_acc(ld_stream1(iv + i, pol)) * gather(x, ld_stream1(a.irreg_cid + i, pol), win);
        } else {
            // Small, L2-resident matrices (latency-bound: one launch is a handful of dependent round trips): B
            // tiles per batch, two batches in flight, software pipelined by hand (the asm loads keep program order) so
            // the streams of batch i+1 are requested before the gathers of batch i are consumed, and the first
            // entries of the irregular tail travel with the first batch.  FMA order = CSR order in both paths.
            // measured on C1/C2 (profiles/r01/README.md): FP64 10.6/10.7/12.1/14.3 us for B = 1/2/3/4, FP16 (blocks are
            // multiples of 4 tiles) 10.2/9.9/10.2/9.2 us
            constexpr int B = PB > 0 ? PB : (sizeof(T) == 8 ? 2 : 4);
            // The column indices of a batch travel as reg_cid (32 bits) or in the compact form (one 64-bit load of four 16-bit
            // offsets + the tile base): 2 instead of 4 index bytes per entry, decoded when the gathers are issued.  The compact
            // form wins where the index bytes dominate the stream (FP16: 6.4 vs 7.2 us on the C2 stand-in) and is chosen per
            // launch (a.keep_compact); a warp with a wide block reads reg_cid.
            auto pipe = [&](auto compact_c) {
                constexpr bool CP = decltype(compact_c)::value;
                constexpr int NI = CP ? 3 : 4; // index registers per tile: two packed words + base, or four columns
                T va[B][4], vb[B][4];
                int ca[B][NI], cb[B][NI];
                auto load = [&](T(&v)[B][4], int(&c)[B][NI], int k) {
#pragma unroll
                    for (int j = 0; j < B; j++) {
                        if (k + j < nt) {
                            ld_stream4<KEEP>(pv + 32 * (k + j), v[j], pol);
                            if constexpr (CP) {
                                asm volatile(DASP_LD_HINT ".v2.u32 {%0,%1}, [%2], %3;"
                                             : "=r"(c[j][0]), "=r"(c[j][1]) : "l"(pd + 32 * (k + j)), "l"(pol.desc));
                                c[j][2] = __ldg(pb + k + j);
                            } else {
                                int t4[4];
                                ld_stream4<KEEP>(pc + 32 * (k + j), t4, pol);
#pragma unroll
                                for (int e = 0; e < 4; e++) c[j][e] = t4[e];
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; e++) v[j][e] = T(0);
#pragma unroll
                            for (int e = 0; e < NI; e++) c[j][e] = CP ? (e < 2 ? -1 : 0) : 0; // 0xFFFF offsets = column 0
                        }
                    }
                };
                auto consume = [&](const T(&v)[B][4], const int(&c)[B][NI]) {
                    A xv[B][4];
#pragma unroll
                    for (int j = 0; j < B; j++)
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            int col;
                            if constexpr (CP) {
                                const unsigned w = (unsigned)c[j][e >> 1];
                                const unsigned h16 = (e & 1) ? (w >> 16) : (w & 0xFFFFu);
                                col = h16 == 0xFFFFu ? 0 : c[j][2] + (int)h16;
                            } else col = c[j][e];
                            xv[j][e] = gather(x, col, win);
                        }
#pragma unroll
                    for (int j = 0; j < B; j++)
#pragma unroll
                        for (int e = 0; e < 4; e++) acc += to_acc(v[j][e]) * xv[j][e];
                };
                load(va, ca, 0);
                constexpr int IR = 2;
                T wv[IR];
                int wc[IR];
#pragma unroll
                for (int j = 0; j < IR; j++) {
                    const bool ok = lo + j < hi;
                    wv[j] = ok ? ld_stream1(iv + lo + j, pol) : T(0);
                    wc[j] = ok ? ld_stream1(a.irreg_cid + lo + j, pol) : 0;
                }
                pdl_wait(); // streams of the first batch are in flight; x and y may belong to the previous kernel
                window_ready();
                for (int k = 0; k < nt; k += 2 * B) {
                    load(vb, cb, k + B);
                    consume(va, ca);
                    load(va, ca, k + 2 * B);
                    consume(vb, cb);
                }
                A xw[IR];
#pragma unroll
                for (int j = 0; j < IR; j++) xw[j] = gather(x, wc[j], win);
#pragma unroll
                for (int j = 0; j < IR; j++) acc += to_acc(wv[j]) * xw[j];
                for (int i = lo + IR; i < hi; i++) acc += to_acc(ld_stream1(iv + i, pol)) * gather(x, ld_stream1(a.irreg_cid + i, pol), win);
            };
            if (a.keep_compact && compact) pipe(std::true_type{});
            else pipe(std::false_type{});
        }
        if (g < a.row_block) store_y<T>(a, (long)a.row_long + g, acc);
        return;
    }
    // DMMA variant: irregular tail and store
    if (g >= a.row_block) return;
    if (a.has_irreg[group]) {
        const T *iv = static_cast<const T *>(a.irreg_val);
        const int lo = __ldg(a.irreg_rpt + g), hi = __ldg(a.irreg_rpt + g + 1);
        for (int i = lo; i < hi; i++) acc += to_acc(ld_stream1(iv + i, pol)) * gather(x, ld_stream1(a.irreg_cid + i, pol));
    }
    store_y<T>(a, (long)a.row_long + g, acc);
}

// Split variant for small matrices (latency-bound: fewer rows than the machine has threads).  One warp per
// 8-row block, four lanes per row: lane 4r+q walks tiles q, q+4, ... of row r and every fourth entry of the
// row's irregular tail; the four partial sums are folded with two shuffles.  The dependent chain is three
// memory round trips for rows of up to 32 regular + 8 irregular entries:
//   {blockPtr, irreg_rpt}  ->  {tile values/indices, irregular values/indices}  ->  {x gathers}
template <typename T, bool KEEP>
__device__ __forceinline__ void medium_rows_split(const SpmvArgs &a, long w)
{
    const StreamPol pol = make_stream_policy<KEEP>();
    using A = typename Acc<T>::type;
    const int lane = threadIdx.x & 31;
    if (w >= a.blocknum) return;
    const int b = (int)w;
    const int r = lane >> 2, q = lane & 3;
    const int g = b * 8 + r;
    const T *x = static_cast<const T *>(a.x);
    const T *iv = static_cast<const T *>(a.irreg_val);
    const int bp0 = __ldg(a.blockPtr + b), bp1 = __ldg(a.blockPtr + b + 1);
    int lo = 0, hi = 0;
    if (g < a.row_block) { lo = __ldg(a.irreg_rpt + g); hi = __ldg(a.irreg_rpt + g + 1); }
    const T *pv = static_cast<const T *>(a.reg_val) + bp0 + 4 * r;
    const int *pc = a.reg_cid + bp0 + 4 * r;
    const int nt = (bp1 - bp0) >> 5;
    A acc = 0;
    int k = q, i = lo + q;
    do {
        T v[2][4], w[2];
        int c[2][4], d[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (k + 4 * j < nt) { ld_stream4<KEEP>(pv + 32 * (k + 4 * j), v[j], pol); ld_stream4<KEEP>(pc + 32 * (k + 4 * j), c[j], pol); }
            else {
#pragma unroll
                for (int e = 0; e < 4; e++) { v[j][e] = T(0); c[j][e] = 0; }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const bool ok = i + 4 * j < hi;
            w[j] = ok ? ld_stream1(iv + i + 4 * j, pol) : T(0);
            d[j] = ok ? ld_stream1(a.irreg_cid + i + 4 * j, pol) : 0;
        }
        A xv[2][4], xw[2];
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) xv[j][e] = gather(x, c[j][e]);
#pragma unroll
        for (int j = 0; j < 2; j++) xw[j] = gather(x, d[j]);
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc += to_acc(v[j][e]) * xv[j][e];
#pragma unroll
        for (int j = 0; j < 2; j++) acc += to_acc(w[j]) * xw[j];
        k += 8;
        i += 8;
    } while (__any_sync(0xffffffffu, k < nt || i < hi));
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0 && g < a.row_block) store_y<T>(a, (long)a.row_long + g, acc);
}

// ------------------------------------------------------------------------------------------------
// short rows

// values and column indices of one warp item of a short segment (4 slots per lane in every segment), so that a kernel
// can request the streams of its next item before it multiplies the current one (short-band kernel)
template <typename T> struct ShortRegs {
    T v[4];
    int c[4];
};
static_assert(SINGLES_PER_THREAD == 4 && SHORT_TILES_PER_WARP == 4, "ShortRegs holds 4 slots per lane");

template <typename T>
__device__ __forceinline__ void short_load(const SpmvArgs &a, int seg, long w, ShortRegs<T> &R, const StreamPol &pol)
{
    const int lane = threadIdx.x & 31;
    if (seg == 2) { // singles
        const T *val = static_cast<const T *>(a.short_val) + a.s1;
        const int *cid = a.short_cid + a.s1;
        const long base = w * 32 * SINGLES_PER_THREAD + lane;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long i = base + j * 32;
            const bool ok = i < a.n1;
            R.v[j] = ok ? ld_stream1(val + i, pol) : T(0);
            R.c[j] = ok ? ld_stream1(cid + i, pol) : 0;
        }
        return;
    }
    constexpr int G = sizeof(T) == 8 ? 8 : 32;
    const int sbase = seg == 3 ? a.s13 : (seg == 4 ? a.s34 : a.s22);
    const int nrows = seg == 3 ? a.c13 : (seg == 4 ? a.n34 : a.n2);
    const T *val = static_cast<const T *>(a.short_val) + sbase;
    const int *cid = a.short_cid + sbase;
    const int tile0 = (int)w * SHORT_TILES_PER_WARP;
    const int tiles_avail = seg == 5 ? (int)(((long)nrows + 2 * G - 1) / (2 * G)) * (G >> 3) : (nrows + 7) / 8;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool ok = tile0 + j < tiles_avail;
        const long s = (long)(tile0 + j) * 32 + lane;
        R.v[j] = ok ? ld_stream1(val + s, pol) : T(0);
        R.c[j] = ok ? ld_stream1(cid + s, pol) : 0;
    }
}

template <typename T, bool KEEP>
__device__ __forceinline__ void short_singles(const SpmvArgs &a, long w, const XWin<T> *win = nullptr, const ShortRegs<T> *pre = nullptr)
{
    const StreamPol pol = make_stream_policy<KEEP>();
    const T *x = static_cast<const T *>(a.x);
    const T *val = static_cast<const T *>(a.short_val) + a.s1;
    const int *cid = a.short_cid + a.s1;
    const long base = w * 32 * SINGLES_PER_THREAD + (threadIdx.x & 31);
    T v[SINGLES_PER_THREAD];
    int c[SINGLES_PER_THREAD];
#pragma unroll
    for (int j = 0; j < SINGLES_PER_THREAD; j++) {
        long i = base + j * 32;
        if (pre) { v[j] = pre->v[j]; c[j] = pre->c[j]; }
        else if (i < a.n1) { v[j] = ld_stream1(val + i, pol); c[j] = ld_stream1(cid + i, pol); }
    }
#pragma unroll
    for (int j = 0; j < SINGLES_PER_THREAD; j++) {
        long i = base + j * 32;
        if (i < a.n1) store_y<T>(a, a.y1 + i, to_acc(v[j]) * gather(x, c[j], win));
    }
}

// y index of tile `tile` (8 tile rows), row r, half h inside a 1&3 or 2&2 segment: FP64 interleaves per tile (G = 8 rows),
// FP16 per group of 4 tiles (G = 32); K11 / P10 of SURVEY §8a.  G is a compile-time constant of the value type, so this
// is shifts and adds (the run-time form cost two 64-bit divisions per tile: half of the short kernels' instructions).
template <typename T> __device__ __forceinline__ int paired_y(int tile, int r, int h)
{
    if (sizeof(T) == 8) return tile * 16 + h * 8 + r;
    return (tile >> 2) * 64 + h * 32 + (tile & 3) * 8 + r;
}

// MODE 0: 1&3 tiles   MODE 1: 3/4 rows   MODE 2: 2&2 tiles
// SMMA: the reference's formulation (src/dasp_f64.h:296-483): one DMMA m8n8k4 per 8x4 tile with B masked to the slots
// that belong to the first / second row of a tile row; the useful results sit on the diagonal of C.
template <typename T, int MODE, bool KEEP, bool SMMA>
__device__ __forceinline__ void short_tiles(const SpmvArgs &a, long w, const XWin<T> *win = nullptr, const ShortRegs<T> *pre = nullptr)
{
    const StreamPol pol = make_stream_policy<KEEP>();
    using A = typename Acc<T>::type;
    const int lane = threadIdx.x & 31;
    const T *x = static_cast<const T *>(a.x);
    constexpr int G = sizeof(T) == 8 ? 8 : 32; // == a.G
    const int sbase = MODE == 0 ? a.s13 : (MODE == 1 ? a.s34 : a.s22);
    const int nrows = MODE == 0 ? a.c13 : (MODE == 1 ? a.n34 : a.n2); // rows (pairs for MODE 0)
    const T *val = static_cast<const T *>(a.short_val) + sbase;
    const int *cid = a.short_cid + sbase;
    const int tile0 = (int)w * SHORT_TILES_PER_WARP;
    // 2&2 packs 2G rows per G/8 tiles (16 per tile in FP64, 64 per 4 tiles in FP16)
    const int tiles_avail = MODE == 2 ? (int)(((long)nrows + 2 * G - 1) / (2 * G)) * (G >> 3) : (nrows + 7) / 8;
    A p[SHORT_TILES_PER_WARP];
    T v[SHORT_TILES_PER_WARP];
    int c[SHORT_TILES_PER_WARP];
#pragma unroll
    for (int j = 0; j < SHORT_TILES_PER_WARP; j++) {
        bool ok = tile0 + j < tiles_avail;
        long s = (long)(tile0 + j) * 32 + lane;
        if (pre) { v[j] = pre->v[j]; c[j] = pre->c[j]; }
        else { v[j] = ok ? ld_stream1(val + s, pol) : T(0); c[j] = ok ? ld_stream1(cid + s, pol) : 0; }
    }
    const int r = lane >> 2, q = lane & 3;
    if constexpr (SMMA && sizeof(T) == 8) {
        // lane = slot: A[r][q] = value, B[q][r] = x of the same slot (masked), C[r][r] = dot product of tile row r.
        // C[r][r] is held by lane 4r + (r>>1) in register r&1: that lane stores it, no shuffle.
        const bool holder = q == (r >> 1);
#pragma unroll
        for (int j = 0; j < SHORT_TILES_PER_WARP; j++) {
            const int tile = tile0 + j;
            const double av = (double)to_acc(v[j]);
            const double xg = (double)gather(x, c[j]);
            double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
            if (MODE == 1) {
                dmma884(c0, av, xg);
                const int row = tile * 8 + r;
                if (holder && row < nrows) store_y<T>(a, a.y34 + row, (A)c0[r & 1]);
            } else {
                const bool first = MODE == 0 ? q == 0 : q < 2; // slots of the first row of the tile row
                dmma884(c0, av, first ? xg : 0.0);
                dmma884(c1, av, first ? 0.0 : xg);
                if (MODE == 0) {
                    if (holder && tile * 8 + r < nrows) {
                        store_y<T>(a, a.y13 + paired_y<T>(tile, r, 0), (A)c0[r & 1]);
                        store_y<T>(a, a.y13 + paired_y<T>(tile, r, 1), (A)c1[r & 1]);
                    }
                } else if (holder) {
                    const int y0 = paired_y<T>(tile, r, 0), y1 = paired_y<T>(tile, r, 1);
                    if (y0 < nrows) store_y<T>(a, a.y22 + y0, (A)c0[r & 1]);
                    if (y1 < nrows) store_y<T>(a, a.y22 + y1, (A)c1[r & 1]);
                }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < SHORT_TILES_PER_WARP; j++) p[j] = to_acc(v[j]) * gather(x, c[j], win);
#pragma unroll
        for (int j = 0; j < SHORT_TILES_PER_WARP; j++) {
            const int tile = tile0 + j;
            if (MODE == 1) {
                A s = p[j] + __shfl_xor_sync(0xffffffffu, p[j], 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                int row = tile * 8 + r;
                if (q == 0 && row < nrows) store_y<T>(a, a.y34 + row, s);
            } else if (MODE == 0) {
                A d1 = __shfl_down_sync(0xffffffffu, p[j], 1), d2 = __shfl_down_sync(0xffffffffu, p[j], 2);
                int pair = tile * 8 + r;
                if (pair < nrows) {
                    if (q == 0) store_y<T>(a, a.y13 + paired_y<T>(tile, r, 0), p[j]);
                    if (q == 1) store_y<T>(a, a.y13 + paired_y<T>(tile, r, 1), p[j] + d1 + d2);
                }
            } else {
                A d1 = __shfl_down_sync(0xffffffffu, p[j], 1);
                if ((q & 1) == 0) {
                    int yi = paired_y<T>(tile, r, q >> 1);
                    if (yi < nrows) store_y<T>(a, a.y22 + yi, p[j] + d1);
                }
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ void zero_rows(const SpmvArgs &a, long w)
{
    long i = w * 32 + (threadIdx.x & 31);
    if (i < a.row_zero) store_y<T>(a, a.y0 + i, typename Acc<T>::type(0));
}

template <typename T> __device__ __forceinline__ void lcb_finalize(const SpmvArgs &a, long w); // below, with the LCB kernel

// MED: 0 one lane per row (large matrices), 1 DMMA tiles, 2 four lanes per row (small matrices)
// KEEP: the layout fits in L2, streams stay at normal L2 priority (small matrices iterated back to back)
template <typename T, int MED, int LONGV, bool KEEP, bool SMMA>
__device__ __forceinline__ void run_category(const SpmvArgs &a, int cat, long w, unsigned char *smem)
{
    switch (cat) {
    case 0:
        if (a.lcb_nctas > 0) lcb_finalize<T>(a, w); // the column-blocked kernel ran just before on this stream
        else long_rows<T, LONGV, KEEP>(a, w, smem);
        break;
    case 1:
        if constexpr (MED == 2) medium_rows_split<T, KEEP>(a, w);
        else if constexpr (MED == 3) medium_rows<T, false, KEEP, true>(a, w); // small matrices, register-lean loop
        else if constexpr (MED == 4) medium_rows<T, false, KEEP, false, 1>(a, w); // A/B aid: pipelined loop, one tile per batch, 4 CTAs per SM
        else if constexpr (MED == 5) medium_rows<T, false, KEEP>(a, w);           // A/B aid: pipelined loop compiled for 3 CTAs per SM
        else if constexpr (MED == 6) medium_rows<T, false, KEEP, true, 3>(a, w);  // A/B aid: register-lean loop, 3 tiles per batch
        else if constexpr (MED == 7) medium_rows<T, false, KEEP, true, 2>(a, w);  // A/B aid: register-lean loop, 2 tiles per batch
        else medium_rows<T, MED == 1, KEEP>(a, w);
        break;
    case 2: short_singles<T, KEEP>(a, w); break;
    case 3: short_tiles<T, 0, KEEP, SMMA>(a, w); break;
    case 4: short_tiles<T, 1, KEEP, SMMA>(a, w); break;
    case 5: short_tiles<T, 2, KEEP, SMMA>(a, w); break;
    default: zero_rows<T>(a, w); break;
    }
}

// One warp per work item; the block index selects the category (grid = sum of the per-category CTA counts).
// NT = threads per CTA: 256 for the bandwidth-bound kernels; the small-matrix (KEEP) kernels also exist as 128-thread CTAs
// compiled for 7 CTAs per SM (<= 72 registers, 4144 warp slots) so that every warp of an L2-resident matrix is resident at once: a 121 k-row
// matrix is 3788 warps, the 256-thread / 84-register form holds 3552 and leaves a 60 %-empty second wave.
template <typename T, int MED, int LONGV, bool KEEP, bool SMMA = false, int NT = CTA>
__global__ void __launch_bounds__(NT, KEEP ? (NT == 128 ? 7 : (MED == 5 ? 3 : (MED >= 3 ? 4 : 1))) : MED_MINB) spmv_kernel(const __grid_constant__ SpmvArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn_smem[]; // only the TMA long-row variant asks for any
    const int bid = blockIdx.x, warp = threadIdx.x >> 5;
    if constexpr (KEEP) pdl_launch_dependents();
    int cat = 0, first = 0;
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (bid >= a.e[k]) { cat = k + 1; first = a.e[k]; }
    int local = bid - first;
    if constexpr (!KEEP) {
        if (a.short_map && cat >= 2 && cat <= 5) { // the four short segments interleaved by row band
            const int e = __ldg(a.short_map + (bid - a.e[1]));
            cat = e >> 28;
            local = e & 0x0FFFFFFF;
        }
    }
    // the medium-row path (MED == 0) waits for the predecessor itself, after it has requested its first tiles
    if constexpr (KEEP) { if (!(cat == 1 && (MED == 0 || MED >= 3))) pdl_wait(); }
    run_category<T, MED, LONGV, KEEP, SMMA>(a, cat, (long)local * (NT / 32) + warp, dyn_smem);
}

// Small matrices, medium rows by SM (DASP_SMQ=1; measured and NOT chosen by AUTO, see the end of this comment).  One product of an L2-resident matrix is bound by the
// L2 -> SM sector rate (~6300 B/clk for the chip): next to the 12 bytes per entry of the streams every scattered x gather
// moves its own 32-byte sector.  x itself is small; what an SM needs of it fits in its L1 if the SM works on NEIGHBOURING
// rows.  So the chunks of 8 groups (one 32-row group per warp), in order of the original id of their first row, are dealt
// smq_k consecutive chunks per SM - as many CTAs as are resident per SM at once - and a CTA asks for the next chunk of the
// SM it happens to run on (%smid); the chunks beyond smq_n * smq_k are taken by block index.  A CTA whose SM has no chunk
// left looks for a queue that has (one warp reads 32 counters per round trip).  Exactly one CTA per chunk is launched, so
// every chunk is taken once.  The last CTA to take its chunk zeroes the counters for the next launch, and only then lets
// the dependent launch start.
// Measured (profiles/r02/README.md §3): L1 hit rate 48 -> 64 % (FP16 81 %), L2 sectors per product 1.89 M -> 1.32 M, and yet
// 12.9 vs 9.3 us (FP64) / 10.8 vs 7.2 us (FP16) back to back: the chunk hand-out and the order indirection add three dependent
// round trips to a kernel that is six long, which costs more than the sectors saved.
template <typename T, int MINB, bool LEAN>
__global__ void __launch_bounds__(CTA, MINB) smq_kernel(const __grid_constant__ SpmvArgs a)
{
    __shared__ int s_chunk;
    const int queued = a.smq_n * a.smq_k;
    if ((int)blockIdx.x >= queued) {
        if (threadIdx.x == 0) s_chunk = blockIdx.x;
    } else if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        int q = (int)(smid % (unsigned)a.smq_n), slot = 0;
        if (lane == 0) slot = atomicAdd(a.smq_cnt + q, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        for (int tries = 0; slot >= a.smq_k; tries++) {
            if (tries > a.smq_n) __trap(); // counters left over by an aborted / concurrent launch on this handle: fault, do not spin
            // this SM's chunks are gone: find a queue that still has one, 32 queues per round trip starting after q
            int found = -1;
            for (int base = 1; base < a.smq_n && found < 0; base += 32) {
                const int t = base + lane;
                int qq = q + t;
                if (qq >= a.smq_n) qq -= a.smq_n;
                const bool free_slot = t < a.smq_n && *((volatile int *)a.smq_cnt + qq) < a.smq_k;
                const unsigned m = __ballot_sync(0xffffffffu, free_slot);
                if (m) found = __shfl_sync(0xffffffffu, qq, __ffs(m) - 1);
            }
            if (found < 0) found = q + 1 == a.smq_n ? 0 : q + 1; // (raced: somebody took it in between) try again
            q = found;
            if (lane == 0) slot = atomicAdd(a.smq_cnt + q, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0);
        }
        if (lane == 0) {
            s_chunk = q * a.smq_k + slot;
            if (atomicAdd(a.smq_cnt + SMQ_MAX, 1) == queued - 1) { // every queued CTA of this launch has its chunk
                for (int i = 0; i < a.smq_n; i++) a.smq_cnt[i] = 0;
                a.smq_cnt[SMQ_MAX] = 0;
                __threadfence();
            }
        }
    }
    __syncthreads();
    pdl_launch_dependents();
    medium_rows<T, false, true, LEAN>(a, (long)s_chunk * (CTA / 32) + (threadIdx.x >> 5));
}

// ------------------------------------------------------------------------------------------------
// long rows, column-blocked (LCB): for long rows whose columns are scattered over x every gather of the chunked kernel
// above costs its own 128-byte L1 wavefront and its own 32-byte DRAM sector.  Here the live entries of ALL long rows
// are sorted by (column block, row) at preprocessing (derive.cu); a CTA owns up to LCB_PART consecutive entries of one
// block, stages that block of x (64 KB) in shared memory with TMA bulk copies (cp.async.bulk + mbarrier) while its
// first value / index loads are in flight, and gathers from shared memory.  A lane owns FOUR consecutive entries per step
// (one 256-bit value load + one 128-bit load of packed row<<16|column indices): the load/store unit, not DRAM, bounds this
// kernel, so instructions per entry are what counts.  Rows ascend inside a block: a lane first folds its own entries
// (row changes inside a lane go straight to an atomic add), then while the whole warp stays on one row the lanes
// accumulate privately; a row change costs ONE warp reduction + one atomic add into the per-row accumulator (the atomic
// merge of split rows, cf. longPart_sum src/dasp_f64.h:53-75).  The last CTA to finish turns the accumulators into y (K11
// placement, scatter / axpby forms included) and re-zeroes them.
template <typename A> __device__ __forceinline__ void red_add(A *p, A v) { atomicAdd(p, v); }

__device__ __forceinline__ void ld_lcb4(const double *p, double (&v)[4], const StreamPol &) { ld_stream4d<false>(p, v); }
__device__ __forceinline__ void ld_lcb4(const __half *p, __half (&v)[4], const StreamPol &pol) { ld_stream4<false>(p, v, pol); }

// IDX16 (FP64): the indices are streamed in the 16-bit form (derive.cu: lcb_encode16) - 10 instead of 12 bytes per entry, of a
// kernel that moves 6.3 TB/s.  The rows of a step are rebuilt from the deltas with one warp scan; a warp walks whole chunks
// of 1024 entries (8 steps), each of which restarts from its own row, and reads a chunk flagged wide through lcb_idx.
template <typename T, bool IDX16>
__global__ void __launch_bounds__(CTA, 3) lcb_kernel(const __grid_constant__ SpmvArgs a)
{
    using A = typename Acc<T>::type;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    T *xs = reinterpret_cast<T *>(dyn_smem);
    __shared__ __align__(8) unsigned long long bar;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int blo = 0, bhi = a.lcb_nblk; // last block b with cta_first[b] <= c (empty blocks share their successor's value)
    while (bhi - blo > 1) {
        const int mid = (blo + bhi) >> 1;
        if (__ldg(a.lcb_cta_first + mid) <= c) blo = mid; else bhi = mid;
    }
    const int b = blo;
    const long col0 = (long)b << a.lcb_bw_log2;
    const int cnt = (int)min(1L << a.lcb_bw_log2, (long)a.ncols - col0);
    const T *xg = static_cast<const T *>(a.x) + col0;
    const uint32_t bytes16 = ((uint32_t)cnt * (uint32_t)sizeof(T)) & ~15u;
    const uint32_t bar_addr = smem_u32(&bar);
    if (tid == 0) {
        mbar_init(bar_addr, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        uint64_t keep;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep)); // other CTAs read the same block
        mbar_expect_tx(bar_addr, bytes16);
        for (uint32_t off = 0; off < bytes16; off += 32768u)
            bulk_g2s(smem_u32(xs) + off, reinterpret_cast<const char *>(xg) + off, min(32768u, bytes16 - off), bar_addr, keep);
    }
    for (int i = (int)(bytes16 / sizeof(T)) + tid; i < cnt; i += CTA) xs[i] = xg[i]; // tail that is not a 16-byte multiple

    const int p0 = __ldg(a.lcb_blk_ptr + b), p1 = __ldg(a.lcb_blk_ptr + b + 1); // multiples of 4
    // the block's entries are cut into equal parts (multiples of 1024) for its CTAs
    const int c0 = __ldg(a.lcb_cta_first + b), nparts = __ldg(a.lcb_cta_first + b + 1) - c0;
    const int psize = (((p1 - p0 + nparts - 1) / nparts) + 1023) & ~1023;
    const int beg = min(p0 + (c - c0) * psize, p1), end = min(beg + psize, p1);
    // The part is walked in steps of 128 entries; the warps take chunks of 8 consecutive steps round-robin (step t of the part
    // belongs to warp (t / 8) % WARPS), so the CTA as a whole streams sequentially and every warp samples the whole part
    // (slices of very different row-change density would leave most warps waiting for the slowest one).
    const int nsteps_part = (end - beg + 127) >> 7;
    const T *val = static_cast<const T *>(a.lcb_val);
    const StreamPol pol = make_stream_policy<false>();
    const int acc_stride = (a.row_long + 31) & ~31;
    A *acc = static_cast<A *>(a.lcb_acc) + (size_t)(c & (LCB_COPIES - 1)) * acc_stride; // this CTA's private copy
    auto step_addr = [&](int s) { return ((s >> 3) * (8 * WARPS) + warp * 8 + (s & 7)); }; // s-th step of this warp
    if (step_addr(0) < nsteps_part) {
        T v0[4], v1[4], v2[4];
        int k0[4], k1[4], k2[4]; // IDX16: [0], [1] four packed 16-bit indices, [2] restart row of the chunk, [3] its wide flag
        bool ok0, ok1, ok2;
        // a lane past the end of the part re-reads the indices of the part's last four entries (rows stay ascending)
        // and takes zeros as values (parts are multiples of 1024 entries: does not happen with IDX16)
        auto load = [&](T(&v)[4], int(&k)[4], bool &ok, int s) {
            const int t = step_addr(s);
            const int q = beg + 128 * t + 4 * lane;
            ok = q < end;
            const int qi = ok ? q : end - 4;
            if (t < nsteps_part) {
                if constexpr (IDX16) {
                    asm volatile(DASP_LD_HINT ".v2.u32 {%0,%1}, [%2], %3;" : "=r"(k[0]), "=r"(k[1]) : "l"(a.lcb_idx16 + qi), "l"(pol.desc));
                    const int chunk = (beg >> 10) + (t >> 3);
                    k[2] = __ldg(a.lcb_chunk_row + chunk);
                    k[3] = __ldg(a.lcb_chunk_wide + chunk);
                } else ld_stream4<false>(reinterpret_cast<const int *>(a.lcb_idx) + qi, k, pol);
                if (ok) ld_lcb4(val + q, v, pol);
            }
        };
        A lane_acc = 0;
        int cur = -1; // row the private accumulators belong to
        int prev_row = 0; // IDX16: row of the last entry of the warp's previous step
        auto consume = [&](const T(&v)[4], const int(&k)[4], bool ok, int s) {
            A p[4];
            int r[4];
            if constexpr (IDX16) {
                const int t = step_addr(s);
                if (k[3]) { // wide chunk (a row delta > 7 somewhere in it): the 32-bit indices of this step, read now
                    int w[4];
                    ld_stream4<false>(reinterpret_cast<const int *>(a.lcb_idx) + beg + 128 * t + 4 * lane, w, pol);
#pragma unroll
                    for (int j = 0; j < 4; j++) { r[j] = (int)((unsigned)w[j] >> 16); p[j] = to_acc(v[j]) * to_acc(xs[w[j] & 0xFFFF]); }
                } else {
                    unsigned h[4] = {(unsigned)k[0] & 0xFFFFu, (unsigned)k[0] >> 16, (unsigned)k[1] & 0xFFFFu, (unsigned)k[1] >> 16};
                    const int d0 = h[0] >> 13, d1 = h[1] >> 13, d2 = h[2] >> 13, d3 = h[3] >> 13;
                    int incl = d0 + d1 + d2 + d3; // rows this lane advances by; inclusive scan over the lanes
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int up = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += up;
                    }
                    const int start = ((t & 7) == 0) ? k[2] : prev_row; // a chunk restarts from its own row (its first delta is 0)
                    r[0] = start + incl - (d1 + d2 + d3);
                    r[1] = r[0] + d1; r[2] = r[1] + d2; r[3] = r[2] + d3;
#pragma unroll
                    for (int j = 0; j < 4; j++) p[j] = to_acc(v[j]) * to_acc(xs[h[j] & 0x1FFFu]);
                }
                prev_row = __shfl_sync(0xffffffffu, r[3], 31);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    r[j] = (int)((unsigned)(ok ? k[j] : k[3]) >> 16); // past the end: the row of the part's last entry
                    p[j] = ok ? to_acc(v[j]) * to_acc(xs[k[j] & 0xFFFF]) : A(0);
                }
            }
            if (__all_sync(0xffffffffu, r[0] == cur && r[3] == cur)) { lane_acc += (p[0] + p[1]) + (p[2] + p[3]); return; }
            // Some row other than `cur` appears in this step.  Fold the lane's own entries: `first` = its leading run,
            // `last` = its trailing run; runs in between (rows of fewer than 4 entries) go straight to the accumulators.
            A first = p[0], last = p[0];
            int last_row = r[0];
            bool split = false; // the lane holds more than one row
#pragma unroll
            for (int j = 1; j < 4; j++) {
                if (r[j] == last_row) { last += p[j]; if (!split) first = last; }
                else {
                    if (split) red_add(acc + last_row, last); // a middle run
                    split = true;
                    last_row = r[j];
                    last = p[j];
                }
            }
            // everything that belongs to `cur`: the private accumulators and the leading runs that continue it
            const A t = lane_acc + (r[0] == cur ? first : A(0));
            const A total = warp_sum(t);
            if (cur >= 0 && lane == 0) red_add(acc + cur, total);
            const int new_cur = __shfl_sync(0xffffffffu, last_row, 31);
            // leading run of a row other than cur / new_cur, or of new_cur when the lane continues with another row
            if (r[0] != cur && (split || r[0] != new_cur)) red_add(acc + r[0], first);
            if (split && last_row != new_cur) red_add(acc + last_row, last);
            lane_acc = (last_row == new_cur) ? ((split || r[0] != cur) ? last : A(0)) : A(0);
            cur = new_cur;
        };
        // three steps (3 x 128 entries per warp) in flight: the kernel is bound by DRAM latency x bytes in flight
        load(v0, k0, ok0, 0);
        load(v1, k1, ok1, 1);
        mbar_wait(bar_addr, 0);
        __syncthreads(); // the tail elements written with plain stores
        for (int s = 0;; s += 3) {
            load(v2, k2, ok2, s + 2);
            consume(v0, k0, ok0, s);
            if (step_addr(s + 1) >= nsteps_part) break;
            load(v0, k0, ok0, s + 3);
            consume(v1, k1, ok1, s + 1);
            if (step_addr(s + 2) >= nsteps_part) break;
            load(v1, k1, ok1, s + 4);
            consume(v2, k2, ok2, s + 2);
            if (step_addr(s + 3) >= nsteps_part) break;
        }
        if (cur >= 0) {
            const A total = warp_sum(lane_acc);
            if (lane == 0) red_add(acc + cur, total);
        }
    } else {
        mbar_wait(bar_addr, 0);
        __syncthreads();
    }
    // The accumulators become y in the launch that follows on the stream (lcb_finalize below: the long-row slot of the fused
    // kernel), so this kernel needs neither a grid-wide completion count nor fences.
}

// y[r] = accumulator of long row r (K11 placement, scatter / axpby forms through store_y), accumulator re-zeroed
template <typename T> __device__ __forceinline__ void lcb_finalize(const SpmvArgs &a, long w)
{
    using A = typename Acc<T>::type;
    const long r = w * 32 + (threadIdx.x & 31);
    if (r >= a.row_long) return;
    A *acc = static_cast<A *>(a.lcb_acc);
    const int acc_stride = (a.row_long + 31) & ~31;
    A t = 0;
#pragma unroll
    for (int k = 0; k < LCB_COPIES; k++) { // the private copies of the CTAs, in a fixed order
        t += __ldcg(acc + (size_t)k * acc_stride + r);
        __stcg(acc + (size_t)k * acc_stride + r, A(0));
    }
    store_y<T>(a, r, t);
}

// the same as a launch of its own: used when the column-blocked kernel ran BESIDE the fused kernel (launch_spmv)
template <typename T> __global__ void __launch_bounds__(256) lcb_finalize_kernel(const __grid_constant__ SpmvArgs a)
{
    lcb_finalize<T>(a, (long)blockIdx.x * 8 + (threadIdx.x >> 5));
}

// ------------------------------------------------------------------------------------------------
// short rows by row band (SB): a scattered gather costs the L1 one wavefront per lane (measured: ~0.72 gathers per clock
// and SM, which caps rows of 1-4 entries far below the HBM roofline).  Here the warp items of the four short segments
// are sorted by the band of ORIGINAL row ids they start in (derive.cu); a persistent CTA per SM walks bands, stages the
// band's window of x (192 KB: the band's 16384 columns plus 4096 either side) in shared memory with TMA bulk copies and
// runs the same per-segment code as the fused kernel with the gathers served from shared memory (a column outside the
// window falls back to global memory, so any matrix is handled).
constexpr int SB_THREADS = 1024;
constexpr int SB_WARPS = SB_THREADS / 32;
constexpr int SB_WIN_BYTES = SB_WINDOW_BYTES;

template <typename T>
__global__ void __launch_bounds__(SB_THREADS, 1) sb_kernel(const __grid_constant__ SpmvArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const T *x = static_cast<const T *>(a.x);
    const StreamPol pol = make_stream_policy<false>();
    const uint32_t bar0 = smem_u32(&bar);
    if (tid == 0) {
        mbar_init(bar0, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t keep = 0;
    if (tid == 0) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep)); // neighbouring bands overlap
    int it = 0;
    for (int b = blockIdx.x; b < a.sb_nbands; b += gridDim.x, it++) {
        // window of this band: columns [lo, lo + len), lo a multiple of 8, len * sizeof(T) a multiple of 16
        XWin<T> win;
        win.lo = __ldg(a.sb_lo + b);
        const long room = (long)a.ncols - win.lo;
        const long n = room < a.sb_wcap ? room : a.sb_wcap;
        win.len = n > 0 ? (unsigned)(n & ~(long)(16 / sizeof(T) - 1)) : 0u;
        win.xs = reinterpret_cast<const T *>(dyn_smem);
        if (tid == 0) {
            const uint32_t bytes = win.len * (uint32_t)sizeof(T), dst = smem_u32(dyn_smem);
            mbar_expect_tx(bar0, bytes);
            for (uint32_t off = 0; off < bytes; off += 32768u)
                bulk_g2s(dst + off, reinterpret_cast<const char *>(x + win.lo) + off, min(32768u, bytes - off), bar0, keep);
        }
        const int i0 = __ldg(a.sb_band_ptr + b), i1 = __ldg(a.sb_band_ptr + b + 1);
        // Software pipeline over the warp's items: the value / index streams of item k+1 (and the id of item k+2) are
        // requested before item k is multiplied, so loads are in flight all the time; the first item's streams are
        // requested while the window itself is still arriving.
        auto multiply = [&](int e, const ShortRegs<T> &R) {
            const long w = e & 0x0FFFFFFF;
            switch (e >> 28) {
            case 2: short_singles<T, false>(a, w, &win, &R); break;
            case 3: short_tiles<T, 0, false, false>(a, w, &win, &R); break;
            case 4: short_tiles<T, 1, false, false>(a, w, &win, &R); break;
            default: short_tiles<T, 2, false, false>(a, w, &win, &R); break;
            }
        };
        ShortRegs<T> R0, R1;
        int i = i0 + warp;
        int e0 = i < i1 ? __ldg(a.sb_item + i) : 0, e1 = i + SB_WARPS < i1 ? __ldg(a.sb_item + i + SB_WARPS) : 0;
        if (i < i1) short_load<T>(a, e0 >> 28, e0 & 0x0FFFFFFF, R0, pol);
        mbar_wait(bar0, (unsigned)it & 1u);
        for (; i < i1; i += 2 * SB_WARPS) {
            const bool has1 = i + SB_WARPS < i1;
            if (has1) short_load<T>(a, e1 >> 28, e1 & 0x0FFFFFFF, R1, pol);
            const int e2 = i + 2 * SB_WARPS < i1 ? __ldg(a.sb_item + i + 2 * SB_WARPS) : 0;
            multiply(e0, R0);
            if (!has1) break;
            const bool has2 = i + 2 * SB_WARPS < i1;
            if (has2) short_load<T>(a, e2 >> 28, e2 & 0x0FFFFFFF, R0, pol);
            const int e3 = i + 3 * SB_WARPS < i1 ? __ldg(a.sb_item + i + 3 * SB_WARPS) : 0;
            multiply(e1, R1);
            e0 = e2;
            e1 = e3;
        }
        __syncthreads(); // every warp is done with the window before the next band overwrites it
    }
}

// ------------------------------------------------------------------------------------------------
// medium rows with the x window of the CTA staged in shared memory (MB).  In the lane-per-row mapping every gather
// instruction touches 32 different rows; when their columns do not share cache lines the L1 retires ONE gather per clock
// and SM from an L2-resident x (tools/gather_bench.cu), which is exactly what a small matrix like cop20k_A costs (2.6 M
// gathers / 148 SMs = 9.1 us) and what caps the length-sorted medium rows of the power-law shape.  With the groups walked
// in order of original row id (med_order) the 8 groups of a CTA sit in one band of the matrix; a ninth warp copies the
// band's window of x (64 KB, placed on the median column of the CTA's rows by derive.cu) into shared memory with TMA bulk
// copies while the eight consumer warps already load their tiles, and the gathers are served from shared memory (5 per
// clock and SM); columns outside the window fall back to global memory.
constexpr int MB_THREADS = CTA + 32;
constexpr int MB_WIN_BYTES = MB_WINDOW_BYTES;

template <typename T, bool KEEP>
__global__ void __launch_bounds__(MB_THREADS, KEEP ? 2 : 3) mb_kernel(const __grid_constant__ SpmvArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) unsigned long long bar;
    if constexpr (KEEP) pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar0 = smem_u32(&bar);
    XWin<T> win;
    win.lo = __ldg(a.mb_lo + blockIdx.x);
    const long room = (long)a.ncols - win.lo;
    const long n = room < a.mb_wcap ? room : a.mb_wcap;
    win.len = n > 0 ? (unsigned)(n & ~(long)(16 / sizeof(T) - 1)) : 0u;
    win.xs = reinterpret_cast<const T *>(dyn_smem);
    win.bar = bar0;
    if (tid == CTA) { // producer: first lane of the ninth warp
        mbar_init(bar0, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == WARPS) {
        if (tid == CTA) {
            if constexpr (KEEP) pdl_wait(); // x may be the previous product's y
            uint64_t keep;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
            const uint32_t bytes = win.len * (uint32_t)sizeof(T), dst = smem_u32(dyn_smem);
            mbar_expect_tx(bar0, bytes);
            const T *x = static_cast<const T *>(a.x);
            for (uint32_t off = 0; off < bytes; off += 32768u)
                bulk_g2s(dst + off, reinterpret_cast<const char *>(x + win.lo) + off, min(32768u, bytes - off), bar0, keep);
        }
        return;
    }
    medium_rows<T, false, KEEP>(a, (long)blockIdx.x * WARPS + warp, &win);
}

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- power-iteration helpers ---------------------------------------------------------------------
constexpr int RED_CTAS = 592; // 4 per SM

__global__ void __launch_bounds__(256) sumsq_partial(const double *__restrict__ v, long n, double *__restrict__ part)
{
    double s = 0.0;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += (long)gridDim.x * 256L) { double t = v[i]; s += t * t; }
    s = warp_sum(s);
    __shared__ double w[8];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += w[k];
        part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) sumsq_final(const double *__restrict__ part, int n, double *__restrict__ out)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += part[i];
    s = warp_sum(s);
    __shared__ double w[8];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += w[k];
        *out = t;
    }
}

__global__ void __launch_bounds__(256) scale_rsqrt(double *__restrict__ v, long n, const double *__restrict__ norm2)
{
    const double f = 1.0 / sqrt(*norm2);
    for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += (long)gridDim.x * 256L) v[i] *= f;
}

} // namespace

namespace {
struct UnpermArgs {
    const void *y_perm;
    const int *inv;
    void *dest[8];
    int n_dest;
    long m, row_offset;
    const double *rs_ptr;
};

// original row i (coalesced over i) <- y_perm[inv[i]] (scattered local read), scaled, stored to every destination
template <typename T>
__global__ void __launch_bounds__(256) unpermute_kernel(const __grid_constant__ UnpermArgs a)
{
    using A = typename Acc<T>::type;
    const T *yp = static_cast<const T *>(a.y_perm);
    const A f = a.rs_ptr ? (A)rsqrt(__ldg(a.rs_ptr)) : A(1);
    for (long i = blockIdx.x * 256L + threadIdx.x; i < a.m; i += (long)gridDim.x * 256L) {
        const A v = to_acc(yp[__ldg(a.inv + i)]) * f;
#pragma unroll 1
        for (int p = 0; p < a.n_dest; p++) from_acc(static_cast<T *>(a.dest[p]) + a.row_offset + i, v);
    }
}
} // namespace

int unpermute_to(dasp_handle *h, const void *d_y_perm, const ScatterTo &dst, void *first, cudaStream_t st)
{
    Layout &L = h->L;
    const int m = L.s.m;
    if (m == 0) return DASP_OK;
    UnpermArgs a{};
    a.y_perm = d_y_perm; a.inv = L.inv_order; a.m = m; a.row_offset = (long)dst.row_offset; a.rs_ptr = dst.norm2;
    a.dest[0] = first;
    for (int p = 0; p < dst.n_extra; p++) a.dest[p + 1] = dst.extra[p];
    a.n_dest = dst.n_extra + 1;
    const int grid = min(cdiv(m, 256), (h->sm_count > 0 ? h->sm_count : 148) * 8);
    if (h->dtype == DASP_F16) unpermute_kernel<__half><<<grid, 256, 0, st>>>(a);
    else unpermute_kernel<double><<<grid, 256, 0, st>>>(a);
    DASP_CUDA(cudaGetLastError());
    return DASP_OK;
}

static bool lcb_selected(const dasp_handle *h)
{
    return h->L.lcb_nctas > 0 && (h->category_mask & 1) &&
           (h->var_long == DASP_VARIANT_BLOCKED || (h->var_long == DASP_VARIANT_AUTO && h->lcb_auto));
}

static bool mb_selected(const dasp_handle *h)
{
    return h->L.mb_lo != nullptr && h->L.s.blocknum > 0 &&
           (h->var_medium == DASP_VARIANT_BANDED || (h->var_medium == DASP_VARIANT_AUTO && h->L.mb_auto));
}

static bool sb_selected(const dasp_handle *h)
{
    return h->L.sb_nbands > 0 && h->L.sb_nitems > 0 &&
           (h->var_short == DASP_VARIANT_BANDED || (h->var_short == DASP_VARIANT_AUTO && h->L.sb_auto));
}

// small (L2-resident) matrix whose medium rows go through the SM-affine queue kernel (same conditions as launch_spmv)
static bool smq_selected(const dasp_handle *h)
{
    static const int use_smq = getenv("DASP_SMQ") ? atoi(getenv("DASP_SMQ")) : 0; // measured slower (profiles/r02/README.md §3): off unless DASP_SMQ=1
    static const int keep_shape = getenv("DASP_KEEP_CTA") ? atoi(getenv("DASP_KEEP_CTA")) : 256;
    const dasp_stats_t &s = h->L.s;
    const bool small = s.data_X <= ((int64_t)48 << 20);
    const bool plain = (h->var_medium == DASP_VARIANT_AUTO || h->var_medium == DASP_VARIANT_CUDA_CORE) &&
                       h->var_long != DASP_VARIANT_MMA && h->var_long != DASP_VARIANT_TMA &&
                       !(h->dtype != DASP_F16 && h->var_short == DASP_VARIANT_MMA);
    return use_smq && keep_shape != 128 && small && plain && (h->category_mask & 2) && s.blocknum > 0 && h->L.smq_cnt != nullptr;
}

int launches_per_spmv(const dasp_handle *h)
{
    // column-blocked long rows, band-staged short / medium rows and the SM-affine medium rows of small matrices are launches of
    // their own, followed by the fused kernel for whatever is left (it also turns the long-row accumulators into y)
    const dasp_stats_t &s = h->L.s;
    const int cm = h->category_mask;
    const bool sb = (cm & 4) && sb_selected(h), mb = (cm & 2) && mb_selected(h), smq = !mb && smq_selected(h);
    const bool short_rows = (long)s.short_row_1 + s.common_13 + s.short_row_34 + s.short_row_2 > 0;
    const bool fused = ((cm & 1) && s.row_long > 0) || ((cm & 2) && s.blocknum > 0 && !mb && !smq) || ((cm & 4) && short_rows && !sb) ||
                       ((cm & 8) && s.row_zero > 0);
    return (fused ? 1 : 0) + (lcb_selected(h) ? 1 : 0) + (sb ? 1 : 0) + (mb ? 1 : 0) + (smq ? 1 : 0);
}

int sumsq(const double *d_v, int64_t count, double *d_out, cudaStream_t st)
{
    // per-call, stream-ordered scratch for the block partials: concurrent calls on different streams do not share it
    double *scratch = nullptr;
    DASP_CUDA(cudaMallocAsync((void **)&scratch, sizeof(double) * RED_CTAS, st));
    sumsq_partial<<<RED_CTAS, 256, 0, st>>>(d_v, (long)count, scratch);
    sumsq_final<<<1, 256, 0, st>>>(scratch, RED_CTAS, d_out);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) { set_error("dasp_sumsq: %s", cudaGetErrorString(e)); return DASP_ERR_CUDA; }
    return DASP_OK;
}

namespace {
struct CopyToArgs {
    const double *src;
    void *dest[8];
    int n_dest;
    long n, offset;
    const double *norm2;
    int vec_ok;
};
__global__ void __launch_bounds__(256) scale_copy_kernel(const __grid_constant__ CopyToArgs a)
{
    const double f = a.norm2 ? 1.0 / sqrt(*a.norm2) : 1.0;
    // two doubles per thread when everything is 16-byte aligned (128-bit multicast / peer stores)
    const bool vec = a.vec_ok && ((a.offset | a.n) & 1) == 0 && ((uintptr_t)a.src & 15) == 0;
    if (vec) {
        const long n2 = a.n >> 1;
        for (long i = blockIdx.x * 256L + threadIdx.x; i < n2; i += (long)gridDim.x * 256L) {
            double2 v = reinterpret_cast<const double2 *>(a.src)[i];
            v.x *= f; v.y *= f;
#pragma unroll 1
            for (int p = 0; p < a.n_dest; p++) reinterpret_cast<double2 *>(static_cast<double *>(a.dest[p]) + a.offset)[i] = v;
        }
    } else {
        for (long i = blockIdx.x * 256L + threadIdx.x; i < a.n; i += (long)gridDim.x * 256L) {
            const double v = a.src[i] * f;
#pragma unroll 1
            for (int p = 0; p < a.n_dest; p++) static_cast<double *>(a.dest[p])[a.offset + i] = v;
        }
    }
}
} // namespace

int scale_copy_to(const double *d_v, int64_t count, void *const *dests, int n_dests, int64_t offset, const double *d_norm2,
                  cudaStream_t st)
{
    if (count <= 0) return DASP_OK;
    CopyToArgs a{};
    a.src = d_v; a.n = (long)count; a.offset = (long)offset; a.norm2 = d_norm2; a.n_dest = n_dests;
    a.vec_ok = 1;
    for (int p = 0; p < n_dests; p++) {
        a.dest[p] = dests[p];
        if ((uintptr_t)dests[p] & 15) a.vec_ok = 0; // 128-bit stores need 16-byte aligned destinations
    }
    scale_copy_kernel<<<RED_CTAS * 2, 256, 0, st>>>(a);
    DASP_CUDA(cudaGetLastError());
    return DASP_OK;
}

int scale_by_rsqrt(double *d_v, int64_t count, const double *d_norm2, cudaStream_t st)
{
    if (count > 0) scale_rsqrt<<<RED_CTAS * 2, 256, 0, st>>>(d_v, (long)count, d_norm2);
    DASP_CUDA(cudaGetLastError());
    return DASP_OK;
}

int launch_spmv(dasp_handle *h, const void *d_x, void *d_y, const int *scatter, cudaStream_t st, const double *alpha_beta,
                const ScatterTo *multi, bool out_f32)
{
    const Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    const bool f16 = h->dtype == DASP_F16;
    SpmvArgs a{};
    a.x = d_x; a.y = d_y; a.scatter = scatter;
    a.axpby = (alpha_beta || multi) ? 1 : 0;
    a.out_f32 = (out_f32 && h->dtype == DASP_F16) ? 1 : 0;
    a.alpha = alpha_beta ? alpha_beta[0] : 1.0;
    a.beta = alpha_beta ? alpha_beta[1] : 0.0;
    if (multi) {
        a.n_extra = multi->n_extra;
        for (int p = 0; p < multi->n_extra; p++) a.y_extra[p] = multi->extra[p];
        a.row_offset = (long)multi->row_offset;
        a.rs_ptr = multi->norm2;
    }
    a.long_val = L.long_val; a.long_cid = L.k_long_cid; a.long_rpt_new = L.long_rpt_new;
    a.unit_row = L.long_unit_row; a.unit_chunk = L.long_unit_chunk; a.unit_first = L.long_unit_first; a.partial = L.long_partial; a.done = L.long_done;
    a.n_units = L.n_long_units; a.longw = f16 ? 256 : 64; a.unit_warps = L.long_unit_warps;
    a.long_cbase = L.long_cbase; a.long_cdelta = L.long_cdelta; a.long_wide = h->index_compression ? L.long_wide : nullptr;
    a.reg_val = L.reg_val; a.reg_cid = L.k_reg_cid; a.blockPtr = L.blockPtr; a.irreg_rpt = L.irreg_rpt;
    a.irreg_val = L.irreg_val; a.irreg_cid = L.k_irreg_cid; a.has_irreg = L.med_has_irreg;
    a.reg_cbase = L.reg_cbase; a.reg_cdelta = L.reg_cdelta; a.blk_wide = h->index_compression ? L.blk_wide : nullptr;
    a.blk_live = L.blk_live;
    a.row_long = s.row_long; a.row_block = s.row_block; a.blocknum = s.blocknum;
    a.short_val = L.short_val; a.short_cid = L.k_short_cid;
    a.n1 = s.short_row_1; a.c13 = s.common_13; a.n34 = s.short_row_34; a.n2 = s.short_row_2;
    const int f13 = s.fill0_nnz_short13, f34 = s.fill0_nnz_short34, f22 = s.fill0_nnz_short22;
    a.s1 = f16 ? f13 + f34 + f22 : 0;
    a.s13 = f16 ? 0 : s.short_row_1;
    a.s34 = a.s13 + f13;
    a.s22 = a.s34 + f34;
    const int ybase = s.row_long + s.row_block;
    a.y13 = ybase + (f16 ? 0 : s.short_row_1);
    a.y34 = a.y13 + 2 * s.common_13;
    a.y22 = a.y34 + s.short_row_34;
    a.y1 = f16 ? a.y22 + s.short_row_2 : ybase;
    a.y0 = s.m - s.row_zero;
    a.G = f16 ? 32 : 8;
    a.row_zero = s.row_zero;

    const int tiles13 = cdiv(s.common_13, 8), tiles34 = cdiv(s.short_row_34, 8);
    const int tiles22 = cdiv(s.short_row_2, 2 * a.G) * (a.G / 8);
    const int cm = h->category_mask;
    // scattered long rows: the column-blocked kernel (its own launch: 64 KB of shared memory per CTA would take the L1
    // away from the x gathers of the other categories); needs x on a 16-byte boundary for the TMA copies
    const bool use_lcb = lcb_selected(h) && ((uintptr_t)d_x & 15) == 0;
    if (use_lcb) {
        a.lcb_val = L.lcb_val; a.lcb_idx = L.lcb_idx; a.lcb_blk_ptr = L.lcb_blk_ptr;
        a.lcb_cta_first = L.lcb_cta_first; a.lcb_acc = L.lcb_acc; a.lcb_done = L.lcb_done;
        a.lcb_bw_log2 = L.lcb_bw_log2; a.lcb_nblk = L.lcb_nblk; a.lcb_nctas = L.lcb_nctas; a.ncols = L.x_len;
        if (!h->lcb_attr_set) {
            DASP_CUDA(cudaFuncSetAttribute(lcb_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            DASP_CUDA(cudaFuncSetAttribute(lcb_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            DASP_CUDA(cudaFuncSetAttribute(lcb_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            h->lcb_attr_set = 1;
        }
    }
    int on_long = cm & 1;
    const int on_zero = (cm >> 3) & 1;
    int on_med = (cm >> 1) & 1, on_short = (cm >> 2) & 1;
    // The column-blocked kernel is bound by DRAM and the shared-memory gathers, the medium / short rows of the same matrix by
    // the L1 miss path (one scattered gather per clock and SM): different resources, so the two launches run BESIDE each
    // other - the column-blocked kernel on a side stream forked from the caller's, limited to `overlap` CTAs per SM by
    // its shared-memory request so that CTAs of the fused kernel fit next to them, the long-row accumulators turned into y
    // by a small launch after the join.  Not with the short-band kernel (192 KB of shared memory per SM).
    // Measured on C3 (profiles/r02/README.md section 4): 0.586 ms serial, 0.576 / 0.680 / 0.980 ms with 3 / 2 / 1 CTAs per SM left to
    // the column-blocked kernel - both launches want the register file, and lcb_kernel loses more from fewer bytes in flight than
    // the overlap gains.  Off unless DASP_LCB_OVERLAP is set.
    static const int overlap_env = getenv("DASP_LCB_OVERLAP") ? atoi(getenv("DASP_LCB_OVERLAP")) : 0;
    const bool sb_would_run = on_short && sb_selected(h) && ((uintptr_t)d_x & 15) == 0;
    const int overlap = (use_lcb && on_long && !sb_would_run && (on_med || on_short)) ? overlap_env : 0;
    bool joined_later = false;
    if (use_lcb) {
        cudaStream_t lcb_stream = st;
        size_t lcb_smem = LCB_BYTES;
        if (overlap > 0) {
            if (!h->side_stream) {
                DASP_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
                DASP_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
                DASP_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
            }
            DASP_CUDA(cudaEventRecord(h->ev_fork, st));
            DASP_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
            lcb_stream = h->side_stream;
            if (overlap == 2) lcb_smem = 100 * 1024; // two CTAs per SM instead of three
            else if (overlap == 1) lcb_smem = 200 * 1024;
            joined_later = true;
        }
        a.lcb_idx16 = L.lcb_idx16; a.lcb_chunk_row = L.lcb_chunk_row; a.lcb_chunk_wide = L.lcb_chunk_wide; // built only with DASP_LCB_IDX16=1
        if (f16) lcb_kernel<__half, false><<<L.lcb_nctas, CTA, lcb_smem, lcb_stream>>>(a);
        else if (L.lcb_idx16) lcb_kernel<double, true><<<L.lcb_nctas, CTA, lcb_smem, lcb_stream>>>(a);
        else lcb_kernel<double, false><<<L.lcb_nctas, CTA, lcb_smem, lcb_stream>>>(a);
        DASP_CUDA(cudaGetLastError());
        if (joined_later) {
            DASP_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
            on_long = 0; // the fused kernel does not touch the long rows; lcb_finalize_kernel does after the join
        }
    }
    auto finish_overlap = [&]() -> int {
        if (!joined_later) return DASP_OK;
        DASP_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
        const int nb = cdiv(s.row_long, 256);
        if (f16) lcb_finalize_kernel<__half><<<nb, 256, 0, st>>>(a);
        else lcb_finalize_kernel<double><<<nb, 256, 0, st>>>(a);
        DASP_CUDA(cudaGetLastError());
        return DASP_OK;
    };
    // short rows by row band with x staged in shared memory (its own launch, 192 KB of shared memory per SM)
    const bool use_sb = on_short && sb_selected(h) && ((uintptr_t)d_x & 15) == 0;
    if (use_sb) {
        a.sb_item = L.sb_item; a.sb_band_ptr = L.sb_band_ptr; a.sb_lo = L.sb_lo; a.sb_nbands = L.sb_nbands;
        a.sb_wcap = SB_WIN_BYTES / (f16 ? 2 : 8); a.ncols = L.x_len;
        if (!h->sb_attr_set) {
            DASP_CUDA(cudaFuncSetAttribute(sb_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_WIN_BYTES));
            DASP_CUDA(cudaFuncSetAttribute(sb_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_WIN_BYTES));
            h->sb_attr_set = 1;
        }
        const int sb_grid = min(L.sb_nbands, h->sm_count > 0 ? h->sm_count : 148);
        if (f16) sb_kernel<__half><<<sb_grid, SB_THREADS, SB_WIN_BYTES, st>>>(a);
        else sb_kernel<double><<<sb_grid, SB_THREADS, SB_WIN_BYTES, st>>>(a);
        DASP_CUDA(cudaGetLastError());
        on_short = 0; // the fused kernel skips the four short segments
    }
    // medium variant: AUTO = one lane per row.  The 4-lanes-per-row split (the analogue of the reference's
    // rowloop=1 geometry for small matrices, src/dasp_f64.h:533-536) and the DMMA tiles are kept as measured
    // alternatives: both lose on B200 (profiles/r01/variants.md).
    const bool small = s.data_X <= ((int64_t)48 << 20); // B200 L2: 126 MB over two dies
    int med = 0;
    if (h->var_medium == DASP_VARIANT_MMA) med = 1; // FP64: DMMA m8n8k4; FP16: HMMA m16n8k16
    else if (h->var_medium == DASP_VARIANT_SPLIT) med = 2;
    const bool mma_long = h->var_long == DASP_VARIANT_MMA && !use_lcb;
    const bool tma_long = h->var_long == DASP_VARIANT_TMA && !use_lcb;
    const bool mma_short = !f16 && h->var_short == DASP_VARIANT_MMA;
    const bool keep_all = small && med != 1 && !mma_long && !tma_long && !mma_short;
    // medium rows with the CTA's window of x in shared memory (its own launch: 64 KB of shared memory per CTA)
    const bool use_mb = on_med && med == 0 && mb_selected(h) && ((uintptr_t)d_x & 15) == 0;
    if (use_mb) {
        a.mb_lo = L.mb_lo; a.mb_wcap = MB_WIN_BYTES / (f16 ? 2 : 8); a.ncols = L.x_len;
        a.med_order = L.med_order; // may be null: identity
        if (!h->mb_attr_set) {
            DASP_CUDA(cudaFuncSetAttribute(mb_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_WIN_BYTES));
            DASP_CUDA(cudaFuncSetAttribute(mb_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_WIN_BYTES));
            DASP_CUDA(cudaFuncSetAttribute(mb_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_WIN_BYTES));
            DASP_CUDA(cudaFuncSetAttribute(mb_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_WIN_BYTES));
            h->mb_attr_set = 1;
        }
        const int mb_grid = cdiv(s.blocknum / 4, WARPS);
        if (keep_all) { // small matrices: programmatic dependent launch, as the fused KEEP kernels
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(mb_grid); cfg.blockDim = dim3(MB_THREADS); cfg.dynamicSmemBytes = MB_WIN_BYTES; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (f16) DASP_CUDA(cudaLaunchKernelEx(&cfg, mb_kernel<__half, true>, a));
            else DASP_CUDA(cudaLaunchKernelEx(&cfg, mb_kernel<double, true>, a));
        } else {
            if (f16) mb_kernel<__half, false><<<mb_grid, MB_THREADS, MB_WIN_BYTES, st>>>(a);
            else mb_kernel<double, false><<<mb_grid, MB_THREADS, MB_WIN_BYTES, st>>>(a);
        }
        DASP_CUDA(cudaGetLastError());
        a.med_order = nullptr;
        on_med = 0; // the fused kernel skips the medium rows
    }
    a.items[0] = on_long * (use_lcb ? (long)cdiv(s.row_long, 32) : (long)L.n_long_units); // LCB: one lane per long row turns its accumulator into y
    a.items[1] = on_med * (long)(med == 2 ? s.blocknum : s.blocknum / 4);
    a.items[2] = on_short * (long)cdiv(s.short_row_1, 32 * SINGLES_PER_THREAD);
    a.items[3] = on_short * (long)cdiv(tiles13, SHORT_TILES_PER_WARP);
    a.items[4] = on_short * (long)cdiv(tiles34, SHORT_TILES_PER_WARP);
    a.items[5] = on_short * (long)cdiv(tiles22, SHORT_TILES_PER_WARP);
    a.items[6] = on_zero * (long)cdiv(s.row_zero, 32);
    // keep the streams at normal L2 priority only when the whole working set is well below the L2 capacity
    const bool keep = small && med != 1 && !mma_long && !tma_long && !mma_short;
    // measured (profiles/r02): the 128-thread / 7-CTA form is SLOWER on C1 / C2 (17.9 vs 9.3 us, 12.2 vs 7.2 us): kept as an A/B aid
    static const int keep_shape = getenv("DASP_KEEP_CTA") ? atoi(getenv("DASP_KEEP_CTA")) : 0; // 0: the measured choice per dtype
    const bool narrow = keep && keep_shape == 128;
    static const int keep_compact_env = getenv("DASP_KEEP_COMPACT") ? atoi(getenv("DASP_KEEP_COMPACT")) : -1; // A/B aid
    a.keep_compact = keep_compact_env >= 0 ? keep_compact_env : 0;
    // Small matrices go through the register-lean medium loop (64 registers, 4 CTAs per SM: the whole matrix is one wave) and
    // walk their 32-row groups in locality order (below), so that the CTAs resident together gather from one neighbourhood of
    // x and the L1 serves most gathers.  Measured on the cop20k_A stand-in (profiles/r02/README.md section 3), back to back:
    //   FP64  pipelined loop 9.0 us | + order 9.2 | lean 11.4 | lean + order 7.4 (256-thread CTAs; 3 tiles per batch: 7.36, 4: 7.43, 2: 7.53)
    //   FP16  pipelined loop 7.0 us | + order 7.1 | lean  6.4 | lean + order 4.65 (4 tiles per batch; 3: 5.0, 2: 5.4)
    // DASP_KEEP_LEAN = 0 (pipelined loop) / 1, 6, 7 (lean, 4 / 3 / 2 tiles per batch) / 2, 3 (A/B shapes) and DASP_KEEP_ORDER = 0 / 1 override.
    static const int keep_lean_env = getenv("DASP_KEEP_LEAN") ? atoi(getenv("DASP_KEEP_LEAN")) : -1;
    // (FP64 matrices whose layout order is already the local one - stencils: no med_order - keep the pipelined loop: 3.3 vs 3.5 us
    // on a 40^3 stencil; FP16 prefers the lean loop there too, 3.4 vs 3.5 us)
    const bool keep_lean = keep_lean_env >= 0 ? keep_lean_env != 0 : (f16 || L.med_order != nullptr);
    const int lean_shape = keep_lean_env > 0 ? keep_lean_env : 1; // four tiles per batch (FP64 in 224-thread CTAs: 6.93 vs 7.07 us with three)
    // FP64 lean loop in 224-thread CTAs (7 warps; 4 CTAs per SM leave 72 registers per thread instead of 64): 7.07 vs 7.19 us
    // on the C1 stand-in; DASP_KEEP_CTA=256 is the A/B switch
    const bool nt224 = keep && (keep_shape == 224 || keep_shape == 0) && !f16 && keep_lean && (lean_shape == 6 || lean_shape == 1) && med == 0 && !mma_long && !tma_long && !mma_short;
    const int nw = narrow ? 4 : (nt224 ? 7 : WARPS);
    // small matrices: medium rows handed out by SM (smq_kernel, its own launch)
    static const int use_smq = getenv("DASP_SMQ") ? atoi(getenv("DASP_SMQ")) : 0; // measured slower (profiles/r02/README.md §3): off unless DASP_SMQ=1
    if (keep && use_smq && !narrow && med == 0 && a.items[1] > 0 && L.smq_cnt) {
        static const int lean = getenv("DASP_SMQ_LEAN") ? atoi(getenv("DASP_SMQ_LEAN")) : 0; // A/B aid: 1 = register-lean loop, 4 CTAs per SM
        const int resident = lean ? 4 : KEEP_MINB; // CTAs per SM the kernel is compiled for
        const int chunks = cdiv(a.items[1], WARPS);
        a.smq_n = min(h->sm_count > 0 ? h->sm_count : 148, SMQ_MAX);
        a.smq_k = min(chunks / a.smq_n, resident); // chunks dealt by SM; the rest goes by block index
        a.smq_cnt = L.smq_cnt;
        a.med_order = L.med_order; // may be null: the layout order is already the local one
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(chunks); cfg.blockDim = dim3(CTA); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (!h->smq_attr_set) {
            cudaFuncSetAttribute(smq_kernel<double, KEEP_MINB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
            cudaFuncSetAttribute(smq_kernel<__half, KEEP_MINB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
            cudaFuncSetAttribute(smq_kernel<double, 4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
            cudaFuncSetAttribute(smq_kernel<__half, 4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
            h->smq_attr_set = 1;
        }
        if (lean) {
            if (f16) DASP_CUDA(cudaLaunchKernelEx(&cfg, smq_kernel<__half, 4, true>, a));
            else DASP_CUDA(cudaLaunchKernelEx(&cfg, smq_kernel<double, 4, true>, a));
        } else {
            if (f16) DASP_CUDA(cudaLaunchKernelEx(&cfg, smq_kernel<__half, KEEP_MINB, false>, a));
            else DASP_CUDA(cudaLaunchKernelEx(&cfg, smq_kernel<double, KEEP_MINB, false>, a));
        }
        a.med_order = nullptr;
        a.items[1] = 0; // the fused kernel skips the medium rows
    }
    long total_items = 0;
    int acc_ctas = 0;
    for (int k = 0; k < 7; k++) {
        acc_ctas += cdiv(a.items[k], nw);
        a.e[k] = acc_ctas;
        total_items += a.items[k];
    }
    if (total_items == 0) return finish_overlap();
    const int grid = a.e[6];
    // locality-ordered work lists: large matrices, everything on, the CTA counts the lists were built for
    static const int use_order = getenv("DASP_NO_LOCALITY_ORDER") ? 0 : 1; // A/B aid
    if (!keep && use_order) {
        if (med == 0 && on_med) a.med_order = L.med_order;
        const int c2 = a.e[2] - a.e[1], c3 = a.e[3] - a.e[2], c4 = a.e[4] - a.e[3], c5 = a.e[5] - a.e[4];
        if (on_short && !mma_short && L.short_map_n > 0 && c2 == L.short_ctas[0] && c3 == L.short_ctas[1] && c4 == L.short_ctas[2] &&
            c5 == L.short_ctas[3])
            a.short_map = L.short_map;
    }
    // small matrices walk their medium groups in locality order too when they run the lean loop (see keep_lean above)
    static const int keep_order_env = getenv("DASP_KEEP_ORDER") ? atoi(getenv("DASP_KEEP_ORDER")) : -1;
    const bool keep_order = keep_order_env >= 0 ? keep_order_env != 0 : (keep_lean && (lean_shape == 1 || lean_shape >= 6));
    if (keep && keep_order && med == 0 && on_med && a.items[1] > 0) a.med_order = L.med_order;

    // the kernels that use no shared memory ask for the whole unified array as L1 (x gathers live there), as the
    // reference does (src/dasp_f64.h:1280-1283)
#define DASP_LAUNCH(T, MED, LV, KEEP)                                                                              \
    do {                                                                                                           \
        const void *fn = (const void *)spmv_kernel<T, MED, LV, KEEP>;                                              \
        if (h->carved_kernel != fn && LV != 2) { /* once per handle (= per device) and kernel */                   \
            cudaFuncSetAttribute(spmv_kernel<T, MED, LV, KEEP>, cudaFuncAttributePreferredSharedMemoryCarveout, 0); \
            h->carved_kernel = fn;                                                                                 \
        }                                                                                                          \
        if (KEEP) { /* small matrices: programmatic dependent launch hides the launch gap between products */     \
            cudaLaunchConfig_t cfg = {};                                                                            \
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(narrow ? 128 : CTA); cfg.dynamicSmemBytes = 0; cfg.stream = st; \
            cudaLaunchAttribute at[1];                                                                              \
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                          \
            at[0].val.programmaticStreamSerializationAllowed = 1;                                                   \
            cfg.attrs = at; cfg.numAttrs = 1;                                                                       \
            if (narrow) {                                                                                           \
                if (h->carved_narrow != (const void *)spmv_kernel<T, MED, LV, KEEP, false, 128>) {                  \
                    cudaFuncSetAttribute(spmv_kernel<T, MED, LV, KEEP, false, 128>, cudaFuncAttributePreferredSharedMemoryCarveout, 0); \
                    h->carved_narrow = (const void *)spmv_kernel<T, MED, LV, KEEP, false, 128>;                     \
                }                                                                                                   \
                DASP_CUDA(cudaLaunchKernelEx(&cfg, spmv_kernel<T, MED, LV, KEEP, false, 128>, a));                  \
            } else                                                                                                  \
                DASP_CUDA(cudaLaunchKernelEx(&cfg, spmv_kernel<T, MED, LV, KEEP>, a));                              \
        } else                                                                                                     \
            spmv_kernel<T, MED, LV, KEEP><<<grid, CTA, LV == 2 ? TMA_SMEM_BYTES : 0, st>>>(a);                     \
    } while (0)
    if (nt224) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(224); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (lean_shape == 1) { // A/B aid: four tiles per batch
            if (h->carved_narrow != (const void *)spmv_kernel<double, 3, 0, true, false, 224>) {
                cudaFuncSetAttribute(spmv_kernel<double, 3, 0, true, false, 224>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
                h->carved_narrow = (const void *)spmv_kernel<double, 3, 0, true, false, 224>;
            }
            DASP_CUDA(cudaLaunchKernelEx(&cfg, spmv_kernel<double, 3, 0, true, false, 224>, a));
        } else {
            if (h->carved_narrow != (const void *)spmv_kernel<double, 6, 0, true, false, 224>) {
                cudaFuncSetAttribute(spmv_kernel<double, 6, 0, true, false, 224>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
                h->carved_narrow = (const void *)spmv_kernel<double, 6, 0, true, false, 224>;
            }
            DASP_CUDA(cudaLaunchKernelEx(&cfg, spmv_kernel<double, 6, 0, true, false, 224>, a));
        }
    } else if (f16) {
        if (tma_long) DASP_LAUNCH(__half, 0, 2, false);
        else if (mma_long) { if (med == 1) DASP_LAUNCH(__half, 1, 1, false); else DASP_LAUNCH(__half, 0, 1, false); }
        else if (med == 1) DASP_LAUNCH(__half, 1, 0, false);
        else if (med == 2) { if (keep) DASP_LAUNCH(__half, 2, 0, true); else DASP_LAUNCH(__half, 2, 0, false); }
        else { if (keep && keep_lean && lean_shape == 6) DASP_LAUNCH(__half, 6, 0, true); else if (keep && keep_lean && lean_shape == 7) DASP_LAUNCH(__half, 7, 0, true); else if (keep && keep_lean) DASP_LAUNCH(__half, 3, 0, true); else if (keep) DASP_LAUNCH(__half, 0, 0, true); else DASP_LAUNCH(__half, 0, 0, false); }
    } else if (mma_short) { // DMMA short rows (comparison variant): with the plain or the DMMA long / medium paths
        if (med == 1 && mma_long) spmv_kernel<double, 1, 1, false, true><<<grid, CTA, 0, st>>>(a);
        else spmv_kernel<double, 0, 0, false, true><<<grid, CTA, 0, st>>>(a);
    } else if (tma_long) {
        DASP_LAUNCH(double, 0, 2, false);
    } else if (mma_long) {
        if (med == 2) DASP_LAUNCH(double, 2, 1, false);
        else if (med == 1) DASP_LAUNCH(double, 1, 1, false);
        else DASP_LAUNCH(double, 0, 1, false);
    } else {
        if (med == 2) { if (keep) DASP_LAUNCH(double, 2, 0, true); else DASP_LAUNCH(double, 2, 0, false); }
        else if (med == 1) DASP_LAUNCH(double, 1, 0, false);
        else { if (keep && keep_lean && lean_shape == 6) DASP_LAUNCH(double, 6, 0, true); else if (keep && keep_lean && lean_shape == 7) DASP_LAUNCH(double, 7, 0, true); else if (keep && lean_shape == 3) DASP_LAUNCH(double, 5, 0, true); else if (keep && lean_shape == 2) DASP_LAUNCH(double, 4, 0, true); else if (keep && keep_lean) DASP_LAUNCH(double, 3, 0, true); else if (keep) DASP_LAUNCH(double, 0, 0, true); else DASP_LAUNCH(double, 0, 0, false); }
    }
#undef DASP_LAUNCH
    DASP_CUDA(cudaGetLastError());
    return finish_overlap();
}

} // namespace dasp
