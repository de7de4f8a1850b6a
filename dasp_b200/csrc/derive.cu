// derive.cu — everything the SpMV kernels read that is NOT part of the reference's layout, derived on the GPU from the
// bit-exact arrays of preprocess.cu (or of a file read by dasp_load): long-row work units and their merge scratch,
// compact 16-bit column indices, irregular-tail flags, the inverse permutation, and — for long rows whose gathers are
// scattered over x — a column-blocked copy of the long part (LCB) whose kernel stages x in shared memory with TMA.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "dasp_internal.h"

namespace dasp {

namespace {

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline unsigned grid_for(long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// work units per long row: one unit = LONG_UNIT_WARPS reference warps (src/dasp_f64.h:1006 defines the warp)
__global__ void units_per_row(const int *__restrict__ long_rpt_new, int row_long, int unit_warps, int *__restrict__ upr)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_long) return;
    int w = long_rpt_new[i + 1] - long_rpt_new[i];
    upr[i] = (w + unit_warps - 1) / unit_warps;
}

__global__ void __launch_bounds__(256) invert_order(const int *__restrict__ order, int m, int *__restrict__ inv)
{
    int k = blockIdx.x * 256 + threadIdx.x;
    if (k < m) inv[order[k]] = k;
}

// Execution order of the long-row work units.  Inside every group of 8 consecutive long rows the units are
// enumerated chunk-major (chunk c of rows 8g..8g+7, then chunk c+1, ...), so the 8 warps of one CTA work on the
// SAME slot range of 8 neighbouring long rows: rows that are long because they touch the same dense column range
// (borders, constraints) then share their x sectors through that SM's L1.  Rows with fewer chunks simply drop out.
__global__ void fill_long_units(const int *__restrict__ unit_first, int row_long, int *__restrict__ unit_row,
                                int *__restrict__ unit_chunk)
{
    const int r0 = blockIdx.x * 8;
    int n[8], base = unit_first[r0], most = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        n[r] = (r0 + r < row_long) ? unit_first[r0 + r + 1] - unit_first[r0 + r] : 0;
        most = max(most, n[r]);
    }
    for (int c = threadIdx.x; c < most; c += blockDim.x) {
        int pos = base;
#pragma unroll
        for (int r = 0; r < 8; r++) pos += min(n[r], c);
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (n[r] > c) { unit_row[pos] = r0 + r; unit_chunk[pos] = c; pos++; }
    }
}

// one flag per 32 sorted medium rows: does any of them own an irregular tail?
__global__ void flag_irreg(const int *__restrict__ irreg_rpt, int row_block, int ngroups, unsigned char *__restrict__ flag)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    int hi = min(32 * (g + 1), row_block);
    flag[g] = irreg_rpt[hi] != irreg_rpt[32 * g];
}

// Resident compact form of reg_cid: one warp per 8-row block walks its tiles; per tile the smallest non-zero
// column is the base and every slot stores (column - base) in 16 bits.  Column 0 (all padding slots, and genuine
// entries of column 0) is the sentinel 0xFFFF, so the kernels gather exactly the x entries the reference layout
// names.  A block with a tile spanning >= 65535 columns is flagged wide and keeps using reg_cid.
template <typename T>
__global__ void compress_cid(const int *__restrict__ blockPtr, const int *__restrict__ reg_cid, const T *__restrict__ reg_val,
                             int blocknum, int *__restrict__ cbase, unsigned short *__restrict__ cdelta,
                             unsigned char *__restrict__ wide, unsigned short *__restrict__ live)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= blocknum) return;
    const int lane = threadIdx.x & 31;
    const int bp0 = blockPtr[b], bp1 = blockPtr[b + 1];
    bool any_wide = false;
    int nlive = 0; // tiles up to the last one that holds a non-zero value: the kernels stop there (the FP16 layout pads
                   // every block to a multiple of 4 tiles, src/dasp_f16.h:1356; a tile of zeros contributes nothing)
    for (int p = bp0; p < bp1; p += 32) {
        if (sizeof(T) == 2) { // only the FP16 layout pads blocks with whole tiles; the FP64 kernels do not read `live`
            if (__any_sync(0xffffffffu, reg_val[p + lane] != T(0))) nlive = ((p - bp0) >> 5) + 1;
        } else nlive = ((p - bp0) >> 5) + 1;
        const int c = reg_cid[p + lane];
        int mn = c ? c : INT32_MAX, mx = c;
        for (int o = 16; o; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (mn == INT32_MAX) mn = 0; // tile of zeros only
        const bool ok = (mx - mn) < 65535;
        any_wide |= !ok;
        cdelta[p + lane] = (unsigned short)(c == 0 ? 0xFFFF : (ok ? c - mn : 0));
        if (lane == 0) cbase[p >> 5] = mn;
    }
    if (lane == 0) { wide[b] = any_wide ? 1 : 0; live[b] = (unsigned short)min(nlive, 65535); }
}

// The same compact index form for the long part: one warp per work unit (execution order), one base per 32-slot
// group; a unit with a group spanning >= 65535 columns is flagged wide and keeps using long_cid.
__global__ void compress_long_cid(const int *__restrict__ unit_row, const int *__restrict__ unit_chunk,
                                  const int *__restrict__ long_rpt_new, const int *__restrict__ long_cid, int n_units,
                                  int longw, int unit_warps, int esz, int *__restrict__ cbase, unsigned short *__restrict__ cdelta,
                                  unsigned char *__restrict__ wide, unsigned long long *__restrict__ lines)
{
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= n_units) return;
    const int lane = threadIdx.x & 31;
    const int row = unit_row[u];
    const long row_end = (long)long_rpt_new[row + 1] * longw;
    const long beg = (long)long_rpt_new[row] * longw + (long)unit_chunk[u] * unit_warps * longw;
    const long end = min(beg + (long)unit_warps * longw, row_end);
    bool any_wide = false;
    unsigned long long nlines = 0;
    for (long p = beg; p < end; p += 32) {
        const int c = long_cid[p + lane];
        int mn = c ? c : INT32_MAX, mx = c;
        for (int o = 16; o; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (mn == INT32_MAX) mn = 0;
        // 128-byte lines of x one 32-lane gather of this group touches: lanes whose line differs from their left
        // neighbour's (exact for ascending columns, ~32 for scattered ones)
        const int line = (int)(((long)c * esz) >> 7);
        const int left = __shfl_up_sync(0xffffffffu, line, 1);
        nlines += (unsigned long long)__popc(__ballot_sync(0xffffffffu, lane == 0 || line != left));
        const bool ok = (mx - mn) < 65535;
        any_wide |= !ok;
        cdelta[p + lane] = (unsigned short)(c == 0 ? 0xFFFF : (ok ? c - mn : 0));
        if (lane == 0) cbase[p >> 5] = mn;
    }
    if (lane == 0) wide[u] = any_wide ? 1 : 0;
    if (lane == 0 && nlines) atomicAdd(lines, nlines);
}


// ---- locality-ordered work lists ------------------------------------------------------------------------------
__global__ void med_group_keys(const int *__restrict__ order_rid, int row_long, int row_block, int ngroups, int *__restrict__ key,
                               int *__restrict__ val)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const long r = 32L * g;
    key[g] = r < row_block ? order_rid[row_long + r] : INT32_MAX; // padding groups last
    val[g] = g;
}

// positions where the identity order of the groups steps DOWN in original row id: out[0] = count, out[1..] = positions
__global__ void med_descents(const int *__restrict__ key, int ngroups, int *__restrict__ out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < 1 || g >= ngroups) return;
    if (key[g] < key[g - 1]) {
        const int slot = atomicAdd(out, 1);
        if (slot < 4096) out[1 + slot] = g;
    }
}

struct ShortOrderGeom {
    int ctas[4];  // CTAs of singles, 1&3, 3&4, 2&2
    int rows[4];  // y entries one CTA covers in each
    int ybase[4]; // first permuted index of each segment
    int count[4]; // y entries of each segment
    int pair_group; // 8 (f64) / 32 (f16): one-rows of a 1&3 group precede its three-rows
};
// one key per short CTA: original id of the first row it produces
__global__ void short_cta_keys(const int *__restrict__ order_rid, ShortOrderGeom g, int total, int *__restrict__ key, int *__restrict__ val)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = 0, local = i;
    while (c < 3 && local >= g.ctas[c]) { local -= g.ctas[c]; c++; }
    long first = (long)local * g.rows[c];
    if (c == 1 && first + g.pair_group < g.count[c]) first += g.pair_group; // the 3-row of the first pair (3 of its 4 entries)
    key[i] = first < g.count[c] ? order_rid[g.ybase[c] + first] : INT32_MAX;
    val[i] = ((c + 2) << 28) | local;
}

// ---- medium-band kernel: x window per CTA (8 consecutive groups of the processing order) ----------------------------
// One CTA per chunk of 8 groups (256 rows).  Every row contributes the column of its first entry as a sample; the window
// is centred on the MEDIAN sample (robust against the scattered columns a power-law row mixes into its window).
// The same pass counts, for the first gather instruction of every group (tile 0, element 0), how many distinct 128-byte
// lines of x its 32 lanes touch.
__global__ void __launch_bounds__(256) mb_place(const int *__restrict__ order, int ngroups4, int row_block, const int *__restrict__ blockPtr,
                                                const int *__restrict__ reg_cid, const int *__restrict__ irreg_rpt,
                                                const int *__restrict__ irreg_cid, int esz, int wcap, int ncols, int *__restrict__ mb_lo,
                                                unsigned long long *__restrict__ lines)
{
    __shared__ int samp[256];
    __shared__ int med;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = blockIdx.x * 8 + warp;
    int sample = INT32_MAX;
    if (item < ngroups4) {
        const int G = order ? order[item] : item;
        const int g = 32 * G + lane;
        if (g < row_block) {
            const int b = g >> 3, r = g & 7;
            const int bp0 = blockPtr[b], bp1 = blockPtr[b + 1];
            if (bp1 > bp0) sample = reg_cid[bp0 + 4 * r];
            else if (irreg_rpt[g + 1] > irreg_rpt[g]) sample = irreg_cid[irreg_rpt[g]];
        }
        // distinct lines among the 32 lanes of this gather
        const int line = sample == INT32_MAX ? -1 - lane : (int)(((long)sample * esz) >> 7);
        const unsigned peers = __match_any_sync(0xffffffffu, line);
        const unsigned leaders = __ballot_sync(0xffffffffu, (peers & ((1u << lane) - 1u)) == 0 && sample != INT32_MAX);
        if (lane == 0) atomicAdd(lines, ((unsigned long long)__popc(leaders) << 32) | 1ull); // sum of lines in the high word, gathers in the low
    }
    samp[tid] = sample;
    if (tid == 0) med = 0;
    __syncthreads();
    int valid = 0, rank = 0;
    for (int k = 0; k < 256; k++) {
        const int v = samp[k];
        valid += v != INT32_MAX;
        rank += (v < sample) || (v == sample && k < tid);
    }
    if (sample != INT32_MAX && rank == valid / 2) med = sample;
    __syncthreads();
    if (tid == 0) {
        long lo = (long)med - wcap / 2;
        if (lo > (long)ncols - wcap) lo = (long)ncols - wcap;
        if (lo < 0) lo = 0;
        mb_lo[blockIdx.x] = (int)(lo & ~7L);
    }
}

// entries of the chunk inside / outside its window (regular tiles without padding, irregular tails)
template <typename T>
__global__ void __launch_bounds__(256) mb_hits(const int *__restrict__ order, int ngroups4, int row_block, const int *__restrict__ blockPtr,
                                               const int *__restrict__ reg_cid, const T *__restrict__ reg_val, const int *__restrict__ irreg_rpt,
                                               const int *__restrict__ irreg_cid, const int *__restrict__ mb_lo, int wcap,
                                               unsigned long long *__restrict__ counts)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = blockIdx.x * 8 + warp;
    if (item >= ngroups4) return;
    const int G = order ? order[item] : item;
    const int g = 32 * G + lane;
    const int lo = mb_lo[blockIdx.x];
    int in = 0, all = 0;
    if (g < row_block) {
        const int b = g >> 3, r = g & 7;
        const int bp0 = blockPtr[b], bp1 = blockPtr[b + 1];
        for (int p = bp0 + 4 * r; p < bp1; p += 32)
            for (int e = 0; e < 4; e++) {
                const int c = reg_cid[p + e];
                if (!(c == 0 && reg_val[p + e] == T(0))) { all++; in += (unsigned)(c - lo) < (unsigned)wcap; }
            }
        for (int i = irreg_rpt[g]; i < irreg_rpt[g + 1]; i++) { all++; in += (unsigned)(irreg_cid[i] - lo) < (unsigned)wcap; }
    }
    for (int o = 16; o; o >>= 1) { in += __shfl_xor_sync(0xffffffffu, in, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
    if (lane == 0) { atomicAdd(counts, (unsigned long long)in); atomicAdd(counts + 1, (unsigned long long)all); }
}

// ---- short-band kernel: warp items of the short segments by row band, x window per band --------------------------
struct SbGeom {
    int items[4];   // warp items of singles, 1&3, 3&4, 2&2 (as the fused kernel counts them)
    int yrows[4];   // y entries one item covers
    int ybase[4], count[4];
    int sbase[4];   // first slot of each segment in short_val / short_cid
    int slots[4];   // slots one item covers
    int nslots[4];  // slots of the segment that exist
    int pair_group;
    int nbands, band_rows;
};
__device__ __forceinline__ int sb_locate(const SbGeom &g, int i, int &local)
{
    int c = 0;
    local = i;
    while (c < 3 && local >= g.items[c]) { local -= g.items[c]; c++; }
    return c;
}
// key = band of the original row the item starts with
__global__ void sb_item_keys(const int *__restrict__ order_rid, SbGeom g, int total, int *__restrict__ key, int *__restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int local;
    const int c = sb_locate(g, i, local);
    long first = (long)local * g.yrows[c];
    if (c == 1 && first + g.pair_group < g.count[c]) first += g.pair_group;
    key[i] = first < g.count[c] ? order_rid[g.ybase[c] + first] / g.band_rows : g.nbands;
    val[i] = ((c + 2) << 28) | local;
}
// smallest column of every band (padding slots, value 0 and column 0, ignored): one warp per item
template <typename T>
__global__ void sb_min_col(const int *__restrict__ sorted_item, const int *__restrict__ sorted_band, int nitems, SbGeom g,
                           const T *__restrict__ short_val, const int *__restrict__ short_cid, int *__restrict__ lo,
                           int *__restrict__ hi)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= nitems) return;
    const int e = sorted_item[i], c = (e >> 28) - 2, local = e & 0x0FFFFFFF;
    const long s0 = (long)local * g.slots[c];
    int mn = INT32_MAX, mx = -1;
    for (int k = lane; k < g.slots[c]; k += 32) {
        const long p = s0 + k;
        if (p < g.nslots[c]) {
            const int col = short_cid[g.sbase[c] + p];
            // (the single of a 1&3 pair, slot 0 of its tile row, belongs to a row from elsewhere: it does not place the window)
            if (!(col == 0 && short_val[g.sbase[c] + p] == T(0)) && !(c == 1 && (p & 3) == 0)) { mn = min(mn, col); mx = max(mx, col); }
        }
    }
    for (int o = 16; o; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0 && mx >= 0) { atomicMin(lo + sorted_band[i], mn); atomicMax(hi + sorted_band[i], mx); }
}
// window start of every band: the smallest column when all columns of the band fit one window (any banded structure,
// row slabs included); otherwise the smallest column but not left of (diagonal position of the band - half a window), so
// that a few outlying columns cannot drag the window away from where a square matrix keeps the band's entries; multiple of 8
__global__ void sb_place_windows(int *__restrict__ lo, const int *__restrict__ hi, int nbands, int band_rows, int m, int n, int wcap)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbands) return;
    long v = lo[b] == INT32_MAX ? 0 : lo[b];
    if ((long)hi[b] - v >= wcap) {
        const long centre = (long)(((double)b * band_rows + band_rows / 2) * (double)n / (double)(m > 0 ? m : 1));
        v = max(v, centre - wcap / 2);
    }
    v = min(v, (long)max(0, n - 1));
    lo[b] = (int)(v & ~7L);
}
__global__ void sb_fix_unset(int *__restrict__ lo, int n)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n && lo[b] == 0x7f7f7f7f) lo[b] = INT32_MAX;
}
// entries inside / outside their band's window
template <typename T>
__global__ void sb_hits(const int *__restrict__ sorted_item, const int *__restrict__ sorted_band, int nitems, SbGeom g,
                        const T *__restrict__ short_val, const int *__restrict__ short_cid, const int *__restrict__ lo, int wcap,
                        unsigned long long *__restrict__ counts)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= nitems) return;
    const int e = sorted_item[i], c = (e >> 28) - 2, local = e & 0x0FFFFFFF;
    const long s0 = (long)local * g.slots[c];
    const int wlo = lo[sorted_band[i]];
    int in = 0, all = 0;
    for (int k = lane; k < g.slots[c]; k += 32) {
        const long p = s0 + k;
        if (p < g.nslots[c]) {
            const int col = short_cid[g.sbase[c] + p];
            if (!(col == 0 && short_val[g.sbase[c] + p] == T(0))) { all++; in += (unsigned)(col - wlo) < (unsigned)wcap; }
        }
    }
    for (int o = 16; o; o >>= 1) { in += __shfl_xor_sync(0xffffffffu, in, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
    if (lane == 0) { atomicAdd(counts, (unsigned long long)in); atomicAdd(counts + 1, (unsigned long long)all); }
}

// ---- LCB: column-blocked copy of the long part -----------------------------------------------------------------

// long row of every reference warp (the reference's rid_by_warp, src/dasp_f64.h:1021-1031)
__global__ void lcb_warp_rows(const int *__restrict__ long_rpt_new, int row_long, int *__restrict__ warp_row)
{
    const int i = blockIdx.x;
    for (int w = long_rpt_new[i] + threadIdx.x; w < long_rpt_new[i + 1]; w += blockDim.x) warp_row[w] = i;
}

// sort key of every slot of the padded long part: its column block; padding (value 0 AND column 0) sorts last
template <typename T>
__global__ void lcb_keys(const T *__restrict__ long_val, const int *__restrict__ ref_cid, const int *__restrict__ long_cid, int slots,
                         int bw_log2, int nblk, int *__restrict__ key, int *__restrict__ idx)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= slots) return;
    const bool live = !(ref_cid[p] == 0 && long_val[p] == T(0)); // padding as the reference writes it
    key[p] = live ? (long_cid[p] >> bw_log2) : nblk;
    idx[p] = p;
}

// first entry of every block in the sorted key sequence (lower bound), one thread per block
__global__ void lcb_block_ptr(const int *__restrict__ sorted_key, int slots, int nblk, int *__restrict__ blk_ptr)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblk) return;
    int lo = 0, hi = slots;
    while (lo < hi) {
        const int mid = (int)(((long)lo + hi) >> 1);
        if (sorted_key[mid] < b) lo = mid + 1; else hi = mid;
    }
    blk_ptr[b] = lo;
}

// padded entry count of every block: a multiple of 1024 (the chunk one warp walks: 8 steps of 128 entries, a lane loads four
// values and four indices per step), so that chunks - and the parts of the CTAs, multiples of 1024 too - start on global
// multiples of 1024 and a chunk's restart row is found at entry >> 10
__global__ void lcb_padded_counts(const int *__restrict__ blk_ptr, int nblk, int part, int *__restrict__ pad_ptr,
                                  int *__restrict__ cta_first)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblk) return;
    const int cnt = b < nblk ? ((blk_ptr[b + 1] - blk_ptr[b] + 1023) & ~1023) : 0;
    pad_ptr[b] = cnt;
    cta_first[b] = (cnt + part - 1) / part;
}

// entry i of the padded, blocked sequence: value, and (long row << 16 | column inside the block); pad entries repeat the
// row of the last real entry of their block with value 0 and column 0
template <typename T>
__global__ void lcb_gather(const T *__restrict__ long_val, const int *__restrict__ long_cid, const int *__restrict__ src,
                           const int *__restrict__ warp_row, const int *__restrict__ blk_ptr, const int *__restrict__ pad_ptr,
                           int nblk, int total, int longw, int bw_mask, T *__restrict__ val, unsigned *__restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = nblk; // last block b with pad_ptr[b] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pad_ptr[mid] <= i) lo = mid; else hi = mid;
    }
    const int b = lo, j = i - pad_ptr[b], cnt = blk_ptr[b + 1] - blk_ptr[b];
    const int p = src[blk_ptr[b] + min(j, cnt - 1)];
    const unsigned row = (unsigned)warp_row[p / longw];
    if (j < cnt) { val[i] = long_val[p]; idx[i] = (row << 16) | (unsigned)(long_cid[p] & bw_mask); }
    else { val[i] = T(0); idx[i] = row << 16; }
}

// Bank-aware order inside the runs (FP64).  lcb_kernel gathers x from shared memory with 64-bit loads; the hardware serves a
// quarter warp (8 lanes x 8 bytes) per pass, and two of its lanes conflict when their columns are equal modulo 16.  Lane l of
// a step owns entries 4l .. 4l + 3, so gather j of lanes 8g .. 8g + 7 reads the stride-4 subsequence {32g + 4i + j, i < 8} of an
// aligned 32-entry window: with random columns that is 1.9 passes instead of 1 (measured: 130 M of 264 M shared-memory
// wavefronts on C5 are conflicts, and the load/store unit is what bounds the kernel).  The order of the entries INSIDE a run
// (same row, same block) is free - the kernel sums a run in any order - so every run (cut at multiples of 1024 entries) is
// re-ordered here: entries ranked inside their bank, sorted by (rank, bank) - any 16 consecutive ones then sit in distinct
// banks as long as the banks are evenly filled - and dealt to the run's positions window by window, column j by column j.
// One CTA per 1024-entry chunk, in place (the chunk is staged in shared memory); runs shorter than 16 entries stay as they are.
__global__ void __launch_bounds__(128) lcb_bank_order(double *__restrict__ val, unsigned *__restrict__ idx, int total)
{
    __shared__ double sv[1024];
    __shared__ unsigned si[1024];
    __shared__ unsigned short dest[1024], seg[1025];
    __shared__ int cnt[4][16], nseg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long base = (long)blockIdx.x * 1024;
    if (base >= total) return;
    for (int e = tid; e < 1024; e += 128) { sv[e] = val[base + e]; si[e] = idx[base + e]; dest[e] = (unsigned short)e; }
    __syncthreads();
    if (warp == 0) { // list of run starts
        int n = 0;
        for (int k = 0; k < 32; k++) {
            const int e = 32 * k + lane;
            const bool head = e == 0 || (si[e] >> 16) != (si[e - 1] >> 16);
            const unsigned m = __ballot_sync(0xffffffffu, head);
            if (head) seg[n + __popc(m & ((1u << lane) - 1))] = (unsigned short)e;
            n += __popc(m);
        }
        if (lane == 0) { seg[n] = 1024; nseg = n; }
    }
    __syncthreads();
    for (int s = warp; s < nseg; s += 4) {
        const int s0 = seg[s], s1 = seg[s + 1], len = s1 - s0;
        if (len < 16) continue;
        if (lane < 16) cnt[warp][lane] = 0;
        __syncwarp();
        for (int e = s0 + lane; e < s1; e += 32) atomicAdd(&cnt[warp][si[e] & 15], 1);
        __syncwarp();
        int run[16]; // entries of every bank seen so far (the same in every lane)
#pragma unroll
        for (int b = 0; b < 16; b++) run[b] = 0;
        for (int e0 = s0; e0 < s1; e0 += 32) {
            const int e = e0 + lane;
            const bool on = e < s1;
            const int b = on ? (int)(si[e] & 15) : 16;
            int rank = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const unsigned m = __ballot_sync(0xffffffffu, b == k);
                if (b == k) { rank = run[k] + __popc(m & ((1u << lane) - 1)); }
                run[k] += __popc(m);
            }
            if (on) {
                // position in the (rank, bank) order: everything of a smaller rank, and the lower banks of the same rank
                int q = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int c = cnt[warp][k];
                    q += min(c, rank) + ((k < b && c > rank) ? 1 : 0);
                }
                // the q-th position of the run, positions taken window by window (aligned 32 entries) and inside a window by
                // (p % 4, p / 4): the entries of gather j of a quarter warp are then consecutive in the (rank, bank) order
                int lo = s0, p = -1;
                while (p < 0) {
                    const int hi = min(s1, (lo & ~31) + 32), n = hi - lo;
                    if (q < n) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int first = lo + ((j - lo) & 3); // first position >= lo with position % 4 == j
                            const int cj = first < hi ? ((hi - 1 - first) >> 2) + 1 : 0;
                            if (p < 0) { if (q < cj) p = first + 4 * q; else q -= cj; }
                        }
                    } else { q -= n; lo = hi; }
                }
                dest[e] = (unsigned short)p;
            }
        }
        __syncwarp(); // every lane is done with cnt before the next run resets it
    }
    __syncthreads();
    for (int e = tid; e < 1024; e += 128) { const int p = dest[e]; val[base + p] = sv[e]; idx[base + p] = si[e]; }
}

// 16-bit form of the packed indices (FP64 blocks are 8192 columns wide: 13 bits): column | row delta << 13, restart row per
// chunk of 1024 entries, chunks with a delta > 7 flagged wide
__global__ void lcb_encode16(const unsigned *__restrict__ idx, int total, unsigned short *__restrict__ idx16,
                             int *__restrict__ chunk_row, unsigned char *__restrict__ chunk_wide)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned e = idx[i], row = e >> 16;
    unsigned delta = 0;
    if (i & 1023) delta = row - (idx[i - 1] >> 16); // rows ascend inside a block, and blocks start on multiples of 1024
    else chunk_row[i >> 10] = (int)row;
    if (delta > 7) { chunk_wide[i >> 10] = 1; delta = 7; }
    idx16[i] = (unsigned short)((e & 0x1FFFu) | (delta << 13));
}

template <typename T> int build_lcb_t(dasp_handle *h, cudaStream_t st)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    if (L.lcb_val || s.row_long == 0 || s.row_long > 65535 || s.fill0_nnz_long == 0) return DASP_OK;
    DevicePool tmp;
    struct Guard { DevicePool &p; ~Guard() { p.free_all(); } } guard{tmp};
    DevicePool &pool = h->pool;
    const int slots = s.fill0_nnz_long, longw = sizeof(T) == 2 ? 256 : 64;
    int bw_log2 = 0;
    while ((sizeof(T) << (bw_log2 + 1)) <= (size_t)LCB_BYTES) bw_log2++; // 8192 doubles / 32768 halves
    const int nblk = L.x_len > 0 ? (int)((((int64_t)L.x_len - 1) >> bw_log2) + 1) : 1;
    int *warp_row = nullptr, *key = nullptr, *idx = nullptr, *skey = nullptr, *sidx = nullptr, *blk_ptr = nullptr;
    // the sort's scratch in one allocation: four key / index arrays, the radix sort's two ping-pong buffers and histograms
    tmp.reserve(sizeof(int) * (6 * (size_t)slots + (size_t)s.warp_number + (size_t)nblk + (size_t)slots / 4 + 65536) + (8u << 20));
    DASP_TRY(tmp.alloc((void **)&warp_row, sizeof(int) * (size_t)s.warp_number));
    DASP_TRY(tmp.alloc((void **)&key, sizeof(int) * (size_t)slots));
    DASP_TRY(tmp.alloc((void **)&idx, sizeof(int) * (size_t)slots));
    DASP_TRY(tmp.alloc((void **)&skey, sizeof(int) * (size_t)slots));
    DASP_TRY(tmp.alloc((void **)&sidx, sizeof(int) * (size_t)slots));
    DASP_TRY(tmp.alloc((void **)&blk_ptr, sizeof(int) * (size_t)(nblk + 1)));
    DASP_CUDA(cudaMemsetAsync(warp_row, 0, sizeof(int) * (size_t)s.warp_number, st));
    lcb_warp_rows<<<s.row_long, 256, 0, st>>>(L.long_rpt_new, s.row_long, warp_row);
    lcb_keys<T><<<grid_for(slots, 256), 256, 0, st>>>((const T *)L.long_val, L.long_cid, L.k_long_cid, slots, bw_log2, nblk, key, idx);
    int bits = 1;
    while ((1 << bits) <= nblk) bits++;
    DASP_TRY(radix_sort_pairs(tmp, key, idx, skey, sidx, slots, bits, false, st));
    DASP_TRY(pool.alloc((void **)&L.lcb_blk_ptr, sizeof(int) * (size_t)(nblk + 1)));
    DASP_TRY(pool.alloc((void **)&L.lcb_cta_first, sizeof(int) * (size_t)(nblk + 1)));
    lcb_block_ptr<<<grid_for(nblk + 1, 256), 256, 0, st>>>(skey, slots, nblk, blk_ptr);
    // entries per CTA: a bigger part amortises the 64 KB block of x a CTA stages, a smaller one gives a small matrix enough
    // CTAs to balance (measured, profiles/r02: C5 1e9 entries 2.10 -> 1.99 ms with 65536, C3 1.4e8 entries 0.339 -> 0.354 ms)
    const int part = (long)s.nnz_long >= 16L * 444 * 65536 ? 65536 : LCB_PART;
    lcb_padded_counts<<<grid_for(nblk + 1, 256), 256, 0, st>>>(blk_ptr, nblk, part, L.lcb_blk_ptr, L.lcb_cta_first);
    DASP_TRY(scan_inplace(tmp, L.lcb_blk_ptr, nblk + 1, st));
    DASP_TRY(scan_inplace(tmp, L.lcb_cta_first, nblk + 1, st));
    int tot[2] = {0, 0};
    DASP_CUDA(cudaMemcpyAsync(&tot[0], L.lcb_blk_ptr + nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    DASP_CUDA(cudaMemcpyAsync(&tot[1], L.lcb_cta_first + nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    DASP_CUDA(cudaStreamSynchronize(st));
    const int total = tot[0];
    if (total < 0) { set_error("column-blocked long part exceeds 32-bit offsets"); return DASP_ERR_RANGE; }
    DASP_TRY(pool.alloc(&L.lcb_val, sizeof(T) * (size_t)total));
    DASP_TRY(pool.alloc((void **)&L.lcb_idx, sizeof(unsigned) * (size_t)total));
    const size_t acc_stride = ((size_t)s.row_long + 31) & ~(size_t)31; // copies start on their own 256-byte boundary
    DASP_TRY(pool.alloc((void **)&L.lcb_done, sizeof(unsigned) * 4));
    DASP_TRY(pool.alloc(&L.lcb_acc, 8 * acc_stride * LCB_COPIES));
    DASP_CUDA(cudaMemsetAsync(L.lcb_acc, 0, 8 * acc_stride * LCB_COPIES, st));
    DASP_CUDA(cudaMemsetAsync(L.lcb_done, 0, sizeof(unsigned) * 4, st));
    if (total > 0)
        lcb_gather<T><<<grid_for(total, 256), 256, 0, st>>>((const T *)L.long_val, L.k_long_cid, sidx, warp_row, blk_ptr, L.lcb_blk_ptr,
                                                           nblk, total, longw, (1 << bw_log2) - 1, (T *)L.lcb_val, L.lcb_idx);
    if constexpr (sizeof(T) == 8) {
        // Measured (profiles/r02/README.md section 4): bank conflicts 130 M -> 86 M wavefronts, 1.99 -> 1.90 ms under ncu, but no
        // difference between back-to-back launches (2.446 vs 2.454 ms per C5 product), and the pass costs 80 ms on C3: opt-in only
        static const int bank_env = getenv("DASP_LCB_BANK_ORDER") ? atoi(getenv("DASP_LCB_BANK_ORDER")) : 0;
        if (bank_env && total > 0) lcb_bank_order<<<total >> 10, 128, 0, st>>>((double *)L.lcb_val, L.lcb_idx, total);
    }
    // FP64, DASP_LCB_IDX16=1 only: the 16-bit index stream.  Measured (profiles/r02/README.md section 4): the kernel then moves 10.4
    // instead of 12.6 GB on C5 (DRAM 66 vs 78 %) and is NOT faster (2.07 vs 2.00 ms; C3 0.37 vs 0.33 ms with its many wide
    // chunks): the row reconstruction's warp scan costs six more load/store-unit wavefronts per step in a kernel whose
    // load/store unit is already 75 % busy.  Not built unless asked for.
    static const int idx16_env = getenv("DASP_LCB_IDX16") ? atoi(getenv("DASP_LCB_IDX16")) : 0;
    if (idx16_env && sizeof(T) == 8 && bw_log2 <= 13 && total > 0) {
        const size_t nchunks = (size_t)total >> 10;
        DASP_TRY(pool.alloc((void **)&L.lcb_idx16, sizeof(unsigned short) * (size_t)total));
        DASP_TRY(pool.alloc((void **)&L.lcb_chunk_row, sizeof(int) * nchunks));
        DASP_TRY(pool.alloc((void **)&L.lcb_chunk_wide, nchunks));
        DASP_CUDA(cudaMemsetAsync(L.lcb_chunk_wide, 0, nchunks, st));
        lcb_encode16<<<grid_for(total, 256), 256, 0, st>>>(L.lcb_idx, total, L.lcb_idx16, L.lcb_chunk_row, L.lcb_chunk_wide);
    }
    DASP_CUDA(cudaGetLastError());
    DASP_CUDA(cudaStreamSynchronize(st)); // the scratch is released by the guard
    L.lcb_bw_log2 = bw_log2; L.lcb_nblk = nblk; L.lcb_live = total; L.lcb_nctas = tot[1];
    return DASP_OK;
}

// flag |= 1 unless a[0] == first, a non-decreasing and a[count-1] == last
__global__ void check_offsets(const int *__restrict__ a, long count, int first, int last, int *__restrict__ flag)
{
    bool bad = false;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        const int v = a[i];
        if (i == 0 && v != first) bad = true;
        if (i == count - 1 ? v != last : v > a[i + 1]) bad = true;
    }
    if (bad) atomicOr(flag, 1);
}
// flag |= 1 unless every a[i] lies in [0, hi)
__global__ void check_range(const int *__restrict__ a, long count, int hi, int *__restrict__ flag)
{
    bool bad = false;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        const int v = a[i];
        if (v < 0 || v >= hi) bad = true;
    }
    if (bad) atomicOr(flag, 1);
}
// flag |= 1 unless a[0..count) is a permutation of 0..count-1 (seen: zeroed scratch of count ints)
__global__ void check_permutation(const int *__restrict__ a, int count, int *__restrict__ seen, int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int v = a[i];
    if (v < 0 || v >= count || atomicExch(seen + v, 1) != 0) atomicOr(flag, 1);
}

} // namespace

int validate_layout(dasp_handle *h, cudaStream_t st)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    DevicePool tmp;
    struct Guard { DevicePool &p; ~Guard() { p.free_all(); } } guard{tmp};
    int *flag = nullptr, *seen = nullptr;
    DASP_TRY(tmp.alloc((void **)&flag, sizeof(int)));
    DASP_TRY(tmp.alloc((void **)&seen, sizeof(int) * (size_t)(s.m > 0 ? s.m : 1)));
    DASP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    DASP_CUDA(cudaMemsetAsync(seen, 0, sizeof(int) * (size_t)(s.m > 0 ? s.m : 1), st));
    const int G = 592;
    const int longw = h->dtype == DASP_F16 ? 256 : 64;
    check_offsets<<<G, 256, 0, st>>>(L.blockPtr, (long)s.blocknum + 1, 0, s.fill0_nnz_reg, flag);
    check_offsets<<<G, 256, 0, st>>>(L.irreg_rpt, (long)s.row_block + 1, 0, s.nnz_irreg, flag);
    if (s.row_long > 0) {
        // long_rpt_new[row_long] = warps actually used, rounded up to warp_number by the reference (src/dasp_f64.h:1017)
        int last = 0;
        DASP_CUDA(cudaMemcpyAsync(&last, L.long_rpt_new + s.row_long, sizeof(int), cudaMemcpyDeviceToHost, st));
        DASP_CUDA(cudaStreamSynchronize(st));
        if (last < 0 || last > s.warp_number || (int64_t)last * longw < s.nnz_long) { set_error("dasp_load: long_rpt_new does not match the layout scalars"); return DASP_ERR_INVALID; }
        check_offsets<<<G, 256, 0, st>>>(L.long_rpt_new, (long)s.row_long + 1, 0, last, flag);
    }
    if (s.n > 0) {
        check_range<<<G, 256, 0, st>>>(L.long_cid, s.fill0_nnz_long, s.n, flag);
        check_range<<<G, 256, 0, st>>>(L.reg_cid, s.fill0_nnz_reg, s.n, flag);
        check_range<<<G, 256, 0, st>>>(L.irreg_cid, s.nnz_irreg, s.n, flag);
        check_range<<<G, 256, 0, st>>>(L.short_cid, s.fill0_nnz_short, s.n, flag);
    } else if ((int64_t)s.fill0_nnz_long + s.fill0_nnz_reg + s.nnz_irreg + s.fill0_nnz_short > 0) {
        set_error("dasp_load: entries in a matrix without columns");
        return DASP_ERR_INVALID;
    }
    if (s.m > 0) check_permutation<<<grid_for(s.m, 256), 256, 0, st>>>(L.order_rid, s.m, seen, flag);
    int bad = 0;
    DASP_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    DASP_CUDA(cudaStreamSynchronize(st));
    DASP_CUDA(cudaGetLastError());
    if (bad) { set_error("dasp_load: offset / index arrays of the file are out of range or not monotone"); return DASP_ERR_INVALID; }
    return DASP_OK;
}

int build_lcb(dasp_handle *h, cudaStream_t st)
{
    // T only needs the right size and an exact zero test: half bits as unsigned short (+0; -0 counts as live, harmless)
    int rc = h->dtype == DASP_F16 ? build_lcb_t<unsigned short>(h, st) : build_lcb_t<double>(h, st);
    h->L.s.device_bytes = h->pool.bytes;
    return rc;
}

namespace {

template <typename T>
__global__ void map_cid(const int *__restrict__ src, const T *__restrict__ val, long count, const int *__restrict__ map,
                        int *__restrict__ dst)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        const int c = src[i];
        // padding slots (column 0, value 0) keep pointing at x[0]: they must stay recognisable and inside x
        dst[i] = (c == 0 && val && val[i] == T(0)) ? 0 : map[c];
    }
}

// (re)build the short-band work list and windows from the kernel-facing short column indices
template <typename T> int build_short_bands_t(dasp_handle *h, cudaStream_t st, bool force)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    DevicePool &pool = h->pool;
    DevicePool tmp;
    struct Guard { DevicePool &p; ~Guard() { p.free_all(); } } guard{tmp};
    if (L.sb_item) { pool.release(L.sb_item); pool.release(L.sb_band_ptr); pool.release(L.sb_lo); }
    L.sb_item = nullptr; L.sb_band_ptr = nullptr; L.sb_lo = nullptr; L.sb_nbands = 0; L.sb_nitems = 0; L.sb_auto = 0; L.sb_hit_rate = 0.0;
    const long short_rows = (long)s.short_row_1 + 2L * s.common_13 + s.short_row_34 + s.short_row_2;
    if ((!force && short_rows < (1 << 20)) || short_rows == 0 || s.m <= 0) return DASP_OK; // small short parts stay in the fused kernel
    const int f16 = h->dtype == DASP_F16, G = f16 ? 32 : 8;
    const int tiles13 = ceil_div(s.common_13, 8), tiles34 = ceil_div(s.short_row_34, 8);
    const int tiles22 = ceil_div(s.short_row_2, 2 * G) * (G / 8);
    SbGeom g;
    g.items[0] = ceil_div(s.short_row_1, 32 * SINGLES_PER_THREAD);
    g.items[1] = ceil_div(tiles13, SHORT_TILES_PER_WARP);
    g.items[2] = ceil_div(tiles34, SHORT_TILES_PER_WARP);
    g.items[3] = ceil_div(tiles22, SHORT_TILES_PER_WARP);
    g.yrows[0] = 32 * SINGLES_PER_THREAD; g.yrows[1] = SHORT_TILES_PER_WARP * 16; g.yrows[2] = SHORT_TILES_PER_WARP * 8;
    g.yrows[3] = SHORT_TILES_PER_WARP * 16;
    const int ybase = s.row_long + s.row_block;
    g.ybase[1] = ybase + (f16 ? 0 : s.short_row_1);
    g.ybase[2] = g.ybase[1] + 2 * s.common_13;
    g.ybase[3] = g.ybase[2] + s.short_row_34;
    g.ybase[0] = f16 ? g.ybase[3] + s.short_row_2 : ybase;
    g.count[0] = s.short_row_1; g.count[1] = 2 * s.common_13; g.count[2] = s.short_row_34; g.count[3] = s.short_row_2;
    const int f13 = s.fill0_nnz_short13, f34 = s.fill0_nnz_short34, f22 = s.fill0_nnz_short22;
    g.sbase[0] = f16 ? f13 + f34 + f22 : 0;
    g.sbase[1] = f16 ? 0 : s.short_row_1;
    g.sbase[2] = g.sbase[1] + f13;
    g.sbase[3] = g.sbase[2] + f34;
    g.slots[0] = 32 * SINGLES_PER_THREAD; g.slots[1] = g.slots[2] = g.slots[3] = SHORT_TILES_PER_WARP * 32;
    g.nslots[0] = s.short_row_1; g.nslots[1] = f13; g.nslots[2] = f34; g.nslots[3] = f22;
    g.pair_group = G;
    g.band_rows = SB_BAND_ROWS;
    g.nbands = ceil_div(s.m, SB_BAND_ROWS);
    const int total = g.items[0] + g.items[1] + g.items[2] + g.items[3];
    if (total <= 0) return DASP_OK;
    int *k0 = nullptr, *v0 = nullptr, *k1 = nullptr;
    DASP_TRY(tmp.alloc((void **)&k0, sizeof(int) * (size_t)total));
    DASP_TRY(tmp.alloc((void **)&v0, sizeof(int) * (size_t)total));
    DASP_TRY(tmp.alloc((void **)&k1, sizeof(int) * (size_t)total));
    DASP_TRY(pool.alloc((void **)&L.sb_item, sizeof(int) * (size_t)total));
    DASP_TRY(pool.alloc((void **)&L.sb_band_ptr, sizeof(int) * (size_t)(g.nbands + 2)));
    DASP_TRY(pool.alloc((void **)&L.sb_lo, sizeof(int) * (size_t)(g.nbands + 1)));
    sb_item_keys<<<grid_for(total, 256), 256, 0, st>>>(L.order_rid, g, total, k0, v0);
    int bits = 1;
    while ((1 << bits) <= g.nbands) bits++;
    DASP_TRY(radix_sort_pairs(tmp, k0, v0, k1, L.sb_item, total, bits, false, st));
    lcb_block_ptr<<<grid_for(g.nbands + 2, 256), 256, 0, st>>>(k1, total, g.nbands + 1, L.sb_band_ptr); // lower bounds of 0..nbands+1
    DASP_CUDA(cudaMemsetAsync(L.sb_lo, 0x7f, sizeof(int) * (size_t)(g.nbands + 1), st)); // 0x7f7f7f7f: "no entry yet"
    int nitems = 0;
    DASP_CUDA(cudaMemcpyAsync(&nitems, L.sb_band_ptr + g.nbands, sizeof(int), cudaMemcpyDeviceToHost, st));
    DASP_CUDA(cudaStreamSynchronize(st));
    const int wcap = SB_WINDOW_BYTES / (int)sizeof(T);
    const T *sval = (const T *)L.short_val;
    unsigned long long *counts = nullptr, hc[2] = {0, 0};
    DASP_TRY(tmp.alloc((void **)&counts, sizeof(unsigned long long) * 2));
    DASP_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 2, st));
    if (nitems > 0) {
        int *hi = nullptr;
        DASP_TRY(tmp.alloc((void **)&hi, sizeof(int) * (size_t)(g.nbands + 1)));
        DASP_CUDA(cudaMemsetAsync(hi, 0xff, sizeof(int) * (size_t)(g.nbands + 1), st)); // -1
        sb_min_col<T><<<grid_for((long)nitems * 32, 256), 256, 0, st>>>(L.sb_item, k1, nitems, g, sval, L.k_short_cid, L.sb_lo, hi);
        sb_fix_unset<<<grid_for(g.nbands + 1, 256), 256, 0, st>>>(L.sb_lo, g.nbands + 1);
        sb_place_windows<<<grid_for(g.nbands, 256), 256, 0, st>>>(L.sb_lo, hi, g.nbands, SB_BAND_ROWS, s.m, L.x_len, wcap);
        sb_hits<T><<<grid_for((long)nitems * 32, 256), 256, 0, st>>>(L.sb_item, k1, nitems, g, sval, L.k_short_cid, L.sb_lo, wcap, counts);
    }
    DASP_CUDA(cudaMemcpyAsync(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost, st));
    DASP_CUDA(cudaStreamSynchronize(st));
    DASP_CUDA(cudaGetLastError());
    L.sb_nbands = g.nbands; L.sb_nitems = nitems;
    L.sb_hit_rate = hc[1] ? (double)hc[0] / (double)hc[1] : 0.0;
    // worth it when nearly every gather is served from the window AND every SM gets enough bands for the persistent CTAs to
    // balance (measured, profiles/r02/README.md: +5 % on the 3052 bands of C5, -15 % on the 611 bands of C3)
    const int sms = h->sm_count > 0 ? h->sm_count : 148;
    L.sb_auto = nitems > 0 && L.sb_hit_rate >= 0.8 && g.nbands >= 16 * sms;
    L.s.short_band_hit_rate = L.sb_hit_rate;
    L.s.short_banded = L.sb_auto;
    return DASP_OK;
}

// compact forms of the kernel-facing column indices (and the estimate that decides chunked vs column-blocked long rows)
int derive_indices(dasp_handle *h, cudaStream_t st, unsigned long long *lines)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    const int cl = s.row_long, blocknum = s.blocknum;
    const int longw = h->dtype == DASP_F16 ? 256 : 64, esz = (int)L.esz;
    if (cl > 0 && L.n_long_units > 0)
        compress_long_cid<<<grid_for((long)L.n_long_units * 32, 256), 256, 0, st>>>(
            L.long_unit_row, L.long_unit_chunk, L.long_rpt_new, L.k_long_cid, L.n_long_units, longw, L.long_unit_warps, esz,
            L.long_cbase, L.long_cdelta, L.long_wide, lines);
    if (L.reg_compact_done) L.reg_compact_done = 0; // written by pack_reg from the same indices (dasp_create): nothing to do this once
    else if (blocknum > 0) {
        if (h->dtype == DASP_F16)
            compress_cid<unsigned short><<<grid_for((long)blocknum * 32, 256), 256, 0, st>>>(
                L.blockPtr, L.k_reg_cid, (const unsigned short *)L.reg_val, blocknum, L.reg_cbase, L.reg_cdelta, L.blk_wide, L.blk_live);
        else
            compress_cid<double><<<grid_for((long)blocknum * 32, 256), 256, 0, st>>>(
                L.blockPtr, L.k_reg_cid, (const double *)L.reg_val, blocknum, L.reg_cbase, L.reg_cdelta, L.blk_wide, L.blk_live);
    }
    DASP_CUDA(cudaGetLastError());
    return DASP_OK;
}

// chunked or column-blocked long rows: average number of distinct 128-byte lines of x that one 32-lane gather of the
// chunked kernel touches (estimated from the column span of every 32-slot group)
int decide_long_variant(dasp_handle *h, cudaStream_t st, const unsigned long long *lines)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    L.long_lines_avg = 0.0;
    h->lcb_auto = 0;
    L.s.long_gather_lines = 0.0;
    L.s.long_blocked = 0;
    if (s.row_long > 0 && s.fill0_nnz_long > 0) {
        unsigned long long nl = 0;
        DASP_CUDA(cudaMemcpyAsync(&nl, lines, sizeof(nl), cudaMemcpyDeviceToHost, st));
        DASP_CUDA(cudaStreamSynchronize(st));
        L.long_lines_avg = (double)nl / ((double)s.fill0_nnz_long / 32.0);
        L.s.long_gather_lines = L.long_lines_avg;
        double thr = 9.0; // measured crossover (profiles/r02/README.md): 5.1 lines (C5 sorted) -> chunked wins 1.69 vs 1.83 ms; 12.5 (C3 sorted) -> column-blocked wins 0.339 vs 0.372 ms; 31.7 / 32 (spec generators) -> column-blocked wins by 1.9x / 8x
        if (const char *e = getenv("DASP_LCB_THRESHOLD")) thr = atof(e);
        // (the column-blocked kernel needs a couple of CTAs per SM of 32768 entries each to fill the machine: a small long part
        // stays with the chunked kernel and its finer units: 20.5 us blocked on the 1.1 M-entry case above)
        static const long lcb_min = getenv("DASP_LCB_MIN_NNZ") ? atol(getenv("DASP_LCB_MIN_NNZ")) : 2L * 148 * LCB_PART;
        if (L.long_lines_avg > thr && s.row_long <= 65535 && s.nnz_long >= lcb_min) {
            DASP_TRY(build_lcb(h, st));
            h->lcb_auto = L.lcb_nctas > 0;
            L.s.long_blocked = h->lcb_auto;
        }
    }
    return DASP_OK;
}

} // namespace

int build_medium_bands(dasp_handle *h, cudaStream_t st)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    DevicePool &pool = h->pool;
    DevicePool tmp;
    struct Guard { DevicePool &p; ~Guard() { p.free_all(); } } guard{tmp};
    if (L.mb_lo) { pool.release(L.mb_lo); L.mb_lo = nullptr; }
    L.mb_auto = 0; L.mb_hit_rate = 0.0; L.med_gather_lines = 0.0;
    L.s.medium_banded = 0; L.s.medium_band_hit_rate = 0.0; L.s.medium_gather_lines = 0.0;
    const int ngroups4 = s.blocknum / 4;
    if (ngroups4 == 0 || s.row_block == 0) return DASP_OK;
    const int nchunks = ceil_div(ngroups4, 8), esz = (int)L.esz, wcap = MB_WINDOW_BYTES / esz;
    DASP_TRY(pool.alloc((void **)&L.mb_lo, sizeof(int) * (size_t)nchunks));
    unsigned long long *counts = nullptr, hc[3] = {0, 0, 0};
    DASP_TRY(tmp.alloc((void **)&counts, sizeof(unsigned long long) * 3));
    DASP_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 3, st));
    mb_place<<<nchunks, 256, 0, st>>>(L.med_order, ngroups4, s.row_block, L.blockPtr, L.k_reg_cid, L.irreg_rpt, L.k_irreg_cid, esz, wcap,
                                      L.x_len, L.mb_lo, counts + 2);
    if (h->dtype == DASP_F16)
        mb_hits<unsigned short><<<nchunks, 256, 0, st>>>(L.med_order, ngroups4, s.row_block, L.blockPtr, L.k_reg_cid,
                                                         (const unsigned short *)L.reg_val, L.irreg_rpt, L.k_irreg_cid, L.mb_lo, wcap, counts);
    else
        mb_hits<double><<<nchunks, 256, 0, st>>>(L.med_order, ngroups4, s.row_block, L.blockPtr, L.k_reg_cid, (const double *)L.reg_val,
                                                 L.irreg_rpt, L.k_irreg_cid, L.mb_lo, wcap, counts);
    DASP_CUDA(cudaMemcpyAsync(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost, st));
    DASP_CUDA(cudaStreamSynchronize(st));
    DASP_CUDA(cudaGetLastError());
    L.mb_hit_rate = hc[1] ? (double)hc[0] / (double)hc[1] : 0.0;
    const unsigned long long gathers = hc[2] & 0xffffffffull, lsum = hc[2] >> 32;
    L.med_gather_lines = gathers ? (double)lsum / (double)gathers : 0.0;
    // Never chosen by AUTO: measured slower than the fused kernel's medium rows on every configuration (C1 10.7 vs 9.3 us,
    // C2 11.3 vs 7.2 us, C3 medium part 0.38 vs 0.21 ms, C4 1.14 vs 0.80 ms; profiles/r02/README.md) - 72-96 registers and
    // the window leave 16-22 % of the warp slots occupied, and the gathers become shared-memory wavefronts at the same
    // one-per-lane rate.  DASP_VARIANT_BANDED keeps it selectable.
    L.mb_auto = 0;
    L.s.medium_banded = L.mb_auto; L.s.medium_band_hit_rate = L.mb_hit_rate; L.s.medium_gather_lines = L.med_gather_lines;
    h->L.s.device_bytes = h->pool.bytes;
    return DASP_OK;
}

int build_short_bands(dasp_handle *h, cudaStream_t st, bool force)
{
    int rc = h->dtype == DASP_F16 ? build_short_bands_t<unsigned short>(h, st, force) : build_short_bands_t<double>(h, st, force);
    h->L.s.device_bytes = h->pool.bytes;
    return rc;
}

int relabel_columns(dasp_handle *h, const int *d_new_index, int n_new, cudaStream_t st)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    DevicePool &pool = h->pool;
    DevicePool tmp;
    struct Guard { DevicePool &p; ~Guard() { p.free_all(); } } guard{tmp};
    if (s.n > 0) { // every new index must lie inside the relabelled x
        int *flag = nullptr, bad = 0;
        DASP_TRY(tmp.alloc((void **)&flag, sizeof(int)));
        DASP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
        check_range<<<592, 256, 0, st>>>(d_new_index, s.n, n_new, flag);
        DASP_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        DASP_CUDA(cudaStreamSynchronize(st));
        if (bad) { set_error("dasp_relabel_columns: a new index lies outside [0, %d)", n_new); return DASP_ERR_INVALID; }
    }
    if (!L.relabelled) {
        int *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr;
        DASP_TRY(pool.alloc((void **)&a, sizeof(int) * (size_t)s.fill0_nnz_long));
        DASP_TRY(pool.alloc((void **)&b, sizeof(int) * (size_t)s.fill0_nnz_reg));
        DASP_TRY(pool.alloc((void **)&c, sizeof(int) * (size_t)s.nnz_irreg));
        DASP_TRY(pool.alloc((void **)&d, sizeof(int) * (size_t)s.fill0_nnz_short));
        L.k_long_cid = a; L.k_reg_cid = b; L.k_irreg_cid = c; L.k_short_cid = d;
        L.relabelled = 1;
    }
    const int G = 1184;
    if (h->dtype == DASP_F16) {
        using H = unsigned short;
        map_cid<H><<<G, 256, 0, st>>>(L.long_cid, (const H *)L.long_val, s.fill0_nnz_long, d_new_index, L.k_long_cid);
        map_cid<H><<<G, 256, 0, st>>>(L.reg_cid, (const H *)L.reg_val, s.fill0_nnz_reg, d_new_index, L.k_reg_cid);
        map_cid<H><<<G, 256, 0, st>>>(L.irreg_cid, (const H *)nullptr, s.nnz_irreg, d_new_index, L.k_irreg_cid);
        map_cid<H><<<G, 256, 0, st>>>(L.short_cid, (const H *)L.short_val, s.fill0_nnz_short, d_new_index, L.k_short_cid);
    } else {
        map_cid<double><<<G, 256, 0, st>>>(L.long_cid, (const double *)L.long_val, s.fill0_nnz_long, d_new_index, L.k_long_cid);
        map_cid<double><<<G, 256, 0, st>>>(L.reg_cid, (const double *)L.reg_val, s.fill0_nnz_reg, d_new_index, L.k_reg_cid);
        map_cid<double><<<G, 256, 0, st>>>(L.irreg_cid, (const double *)nullptr, s.nnz_irreg, d_new_index, L.k_irreg_cid);
        map_cid<double><<<G, 256, 0, st>>>(L.short_cid, (const double *)L.short_val, s.fill0_nnz_short, d_new_index, L.k_short_cid);
    }
    DASP_CUDA(cudaGetLastError());
    L.x_len = n_new;
    L.s.col_min = 0; L.s.col_max = n_new - 1; // the host paths upload the whole relabelled vector
    unsigned long long *lines = nullptr;
    DASP_TRY(tmp.alloc((void **)&lines, sizeof(unsigned long long)));
    DASP_CUDA(cudaMemsetAsync(lines, 0, sizeof(unsigned long long), st));
    DASP_TRY(derive_indices(h, st, lines));
    // the column-blocked copy depends on the labels: drop it and let the decision run again on the new ones
    const bool had_lcb = L.lcb_val != nullptr;
    if (had_lcb) {
        DASP_CUDA(cudaStreamSynchronize(st));
        pool.release(L.lcb_val); pool.release(L.lcb_idx); pool.release(L.lcb_blk_ptr); pool.release(L.lcb_cta_first);
        pool.release(L.lcb_acc); pool.release(L.lcb_done);
        pool.release(L.lcb_idx16); pool.release(L.lcb_chunk_row); pool.release(L.lcb_chunk_wide);
        L.lcb_idx16 = nullptr; L.lcb_chunk_row = nullptr; L.lcb_chunk_wide = nullptr;
        L.lcb_val = nullptr; L.lcb_idx = nullptr; L.lcb_blk_ptr = nullptr; L.lcb_cta_first = nullptr; L.lcb_acc = nullptr;
        L.lcb_done = nullptr; L.lcb_nctas = 0; L.lcb_live = 0;
    }
    DASP_TRY(decide_long_variant(h, st, lines));
    if (h->var_long == DASP_VARIANT_BLOCKED && !L.lcb_val) DASP_TRY(build_lcb(h, st));
    DASP_TRY(build_short_bands(h, st, h->var_short == DASP_VARIANT_BANDED));
    if (h->var_medium == DASP_VARIANT_BANDED) DASP_TRY(build_medium_bands(h, st)); // on demand only (never AUTO)
    else if (L.mb_lo) { pool.release(L.mb_lo); L.mb_lo = nullptr; }
    DASP_CUDA(cudaStreamSynchronize(st));
    h->L.s.device_bytes = h->pool.bytes;
    return DASP_OK;
}

int derive(dasp_handle *h, cudaStream_t st)
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    DevicePool &pool = h->pool;
    DevicePool tmp;
    struct Guard { DevicePool &p; ~Guard() { p.free_all(); } } guard{tmp};
    const int cl = s.row_long, cm = s.row_block, blocknum = s.blocknum, m = s.m;
    tmp.reserve(16 * ((size_t)blocknum / 4 + (size_t)m / 64 + 8192) + (4u << 20)); // sort buffers of the two work lists + radix / scan scratch
    L.k_long_cid = L.long_cid; L.k_reg_cid = L.reg_cid; L.k_irreg_cid = L.irreg_cid; L.k_short_cid = L.short_cid;
    L.relabelled = 0;
    L.x_len = s.n;

    // ---- long rows: work units (execution order), merge scratch, compact indices ----
    DASP_TRY(pool.alloc((void **)&L.long_unit_first, sizeof(int) * (size_t)(cl + 1)));
    DASP_CUDA(cudaMemsetAsync(L.long_unit_first, 0, sizeof(int) * (size_t)(cl + 1), st));
    unsigned long long *lines = nullptr;
    DASP_TRY(tmp.alloc((void **)&lines, sizeof(unsigned long long)));
    DASP_CUDA(cudaMemsetAsync(lines, 0, sizeof(unsigned long long), st));
    L.n_long_units = 0;
    if (cl > 0) {
        // A unit is what ONE warp walks serially.  A long part too small to give every SM a few dozen units of 32 reference
        // warps is cut finer (power of two), so that a small matrix with a handful of long rows is not one long dependent
        // chain per row: 1.1 M long entries in 655 rows took 20 us in units of 2048 entries.
        const int sms = h->sm_count > 0 ? h->sm_count : 148;
        int uw = LONG_UNIT_WARPS;
        while (uw > 1 && (long)s.warp_number / uw < 32L * sms) uw >>= 1;
        if (getenv("DASP_LONG_UNIT_WARPS")) uw = atoi(getenv("DASP_LONG_UNIT_WARPS")); // A/B aid
        if (uw < 1) uw = 1;
        if (uw > LONG_UNIT_WARPS) uw = LONG_UNIT_WARPS;
        L.long_unit_warps = uw;
        units_per_row<<<grid_for(cl, 256), 256, 0, st>>>(L.long_rpt_new, cl, uw, L.long_unit_first);
        DASP_TRY(scan_inplace(tmp, L.long_unit_first, cl + 1, st));
        DASP_CUDA(cudaMemcpyAsync(&L.n_long_units, L.long_unit_first + cl, sizeof(int), cudaMemcpyDeviceToHost, st));
        DASP_CUDA(cudaStreamSynchronize(st));
    }
    DASP_TRY(pool.alloc((void **)&L.long_unit_row, sizeof(int) * (size_t)L.n_long_units));
    DASP_TRY(pool.alloc((void **)&L.long_unit_chunk, sizeof(int) * (size_t)L.n_long_units));
    DASP_TRY(pool.alloc(&L.long_partial, 8 * (size_t)L.n_long_units));
    DASP_TRY(pool.alloc((void **)&L.long_done, sizeof(unsigned) * (size_t)cl));
    DASP_TRY(pool.alloc((void **)&L.long_cbase, sizeof(int) * (size_t)(s.fill0_nnz_long / 32)));
    DASP_TRY(pool.alloc((void **)&L.long_cdelta, sizeof(unsigned short) * (size_t)s.fill0_nnz_long));
    DASP_TRY(pool.alloc((void **)&L.long_wide, (size_t)L.n_long_units));
    DASP_CUDA(cudaMemsetAsync(L.long_done, 0, sizeof(unsigned) * (size_t)cl, st));
    if (cl > 0 && L.n_long_units > 0)
        fill_long_units<<<(cl + 7) / 8, 128, 0, st>>>(L.long_unit_first, cl, L.long_unit_row, L.long_unit_chunk);
    // ---- medium rows: irregular-tail flags, compact indices ----
    const int ngroups = ceil_div(cm, 32);
    DASP_TRY(pool.alloc((void **)&L.med_has_irreg, (size_t)ngroups));
    if (!L.reg_cbase) { // (dasp_create has them already, filled by pack_reg; dasp_load comes here without)
        DASP_TRY(pool.alloc((void **)&L.reg_cbase, sizeof(int) * (size_t)(s.fill0_nnz_reg / 32)));
        DASP_TRY(pool.alloc((void **)&L.reg_cdelta, sizeof(unsigned short) * (size_t)s.fill0_nnz_reg));
        DASP_TRY(pool.alloc((void **)&L.blk_wide, (size_t)blocknum));
        DASP_TRY(pool.alloc((void **)&L.blk_live, sizeof(unsigned short) * (size_t)blocknum));
    }
    if (cm > 0) flag_irreg<<<grid_for(ngroups, 256), 256, 0, st>>>(L.irreg_rpt, cm, ngroups, L.med_has_irreg);
    DASP_TRY(derive_indices(h, st, lines));
    // ---- locality-ordered work lists (see dasp_internal.h) ----
    {
        const int ngroups4 = blocknum / 4;
        DASP_TRY(pool.alloc((void **)&L.med_order, sizeof(int) * (size_t)ngroups4));
        if (ngroups4 > 0) {
            int *k0 = nullptr, *v0 = nullptr, *k1 = nullptr;
            DASP_TRY(tmp.alloc((void **)&k0, sizeof(int) * (size_t)ngroups4));
            DASP_TRY(tmp.alloc((void **)&v0, sizeof(int) * (size_t)ngroups4));
            DASP_TRY(tmp.alloc((void **)&k1, sizeof(int) * (size_t)ngroups4));
            med_group_keys<<<grid_for(ngroups4, 256), 256, 0, st>>>(L.order_rid, cl, cm, ngroups4, k0, v0);
            DASP_TRY(radix_sort_pairs(tmp, k0, v0, k1, L.med_order, ngroups4, 31, false, st));
            // When one ascending run already covers most of the groups (a stencil: one dominant length class, rows in
            // natural order) the identity order IS the local one and the indirection only costs: drop the list.
            int *desc = nullptr;
            DASP_TRY(tmp.alloc((void **)&desc, sizeof(int) * 4100));
            DASP_CUDA(cudaMemsetAsync(desc, 0, sizeof(int) * 4100, st));
            med_descents<<<grid_for(ngroups4, 256), 256, 0, st>>>(k0, ngroups4, desc);
            std::vector<int> hd(4100);
            DASP_CUDA(cudaMemcpyAsync(hd.data(), desc, sizeof(int) * 4100, cudaMemcpyDeviceToHost, st));
            DASP_CUDA(cudaStreamSynchronize(st));
            if (hd[0] <= 4096) {
                std::sort(hd.begin() + 1, hd.begin() + 1 + hd[0]);
                int longest = 0, prev = 0;
                for (int i = 1; i <= hd[0]; i++) { longest = std::max(longest, hd[i] - prev); prev = hd[i]; }
                longest = std::max(longest, ngroups4 - prev);
                if ((double)longest >= 0.9 * ngroups4) { pool.release(L.med_order); L.med_order = nullptr; }
            }
        }
        const int f16 = h->dtype == DASP_F16;
        const int G = f16 ? 32 : 8, warps = SPMV_CTA / 32;
        const int tiles13 = ceil_div(s.common_13, 8), tiles34 = ceil_div(s.short_row_34, 8);
        const int tiles22 = ceil_div(s.short_row_2, 2 * G) * (G / 8);
        ShortOrderGeom g;
        g.ctas[0] = ceil_div(ceil_div(s.short_row_1, 32 * SINGLES_PER_THREAD), warps);
        g.ctas[1] = ceil_div(ceil_div(tiles13, SHORT_TILES_PER_WARP), warps);
        g.ctas[2] = ceil_div(ceil_div(tiles34, SHORT_TILES_PER_WARP), warps);
        g.ctas[3] = ceil_div(ceil_div(tiles22, SHORT_TILES_PER_WARP), warps);
        g.rows[0] = warps * 32 * SINGLES_PER_THREAD;
        g.rows[1] = warps * SHORT_TILES_PER_WARP * 16; // 8 pairs = 16 y entries per tile
        g.rows[2] = warps * SHORT_TILES_PER_WARP * 8;
        g.rows[3] = warps * SHORT_TILES_PER_WARP * 16;
        const int ybase = cl + cm; // K11 placement (see launch_spmv)
        g.ybase[1] = ybase + (f16 ? 0 : s.short_row_1);
        g.ybase[2] = g.ybase[1] + 2 * s.common_13;
        g.ybase[3] = g.ybase[2] + s.short_row_34;
        g.ybase[0] = f16 ? g.ybase[3] + s.short_row_2 : ybase;
        g.pair_group = G;
        g.count[0] = s.short_row_1; g.count[1] = 2 * s.common_13; g.count[2] = s.short_row_34; g.count[3] = s.short_row_2;
        const int total = g.ctas[0] + g.ctas[1] + g.ctas[2] + g.ctas[3];
        L.short_map_n = total;
        for (int c = 0; c < 4; c++) L.short_ctas[c] = g.ctas[c];
        DASP_TRY(pool.alloc((void **)&L.short_map, sizeof(int) * (size_t)total));
        if (total > 0) {
            int *k0 = nullptr, *v0 = nullptr, *k1 = nullptr;
            DASP_TRY(tmp.alloc((void **)&k0, sizeof(int) * (size_t)total));
            DASP_TRY(tmp.alloc((void **)&v0, sizeof(int) * (size_t)total));
            DASP_TRY(tmp.alloc((void **)&k1, sizeof(int) * (size_t)total));
            short_cta_keys<<<grid_for(total, 256), 256, 0, st>>>(L.order_rid, g, total, k0, v0);
            DASP_TRY(radix_sort_pairs(tmp, k0, v0, k1, L.short_map, total, 31, false, st));
        }
    }
    // ---- SM-affine medium-row queues of the small-matrix kernels (spmv.cu): counters start at zero, the kernel resets them ----
    DASP_TRY(pool.alloc((void **)&L.smq_cnt, sizeof(int) * (SMQ_MAX + 1)));
    DASP_CUDA(cudaMemsetAsync(L.smq_cnt, 0, sizeof(int) * (SMQ_MAX + 1), st));
    // ---- inverse permutation (dasp_unpermute_to, relabelled mode) ----
    DASP_TRY(pool.alloc((void **)&L.inv_order, sizeof(int) * (size_t)m));
    if (m > 0) invert_order<<<grid_for(m, 256), 256, 0, st>>>(L.order_rid, m, L.inv_order);
    DASP_CUDA(cudaGetLastError());

    pool.close_slab(); // what follows may be rebuilt later (relabelled mode, variants on demand): allocated on its own
    // ---- scattered long rows: column-blocked copy ----
    DASP_TRY(decide_long_variant(h, st, lines));
    // ---- short rows by row band ----
    DASP_TRY(build_short_bands(h, st, h->var_short == DASP_VARIANT_BANDED));
    if (h->var_medium == DASP_VARIANT_BANDED) DASP_TRY(build_medium_bands(h, st)); // on demand only (never AUTO)
    else if (L.mb_lo) { pool.release(L.mb_lo); L.mb_lo = nullptr; }
    DASP_CUDA(cudaStreamSynchronize(st));
    DASP_CUDA(cudaGetLastError());
    return DASP_OK;
}

} // namespace dasp
