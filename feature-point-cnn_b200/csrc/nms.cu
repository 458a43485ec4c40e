// K3-K5 on the detect path: grid NMS, the exact parallel form of the reference's greedy sweep (python/src/nms.py:4-53
// with the threshold of python/src/netutils.py:59 and the border removal of netutils.py:95-99), followed by the
// descending sort (netutils.py:92-93) - two launches per batch.
//
// The reference visits candidates (heat >= thresh) by descending confidence; a live candidate is kept
// and kills every candidate of its (2r+1)^2 window; a killed candidate kills nothing.  Equivalent
// rounds: every undecided candidate that is the maximum of the undecided candidates in its window is
// kept; every undecided candidate inside the window of a new keeper is suppressed; repeat until none
// is undecided.  (Induction on the visiting order: the superiors of a window maximum are all decided
// and, had one been kept, the maximum would have been suppressed with it.)  Equal confidences are
// ordered by ascending pixel index, the oracle's tie rule.
//
// Round 0 (nms_round0_kernel) works on 64x64-pixel tiles with a 2r halo in shared memory.  On the detect path it
// reads the LOGITS (10x10 cells x 65 channels per tile) and computes the softmax / depth-to-space values itself,
// with the arithmetic of heatmap_kernel (python/src/superpoint.py:111-114, python/src/netutils.py:64-75): the
// full-resolution heatmap is never written unless the caller asks for it.  The tile becomes sortable keys, the
// candidates are compacted, the window scans run per candidate (early exit), keeper flags are row bitmasks and
// suppression is a shifted-word test; it emits the keepers, a bit-per-pixel mask of the still-undecided candidates,
// their keys (scattered into a dense side array) and their compact list.
// After it only a few percent of the candidates are left: nms_finish_kernel, ONE 1024-thread CTA per image, keeps
// them in registers (a thread per candidate scans its window through the bitmask and looks keys up only where a
// bit is set - the mask is copied to shared memory when it fits; keepers clear their window bits with atomics; two
// block barriers per round), then sorts the survivors (LSD radix sort, 8-bit digits, keys in registers, one
// scatter/gather through shared memory per pass, constant digits skipped; through global memory when there are
// more than 8192) and emits (x, y), confidence, count.
#include "kernels.h"
#include "softmax_cell.cuh"
#include "sortkey.cuh"

namespace spb200 {

constexpr int kN0TW = 64, kN0TH = 64;        // interior tile of round 0 (two mask words per row)
constexpr int kN0Threads = 512;
constexpr int kNmsMaxR = 8;
constexpr int kFinThreads = 1024;
constexpr int kFinRegEntries = 8;            // undecided candidates a thread keeps in registers
constexpr int kSortSmemKeys = 8192;          // survivors per image sorted in registers + shared memory (8 per thread)
constexpr int kFinMaskWords = 16384;         // undecided-bit mask words per image held in shared memory (64 KB)

// ------------------------------------------------------------------------------------------------
// Round 0
// ------------------------------------------------------------------------------------------------
// Dense and separable: a candidate is kept in round 0 when no key of its window is larger and no EQUAL key comes earlier
// in raster order.  With F = the row maximum over dx in [-R, R], Lf / Rt = the maxima left / right of the centre:
//     kept  <=>  Lf < key  and  Rt <= key  and  F(rows above) < key  and  F(rows below) <= key.
// One regular pass (a thread = four neighbouring pixels of a row: three 16-byte loads, ~30 max operations) leaves F of
// every loaded row in shared memory, a bit per pixel that passes the row test and a bit per candidate; the vertical test
// then runs only where a row bit is set (2R loads of F).  Keepers are bits; "suppressed" is the keeper mask dilated by R
// (shifted-word ORs); undecided = candidate and not kept and not suppressed, per 32-pixel word.  No candidate lists, no
// per-candidate window scans, no divergence on the candidate density.
// LOGITS: src = detector logits, channels last, `cell_stride` floats per cell (65 real), needs R <= 4 (one halo cell).
// inv (heatmap input only, may be null): [B][H/8][W/8], the value of a pixel is src * inv of its cell - the detector tail's
// fused softmax leaves exp(l) and the per-cell normaliser separately (halo_tc.cu).
template <int R>
struct N0Geom {
    static constexpr int LW = kN0TW + 4 * R, LH = kN0TH + 4 * R;      // loaded region (halo 2R)
    static constexpr int EW = kN0TW + 2 * R, EH = kN0TH + 2 * R;      // region where keepers are evaluated (halo R)
    static constexpr int KP = 4 * ((LW / 4 + 1) | 1);                 // key row pitch: 16-byte aligned rows, an odd number of 16-byte groups
    static constexpr int FP = (EW + 3) / 4 * 4;                       // row pitch of F
    static constexpr int BW = (EW + 31) / 32 + 1;                     // bit words per E row (+1 zero word: funnel reads of word j + 1)
    static constexpr int kMaxKeep = ((kN0TW + R) / (R + 1) + 1) * ((kN0TH + R) / (R + 1) + 1);   // keepers are > R apart
    static constexpr int kListWords = kN0TH * kN0TW + 2 * kMaxKeep + 2;   // undecided list + survivor keys: they reuse F
    static constexpr int kFWords = LH * FP > kListWords ? LH * FP : kListWords;
    static constexpr size_t kSmem = sizeof(unsigned) * ((size_t)LH * KP + kFWords + 4 * EH * BW) + 16;
};

template <int R, bool LOGITS>
__global__ void __launch_bounds__(kN0Threads)
nms_round0_kernel(const float* __restrict__ src, int cell_stride, int H, int W, float thresh, int border, int kcap,
                  unsigned long long* __restrict__ keys, int* __restrict__ counters, unsigned* __restrict__ mask,
                  int mask_w, unsigned* __restrict__ und, unsigned* __restrict__ ukey, const float* __restrict__ inv) {
    using G = N0Geom<R>;
    constexpr int LW = G::LW, LH = G::LH, EW = G::EW, EH = G::EH, KP = G::KP, FP = G::FP, BW = G::BW, kMaxKeep = G::kMaxKeep;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* s_key = reinterpret_cast<unsigned*>(smem_raw);   // [LH][KP]   sortable key, 0 = not a candidate
    unsigned* s_F = s_key + LH * KP;                           // [LH][FP]   row maximum over dx in [-R, R], E columns
    unsigned* s_cb = s_F + G::kFWords;                         // [EH][BW]   candidate bits, bit ex of row ey
    unsigned* s_ok = s_cb + EH * BW;                           // [EH][BW]   passes the row test
    unsigned* s_kb = s_ok + EH * BW;                           // [EH][BW]   keeper bits
    unsigned* s_hd = s_kb + EH * BW;                           // [EH][BW]   keeper bits dilated along x
    unsigned* s_und = s_F;                                     // [TH*TW]    undecided pixel indices (after the vertical test F is dead)
    unsigned long long* s_keep = reinterpret_cast<unsigned long long*>(s_F + kN0TH * kN0TW + (kN0TH * kN0TW & 1));   // [kMaxKeep] survivor keys
    __shared__ int s_nund, s_nkeep, s_base_und, s_base_keep;

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int ty0 = blockIdx.y * kN0TH, tx0 = blockIdx.x * kN0TW;
    pdl_trigger();
    if (tid == 0) { s_nund = 0; s_nkeep = 0; }
    for (int i = tid; i < 4 * EH * BW; i += kN0Threads) s_cb[i] = 0u;
    pdl_wait();                                                // logits / heatmap of the previous kernel; the lists it appends to

    // 1. keys of the loaded region
    if (LOGITS) {
        // eight lanes per cell (softmax_cell.cuh): lane j owns channels 8j..8j+7 = pixel row j of the cell, so its keys
        // are eight consecutive entries of one s_key row; the loads of all the thread's cells are issued first
        static_assert(!LOGITS || R <= 4, "one halo cell");
        const int Hc = H / 8, Wc = W / 8;
        const float* lb = src + (size_t)b * Hc * Wc * cell_stride;
        constexpr int kOff = 8 - 2 * R;                        // window origin inside the 10x10-cell region
        constexpr int kTasks = 100 * 8, kIters = (kTasks + kN0Threads - 1) / kN0Threads;
        float4 la[kIters], lc[kIters];
        float ld[kIters];
        bool in[kIters];
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int t = it * kN0Threads + tid, c = t >> 3, j = t & 7;
            const int gcy = (int)blockIdx.y * 8 - 1 + c / 10, gcx = (int)blockIdx.x * 8 - 1 + c % 10;
            in[it] = t < kTasks && gcy >= 0 && gcy < Hc && gcx >= 0 && gcx < Wc;
            la[it] = lc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            ld[it] = 0.f;
            if (in[it]) {
                const float* p = lb + (size_t)(gcy * Wc + gcx) * cell_stride;
                la[it] = __ldg(reinterpret_cast<const float4*>(p + 8 * j));
                lc[it] = __ldg(reinterpret_cast<const float4*>(p + 8 * j) + 1);
                ld[it] = __ldg(p + 64);
            }
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int t = it * kN0Threads + tid, c = t >> 3, j = t & 7;
            if (it * kN0Threads + (tid & ~31) >= kTasks) break;          // the whole warp is past the last task
            const float l[8] = {la[it].x, la[it].y, la[it].z, la[it].w, lc[it].x, lc[it].y, lc[it].z, lc[it].w};
            float h[8];
            softmax_cell_octet(l, ld[it], j, h);
            const int ly = (c / 10) * 8 + j - kOff, lx0 = (c % 10) * 8 - kOff;
            if (t < kTasks && ly >= 0 && ly < LH) {
                unsigned kk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) kk[k] = (in[it] && h[k] >= thresh) ? sortable_bits(h[k]) : 0u;
                if (kOff == 0) {                                        // R = 4: the ten cells of a row are the loaded row
                    uint4* d = reinterpret_cast<uint4*>(s_key + ly * KP + lx0);
                    d[0] = make_uint4(kk[0], kk[1], kk[2], kk[3]);
                    d[1] = make_uint4(kk[4], kk[5], kk[6], kk[7]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (lx0 + k >= 0 && lx0 + k < LW) s_key[ly * KP + lx0 + k] = kk[k];
                }
            }
        }
    } else {
        // four pixels (one 16-byte load) per thread and step, every load of the thread issued before the first use
        const float* hmap = src + (size_t)b * H * W;
        // per-cell normalisers of the cells the loaded region touches: into shared memory while the pixel loads are in flight
        constexpr int kIC = LW / 8 + 2;                        // cells per axis (the region need not start on a cell boundary)
        __shared__ float s_inv[kIC * kIC];
        const int cy0 = (ty0 - 2 * R) >> 3, cx0 = (tx0 - 2 * R) >> 3;         // arithmetic shifts: floor for the negative halo
        if (inv) {
            const int Hc = H >> 3, Wc = W >> 3;
            for (int i = tid; i < kIC * kIC; i += kN0Threads) {
                const int cy = cy0 + i / kIC, cx = cx0 + i % kIC;
                s_inv[i] = (cy >= 0 && cy < Hc && cx >= 0 && cx < Wc) ? __ldg(inv + ((size_t)b * Hc + cy) * Wc + cx) : 0.f;
            }
        }
        constexpr int LQ = LW / 4;                             // float4 groups per loaded row
        constexpr int kIters = (LH * LQ + kN0Threads - 1) / kN0Threads;
        // groups are then 16-byte aligned and never straddle the image edge
        const bool vec_ok = ((2 * R) % 4 == 0) && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(hmap) & 15) == 0);
        float4 v[kIters];
        int koff[kIters];                                      // s_key offset of the group, -1 past the end
        // (row, group) of the thread's it-th group without a division per step: i = it * kN0Threads + tid
        constexpr int kDq = kN0Threads / LQ, kDr = kN0Threads % LQ;
        const int q0 = tid / LQ, r0 = tid - q0 * LQ;
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            int ly = q0 + it * kDq, lq = r0 + it * kDr;
#pragma unroll
            for (int c = 0; c < (it * kDr + LQ - 1) / LQ; ++c)
                if (lq >= LQ) { lq -= LQ; ++ly; }
            const int lx = lq * 4;
            koff[it] = ly < LH ? ly * KP + lx : -1;
            if (ly < LH) {                                     // + the group's cell in s_inv (bits 16-23) and its column within the cell
                const int gy = ty0 - 2 * R + ly, gx = tx0 - 2 * R + lx;
                koff[it] |= ((((gy >> 3) - cy0) * kIC + ((gx >> 3) - cx0)) << 16) | ((gx & 7) << 24);
            }
            v[it] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);      // outside the image: never a candidate
            if (ly < LH) {
                const int gy = ty0 - 2 * R + ly, gx = tx0 - 2 * R + lx;
                if (gy >= 0 && gy < H) {
                    const float* p = hmap + ((size_t)gy * W + gx);
                    if (vec_ok) {
                        if (gx >= 0 && gx < W) v[it] = __ldg(reinterpret_cast<const float4*>(p));
                    } else {
                        if (gx >= 0 && gx < W) v[it].x = __ldg(p);
                        if (gx + 1 >= 0 && gx + 1 < W) v[it].y = __ldg(p + 1);
                        if (gx + 2 >= 0 && gx + 2 < W) v[it].z = __ldg(p + 2);
                        if (gx + 3 >= 0 && gx + 3 < W) v[it].w = __ldg(p + 3);
                    }
                }
            }
        }
        if (inv) {
            __syncthreads();
#pragma unroll
            for (int it = 0; it < kIters; ++it) {
                if (koff[it] < 0) continue;
                const float* ir = s_inv + ((koff[it] >> 16) & 255);
                if ((2 * R) % 4 == 0) {                        // groups of four never straddle a cell
                    const float sc = ir[0];
                    v[it].x *= sc; v[it].y *= sc; v[it].z *= sc; v[it].w *= sc;
                } else {
                    const int c = (koff[it] >> 24) & 7;
                    v[it].x *= ir[c >> 3]; v[it].y *= ir[(c + 1) >> 3]; v[it].z *= ir[(c + 2) >> 3]; v[it].w *= ir[(c + 3) >> 3];
                }
            }
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            if (koff[it] >= 0)
                *reinterpret_cast<uint4*>(s_key + (koff[it] & 0xffff)) =
                    make_uint4(v[it].x >= thresh ? sortable_bits(v[it].x) : 0u, v[it].y >= thresh ? sortable_bits(v[it].y) : 0u,
                               v[it].z >= thresh ? sortable_bits(v[it].z) : 0u, v[it].w >= thresh ? sortable_bits(v[it].w) : 0u);
        }
    }
    __syncthreads();

    // 2. row pass: F of every loaded row, candidate bits and row-test bits of the E rows
    {
        constexpr int TPR = FP / 4, NV = (4 + 2 * R + 3) / 4;
        for (int t = tid; t < LH * TPR; t += kN0Threads) {
            const int ly = t / TPR, ex0 = (t - ly * TPR) * 4;
            unsigned k[NV * 4];
            const uint4* kp = reinterpret_cast<const uint4*>(s_key + ly * KP + ex0);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const uint4 q = kp[v];
                k[4 * v] = q.x; k[4 * v + 1] = q.y; k[4 * v + 2] = q.z; k[4 * v + 3] = q.w;
            }
            unsigned f[4], cbits = 0u, okbits = 0u;
            unsigned m4[R == 4 ? 9 : 1];                       // R = 4: maxima of four neighbours by doubling, shared between the outputs
            if (R == 4) {
                unsigned m2[11];
#pragma unroll
                for (int i = 0; i < 11; ++i) m2[i] = max(k[i], k[i + 1]);
#pragma unroll
                for (int i = 0; i < 9; ++i) m4[R == 4 ? i : 0] = (i == 4) ? 0u : max(m2[i], m2[i + 2]);
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                unsigned lf = 0u, rt = 0u;
                if (R == 4) {
                    lf = m4[R == 4 ? o : 0]; rt = m4[R == 4 ? o + 5 : 0];
                } else {
#pragma unroll
                    for (int d = 0; d < R; ++d) { lf = max(lf, k[o + d]); rt = max(rt, k[o + R + 1 + d]); }
                }
                const unsigned c = k[o + R];
                f[o] = max(max(lf, rt), c);
                if (c != 0u && (FP == EW || ex0 + o < EW)) {
                    cbits |= 1u << o;
                    if (lf < c && rt <= c) okbits |= 1u << o;
                }
            }
            *reinterpret_cast<uint4*>(s_F + ly * FP + ex0) = make_uint4(f[0], f[1], f[2], f[3]);
            const int ey = ly - R;
            if (cbits && ey >= 0 && ey < EH) {
                atomicOr(&s_cb[ey * BW + (ex0 >> 5)], cbits << (ex0 & 31));
                if (okbits) atomicOr(&s_ok[ey * BW + (ex0 >> 5)], okbits << (ex0 & 31));
            }
        }
    }
    __syncthreads();

    // 3. vertical test where the row test passed: nothing as large above, nothing larger below
    {
        constexpr int CH = (EW + 15) / 16;                     // 16-pixel chunks per E row
        for (int t = tid; t < EH * CH; t += kN0Threads) {
            const int ey = t / CH, ch = t - ey * CH;
            unsigned bits = (s_ok[ey * BW + (ch >> 1)] >> ((ch & 1) * 16)) & 0xffffu;
            unsigned keep = 0u;
            while (bits) {
                const int bp = __ffs(bits) - 1;
                bits &= bits - 1u;
                const int ex = ch * 16 + bp;
                const unsigned c = s_key[(ey + R) * KP + ex + R];
                bool ok = true;
#pragma unroll
                for (int d = 1; d <= R; ++d) {
                    if (s_F[(ey + R - d) * FP + ex] >= c) ok = false;
                    if (s_F[(ey + R + d) * FP + ex] > c) ok = false;
                }
                if (ok) keep |= 1u << bp;
            }
            if (keep) atomicOr(&s_kb[ey * BW + (ch >> 1)], keep << ((ch & 1) * 16));
        }
    }
    __syncthreads();

    // 4. keeper bits dilated along x
    for (int t = tid; t < EH * (BW - 1); t += kN0Threads) {
        const int ey = t / (BW - 1), w = t - ey * (BW - 1);
        const unsigned cur = s_kb[ey * BW + w], prev = w ? s_kb[ey * BW + w - 1] : 0u, next = s_kb[ey * BW + w + 1];
        unsigned o = cur;
#pragma unroll
        for (int d = 1; d <= R; ++d) o |= (cur << d) | (prev >> (32 - d)) | (cur >> d) | (next << (32 - d));
        s_hd[ey * BW + w] = o;
    }
    __syncthreads();

    // 5. interior words: suppressed = dilated along y; undecided = candidate, not kept, not suppressed; lists
    int* cnt = counters + b * kNmsCounters;
    unsigned* uk = ukey + (size_t)b * H * W;
    unsigned* mrow = mask + (size_t)b * H * mask_w;
    for (int t = tid; t < kN0TH * 2 * 4; t += kN0Threads) {
        // four threads per 32-pixel word, each emits the list entries of eight pixels (the bit loops are the serial part)
        const int iy = t >> 3, j = (t >> 2) & 1, part = t & 3, ey = iy + R;
        // the interior starts at bit R of the E row: 32 bits from bit R + 32 j
        auto ext = [&](const unsigned* a, int row) { return __funnelshift_r(a[row * BW + j], a[row * BW + j + 1], R); };
        unsigned sup = 0u;
#pragma unroll
        for (int dy = -R; dy <= R; ++dy) sup |= ext(s_hd, ey + dy);
        const unsigned keep = ext(s_kb, ey);
        const unsigned undw = ext(s_cb, ey) & ~keep & ~sup;
        const int gy = ty0 + iy, wcol = (tx0 >> 5) + j;
        if (part == 0 && gy < H && wcol < mask_w) mrow[(size_t)gy * mask_w + wcol] = undw;
        const unsigned* krow = s_key + (ey + R) * KP + 2 * R + 32 * j;
        const unsigned slice = 0xffu << (8 * part);
        unsigned kb = keep & slice;
        while (kb) {
            const int bp = __ffs(kb) - 1;
            kb &= kb - 1u;
            const int gx = tx0 + 32 * j + bp;
            if (!(gx < border || gx >= W - border || gy < border || gy >= H - border)) {
                const int pos = atomicAdd(&s_nkeep, 1);
                if (pos < kMaxKeep) s_keep[pos] = survivor_key(krow[bp], (unsigned)(gy * W + gx));
            }
        }
        unsigned ub = undw & slice;
        while (ub) {
            const int bp = __ffs(ub) - 1;
            ub &= ub - 1u;
            const unsigned pix = (unsigned)(gy * W + tx0 + 32 * j + bp);
            s_und[atomicAdd(&s_nund, 1)] = pix;
            uk[pix] = krow[bp];
        }
    }
    __syncthreads();

    // 6. the two compact lists behind one atomicAdd each
    const int nund = s_nund, nkeep = min(s_nkeep, kMaxKeep);
    if (tid == 0) s_base_und = nund ? atomicAdd(cnt + 1, nund) : 0;          // two warps: the two round trips overlap
    if (tid == 32) s_base_keep = nkeep ? atomicAdd(cnt, nkeep) : 0;
    __syncthreads();
    unsigned* uout = und + (size_t)b * H * W + s_base_und;
    for (int i = tid; i < nund; i += kN0Threads) uout[i] = s_und[i];
    unsigned long long* kout = keys + (size_t)b * kcap;
    for (int i = tid; i < nkeep; i += kN0Threads)
        if (s_base_keep + i < kcap) kout[s_base_keep + i] = s_keep[i];
}

// ------------------------------------------------------------------------------------------------
// Rounds >= 1, sort, emit: one CTA per image
// ------------------------------------------------------------------------------------------------
constexpr unsigned kDead = 0xffffffffu;

template <bool SMASK>
__device__ __forceinline__ unsigned ld_mask(const unsigned* p) {
    return SMASK ? *reinterpret_cast<const volatile unsigned*>(p) : __ldcg(p);
}

// The usual case: candidates (and their keys) in registers, radius known at compile time; SMASK: the undecided-bit mask
// of the image is a copy in shared memory (else it is read from L2).
// The window rows are fetched first, then every neighbour key is loaded under its mask bit - up to (2R+1)^2
// independent predicated loads, so a round costs about one L2 round trip instead of one per undecided neighbour.
template <int R, bool SMASK>
__device__ __forceinline__ void finish_rounds_fast(int n0, int H, int W, int border, int kcap, const unsigned* __restrict__ list,
                                                   unsigned* s_mask, int mask_w, const unsigned* __restrict__ uk,
                                                   unsigned long long* __restrict__ kdst, int* s_n) {
    constexpr int S = 2 * R + 1;
    const int tid = threadIdx.x;
    unsigned ent[kFinRegEntries], kp[kFinRegEntries];
#pragma unroll
    for (int k = 0; k < kFinRegEntries; ++k) {
        const int idx = tid + k * kFinThreads;
        ent[k] = idx < n0 ? __ldcg(list + idx) : kDead;
    }
#pragma unroll
    for (int k = 0; k < kFinRegEntries; ++k) kp[k] = ent[k] != kDead ? __ldcg(uk + ent[k]) : 0u;
    for (int round = 1; round < (1 << 30); ++round) {
        unsigned keepers = 0u;
        bool live = false;
#pragma unroll
        for (int k = 0; k < kFinRegEntries; ++k) {
            if (ent[k] == kDead) continue;
            const unsigned p = ent[k];
            const int y = (int)(p / (unsigned)W), x = (int)(p - (unsigned)y * (unsigned)W);
            const int x0 = max(x - R, 0), x1 = min(x + R, W - 1);
            const int w0 = x0 >> 5, w1 = x1 >> 5, sh = x0 & 31;
            const unsigned long long wmask = (1ull << (x1 - x0 + 1)) - 1ull;
            unsigned rows[S];
#pragma unroll
            for (int d = 0; d < S; ++d) {
                const int qy = y + d - R;
                rows[d] = 0u;
                if (qy >= 0 && qy < H) {
                    const unsigned* mr = s_mask + (size_t)qy * mask_w;
                    unsigned long long two = (unsigned long long)ld_mask<SMASK>(mr + w0);
                    if (w1 != w0) two |= (unsigned long long)ld_mask<SMASK>(mr + w1) << 32;
                    rows[d] = (unsigned)((two >> sh) & wmask);
                }
            }
            if (!((rows[R] >> (x - x0)) & 1u)) {               // a keeper of the previous round cleared it
                ent[k] = kDead;
                continue;
            }
            rows[R] &= ~(1u << (x - x0));
            unsigned beaten = 0u;
#pragma unroll
            for (int d = 0; d < S; ++d) {
                if (rows[d] == 0u) continue;
                const unsigned qrow = (unsigned)((y + d - R) * W + x0);
                unsigned kq[S];
#pragma unroll
                for (int t = 0; t < S; ++t) kq[t] = 0u;
#pragma unroll
                for (int t = 0; t < S; ++t)
                    if ((rows[d] >> t) & 1u) kq[t] = __ldcg(uk + qrow + t);      // the row's loads are independent
#pragma unroll
                for (int t = 0; t < S; ++t)                                     // kq = 0 (bit clear) never beats: kp > 0
                    beaten |= (unsigned)(kq[t] > kp[k]) | ((unsigned)(kq[t] == kp[k]) & (unsigned)(qrow + t < p));
            }
            if (!beaten) keepers |= 1u << k; else live = true;
        }
        __syncthreads();                 // every test of the round precedes every clearing
#pragma unroll
        for (int k = 0; k < kFinRegEntries; ++k) {
            if (!((keepers >> k) & 1u)) continue;
            const unsigned p = ent[k];
            ent[k] = kDead;
            const int y = (int)(p / (unsigned)W), x = (int)(p - (unsigned)y * (unsigned)W);
            const int x0 = max(x - R, 0), x1 = min(x + R, W - 1);
            const int w0 = x0 >> 5, w1 = x1 >> 5;
            const unsigned lo = 0xffffffffu << (x0 & 31), hi = 0xffffffffu >> (31 - (x1 & 31));
            for (int qy = max(y - R, 0); qy <= min(y + R, H - 1); ++qy) {
                unsigned* qr = s_mask + (size_t)qy * mask_w;
                if (w0 == w1) atomicAnd(qr + w0, ~(lo & hi));
                else { atomicAnd(qr + w0, ~lo); atomicAnd(qr + w1, ~hi); }
            }
            if (!(x < border || x >= W - border || y < border || y >= H - border)) {
                const int pos = atomicAdd(s_n, 1);
                if (pos < kcap) kdst[pos] = survivor_key(kp[k], p);
            }
        }
        if (!SMASK) __threadfence();     // the cleared bits are read back from L2 in the next round
        if (!__syncthreads_or(live ? 1 : 0)) break;
    }
}

// REGS: the thread's candidates live in registers (n0 <= kFinRegEntries * kFinThreads); otherwise they are re-read
// from the list, where a decided entry is overwritten with kDead.  SMASK: the undecided-bit mask of the image is a
// copy in shared memory.  New survivors are appended to the image's keys.
template <bool REGS, bool SMASK>
__device__ __forceinline__ void finish_rounds(int n0, int H, int W, int r, int border, int kcap, unsigned* __restrict__ list,
                                              unsigned* __restrict__ mrow, int mask_w, const unsigned* __restrict__ uk,
                                              unsigned long long* __restrict__ kdst, int* s_n) {
    const int tid = threadIdx.x;
    const int per = REGS ? kFinRegEntries : (n0 + kFinThreads - 1) / kFinThreads;
    unsigned ent[kFinRegEntries];
    if (REGS) {
#pragma unroll
        for (int k = 0; k < kFinRegEntries; ++k) {
            const int idx = tid + k * kFinThreads;
            ent[k] = idx < n0 ? __ldcg(list + idx) : kDead;
        }
    }
    for (int round = 1; round < (1 << 30); ++round) {
        // ---- phase A: suppressed since the last round?  the maximum of the undecided candidates of its window? ----
        unsigned keepers = 0u;          // REGS: bit k = entry k is a new keeper
        bool live = false;
#pragma unroll
        for (int k = 0; k < kFinRegEntries; ++k) {
            if (!REGS && k > 0) break;
            for (int kk = 0; kk < (REGS ? 1 : per); ++kk) {
                const int idx = REGS ? 0 : tid + kk * kFinThreads;
                unsigned p = REGS ? ent[k] : (idx < n0 ? __ldcg(list + idx) : kDead);
                if (p == kDead) continue;
                p &= 0x7fffffffu;
                const int y = (int)(p / (unsigned)W), x = (int)(p - (unsigned)y * (unsigned)W);
                const int x0 = max(x - r, 0), x1 = min(x + r, W - 1);
                const int w0 = x0 >> 5, w1 = x1 >> 5, sh = x0 & 31;
                const unsigned long long wmask = (1ull << (x1 - x0 + 1)) - 1ull;
                const unsigned* mr = mrow + (size_t)y * mask_w;
                unsigned long long two = (unsigned long long)ld_mask<SMASK>(mr + w0);
                if (w1 != w0) two |= (unsigned long long)ld_mask<SMASK>(mr + w1) << 32;
                const unsigned own_row = (unsigned)((two >> sh) & wmask);
                if (!((own_row >> (x - x0)) & 1u)) {               // a keeper of the previous round cleared it
                    if (REGS) ent[k] = kDead; else list[idx] = kDead;
                    continue;
                }
                const unsigned kp = __ldcg(uk + p);
                bool beaten = false;
                for (int dy = -r; dy <= r; ++dy) {
                    const int qy = y + dy;
                    if (qy < 0 || qy >= H) continue;
                    unsigned bits = own_row & ~(1u << (x - x0));
                    if (dy != 0) {
                        const unsigned* qr = mrow + (size_t)qy * mask_w;
                        unsigned long long t2 = (unsigned long long)ld_mask<SMASK>(qr + w0);
                        if (w1 != w0) t2 |= (unsigned long long)ld_mask<SMASK>(qr + w1) << 32;
                        bits = (unsigned)((t2 >> sh) & wmask);
                    }
                    while (bits) {
                        const int t = __ffs(bits) - 1;
                        bits &= bits - 1u;
                        const unsigned q = (unsigned)(qy * W + x0 + t);
                        const unsigned kq = __ldcg(uk + q);
                        if (kq > kp || (kq == kp && q < p)) beaten = true;
                    }
                }
                if (!beaten) {
                    if (REGS) keepers |= 1u << k; else list[idx] = p | 0x80000000u;
                } else {
                    live = true;
                }
            }
        }
        __syncthreads();                 // every test of the round precedes every clearing
        // ---- phase B: new keepers clear their window in the mask and are emitted ----
#pragma unroll
        for (int k = 0; k < kFinRegEntries; ++k) {
            if (!REGS && k > 0) break;
            for (int kk = 0; kk < (REGS ? 1 : per); ++kk) {
                const int idx = REGS ? 0 : tid + kk * kFinThreads;
                unsigned p;
                if (REGS) {
                    if (!((keepers >> k) & 1u)) continue;
                    p = ent[k] & 0x7fffffffu;
                    ent[k] = kDead;
                } else {
                    if (idx >= n0) continue;
                    const unsigned e = list[idx];
                    if (e == kDead || !(e >> 31)) continue;
                    p = e & 0x7fffffffu;
                    list[idx] = kDead;
                }
                const int y = (int)(p / (unsigned)W), x = (int)(p - (unsigned)y * (unsigned)W);
                const int x0 = max(x - r, 0), x1 = min(x + r, W - 1);
                const int w0 = x0 >> 5, w1 = x1 >> 5;
                const unsigned lo = 0xffffffffu << (x0 & 31), hi = 0xffffffffu >> (31 - (x1 & 31));
                for (int qy = max(y - r, 0); qy <= min(y + r, H - 1); ++qy) {
                    unsigned* qr = mrow + (size_t)qy * mask_w;
                    if (w0 == w1) atomicAnd(qr + w0, ~(lo & hi));
                    else { atomicAnd(qr + w0, ~lo); atomicAnd(qr + w1, ~hi); }
                }
                if (!(x < border || x >= W - border || y < border || y >= H - border)) {
                    const int pos = atomicAdd(s_n, 1);
                    if (pos < kcap) kdst[pos] = survivor_key(__ldcg(uk + p), p);
                }
            }
        }
        __threadfence();
        if (!__syncthreads_or(live ? 1 : 0)) break;
    }
}

__global__ void __launch_bounds__(kFinThreads, 1)
nms_finish_kernel(int H, int W, int r, int border, int kcap, unsigned long long* __restrict__ keys,
                  unsigned long long* __restrict__ keys_alt, int* __restrict__ counters, unsigned* __restrict__ mask,
                  int mask_w, unsigned* __restrict__ und, const unsigned* __restrict__ ukey, int cap_out, int top_k,
                  int* __restrict__ count, int* __restrict__ xy, float* __restrict__ conf) {
    extern __shared__ __align__(16) unsigned char fin_smem[];
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(fin_smem);       // [kSortSmemKeys], after the rounds
    unsigned* s_mask = reinterpret_cast<unsigned*>(fin_smem);                          // [H][mask_w], during the rounds
    __shared__ unsigned s_hist[8][256];
    __shared__ unsigned s_base[256];
    __shared__ unsigned s_wcnt[32][256];
    __shared__ unsigned s_warp_tot[8];
    __shared__ int s_n;
    __shared__ unsigned s_orand[4];
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int b = blockIdx.x;
    int* cnt = counters + b * kNmsCounters;
    pdl_trigger();
    pdl_wait();
    const int nkeep0 = min(__ldcg(cnt), kcap);
    const int n0 = (int)min((long)__ldcg(cnt + 1), (long)H * W);
    // the counters are left at zero for the next call's round 0 (no memset node between the detector head and round 0)
    __syncthreads();
    if (tid < kNmsCounters) cnt[tid] = 0;
    unsigned long long* gkeys = keys + (size_t)b * kcap;
    if (tid == 0) s_n = nkeep0;

    if (n0 > 0) {
        unsigned* list = und + (size_t)b * H * W;
        unsigned* mrow = mask + (size_t)b * H * mask_w;
        const unsigned* uk = ukey + (size_t)b * H * W;
        const bool regs = n0 <= kFinRegEntries * kFinThreads;
        if (regs && H * mask_w <= kFinMaskWords) {
            for (int i = tid; i < H * mask_w; i += kFinThreads) s_mask[i] = __ldcg(mrow + i);
            __syncthreads();
            if (r == 4) finish_rounds_fast<4, true>(n0, H, W, border, kcap, list, s_mask, mask_w, uk, gkeys, &s_n);
            else finish_rounds<true, true>(n0, H, W, r, border, kcap, list, s_mask, mask_w, uk, gkeys, &s_n);
        } else {
            __syncthreads();
            if (regs && r == 4) finish_rounds_fast<4, false>(n0, H, W, border, kcap, list, mrow, mask_w, uk, gkeys, &s_n);
            else if (regs) finish_rounds<true, false>(n0, H, W, r, border, kcap, list, mrow, mask_w, uk, gkeys, &s_n);
            else finish_rounds<false, false>(n0, H, W, r, border, kcap, list, mrow, mask_w, uk, gkeys, &s_n);
        }
    }
    __syncthreads();
    int n = min(s_n, kcap);
    unsigned long long* src = gkeys;
    bool inverted = false;

    // ---- top-k with very many survivors (1080p): keep only the keys that can reach the first top_k places ----
    // Two histogram levels over the leading 13 + 13 bits of the inverted key (sign, exponent and 17 mantissa bits of the
    // confidence): at each level the bin in which the running count from the best key reaches the remaining quota
    // is the cut; keys before the cut prefix, and keys on it (ties included), are compacted into the second key buffer
    // and sorted by the usual path.
    if (top_k > 0 && n > kSortSmemKeys && top_k <= kSortSmemKeys / 2) {
        unsigned* hist = &s_wcnt[0][0];                          // 8192 bins
        __shared__ int s_cut, s_sel;
        unsigned prefix = 0u;                                    // cut bins of the levels done so far
        int quota = top_k, above = 0, sel = 0;                   // `above` keys are better than every key of the cut prefix
        bool two = false;
        for (int level = 0; level < 2; ++level) {
            const int sh_hi = 51 - 13 * level;                   // this level's 13 bits start here
            for (int i = tid; i < 8192; i += kFinThreads) hist[i] = 0u;
            if (tid == 0) { s_cut = 8191; s_sel = 0; }
            __syncthreads();
            for (int i = tid; i < n; i += kFinThreads) {
                const unsigned long long k = ~__ldcg(gkeys + i);
                if (level == 0 || (unsigned)(k >> 51) == prefix) atomicAdd(&hist[(unsigned)(k >> sh_hi) & 8191u], 1u);
            }
            __syncthreads();
            unsigned loc[8], sum = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = hist[tid * 8 + j]; sum += loc[j]; }
            unsigned inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_base[warp] = inc;
            __syncthreads();
            unsigned before = inc - sum;
            for (int w = 0; w < warp; ++w) before += s_base[w];
            if (before < (unsigned)quota && before + sum >= (unsigned)quota) {      // exactly one thread
                unsigned c = before;
                for (int j = 0; j < 8; ++j) {
                    if (c + loc[j] >= (unsigned)quota) { s_cut = tid * 8 + j; s_sel = (int)c; break; }
                    c += loc[j];
                }
            }
            __syncthreads();
            const int cut = s_cut, better = s_sel;               // keys of this level strictly before the cut bin
            const int in_cut = (int)hist[cut];
            __syncthreads();
            if (level == 0) prefix = (unsigned)cut; else { prefix = (prefix << 13) | (unsigned)cut; two = true; }
            above += better;
            quota -= better;
            sel = above + in_cut;
            if (sel <= kSortSmemKeys) break;                     // fits: no finer cut needed
        }
        if (sel > 0 && sel <= kSortSmemKeys) {
            unsigned long long* alt = keys_alt + (size_t)b * kcap;
            const int cmp_shift = two ? 38 : 51;                 // compare the leading 26 or 13 bits with the cut prefix
            if (tid == 0) s_sel = 0;
            __syncthreads();
            for (int i0 = 0; i0 < n; i0 += kFinThreads) {
                const int i = i0 + tid;
                const unsigned long long k = i < n ? __ldcg(gkeys + i) : 0ull;
                bool take = false;
                if (i < n) take = (unsigned)((~k) >> cmp_shift) <= prefix;
                const unsigned bal = __ballot_sync(0xffffffffu, take);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_sel, __popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (take) alt[base + __popc(bal & ((1u << lane) - 1u))] = k;
                }
            }
            __syncthreads();
            n = sel;
            gkeys = alt;
            src = alt;
        }
        __syncthreads();
    }

    if (n <= kSortSmemKeys) {
        // ---- the usual case: LSD radix sort with the keys in registers (striped: warp w owns positions
        //      [256 w, 256 w + 256), item i of lane l is position 256 w + 32 i + l, so walking the items in order is
        //      stable), ranks from match_any + per-warp digit counters, one scatter and one gather through shared
        //      memory per pass; only the confidence bytes are sorted (constant ones skipped), a fix-up orders the rare
        //      runs of equal confidence by pixel index.  Keys are held inverted (ascending sort of ~key = descending
        //      sort of key).
        constexpr int IPT = kSortSmemKeys / kFinThreads;
        // a digit position is constant when the OR and the AND of all keys agree on it
        if (tid < 4) s_orand[tid] = tid < 2 ? 0u : 0xffffffffu;
        __syncthreads();
        unsigned long long key[IPT];
        unsigned or_lo = 0u, or_hi = 0u, and_lo = 0xffffffffu, and_hi = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const int pos = warp * (32 * IPT) + i * 32 + lane;
            key[i] = pos < n ? ~__ldcg(gkeys + pos) : ~0ull;
            if (pos < n) {
                or_lo |= (unsigned)key[i]; or_hi |= (unsigned)(key[i] >> 32);
                and_lo &= (unsigned)key[i]; and_hi &= (unsigned)(key[i] >> 32);
            }
        }
        or_lo = __reduce_or_sync(0xffffffffu, or_lo); or_hi = __reduce_or_sync(0xffffffffu, or_hi);
        and_lo = __reduce_and_sync(0xffffffffu, and_lo); and_hi = __reduce_and_sync(0xffffffffu, and_hi);
        if (lane == 0) {
            atomicOr(&s_orand[0], or_lo); atomicOr(&s_orand[1], or_hi);
            atomicAnd(&s_orand[2], and_lo); atomicAnd(&s_orand[3], and_hi);
        }
        __syncthreads();
        const unsigned long long varying = (((unsigned long long)s_orand[1] << 32) | s_orand[0]) ^
                                           (((unsigned long long)s_orand[3] << 32) | s_orand[2]);
        bool in_smem = false;
        for (int pass = 4; pass < 8 && n > 1; ++pass) {
            if (!((varying >> (8 * pass)) & 255ull)) continue;
#pragma unroll
            for (int i = 0; i < 8; ++i) (&s_wcnt[0][0])[tid + i * kFinThreads] = 0u;
            __syncthreads();
            const int shift = 8 * pass;
            unsigned rank[IPT];
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const int pos = warp * (32 * IPT) + i * 32 + lane;
                if (pos - lane >= n) break;                         // the rest of the warp's segment is empty
                const bool valid = pos < n;
                const unsigned dig = (unsigned)(key[i] >> shift) & 255u;
                // lanes holding the same digit: eight ballots (MATCH.ANY has a fraction of the ballot throughput)
                unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) {
                    const unsigned bal = __ballot_sync(0xffffffffu, (dig >> bit) & 1u);
                    peers &= ((dig >> bit) & 1u) ? bal : ~bal;
                }
                const unsigned before = __popc(peers & ((1u << lane) - 1u));
                unsigned base = 0u;
                if (valid) base = s_wcnt[warp][dig];
                __syncwarp();
                if (valid && before == 0u) s_wcnt[warp][dig] = base + __popc(peers);
                __syncwarp();
                rank[i] = base + before;
            }
            __syncthreads();
            // digit totals -> exclusive offsets: s_wcnt[w][d] becomes the offset of warp w inside digit d, s_base[d] the
            // offset of digit d
            if (tid < 256) {
                unsigned total = 0u;
#pragma unroll 8
                for (int w = 0; w < 32; ++w) {
                    const unsigned c = s_wcnt[w][tid];
                    s_wcnt[w][tid] = total;
                    total += c;
                }
                unsigned inc = total;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) s_warp_tot[warp] = inc;
                s_base[tid] = inc - total;
            }
            __syncthreads();
            if (tid < 256) {
                unsigned add = 0u;
                for (int w = 0; w < warp; ++w) add += s_warp_tot[w];
                s_base[tid] += add;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const int pos = warp * (32 * IPT) + i * 32 + lane;
                if (pos < n) {
                    const unsigned dig = (unsigned)(key[i] >> shift) & 255u;
                    s_keys[s_base[dig] + s_wcnt[warp][dig] + rank[i]] = key[i];
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const int pos = warp * (32 * IPT) + i * 32 + lane;
                key[i] = pos < n ? s_keys[pos] : ~0ull;
            }
            in_smem = true;
        }
        if (!in_smem) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const int pos = warp * (32 * IPT) + i * 32 + lane;
                if (pos < n) s_keys[pos] = key[i];
            }
        }
        __syncthreads();
        // runs of equal confidence (flat image regions give them) are still in arrival order: every element of a run
        // counts the smaller keys of its run and moves to that rank - ascending inverted key = ascending pixel index
        {
            int newpos[IPT];
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const int i = tid + k * kFinThreads;
                newpos[k] = -1;
                if (i >= n) continue;
                const unsigned long long v = s_keys[i];
                const unsigned c = (unsigned)(v >> 32);
                const bool left = i > 0 && (unsigned)(s_keys[i - 1] >> 32) == c;
                const bool right = i + 1 < n && (unsigned)(s_keys[i + 1] >> 32) == c;
                if (!left && !right) continue;
                key[k] = v;
                int lo = i, hi = i, rank = 0;
                while (lo > 0 && (unsigned)(s_keys[lo - 1] >> 32) == c) { --lo; ++rank; rank -= s_keys[lo] > v ? 1 : 0; }
                while (hi + 1 < n && (unsigned)(s_keys[hi + 1] >> 32) == c) { ++hi; rank += s_keys[hi] < v ? 1 : 0; }
                newpos[k] = lo + rank;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < IPT; ++k)
                if (newpos[k] >= 0) s_keys[newpos[k]] = key[k];
        }
        __syncthreads();
        src = s_keys;
        inverted = true;

    } else {
    // ---- very many survivors: LSD radix sort through global memory ----
    unsigned long long* dst = keys_alt + (size_t)b * kcap;            // (the preselection above never leads here: it only swaps buffers when the result fits)
    for (int i = tid; i < 8 * 256; i += kFinThreads) (&s_hist[0][0])[i] = 0;
    for (int i = tid; i < 32 * 256; i += kFinThreads) (&s_wcnt[0][0])[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kFinThreads) {
        const unsigned long long v = ~src[i];          // ascending on ~key == descending on key
#pragma unroll
        for (int d = 0; d < 8; ++d) atomicAdd(&s_hist[d][(unsigned)(v >> (8 * d)) & 255u], 1u);
    }
    __syncthreads();
    for (int pass = 0; pass < 8 && n > 1; ++pass) {
        // skip a digit position on which every key agrees (block-uniform decision)
        const int hit = (tid < 256 && s_hist[pass][tid] == (unsigned)n) ? 1 : 0;
        if (__syncthreads_or(hit)) continue;
        // exclusive scan of the 256-bin histogram
        if (tid < 256) {
            const unsigned v = s_hist[pass][tid];
            unsigned inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_warp_tot[warp] = inc;
            s_base[tid] = inc - v;
        }
        __syncthreads();
        if (tid < 256) {
            unsigned add = 0;
            for (int w = 0; w < warp; ++w) add += s_warp_tot[w];
            s_base[tid] += add;
        }
        __syncthreads();
        const int shift = 8 * pass;
        for (int r0 = 0; r0 < n; r0 += kFinThreads) {
            const int i = r0 + tid;
            const bool valid = i < n;
            const unsigned long long key = valid ? src[i] : 0ull;
            const int dig = valid ? (int)((unsigned)((~key) >> shift) & 255u) : 256 + lane;
            const unsigned peers = __match_any_sync(0xffffffffu, dig);
            const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank_in_warp == 0) s_wcnt[warp][dig] = __popc(peers);
            __syncthreads();
            if (tid < 256) {
                unsigned off = s_base[tid];
                for (int w = 0; w < 32; ++w) {
                    const unsigned c = s_wcnt[w][tid];
                    if (c) {                   // untouched entries must stay 0: only a group's leader resets its entry
                        s_wcnt[w][tid] = off;
                        off += c;
                    }
                }
                s_base[tid] = off;
            }
            __syncthreads();
            if (valid) dst[s_wcnt[warp][dig] + rank_in_warp] = key;
            __syncwarp();
            if (valid && rank_in_warp == 0) s_wcnt[warp][dig] = 0;
            __syncthreads();
        }
        unsigned long long* t = src; src = dst; dst = t;
        __syncthreads();
    }

    }

    int nout = min(n, cap_out);
    if (top_k > 0) nout = min(nout, top_k);
    for (int i = tid; i < nout; i += kFinThreads) {
        const unsigned long long key = inverted ? ~src[i] : src[i];
        const unsigned pix = ~(unsigned)(key & 0xffffffffull);
        xy[((size_t)b * cap_out + i) * 2 + 0] = (int)(pix % (unsigned)W);
        xy[((size_t)b * cap_out + i) * 2 + 1] = (int)(pix / (unsigned)W);
        conf[(size_t)b * cap_out + i] = from_sortable_bits((unsigned)(key >> 32));
    }
    if (tid == 0) count[b] = nout;
}

template <int R, bool LOGITS>
static void launch_round0_t(const float* src, int cell_stride, int B, int H, int W, float thresh, int border,
                            const NmsWorkspace& ws, cudaStream_t st, const float* inv = nullptr) {
    const size_t smem = N0Geom<R>::kSmem;
    auto kern = nms_round0_kernel<R, LOGITS>;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + kN0TW - 1) / kN0TW, (H + kN0TH - 1) / kN0TH, B);
    launch_pdl(kern, grid, dim3(kN0Threads), smem, st, src, cell_stride, H, W, thresh, border, ws.kcap, ws.keys, ws.counters, ws.mask,
               ws.mask_w, ws.und, ws.ukey, inv);
}

bool nms_logits_supported(int radius) { return radius >= 0 && radius <= 4; }

void launch_nms_round0(const float* heat, const float* logits, int cell_stride, int B, int H, int W, float thresh, int radius,
                       int border, const NmsWorkspace& ws, bool zero_counters, cudaStream_t st, const float* heat_inv) {
    if (heat_inv && (!heat || H % 8 || W % 8)) throw std::invalid_argument("nms: a normaliser map needs a heatmap of whole cells");
    if (radius < 0 || radius > kNmsMaxR) throw std::invalid_argument("nms_dist must be in [0, 8]");
    if ((long)H * W >= (1l << 31)) throw std::invalid_argument("image too large");
    if (!heat && !(logits && nms_logits_supported(radius))) throw std::invalid_argument("nms: no heatmap given");
    if (zero_counters) SPB_CUDA(cudaMemsetAsync(ws.counters, 0, sizeof(int) * kNmsCounters * B, st));
    if (!heat) {
        switch (radius) {
            case 0: launch_round0_t<0, true>(logits, cell_stride, B, H, W, thresh, border, ws, st); break;
            case 1: launch_round0_t<1, true>(logits, cell_stride, B, H, W, thresh, border, ws, st); break;
            case 2: launch_round0_t<2, true>(logits, cell_stride, B, H, W, thresh, border, ws, st); break;
            case 3: launch_round0_t<3, true>(logits, cell_stride, B, H, W, thresh, border, ws, st); break;
            default: launch_round0_t<4, true>(logits, cell_stride, B, H, W, thresh, border, ws, st); break;
        }
    } else {
        switch (radius) {
            case 0: launch_round0_t<0, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            case 1: launch_round0_t<1, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            case 2: launch_round0_t<2, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            case 3: launch_round0_t<3, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            case 4: launch_round0_t<4, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            case 5: launch_round0_t<5, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            case 6: launch_round0_t<6, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            case 7: launch_round0_t<7, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
            default: launch_round0_t<8, false>(heat, 0, B, H, W, thresh, border, ws, st, heat_inv); break;
        }
    }
}

void launch_nms_finish(int B, int H, int W, int radius, int border, int top_k, int cap, const NmsWorkspace& ws, int* count,
                       int* xy, float* conf, cudaStream_t st) {
    const size_t smem = sizeof(unsigned long long) * kSortSmemKeys;
    SPB_CUDA(cudaFuncSetAttribute(nms_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(nms_finish_kernel, dim3(B), dim3(kFinThreads), smem, st, H, W, radius, border, ws.kcap, ws.keys, ws.keys_alt, ws.counters,
               ws.mask, ws.mask_w, ws.und, (const unsigned*)ws.ukey, cap, top_k, count, xy, conf);
}

}  // namespace spb200
