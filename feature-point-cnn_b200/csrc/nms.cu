// K4: grid NMS, the exact parallel form of the reference's greedy sweep (python/src/nms.py:4-53 with the
// threshold of python/src/netutils.py:59 and the border removal of netutils.py:95-99).
//
// The reference visits candidates (heat >= thresh) by descending confidence; a live candidate is kept
// and kills every candidate of its (2r+1)^2 window; a killed candidate kills nothing.  Equivalent
// rounds: every undecided candidate that is the maximum of the undecided candidates in its window is
// kept; every undecided candidate inside the window of a new keeper is suppressed; repeat until none
// is undecided.  (Induction on the visiting order: the superiors of a window maximum are all decided
// and, had one been kept, the maximum would have been suppressed with it.)  Equal confidences are
// ordered by ascending pixel index, the oracle's tie rule.
//
// Round 0 (nms_round0_kernel) works on 64x64-pixel tiles with a 2r halo in shared memory: one dense pass
// thresholds the tile into sortable keys and compacts the candidates, then the window scans run per
// candidate (early exit), keeper flags are row bitmasks and suppression is a shifted-word test; it emits
// the keepers, a bit-per-pixel mask of the still-undecided candidates and their compact list.  After it
// only a few percent of the candidates are left, so the remaining rounds
// (nms_rounds_kernel, one 8-CTA cluster per image) work on the compact list: a warp per candidate
// scans its window through the bitmask (heat is read only where a bit is set), keepers clear their
// window bits with atomics, the list is compacted, two cluster barriers per round.
#include <cooperative_groups.h>

#include "kernels.h"
#include "sortkey.cuh"

namespace cg = cooperative_groups;

namespace spb200 {

constexpr int kN0TW = 64, kN0TH = 64;        // interior tile of round 0 (two mask words per row)
constexpr int kN0Threads = 256;
constexpr int kNmsMaxR = 8;
constexpr int kNmsCluster = 8;
constexpr int kRoundsThreads = 256;

// ------------------------------------------------------------------------------------------------
// Round 0
// ------------------------------------------------------------------------------------------------
// Candidate-centric: only a few percent of the pixels pass the threshold, so after one dense pass that turns the
// tile (+ 2R halo) into sortable keys and compacts the candidates of the evaluation region (interior + R), the
// window scans run per CANDIDATE with early exit on the first stronger neighbour, keepers become bits, and a
// candidate is suppressed when a keeper bit lies in its window (nine shifted word tests).
template <int R>
__global__ void __launch_bounds__(kN0Threads)
nms_round0_kernel(const float* __restrict__ heat, int H, int W, float thresh, int border, int kcap,
                  unsigned long long* __restrict__ keys, int* __restrict__ counters, unsigned* __restrict__ mask,
                  int mask_w, unsigned* __restrict__ und) {
    constexpr int LW = kN0TW + 4 * R, LH = kN0TH + 4 * R;      // loaded region (halo 2R)
    constexpr int EW = kN0TW + 2 * R, EH = kN0TH + 2 * R;      // region where keepers are evaluated (halo R)
    constexpr int KW = (EW + 31) / 32 + 1;                     // keeper bit words per E row (+1 so a 64-bit window read stays inside)
    constexpr int kMaxKeep = ((kN0TW + R) / (R + 1) + 1) * ((kN0TH + R) / (R + 1) + 1);   // keepers are > R apart
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* s_key = reinterpret_cast<unsigned*>(smem_raw);   // [LH][LW]   sortable key, 0 = not a candidate
    unsigned* s_kb = s_key + LH * LW;                          // [EH][KW]   keeper bits, bit ex of row ey
    unsigned* s_ub = s_kb + EH * KW;                           // [TH][2]    undecided bits of the interior
    unsigned* s_und = s_ub + kN0TH * 2;                        // [TH*TW]    undecided pixel indices
    unsigned long long* s_keep = reinterpret_cast<unsigned long long*>(
        (reinterpret_cast<uintptr_t>(s_und + kN0TH * kN0TW) + 7) & ~(uintptr_t)7);                 // [kMaxKeep] survivor keys
    unsigned short* s_cand = reinterpret_cast<unsigned short*>(s_keep + kMaxKeep);               // [EH*EW] ey << 8 | ex
    __shared__ int s_ncand, s_nund, s_nkeep, s_base_und, s_base_keep;

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z;
    const int ty0 = blockIdx.y * kN0TH, tx0 = blockIdx.x * kN0TW;
    const float* hmap = heat + (size_t)b * H * W;
    if (tid == 0) { s_ncand = 0; s_nund = 0; s_nkeep = 0; }
    for (int i = tid; i < EH * KW; i += kN0Threads) s_kb[i] = 0u;
    for (int i = tid; i < kN0TH * 2; i += kN0Threads) s_ub[i] = 0u;
    __syncthreads();

    // 1. keys of the loaded region, four pixels (one 16-byte load) per thread and step; candidates of the evaluation
    //    region are appended to the list through a warp prefix sum (one shared atomic per warp and step)
    constexpr int LQ = (LW + 3) / 4;                           // float4 groups per loaded row
    // groups are then 16-byte aligned and never straddle the image edge
    const bool vec_ok = ((2 * R) % 4 == 0) && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(hmap) & 15) == 0);
    for (int i0 = 0; i0 < LH * LQ; i0 += kN0Threads) {
        const int i = i0 + tid;
        unsigned key[4] = {0u, 0u, 0u, 0u};
        int ly = 0, lx = 0;
        if (i < LH * LQ) {
            ly = i / LQ; lx = (i % LQ) * 4;
            const int gy = ty0 - 2 * R + ly, gx = tx0 - 2 * R + lx;
            if (gy >= 0 && gy < H) {
                const float* src = hmap + (size_t)gy * W + gx;
                float h[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // outside the image: never a candidate
                if (vec_ok) {
                    if (gx >= 0 && gx < W) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                        h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (gx + e >= 0 && gx + e < W) h[e] = __ldg(src + e);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (h[e] >= thresh) key[e] = sortable_bits(h[e]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (lx + e < LW) s_key[ly * LW + lx + e] = key[e];
        }
        unsigned flags = 0u;
        if (ly >= R && ly < R + EH) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (key[e] != 0u && lx + e >= R && lx + e < R + EW) flags |= 1u << e;
        }
        const int mine = __popc(flags);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            int base = 0;
            if (lane == 31) base = atomicAdd(&s_ncand, total);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (flags & (1u << e)) s_cand[base++] = (unsigned short)(((ly - R) << 8) | (lx + e - R));
        }
    }
    __syncthreads();
    const int ncand = s_ncand;

    // 2. keeper test per candidate: no stronger key in the window, no equal key earlier in raster order
    for (int c = tid; c < ncand; c += kN0Threads) {
        const int ey = s_cand[c] >> 8, ex = s_cand[c] & 255;
        const unsigned* centre = s_key + (ey + R) * LW + ex + R;
        const unsigned me = *centre;
        bool keep = true;
        for (int dy = -R; dy <= R && keep; ++dy) {
            const unsigned* row = centre + dy * LW;
#pragma unroll
            for (int dx = -R; dx <= R; ++dx) {
                const unsigned v = row[dx];
                if (v > me || (v == me && (dy < 0 || (dy == 0 && dx < 0)))) keep = false;
            }
        }
        if (keep) atomicOr(&s_kb[ey * KW + (ex >> 5)], 1u << (ex & 31));
    }
    __syncthreads();

    // 3. interior candidates: kept / suppressed by a keeper in the window / still undecided
    int* cnt = counters + b * kNmsCounters;
    for (int c = tid; c < ncand; c += kN0Threads) {
        const int ey = s_cand[c] >> 8, ex = s_cand[c] & 255;
        const int iy = ey - R, ix = ex - R;
        if (iy < 0 || iy >= kN0TH || ix < 0 || ix >= kN0TW) continue;
        const int gy = ty0 + iy, gx = tx0 + ix;
        const bool keep = (s_kb[ey * KW + (ex >> 5)] >> (ex & 31)) & 1u;
        const unsigned pix = (unsigned)(gy * W + gx);
        if (keep) {
            if (!(gx < border || gx >= W - border || gy < border || gy >= H - border)) {
                const int pos = atomicAdd(&s_nkeep, 1);
                if (pos < kMaxKeep) s_keep[pos] = survivor_key(s_key[(ey + R) * LW + ex + R], pix);
            }
            continue;
        }
        // window columns ex-R .. ex+R of rows ey-R .. ey+R (E coordinates; rows/cols outside E hold no keeper
        // that could matter: a keeper more than R outside the interior cannot cover an interior pixel... but one
        // within R can, which is why keepers are evaluated on the whole E region)
        unsigned any = 0u;
        const int x0 = ex - R;                                  // may be negative by at most R for ix < R? no: ex >= R here
#pragma unroll
        for (int dy = -R; dy <= R; ++dy) {
            const int yy = ey + dy;
            if (yy < 0 || yy >= EH) continue;
            const unsigned* kr = s_kb + yy * KW;
            const int w0 = x0 >> 5, sh = x0 & 31;
            const unsigned long long two = (unsigned long long)kr[w0] | ((unsigned long long)kr[w0 + 1] << 32);
            any |= (unsigned)(two >> sh) & ((1u << (2 * R + 1)) - 1u);
        }
        if (!any) {
            atomicOr(&s_ub[iy * 2 + (ix >> 5)], 1u << (ix & 31));
            s_und[atomicAdd(&s_nund, 1)] = pix;
        }
    }
    __syncthreads();

    // 4. write out: undecided mask words of every interior row, the two compact lists behind one atomicAdd each
    unsigned* mrow = mask + (size_t)b * H * mask_w;
    for (int i = tid; i < kN0TH * 2; i += kN0Threads) {
        const int gy = ty0 + (i >> 1), wcol = (tx0 >> 5) + (i & 1);
        if (gy < H && wcol < mask_w) mrow[(size_t)gy * mask_w + wcol] = s_ub[i];
    }
    const int nund = s_nund, nkeep = min(s_nkeep, kMaxKeep);
    if (tid == 0) {
        s_base_und = nund ? atomicAdd(cnt + 1, nund) : 0;
        s_base_keep = nkeep ? atomicAdd(cnt, nkeep) : 0;
    }
    __syncthreads();
    unsigned* uout = und + (size_t)b * H * W + s_base_und;
    for (int i = tid; i < nund; i += kN0Threads) uout[i] = s_und[i];
    unsigned long long* kout = keys + (size_t)b * kcap;
    for (int i = tid; i < nkeep; i += kN0Threads)
        if (s_base_keep + i < kcap) kout[s_base_keep + i] = s_keep[i];
}

// ------------------------------------------------------------------------------------------------
// Rounds >= 1 on the compact list
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRoundsThreads)
nms_rounds_kernel(const float* __restrict__ heat, int H, int W, int r, int border, int kcap,
                  unsigned long long* __restrict__ keys, int* __restrict__ counters, unsigned* __restrict__ mask,
                  int mask_w, unsigned* __restrict__ und) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int s_warp_tot[kRoundsThreads / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / kNmsCluster;
    const int rank = (int)cluster.block_rank();
    int* cnt = counters + b * kNmsCounters;
    const long n0 = min((long)__ldcg(cnt + 1), (long)H * W);
    if (n0 == 0) return;                                     // uniform over the cluster
    const long seg0 = rank * n0 / kNmsCluster, seg1 = (rank + 1) * n0 / kNmsCluster;
    unsigned* list = und + (size_t)b * H * W + seg0;
    int m = (int)(seg1 - seg0);
    const float* hmap = heat + (size_t)b * H * W;
    unsigned* mrow = mask + (size_t)b * H * mask_w;
    unsigned long long* kout = keys + (size_t)b * kcap;
    const int side = 2 * r + 1;

    for (int k = 1; k < (1 << 30); ++k) {
        // ---- phase A: is the candidate the maximum of the undecided candidates of its window? ----
        for (int i = warp; i < m; i += kRoundsThreads / 32) {
            const unsigned p = list[i] & 0x7fffffffu;
            const int y = (int)(p / (unsigned)W), x = (int)(p % (unsigned)W);
            const unsigned kp = sortable_bits(__ldg(hmap + p));
            bool beaten = false;
            for (int t = lane; t < side * side; t += 32) {
                const int qy = y + t / side - r, qx = x + t % side - r;
                if (qy < 0 || qy >= H || qx < 0 || qx >= W || (qy == y && qx == x)) continue;
                const unsigned word = __ldcg(mrow + (size_t)qy * mask_w + (qx >> 5));
                if ((word >> (qx & 31)) & 1u) {
                    const unsigned q = (unsigned)(qy * W + qx);
                    const unsigned kq = sortable_bits(__ldg(hmap + q));
                    if (kq > kp || (kq == kp && q < p)) beaten = true;
                }
            }
            const bool any = __any_sync(0xffffffffu, beaten);
            if (lane == 0 && !any) list[i] = p | 0x80000000u;
        }
        __syncthreads();
        cluster.sync();
        if (k > 1 && __ldcg(cnt + 2 + (k - 1) % 3) == 0) break;            // nothing was left after the last round
        // ---- phase B: new keepers clear their window in the mask and are emitted ----
        for (int i = warp; i < m; i += kRoundsThreads / 32) {
            const unsigned e = list[i];
            if (!(e >> 31)) continue;
            const unsigned p = e & 0x7fffffffu;
            const int y = (int)(p / (unsigned)W), x = (int)(p % (unsigned)W);
            if (lane < side) {
                const int qy = y + lane - r;
                if (qy >= 0 && qy < H) {
                    const int x0 = max(x - r, 0), x1 = min(x + r, W - 1);
                    const int w0 = x0 >> 5, w1 = x1 >> 5;
                    const unsigned lo = 0xffffffffu << (x0 & 31), hi = 0xffffffffu >> (31 - (x1 & 31));
                    if (w0 == w1) atomicAnd(mrow + (size_t)qy * mask_w + w0, ~(lo & hi));
                    else { atomicAnd(mrow + (size_t)qy * mask_w + w0, ~lo); atomicAnd(mrow + (size_t)qy * mask_w + w1, ~hi); }
                }
            }
            if (lane == 0 && !(x < border || x >= W - border || y < border || y >= H - border)) {
                const int pos = atomicAdd(cnt, 1);
                if (pos < kcap) kout[pos] = survivor_key(sortable_bits(__ldg(hmap + p)), p);
            }
        }
        __threadfence();
        __syncthreads();
        cluster.sync();
        // ---- compaction: keep the entries whose own bit survived ----
        int new_m = 0;
        for (int base = 0; base < m; base += kRoundsThreads) {
            const int i = base + tid;
            unsigned p = 0u;
            bool alive = false;
            if (i < m) {
                const unsigned e = list[i];
                p = e & 0x7fffffffu;
                if (!(e >> 31)) {
                    const int y = (int)(p / (unsigned)W), x = (int)(p % (unsigned)W);
                    alive = (__ldcg(mrow + (size_t)y * mask_w + (x >> 5)) >> (x & 31)) & 1u;
                }
            }
            const unsigned ab = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) s_warp_tot[warp] = __popc(ab);
            __syncthreads();
            int off = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kRoundsThreads / 32; ++w) {
                const int c = s_warp_tot[w];
                if (w < warp) off += c;
                tot += c;
            }
            if (alive) list[new_m + off + __popc(ab & ((1u << lane) - 1u))] = p;
            new_m += tot;
            __syncthreads();
        }
        m = new_m;
        if (tid == 0) {
            if (m) atomicAdd(cnt + 2 + k % 3, m);
            if (rank == 0) cnt[2 + (k + 1) % 3] = 0;
            __threadfence();
        }
        (void)s_total;
    }
}

template <int R>
static void launch_round0_t(const float* heat, int B, int H, int W, float thresh, int border, const NmsWorkspace& ws,
                            cudaStream_t st) {
    constexpr int LW = kN0TW + 4 * R, LH = kN0TH + 4 * R, EW = kN0TW + 2 * R, EH = kN0TH + 2 * R;
    constexpr int KW = (EW + 31) / 32 + 1;
    constexpr int kMaxKeep = ((kN0TW + R) / (R + 1) + 1) * ((kN0TH + R) / (R + 1) + 1);
    const size_t smem = sizeof(unsigned) * ((size_t)LH * LW + EH * KW + kN0TH * 2 + kN0TH * kN0TW) + sizeof(unsigned long long) * kMaxKeep +
                        sizeof(unsigned short) * ((size_t)EH * EW) + 16;
    auto kern = nms_round0_kernel<R>;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + kN0TW - 1) / kN0TW, (H + kN0TH - 1) / kN0TH, B);
    kern<<<grid, kN0Threads, smem, st>>>(heat, H, W, thresh, border, ws.kcap, ws.keys, ws.counters, ws.mask, ws.mask_w, ws.und);
    SPB_CHECK_LAUNCH();
}

void launch_nms(const float* heat, int B, int H, int W, float thresh, int radius, int border, const NmsWorkspace& ws,
                cudaStream_t st) {
    if (radius < 0 || radius > kNmsMaxR) throw std::invalid_argument("nms_dist must be in [0, 8]");
    if ((long)H * W >= (1l << 31)) throw std::invalid_argument("image too large");
    SPB_CUDA(cudaMemsetAsync(ws.counters, 0, sizeof(int) * kNmsCounters * B, st));
    switch (radius) {
        case 0: launch_round0_t<0>(heat, B, H, W, thresh, border, ws, st); break;
        case 1: launch_round0_t<1>(heat, B, H, W, thresh, border, ws, st); break;
        case 2: launch_round0_t<2>(heat, B, H, W, thresh, border, ws, st); break;
        case 3: launch_round0_t<3>(heat, B, H, W, thresh, border, ws, st); break;
        case 4: launch_round0_t<4>(heat, B, H, W, thresh, border, ws, st); break;
        case 5: launch_round0_t<5>(heat, B, H, W, thresh, border, ws, st); break;
        case 6: launch_round0_t<6>(heat, B, H, W, thresh, border, ws, st); break;
        case 7: launch_round0_t<7>(heat, B, H, W, thresh, border, ws, st); break;
        default: launch_round0_t<8>(heat, B, H, W, thresh, border, ws, st); break;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * kNmsCluster);
    cfg.blockDim = dim3(kRoundsThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kNmsCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SPB_CUDA(cudaLaunchKernelEx(&cfg, nms_rounds_kernel, heat, H, W, radius, border, ws.kcap, ws.keys, ws.counters, ws.mask,
                                ws.mask_w, ws.und));
}

}  // namespace spb200
