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
// Round 0 is a dense streaming pass (nms_round0_kernel): 64x32-pixel tiles with a 2r halo in shared
// memory, separable window maxima with register sliding windows, keeper flags as row bitmasks, dilation
// by shifts; it emits the keepers, a bit-per-pixel mask of the still-undecided candidates and their
// compact list.  After it only a few percent of the candidates are left, so the remaining rounds
// (nms_rounds_kernel, one 8-CTA cluster per image) work on the compact list: a warp per candidate
// scans its window through the bitmask (heat is read only where a bit is set), keepers clear their
// window bits with atomics, the list is compacted, two cluster barriers per round.
#include <cooperative_groups.h>

#include "kernels.h"
#include "sortkey.cuh"

namespace cg = cooperative_groups;

namespace spb200 {

constexpr int kN0TW = 64, kN0TH = 32;        // interior tile of round 0 (two mask words per row)
constexpr int kN0Threads = 256;
constexpr int kNmsMaxR = 8;
constexpr int kNmsCluster = 8;
constexpr int kRoundsThreads = 256;

// ------------------------------------------------------------------------------------------------
// Round 0
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kN0Threads)
nms_round0_kernel(const float* __restrict__ heat, int H, int W, float thresh, int border, int kcap,
                  unsigned long long* __restrict__ keys, int* __restrict__ counters, unsigned* __restrict__ mask,
                  int mask_w, unsigned* __restrict__ und) {
    constexpr int LW = kN0TW + 4 * R, LH = kN0TH + 4 * R;      // loaded region (halo 2R)
    constexpr int EW = kN0TW + 2 * R, EH = kN0TH + 2 * R;      // region where keepers are evaluated (halo R)
    constexpr int EWQ = (EW + 3) / 4;                          // 4-wide strips per row
    constexpr int EHQ = (EH + 3) / 4;
    static_assert(EW <= 96, "keeper rows are three 32-bit words");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* s_key = reinterpret_cast<unsigned*>(smem_raw);   // [LH][LW]   sortable key, 0 = not a candidate
    unsigned* s_rm = s_key + LH * LW;                          // [LH][EW]   horizontal window maxima
    unsigned* s_kb = s_rm + LH * EW;                           // [EH][3]    keeper bits, bit ex of row ey
    unsigned* s_dil = s_kb + EH * 3;                           // [EH][2]    keepers dilated horizontally (interior cols)
    unsigned* s_sup = s_dil + EH * 2;                          // [TH][2]    ... and vertically: suppressed bits

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int ty0 = blockIdx.y * kN0TH, tx0 = blockIdx.x * kN0TW;
    const float* hmap = heat + (size_t)b * H * W;

    // 1. keys of the loaded region
    for (int i = tid; i < LH * LW; i += kN0Threads) {
        const int gy = ty0 - 2 * R + i / LW, gx = tx0 - 2 * R + i % LW;
        unsigned key = 0u;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const float h = __ldg(hmap + (size_t)gy * W + gx);
            if (h >= thresh) key = sortable_bits(h);
        }
        s_key[i] = key;
    }
    for (int i = tid; i < EH * 3; i += kN0Threads) s_kb[i] = 0u;
    __syncthreads();

    // 2. horizontal maxima, four adjacent outputs per work item from a (4 + 2R)-wide register window
    for (int it = tid; it < LH * EWQ; it += kN0Threads) {
        const int ly = it / EWQ, ex = (it % EWQ) * 4;
        unsigned v[4 + 2 * R];
#pragma unroll
        for (int j = 0; j < 4 + 2 * R; ++j) v[j] = (ex + j < LW) ? s_key[ly * LW + ex + j] : 0u;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            unsigned m = v[o];
#pragma unroll
            for (int d = 1; d <= 2 * R; ++d) m = max(m, v[o + d]);
            if (ex + o < EW) s_rm[ly * EW + ex + o] = m;
        }
    }
    __syncthreads();

    // 3. vertical maxima -> keeper test.  Work items are (strip of 4 rows, column) with the column padded
    //    to 96 so that a warp owns one 32-bit word of each of its four rows (ballot -> bit row).
    for (int it = tid; it < EHQ * 96; it += kN0Threads) {
        const int eq = it / 96, ex = it % 96;
        const int ey0 = eq * 4;
        unsigned v[4 + 2 * R];
#pragma unroll
        for (int j = 0; j < 4 + 2 * R; ++j) v[j] = (ex < EW && ey0 + j < LH) ? s_rm[(ey0 + j) * EW + ex] : 0u;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int ey = ey0 + o;
            bool keep = false;
            if (ex < EW && ey < EH) {
                unsigned m = v[o];
#pragma unroll
                for (int d = 1; d <= 2 * R; ++d) m = max(m, v[o + d]);
                const unsigned me = s_key[(ey + R) * LW + ex + R];
                if (me != 0u && me == m) {
                    // ties: an equal key earlier in raster order wins
                    keep = true;
                    for (int dy = -R; dy <= R && keep; ++dy)
                        for (int dx = -R; dx <= R; ++dx) {
                            if (dy > 0 || (dy == 0 && dx >= 0)) break;
                            if (s_key[(ey + R + dy) * LW + ex + R + dx] == me) { keep = false; break; }
                        }
                }
            }
            const unsigned bits = __ballot_sync(0xffffffffu, keep);
            if (lane == 0 && ey < EH) s_kb[ey * 3 + (ex >> 5)] = bits;
        }
    }
    __syncthreads();

    // 4. dilate the keeper bits horizontally: interior column ix sees E columns ix .. ix+2R
    for (int it = tid; it < EH * 2; it += kN0Threads) {
        const int ey = it >> 1, ws = it & 1;
        const unsigned w0 = s_kb[ey * 3 + ws], w1 = s_kb[ey * 3 + ws + 1];
        unsigned acc = w0;
#pragma unroll
        for (int d = 1; d <= 2 * R; ++d) acc |= (w0 >> d) | (w1 << (32 - d));
        s_dil[it] = acc;
    }
    __syncthreads();
    //    ... and vertically: interior row iy sees E rows iy .. iy+2R
    for (int it = tid; it < kN0TH * 2; it += kN0Threads) {
        const int iy = it >> 1, ws = it & 1;
        unsigned acc = 0u;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) acc |= s_dil[(iy + d) * 2 + ws];
        s_sup[it] = acc;
    }
    __syncthreads();

    // 5. decide the interior: a warp owns one mask word (32 pixels of one row) at a time
    int* cnt = counters + b * kNmsCounters;
    unsigned long long* kout = keys + (size_t)b * kcap;
    unsigned* uout = und + (size_t)b * H * W;
    unsigned* mrow = mask + (size_t)b * H * mask_w;
    for (int unit = warp; unit < kN0TH * 2; unit += kN0Threads / 32) {
        const int iy = unit >> 1, ws = unit & 1;
        const int ix = ws * 32 + lane;
        const int gy = ty0 + iy, gx = tx0 + ix;
        const unsigned key = s_key[(iy + 2 * R) * LW + ix + 2 * R];          // 0 outside the image
        const int ex = ix + R;
        const bool keep = (s_kb[(iy + R) * 3 + (ex >> 5)] >> (ex & 31)) & 1u;
        const bool sup = (s_sup[unit] >> lane) & 1u;
        const bool cand = key != 0u;
        const bool undecided = cand && !keep && !sup;
        const bool emit = cand && keep && !(gx < border || gx >= W - border || gy < border || gy >= H - border);
        const unsigned pix = (unsigned)(gy * W + gx);
        const unsigned ub = __ballot_sync(0xffffffffu, undecided);
        const unsigned eb = __ballot_sync(0xffffffffu, emit);
        if (lane == 0 && gy < H && (tx0 >> 5) + ws < mask_w) mrow[(size_t)gy * mask_w + (tx0 >> 5) + ws] = ub;
        if (eb) {
            int base = 0;
            if (lane == 0) base = atomicAdd(cnt, __popc(eb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (emit) {
                const int pos = base + __popc(eb & ((1u << lane) - 1u));
                if (pos < kcap) kout[pos] = survivor_key(key, pix);
            }
        }
        if (ub) {
            int base = 0;
            if (lane == 0) base = atomicAdd(cnt + 1, __popc(ub));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (undecided) uout[base + __popc(ub & ((1u << lane) - 1u))] = pix;
        }
    }
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
    const size_t smem = sizeof(unsigned) * ((size_t)LH * LW + (size_t)LH * EW + EH * 3 + EH * 2 + kN0TH * 2);
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
