// kernels_search_big.cu — search passes for windows of 32x32 flow pixels and larger at full flow resolution
// (7 of the 11 iterations at 4K).  Same arithmetic as sadPassKernel (kernels_search.cu), different data movement:
//
// A warp owns a 32x32 tile of one window.  Along the candidate axis every thread owns a run of 32 pixels and
// slides over the 32 + (HI-LO) frame-1 samples that run can ever be compared with: each sample is fetched ONCE
// and feeds every (pixel, candidate) pair it belongs to (up to R of them), all register indices being
// compile-time after unrolling.  Per 512 VABSDIFF4 a thread issues 145 + 32 loads (R = 16) instead of 512 + 32.
// Both steps run the same kernel: the Y step on the row-major planes (lanes = x, run along y), the X step on the
// transposed planes (lanes = y, run along x) — see View in search_common.cuh.  Every fetch is a contiguous row
// segment of 32 words shared by the warp; vertically stacked warps of a CTA share the halo rows through L1.
// The warp then reduces its R sums with a recursive-halving butterfly and either finalizes the window itself
// (ws == 32) or adds R partial sums to the per-window scratch (ws > 32).
#include "search_common.cuh"

namespace hrb {

namespace {

template <int R> struct CandSpan {
    static constexpr int LO = candOffset<R>(0);
    static constexpr int HI = candOffset<R>(R - 1);
    static constexpr int SPAN = HI - LO;   // extra samples along the candidate axis
    static constexpr int LEN = 32 + SPAN;  // samples a 32-pixel run is compared with
};

// acc[z] += sum over the run's pixels p of SAD(frame1[p + d_z], frame2[p]); fetch(j) returns frame1 sample LO + j.
template <int R, bool CHECKED, typename Fetch>
__device__ __forceinline__ void slidingSad(uint32_t (&acc)[16], const uint32_t (&f2)[32], int np, Fetch fetch) {
#pragma unroll
    for (int j = 0; j < CandSpan<R>::LEN; ++j) {
        const uint32_t f1 = fetch(j);
#pragma unroll
        for (int z = 0; z < R; ++z) {
            const int p = j - (candOffset<R>(z) - CandSpan<R>::LO);
            if (p >= 0 && p < 32) {
                if (!CHECKED || p < np) acc[z] = sad4(f1, f2[p], acc[z]);
            }
        }
    }
}

// Warp-level reduction of the R sums of a 32x32 tile, then arg-min (ws == 32) or scratch accumulation (ws > 32).
template <int R, int STEP>
__device__ __forceinline__ void reduceAndEmit(const SearchArgs& a, uint32_t (&acc)[16], int lane, int wx, int wy, int ox, int oy) {
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8;
    bfly<16>(acc, 1, b0);
    bfly<8>(acc, 2, b1);
    bfly<4>(acc, 4, b2);
    bfly<2>(acc, 8, b3);
    const uint32_t s = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 16);
    const int z = (b0 ? 8 : 0) + (b1 ? 4 : 0) + (b2 ? 2 : 0) + (b3 ? 1 : 0);  // the layer this lane ended up with
    uint32_t sum = s;
    if (a.ws > 32) {
        // the window spans several warp tiles: add the partial sums to the per-window scratch; the LAST tile to arrive
        // (ticket counter) reads the complete sums back, finalizes the window and leaves scratch and ticket zeroed for
        // the next pass — no separate finalize kernel, no memset between passes
        const size_t w = (size_t)(wy * a.nWx + wx);
        if (lane < 16 && z < R) atomicAdd(&a.winSums[w * 16 + z], s);
        __threadfence();
        const int x0 = wx << a.wsLog2, y0 = wy << a.wsLog2;
        const unsigned tilesInWindow = (unsigned)(((min(a.ws, a.lw - x0) + 31) >> 5) * ((min(a.ws, a.lh - y0) + 31) >> 5));
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd(&a.winTicket[w], 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket != tilesInWindow - 1) return;
        __threadfence();
        sum = (lane < 16 && z < R) ? __ldcg(&a.winSums[w * 16 + z]) : 0u;
        if (lane < 16 && z < R) a.winSums[w * 16 + z] = 0;
        if (lane == 0) a.winTicket[w] = 0;
    }
    {
        const WindowCtx c = loadWindowCtx<STEP>(a, wx, wy, ox, oy);
        unsigned long long key = ~0ull;
        if (z < R) {
            const uint32_t total = windowTotal<R>(a, c, sum, z);
            if (lane < 16) tapTotal<R>(a, wx, wy, z, total);
            key = layerKey(total, z);
        }
        key = min(key, shflXor64(key, 1));
        key = min(key, shflXor64(key, 2));
        key = min(key, shflXor64(key, 4));
        key = min(key, shflXor64(key, 8));
        if (lane == 0) commitWindow<R, STEP>(a, wx, wy, c.o, (int)(key & 0xff));
    }
}

// ---- the sliding kernel (both steps, see View) -------------------------------------------------------
template <int R, int STEP, int NWARPS> __global__ void __launch_bounds__(32 * NWARPS) sadSlideKernel(const SearchArgs a) {
    constexpr int LO = CandSpan<R>::LO, LEN = CandSpan<R>::LEN;
    const View<STEP> vw(a);
    const int lane = threadIdx.x;
    const int cu = blockIdx.x * 32 + lane;
    const int v0 = (blockIdx.y * NWARPS + threadIdx.y) * 32;
    if (v0 >= vw.lv) return;  // warp-uniform
    const int wu = (blockIdx.x * 32) >> a.wsLog2, wv = v0 >> a.wsLog2;
    const int wx = View<STEP>::wx(wu, wv), wy = View<STEP>::wy(wu, wv);
    int ox, oy;
    loadWindowOffsets<STEP>(a, wx, wy, ox, oy);
    uint32_t acc[16];
#pragma unroll
    for (int z = 0; z < 16; ++z) acc[z] = 0;
    const int np = min(32, vw.lv - v0);
    if (cu < vw.lu) {
        uint32_t f2[32];
        const uint32_t* __restrict__ p2 = vw.p2 + (size_t)v0 * vw.pitch + cu;
        const uint32_t* __restrict__ col = vw.p1 + mirrorSearch(cu + View<STEP>::ou(ox, oy), vw.dimU);
        const int bv = v0 + View<STEP>::ov(ox, oy) + LO;  // frame-1 row of sample 0
        if (np == 32) {
#pragma unroll
            for (int p = 0; p < 32; ++p) f2[p] = __ldg(rowPtr(p2, vw.pitch, p));
            if (bv >= 0 && bv + LEN <= vw.dimV) {
                const uint32_t* __restrict__ p1 = rowPtr(col, vw.pitch, bv);
                slidingSad<R, false>(acc, f2, 32, [&](int j) { return __ldg(rowPtr(p1, vw.pitch, j)); });
            } else {
                slidingSad<R, false>(acc, f2, 32, [&](int j) { return __ldg(rowPtr(col, vw.pitch, mirrorSearch(bv + j, vw.dimV))); });
            }
        } else {
#pragma unroll
            for (int p = 0; p < 32; ++p) f2[p] = __ldg(rowPtr(p2, vw.pitch, min(p, np - 1)));
            slidingSad<R, true>(acc, f2, np, [&](int j) { return __ldg(rowPtr(col, vw.pitch, mirrorSearch(bv + j, vw.dimV))); });
        }
    }
    reduceAndEmit<R, STEP>(a, acc, lane, wx, wy, ox, oy);
}

// ---- the sliding kernel with the frame-1 rows staged in shared memory -------------------------------------------
// CTA = NW warps stacked along v inside ONE window (32 x 32*NW pixels, NW = min(ws, 128) / 32).  The 32*NW + (HI-LO)
// frame-1 rows the tile can touch are copied once with 16-byte loads (through the mirror at the frame border); the
// sliding loop then reads them with LDS at compile-time offsets: no per-fetch address arithmetic is left on the ALU
// pipe, and rows shared by the stacked runs cross L2 -> SM once.
constexpr int SSP = 36;  // staged row pitch in words: 32 + alignment slack, 9 x 16 B

template <int R, int STEP, int NW, int WU> __global__ void __launch_bounds__(32 * NW * WU, 768 / (32 * NW * WU)) sadSlideStagedKernel(const SearchArgs a) {
    // tile = 32*WU columns x 32*NW rows inside ONE window; warp (wu, wv) owns the 32 x 32 sub-tile at (32*wu, 32*wv).
    // Wider tiles (WU = 2) make the staged row segments 272 B long: fewer, longer DRAM bursts per row.
    constexpr int LO = CandSpan<R>::LO, SPAN = CandSpan<R>::SPAN;
    constexpr int ROWS = 32 * NW + SPAN;
    constexpr int PW = 32 * WU + 4;     // staged row pitch in words (alignment slack), a multiple of 4
    constexpr int CH = PW / 4;          // 16-byte chunks per row
    constexpr int NT = 32 * NW * WU;
    extern __shared__ __align__(16) uint32_t s_f1[];  // [ROWS * PW]
    const View<STEP> vw(a);
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * 32 + lane;
    const int wuIdx = warp % WU, wvIdx = warp / WU;
    const int U0 = blockIdx.x * 32 * WU, V0 = blockIdx.y * 32 * NW;
    const int wu = U0 >> a.wsLog2, wv = V0 >> a.wsLog2;
    const int wx = View<STEP>::wx(wu, wv), wy = View<STEP>::wy(wu, wv);
    int ox, oy;
    loadWindowOffsets<STEP>(a, wx, wy, ox, oy);
    const int ou = View<STEP>::ou(ox, oy), ov = View<STEP>::ov(ox, oy);

    // frame-2 run of this thread first: its 32 loads are in flight while the CTA stages frame 1
    const int cu = U0 + wuIdx * 32 + lane, v0 = V0 + wvIdx * 32;
    const bool runOk = v0 < vw.lv && cu < vw.lu;
    const int np = min(32, vw.lv - v0);
    uint32_t f2[32];
    if (runOk) {
        const uint32_t* __restrict__ p2 = rowPtr(vw.p2 + cu, vw.pitch, v0);
#pragma unroll
        for (int p = 0; p < 32; ++p) f2[p] = __ldg(rowPtr(p2, vw.pitch, np == 32 ? p : min(p, np - 1)));
    }

    // stage rows V0+ov+LO .. +ROWS-1, columns U0+ou .. +32*WU-1 (16-byte aligned superset)
    const int cb = U0 + ou, ca = cb & ~3, sh = cb - ca;
    const int rb = V0 + ov + LO;
    const int rowsNeeded = min(ROWS, (vw.lv - V0) + SPAN);  // tiles cut by the flow's last row need fewer rows
    if (ca >= 0 && ca + PW <= vw.pitch && cb + 32 * WU <= vw.dimU && rb >= 0 && rb + rowsNeeded <= vw.dimV) {
        constexpr int RPI = NT / CH;  // rows copied per iteration
        if (tid < RPI * CH) {
            const int c4 = tid % CH;
            const uint32_t* __restrict__ src = vw.p1 + ca + c4 * 4;
            for (int r = tid / CH; r < rowsNeeded; r += RPI) cpAsync16(&s_f1[r * PW + c4 * 4], rowPtr(src, vw.pitch, rb + r));
        }
        cpAsyncWaitAll();
    } else {
        for (int idx = tid; idx < rowsNeeded * PW; idx += NT) {
            const int r = idx / PW, c = idx - r * PW;
            s_f1[idx] = __ldg(rowPtr(vw.p1 + mirrorSearch(ca + c, vw.dimU), vw.pitch, mirrorSearch(rb + r, vw.dimV)));
        }
    }
    __syncthreads();

    uint32_t acc[16];
#pragma unroll
    for (int z = 0; z < 16; ++z) acc[z] = 0;
    if (v0 < vw.lv && U0 + wuIdx * 32 < vw.lu) {  // warp-uniform
        if (runOk) {
            const uint32_t* __restrict__ q = &s_f1[(wvIdx * 32) * PW + wuIdx * 32 + lane + sh];
            if (np == 32)
                slidingSad<R, false>(acc, f2, 32, [&](int j) { return q[j * PW]; });
            else
                slidingSad<R, true>(acc, f2, np, [&](int j) { return q[j * PW]; });  // rows past the staged ones are never consumed
        }
        reduceAndEmit<R, STEP>(a, acc, lane, wx, wy, ox, oy);
    }
}

// ---- the same, software-pipelined ---------------------------------------------------------------------------------
// Every pass streams both 33 MB planes from HBM (the four planes of a ladder do not fit the usable L2), so the
// kernel has to keep loads in flight all the time.  Persistent CTAs (as many as are resident at once) walk over the
// tiles with TWO staging buffers: while the warps slide over tile k, the cp.async copies of tile k+1 are in flight.
template <int R, int STEP, int NW> __global__ void __launch_bounds__(32 * NW) sadSlidePipeKernel(const SearchArgs a, int nTu, int nTv) {
    constexpr int LO = CandSpan<R>::LO, SPAN = CandSpan<R>::SPAN;
    constexpr int ROWS = 32 * NW + SPAN;
    extern __shared__ __align__(16) uint32_t smem[];  // [2][ROWS * SSP]
    const View<STEP> vw(a);
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * 32 + lane;
    const int nTiles = nTu * nTv;

    struct Tile {
        int U0, V0, wx, wy, ox, oy, sh;
    };
    // issue the frame-1 copies of tile `idx` into `buf` (exactly one cp.async group per call)
    auto prefetch = [&](int idx, uint32_t* __restrict__ buf, Tile& t) {
        if (idx < nTiles) {
            t.U0 = (idx / nTv) * 32;  // consecutive tiles walk along v: neighbouring CTAs share halo rows in L2
            t.V0 = (idx % nTv) * 32 * NW;
            const int wu = t.U0 >> a.wsLog2, wv = t.V0 >> a.wsLog2;
            t.wx = View<STEP>::wx(wu, wv);
            t.wy = View<STEP>::wy(wu, wv);
            loadWindowOffsets<STEP>(a, t.wx, t.wy, t.ox, t.oy);
            const int cb = t.U0 + View<STEP>::ou(t.ox, t.oy), ca = cb & ~3;
            t.sh = cb - ca;
            const int rb = t.V0 + View<STEP>::ov(t.ox, t.oy) + LO;
            const int rowsNeeded = min(ROWS, (vw.lv - t.V0) + SPAN);
            if (ca >= 0 && ca + SSP <= vw.pitch && cb + 32 <= vw.dimU && rb >= 0 && rb + rowsNeeded <= vw.dimV) {
                constexpr int RPI = (32 * NW) / 9;
                if (tid < RPI * 9) {
                    const int c4 = tid % 9;
                    const uint32_t* __restrict__ src = vw.p1 + ca + c4 * 4;
                    for (int r = tid / 9; r < rowsNeeded; r += RPI) cpAsync16(&buf[r * SSP + c4 * 4], rowPtr(src, vw.pitch, rb + r));
                }
            } else {
                for (int i = tid; i < rowsNeeded * SSP; i += 32 * NW) {
                    const int r = i / SSP, c = i - r * SSP;
                    buf[i] = __ldg(rowPtr(vw.p1 + mirrorSearch(ca + c, vw.dimU), vw.pitch, mirrorSearch(rb + r, vw.dimV)));
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    Tile tiles[2];
    int idx = blockIdx.x;
    prefetch(idx, smem, tiles[0]);
    for (int k = 0; idx < nTiles; ++k, idx += gridDim.x) {
        const int cur = k & 1;
        uint32_t* __restrict__ buf = smem + cur * (ROWS * SSP);
        prefetch(idx + gridDim.x, smem + (cur ^ 1) * (ROWS * SSP), tiles[cur ^ 1]);
        const Tile t = tiles[cur];
        // frame-2 run of this thread (in flight while the barrier below waits for the frame-1 copies)
        const int cu = t.U0 + lane, v0 = t.V0 + warp * 32;
        const bool runOk = v0 < vw.lv && cu < vw.lu;
        const int np = min(32, vw.lv - v0);
        uint32_t f2[32];
        if (runOk) {
            const uint32_t* __restrict__ p2 = rowPtr(vw.p2 + cu, vw.pitch, v0);
#pragma unroll
            for (int p = 0; p < 32; ++p) f2[p] = __ldg(rowPtr(p2, vw.pitch, np == 32 ? p : min(p, np - 1)));
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the newest group (tile k+1) has landed
        __syncthreads();
        uint32_t acc[16];
#pragma unroll
        for (int z = 0; z < 16; ++z) acc[z] = 0;
        if (v0 < vw.lv) {  // warp-uniform
            if (runOk) {
                const uint32_t* __restrict__ q = &buf[(warp * 32) * SSP + lane + t.sh];
                if (np == 32)
                    slidingSad<R, false>(acc, f2, 32, [&](int j) { return q[j * SSP]; });
                else
                    slidingSad<R, true>(acc, f2, np, [&](int j) { return q[j * SSP]; });
            }
            reduceAndEmit<R, STEP>(a, acc, lane, t.wx, t.wy, t.ox, t.oy);
        }
        __syncthreads();  // the buffer is free for the prefetch of tile k+2
    }
}

template <int R, int STEP, int NW> int launchSlidePipe(hrb_ofc* h, const SearchArgs& a, int lu, int lv) {
    constexpr size_t BYTES = 2 * (size_t)(32 * NW + CandSpan<R>::SPAN) * SSP * 4;
    static int perSmOf[HRB_MAX_DEVICES] = {};  // per device: the shared-memory attribute belongs to the function on one device
    int& perSm = perSmOf[h->device & (HRB_MAX_DEVICES - 1)];
    if (perSm == 0) {
        HRB_CUDA(cudaFuncSetAttribute(sadSlidePipeKernel<R, STEP, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BYTES));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, sadSlidePipeKernel<R, STEP, NW>, 32 * NW, BYTES) != cudaSuccess || perSm < 1) perSm = 1;
    }
    const int nTu = (lu + 31) / 32, nTv = (lv + 32 * NW - 1) / (32 * NW);
    const int grid = min(nTu * nTv, h->smCount * perSm);
    sadSlidePipeKernel<R, STEP, NW><<<grid, dim3(32, NW, 1), BYTES, h->stream>>>(a, nTu, nTv);
    return HRB_OK;
}

template <int R, int STEP, int NW, int WU> int launchSlideStaged(hrb_ofc* h, const SearchArgs& a, int lu, int lv) {
    constexpr size_t BYTES = (size_t)(32 * NW + CandSpan<R>::SPAN) * (32 * WU + 4) * 4;
    static bool configured[HRB_MAX_DEVICES] = {};  // the attribute is per device
    if (!configured[h->device & (HRB_MAX_DEVICES - 1)]) {
        HRB_CUDA(cudaFuncSetAttribute(sadSlideStagedKernel<R, STEP, NW, WU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BYTES));
        configured[h->device & (HRB_MAX_DEVICES - 1)] = true;
    }
    sadSlideStagedKernel<R, STEP, NW, WU><<<dim3((lu + 32 * WU - 1) / (32 * WU), (lv + 32 * NW - 1) / (32 * NW), 1), dim3(32, NW * WU, 1), BYTES, h->stream>>>(a);
    return HRB_OK;
}

template <int R, int STEP> int launchBigStep(hrb_ofc* h, const SearchArgs& a) {
    const int lu = STEP == 1 ? a.lw : a.lh, lv = STEP == 1 ? a.lh : a.lw;
    if (h->searchVariant == 2) {  // the L1-fed variant (A/B)
        constexpr int NWARPS = 4;
        sadSlideKernel<R, STEP, NWARPS><<<dim3((lu + 31) / 32, (lv + 32 * NWARPS - 1) / (32 * NWARPS), 1), dim3(32, NWARPS, 1), 0, h->stream>>>(a);
    } else if (h->searchVariant == 3) {  // persistent, double-buffered variant (A/B: slower — fewer resident warps)
        int rc;
        if (a.ws >= 128)
            rc = launchSlidePipe<R, STEP, 4>(h, a, lu, lv);
        else if (a.ws == 64)
            rc = launchSlidePipe<R, STEP, 2>(h, a, lu, lv);
        else
            rc = launchSlidePipe<R, STEP, 1>(h, a, lu, lv);
        if (rc) return rc;
    } else {
        int rc;
        // tile shapes were swept on B200 (32/64 columns x 32..256 rows): all within 10 %; 32 x 128 is the fastest
        if (a.ws >= 128)
            rc = launchSlideStaged<R, STEP, 4, 1>(h, a, lu, lv);
        else if (a.ws == 64)
            rc = launchSlideStaged<R, STEP, 2, 1>(h, a, lu, lv);
        else
            rc = launchSlideStaged<R, STEP, 1, 1>(h, a, lu, lv);
        if (rc) return rc;
    }
    HRB_LAUNCH_CHECK();
    return HRB_OK;
}

template <int R> int launchBigR(hrb_ofc* h, const SearchArgs& a, int step) { return step == 1 ? launchBigStep<R, 1>(h, a) : launchBigStep<R, 0>(h, a); }

}  // namespace

// One whole pass for ws >= 32 at full flow resolution (windows larger than a tile are finalized by their last CTA).
// Like kernels_search_cand.cu this file is compiled three times (-DHRB_BIG_PART=0/1/2), four search radii per part.
#ifndef HRB_BIG_PART
#error "compile with -DHRB_BIG_PART=0, 1 or 2"
#endif
#define HRB_BIG_CONCAT2(a, b) a##b
#define HRB_BIG_CONCAT(a, b) HRB_BIG_CONCAT2(a, b)
int HRB_BIG_CONCAT(launchSearchPassBigPart, HRB_BIG_PART)(hrb_ofc* h, const SearchArgs& a, int R, int step) {
    switch (R) {
#define HRB_CASE(N) case N: return launchBigR<N>(h, a, step);
#if HRB_BIG_PART == 0
        HRB_CASE(5) HRB_CASE(6) HRB_CASE(7) HRB_CASE(8)
#elif HRB_BIG_PART == 1
        HRB_CASE(9) HRB_CASE(10) HRB_CASE(11) HRB_CASE(12)
#else
        HRB_CASE(13) HRB_CASE(14) HRB_CASE(15) HRB_CASE(16)
#endif
#undef HRB_CASE
        default: return -1;  // not handled here
    }
}

}  // namespace hrb
