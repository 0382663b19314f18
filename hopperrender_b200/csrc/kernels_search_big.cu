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
    if (a.ws > 32) {
        if (lane < 16 && z < R) atomicAdd(&a.winSums[(size_t)(wy * a.nWx + wx) * 16 + z], s);
    } else {
        const WindowCtx c = loadWindowCtx<STEP>(a, wx, wy, ox, oy);
        unsigned long long key = ~0ull;
        if (z < R) {
            const uint32_t total = windowTotal<R>(a, c, s, z);
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

template <int R> int launchBigR(hrb_ofc* h, const SearchArgs& a, int step) {
    constexpr int NWARPS = 4;
    const dim3 block(32, NWARPS, 1);
    if (step == 1) {
        const dim3 grid((a.lw + 31) / 32, (a.lh + 32 * NWARPS - 1) / (32 * NWARPS), 1);
        sadSlideKernel<R, 1, NWARPS><<<grid, block, 0, h->stream>>>(a);
    } else {
        const dim3 grid((a.lh + 31) / 32, (a.lw + 32 * NWARPS - 1) / (32 * NWARPS), 1);
        sadSlideKernel<R, 0, NWARPS><<<grid, block, 0, h->stream>>>(a);
    }
    HRB_LAUNCH_CHECK();
    return HRB_OK;
}

}  // namespace

// SAD part of one pass for ws >= 32 at full flow resolution; the caller zeroes winSums before and runs the
// large-window finalize after when ws > 32.
int launchSearchPassBig(hrb_ofc* h, const SearchArgs& a, int R, int step) {
    switch (R) {
#define HRB_CASE(N) case N: return launchBigR<N>(h, a, step);
        HRB_CASE(5) HRB_CASE(6) HRB_CASE(7) HRB_CASE(8) HRB_CASE(9) HRB_CASE(10) HRB_CASE(11) HRB_CASE(12) HRB_CASE(13) HRB_CASE(14)
        HRB_CASE(15) HRB_CASE(16)
#undef HRB_CASE
        default: return -1;  // not handled here
    }
}

}  // namespace hrb
