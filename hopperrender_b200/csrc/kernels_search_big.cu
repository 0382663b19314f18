// kernels_search_big.cu — search passes for windows of 32x32 flow pixels and larger at full flow resolution
// (7 of the 11 iterations at 4K).  Same arithmetic as sadPassKernel (kernels_search.cu), different data movement:
//
// A warp owns a 32x32 tile of one window.  Along the candidate axis every thread owns a run of 32 pixels and
// slides over the 32 + (HI-LO) frame-1 samples that run can ever be compared with: each sample is fetched ONCE
// and feeds every (pixel, candidate) pair it belongs to (up to R of them), all register indices being
// compile-time after unrolling.  Per 512 VABSDIFF4 a thread issues 145 + 32 loads (R = 16) instead of 512 + 32.
//   * Y step (candidates along rows): lanes = columns, the fetches are coalesced row segments straight from L2.
//   * X step (candidates along columns): lanes = rows; the CTA first stages the frame-1 rows (+halo, mirrored)
//     and the frame-2 tile in shared memory with coalesced loads, then reads them transposed through an odd
//     row pitch (bank-conflict free).
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

// ---- Y step -----------------------------------------------------------------------------------------
template <int R> __global__ void __launch_bounds__(128) sadBigYKernel(const SearchArgs a) {
    constexpr int LO = CandSpan<R>::LO, LEN = CandSpan<R>::LEN;
    const int lane = threadIdx.x;
    const int cx = blockIdx.x * 32 + lane;
    const int r0 = (blockIdx.y * 4 + threadIdx.y) * 32;
    if (r0 >= a.lh) return;  // warp-uniform
    const int wx = (blockIdx.x * 32) >> a.wsLog2, wy = r0 >> a.wsLog2;
    int ox, oy;
    loadWindowOffsets<1>(a, wx, wy, ox, oy);
    uint32_t acc[16];
#pragma unroll
    for (int z = 0; z < 16; ++z) acc[z] = 0;
    const int np = min(32, a.lh - r0);
    if (cx < a.lw) {
        uint32_t f2[32];
        const uint32_t* __restrict__ p2 = a.plane2 + (size_t)r0 * a.pitch + cx;
        const uint32_t* __restrict__ col = a.plane1 + mirrorSearch(cx + ox, a.W);
        const int by = r0 + oy + LO;  // frame-1 row of sample 0
        if (np == 32) {
#pragma unroll
            for (int p = 0; p < 32; ++p) f2[p] = __ldg(p2 + (size_t)p * a.pitch);
            if (by >= 0 && by + LEN <= a.H) {
                const uint32_t* __restrict__ p1 = col + (size_t)by * a.pitch;
                slidingSad<R, false>(acc, f2, 32, [&](int j) { return __ldg(p1 + (size_t)j * a.pitch); });
            } else {
                slidingSad<R, false>(acc, f2, 32, [&](int j) { return __ldg(col + (size_t)mirrorSearch(by + j, a.H) * a.pitch); });
            }
        } else {
#pragma unroll
            for (int p = 0; p < 32; ++p) f2[p] = __ldg(p2 + (size_t)min(p, np - 1) * a.pitch);
            slidingSad<R, true>(acc, f2, np, [&](int j) { return __ldg(col + (size_t)mirrorSearch(by + j, a.H) * a.pitch); });
        }
    }
    reduceAndEmit<R, 1>(a, acc, lane, wx, wy, ox, oy);
}

// ---- X step -----------------------------------------------------------------------------------------
template <int R, int NW> struct XLayout {
    static constexpr int RW = 32 * NW + CandSpan<R>::SPAN;  // staged frame-1 words per row
    static constexpr int RP = RW | 1;                       // odd pitch: lanes (= rows) fall into distinct banks
    static constexpr int FP = 32 * NW + 1;                  // frame-2 pitch
    static constexpr int BYTES = (32 * RP + 32 * FP) * 4;
};

template <int R, int NW> __global__ void __launch_bounds__(32 * NW) sadBigXKernel(const SearchArgs a) {
    constexpr int LO = CandSpan<R>::LO;
    using L = XLayout<R, NW>;
    extern __shared__ uint32_t smem[];
    uint32_t* __restrict__ s1 = smem;
    uint32_t* __restrict__ s2 = smem + 32 * L::RP;
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * 32 + lane;
    const int X0 = blockIdx.x * 32 * NW, Y0 = blockIdx.y * 32;
    const int wx = X0 >> a.wsLog2, wy = Y0 >> a.wsLog2;
    int ox, oy;
    loadWindowOffsets<0>(a, wx, wy, ox, oy);

    // stage frame 1: rows Y0+oy .. +31, columns X0+ox+LO .. +RW-1 (mirrored), coalesced along the row
    const int bx = X0 + ox + LO;
    const bool interior = bx >= 0 && bx + L::RW <= a.W;
    for (int idx = tid; idx < 32 * L::RW; idx += 32 * NW) {
        const int l = idx / L::RW, i = idx - l * L::RW;
        const int ny = mirrorSearch(Y0 + l + oy, a.H);
        const int nx = interior ? bx + i : mirrorSearch(bx + i, a.W);
        s1[l * L::RP + i] = __ldg(a.plane1 + (size_t)ny * a.pitch + nx);
    }
    // stage frame 2: the tile itself
    for (int idx = tid; idx < 32 * 32 * NW; idx += 32 * NW) {
        const int l = idx / (32 * NW), i = idx - l * (32 * NW);
        const int y = min(Y0 + l, a.lh - 1), x = min(X0 + i, a.lw - 1);
        s2[l * L::FP + i] = __ldg(a.plane2 + (size_t)y * a.pitch + x);
    }
    __syncthreads();

    const int cy = Y0 + lane, c0 = X0 + warp * 32;
    uint32_t acc[16];
#pragma unroll
    for (int z = 0; z < 16; ++z) acc[z] = 0;
    if (cy < a.lh && c0 < a.lw) {
        const int np = min(32, a.lw - c0);
        uint32_t f2[32];
        const uint32_t* __restrict__ q2 = s2 + lane * L::FP + warp * 32;
#pragma unroll
        for (int p = 0; p < 32; ++p) f2[p] = q2[p];
        const uint32_t* __restrict__ row = s1 + lane * L::RP + warp * 32;
        if (np == 32)
            slidingSad<R, false>(acc, f2, 32, [&](int j) { return row[j]; });
        else
            slidingSad<R, true>(acc, f2, np, [&](int j) { return row[j]; });
    }
    reduceAndEmit<R, 0>(a, acc, lane, wx, wy, ox, oy);
}

template <int R> int launchBigR(hrb_ofc* h, const SearchArgs& a, int step) {
    if (step == 1) {
        const dim3 block(32, 4, 1);
        const dim3 grid((a.lw + 31) / 32, (a.lh + 127) / 128, 1);
        sadBigYKernel<R><<<grid, block, 0, h->stream>>>(a);
    } else if (a.ws >= 128) {
        const dim3 block(32, 4, 1);
        const dim3 grid((a.lw + 127) / 128, (a.lh + 31) / 32, 1);
        sadBigXKernel<R, 4><<<grid, block, XLayout<R, 4>::BYTES, h->stream>>>(a);
    } else if (a.ws == 64) {
        const dim3 block(32, 2, 1);
        const dim3 grid((a.lw + 63) / 64, (a.lh + 31) / 32, 1);
        sadBigXKernel<R, 2><<<grid, block, XLayout<R, 2>::BYTES, h->stream>>>(a);
    } else {
        const dim3 block(32, 1, 1);
        const dim3 grid((a.lw + 31) / 32, (a.lh + 31) / 32, 1);
        sadBigXKernel<R, 1><<<grid, block, XLayout<R, 1>::BYTES, h->stream>>>(a);
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
