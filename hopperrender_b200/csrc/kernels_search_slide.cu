// kernels_search_slide.cu — search passes for windows of 64x64 flow pixels and larger at full flow resolution
// (12 of the 22 passes at 4K).  Same arithmetic as sadPassKernel (kernels_search.cu), different data movement.
//
// Data: the 8-bit planar search planes (luma + NV12-style chroma, see SearchArgs), so one VABSDIFF4 covers FOUR
// luma pixels, or the U,V samples of four pixels, instead of one pixel: half the SAD instructions of a {Y,U,V,0}
// word per pixel, and the two frames of a pass are 25 MB — they stay in the 126 MB L2 over the whole ladder.
//
// Work split (u = contiguous axis, v = candidate axis, see View): a CTA owns a tile of 128 (u) x 32*NWV (v) flow
// pixels inside ONE window row; a warp owns 128 x 32 of it, a lane a column of 4 pixels (one word) x 32 rows.
//   * frame 1: the (32*NWV + HI-LO) luma rows and the matching chroma rows every candidate of the tile can touch
//     are staged in shared memory by TMA (cp.async.bulk.tensor.2d, one elected thread, completion on an mbarrier);
//     the box starts at the 16-byte boundary below the window's displaced column (TMA needs that alignment), each
//     lane then re-aligns its words with one funnel shift per sample (skipped when the displacement is a multiple
//     of 4).  Boxes that leave the frame come back zero-filled there; the CTA patches those bytes through the
//     reference's mirror (calcDeltaSumsKernelSDR.h:86-95).
//   * along v every lane slides over the staged column: each frame-1 sample is fetched ONCE and feeds every
//     (row, candidate) pair it belongs to (up to R of them), all register indices being compile-time.
//   * chroma: the U,V pair of luma (v, u) is c[v >> 1][u & ~1].  Along v, luma rows 2k and 2k+1 with displacement s
//     read chroma rows k + ((e + s) >> 1), e = 0, 1: the same row when s is even (one SAD, weight 2), two rows when
//     it is odd.  Along u, a displacement ou maps a pixel pair onto one chroma pair when ou is even (weight 2) and
//     onto two neighbouring pairs when it is odd (two staged boxes, ou - 1 and ou + 1, weight 1 each).
//   * the warp reduces its R sums with a recursive-halving butterfly, adds them to the per-window scratch and the
//     last contributor of a window (arrival ticket) finalizes it: arg-min + offset update, no extra launch.
#include <cuda.h>

#include <cstring>
#include <mutex>

#include "search_common.cuh"

namespace hrb {

namespace {

template <int R> struct CandSpan {
    static constexpr int LO = candOffset<R>(0);
    static constexpr int HI = candOffset<R>(R - 1);
    static constexpr int SPAN = HI - LO;  // extra samples along the candidate axis
};

__host__ __device__ constexpr int floorHalf(int v) { return v >= 0 ? (v >> 1) : -((1 - v) >> 1); }
__host__ __device__ constexpr int align128(int v) { return (v + 127) & ~127; }

// geometry of the staged boxes of one window column of a tile
template <int R, int NWV, int NWIN> struct SlideGeom {
    static constexpr int SEGW = 128 / NWIN;                              // pixels (bytes) of a window column inside the tile
    static constexpr int BW = SEGW + 16;                                 // box width: 16-byte aligned superset
    static constexpr int BWW = BW / 4;                                   // ... in words
    static constexpr int ROWS = 32 * NWV + CandSpan<R>::SPAN;            // luma rows
    static constexpr int ROWSC = (32 * NWV + CandSpan<R>::SPAN) / 2 + 2; // chroma rows
    static constexpr int LUMA_BYTES = align128(ROWS * BW);
    static constexpr int CHROMA_BYTES = align128(ROWSC * BW);
    static constexpr int SMEM = NWIN * (LUMA_BYTES + CHROMA_BYTES) + 128;  // + slack to align the base to 128 bytes
};

// ---- mbarrier / TMA -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmaLoad2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(map), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// ---- the sliding loops ---------------------------------------------------------------------------------------------
// luma: acc[z] += sum over the run's rows p of SAD4(frame1 row p + d_z, frame2 row p); fetch(j) returns the lane's
// frame-1 word of staged row j = p + d_z - LO.
template <int R, typename Fetch> __device__ __forceinline__ void slideLuma(uint32_t (&acc)[16], const uint32_t (&f2)[32], Fetch fetch) {
    constexpr int LO = CandSpan<R>::LO;
#pragma unroll
    for (int j = 0; j < 32 + CandSpan<R>::SPAN; ++j) {
        const uint32_t f1 = fetch(j);
#pragma unroll
        for (int z = 0; z < R; ++z) {
            const int p = j - (candOffset<R>(z) - LO);
            if (p >= 0 && p < 32) acc[z] = sad4(f1, f2[p], acc[z]);
        }
    }
}

// chroma: PI = parity of the window's displacement along v.  Luma rows 2k + e of the run with candidate z read
// staged chroma row k + floorHalf(e + PI + d_z) - AMIN; e = 0 and e = 1 coincide when PI + d_z is even (one SAD,
// the caller doubles the sum, chromaShift).
template <int R, int PI, typename Fetch> __device__ __forceinline__ void slideChroma(uint32_t (&acc)[16], const uint32_t (&f2c)[16], Fetch fetch) {
    constexpr int LO = CandSpan<R>::LO, HI = CandSpan<R>::HI;
    constexpr int AMIN = floorHalf(PI + LO);
    constexpr int CLEN = 16 + floorHalf(1 + PI + HI) - AMIN;
#pragma unroll
    for (int jc = 0; jc < CLEN; ++jc) {
        const uint32_t c1 = fetch(jc);
#pragma unroll
        for (int z = 0; z < R; ++z) {
            const int a0 = floorHalf(PI + candOffset<R>(z)) - AMIN;
            const int a1 = floorHalf(1 + PI + candOffset<R>(z)) - AMIN;
            const int k0 = jc - a0, k1 = jc - a1;
            if (k0 >= 0 && k0 < 16) acc[z] = sad4(c1, f2c[k0], acc[z]);
            if (a1 != a0 && k1 >= 0 && k1 < 16) acc[z] = sad4(c1, f2c[k1], acc[z]);
        }
    }
}
template <int R, int PI> __device__ __forceinline__ int chromaShift(int z) { return ((PI + candOffset<R>(z)) & 1) ? 0 : 1; }

// Patches the bytes of a staged box that lie outside the plane (TMA zero-fills them) with the reference's mirrored
// samples; `all` stages every byte this way (no TMA).  PAIRS: chroma plane — columns are mirrored as (U,V) pairs.
// Box row r / byte c <-> plane row rb + r / byte ca + c; only rows [0, rows) and bytes [c0, c1) are ever read.
template <bool PAIRS>
__device__ __forceinline__ void patchBox(uint8_t* __restrict__ buf, int bw, const uint8_t* __restrict__ plane, int pitch, int dimU, int dimV, int rb, int ca, int rows,
                                         int c0, int c1, bool all, int tid, int nThreads) {
    const int cols = c1 - c0;
    auto fix = [&](int r, int c) {
        const int vr = mirrorSearch(rb + r, dimV);
        const int vcRaw = ca + c;
        int vc;
        if (PAIRS)
            vc = 2 * mirrorSearch(vcRaw >> 1, dimU >> 1) + (vcRaw & 1);
        else
            vc = mirrorSearch(vcRaw, dimU);
        buf[r * bw + c] = __ldg(rowPtr(plane, pitch, vr) + vc);
    };
    if (all) {
        for (int i = tid; i < rows * cols; i += nThreads) {
            const int r = i / cols;
            fix(r, c0 + (i - r * cols));
        }
        return;
    }
    const int rTop = min(max(-rb, 0), rows);            // rows [0, rTop) lie above the plane
    const int rBot = min(max(dimV - rb, 0), rows);      // rows [rBot, rows) lie below it
    const int nOutRows = rTop + (rows - rBot);
    for (int i = tid; i < nOutRows * cols; i += nThreads) {
        int r = i / cols;
        const int c = c0 + (i - r * cols);
        if (r >= rTop) r += rBot - rTop;
        fix(r, c);
    }
    const int cLeft = min(max(-ca, c0), c1);            // bytes [c0, cLeft) lie left of the plane
    const int cRight = min(max(dimU - ca, c0), c1);     // bytes [cRight, c1) lie right of it
    const int nOutCols = (cLeft - c0) + (c1 - cRight);
    const int nInRows = rBot - rTop;
    if (nOutCols > 0 && nInRows > 0) {
        for (int i = tid; i < nInRows * nOutCols; i += nThreads) {
            const int r = rTop + i / nOutCols;
            int c = c0 + (i % nOutCols);
            if (c >= cLeft) c += cRight - cLeft;
            fix(r, c);
        }
    }
}

struct SlideFlags {
    int useTma;       // 0: stage every box with plain loads (A/B, and when no tensor map could be encoded)
    int forceFunnel;  // 1: always take the funnel-shift fetch (A/B of the aligned fast path)
};

// ---- the kernel ------------------------------------------------------------------------------------------------------
// NWV warps stacked along v, NWIN window columns inside the 128-pixel tile (1: ws >= 128, 2: ws == 64).
template <int R, int STEP, int NWV, int NWIN>
__global__ void __launch_bounds__(32 * NWV, NWIN == 1 ? 4 : 5)
    sadSlidePlanarKernel(const SearchArgs a, const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapC, const SlideFlags flags) {
    using G = SlideGeom<R, NWV, NWIN>;
    constexpr int LO = CandSpan<R>::LO, SPAN = CandSpan<R>::SPAN;
    constexpr int NT = 32 * NWV;
    extern __shared__ uint8_t smemRaw[];
    __shared__ uint64_t barY, barC;
    uint8_t* const smem = smemRaw + ((128u - ((unsigned)__cvta_generic_to_shared(smemRaw) & 127u)) & 127u);
    auto bufY = [&](int win) { return smem + win * G::LUMA_BYTES; };
    auto bufC = [&](int win) { return smem + NWIN * G::LUMA_BYTES + win * G::CHROMA_BYTES; };

    const View<STEP> vw(a);
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * 32 + lane;
    const int U0 = blockIdx.x * 128, V0 = blockIdx.y * 32 * NWV;
    const int wv = V0 >> a.wsLog2;
    const int rowsLeft = vw.lv - V0;                                  // > 0 by the grid
    const int rowsNeeded = min(G::ROWS, rowsLeft + SPAN);            // luma rows of the boxes that are ever read

    // per window column of the tile: displacement, box origins, border flags (every thread computes both: cheap, uniform)
    int ouW[NWIN], ovW[NWIN];
    bool existW[NWIN];
    bool anyOdd = false, border = false;
#pragma unroll
    for (int w = 0; w < NWIN; ++w) {
        const int uw = U0 + w * G::SEGW;
        existW[w] = uw < vw.lu;
        ouW[w] = ovW[w] = 0;
        if (existW[w]) {
            const int wu = uw >> a.wsLog2;
            int ox, oy;
            loadWindowOffsets<STEP>(a, View<STEP>::wx(wu, wv), View<STEP>::wy(wu, wv), ox, oy);
            ouW[w] = View<STEP>::ou(ox, oy);
            ovW[w] = View<STEP>::ov(ox, oy);
            anyOdd |= (ouW[w] & 1) != 0;
            const int segValid = min(G::SEGW, vw.lu - uw);
            const int cb = uw + ouW[w], rb = V0 + ovW[w] + LO;
            border |= cb < 0 || cb + segValid + 2 > vw.dimU || rb < 0 || rb + rowsNeeded > vw.dimV;  // + 2: the box of ou + 1
        }
    }
    const bool manual = !flags.useTma;

    if (tid == 0) {
        mbarInit(&barY, 1);
        mbarInit(&barC, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // chroma box of pass `pass` (0: displacement ou & ~1, 1: + 2) and the luma box of window column w
    auto chromaOrigin = [&](int w, int pass, int& ca, int& sh, int& rbc) {
        const int cb = U0 + w * G::SEGW + (ouW[w] & ~1) + 2 * pass;
        ca = cb & ~15;
        sh = cb - ca;
        rbc = (V0 + ovW[w] + LO) >> 1;
    };
    auto lumaOrigin = [&](int w, int& ca, int& sh, int& rb) {
        const int cb = U0 + w * G::SEGW + ouW[w];
        ca = cb & ~15;
        sh = cb - ca;
        rb = V0 + ovW[w] + LO;
    };
    auto issueChroma = [&](int pass) {  // thread 0
        unsigned bytes = 0;
#pragma unroll
        for (int w = 0; w < NWIN; ++w)
            if (existW[w] && (pass == 0 || (ouW[w] & 1))) bytes += G::ROWSC * G::BW;
        mbarExpectTx(&barC, bytes);
#pragma unroll
        for (int w = 0; w < NWIN; ++w)
            if (existW[w] && (pass == 0 || (ouW[w] & 1))) {
                int ca, sh, rbc;
                chromaOrigin(w, pass, ca, sh, rbc);
                tmaLoad2d(bufC(w), &mapC, ca, rbc, &barC);
            }
    };
    auto stageChromaManual = [&](int pass, bool all) {  // all threads: patch (or fully stage) the chroma boxes of `pass`
#pragma unroll
        for (int w = 0; w < NWIN; ++w)
            if (existW[w] && (pass == 0 || (ouW[w] & 1))) {
                int ca, sh, rbc;
                chromaOrigin(w, pass, ca, sh, rbc);
                const int segValid = min(G::SEGW, vw.lu - (U0 + w * G::SEGW));
                const int rowsC = min(G::ROWSC, ((V0 + ovW[w] + LO + rowsNeeded - 1) >> 1) - rbc + 1);
                patchBox<true>(bufC(w), G::BW, vw.c1, vw.pitch, vw.dimU, vw.dimV >> 1, rbc, ca, rowsC, sh, min(sh + segValid + 4, G::BW), all, tid, NT);
            }
    };

    if (!manual && tid == 0) {
        issueChroma(0);
        unsigned bytes = 0;
#pragma unroll
        for (int w = 0; w < NWIN; ++w)
            if (existW[w]) bytes += G::ROWS * G::BW;
        mbarExpectTx(&barY, bytes);
#pragma unroll
        for (int w = 0; w < NWIN; ++w)
            if (existW[w]) {
                int ca, sh, rb;
                lumaOrigin(w, ca, sh, rb);
                tmaLoad2d(bufY(w), &mapY, ca, rb, &barY);
            }
    }

    // this lane's column and run
    const int win = NWIN == 1 ? 0 : lane >> 4;
    const int lw = NWIN == 1 ? lane : lane & 15;
    const int cu = U0 + 4 * lane, v0 = V0 + 32 * warp;
    const bool segOk = existW[win] && v0 < vw.lv;                  // this lane's window column has rows in this warp
    const bool runOk = segOk && cu < vw.lu;                        // (lu is a multiple of 4: words are never partial)
    const int np = min(32, vw.lv - v0);                            // rows of the run (even)
    const int ou = ouW[win], ov = ovW[win];
    const int pi = ov & 1;
    const bool segOdd = (ou & 1) != 0;

    uint32_t accY[16], accC[16];
#pragma unroll
    for (int z = 0; z < 16; ++z) accY[z] = accC[z] = 0;

    if (manual || border) {
        if (!manual) {
            mbarWait(&barC, 0);
            mbarWait(&barY, 0);
        }
        stageChromaManual(0, manual);
#pragma unroll
        for (int w = 0; w < NWIN; ++w)
            if (existW[w]) {
                int ca, sh, rb;
                lumaOrigin(w, ca, sh, rb);
                const int segValid = min(G::SEGW, vw.lu - (U0 + w * G::SEGW));
                patchBox<false>(bufY(w), G::BW, vw.y1, vw.pitch, vw.dimU, vw.dimV, rb, ca, rowsNeeded, sh, min(sh + segValid + 4, G::BW), manual, tid, NT);
            }
        __syncthreads();
    }

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            if (!anyOdd) break;  // CTA-uniform
            if (manual || border) {
                // border tiles stage the ou + 1 boxes with plain loads (they are a small share of the tiles)
                __syncthreads();  // everybody is done with the chroma boxes of pass 0
                stageChromaManual(1, true);
                __syncthreads();
            }
        }
        // ---- chroma of this pass ----
        if (runOk && (pass == 0 || segOdd)) {
            if (!manual && !border) mbarWait(&barC, (unsigned)pass);
            int ca, sh, rbc;
            chromaOrigin(win, pass, ca, sh, rbc);
            const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(bufC(win)) + (16 * warp) * G::BWW + (sh >> 2) + lw;
            const int shBits = (sh & 3) * 8;
            if (np == 32) {
                uint32_t f2c[16];
                const uint8_t* __restrict__ p2 = rowPtr(vw.c2 + cu, vw.pitch, v0 >> 1);
#pragma unroll
                for (int k = 0; k < 16; ++k) f2c[k] = __ldg(reinterpret_cast<const uint32_t*>(rowPtr(p2, vw.pitch, k)));
                if (shBits == 0 && !flags.forceFunnel) {
                    if (pi)
                        slideChroma<R, 1>(accC, f2c, [&](int j) { return q[j * G::BWW]; });
                    else
                        slideChroma<R, 0>(accC, f2c, [&](int j) { return q[j * G::BWW]; });
                } else {
                    if (pi)
                        slideChroma<R, 1>(accC, f2c, [&](int j) { return __funnelshift_r(q[j * G::BWW], q[j * G::BWW + 1], shBits); });
                    else
                        slideChroma<R, 0>(accC, f2c, [&](int j) { return __funnelshift_r(q[j * G::BWW], q[j * G::BWW + 1], shBits); });
                }
            } else {
                // partial run at the last rows of the flow field: compact loop, one (row pair, candidate) at a time
                const int amin = (pi + LO) >> 1;
                for (int k = 0; k < (np >> 1); ++k) {
                    const uint32_t f2 = __ldg(reinterpret_cast<const uint32_t*>(rowPtr(vw.c2 + cu, vw.pitch, (v0 >> 1) + k)));
#pragma unroll
                    for (int z = 0; z < R; ++z) {
                        const int d = pi + candOffset<R>(z);
                        const int j = k + (d >> 1) - amin;
                        accC[z] = sad4(__funnelshift_r(q[j * G::BWW], q[j * G::BWW + 1], shBits), f2, accC[z]);
                        if (d & 1) accC[z] = sad4(__funnelshift_r(q[(j + 1) * G::BWW], q[(j + 1) * G::BWW + 1], shBits), f2, accC[z]);
                    }
                }
            }
        }
        if (pass == 0) {
            if (anyOdd && !(manual || border)) {
                __syncthreads();  // everybody is done with the chroma boxes of pass 0: the ou + 1 boxes may overwrite them
                if (tid == 0) issueChroma(1);
            }
            // ---- luma (the copy of the pass-1 chroma boxes is in flight meanwhile) ----
            if (runOk) {
                if (!manual && !border) mbarWait(&barY, 0);
                int ca, sh, rb;
                lumaOrigin(win, ca, sh, rb);
                const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(bufY(win)) + (32 * warp) * G::BWW + (sh >> 2) + lw;
                const int shBits = (sh & 3) * 8;
                if (np == 32) {
                    uint32_t f2[32];
                    const uint8_t* __restrict__ p2 = rowPtr(vw.y2 + cu, vw.pitch, v0);
#pragma unroll
                    for (int p = 0; p < 32; ++p) f2[p] = __ldg(reinterpret_cast<const uint32_t*>(rowPtr(p2, vw.pitch, p)));
                    if (shBits == 0 && !flags.forceFunnel)
                        slideLuma<R>(accY, f2, [&](int j) { return q[j * G::BWW]; });
                    else
                        slideLuma<R>(accY, f2, [&](int j) { return __funnelshift_r(q[j * G::BWW], q[j * G::BWW + 1], shBits); });
                } else {
                    for (int p = 0; p < np; ++p) {
                        const uint32_t f2 = __ldg(reinterpret_cast<const uint32_t*>(rowPtr(vw.y2 + cu, vw.pitch, v0 + p)));
#pragma unroll
                        for (int z = 0; z < R; ++z) {
                            const int j = p + candOffset<R>(z) - LO;
                            accY[z] = sad4(__funnelshift_r(q[j * G::BWW], q[j * G::BWW + 1], shBits), f2, accY[z]);
                        }
                    }
                }
            }
        }
    }

    // ---- combine luma and chroma: chroma weight = (rows coincide ? 2 : 1) * (columns coincide ? 2 : 1) ----
    const int evenU = segOdd ? 0 : 1;
    uint32_t acc[16];
#pragma unroll
    for (int z = 0; z < 16; ++z) {
        const int sh = z < R ? (pi ? chromaShift<R, 1>(z < R ? z : 0) : chromaShift<R, 0>(z < R ? z : 0)) + evenU : 0;
        acc[z] = accY[z] + (accC[z] << sh);
    }

    // ---- reduce over the lanes of the window column, add to the window's scratch, last contributor finalizes ----
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8;
    bfly<16>(acc, 1, b0);
    bfly<8>(acc, 2, b1);
    bfly<4>(acc, 4, b2);
    bfly<2>(acc, 8, b3);
    uint32_t s = acc[0];
    if (NWIN == 1) s += __shfl_xor_sync(0xffffffffu, s, 16);
    const int z = (b0 ? 8 : 0) + (b1 ? 4 : 0) + (b2 ? 2 : 0) + (b3 ? 1 : 0);  // the layer this lane ended up with
    const bool holder = (NWIN == 1 ? lane < 16 : true) && z < R;             // one lane per (window column, layer)
    const int uw = U0 + win * G::SEGW;
    const int wu = uw >> a.wsLog2;
    const int wx = View<STEP>::wx(wu, wv), wy = View<STEP>::wy(wu, wv);
    const size_t widx = (size_t)(wy * a.nWx + wx);
    bool last = false;
    if (segOk) {
        if (holder) atomicAdd(&a.winSums[widx * 16 + z], s);
        __threadfence();
    }
    __syncwarp();
    {
        unsigned ticket = 0, need = 1;
        if (segOk && lw == 0) {
            const int uext = min(a.ws, vw.lu - (wu << a.wsLog2)), vext = min(a.ws, vw.lv - (wv << a.wsLog2));
            need = (unsigned)(((uext + G::SEGW - 1) / G::SEGW) * ((vext + 31) >> 5));
            ticket = atomicAdd(&a.winTicket[widx], 1u);
        }
        const int leader = NWIN == 1 ? 0 : (lane & 16);
        ticket = __shfl_sync(0xffffffffu, ticket, leader);
        need = __shfl_sync(0xffffffffu, need, leader);
        last = segOk && ticket == need - 1;
    }
    uint32_t sum = 0;
    if (last) {
        __threadfence();
        if (holder) {
            sum = __ldcg(&a.winSums[widx * 16 + z]);
            a.winSums[widx * 16 + z] = 0;  // scratch and ticket are left zeroed for the next pass
        }
        if (lw == 0) a.winTicket[widx] = 0;
    }
    WindowCtx c;
    c.o = 0; c.nw = 0; c.useNb = false;
    c.nb[0] = c.nb[1] = c.nb[2] = c.nb[3] = 0;
    unsigned long long key = ~0ull;
    if (last) {
        int ox, oy;
        loadWindowOffsets<STEP>(a, wx, wy, ox, oy);
        c = loadWindowCtx<STEP>(a, wx, wy, ox, oy);
        if (z < R) {
            const uint32_t total = windowTotal<R>(a, c, sum, z);
            if (holder) {
                tapTotal<R>(a, wx, wy, z, total);
                key = layerKey(total, z);
            }
        }
    }
    key = min(key, shflXor64(key, 1));
    key = min(key, shflXor64(key, 2));
    key = min(key, shflXor64(key, 4));
    key = min(key, shflXor64(key, 8));
    if (last && lw == 0) commitWindow<R, STEP>(a, wx, wy, c.o, (int)(key & 0xff));
}

// ---- tensor maps ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled() {
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        cudaGetLastError();
    });
    return fn;
}

}  // namespace

struct TmaCache {
    struct alignas(64) Entry {
        CUtensorMap map;
        const void* base;
        int dimU, dimV, pitch, boxW, boxH;
    };
    std::vector<Entry*> entries;
};

#if HRB_SLIDE_PART == 0
void freeTmaCache(hrb_ofc* h) {
    if (!h->tmaCache) return;
    for (auto* e : h->tmaCache->entries) delete e;
    delete h->tmaCache;
    h->tmaCache = nullptr;
}
#endif

namespace {

// u8 plane [dimV][pitch] seen as a 2-D tensor of dimU x dimV bytes; boxes of boxW x boxH.  nullptr: not encodable.
const CUtensorMap* tensorMapFor(hrb_ofc* h, const uint8_t* base, int dimU, int dimV, int pitch, int boxW, int boxH) {
    if (!h->tmaCache) h->tmaCache = new TmaCache();
    for (auto* e : h->tmaCache->entries)
        if (e->base == base && e->dimU == dimU && e->dimV == dimV && e->pitch == pitch && e->boxW == boxW && e->boxH == boxH) return &e->map;
    EncodeTiledFn enc = encodeTiled();
    if (!enc) return nullptr;
    auto* e = new TmaCache::Entry();
    const cuuint64_t dims[2] = {(cuuint64_t)dimU, (cuuint64_t)dimV};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH};
    const cuuint32_t es[2] = {1, 1};
    if (enc(&e->map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        delete e;
        return nullptr;
    }
    e->base = base; e->dimU = dimU; e->dimV = dimV; e->pitch = pitch; e->boxW = boxW; e->boxH = boxH;
    if (h->tmaCache->entries.size() >= 256) {  // geometry and slots are fixed per handle: this never grows in practice
        for (auto* o : h->tmaCache->entries) delete o;
        h->tmaCache->entries.clear();
    }
    h->tmaCache->entries.push_back(e);
    return &e->map;
}

template <int R, int STEP, int NWV, int NWIN> int launchSlide(hrb_ofc* h, const SearchArgs& a) {
    using G = SlideGeom<R, NWV, NWIN>;
    static std::once_flag configured[HRB_MAX_DEVICES];  // the attribute is per device
    cudaError_t cfgErr = cudaSuccess;
    std::call_once(configured[h->device & (HRB_MAX_DEVICES - 1)],
                   [&] { cfgErr = cudaFuncSetAttribute(sadSlidePlanarKernel<R, STEP, NWV, NWIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM); });
    HRB_CUDA(cfgErr);
    const int lu = STEP == 1 ? a.lw : a.lh, lv = STEP == 1 ? a.lh : a.lw;
    const int dimU = STEP == 1 ? a.W : a.H, dimV = STEP == 1 ? a.H : a.W;
    const uint8_t* y1 = STEP == 1 ? a.y1 : a.yT1;
    const uint8_t* c1 = STEP == 1 ? a.c1 : a.cT1;
    const int pitch = STEP == 1 ? a.pitch : a.pitchT;
    SlideFlags flags;
    flags.useTma = h->searchVariant != 2;
    flags.forceFunnel = h->searchVariant == 3;
    const CUtensorMap* mY = flags.useTma ? tensorMapFor(h, y1, dimU, dimV, pitch, G::BW, G::ROWS) : nullptr;
    const CUtensorMap* mC = flags.useTma ? tensorMapFor(h, c1, dimU, dimV / 2, pitch, G::BW, G::ROWSC) : nullptr;
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof(dummy));
    if (!mY || !mC) {
        flags.useTma = 0;
        mY = mC = &dummy;
    }
    const dim3 grid((lu + 127) / 128, (lv + 32 * NWV - 1) / (32 * NWV), 1);
    sadSlidePlanarKernel<R, STEP, NWV, NWIN><<<grid, dim3(32, NWV, 1), G::SMEM, h->stream>>>(a, *mY, *mC, flags);
    HRB_LAUNCH_CHECK();
    return HRB_OK;
}

template <int R, int STEP> int launchSlideStep(hrb_ofc* h, const SearchArgs& a) {
    if (a.ws >= 128) return launchSlide<R, STEP, 4, 1>(h, a);
    return launchSlide<R, STEP, 2, 2>(h, a);
}

template <int R> int launchSlideR(hrb_ofc* h, const SearchArgs& a, int step) { return step == 1 ? launchSlideStep<R, 1>(h, a) : launchSlideStep<R, 0>(h, a); }

}  // namespace

// One whole pass for ws >= 64 at full flow resolution.  -1: this (R, geometry) is not covered here.
// Compiled three times (-DHRB_SLIDE_PART=0/1/2), four search radii per part, so the parts build in parallel.
#ifndef HRB_SLIDE_PART
#error "compile with -DHRB_SLIDE_PART=0, 1 or 2"
#endif
#define HRB_SLIDE_CONCAT2(a, b) a##b
#define HRB_SLIDE_CONCAT(a, b) HRB_SLIDE_CONCAT2(a, b)
int HRB_SLIDE_CONCAT(launchSearchPassSlidePart, HRB_SLIDE_PART)(hrb_ofc* h, const SearchArgs& a, int R, int step) {
    const int lu = step == 1 ? a.lw : a.lh;
    if (a.rs != 0 || a.ws < 64 || (lu & 3) != 0) return -1;
    switch (R) {
#define HRB_CASE(N) case N: return launchSlideR<N>(h, a, step);
#if HRB_SLIDE_PART == 0
        HRB_CASE(5) HRB_CASE(6) HRB_CASE(7) HRB_CASE(8)
#elif HRB_SLIDE_PART == 1
        HRB_CASE(9) HRB_CASE(10) HRB_CASE(11) HRB_CASE(12)
#else
        HRB_CASE(13) HRB_CASE(14) HRB_CASE(15) HRB_CASE(16)
#endif
#undef HRB_CASE
        default: return -1;  // not handled here
    }
}

}  // namespace hrb
