// hrb_internal.cuh — shared declarations of the hopperrender_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "hrb.h"

namespace hrb {

// ------------------------------------------------------------------------------------------------
// error plumbing (replaces CHECK_ERROR, HopperRender/opticalFlowCalc.h:15-22: no exceptions here,
// the C++ shim in include/opticalFlowCalc.h re-throws)
// ------------------------------------------------------------------------------------------------
void setLastError(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launchCount;  // kernels launched by this process (handles may live on several threads)
extern thread_local unsigned long long t_launchCount;     // ... by this thread: lets a graph capture count its own nodes
constexpr int HRB_MAX_DEVICES = 64;  // per-device launch configuration caches (power of two)

#define HRB_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            ::hrb::setLastError("[hopperrender_b200] CUDA error %d (%s) in %s at %s:%d", (int)_e,        \
                                cudaGetErrorString(_e), __func__, __FILE__, __LINE__);                   \
            return HRB_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

#define HRB_LAUNCH_CHECK()                                                                               \
    do {                                                                                                 \
        ::hrb::g_launchCount.fetch_add(1, std::memory_order_relaxed);                                     \
        ::hrb::t_launchCount++;                                                                          \
        HRB_CUDA(cudaGetLastError());                                                                    \
    } while (0)

// ------------------------------------------------------------------------------------------------
// geometry of one search pass (one (iteration, step) of HopperRender/opticalFlowCalcSDR.cpp:72-107)
// ------------------------------------------------------------------------------------------------
struct PassGeom {
    int windowSize;  // ws (power of two >= 2)
    int wsLog2;
    int iteration;
    int step;
    int nWx, nWy;        // windows of this level: ceil(lw/ws), ceil(lh/ws)
    int prevNWx, prevNWy;  // windows of the parent level (ws*2); 0 at iteration 0
};

// Arguments of the search kernels.  Window-level offset arrays replace the per-pixel offsetArray of the
// reference (offsets are constant inside each window).
struct SearchArgs {
    // 8-bit search planes (HDR samples >> 8, calcDeltaSumsKernelHDR.h:98-100) of frame N-1 (m_inputFrameArray[1], "1")
    // and frame N (m_inputFrameArray[2], "2"): luma [H][pitch] and the interleaved U,V plane [H/2][pitch] in the
    // NV12 layout; the T planes are the same data transposed, luma [W][pitchT] and chroma [W/2][pitchT] with
    // byte 2*(y>>1)+ch of row x>>1 = channel ch of chroma sample (y>>1, x>>1).  X steps read the T planes.
    const uint8_t *y1, *c1, *y2, *c2;
    const uint8_t *yT1, *cT1, *yT2, *cT2;
    int pitch;               // bytes per row of y / c (multiple of 128)
    int pitchT;              // bytes per row of yT / cT (multiple of 128)
    int W, H;                // frame size
    int lw, lh;              // flow size
    int rs;                  // resolution scalar
    int ws, wsLog2, iteration;
    int deltaScalar, neighborBiasScalar;
    int nWx, nWy, prevNWx;
    const int16_t* prevX;  // level i-1 offsets (nullptr at iteration 0: all zero)
    const int16_t* prevY;
    int16_t* curX;         // level i offsets: written by step 0, read by step 1
    int16_t* curY;         // level i offsets: written by step 1
    uint32_t* winSums;     // [nW][16] scratch for windows larger than one CTA tile
    unsigned* winTicket;   // [nW] arrival counters of the sliding kernels' fused finalize (zero between passes)
    uint32_t* rawDelta;    // non-null in pass 0: receives sums[R/2-1][window 0]
    uint32_t* tapSums;     // optional [R][nWy][nWx]
    uint8_t* tapLayer;     // optional [nWy][nWx]
    unsigned long long* dbg;  // optional per-CTA timeline of this pass (hrb_ofc_debug_timeline): 8 words per CTA
    bool winLanes;         // windows of 2 and 4: one lane per window (off: search variant 4, the butterfly form, for A/B runs)
};

// The search representation of one input slot: 8-bit planes in both orientations, one allocation.
struct SearchPlanes {
    uint8_t* base = nullptr;   // the allocation
    uint8_t *y = nullptr, *c = nullptr, *yT = nullptr, *cT = nullptr;
};

struct TmaCache;

struct PassTapDev {
    PassGeom g;
    uint32_t* sums = nullptr;
    uint8_t* layer = nullptr;
    int16_t* offX = nullptr;  // snapshot of the X offsets after the pass
    int16_t* offY = nullptr;
    int wsX = 0, nWxX = 0, nWyX = 0;  // level geometry of the X snapshot
    int wsY = 0, nWxY = 0, nWyY = 0;  // level geometry of the Y snapshot (0 = all zero)
};

struct Profile {
    bool on = false;
    struct Pending {
        cudaEvent_t a, b;
        int cls;
    };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    double ms[5] = {0, 0, 0, 0, 0};
    unsigned long long n[5] = {0, 0, 0, 0, 0};
};
enum { CLS_INGEST = 0, CLS_SEARCH = 1, CLS_BLUR = 2, CLS_WARP = 3, CLS_COPY = 4 };

}  // namespace hrb

// The handle.  Field names mirror HopperRender/opticalFlowCalc.h:26-78 where a counterpart exists.
struct hrb_ofc {
    // video properties
    int frameWidth, frameHeight, inputStride, outputStride;
    float outputBlackLevel, outputWhiteLevel;
    int hdr, bpp;
    // optical flow calculation
    int resScalar, flowWidth, flowHeight, searchRadius;
    double ofcCalcTime, ofcAvgCalcTime, ofcPeakCalcTime;
    int ofcCalcCount;
    double ofcCalcTimeSum, warpCalcTime;
    int deltaScalar, neighborBiasScalar;
    unsigned int totalFrameDelta, frameCount;

    // CUDA
    int device;
    cudaStream_t stream;
    bool ownStream;
    // one record per calculateOpticalFlow in flight: m_ofcStartedEvent / ofcEndEvent of the reference plus the
    // pinned word that receives the raw frame delta; a small ring lets the asynchronous path run ahead of the GPU
    struct FlowRecord {
        cudaEvent_t start, end;
        uint32_t* rawDeltaHost;
        bool startValid, pending;
    };
    static constexpr int kFlowRecords = 4;
    FlowRecord flowRec[kFlowRecords];
    int curRec;
    cudaEvent_t warpStartedEvent, warpEndEvent, uploadDoneEvent;
    bool warpStartedValid;
    // transfers run on their own streams so that the upload of frame N+1 and the downloads of the outputs of frame N
    // overlap the kernels: input slot [3] is the upload target (rotated in by updateFrame), the output is a ring of 3
    cudaStream_t upStream, downStream;
    cudaStream_t flowStream;             // asynchronous flow calculations run here, beside the warps of the same source frame
    cudaEvent_t flowForkEvent, flowJoinEvent;
    bool flowJoinPending, flowOverlap;
    int updatesSinceFlow = 0;      // update_frame calls since the flow in flight was enqueued (see finishUpdate)
    cudaEvent_t spareFreeEvent;          // compute stream: last readers of the buffer that became input slot [3] are enqueued
    static constexpr int kOutRing = 2 * HRB_WARP_BATCH_MAX;  // the batch being written plus the previous one still being downloaded
    cudaEvent_t outReady[kOutRing];      // compute stream: warp/copy into ring slot i finished
    cudaEvent_t outFree[kOutRing];       // download stream: D2H of ring slot i finished
    int outCur;                          // ring slot warpFrames / copyFrame write next (advances on download)
    int outView;                         // ring slot written most recently
    static constexpr int kTickets = 16;
    cudaEvent_t ticketEvent[kTickets];
    unsigned long long downloadSeq;      // downloads enqueued so far (ticket = sequence number, 1-based)

    // device arrays
    size_t inFrameBytes, outFrameBytes;
    uint8_t* inputFrameArray[4];   // raw NV12 / P010 frames, [0..2] rotated like m_inputFrameArray, [3] = upload target
    hrb::SearchPlanes searchPlane[4];  // 8-bit planar search representation of the same slot (row-major + transposed)
    int planePitch;                // bytes per row of the row-major planes
    int planePitchT;               // bytes per row of the transposed planes
    int stripeY0, stripeY1;        // output stripe (luma rows) warpFrames / copyFrame / downloadFrame work on; default the whole frame
    uint8_t* outputRing[kOutRing]; // m_outputFrameArray, as a ring so that downloads overlap the next warps
    int16_t* levelOffsets[2][2];   // [iteration parity][axis] window-level offsets
    size_t levelCapacity;          // entries per level array
    uint32_t* winSums;
    unsigned* winTicket;
    int16_t* offsetArrayScratch;   // [2][lh][lw], materialised on demand for taps
    int16_t* blurredOffsetArray[2];
    uint32_t* flowMaxDev[2];       // max |value| of blurredOffsetArray[i] (rotates with it): lets warpFrames skip the mirror in the interior
    uint32_t* rawDeltaDev;
    // final level geometry after the last calculate (input of blur / offset tap)
    int lastIterParity, lastNWx, lastNWy, lastWs;
    bool haveFlowLevels;

    // taps / profiling
    int searchVariant;  // 0: automatic kernel selection, 1: generic sadPassKernel for every pass, 2: sliding kernel without TMA, 3: ... without the aligned fast path
    int smCount;
    bool tapMode;
    // instantiated CUDA graphs of the flow calculation (search ladder + blur), one per combination of buffers and
    // parameters it was captured with: the input slots rotate with period 4, so a handful of graphs serve a stream
    struct FlowGraph {
        const void* plane1;
        const void* plane2;
        const void* blurOut;
        int R, deltaScalar, neighborBiasScalar, variant;
        cudaGraphExec_t exec;
        unsigned launches;
    };
    std::vector<FlowGraph> flowGraphs;
    bool flowGraphsOn;
    std::vector<hrb::PassTapDev> taps;
    hrb::Profile prof;
    hrb::TmaCache* tmaCache = nullptr;  // TMA descriptors of the search planes, encoded on first use
    struct SideBlob {
        uint8_t guid[16];
        std::vector<uint8_t> bytes;
    };
    std::vector<SideBlob> sideData;        // IMediaSideData blobs of the newest source frame (opaque, host memory)
    unsigned long long* dbgDev = nullptr;  // per-CTA timelines of the search passes (debug aid, off by default)
    size_t dbgStride = 0;                  // words per pass
};

namespace hrb {

// kernels_frame.cu
int launchPackFrame(hrb_ofc* h, int slot);
int launchCopyFrame(hrb_ofc* h, int slot);
int launchWarpFrames(hrb_ofc* h, int n, const float* t, uint8_t* const* out, int mode);
// kernels_search.cu
int launchSearchPass(hrb_ofc* h, const SearchArgs& a, int R, int step);
// kernels_search_slide.cu: HRB_OK, an error code, or -1 when this (R, geometry) is not covered
int launchSearchPassSlidePart0(hrb_ofc* h, const SearchArgs& a, int R, int step);  // R 5..8   (kernels_search_slide.cu, three builds)
int launchSearchPassSlidePart1(hrb_ofc* h, const SearchArgs& a, int R, int step);  // R 9..12
int launchSearchPassSlidePart2(hrb_ofc* h, const SearchArgs& a, int R, int step);  // R 13..16
inline int launchSearchPassSlide(hrb_ofc* h, const SearchArgs& a, int R, int step) {
    if (R < 5 || R > 16) return -1;
    return R <= 8 ? launchSearchPassSlidePart0(h, a, R, step) : R <= 12 ? launchSearchPassSlidePart1(h, a, R, step) : launchSearchPassSlidePart2(h, a, R, step);
}
void freeTmaCache(hrb_ofc* h);  // tensor maps of the handle's search planes (kernels_search_slide.cu)
// kernels_search_cand.cu: a whole pass for 2 <= ws <= 16 at full flow resolution; same return convention
int launchSearchPassCandPart0(hrb_ofc* h, const SearchArgs& a, int R, int step);  // R 5..8   (kernels_search_cand.cu, three builds)
int launchSearchPassCandPart1(hrb_ofc* h, const SearchArgs& a, int R, int step);  // R 9..12
int launchSearchPassCandPart2(hrb_ofc* h, const SearchArgs& a, int R, int step);  // R 13..16
inline int launchSearchPassCand(hrb_ofc* h, const SearchArgs& a, int R, int step) {
    if (R < 5 || R > 16) return -1;
    return R <= 8 ? launchSearchPassCandPart0(h, a, R, step) : R <= 12 ? launchSearchPassCandPart1(h, a, R, step) : launchSearchPassCandPart2(h, a, R, step);
}
int launchBlurFlow(hrb_ofc* h, const int16_t* lvlX, const int16_t* lvlY, int nWx, int wsLog2, int16_t* out, uint32_t* flowMax);
int launchExpandOffsets(hrb_ofc* h, const int16_t* lvlX, int nWxX, int wsLog2X, const int16_t* lvlY, int nWxY, int wsLog2Y,
                        int16_t* out);
int microbenchSad(int device, double* gigaAbsdiffPerSec);

// profiling helpers (hrb_api.cu)
void profBegin(hrb_ofc* h, int cls);
void profEnd(hrb_ofc* h, int cls, unsigned launches);

}  // namespace hrb
