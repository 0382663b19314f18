// hrb_api.cu — handle management, the per-frame host schedule and the C ABI (include/hrb.h).
// The schedule restates HopperRender/opticalFlowCalcSDR.cpp / opticalFlowCalcHDR.cpp on one CUDA stream.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "hrb_internal.cuh"

namespace hrb {

static thread_local std::string t_lastError;
std::atomic<unsigned long long> g_launchCount{0};
thread_local unsigned long long t_launchCount = 0;

void setLastError(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_lastError = buf;
    // the reference also reports on stderr (opticalFlowCalc.h:19-20)
    fputs(buf, stderr);
    fputc('\n', stderr);
}

// ---- profiling -----------------------------------------------------------------------------------
static cudaEvent_t profEvent(hrb_ofc* h) {
    if (!h->prof.pool.empty()) {
        cudaEvent_t e = h->prof.pool.back();
        h->prof.pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void profBegin(hrb_ofc* h, int cls) {
    if (!h->prof.on) return;
    Profile::Pending p;
    p.a = profEvent(h);
    p.b = nullptr;
    p.cls = cls;
    cudaEventRecord(p.a, h->stream);
    h->prof.pending.push_back(p);
}

void profEnd(hrb_ofc* h, int cls, unsigned launches) {
    h->prof.n[cls] += launches;
    if (!h->prof.on || h->prof.pending.empty()) return;
    Profile::Pending& p = h->prof.pending.back();
    p.b = profEvent(h);
    cudaEventRecord(p.b, h->stream);
}

static int profResolve(hrb_ofc* h) {
    if (h->prof.pending.empty()) return HRB_OK;
    HRB_CUDA(cudaStreamSynchronize(h->stream));
    for (auto& p : h->prof.pending) {
        if (p.a && p.b) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) h->prof.ms[p.cls] += ms;
        }
        if (p.a) h->prof.pool.push_back(p.a);
        if (p.b) h->prof.pool.push_back(p.b);
    }
    h->prof.pending.clear();
    return HRB_OK;
}

// ---- taps ----------------------------------------------------------------------------------------
static void freeTaps(hrb_ofc* h) {
    for (auto& t : h->taps) {
        cudaFree(t.sums);
        cudaFree(t.layer);
        cudaFree(t.offX);
        cudaFree(t.offY);
    }
    h->taps.clear();
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) ++l;
    return l;
}

// First window size and iteration count — opticalFlowCalcSDR.cpp:49-65 (NUM_ITERATIONS == 0)
static void ladder(int lw, int lh, int* ws0, int* iterations) {
    int windowSize = 1;
    int maxDim = lw > lh ? lw : lh;
    if (maxDim && !(maxDim & (maxDim - 1))) {
        windowSize = maxDim;
    } else {
        while (maxDim & (maxDim - 1)) maxDim &= (maxDim - 1);
        windowSize = maxDim << 1;
    }
    windowSize /= 2;
    *ws0 = windowSize;
    *iterations = ilog2(windowSize);
}

// The stats block at the end of calculateOpticalFlow — opticalFlowCalcSDR.cpp:119-138
static int resolveRecord(hrb_ofc* h, hrb_ofc::FlowRecord& r) {
    if (!r.pending) return HRB_OK;
    HRB_CUDA(cudaEventSynchronize(r.end));
    r.pending = false;
    // m_totalFrameDelta — opticalFlowCalcSDR.cpp:92-93 (divisor 10) / opticalFlowCalcHDR.cpp:92-93 (divisor 6)
    h->totalFrameDelta = *r.rawDeltaHost;
    h->totalFrameDelta /= (unsigned)(h->flowHeight * h->flowWidth * (h->hdr ? 6 : 10));
    float ms = 0;
    HRB_CUDA(cudaEventElapsedTime(&ms, r.start, r.end));
    h->ofcCalcTime = (double)ms / 1e3;
    if (h->ofcCalcCount >= 240) {  // CALC_TIME_INTERVAL, config.h:17
        h->ofcAvgCalcTime = h->ofcCalcTimeSum / h->ofcCalcCount;
        h->ofcCalcCount = 0;
        h->ofcCalcTimeSum = 0.0;
        h->ofcPeakCalcTime = h->ofcCalcTime;
    }
    h->ofcCalcCount++;
    h->ofcCalcTimeSum += h->ofcCalcTime;
    if (h->ofcCalcTime > h->ofcPeakCalcTime) h->ofcPeakCalcTime = h->ofcCalcTime;
    return HRB_OK;
}

// resolve every flow in flight, oldest first
static int resolveFlow(hrb_ofc* h) {
    for (int i = 1; i <= hrb_ofc::kFlowRecords; ++i) {
        const int rc = resolveRecord(h, h->flowRec[(h->curRec + i) % hrb_ofc::kFlowRecords]);
        if (rc) return rc;
    }
    return HRB_OK;
}

// The search ladder and the blur of one flow calculation.  issue = false only replays the host-side bookkeeping
// (the kernels are then launched from a captured graph).
static int issueFlowKernels(hrb_ofc* h, int R, int ws0, int iterations, bool issue) {
    const int lw = h->flowWidth, lh = h->flowHeight;
    SearchArgs a;
    memset(&a, 0, sizeof(a));
    const SearchPlanes& f1 = h->searchPlane[1];  // frame1 = m_inputFrameArray[1], opticalFlowCalcSDR.cpp:79
    const SearchPlanes& f2 = h->searchPlane[2];  // frame2 = m_inputFrameArray[2], opticalFlowCalcSDR.cpp:80
    a.y1 = f1.y; a.c1 = f1.c; a.yT1 = f1.yT; a.cT1 = f1.cT;
    a.y2 = f2.y; a.c2 = f2.c; a.yT2 = f2.yT; a.cT2 = f2.cT;
    a.pitch = h->planePitch;
    a.pitchT = h->planePitchT;
    a.W = h->frameWidth;
    a.H = h->frameHeight;
    a.lw = lw;
    a.lh = lh;
    a.rs = h->resScalar;
    a.deltaScalar = h->deltaScalar;
    a.neighborBiasScalar = h->neighborBiasScalar;
    a.winSums = h->winSums;
    a.winTicket = h->winTicket;

    int prevNWx = 0, prevNWy = 0;
    for (int iter = 0; iter < iterations; ++iter) {
        const int ws = ws0 >> iter;
        const int par = iter & 1;
        a.ws = ws;
        a.wsLog2 = ilog2(ws);
        a.iteration = iter;
        a.nWx = (lw + ws - 1) / ws;
        a.nWy = (lh + ws - 1) / ws;
        a.prevNWx = prevNWx;
        a.prevX = iter ? h->levelOffsets[par ^ 1][0] : nullptr;
        a.prevY = iter ? h->levelOffsets[par ^ 1][1] : nullptr;
        a.curX = h->levelOffsets[par][0];
        a.curY = h->levelOffsets[par][1];
        for (int step = 0; step < 2; ++step) {
            a.rawDelta = (iter == 0 && step == 0) ? h->rawDeltaDev : nullptr;
            a.tapSums = nullptr;
            a.tapLayer = nullptr;
            a.dbg = h->dbgDev ? h->dbgDev + (size_t)(iter * 2 + step) * h->dbgStride : nullptr;
            a.winLanes = h->searchVariant != 4;
            if (h->tapMode) {
                PassTapDev t;
                t.g.windowSize = ws;
                t.g.wsLog2 = a.wsLog2;
                t.g.iteration = iter;
                t.g.step = step;
                t.g.nWx = a.nWx;
                t.g.nWy = a.nWy;
                t.g.prevNWx = prevNWx;
                t.g.prevNWy = prevNWy;
                const size_t nW = (size_t)a.nWx * a.nWy;
                HRB_CUDA(cudaMalloc(&t.sums, nW * R * sizeof(uint32_t)));
                HRB_CUDA(cudaMalloc(&t.layer, nW));
                HRB_CUDA(cudaMalloc(&t.offX, nW * sizeof(int16_t)));
                HRB_CUDA(cudaMalloc(&t.offY, nW * sizeof(int16_t)));
                HRB_CUDA(cudaMemsetAsync(t.sums, 0, nW * R * sizeof(uint32_t), h->stream));
                a.tapSums = t.sums;
                a.tapLayer = t.layer;
                h->taps.push_back(t);
            }
            if (issue) {
                const int rc = launchSearchPass(h, a, R, step);
                if (rc) return rc;
            }
            if (h->tapMode) {
                PassTapDev& t = h->taps.back();
                const size_t nW = (size_t)a.nWx * a.nWy;
                HRB_CUDA(cudaMemcpyAsync(t.offX, a.curX, nW * sizeof(int16_t), cudaMemcpyDeviceToDevice, h->stream));
                t.wsX = ws;
                t.nWxX = a.nWx;
                t.nWyX = a.nWy;
                if (step == 1) {
                    HRB_CUDA(cudaMemcpyAsync(t.offY, a.curY, nW * sizeof(int16_t), cudaMemcpyDeviceToDevice, h->stream));
                    t.wsY = ws;
                    t.nWxY = a.nWx;
                    t.nWyY = a.nWy;
                } else if (iter > 0) {
                    HRB_CUDA(cudaMemcpyAsync(t.offY, a.prevY, (size_t)prevNWx * prevNWy * sizeof(int16_t), cudaMemcpyDeviceToDevice, h->stream));
                    t.wsY = ws * 2;
                    t.nWxY = prevNWx;
                    t.nWyY = prevNWy;
                } else {
                    t.wsY = 0;  // all zero
                }
            }
        }
        prevNWx = a.nWx;
        prevNWy = a.nWy;
        h->lastIterParity = par;
        h->lastNWx = a.nWx;
        h->lastNWy = a.nWy;
        h->lastWs = ws;
    }
    h->haveFlowLevels = true;

    // blur into m_blurredOffsetArray[0], then swap (opticalFlowCalcSDR.cpp:113-123)
    if (issue) {
        const int rc = launchBlurFlow(h, h->levelOffsets[h->lastIterParity][0], h->levelOffsets[h->lastIterParity][1], h->lastNWx, ilog2(h->lastWs),
                                      h->blurredOffsetArray[0], h->flowMaxDev[0]);
        if (rc) return rc;
    }
    return HRB_OK;
}

static void dropFlowGraphs(hrb_ofc* h) {
    for (auto& g : h->flowGraphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->flowGraphs.clear();
}

// Launches the flow kernels through a CUDA graph captured on first use for this combination of buffers and parameters
// (24 dependent launches at 4K become one); falls back to plain launches whenever capture is not possible.
static int launchFlowKernels(hrb_ofc* h, int R, int ws0, int iterations) {
    if (!h->flowGraphsOn || h->tapMode || h->prof.on || h->dbgDev) return issueFlowKernels(h, R, ws0, iterations, true);
    hrb_ofc::FlowGraph key;
    key.plane1 = h->searchPlane[1].base;
    key.plane2 = h->searchPlane[2].base;
    key.blurOut = h->blurredOffsetArray[0];
    key.R = R;
    key.deltaScalar = h->deltaScalar;
    key.neighborBiasScalar = h->neighborBiasScalar;
    key.variant = h->searchVariant;
    hrb_ofc::FlowGraph* hit = nullptr;
    for (auto& g : h->flowGraphs)
        if (g.plane1 == key.plane1 && g.plane2 == key.plane2 && g.blurOut == key.blurOut && g.R == key.R && g.deltaScalar == key.deltaScalar &&
            g.neighborBiasScalar == key.neighborBiasScalar && g.variant == key.variant)
            hit = &g;
    if (!hit) {
        if (h->flowGraphs.size() >= 64) dropFlowGraphs(h);  // 12 radii x 4 slot rotations fit; beyond that parameters keep changing (UI): start over
        const unsigned long long before = t_launchCount;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            h->flowGraphsOn = false;  // e.g. the caller's stream is the legacy default stream
            return issueFlowKernels(h, R, ws0, iterations, true);
        }
        const int rc = issueFlowKernels(h, R, ws0, iterations, true);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        key.launches = (unsigned)(t_launchCount - before);
        g_launchCount.fetch_sub(key.launches, std::memory_order_relaxed);  // captured, not run yet
        key.exec = nullptr;
        if (rc == HRB_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&key.exec, graph, 0) == cudaSuccess) {
            cudaGraphDestroy(graph);
            h->flowGraphs.push_back(key);
            hit = &h->flowGraphs.back();
        } else {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            h->flowGraphsOn = false;
            if (rc) return rc;
            return issueFlowKernels(h, R, ws0, iterations, true);
        }
    } else {
        const int rc = issueFlowKernels(h, R, ws0, iterations, false);
        if (rc) return rc;
    }
    HRB_CUDA(cudaGraphLaunch(hit->exec, h->stream));
    g_launchCount.fetch_add(hit->launches, std::memory_order_relaxed);
    return HRB_OK;
}

static int enqueueFlow(hrb_ofc* h) {
    HRB_CUDA(cudaSetDevice(h->device));
    hrb_ofc::FlowRecord& rec = h->flowRec[h->curRec];
    if (rec.pending) {  // a second calculate on the same upload: its end event is about to be re-recorded
        const int rc = resolveFlow(h);
        if (rc) return rc;
    }
    const int R = h->searchRadius;  // m_lowGrid8x8xL[2] = m_opticalFlowSearchRadius, opticalFlowCalcSDR.cpp:46
    if (R < 2 || R > 16) {
        setLastError("[hopperrender_b200] search radius %d outside 2..16", R);
        return HRB_ERR_INVALID_ARG;
    }
    if (h->deltaScalar < 0 || h->deltaScalar > 31 || h->neighborBiasScalar < 0 || h->neighborBiasScalar > 31) {
        setLastError("[hopperrender_b200] delta/neighbor scalar outside 0..31");
        return HRB_ERR_INVALID_ARG;
    }
    const int lw = h->flowWidth, lh = h->flowHeight;
    int ws0, iterations;
    ladder(lw, lh, &ws0, &iterations);
    if (!rec.startValid) {  // calculate without a preceding updateFrame: time from here
        HRB_CUDA(cudaEventRecord(rec.start, h->stream));
        rec.startValid = true;
    }
    if (h->tapMode) freeTaps(h);

    {
        const int rc = launchFlowKernels(h, R, ws0, iterations);
        if (rc) return rc;
    }
    HRB_CUDA(cudaMemcpyAsync(rec.rawDeltaHost, h->rawDeltaDev, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    HRB_CUDA(cudaEventRecord(rec.end, h->stream));
    int16_t* t0 = h->blurredOffsetArray[0];
    h->blurredOffsetArray[0] = h->blurredOffsetArray[1];
    h->blurredOffsetArray[1] = t0;
    uint32_t* m0 = h->flowMaxDev[0];
    h->flowMaxDev[0] = h->flowMaxDev[1];
    h->flowMaxDev[1] = m0;
    rec.pending = true;
    return HRB_OK;
}

static int finishUpdate(hrb_ofc* h) {
    // the new frame sits in slot [3]; rotate [0] <- [1] <- [2] <- [3], the old [0] becomes the next upload target
    // (same result as the reference's write-into-[0]-then-rotate, opticalFlowCalcSDR.cpp:20-26)
    auto rot = [](auto** v) {
        auto* f0 = v[0];
        v[0] = v[1];
        v[1] = v[2];
        v[2] = v[3];
        v[3] = f0;
    };
    rot(h->inputFrameArray);
    {
        const SearchPlanes s0 = h->searchPlane[0];
        h->searchPlane[0] = h->searchPlane[1];
        h->searchPlane[1] = h->searchPlane[2];
        h->searchPlane[2] = h->searchPlane[3];
        h->searchPlane[3] = s0;
    }
    h->frameCount++;
    // every reader of the buffer that just became slot [3] (the warps of the previous source frame) is already enqueued
    HRB_CUDA(cudaEventRecord(h->spareFreeEvent, h->stream));
    // an asynchronous flow still in flight reads the search planes of slots [1] and [2] as they were when it was
    // enqueued; the third update after it writes one of them and is ordered behind it.  The first two do not touch
    // them, so in the usual update / calculate alternation the ingest of frame N+1 runs beside the search of frame N.
    if (h->flowJoinPending && ++h->updatesSinceFlow >= 3) HRB_CUDA(cudaStreamWaitEvent(h->stream, h->flowJoinEvent, 0));
    return launchPackFrame(h, 2);
}

// `onStream`: the stream the frame copy will be enqueued on; the flow timer starts there (m_ofcStartedEvent is the
// upload event in the reference, opticalFlowCalcSDR.cpp:20)
static int beginUpdate(hrb_ofc* h, cudaStream_t onStream) {
    HRB_CUDA(cudaSetDevice(h->device));
    h->curRec = (h->curRec + 1) % hrb_ofc::kFlowRecords;
    hrb_ofc::FlowRecord& rec = h->flowRec[h->curRec];
    if (rec.pending) {  // the ring is full: wait for the oldest flow before reusing its events
        const int rc = resolveRecord(h, rec);
        if (rc) return rc;
    }
    if (onStream != h->stream) HRB_CUDA(cudaStreamWaitEvent(onStream, h->spareFreeEvent, 0));
    HRB_CUDA(cudaEventRecord(rec.start, onStream));
    rec.startValid = true;
    return HRB_OK;
}

static int uploadFrame(hrb_ofc* h, const uint8_t* input_planes, bool wait) {
    int rc = beginUpdate(h, h->upStream);
    if (rc) return rc;
    // write into the upload slot on the upload stream (opticalFlowCalcSDR.cpp:20), then hand it to the compute stream
    HRB_CUDA(cudaMemcpyAsync(h->inputFrameArray[3], input_planes, h->inFrameBytes, cudaMemcpyHostToDevice, h->upStream));
    HRB_CUDA(cudaEventRecord(h->uploadDoneEvent, h->upStream));
    HRB_CUDA(cudaStreamWaitEvent(h->stream, h->uploadDoneEvent, 0));
    rc = finishUpdate(h);
    if (rc) return rc;
    if (wait) HRB_CUDA(cudaEventSynchronize(h->uploadDoneEvent));  // the caller's buffer is free again; everything else continues asynchronously
    return HRB_OK;
}

static int enqueueDownload(hrb_ofc* h, uint8_t* dst, unsigned long long* ticket) {
    HRB_CUDA(cudaSetDevice(h->device));
    const int slot = h->outCur;
    HRB_CUDA(cudaStreamWaitEvent(h->downStream, h->outReady[slot], 0));
    if (h->stripeY0 == 0 && h->stripeY1 == h->frameHeight) {
        HRB_CUDA(cudaMemcpyAsync(dst, h->outputRing[slot], h->outFrameBytes, cudaMemcpyDeviceToHost, h->downStream));
    } else {
        // only the stripe's luma rows and their chroma rows travel, to the same offsets of the caller's full-frame buffer
        const size_t rowBytes = (size_t)h->outputStride * h->bpp;
        const size_t lumaOff = (size_t)h->stripeY0 * rowBytes, lumaLen = (size_t)(h->stripeY1 - h->stripeY0) * rowBytes;
        const size_t chromaOff = ((size_t)h->frameHeight + (h->stripeY0 >> 1)) * rowBytes, chromaLen = lumaLen / 2;
        HRB_CUDA(cudaMemcpyAsync(dst + lumaOff, h->outputRing[slot] + lumaOff, lumaLen, cudaMemcpyDeviceToHost, h->downStream));
        HRB_CUDA(cudaMemcpyAsync(dst + chromaOff, h->outputRing[slot] + chromaOff, chromaLen, cudaMemcpyDeviceToHost, h->downStream));
    }
    HRB_CUDA(cudaEventRecord(h->outFree[slot], h->downStream));
    const unsigned long long seq = ++h->downloadSeq;
    HRB_CUDA(cudaEventRecord(h->ticketEvent[seq % hrb_ofc::kTickets], h->downStream));
    if (ticket) *ticket = seq;
    h->outCur = (slot + 1) % hrb_ofc::kOutRing;
    return HRB_OK;
}

// warpFrames / copyFrame are about to overwrite ring slot outCur: its previous download must have finished
static int beginOutput(hrb_ofc* h) {
    HRB_CUDA(cudaSetDevice(h->device));
    HRB_CUDA(cudaStreamWaitEvent(h->stream, h->outFree[h->outCur], 0));
    HRB_CUDA(cudaEventRecord(h->warpStartedEvent, h->stream));  // m_warpStartedEvent, opticalFlowCalcSDR.cpp:164
    h->warpStartedValid = true;
    h->outView = h->outCur;
    return HRB_OK;
}

}  // namespace hrb

using namespace hrb;

#define HRB_REQUIRE(cond, msg)                                       \
    do {                                                             \
        if (!(cond)) {                                               \
            setLastError("[hopperrender_b200] %s: %s", __func__, msg); \
            return HRB_ERR_INVALID_ARG;                              \
        }                                                            \
    } while (0)

extern "C" {

int hrb_ofc_create(hrb_ofc** out, const hrb_ofc_desc* d) {
    HRB_REQUIRE(out && d, "null argument");
    *out = nullptr;
    HRB_REQUIRE(d->frame_width >= 16 && d->frame_height >= 16, "frame must be at least 16x16");
    HRB_REQUIRE((d->frame_width % 2) == 0 && (d->frame_height % 2) == 0, "NV12/P010 frame dimensions must be even");
    HRB_REQUIRE(d->max_calc_res >= 1, "max_calc_res must be positive");
    HRB_REQUIRE(d->input_stride <= 0 || d->input_stride >= d->frame_width, "stride smaller than the frame width");
    HRB_REQUIRE(d->output_stride <= 0 || d->output_stride >= d->frame_width, "stride smaller than the frame width");
    {
        int rs = 0;
        while ((d->frame_height >> rs) > d->max_calc_res) rs++;
        HRB_REQUIRE((d->frame_width >> rs) >= 4 && (d->frame_height >> rs) >= 4, "flow resolution below 4x4 (raise max_calc_res)");
    }
    HRB_REQUIRE(d->delta_scalar >= 0 && d->delta_scalar <= 31 && d->neighbor_scalar >= 0 && d->neighbor_scalar <= 31, "delta/neighbor scalar outside 0..31");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) {
        setLastError("[hopperrender_b200] no CUDA device (this library has no CPU fallback)");
        return HRB_ERR_NO_DEVICE;
    }
    HRB_REQUIRE(d->device_ordinal >= 0 && d->device_ordinal < nDev, "device_ordinal out of range");
    HRB_CUDA(cudaSetDevice(d->device_ordinal));

    hrb_ofc* h = new (std::nothrow) hrb_ofc();
    HRB_REQUIRE(h, "out of host memory");
    // opticalFlowCalcSDR.cpp:208-232
    h->frameWidth = d->frame_width;
    h->frameHeight = d->frame_height;
    h->inputStride = d->input_stride > 0 ? d->input_stride : d->frame_width;
    h->outputStride = d->output_stride > 0 ? d->output_stride : d->frame_width;
    h->outputBlackLevel = d->black_level;
    h->outputWhiteLevel = d->white_level;
    h->hdr = d->is_hdr ? 1 : 0;
    h->bpp = h->hdr ? 2 : 1;
    h->searchRadius = 5;  // MIN_SEARCH_RADIUS, config.h:8
    h->resScalar = 0;
    while ((d->frame_height >> h->resScalar) > d->max_calc_res) h->resScalar++;
    h->flowWidth = (int)std::ceil(h->frameWidth / std::pow(2, h->resScalar));
    h->flowHeight = (int)std::ceil(h->frameHeight / std::pow(2, h->resScalar));
    h->ofcCalcTime = h->ofcAvgCalcTime = h->ofcPeakCalcTime = 0.0;
    h->ofcCalcCount = 0;
    h->ofcCalcTimeSum = 0.0;
    h->warpCalcTime = 0.0;
    h->deltaScalar = d->delta_scalar;
    h->neighborBiasScalar = d->neighbor_scalar;
    h->totalFrameDelta = 0;
    h->frameCount = 0;
    h->device = d->device_ordinal;
    h->stream = nullptr;
    h->ownStream = false;
    h->warpStartedValid = false;
    h->curRec = 0;
    for (auto& r : h->flowRec) {
        r.start = r.end = nullptr;
        r.rawDeltaHost = nullptr;
        r.startValid = r.pending = false;
    }
    h->haveFlowLevels = false;
    h->tapMode = false;
    h->flowGraphsOn = true;
    h->searchVariant = 0;
    h->smCount = 148;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d->device_ordinal) == cudaSuccess) h->smCount = prop.multiProcessorCount;
    }
    h->lastIterParity = 0;
    h->lastNWx = h->lastNWy = h->lastWs = 0;
    for (int i = 0; i < 4; ++i) {
        h->inputFrameArray[i] = nullptr;
        h->searchPlane[i] = SearchPlanes();
    }
    for (int i = 0; i < hrb_ofc::kOutRing; ++i) {
        h->outputRing[i] = nullptr;
        h->outReady[i] = h->outFree[i] = nullptr;
    }
    for (auto& e : h->ticketEvent) e = nullptr;
    h->upStream = h->downStream = h->flowStream = nullptr;
    h->flowForkEvent = h->flowJoinEvent = nullptr;
    h->spareFreeEvent = nullptr;
    h->outCur = h->outView = 0;
    h->downloadSeq = 0;
    h->stripeY0 = 0;
    h->stripeY1 = d->frame_height;
    h->levelOffsets[0][0] = h->levelOffsets[0][1] = h->levelOffsets[1][0] = h->levelOffsets[1][1] = nullptr;
    h->winSums = nullptr;
    h->winTicket = nullptr;
    h->offsetArrayScratch = nullptr;
    h->blurredOffsetArray[0] = h->blurredOffsetArray[1] = nullptr;
    h->flowMaxDev[0] = h->flowMaxDev[1] = nullptr;
    h->rawDeltaDev = nullptr;
    h->warpStartedEvent = h->warpEndEvent = h->uploadDoneEvent = nullptr;

    auto fail = [&](int rc) {
        hrb_ofc_destroy(h);
        return rc;
    };
    if (h->inputStride < h->frameWidth || h->outputStride < h->frameWidth) {
        setLastError("[hopperrender_b200] hrb_ofc_create: stride smaller than the frame width");
        return fail(HRB_ERR_INVALID_ARG);
    }
    if (h->flowWidth < 4 || h->flowHeight < 4) {
        setLastError("[hopperrender_b200] hrb_ofc_create: flow resolution %dx%d is below 4x4 (raise max_calc_res)", h->flowWidth, h->flowHeight);
        return fail(HRB_ERR_INVALID_ARG);
    }

    const size_t lw = h->flowWidth, lh = h->flowHeight;
    h->inFrameBytes = ((size_t)h->frameHeight * h->inputStride + (size_t)(h->frameHeight / 2) * h->inputStride) * h->bpp;
    h->outFrameBytes = ((size_t)h->frameHeight * h->outputStride + (size_t)(h->frameHeight / 2) * h->outputStride) * h->bpp;
    // planar 8-bit search planes: luma + NV12-style chroma, row-major and transposed (rows padded to 128 bytes: the
    // TMA row stride must be a multiple of 16 and a warp's 128-byte row segment never leaves the allocation)
    h->planePitch = (h->frameWidth + 127) & ~127;
    h->planePitchT = (h->frameHeight + 127) & ~127;
    const size_t planeYBytes = (size_t)h->planePitch * h->frameHeight, planeCBytes = (size_t)h->planePitch * (h->frameHeight / 2);
    const size_t planeYTBytes = (size_t)h->planePitchT * h->frameWidth, planeCTBytes = (size_t)h->planePitchT * (h->frameWidth / 2);
    const size_t planeBytes = planeYBytes + planeCBytes + 256, planeTBytes = planeYTBytes + planeCTBytes + 256;
    h->levelCapacity = ((lw + 1) / 2) * ((lh + 1) / 2);
    const size_t winSumEntries = ((lw + 63) / 64) * ((lh + 63) / 64) * 16 + 16;
    const size_t need = 4 * (h->inFrameBytes + planeBytes + planeTBytes) + hrb_ofc::kOutRing * h->outFrameBytes + 4 * h->levelCapacity * 2 + winSumEntries * 4 + 3 * 2 * lw * lh * 2;

    // replaces detectDevices' memory check (opticalFlowCalc.cpp:48-51,86-96)
    size_t freeB = 0, totalB = 0;
    if (cudaMemGetInfo(&freeB, &totalB) != cudaSuccess) {
        setLastError("[hopperrender_b200] device %d is not usable: %s", h->device, cudaGetErrorString(cudaGetLastError()));
        return fail(HRB_ERR_CUDA);
    }
    if (freeB < need + (64u << 20)) {
        setLastError("[hopperrender_b200] device %d has %zu MiB free, %zu MiB needed", h->device, freeB >> 20, need >> 20);
        return fail(HRB_ERR_NO_DEVICE);
    }

#define HRB_TRY(call)                                                                                                      \
    do {                                                                                                                   \
        cudaError_t _e = (call);                                                                                           \
        if (_e != cudaSuccess) {                                                                                           \
            setLastError("[hopperrender_b200] CUDA error %d (%s) in hrb_ofc_create at %s:%d", (int)_e, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                                              \
            return fail(HRB_ERR_CUDA);                                                                                     \
        }                                                                                                                  \
    } while (0)

    if (d->cuda_stream) {
        h->stream = (cudaStream_t)d->cuda_stream;
    } else {
        HRB_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->ownStream = true;
    }
    for (auto& r : h->flowRec) {
        HRB_TRY(cudaEventCreate(&r.start));
        HRB_TRY(cudaEventCreate(&r.end));
        HRB_TRY(cudaMallocHost(&r.rawDeltaHost, sizeof(uint32_t)));
        *r.rawDeltaHost = 0;
    }
    HRB_TRY(cudaEventCreate(&h->warpStartedEvent));
    HRB_TRY(cudaEventCreate(&h->warpEndEvent));
    HRB_TRY(cudaEventCreateWithFlags(&h->uploadDoneEvent, cudaEventBlockingSync | cudaEventDisableTiming));
    HRB_TRY(cudaStreamCreateWithFlags(&h->upStream, cudaStreamNonBlocking));
    {
        // the search ladder is the critical path of a source frame: its CTAs go first when warps compete for the SMs
        int prioLow = 0, prioHigh = 0;
        HRB_TRY(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
        HRB_TRY(cudaStreamCreateWithPriority(&h->flowStream, cudaStreamNonBlocking, prioHigh));
    }
    HRB_TRY(cudaEventCreateWithFlags(&h->flowForkEvent, cudaEventDisableTiming));
    HRB_TRY(cudaEventCreateWithFlags(&h->flowJoinEvent, cudaEventDisableTiming));
    h->flowJoinPending = false;
    h->flowOverlap = true;
    HRB_TRY(cudaStreamCreateWithFlags(&h->downStream, cudaStreamNonBlocking));
    HRB_TRY(cudaEventCreateWithFlags(&h->spareFreeEvent, cudaEventDisableTiming));
    for (int i = 0; i < hrb_ofc::kOutRing; ++i) {
        HRB_TRY(cudaEventCreateWithFlags(&h->outReady[i], cudaEventDisableTiming));
        HRB_TRY(cudaEventCreate(&h->outFree[i]));
        HRB_TRY(cudaMalloc(&h->outputRing[i], h->outFrameBytes));
        HRB_TRY(cudaMemsetAsync(h->outputRing[i], 0, h->outFrameBytes, h->stream));
    }
    for (auto& e : h->ticketEvent) HRB_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) {
        HRB_TRY(cudaMalloc(&h->inputFrameArray[i], h->inFrameBytes));
        HRB_TRY(cudaMemsetAsync(h->inputFrameArray[i], 0, h->inFrameBytes, h->stream));
        SearchPlanes& sp = h->searchPlane[i];
        HRB_TRY(cudaMalloc(&sp.base, planeBytes + planeTBytes));
        HRB_TRY(cudaMemsetAsync(sp.base, 0, planeBytes + planeTBytes, h->stream));
        sp.y = sp.base;                       // cudaMalloc returns 256-byte aligned memory; every plane starts 128-byte aligned
        sp.c = sp.y + planeYBytes;
        sp.yT = sp.c + planeCBytes + 256;       // plane sizes are multiples of 128 (padded pitches)
        sp.cT = sp.yT + planeYTBytes;
    }
    for (int p = 0; p < 2; ++p)
        for (int ax = 0; ax < 2; ++ax) {
            HRB_TRY(cudaMalloc(&h->levelOffsets[p][ax], h->levelCapacity * sizeof(int16_t)));
            HRB_TRY(cudaMemsetAsync(h->levelOffsets[p][ax], 0, h->levelCapacity * sizeof(int16_t), h->stream));
        }
    HRB_TRY(cudaMalloc(&h->winSums, winSumEntries * sizeof(uint32_t)));
    HRB_TRY(cudaMemsetAsync(h->winSums, 0, winSumEntries * sizeof(uint32_t), h->stream));
    HRB_TRY(cudaMalloc(&h->winTicket, (winSumEntries / 16 + 1) * sizeof(unsigned)));
    HRB_TRY(cudaMemsetAsync(h->winTicket, 0, (winSumEntries / 16 + 1) * sizeof(unsigned), h->stream));
    HRB_TRY(cudaMalloc(&h->offsetArrayScratch, 2 * lw * lh * sizeof(int16_t)));
    for (int i = 0; i < 2; ++i) {
        HRB_TRY(cudaMalloc(&h->blurredOffsetArray[i], 2 * lw * lh * sizeof(int16_t)));
        HRB_TRY(cudaMemsetAsync(h->blurredOffsetArray[i], 0, 2 * lw * lh * sizeof(int16_t), h->stream));
        HRB_TRY(cudaMalloc(&h->flowMaxDev[i], sizeof(uint32_t)));
        HRB_TRY(cudaMemsetAsync(h->flowMaxDev[i], 0, sizeof(uint32_t), h->stream));
    }
    HRB_TRY(cudaMalloc(&h->rawDeltaDev, sizeof(uint32_t)));
    HRB_TRY(cudaMemsetAsync(h->rawDeltaDev, 0, sizeof(uint32_t), h->stream));
    HRB_TRY(cudaStreamSynchronize(h->stream));
#undef HRB_TRY
    *out = h;
    return HRB_OK;
}

void hrb_ofc_destroy(hrb_ofc* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);  // clFinish, opticalFlowCalcSDR.cpp:186
    if (h->upStream) cudaStreamSynchronize(h->upStream);
    if (h->flowStream) cudaStreamSynchronize(h->flowStream);
    if (h->downStream) cudaStreamSynchronize(h->downStream);
    freeTaps(h);
    for (auto& p : h->prof.pending) {
        if (p.a) cudaEventDestroy(p.a);
        if (p.b) cudaEventDestroy(p.b);
    }
    for (auto e : h->prof.pool) cudaEventDestroy(e);
    for (int i = 0; i < 4; ++i) {
        cudaFree(h->inputFrameArray[i]);
        cudaFree(h->searchPlane[i].base);
    }
    for (int i = 0; i < hrb_ofc::kOutRing; ++i) {
        cudaFree(h->outputRing[i]);
        if (h->outReady[i]) cudaEventDestroy(h->outReady[i]);
        if (h->outFree[i]) cudaEventDestroy(h->outFree[i]);
    }
    for (auto e : h->ticketEvent)
        if (e) cudaEventDestroy(e);
    if (h->spareFreeEvent) cudaEventDestroy(h->spareFreeEvent);
    dropFlowGraphs(h);
    freeTmaCache(h);
    cudaFree(h->dbgDev);
    if (h->upStream) cudaStreamDestroy(h->upStream);
    if (h->flowStream) cudaStreamDestroy(h->flowStream);
    if (h->flowForkEvent) cudaEventDestroy(h->flowForkEvent);
    if (h->flowJoinEvent) cudaEventDestroy(h->flowJoinEvent);
    if (h->downStream) cudaStreamDestroy(h->downStream);
    for (int p = 0; p < 2; ++p)
        for (int ax = 0; ax < 2; ++ax) cudaFree(h->levelOffsets[p][ax]);
    cudaFree(h->winSums);
    cudaFree(h->winTicket);
    cudaFree(h->offsetArrayScratch);
    cudaFree(h->blurredOffsetArray[0]);
    cudaFree(h->blurredOffsetArray[1]);
    cudaFree(h->flowMaxDev[0]);
    cudaFree(h->flowMaxDev[1]);
    cudaFree(h->rawDeltaDev);
    for (auto& r : h->flowRec) {
        if (r.rawDeltaHost) cudaFreeHost(r.rawDeltaHost);
        if (r.start) cudaEventDestroy(r.start);
        if (r.end) cudaEventDestroy(r.end);
    }
    if (h->warpStartedEvent) cudaEventDestroy(h->warpStartedEvent);
    if (h->warpEndEvent) cudaEventDestroy(h->warpEndEvent);
    if (h->uploadDoneEvent) cudaEventDestroy(h->uploadDoneEvent);
    if (h->ownStream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int hrb_ofc_update_frame(hrb_ofc* h, const uint8_t* input_planes) {
    HRB_REQUIRE(h && input_planes, "null argument");
    return uploadFrame(h, input_planes, true);
}

int hrb_ofc_update_frame_async(hrb_ofc* h, const uint8_t* pinned_input_planes) {
    HRB_REQUIRE(h && pinned_input_planes, "null argument");
    return uploadFrame(h, pinned_input_planes, false);
}

int hrb_ofc_wait_upload(hrb_ofc* h) {
    HRB_REQUIRE(h, "null handle");
    HRB_CUDA(cudaSetDevice(h->device));
    HRB_CUDA(cudaEventSynchronize(h->uploadDoneEvent));
    return HRB_OK;
}

int hrb_ofc_update_frame_device(hrb_ofc* h, const void* device_planes) {
    HRB_REQUIRE(h && device_planes, "null argument");
    int rc = beginUpdate(h, h->stream);
    if (rc) return rc;
    HRB_CUDA(cudaMemcpyAsync(h->inputFrameArray[3], device_planes, h->inFrameBytes, cudaMemcpyDeviceToDevice, h->stream));
    return finishUpdate(h);
}

// Orders the compute stream behind the last flow calculation that was enqueued on the flow stream.
static int joinFlow(hrb_ofc* h) {
    if (h->flowJoinPending) {
        HRB_CUDA(cudaSetDevice(h->device));
        HRB_CUDA(cudaStreamWaitEvent(h->stream, h->flowJoinEvent, 0));
        h->flowJoinPending = false;
    }
    return HRB_OK;
}

// warpFrames reads the flow of the PREVIOUS calculation (m_blurredOffsetArray12[0] after the swap,
// opticalFlowCalcSDR.cpp:113-123,150) and the search touches nothing the warp writes, so the calculation of source
// frame N and the warps issued after it are independent: the asynchronous entry point runs the search ladder on its
// own stream, forked behind everything already enqueued (the ingest of frame N, the last readers of the flow buffer
// it overwrites) and joined before the next calculation — i.e. before any warp that needs its result.
int hrb_ofc_calculate_optical_flow_async(hrb_ofc* h) {
    HRB_REQUIRE(h, "null handle");
    int rc = joinFlow(h);
    if (rc) return rc;
    if (!h->flowOverlap || h->tapMode || h->prof.on) return enqueueFlow(h);
    HRB_CUDA(cudaSetDevice(h->device));
    HRB_CUDA(cudaEventRecord(h->flowForkEvent, h->stream));
    HRB_CUDA(cudaStreamWaitEvent(h->flowStream, h->flowForkEvent, 0));
    cudaStream_t compute = h->stream;
    h->stream = h->flowStream;
    rc = enqueueFlow(h);
    h->stream = compute;
    if (rc) return rc;
    HRB_CUDA(cudaEventRecord(h->flowJoinEvent, h->flowStream));
    h->flowJoinPending = true;
    h->updatesSinceFlow = 0;
    return HRB_OK;
}

int hrb_ofc_calculate_optical_flow(hrb_ofc* h) {
    HRB_REQUIRE(h, "null handle");
    int rc = joinFlow(h);
    if (rc) return rc;
    rc = enqueueFlow(h);
    if (rc) return rc;
    return resolveFlow(h);
}

int hrb_ofc_warp_frames(hrb_ofc* h, float blending_scalar, int frame_output_mode) {
    HRB_REQUIRE(h, "null handle");
    if (blending_scalar > 1.0f) {  // opticalFlowCalcSDR.cpp:143-146
        setLastError("[HopperRender] Error in function warpFrames: Blending scalar is greater than 1.0");
        return HRB_ERR_BLEND_RANGE;
    }
    HRB_REQUIRE(frame_output_mode >= 0 && frame_output_mode <= 6, "frame_output_mode outside 0..6");
    int rc = beginOutput(h);
    if (rc) return rc;
    uint8_t* out = h->outputRing[h->outCur];
    rc = launchWarpFrames(h, 1, &blending_scalar, &out, frame_output_mode);
    if (rc) return rc;
    HRB_CUDA(cudaEventRecord(h->outReady[h->outCur], h->stream));
    return HRB_OK;
}

int hrb_ofc_warp_frames_batch(hrb_ofc* h, int n, const float* blending_scalars, int frame_output_mode) {
    HRB_REQUIRE(h && blending_scalars, "null argument");
    HRB_REQUIRE(n >= 1 && n <= HRB_WARP_BATCH_MAX, "batch size outside 1..HRB_WARP_BATCH_MAX");
    for (int i = 0; i < n; ++i)
        if (blending_scalars[i] > 1.0f) {  // opticalFlowCalcSDR.cpp:143-146
            setLastError("[HopperRender] Error in function warpFrames: Blending scalar is greater than 1.0");
            return HRB_ERR_BLEND_RANGE;
        }
    HRB_REQUIRE(frame_output_mode >= 0 && frame_output_mode <= 6, "frame_output_mode outside 0..6");
    int rc = beginOutput(h);  // waits for the download of the first slot, starts the warp timer
    if (rc) return rc;
    uint8_t* outs[HRB_WARP_BATCH_MAX];
    for (int i = 0; i < n; ++i) {
        const int slot = (h->outCur + i) % hrb_ofc::kOutRing;
        if (i) HRB_CUDA(cudaStreamWaitEvent(h->stream, h->outFree[slot], 0));
        outs[i] = h->outputRing[slot];
    }
    rc = launchWarpFrames(h, n, blending_scalars, outs, frame_output_mode);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) HRB_CUDA(cudaEventRecord(h->outReady[(h->outCur + i) % hrb_ofc::kOutRing], h->stream));
    return HRB_OK;
}

int hrb_ofc_copy_frame(hrb_ofc* h) {
    HRB_REQUIRE(h, "null handle");
    const int frameIndex = h->frameCount >= 3 ? 0 : h->frameCount >= 2 ? 1 : 2;  // opticalFlowCalcSDR.cpp:173
    int rc = beginOutput(h);
    if (rc) return rc;
    rc = launchCopyFrame(h, frameIndex);
    if (rc) return rc;
    HRB_CUDA(cudaEventRecord(h->outReady[h->outCur], h->stream));
    return HRB_OK;
}

int hrb_ofc_download_frame(hrb_ofc* h, uint8_t* output_planes) {
    HRB_REQUIRE(h && output_planes, "null argument");
    const int slot = h->outCur;
    const int rc = enqueueDownload(h, output_planes, nullptr);
    if (rc) return rc;
    HRB_CUDA(cudaEventSynchronize(h->outFree[slot]));  // blocking read, opticalFlowCalcSDR.cpp:33
    if (h->warpStartedValid) {  // opticalFlowCalcSDR.cpp:36-41
        float ms = 0;
        HRB_CUDA(cudaEventElapsedTime(&ms, h->warpStartedEvent, h->outFree[slot]));
        h->warpCalcTime = (double)ms / 1e3;
    }
    return HRB_OK;
}

int hrb_ofc_download_frame_async(hrb_ofc* h, uint8_t* pinned_output_planes, unsigned long long* ticket) {
    HRB_REQUIRE(h && pinned_output_planes, "null argument");
    return enqueueDownload(h, pinned_output_planes, ticket);
}

int hrb_ofc_wait_download(hrb_ofc* h, unsigned long long ticket) {
    HRB_REQUIRE(h, "null handle");
    HRB_REQUIRE(ticket >= 1 && ticket <= h->downloadSeq, "unknown ticket");
    HRB_CUDA(cudaSetDevice(h->device));
    // The slot may have been re-recorded by a later download (tickets wrap every kTickets): the download stream is in
    // order, so waiting for that later record also proves that this ticket's copy has landed.  Never return unwaited.
    HRB_CUDA(cudaEventSynchronize(h->ticketEvent[ticket % hrb_ofc::kTickets]));
    return HRB_OK;
}

int hrb_ofc_synchronize(hrb_ofc* h) {
    HRB_REQUIRE(h, "null handle");
    HRB_CUDA(cudaSetDevice(h->device));
    HRB_CUDA(cudaStreamSynchronize(h->upStream));
    HRB_CUDA(cudaStreamSynchronize(h->flowStream));
    HRB_CUDA(cudaStreamSynchronize(h->stream));
    HRB_CUDA(cudaStreamSynchronize(h->downStream));
    h->flowJoinPending = false;
    return resolveFlow(h);
}

int hrb_ofc_stream(hrb_ofc* h, void** out) {
    HRB_REQUIRE(h && out, "null argument");
    *out = (void*)h->stream;
    return HRB_OK;
}

int hrb_ofc_output_device_ptr(hrb_ofc* h, void** out) {
    HRB_REQUIRE(h && out, "null argument");
    *out = h->outputRing[h->outView];
    return HRB_OK;
}

int hrb_host_register(void* ptr, size_t bytes) {
    HRB_REQUIRE(ptr && bytes, "null argument");
    HRB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return HRB_OK;
}
int hrb_host_unregister(void* ptr) {
    HRB_REQUIRE(ptr, "null argument");
    HRB_CUDA(cudaHostUnregister(ptr));
    return HRB_OK;
}
int hrb_host_alloc(void** out, size_t bytes) {
    HRB_REQUIRE(out && bytes, "null argument");
    HRB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return HRB_OK;
}
int hrb_host_free(void* ptr) {
    HRB_REQUIRE(ptr, "null argument");
    HRB_CUDA(cudaFreeHost(ptr));
    return HRB_OK;
}

// folds in the flow calculations that have FINISHED (oldest first), without waiting for the ones still running
static int resolveFinishedFlows(hrb_ofc* h) {
    for (int i = 1; i <= hrb_ofc::kFlowRecords; ++i) {
        hrb_ofc::FlowRecord& r = h->flowRec[(h->curRec + i) % hrb_ofc::kFlowRecords];
        if (!r.pending) continue;
        const cudaError_t q = cudaEventQuery(r.end);
        if (q == cudaErrorNotReady) {
            cudaGetLastError();
            break;  // later records are younger: still running too
        }
        HRB_CUDA(q);
        const int rc = resolveRecord(h, r);
        if (rc) return rc;
    }
    return HRB_OK;
}

static void fillState(const hrb_ofc* h, hrb_ofc_state* s);

int hrb_ofc_peek_state(hrb_ofc* h, hrb_ofc_state* s) {
    HRB_REQUIRE(h && s, "null argument");
    HRB_CUDA(cudaSetDevice(h->device));
    const int rc = resolveFinishedFlows(h);
    if (rc) return rc;
    fillState(h, s);
    return HRB_OK;
}

int hrb_ofc_get_state(hrb_ofc* h, hrb_ofc_state* s) {
    HRB_REQUIRE(h && s, "null argument");
    const int rc = resolveFlow(h);
    if (rc) return rc;
    fillState(h, s);
    return HRB_OK;
}

static void fillState(const hrb_ofc* h, hrb_ofc_state* s) {
    s->frame_width = h->frameWidth;
    s->frame_height = h->frameHeight;
    s->input_stride = h->inputStride;
    s->output_stride = h->outputStride;
    s->output_black_level = h->outputBlackLevel;
    s->output_white_level = h->outputWhiteLevel;
    s->res_scalar = h->resScalar;
    s->flow_width = h->flowWidth;
    s->flow_height = h->flowHeight;
    s->search_radius = h->searchRadius;
    s->ofc_calc_time = h->ofcCalcTime;
    s->ofc_avg_calc_time = h->ofcAvgCalcTime;
    s->ofc_peak_calc_time = h->ofcPeakCalcTime;
    s->ofc_calc_count = h->ofcCalcCount;
    s->ofc_calc_time_sum = h->ofcCalcTimeSum;
    s->warp_calc_time = h->warpCalcTime;
    s->delta_scalar = h->deltaScalar;
    s->neighbor_bias_scalar = h->neighborBiasScalar;
    s->total_frame_delta = h->totalFrameDelta;
    s->frame_count = h->frameCount;
}

int hrb_ofc_set_params(hrb_ofc* h, const hrb_ofc_params* p) {
    HRB_REQUIRE(h && p, "null argument");
    // the same ranges calculateOpticalFlow accepts (the filter keeps the radius in 5..16, config.h:8-9; 2..4 exist for tests)
    HRB_REQUIRE(p->search_radius >= 2 && p->search_radius <= 16, "search radius outside 2..16");
    HRB_REQUIRE(p->delta_scalar >= 0 && p->delta_scalar <= 31 && p->neighbor_bias_scalar >= 0 && p->neighbor_bias_scalar <= 31, "delta/neighbor scalar outside 0..31");
    HRB_REQUIRE(p->black_level == p->black_level && p->white_level == p->white_level, "levels must be numbers");
    h->searchRadius = p->search_radius;
    h->deltaScalar = p->delta_scalar;
    h->neighborBiasScalar = p->neighbor_bias_scalar;
    h->outputBlackLevel = p->black_level;
    h->outputWhiteLevel = p->white_level;
    return HRB_OK;
}

int hrb_ofc_set_frame_count(hrb_ofc* h, unsigned int n) {
    HRB_REQUIRE(h, "null handle");
    h->frameCount = n;
    return HRB_OK;
}

int hrb_ofc_reset(hrb_ofc* h) { return hrb_ofc_set_frame_count(h, 0); }

// ---- taps ----------------------------------------------------------------------------------------
int hrb_ofc_set_tap_mode(hrb_ofc* h, int on) {
    HRB_REQUIRE(h, "null handle");
    HRB_CUDA(cudaSetDevice(h->device));
    if (joinFlow(h)) return HRB_ERR_CUDA;
    HRB_CUDA(cudaStreamSynchronize(h->stream));
    h->tapMode = on != 0;
    if (!h->tapMode) freeTaps(h);
    return HRB_OK;
}

int hrb_ofc_num_passes(hrb_ofc* h, int* out) {
    HRB_REQUIRE(h && out, "null argument");
    *out = (int)h->taps.size();
    return HRB_OK;
}

int hrb_ofc_pass_info(hrb_ofc* h, int pass, int* window_size, int* iteration, int* step, int* windows_x, int* windows_y) {
    HRB_REQUIRE(h, "null handle");
    HRB_REQUIRE(pass >= 0 && pass < (int)h->taps.size(), "pass out of range (is tap mode on?)");
    const PassGeom& g = h->taps[pass].g;
    if (window_size) *window_size = g.windowSize;
    if (iteration) *iteration = g.iteration;
    if (step) *step = g.step;
    if (windows_x) *windows_x = g.nWx;
    if (windows_y) *windows_y = g.nWy;
    return HRB_OK;
}

int hrb_ofc_read_pass_tap(hrb_ofc* h, int pass, int which, void* dst, size_t bytes) {
    HRB_REQUIRE(h && dst, "null argument");
    if (pass < 0 || pass >= (int)h->taps.size()) {
        setLastError("[hopperrender_b200] hrb_ofc_read_pass_tap: no such pass (tap mode must be on before calculate)");
        return HRB_ERR_STATE;
    }
    HRB_CUDA(cudaSetDevice(h->device));
    const PassTapDev& t = h->taps[pass];
    const size_t nW = (size_t)t.g.nWx * t.g.nWy;
    const size_t lwlh = (size_t)h->flowWidth * h->flowHeight;
    if (which == HRB_TAP_WINDOW_SUMS) {
        HRB_REQUIRE(bytes <= nW * 16 * sizeof(uint32_t) && bytes % (nW * sizeof(uint32_t)) == 0, "size must be R*windows*4");
        HRB_CUDA(cudaMemcpyAsync(dst, t.sums, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else if (which == HRB_TAP_WINDOW_LAYER) {
        HRB_REQUIRE(bytes == nW, "size must be windows");
        HRB_CUDA(cudaMemcpyAsync(dst, t.layer, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else if (which == HRB_TAP_OFFSETS) {
        HRB_REQUIRE(bytes == 2 * lwlh * sizeof(int16_t), "size must be 2*flow_w*flow_h*2");
        const int rc = launchExpandOffsets(h, t.offX, t.nWxX, ilog2(t.wsX), t.wsY ? t.offY : nullptr, t.nWxY, t.wsY ? ilog2(t.wsY) : 0,
                                           h->offsetArrayScratch);
        if (rc) return rc;
        HRB_CUDA(cudaMemcpyAsync(dst, h->offsetArrayScratch, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else {
        HRB_REQUIRE(false, "unknown tap");
    }
    HRB_CUDA(cudaStreamSynchronize(h->stream));
    return HRB_OK;
}

int hrb_ofc_read_buffer(hrb_ofc* h, int which, void* dst, size_t bytes) {
    HRB_REQUIRE(h && dst, "null argument");
    HRB_CUDA(cudaSetDevice(h->device));
    if (joinFlow(h)) return HRB_ERR_CUDA;
    const size_t flowBytes = 2 * (size_t)h->flowWidth * h->flowHeight * sizeof(int16_t);
    if (which == HRB_BUF_OFFSET_ARRAY) {
        HRB_REQUIRE(bytes == flowBytes, "size must be 2*flow_w*flow_h*2");
        if (!h->haveFlowLevels) {
            setLastError("[hopperrender_b200] hrb_ofc_read_buffer: no flow has been calculated yet");
            return HRB_ERR_STATE;
        }
        const int s = ilog2(h->lastWs);
        const int rc = launchExpandOffsets(h, h->levelOffsets[h->lastIterParity][0], h->lastNWx, s, h->levelOffsets[h->lastIterParity][1], h->lastNWx, s,
                                           h->offsetArrayScratch);
        if (rc) return rc;
        HRB_CUDA(cudaMemcpyAsync(dst, h->offsetArrayScratch, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else if (which == HRB_BUF_FLOW_FOR_WARP || which == HRB_BUF_FLOW_LATEST) {
        HRB_REQUIRE(bytes == flowBytes, "size must be 2*flow_w*flow_h*2");
        HRB_CUDA(cudaMemcpyAsync(dst, h->blurredOffsetArray[which == HRB_BUF_FLOW_FOR_WARP ? 0 : 1], bytes, cudaMemcpyDeviceToHost, h->stream));
    } else if (which == HRB_BUF_OUTPUT_FRAME) {
        HRB_REQUIRE(bytes <= h->outFrameBytes, "size larger than the output frame");
        HRB_CUDA(cudaMemcpyAsync(dst, h->outputRing[h->outView], bytes, cudaMemcpyDeviceToHost, h->stream));
    } else if (which == HRB_BUF_RAW_FRAME_DELTA) {
        HRB_REQUIRE(bytes == sizeof(uint32_t), "size must be 4");
        HRB_CUDA(cudaMemcpyAsync(dst, h->rawDeltaDev, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else if (which == HRB_BUF_FLOW_PEAK) {
        HRB_REQUIRE(bytes == 2 * sizeof(uint32_t), "size must be 8");
        for (int i = 0; i < 2; ++i)
            HRB_CUDA(cudaMemcpyAsync(static_cast<uint32_t*>(dst) + i, h->flowMaxDev[i], sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    } else {
        HRB_REQUIRE(false, "unknown buffer");
    }
    HRB_CUDA(cudaStreamSynchronize(h->stream));
    return resolveFlow(h);
}

int hrb_ofc_write_flow(hrb_ofc* h, int which, const int16_t* src, size_t count) {
    HRB_REQUIRE(h && src, "null argument");
    HRB_REQUIRE(which == HRB_BUF_FLOW_FOR_WARP || which == HRB_BUF_FLOW_LATEST, "only the blurred flows are writable");
    HRB_REQUIRE(count == 2 * (size_t)h->flowWidth * h->flowHeight, "count must be 2*flow_w*flow_h");
    HRB_CUDA(cudaSetDevice(h->device));
    if (joinFlow(h)) return HRB_ERR_CUDA;
    const int slot = which == HRB_BUF_FLOW_FOR_WARP ? 0 : 1;
    HRB_CUDA(cudaMemcpyAsync(h->blurredOffsetArray[slot], src, count * sizeof(int16_t), cudaMemcpyHostToDevice, h->stream));
    uint32_t peak = 0;
    for (size_t i = 0; i < count; ++i) {
        const uint32_t v = (uint32_t)(src[i] < 0 ? -(int)src[i] : (int)src[i]);
        if (v > peak) peak = v;
    }
    HRB_CUDA(cudaMemcpyAsync(h->flowMaxDev[slot], &peak, sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    HRB_CUDA(cudaStreamSynchronize(h->stream));
    return HRB_OK;
}

// ---- measurement -----------------------------------------------------------------------------------
int hrb_ofc_set_profile(hrb_ofc* h, int on) {
    HRB_REQUIRE(h, "null handle");
    HRB_CUDA(cudaSetDevice(h->device));
    if (joinFlow(h)) return HRB_ERR_CUDA;
    const int rc = profResolve(h);
    if (rc) return rc;
    h->prof.on = on != 0;
    return HRB_OK;
}

int hrb_ofc_profile_read(hrb_ofc* h, hrb_ofc_profile* out) {
    HRB_REQUIRE(h && out, "null argument");
    HRB_CUDA(cudaSetDevice(h->device));
    const int rc = profResolve(h);
    if (rc) return rc;
    out->ms_ingest = h->prof.ms[CLS_INGEST];
    out->ms_search = h->prof.ms[CLS_SEARCH];
    out->ms_blur = h->prof.ms[CLS_BLUR];
    out->ms_warp = h->prof.ms[CLS_WARP];
    out->ms_copy = h->prof.ms[CLS_COPY];
    out->n_ingest = h->prof.n[CLS_INGEST];
    out->n_search = h->prof.n[CLS_SEARCH];
    out->n_blur = h->prof.n[CLS_BLUR];
    out->n_warp = h->prof.n[CLS_WARP];
    out->n_copy = h->prof.n[CLS_COPY];
    return HRB_OK;
}

int hrb_ofc_profile_reset(hrb_ofc* h) {
    HRB_REQUIRE(h, "null handle");
    const int rc = profResolve(h);
    if (rc) return rc;
    for (int i = 0; i < 5; ++i) {
        h->prof.ms[i] = 0;
        h->prof.n[i] = 0;
    }
    return HRB_OK;
}

int hrb_ofc_set_side_data(hrb_ofc* h, const hrb_side_data* items, int count) {
    HRB_REQUIRE(h && (items || count == 0), "null argument");
    HRB_REQUIRE(count >= 0 && count <= 64, "side-data count outside 0..64");
    for (int i = 0; i < count; ++i) HRB_REQUIRE(items[i].data || items[i].bytes == 0, "null blob with a size");
    h->sideData.clear();
    h->sideData.resize((size_t)count);
    for (int i = 0; i < count; ++i) {
        memcpy(h->sideData[i].guid, items[i].guid, 16);
        const uint8_t* p = static_cast<const uint8_t*>(items[i].data);
        h->sideData[i].bytes.assign(p, p + items[i].bytes);
    }
    return HRB_OK;
}

int hrb_ofc_get_side_data(hrb_ofc* h, hrb_side_data* items, int capacity, int* count) {
    HRB_REQUIRE(h && count, "null argument");
    HRB_REQUIRE(capacity >= 0 && (items || capacity == 0), "null array with a capacity");
    *count = (int)h->sideData.size();
    for (int i = 0; i < capacity && i < *count; ++i) {
        memcpy(items[i].guid, h->sideData[i].guid, 16);
        items[i].data = h->sideData[i].bytes.data();
        items[i].bytes = h->sideData[i].bytes.size();
    }
    return HRB_OK;
}

int hrb_ofc_set_output_stripe(hrb_ofc* h, int row_begin, int row_end) {
    HRB_REQUIRE(h, "null handle");
    HRB_REQUIRE(row_begin >= 0 && row_end <= h->frameHeight && row_begin < row_end, "stripe outside the frame");
    HRB_REQUIRE((row_begin % 2) == 0 && (row_end % 2) == 0, "stripe bounds must be even (chroma rows are shared by luma row pairs)");
    h->stripeY0 = row_begin;
    h->stripeY1 = row_end;
    return HRB_OK;
}

int hrb_ofc_set_flow_overlap(hrb_ofc* h, int on) {
    HRB_REQUIRE(h, "null handle");
    const int rc = joinFlow(h);
    if (rc) return rc;
    h->flowOverlap = on != 0;
    return HRB_OK;
}

int hrb_ofc_join_flow(hrb_ofc* h) {
    HRB_REQUIRE(h, "null handle");
    return joinFlow(h);
}

int hrb_ofc_set_search_variant(hrb_ofc* h, int variant) {
    HRB_REQUIRE(h, "null handle");
    HRB_REQUIRE(variant >= 0 && variant <= 4, "variant must be 0 (automatic), 1 (generic kernel only), 2 (tile kernel without TMA), 3 (tile kernel, per-pixel path forced) or 4 (butterfly form for windows of 2 and 4)");
    h->searchVariant = variant;
    return HRB_OK;
}

// Debug aid: per-CTA timelines of the tile search kernels (8 words per CTA: start, boxes landed, runs done, end
// [globaltimer ns], SM id, block x, block y, rounds), `words_per_pass` words for each of up to 32 passes.
int hrb_ofc_debug_timeline(hrb_ofc* h, size_t words_per_pass) {
    HRB_REQUIRE(h, "null handle");
    HRB_CUDA(cudaSetDevice(h->device));
    HRB_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(h->dbgDev);
    h->dbgDev = nullptr;
    h->dbgStride = words_per_pass;
    if (words_per_pass) {
        HRB_CUDA(cudaMalloc(&h->dbgDev, words_per_pass * 32 * sizeof(unsigned long long)));
        HRB_CUDA(cudaMemset(h->dbgDev, 0, words_per_pass * 32 * sizeof(unsigned long long)));
    }
    return HRB_OK;
}
int hrb_ofc_debug_timeline_read(hrb_ofc* h, void* dst, size_t bytes) {
    HRB_REQUIRE(h && dst && h->dbgDev, "no timeline buffer");
    HRB_REQUIRE(bytes <= h->dbgStride * 32 * sizeof(unsigned long long), "size larger than the buffer");
    HRB_CUDA(cudaSetDevice(h->device));
    HRB_CUDA(cudaDeviceSynchronize());
    HRB_CUDA(cudaMemcpy(dst, h->dbgDev, bytes, cudaMemcpyDeviceToHost));
    return HRB_OK;
}

uint64_t hrb_kernel_launch_count(void) { return g_launchCount.load(std::memory_order_relaxed); }

int hrb_microbench_sad_peak(int device_ordinal, double* giga_absdiff_per_s) {
    HRB_REQUIRE(giga_absdiff_per_s, "null argument");
    return microbenchSad(device_ordinal, giga_absdiff_per_s);
}

const char* hrb_last_error(void) { return t_lastError.c_str(); }
const char* hrb_version(void) { return "hopperrender_b200 0.1 (sm_100a)"; }

}  // extern "C"
