// ref_harness.cpp — C entry points over the UNMODIFIED reference classes OpticalFlowCalcSDR / OpticalFlowCalcHDR
// (compiled from /root/reference/HopperRender/opticalFlowCalc*.cpp with their kernel strings), so that tests can
// drive the reference itself on an OpenCL device and read its buffers.  Test infrastructure (oracle/_ref).
#include <string.h>

#include <string>
#include <vector>

#include "opticalFlowCalcHDR.h"
#include "opticalFlowCalcSDR.h"

extern "C" void OutputDebugStringA(const char*) {}

extern "C" {
extern void (*hrref_on_set_kernel_arg)(cl_kernel, cl_uint, size_t, const void*);
extern void (*hrref_on_enqueue_kernel)(cl_command_queue, cl_kernel);
}

namespace {

struct PassTap {
    int windowSize, iteration, step;
    std::vector<uint32_t> sums;
    std::vector<uint8_t> layers;
    std::vector<int16_t> offsets;
};

struct Ref {
    OpticalFlowCalc* c = nullptr;
    bool hdr = false;
    bool tapsOn = false;
    std::vector<PassTap> taps;
    int curWs = 0, curIter = 0, curStep = 0;
};

thread_local std::string t_err;
Ref* g_active = nullptr;  // object inside calculateOpticalFlow (taps)

void onSetArg(cl_kernel k, cl_uint i, size_t n, const void* p) {
    Ref* r = g_active;
    if (!r || k != r->c->m_calcDeltaSumsKernel || n != sizeof(int)) return;
    int v;
    memcpy(&v, p, sizeof(int));
    if (i == 9) r->curWs = v;          // windowSize   (opticalFlowCalcSDR.cpp:81)
    else if (i == 12) r->curIter = v;  // iteration    (:83)
    else if (i == 13) r->curStep = v;  // step         (:84)
}

void onEnqueue(cl_command_queue q, cl_kernel k) {
    Ref* r = g_active;
    if (!r || !r->tapsOn || k != r->c->m_adjustOffsetArrayKernel) return;
    const size_t lw = r->c->m_opticalFlowFrameWidth, lh = r->c->m_opticalFlowFrameHeight, R = r->c->m_opticalFlowSearchRadius;
    PassTap t;
    t.windowSize = r->curWs;
    t.iteration = r->curIter;
    t.step = r->curStep;
    t.sums.resize(R * lw * lh);
    t.layers.resize(lw * lh);
    t.offsets.resize(2 * lw * lh);
    clFinish(q);
    clEnqueueReadBuffer(q, r->c->m_summedDeltaValuesArray, CL_TRUE, 0, t.sums.size() * 4, t.sums.data(), 0, NULL, NULL);
    clEnqueueReadBuffer(q, r->c->m_lowestLayerArray, CL_TRUE, 0, t.layers.size(), t.layers.data(), 0, NULL, NULL);
    clEnqueueReadBuffer(q, r->c->m_offsetArray, CL_TRUE, 0, t.offsets.size() * 2, t.offsets.data(), 0, NULL, NULL);
    r->taps.push_back(std::move(t));
}

template <typename F> int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        t_err = e.what();
        return 1;
    } catch (...) {
        t_err = "unknown exception";
        return 1;
    }
}

}  // namespace

extern "C" {

const char* hrref_last_error() { return t_err.c_str(); }

void* hrref_ofc_create(int frameHeight, int frameWidth, int inputStride, int outputStride, int deltaScalar, int neighborScalar, float blackLevel,
                       float whiteLevel, int maxCalcRes, int hdr) {
    hrref_on_set_kernel_arg = onSetArg;
    hrref_on_enqueue_kernel = onEnqueue;
    Ref* r = new Ref();
    r->hdr = hdr != 0;
    const int rc = guarded([&] {
        if (hdr)
            r->c = new OpticalFlowCalcHDR(frameHeight, frameWidth, inputStride, outputStride, deltaScalar, neighborScalar, blackLevel, whiteLevel, maxCalcRes);
        else
            r->c = new OpticalFlowCalcSDR(frameHeight, frameWidth, inputStride, outputStride, deltaScalar, neighborScalar, blackLevel, whiteLevel, maxCalcRes);
    });
    if (rc) {
        delete r;
        return nullptr;
    }
    // the reference leaves its buffers uninitialised; tests read flows before the first calculate, so zero them once
    const size_t n = 2 * (size_t)r->c->m_opticalFlowFrameWidth * r->c->m_opticalFlowFrameHeight * sizeof(short);
    const cl_uint zero = 0;
    clEnqueueFillBuffer(r->c->m_queue, r->c->m_blurredOffsetArray[0], &zero, sizeof(short), 0, n, 0, NULL, NULL);
    clEnqueueFillBuffer(r->c->m_queue, r->c->m_blurredOffsetArray[1], &zero, sizeof(short), 0, n, 0, NULL, NULL);
    clFinish(r->c->m_queue);
    return r;
}

void hrref_ofc_destroy(void* h) {
    Ref* r = (Ref*)h;
    if (!r) return;
    guarded([&] { delete r->c; });
    delete r;
}

int hrref_ofc_update_frame(void* h, uint8_t* p) { return guarded([&] { ((Ref*)h)->c->updateFrame(p); }); }
int hrref_ofc_download_frame(void* h, uint8_t* p) { return guarded([&] { ((Ref*)h)->c->downloadFrame(p); }); }
int hrref_ofc_calculate_optical_flow(void* h) {
    Ref* r = (Ref*)h;
    r->taps.clear();
    g_active = r;
    const int rc = guarded([&] { r->c->calculateOpticalFlow(); });
    g_active = nullptr;
    return rc;
}
int hrref_ofc_warp_frames(void* h, float t, int mode) { return guarded([&] { ((Ref*)h)->c->warpFrames(t, mode); }); }
int hrref_ofc_copy_frame(void* h) { return guarded([&] { ((Ref*)h)->c->copyFrame(); }); }

struct hrref_state {
    int frameWidth, frameHeight, inputStride, outputStride;
    float outputBlackLevel, outputWhiteLevel;
    int resScalar, flowWidth, flowHeight, searchRadius;
    double ofcCalcTime, ofcAvgCalcTime, ofcPeakCalcTime, warpCalcTime;
    int deltaScalar, neighborBiasScalar;
    unsigned int totalFrameDelta, frameCount;
};

void hrref_ofc_get_state(void* h, hrref_state* s) {
    OpticalFlowCalc* c = ((Ref*)h)->c;
    s->frameWidth = c->m_frameWidth;
    s->frameHeight = c->m_frameHeight;
    s->inputStride = c->m_inputStride;
    s->outputStride = c->m_outputStride;
    s->outputBlackLevel = c->m_outputBlackLevel;
    s->outputWhiteLevel = c->m_outputWhiteLevel;
    s->resScalar = c->m_opticalFlowResScalar;
    s->flowWidth = c->m_opticalFlowFrameWidth;
    s->flowHeight = c->m_opticalFlowFrameHeight;
    s->searchRadius = c->m_opticalFlowSearchRadius;
    s->ofcCalcTime = c->m_ofcCalcTime;
    s->ofcAvgCalcTime = c->m_ofcAvgCalcTime;
    s->ofcPeakCalcTime = c->m_ofcPeakCalcTime;
    s->warpCalcTime = c->m_warpCalcTime;
    s->deltaScalar = c->m_deltaScalar;
    s->neighborBiasScalar = c->m_neighborBiasScalar;
    s->totalFrameDelta = c->m_totalFrameDelta;
    s->frameCount = c->m_frameCount;
}

void hrref_ofc_set_params(void* h, int searchRadius, int deltaScalar, int neighborBiasScalar, float black, float white) {
    OpticalFlowCalc* c = ((Ref*)h)->c;
    c->m_opticalFlowSearchRadius = searchRadius;
    c->m_deltaScalar = deltaScalar;
    c->m_neighborBiasScalar = neighborBiasScalar;
    c->m_outputBlackLevel = black;
    c->m_outputWhiteLevel = white;
}
void hrref_ofc_set_frame_count(void* h, unsigned int n) { ((Ref*)h)->c->m_frameCount = n; }

void hrref_ofc_enable_taps(void* h, int on) { ((Ref*)h)->tapsOn = on != 0; }
int hrref_ofc_num_passes(void* h) { return (int)((Ref*)h)->taps.size(); }
int hrref_ofc_pass_info(void* h, int pass, int* windowSize, int* iteration, int* step) {
    Ref* r = (Ref*)h;
    if (pass < 0 || pass >= (int)r->taps.size()) return 1;
    *windowSize = r->taps[pass].windowSize;
    *iteration = r->taps[pass].iteration;
    *step = r->taps[pass].step;
    return 0;
}
int hrref_ofc_read_pass_tap(void* h, int pass, int which, void* dst, size_t bytes) {
    Ref* r = (Ref*)h;
    if (pass < 0 || pass >= (int)r->taps.size()) return 1;
    const PassTap& t = r->taps[pass];
    const void* src;
    size_t n;
    if (which == 0) { src = t.sums.data(); n = t.sums.size() * 4; }
    else if (which == 1) { src = t.layers.data(); n = t.layers.size(); }
    else if (which == 2) { src = t.offsets.data(); n = t.offsets.size() * 2; }
    else return 2;
    if (bytes != n) return 3;
    memcpy(dst, src, n);
    return 0;
}
// which: 0 offsetArray, 1 m_blurredOffsetArray[0], 2 m_blurredOffsetArray[1], 3 output frame
int hrref_ofc_read_buffer(void* h, int which, void* dst, size_t bytes) {
    OpticalFlowCalc* c = ((Ref*)h)->c;
    cl_mem m = which == 0 ? c->m_offsetArray : which == 1 ? c->m_blurredOffsetArray[0] : which == 2 ? c->m_blurredOffsetArray[1] : which == 3 ? c->m_outputFrameArray : nullptr;
    if (!m) return 2;
    clFinish(c->m_queue);
    return clEnqueueReadBuffer(c->m_queue, m, CL_TRUE, 0, bytes, dst, 0, NULL, NULL) == CL_SUCCESS ? 0 : 3;
}
int hrref_ofc_write_flow(void* h, int which, const int16_t* src, size_t count) {
    OpticalFlowCalc* c = ((Ref*)h)->c;
    cl_mem m = which == 1 ? c->m_blurredOffsetArray[0] : c->m_blurredOffsetArray[1];
    const cl_int rc = clEnqueueWriteBuffer(c->m_queue, m, CL_TRUE, 0, count * 2, src, 0, NULL, NULL);
    clFinish(c->m_queue);
    return rc == CL_SUCCESS ? 0 : 3;
}

int hrref_device_name(void* h, char* dst, size_t n) {
    OpticalFlowCalc* c = ((Ref*)h)->c;
    return clGetDeviceInfo(c->m_clDeviceId, CL_DEVICE_NAME, n, dst, NULL) == CL_SUCCESS ? 0 : 1;
}

}  // extern "C"
