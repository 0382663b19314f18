// opticalFlowCalc.h — header-compatible re-creation of HopperRender's calculator base class on top of the
// hopperrender_b200 C ABI (hrb.h).  A filter built against HopperRender/opticalFlowCalc.h:24-138 compiles
// against this header unchanged: same class name, same five virtuals, same public data members with the
// same meaning.  What is gone are the OpenCL members (cl_mem / cl_kernel / grids), which no caller touches
// (HopperRender/HopperRender.cpp uses only the members kept here — SURVEY.md §8b).
//
// The members are plain fields, exactly as in the reference, because the filter reads and writes them
// directly (m_opticalFlowSearchRadius++, m_frameCount = 0, ...).  Every method pushes the writable ones to
// the library before the call and pulls all of them back after it.  A non-zero hrb return code becomes the
// std::runtime_error the reference throws from CHECK_ERROR (opticalFlowCalc.h:15-22).
#pragma once

#include <stdexcept>
#include <string>

#include "hrb.h"

class OpticalFlowCalc {
public:
    // Video properties (opticalFlowCalc.h:27-32)
    int m_frameWidth = 0;
    int m_frameHeight = 0;
    int m_inputStride = 0;
    int m_outputStride = 0;
    float m_outputBlackLevel = 0.0f;
    float m_outputWhiteLevel = 255.0f;

    // Optical flow calculation (opticalFlowCalc.h:35-50)
    int m_opticalFlowResScalar = 0;
    int m_opticalFlowFrameWidth = 0;
    int m_opticalFlowFrameHeight = 0;
    int m_opticalFlowSearchRadius = 0;
    double m_ofcCalcTime = 0.0;
    double m_ofcAvgCalcTime = 0.0;
    double m_ofcPeakCalcTime = 0.0;
    int m_ofcCalcCount = 0;
    double m_ofcCalcTimeSum = 0.0;
    double m_warpCalcTime = 0.0;
    int m_deltaScalar = 0;
    int m_neighborBiasScalar = 0;
    unsigned int m_totalFrameDelta = 0;
    unsigned int m_frameCount = 0;

    OpticalFlowCalc() = default;
    OpticalFlowCalc(const OpticalFlowCalc&) = delete;
    OpticalFlowCalc& operator=(const OpticalFlowCalc&) = delete;

    virtual ~OpticalFlowCalc() {
        if (m_handle) hrb_ofc_destroy(m_handle);
    }

    // opticalFlowCalc.h:100-132
    virtual void updateFrame(unsigned char* inputPlanes) {
        push();
        check(hrb_ofc_update_frame(m_handle, inputPlanes));
        pull();
    }
    virtual void downloadFrame(unsigned char* outputPlanes) {
        push();
        check(hrb_ofc_download_frame(m_handle, outputPlanes));
        pull();
    }
    virtual void calculateOpticalFlow() {
        push();
        check(hrb_ofc_calculate_optical_flow(m_handle));
        pull();
    }
    virtual void warpFrames(const float blendingScalar, const int frameOutputMode) {
        push();
        check(hrb_ofc_warp_frames(m_handle, blendingScalar, frameOutputMode));
        pull();
    }
    virtual void copyFrame() {
        push();
        check(hrb_ofc_copy_frame(m_handle));
        pull();
    }

    hrb_ofc* handle() const { return m_handle; }

protected:
    void create(int frameHeight, int frameWidth, int inputStride, int outputStride, int deltaScalar, int neighborScalar, float blackLevel,
                float whiteLevel, int maxCalcRes, bool hdr) {
        hrb_ofc_desc d{};
        d.frame_height = frameHeight;
        d.frame_width = frameWidth;
        d.input_stride = inputStride;
        d.output_stride = outputStride;
        d.delta_scalar = deltaScalar;
        d.neighbor_scalar = neighborScalar;
        d.black_level = blackLevel;
        d.white_level = whiteLevel;
        d.max_calc_res = maxCalcRes;
        d.is_hdr = hdr ? 1 : 0;
        d.device_ordinal = 0;
        d.cuda_stream = nullptr;
        check(hrb_ofc_create(&m_handle, &d));
        pull();
    }

    static void check(int rc) {
        if (rc != HRB_OK) throw std::runtime_error(std::string(hrb_last_error()) + "\n");
    }

    // members the filter writes on a live object (HopperRender.cpp:840, 1386-1389, 1448, 1457)
    void push() {
        hrb_ofc_params p{};
        p.search_radius = m_opticalFlowSearchRadius;
        p.delta_scalar = m_deltaScalar;
        p.neighbor_bias_scalar = m_neighborBiasScalar;
        p.black_level = m_outputBlackLevel;
        p.white_level = m_outputWhiteLevel;
        check(hrb_ofc_set_params(m_handle, &p));
        check(hrb_ofc_set_frame_count(m_handle, m_frameCount));
    }

    void pull() {
        hrb_ofc_state s{};
        check(hrb_ofc_get_state(m_handle, &s));
        m_frameWidth = s.frame_width;
        m_frameHeight = s.frame_height;
        m_inputStride = s.input_stride;
        m_outputStride = s.output_stride;
        m_outputBlackLevel = s.output_black_level;
        m_outputWhiteLevel = s.output_white_level;
        m_opticalFlowResScalar = s.res_scalar;
        m_opticalFlowFrameWidth = s.flow_width;
        m_opticalFlowFrameHeight = s.flow_height;
        m_opticalFlowSearchRadius = s.search_radius;
        m_ofcCalcTime = s.ofc_calc_time;
        m_ofcAvgCalcTime = s.ofc_avg_calc_time;
        m_ofcPeakCalcTime = s.ofc_peak_calc_time;
        m_ofcCalcCount = s.ofc_calc_count;
        m_ofcCalcTimeSum = s.ofc_calc_time_sum;
        m_warpCalcTime = s.warp_calc_time;
        m_deltaScalar = s.delta_scalar;
        m_neighborBiasScalar = s.neighbor_bias_scalar;
        m_totalFrameDelta = s.total_frame_delta;
        m_frameCount = s.frame_count;
    }

    hrb_ofc* m_handle = nullptr;
};
