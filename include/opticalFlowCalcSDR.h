// opticalFlowCalcSDR.h — NV12 (8-bit) calculator; drop-in for HopperRender/opticalFlowCalcSDR.h:13-56.
#pragma once

#include "opticalFlowCalc.h"

class OpticalFlowCalcSDR : public OpticalFlowCalc {
public:
    // same argument order as HopperRender/opticalFlowCalcSDR.h:15-17
    OpticalFlowCalcSDR(const int frameHeight, const int frameWidth, const int inputStride, const int outputStride, int deltaScalar,
                       int neighborScalar, float blackLevel, float whiteLevel, int maxCalcRes) {
        create(frameHeight, frameWidth, inputStride, outputStride, deltaScalar, neighborScalar, blackLevel, whiteLevel, maxCalcRes, false);
    }
};
