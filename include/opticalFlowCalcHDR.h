// opticalFlowCalcHDR.h — P010 (10-bit in 16) calculator; drop-in for HopperRender/opticalFlowCalcHDR.h:13-56.
#pragma once

#include "opticalFlowCalc.h"

class OpticalFlowCalcHDR : public OpticalFlowCalc {
public:
    // same argument order as HopperRender/opticalFlowCalcHDR.h:15-17
    OpticalFlowCalcHDR(const int frameHeight, const int frameWidth, const int inputStride, const int outputStride, int deltaScalar,
                       int neighborScalar, float blackLevel, float whiteLevel, int maxCalcRes) {
        create(frameHeight, frameWidth, inputStride, outputStride, deltaScalar, neighborScalar, blackLevel, whiteLevel, maxCalcRes, true);
    }
};
