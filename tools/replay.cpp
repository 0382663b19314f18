// replay.cpp — headless C++ driver: the call sequence of CHopperRender::DeliverToRenderer
// (HopperRender/HopperRender.cpp:944-1211) against the header-compatible classes of include/, i.e. the same
// host code a DirectShow build would run, minus COM.  Reads raw NV12/P010 frames from a file, writes one line
// per delivered frame: "<source> <index> <blend> <warped> <radius> <crc32>".
//
//   replay <frames.raw> <width> <height> <hdr 0|1> <n_frames> <target_frame_time> <max_calc_res> <radius> [auto]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <vector>

#include "opticalFlowCalcHDR.h"
#include "opticalFlowCalcSDR.h"

static uint32_t crc32(const unsigned char* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

struct DeltaEntry {
    unsigned int frameNumber, totalDelta;
};

int main(int argc, char** argv) {
    if (argc < 9) {
        fprintf(stderr, "usage: %s frames.raw width height hdr n_frames target_frame_time max_calc_res radius [auto]\n", argv[0]);
        return 2;
    }
    const char* path = argv[1];
    const int W = atoi(argv[2]), H = atoi(argv[3]), hdr = atoi(argv[4]), nFrames = atoi(argv[5]);
    const long long targetFrameTime = atoll(argv[6]);
    const int maxCalcRes = atoi(argv[7]), radius = atoi(argv[8]);
    const bool autoAdjust = argc > 9;
    const long long sourceFrameTime = 417083;  // 23.976 fps in 100 ns units
    const size_t frameBytes = (size_t)W * H * 3 / 2 * (hdr ? 2 : 1);
    const unsigned sceneChangeThreshold = 200;  // DEFAULT_SCENE_CHANGE_THRESHOLD, config.h:28

    FILE* f = fopen(path, "rb");
    if (!f) {
        perror(path);
        return 2;
    }
    std::vector<unsigned char> in(frameBytes), out(frameBytes);
    try {
        OpticalFlowCalc* calc = hdr ? (OpticalFlowCalc*)new OpticalFlowCalcHDR(H, W, W, W, 8, 6, 0.0f, 255.0f, maxCalcRes)
                                    : (OpticalFlowCalc*)new OpticalFlowCalcSDR(H, W, W, W, 8, 6, 0.0f, 255.0f, maxCalcRes);
        calc->m_opticalFlowSearchRadius = radius;
        double blend = 0.0, totalWarpDuration = 0.0;
        std::deque<DeltaEntry> history;
        for (int n = 0; n < nFrames; ++n) {
            if (fread(in.data(), 1, frameBytes, f) != frameBytes) break;
            // number of output frames for this source frame (HopperRender.cpp:945)
            const int numInt = (int)std::fmax(std::ceil((1.0 - blend) / ((double)targetFrameTime / (double)sourceFrameTime)), 1.0);
            if (autoAdjust) {  // HopperRender.cpp:1438-1463
                const double budget = (double)sourceFrameTime / 10000000.0;
                const double used = calc->m_ofcCalcTime + totalWarpDuration;
                if (used * 1.4 > budget) {
                    if (calc->m_opticalFlowSearchRadius > 5) calc->m_opticalFlowSearchRadius--;
                } else if (used * 1.6 < budget) {
                    if (calc->m_opticalFlowSearchRadius < 16) calc->m_opticalFlowSearchRadius++;
                }
                totalWarpDuration = 0.0;
            }
            calc->updateFrame(in.data());
            if (calc->m_frameCount >= 3) {
                calc->calculateOpticalFlow();
                const unsigned framesIn3s = (unsigned)(3.0 * 10000000.0 / sourceFrameTime);
                history.push_back({calc->m_frameCount, calc->m_totalFrameDelta});
                while (!history.empty() && calc->m_frameCount - history.front().frameNumber > framesIn3s) history.pop_front();
            }
            for (int i = 0; i < numInt; ++i) {
                bool sceneChange = false;  // HopperRender.cpp:1126-1176
                if (history.size() >= 3) {
                    const size_t hs = history.size();
                    const size_t count = hs - 2 < 10 ? hs - 2 : 10;
                    unsigned long long sum = 0;
                    for (size_t k = 0; k < count; ++k) sum += history[hs - 2 - k].totalDelta;
                    const int average = (int)(sum / count);
                    const int next = (int)history[hs - 1].totalDelta, cur = (int)history[hs - 2].totalDelta;
                    const int d1 = cur - average, d2 = cur - next;
                    sceneChange = d1 > 0 && d2 > 0 && (unsigned)d1 >= sceneChangeThreshold && (unsigned)d2 >= sceneChangeThreshold;
                }
                const bool warped = calc->m_frameCount >= 3 && !sceneChange;
                if (warped)
                    calc->warpFrames((float)blend, 2);
                else
                    calc->copyFrame();
                calc->downloadFrame(out.data());
                totalWarpDuration += calc->m_warpCalcTime;
                printf("%u %d %.9f %d %d %08x\n", calc->m_frameCount, i, blend, warped ? 1 : 0, calc->m_opticalFlowSearchRadius, crc32(out.data(), frameBytes));
                blend += (double)targetFrameTime / (double)sourceFrameTime;  // HopperRender.cpp:1192-1197
                if (blend >= 1.0) blend -= 1.0;
            }
        }
        fprintf(stderr, "ofcCalcTime %.6f s, warpCalcTime %.6f s, flow %dx%d\n", calc->m_ofcCalcTime, calc->m_warpCalcTime, calc->m_opticalFlowFrameWidth,
                calc->m_opticalFlowFrameHeight);
        // the reference throws on an invalid blending scalar (opticalFlowCalcSDR.cpp:143-146); so must the drop-in
        bool threw = false;
        try {
            calc->warpFrames(1.5f, 2);
        } catch (const std::runtime_error&) {
            threw = true;
        }
        if (!threw) {
            fprintf(stderr, "warpFrames(1.5) did not throw\n");
            return 3;
        }
        delete calc;
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        fclose(f);
        return 1;
    }
    fclose(f);
    return 0;
}
