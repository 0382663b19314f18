// hr_oracle.cpp — CPU restatement of HopperRender's optical-flow hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (hopperrender_b200/, include/) may link,
// import or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs use it, and only as the checker / the CPU timing baseline.
//
// What it is: a literal, per-work-item restatement of the six OpenCL kernels of the reference
// plus the host schedule of OpticalFlowCalcSDR/HDR.  Every function cites the reference lines it
// follows (paths relative to /root/reference, HR/ = HopperRender/).
//
// Semantics fixed here where the reference is racy or undefined:
//   * calcDeltaSums' barrier-free local reduction is given its lock-step meaning: the window
//     representative receives the exact sum over the window's in-range work-items, mod 2^32.
//   * the single-reflection mirror of calcDeltaSums is followed by a clamp (only differs where
//     the reference would read out of bounds).
//   * fp32 is IEEE with correctly rounded division, float->integer conversions truncate, round() is
//     half-away-from-zero.  No implicit FMA contraction (compile with -ffp-contract=off); the ONE
//     contraction the reference's OpenCL build performs on NVIDIA GPUs — the blend a*t21 + b*t12 ->
//     fma(a, t21, b*t12) — is written explicitly.  What is left is the OpenCL compiler's approximate
//     division in apply_levelsY (not reproducible off the GPU): <= 1 LSB, luma only.
//   * atan2 (HSV visualisation only) is evaluated in double and rounded to float.
//
// Parity pin: the reference ships no tests or golden vectors.  This restatement is
// pinned against the reference's own kernel strings executed by oracle/_ref (a minimal CPU OpenCL
// shim with lock-step work-group emulation that compiles the unmodified reference sources); see
// oracle/README.md and tests/test_oracle_vs_ref.py, and the committed vectors in tests/golden/.
//
// Build: g++ -O3 -march=native -fopenmp -ffp-contract=off -shared -fPIC (oracle/Makefile).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------

// sq(d) = d*d*sign(d)  — HR/calcDeltaSumsKernelSDR.h:70-71,73-74 ; HR/adjustOffsetArrayKernelSDR.h:18
inline int signedSquare(int d) { return d * d * (d > 0 ? 1 : -1); }

// HR/calcDeltaSumsKernelSDR.h:86-95 (single reflection) + the clamp documented above
inline int mirrorSearch(int n, int dim) {
    if (n >= dim) {
        n = dim - (n - dim + 1);
    } else if (n < 0) {
        n = -n - 1;
    }
    return std::min(std::max(n, 0), dim - 1);
}

// HR/blurFlowKernelSDR.h:7-14
inline int mirrorBlur(int pos, int dim) {
    if (pos >= dim) return dim - (pos - dim + 1);
    if (pos < 0) return -pos - 1;
    return pos;
}

// HR/warpFrameKernelSDR.h:12-20
inline int mirrorWarp(int pos, int dim) {
    int res = pos;
    if (pos >= dim - 1) {
        res = pos - ((pos - (dim - 2)) * 2);
    } else if (pos < 1) {
        res = -pos + 1;
    }
    return std::min(std::max(res, 1), dim - 2);
}

template <typename T> struct PixelTraits;
template <> struct PixelTraits<uint8_t> {
    static constexpr bool hdr = false;
    static constexpr float maxv = 255.0f;    // HR/warpFrameKernelSDR.h:4
    static constexpr float mid = 128.0f;     // HR/warpFrameKernelSDR.h:8
    static constexpr int midInt = 128;       // HR/warpFrameKernelSDR.h:147,162
    static constexpr int greyShift = 2;      // HR/warpFrameKernelSDR.h:162
    static constexpr unsigned greyMax = 255u;
    static inline unsigned searchSample(uint8_t v) { return v; }  // HR/calcDeltaSumsKernelSDR.h:98-100
};
template <> struct PixelTraits<uint16_t> {
    static constexpr bool hdr = true;
    static constexpr float maxv = 65535.0f;  // HR/warpFrameKernelHDR.h:4
    static constexpr float mid = 32768.0f;   // HR/warpFrameKernelHDR.h:8
    static constexpr int midInt = 32768;     // HR/warpFrameKernelHDR.h:147,162
    static constexpr int greyShift = 10;     // HR/warpFrameKernelHDR.h:162
    static constexpr unsigned greyMax = 65535u;
    static inline unsigned searchSample(uint16_t v) { return v >> 8; }  // HR/calcDeltaSumsKernelHDR.h:98-100
};

// HR/warpFrameKernelSDR.h:3-5 / HDR :3-5 ; HR/copyFrameKernelSDR.h:3-5.  Return type unsigned short.
template <typename T> inline uint16_t applyLevelsY(float value, float black, float white) {
    float r = std::fmax(std::fmin((value - black) / (white - black) * PixelTraits<T>::maxv, PixelTraits<T>::maxv), 0.0f);
    return (uint16_t)r;
}
// HR/warpFrameKernelSDR.h:7-9 / HDR :7-9
template <typename T> inline uint16_t applyLevelsUV(float value, float white) {
    float r = std::fmax(std::fmin((value - PixelTraits<T>::mid) / white * PixelTraits<T>::maxv + PixelTraits<T>::mid, PixelTraits<T>::maxv), 0.0f);
    return (uint16_t)r;
}

// HR/warpFrameKernelSDR.h:23-113 / HR/warpFrameKernelHDR.h:23-113.
// Returns unsigned char (SDR) / unsigned short (HDR); currPixel has the same type.
template <typename T> inline T visualizeFlow(short offsetX, short offsetY, T currPixel, int channel, int resImpact) {
    uint8_t r, g, b;
    const int ax = std::abs((int)offsetX), ay = std::abs((int)offsetY);
    if ((float)ax < 1.0f && (float)ay < 1.0f) {  // :32
        r = g = b = 0;
    } else {
        const float angle_rad = (float)std::atan2((double)offsetY, (double)offsetX);  // :38
        float angle_deg = angle_rad * (180.0f / 3.14159274101257f);                   // :41
        if (angle_deg < 0) angle_deg += 360.0f;                                        // :44-46
        angle_deg = std::fmod(angle_deg, 360.0f);                                      // :49
        if (angle_deg < 0) angle_deg += 360.0f;                                        // :50-52
        const float hue = angle_deg / 360.0f;                                          // :55
        const int h_i = (int)(hue * 6.0f);                                             // :58
        const float f = hue * 6.0f - h_i;                                              // :59
        const float q = 1.0f - f;                                                      // :60
        switch (h_i % 6) {                                                             // :62-98
            case 0: r = 255; g = (uint8_t)(f * 255.0f); b = 0; break;
            case 1: r = (uint8_t)(q * 255.0f); g = 255; b = 0; break;
            case 2: r = 0; g = 255; b = (uint8_t)(f * 255.0f); break;
            case 3: r = 0; g = (uint8_t)(q * 255.0f); b = 255; break;
            case 4: r = (uint8_t)(f * 255.0f); g = 0; b = 255; break;
            case 5: r = 255; g = 0; b = (uint8_t)(q * 255.0f); break;
            default: r = g = b = 0; break;
        }
        // :101-103   (float)c / 255.0f * (int) * (float)resImpact, left to right
        r = (uint8_t)std::fmax(std::fmin((float)r / 255.0f * (float)(ax + ay) * (float)resImpact, 255.0f), 0.0f);
        g = (uint8_t)std::fmax(std::fmin((float)g / 255.0f * (float)ay * 2.0f * (float)resImpact, 255.0f), 0.0f);
        b = (uint8_t)std::fmax(std::fmin((float)b / 255.0f * (float)(ax + ay) * (float)resImpact, 255.0f), 0.0f);
    }
    if (channel == 0) {  // :107
        const float y = std::fmax(std::fmin((float)r * 0.299f + (float)g * 0.587f + (float)b * 0.114f, 255.0f), 0.0f);
        if (PixelTraits<T>::hdr) return (T)(((int)(uint16_t)y << 7) + ((int)currPixel >> 1));
        return (T)(((int)(uint8_t)y >> 1) + ((int)currPixel >> 1));
    } else if (channel == 1) {  // :109
        const float u = std::fmax(std::fmin((float)r * -0.168736f + (float)g * -0.331264f + (float)b * 0.5f + 128.0f, 255.0f), 0.0f);
        if (PixelTraits<T>::hdr) return (T)((int)(uint16_t)u << 8);
        return (T)u;
    } else {  // :111
        const float v = std::fmax(std::fmin((float)r * 0.5f + (float)g * -0.418688f + (float)b * -0.081312f + 128.0f, 255.0f), 0.0f);
        if (PixelTraits<T>::hdr) return (T)((int)(uint16_t)v << 8);
        return (T)v;
    }
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// Value one work-item (cx,cy,cz) contributes.  HR/calcDeltaSumsKernelSDR.h:61-143 (HDR: same lines).
template <typename T>
inline uint32_t deltaOfWorkItem(const T* frame1, const T* frame2, const int16_t* offsetArray, int cx, int cy, int cz,
                                int dimY, int dimX, int inputStride, int lowDimY, int lowDimX, int windowSize,
                                int searchWindowSize, int resolutionScalar, int iteration, int step, int deltaScalar,
                                int neighborBiasScalar) {
    const int scaledCx = cx << resolutionScalar;  // :50
    const int scaledCy = cy << resolutionScalar;  // :51
    const int threadIndex2D = cy * lowDimX + cx;  // :52
    uint32_t delta = 0, offsetBias = 0, neighborBias = 0;

    const int16_t idealOffsetX = offsetArray[threadIndex2D];                      // :65
    const int16_t idealOffsetY = offsetArray[lowDimY * lowDimX + threadIndex2D];  // :66
    int16_t relX = 0, relY = 0;
    if (!(step & 1)) {  // :69-75
        relX = (int16_t)((cz % searchWindowSize) - (searchWindowSize / 2));
        relX = (int16_t)signedSquare(relX);
    } else {
        relY = (int16_t)((cz % searchWindowSize) - (searchWindowSize / 2));
        relY = (int16_t)signedSquare(relY);
    }
    const int16_t offsetX = (int16_t)(idealOffsetX + relX);  // :76
    const int16_t offsetY = (int16_t)(idealOffsetY + relY);  // :77
    int newCx = scaledCx + offsetX;                          // :78
    int newCy = scaledCy + offsetY;                          // :79

    if (scaledCx < 0 || scaledCx >= dimX || scaledCy < 0 || scaledCy >= dimY) {  // :82
        delta = 0;
    } else {
        newCx = mirrorSearch(newCx, dimX);  // :86-90
        newCy = mirrorSearch(newCy, dimY);  // :91-95
        const T* uv1 = frame1 + (size_t)dimY * inputStride;
        const T* uv2 = frame2 + (size_t)dimY * inputStride;
        auto S = [](T v) { return (int)PixelTraits<T>::searchSample(v); };
        // :98-100
        delta = (uint32_t)std::abs(S(frame1[(size_t)newCy * inputStride + newCx]) - S(frame2[(size_t)scaledCy * inputStride + scaledCx])) +
                (uint32_t)std::abs(S(uv1[(size_t)(newCy >> 1) * inputStride + (newCx & ~1)]) - S(uv2[(size_t)(scaledCy >> 1) * inputStride + (scaledCx & ~1)])) +
                (uint32_t)std::abs(S(uv1[(size_t)(newCy >> 1) * inputStride + (newCx & ~1) + 1]) - S(uv2[(size_t)(scaledCy >> 1) * inputStride + (scaledCx & ~1) + 1]));
        delta <<= deltaScalar;  // :101
    }

    // :105-109  abs(short) -> ushort
    offsetBias = (uint32_t)(uint16_t)std::abs((int)(!step ? offsetX : offsetY));

    if (iteration >= 4) {  // FIRST_NEIGHBOR_ITERATION :3,112
        const int nb[4][2] = {{0, 2 * windowSize}, {2 * windowSize, 0}, {-2 * windowSize, 0}, {0, -2 * windowSize}};  // :114-119
        for (int i = 0; i < 4; ++i) {
            const int nx = std::min(std::max(cx + nb[i][0], 0), lowDimX - 1);  // :6-9
            const int ny = std::min(std::max(cy + nb[i][1], 0), lowDimY - 1);
            const int16_t nOff = offsetArray[(!step ? 0 : lowDimY * lowDimX) + ny * lowDimX + nx];  // :127-131
            const uint16_t diff = (uint16_t)std::abs((int)nOff - (int)(!step ? offsetX : offsetY));    // :134-138
            neighborBias += diff;                                                                      // :141
        }
        neighborBias <<= neighborBiasScalar;  // :143
    }
    return delta + offsetBias + neighborBias;  // :148,151
}

// HR/calcDeltaSumsKernelSDR.h:36-191, launched over (ceil(lw/8)*8, ceil(lh/8)*8, R) (HR/opticalFlowCalcSDR.cpp:88).
// Net effect under lock-step semantics: sums[cz][window representative] += v for every in-range work-item.
// The caller zero-fills `sums` first (HR/opticalFlowCalcSDR.cpp:75-76).
template <typename T>
void calcDeltaSums(uint32_t* sums, const T* frame1, const T* frame2, const int16_t* offsetArray, int dimY, int dimX,
                   int inputStride, int lowDimY, int lowDimX, int windowSize, int searchWindowSize, int resolutionScalar,
                   int iteration, int step, int deltaScalar, int neighborBiasScalar) {
    const int nWy = (lowDimY + windowSize - 1) / windowSize;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int cz = 0; cz < searchWindowSize; ++cz) {
        for (int wy = 0; wy < nWy; ++wy) {
            const int yEnd = std::min((wy + 1) * windowSize, lowDimY);
            for (int cy = wy * windowSize; cy < yEnd; ++cy) {
                for (int cx = 0; cx < lowDimX; ++cx) {
                    const uint32_t v = deltaOfWorkItem<T>(frame1, frame2, offsetArray, cx, cy, cz, dimY, dimX, inputStride,
                                                          lowDimY, lowDimX, windowSize, searchWindowSize, resolutionScalar,
                                                          iteration, step, deltaScalar, neighborBiasScalar);
                    if (windowSize == 1) {  // :146-149
                        sums[(size_t)cz * lowDimY * lowDimX + (size_t)cy * lowDimX + cx] = v;
                    } else {  // :184-190
                        const int wxr = (cx / windowSize) * windowSize;
                        const int wyr = (cy / windowSize) * windowSize;
                        sums[(size_t)cz * lowDimY * lowDimX + (size_t)wyr * lowDimX + wxr] += v;
                    }
                }
            }
        }
    }
}

// HR/determineLowestLayerKernelSDR.h:4-28
void determineLowestLayer(const uint32_t* sums, uint8_t* lowestLayerArray, int windowSize, int searchWindowSize, int lowDimY,
                          int lowDimX) {
    const size_t plane = (size_t)lowDimY * lowDimX;
#pragma omp parallel for schedule(static)
    for (int cy = 0; cy < lowDimY; ++cy) {
        if (cy % windowSize != 0) continue;
        for (int cx = 0; cx < lowDimX; cx += windowSize) {
            uint8_t lowestLayer = 0;
            for (int z = 1; z < searchWindowSize; ++z) {
                if (sums[z * plane + (size_t)cy * lowDimX + cx] < sums[lowestLayer * plane + (size_t)cy * lowDimX + cx]) lowestLayer = (uint8_t)z;
            }
            lowestLayerArray[(size_t)cy * lowDimX + cx] = lowestLayer;
        }
    }
}

// HR/adjustOffsetArrayKernelSDR.h:4-21
void adjustOffsetArray(int16_t* offsetArray, const uint8_t* lowestLayerArray, int windowSize, int searchWindowSize, int lowDimY,
                       int lowDimX, int step) {
#pragma omp parallel for schedule(static)
    for (int cy = 0; cy < lowDimY; ++cy) {
        for (int cx = 0; cx < lowDimX; ++cx) {
            const int wx = (cx / windowSize) * windowSize;
            const int wy = (cy / windowSize) * windowSize;
            const uint8_t lowestLayer = lowestLayerArray[(size_t)wy * lowDimX + wx];
            const int16_t idealRelOffset = (int16_t)((lowestLayer % searchWindowSize) - (searchWindowSize / 2));
            int16_t& o = offsetArray[(size_t)(step & 1) * lowDimY * lowDimX + (size_t)cy * lowDimX + cx];
            o = (int16_t)(o + signedSquare(idealRelOffset));
        }
    }
}

// HR/blurFlowKernelSDR.h:17-92 — the local tile only caches mirrored loads; the result is this formula (:80-90).
void blurFlow(const int16_t* offsetArray, int16_t* blurred, int dimY, int dimX) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int gz = 0; gz < 2; ++gz) {
        for (int gy = 0; gy < dimY; ++gy) {
            for (int gx = 0; gx < dimX; ++gx) {
                int sum = 0;
                for (int ky = -4; ky < 4; ++ky)
                    for (int kx = -4; kx < 4; ++kx)
                        sum += offsetArray[(size_t)gz * dimX * dimY + (size_t)mirrorBlur(gy + ky, dimY) * dimX + mirrorBlur(gx + kx, dimX)];
                blurred[(size_t)gz * dimX * dimY + (size_t)gy * dimX + gx] = (int16_t)(sum / 64);
            }
        }
    }
}

// HR/warpFrameKernelSDR.h:116-184 / HR/warpFrameKernelHDR.h:116-184 — one launch (cz = 0 luma, 1 chroma).
template <typename T>
void warpFrame(const T* sourceFrame12, const T* sourceFrame21, const int16_t* offsetArray, T* outputFrame, float frameScalar12,
               float frameScalar21, int lowDimY, int lowDimX, int dimY, int dimX, int inputStride, int outputStride,
               int resolutionScalar, int frameOutputMode, float black_level, float white_level, int cz) {
    const int verticalOffset = dimY >> 2;  // :125
    const int rows = dimY >> cz;           // :128
#pragma omp parallel for schedule(static)
    for (int cy = 0; cy < rows; ++cy) {
        for (int cx = 0; cx < dimX; ++cx) {
            int adjCx = cx, adjCy = cy;
            T* out = &outputFrame[(size_t)cz * dimY * outputStride + (size_t)cy * outputStride + cx];
            const size_t inPlane = (size_t)cz * dimY * inputStride;
            if (frameOutputMode == 5 && cx < (dimX >> 1)) {  // :133-135
                *out = sourceFrame12[inPlane + (size_t)cy * inputStride + cx];
                continue;
            } else if (frameOutputMode == 6) {  // :136-149
                const bool inBand = cy >= (verticalOffset >> cz) && cy < ((verticalOffset >> cz) + (dimY >> (1 + cz)));
                const bool isInLeftSide = inBand && cx < (dimX >> 1);
                const bool isInRightSide = inBand && cx >= (dimX >> 1) && cx < dimX;
                if (isInLeftSide) {
                    *out = sourceFrame12[inPlane + (size_t)((cy - (verticalOffset >> cz)) << 1) * inputStride + (cx << 1) + (cz ? (cx & 1) : 0)];
                    continue;
                } else if (isInRightSide) {
                    adjCx = (cx - (dimX >> 1)) << 1;
                    adjCy = (cy - (verticalOffset >> cz)) << 1;
                } else {
                    *out = (T)(cz ? PixelTraits<T>::midInt : 0);
                    continue;
                }
            }
            // :153-158
            const int scaledCx = cz ? ((adjCx >> resolutionScalar) & ~1) : (adjCx >> resolutionScalar);
            const int scaledCy = cz ? ((adjCy >> resolutionScalar) << 1) : (adjCy >> resolutionScalar);
            const size_t lowPlane = (size_t)lowDimY * lowDimX;
            const int offsetX12 = offsetArray[(size_t)scaledCy * lowDimX + scaledCx];
            const int offsetY12 = offsetArray[lowPlane + (size_t)scaledCy * lowDimX + scaledCx];
            const int gy = std::min(std::max(scaledCy - (offsetY12 >> resolutionScalar), 0), lowDimY - 1);
            const int gx = std::min(std::max(scaledCx - (offsetX12 >> resolutionScalar), 0), lowDimX - 1);
            const int offsetX21 = offsetArray[(size_t)gy * lowDimX + gx];
            const int offsetY21 = offsetArray[lowPlane + (size_t)gy * lowDimX + gx];

            if (frameOutputMode == 4) {  // :161-164
                const unsigned m = (unsigned)(std::abs(offsetX12) + std::abs(offsetY12)) << PixelTraits<T>::greyShift;
                *out = (T)(cz ? (unsigned)PixelTraits<T>::midInt : std::min(m, PixelTraits<T>::greyMax));
                continue;
            }

            // :167-170
            const float vs = cz ? 0.5f : 1.0f;
            const int dimYc = cz ? (dimY >> 1) : dimY;
            const int newCx12 = mirrorWarp(adjCx + (int)std::round((float)offsetX12 * frameScalar12), dimX);
            const int newCy12 = mirrorWarp(adjCy + (int)std::round((float)offsetY12 * frameScalar12 * vs), dimYc);
            const int newCx21 = mirrorWarp(adjCx - (int)std::round((float)offsetX21 * frameScalar21), dimX);
            const int newCy21 = mirrorWarp(adjCy - (int)std::round((float)offsetY21 * frameScalar21 * vs), dimYc);

            const int xmask = cz ? ~1 : ~0;
            const int xpar = cx & (cz ? 1 : 0);
            const T a = sourceFrame12[inPlane + (size_t)newCy12 * inputStride + (newCx12 & xmask) + xpar];
            const T b = sourceFrame21[inPlane + (size_t)newCy21 * inputStride + (newCx21 & xmask) + xpar];
            if (frameOutputMode == 0) {  // :172-173
                *out = a;
            } else if (frameOutputMode == 1) {  // :174-175
                *out = b;
            } else {  // :176-183
                // OpenCL C contracts a*b+c by default; the NVIDIA OpenCL compiler emits fma(a, t21, b*t12) for this
                // line (established against the reference run on a B200: bit-exact with this form, up to 2 LSB off
                // without it once the levels gain exceeds 1 — see tests/test_oracle_golden.py).
                uint16_t blendedValue = (uint16_t)fmaf((float)a, frameScalar21, (float)b * frameScalar12);
                if (frameOutputMode == 3) {
                    blendedValue = visualizeFlow<T>((short)-offsetX12, (short)-offsetY12, (T)blendedValue, cz + (cx & (cz ? 1 : 0)),
                                                    resolutionScalar <= 2 ? 4 : 1);
                }
                *out = (T)(cz ? applyLevelsUV<T>((float)blendedValue, white_level) : applyLevelsY<T>((float)blendedValue, black_level, white_level));
            }
        }
    }
}

// HR/copyFrameKernelSDR.h:12-25 / HDR :12-25
template <typename T>
void copyFrameKernel(const T* sourceFrame, T* outputFrame, int dimY, int dimX, int inputStride, int outputStride, float black_level,
                     float white_level, int cz) {
    const int rows = dimY >> cz;
#pragma omp parallel for schedule(static)
    for (int cy = 0; cy < rows; ++cy) {
        for (int cx = 0; cx < dimX; ++cx) {
            const T value = sourceFrame[(size_t)cz * dimY * inputStride + (size_t)cy * inputStride + cx];
            outputFrame[(size_t)cz * dimY * outputStride + (size_t)cy * outputStride + cx] =
                (T)(cz ? applyLevelsUV<T>((float)value, white_level) : applyLevelsY<T>((float)value, black_level, white_level));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host schedule — HR/opticalFlowCalcSDR.cpp / HR/opticalFlowCalcHDR.cpp
// ---------------------------------------------------------------------------------------------

struct PassTap {
    int windowSize, iteration, step;
    std::vector<uint32_t> sums;     // [R][lh][lw] after calcDeltaSums
    std::vector<uint8_t> layers;    // [lh][lw] after determineLowestLayer (only representatives are meaningful)
    std::vector<int16_t> offsets;   // [2][lh][lw] after adjustOffsetArray
};

struct Calc {
    bool hdr;
    int bpp;
    // public fields of HR/opticalFlowCalc.h:26-50
    int m_frameWidth, m_frameHeight, m_inputStride, m_outputStride;
    float m_outputBlackLevel, m_outputWhiteLevel;
    int m_opticalFlowResScalar, m_opticalFlowFrameWidth, m_opticalFlowFrameHeight, m_opticalFlowSearchRadius;
    double m_ofcCalcTime, m_ofcAvgCalcTime, m_ofcPeakCalcTime;
    int m_ofcCalcCount;
    double m_ofcCalcTimeSum, m_warpCalcTime;
    int m_deltaScalar, m_neighborBiasScalar;
    unsigned int m_totalFrameDelta;
    unsigned int m_frameCount;
    // buffers (HR/opticalFlowCalcSDR.cpp:272-280)
    std::vector<uint8_t> input[3];
    std::vector<uint8_t> output;
    std::vector<int16_t> offsetArray, blurred[2];
    std::vector<uint32_t> summedDelta;
    std::vector<uint8_t> lowestLayer;
    int in[3] = {0, 1, 2};  // rotation of m_inputFrameArray
    int bl[2] = {0, 1};     // rotation of m_blurredOffsetArray
    bool tapsEnabled = false;
    std::vector<PassTap> taps;
    std::chrono::steady_clock::time_point ofcStart, warpStart;
};

double secondsSince(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

extern "C" {

// ---- kernel-level entry points (hdr: 0 = uchar pixels, 1 = ushort pixels) ---------------------

void orc_calc_delta_sums(uint32_t* sums, const void* frame1, const void* frame2, const int16_t* offsetArray, int dimY, int dimX,
                         int inputStride, int lowDimY, int lowDimX, int windowSize, int searchWindowSize, int resolutionScalar,
                         int iteration, int step, int deltaScalar, int neighborBiasScalar, int hdr) {
    if (hdr)
        calcDeltaSums<uint16_t>(sums, (const uint16_t*)frame1, (const uint16_t*)frame2, offsetArray, dimY, dimX, inputStride, lowDimY,
                                lowDimX, windowSize, searchWindowSize, resolutionScalar, iteration, step, deltaScalar, neighborBiasScalar);
    else
        calcDeltaSums<uint8_t>(sums, (const uint8_t*)frame1, (const uint8_t*)frame2, offsetArray, dimY, dimX, inputStride, lowDimY,
                               lowDimX, windowSize, searchWindowSize, resolutionScalar, iteration, step, deltaScalar, neighborBiasScalar);
}

void orc_determine_lowest_layer(const uint32_t* sums, uint8_t* lowestLayerArray, int windowSize, int searchWindowSize, int lowDimY,
                                int lowDimX) {
    determineLowestLayer(sums, lowestLayerArray, windowSize, searchWindowSize, lowDimY, lowDimX);
}

void orc_adjust_offset_array(int16_t* offsetArray, const uint8_t* lowestLayerArray, int windowSize, int searchWindowSize,
                             int lowDimY, int lowDimX, int step) {
    adjustOffsetArray(offsetArray, lowestLayerArray, windowSize, searchWindowSize, lowDimY, lowDimX, step);
}

void orc_blur_flow(const int16_t* offsetArray, int16_t* blurred, int dimY, int dimX) { blurFlow(offsetArray, blurred, dimY, dimX); }

void orc_warp_frame(const void* src12, const void* src21, const int16_t* offsetArray, void* out, float frameScalar12,
                    float frameScalar21, int lowDimY, int lowDimX, int dimY, int dimX, int inputStride, int outputStride,
                    int resolutionScalar, int frameOutputMode, float black, float white, int cz, int hdr) {
    if (hdr)
        warpFrame<uint16_t>((const uint16_t*)src12, (const uint16_t*)src21, offsetArray, (uint16_t*)out, frameScalar12, frameScalar21,
                            lowDimY, lowDimX, dimY, dimX, inputStride, outputStride, resolutionScalar, frameOutputMode, black, white, cz);
    else
        warpFrame<uint8_t>((const uint8_t*)src12, (const uint8_t*)src21, offsetArray, (uint8_t*)out, frameScalar12, frameScalar21,
                           lowDimY, lowDimX, dimY, dimX, inputStride, outputStride, resolutionScalar, frameOutputMode, black, white, cz);
}

void orc_copy_frame_kernel(const void* src, void* out, int dimY, int dimX, int inputStride, int outputStride, float black,
                           float white, int cz, int hdr) {
    if (hdr)
        copyFrameKernel<uint16_t>((const uint16_t*)src, (uint16_t*)out, dimY, dimX, inputStride, outputStride, black, white, cz);
    else
        copyFrameKernel<uint8_t>((const uint8_t*)src, (uint8_t*)out, dimY, dimX, inputStride, outputStride, black, white, cz);
}

// ---- calculator (host schedule) ---------------------------------------------------------------

// HR/opticalFlowCalcSDR.cpp:206-280 / HR/opticalFlowCalcHDR.cpp:211-289
void* orc_ofc_create(int frameHeight, int frameWidth, int inputStride, int outputStride, int deltaScalar, int neighborScalar,
                     float blackLevel, float whiteLevel, int maxCalcRes, int hdr) {
    Calc* c = new Calc();
    c->hdr = hdr != 0;
    c->bpp = hdr ? 2 : 1;
    c->m_frameWidth = frameWidth;
    c->m_frameHeight = frameHeight;
    c->m_inputStride = inputStride > 0 ? inputStride : frameWidth;
    c->m_outputStride = outputStride > 0 ? outputStride : frameWidth;
    c->m_outputBlackLevel = blackLevel;
    c->m_outputWhiteLevel = whiteLevel;
    c->m_opticalFlowSearchRadius = 5;  // MIN_SEARCH_RADIUS, HR/config.h:8
    c->m_opticalFlowResScalar = 0;
    while ((frameHeight >> c->m_opticalFlowResScalar) > maxCalcRes) c->m_opticalFlowResScalar++;
    c->m_opticalFlowFrameWidth = (int)std::ceil(frameWidth / std::pow(2, c->m_opticalFlowResScalar));
    c->m_opticalFlowFrameHeight = (int)std::ceil(frameHeight / std::pow(2, c->m_opticalFlowResScalar));
    c->m_ofcCalcTime = c->m_ofcAvgCalcTime = c->m_ofcPeakCalcTime = 0.0;
    c->m_ofcCalcCount = 0;
    c->m_ofcCalcTimeSum = 0.0;
    c->m_warpCalcTime = 0.0;
    c->m_deltaScalar = deltaScalar;
    c->m_neighborBiasScalar = neighborScalar;
    c->m_totalFrameDelta = 0;
    c->m_frameCount = 0;
    const size_t lw = c->m_opticalFlowFrameWidth, lh = c->m_opticalFlowFrameHeight;
    const size_t inBytes = (size_t)(1.5 * frameHeight * c->m_inputStride) * c->bpp;
    const size_t outBytes = (size_t)(1.5 * frameHeight * c->m_outputStride) * c->bpp;
    for (auto& v : c->input) v.assign(inBytes, 0);
    c->output.assign(outBytes, 0);
    c->offsetArray.assign(2 * lw * lh, 0);
    c->blurred[0].assign(2 * lw * lh, 0);
    c->blurred[1].assign(2 * lw * lh, 0);
    c->summedDelta.assign(16 * lw * lh, 0);  // MAX_SEARCH_RADIUS layers, HR/opticalFlowCalcSDR.cpp:279
    c->lowestLayer.assign(lw * lh, 0);
    return c;
}

void orc_ofc_destroy(void* h) { delete (Calc*)h; }

// HR/opticalFlowCalcSDR.cpp:19-29 / HDR :19-29
int orc_ofc_update_frame(void* h, const uint8_t* inputPlanes) {
    Calc* c = (Calc*)h;
    c->ofcStart = std::chrono::steady_clock::now();
    const size_t n = (size_t)c->bpp * ((size_t)c->m_frameHeight * c->m_inputStride + (size_t)(c->m_frameHeight / 2) * c->m_inputStride);
    std::memcpy(c->input[c->in[0]].data(), inputPlanes, n);
    const int t = c->in[0];
    c->in[0] = c->in[1];
    c->in[1] = c->in[2];
    c->in[2] = t;
    c->m_frameCount++;
    return 0;
}

// HR/opticalFlowCalcSDR.cpp:31-42
int orc_ofc_download_frame(void* h, uint8_t* outputPlanes) {
    Calc* c = (Calc*)h;
    const size_t n = (size_t)c->bpp * ((size_t)c->m_frameHeight * c->m_outputStride + (size_t)(c->m_frameHeight / 2) * c->m_outputStride);
    std::memcpy(outputPlanes, c->output.data(), n);
    c->m_warpCalcTime = secondsSince(c->warpStart);
    return 0;
}

// HR/opticalFlowCalcSDR.cpp:44-139 / HDR :44-139
int orc_ofc_calculate_optical_flow(void* h) {
    Calc* c = (Calc*)h;
    const int lw = c->m_opticalFlowFrameWidth, lh = c->m_opticalFlowFrameHeight;
    const int R = c->m_opticalFlowSearchRadius;  // :46
    // :49-59
    int windowSize = 1;
    int maxDim = std::max(lw, lh);
    if (maxDim && !(maxDim & (maxDim - 1))) {
        windowSize = maxDim;
    } else {
        while (maxDim & (maxDim - 1)) maxDim &= (maxDim - 1);
        windowSize = maxDim << 1;
    }
    windowSize /= 2;
    const int iterations = (int)std::log2(windowSize);  // :62-65 with NUM_ITERATIONS == 0
    std::fill(c->offsetArray.begin(), c->offsetArray.end(), 0);  // :68-69
    c->taps.clear();
    const void* f1 = c->input[c->in[1]].data();
    const void* f2 = c->input[c->in[2]].data();
    for (int iter = 0; iter < iterations; iter++) {
        for (int step = 0; step < 2; step++) {
            std::fill(c->summedDelta.begin(), c->summedDelta.begin() + (size_t)R * lw * lh, 0u);  // :75-76
            orc_calc_delta_sums(c->summedDelta.data(), f1, f2, c->offsetArray.data(), c->m_frameHeight, c->m_frameWidth, c->m_inputStride,
                                lh, lw, windowSize, R, c->m_opticalFlowResScalar, iter, step, c->m_deltaScalar, c->m_neighborBiasScalar,
                                c->hdr);  // :79-88
            if (iter == 0 && step == 0) {  // :91-94
                c->m_totalFrameDelta = c->summedDelta[(size_t)((R / 2) - 1) * lh * lw];
                c->m_totalFrameDelta /= (unsigned)(lh * lw * (c->hdr ? 6 : 10));
            }
            determineLowestLayer(c->summedDelta.data(), c->lowestLayer.data(), windowSize, R, lh, lw);  // :97-99
            adjustOffsetArray(c->offsetArray.data(), c->lowestLayer.data(), windowSize, R, lh, lw, step);  // :102-106
            if (c->tapsEnabled) {
                PassTap t;
                t.windowSize = windowSize;
                t.iteration = iter;
                t.step = step;
                t.sums.assign(c->summedDelta.begin(), c->summedDelta.begin() + (size_t)R * lw * lh);
                t.layers = c->lowestLayer;
                t.offsets = c->offsetArray;
                c->taps.push_back(std::move(t));
            }
        }
        windowSize = std::max(windowSize >> 1, 1);  // :110
    }
    blurFlow(c->offsetArray.data(), c->blurred[c->bl[0]].data(), lh, lw);  // :113-116
    std::swap(c->bl[0], c->bl[1]);                                          // :121-123
    // :124-138
    c->m_ofcCalcTime = secondsSince(c->ofcStart);
    if (c->m_ofcCalcCount >= 240) {  // CALC_TIME_INTERVAL
        c->m_ofcAvgCalcTime = c->m_ofcCalcTimeSum / c->m_ofcCalcCount;
        c->m_ofcCalcCount = 0;
        c->m_ofcCalcTimeSum = 0.0;
        c->m_ofcPeakCalcTime = c->m_ofcCalcTime;
    }
    c->m_ofcCalcCount++;
    c->m_ofcCalcTimeSum += c->m_ofcCalcTime;
    if (c->m_ofcCalcTime > c->m_ofcPeakCalcTime) c->m_ofcPeakCalcTime = c->m_ofcCalcTime;
    return 0;
}

// HR/opticalFlowCalcSDR.cpp:141-168 / HDR :141-170.  Returns non-zero where the reference throws.
int orc_ofc_warp_frames(void* h, float blendingScalar, int frameOutputMode) {
    Calc* c = (Calc*)h;
    if (blendingScalar > 1.0f) return 1;  // :143-146
    const float frameScalar12 = blendingScalar;
    const float frameScalar21 = 1.0f - blendingScalar;
    float black = c->m_outputBlackLevel, white = c->m_outputWhiteLevel;
    if (c->hdr) {  // HR/opticalFlowCalcHDR.cpp:151-152
        black = c->m_outputBlackLevel * 256.0f;
        white = c->m_outputWhiteLevel * 256.0f;
    }
    c->warpStart = std::chrono::steady_clock::now();
    for (int cz = 0; cz < 2; ++cz)
        orc_warp_frame(c->input[c->in[0]].data(), c->input[c->in[1]].data(), c->blurred[c->bl[0]].data(), c->output.data(), frameScalar12,
                       frameScalar21, c->m_opticalFlowFrameHeight, c->m_opticalFlowFrameWidth, c->m_frameHeight, c->m_frameWidth,
                       c->m_inputStride, c->m_outputStride, c->m_opticalFlowResScalar, frameOutputMode, black, white, cz, c->hdr);
    return 0;
}

// HR/opticalFlowCalcSDR.cpp:170-183 / HDR :172-188
int orc_ofc_copy_frame(void* h) {
    Calc* c = (Calc*)h;
    const int frameIndex = c->m_frameCount >= 3 ? 0 : c->m_frameCount >= 2 ? 1 : 2;
    float black = c->m_outputBlackLevel, white = c->m_outputWhiteLevel;
    if (c->hdr) {
        black = c->m_outputBlackLevel * 256.0f;
        white = c->m_outputWhiteLevel * 256.0f;
    }
    c->warpStart = std::chrono::steady_clock::now();
    for (int cz = 0; cz < 2; ++cz)
        orc_copy_frame_kernel(c->input[c->in[frameIndex]].data(), c->output.data(), c->m_frameHeight, c->m_frameWidth, c->m_inputStride,
                              c->m_outputStride, black, white, cz, c->hdr);
    return 0;
}

// ---- field access -------------------------------------------------------------------------------
struct orc_state {
    int frameWidth, frameHeight, inputStride, outputStride;
    float outputBlackLevel, outputWhiteLevel;
    int resScalar, flowWidth, flowHeight, searchRadius;
    double ofcCalcTime, ofcAvgCalcTime, ofcPeakCalcTime, warpCalcTime;
    int deltaScalar, neighborBiasScalar;
    unsigned int totalFrameDelta, frameCount;
};

void orc_ofc_get_state(void* h, orc_state* s) {
    Calc* c = (Calc*)h;
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

// the fields the filter writes (HR/HopperRender.cpp:840,1386-1389,1448,1457)
void orc_ofc_set_params(void* h, int searchRadius, int deltaScalar, int neighborBiasScalar, float black, float white) {
    Calc* c = (Calc*)h;
    c->m_opticalFlowSearchRadius = searchRadius;
    c->m_deltaScalar = deltaScalar;
    c->m_neighborBiasScalar = neighborBiasScalar;
    c->m_outputBlackLevel = black;
    c->m_outputWhiteLevel = white;
}
void orc_ofc_set_frame_count(void* h, unsigned int n) { ((Calc*)h)->m_frameCount = n; }

// ---- taps ---------------------------------------------------------------------------------------
void orc_ofc_enable_taps(void* h, int on) { ((Calc*)h)->tapsEnabled = on != 0; }
int orc_ofc_num_passes(void* h) { return (int)((Calc*)h)->taps.size(); }
int orc_ofc_pass_info(void* h, int pass, int* windowSize, int* iteration, int* step) {
    Calc* c = (Calc*)h;
    if (pass < 0 || pass >= (int)c->taps.size()) return 1;
    *windowSize = c->taps[pass].windowSize;
    *iteration = c->taps[pass].iteration;
    *step = c->taps[pass].step;
    return 0;
}
// which: 0 sums (uint32 [R][lh][lw]), 1 layers (u8 [lh][lw]), 2 offsets (int16 [2][lh][lw])
int orc_ofc_read_pass_tap(void* h, int pass, int which, void* dst, size_t bytes) {
    Calc* c = (Calc*)h;
    if (pass < 0 || pass >= (int)c->taps.size()) return 1;
    const PassTap& t = c->taps[pass];
    const void* src = nullptr;
    size_t n = 0;
    if (which == 0) { src = t.sums.data(); n = t.sums.size() * 4; }
    else if (which == 1) { src = t.layers.data(); n = t.layers.size(); }
    else if (which == 2) { src = t.offsets.data(); n = t.offsets.size() * 2; }
    else return 2;
    if (bytes != n) return 3;
    std::memcpy(dst, src, n);
    return 0;
}
// which: 0 offsetArray, 1 blurred flow read by warpFrames (blurred[0]), 2 freshest blurred flow (blurred[1]), 3 output frame
int orc_ofc_read_buffer(void* h, int which, void* dst, size_t bytes) {
    Calc* c = (Calc*)h;
    const void* src = nullptr;
    size_t n = 0;
    if (which == 0) { src = c->offsetArray.data(); n = c->offsetArray.size() * 2; }
    else if (which == 1) { src = c->blurred[c->bl[0]].data(); n = c->blurred[0].size() * 2; }
    else if (which == 2) { src = c->blurred[c->bl[1]].data(); n = c->blurred[1].size() * 2; }
    else if (which == 3) { src = c->output.data(); n = c->output.size(); }
    else return 2;
    if (bytes > n) return 3;
    std::memcpy(dst, src, bytes);
    return 0;
}
// write a blurred flow directly (lets tests drive warpFrames with a chosen flow). which as above (1 or 2).
int orc_ofc_write_flow(void* h, int which, const int16_t* src, size_t count) {
    Calc* c = (Calc*)h;
    std::vector<int16_t>& v = c->blurred[which == 1 ? c->bl[0] : c->bl[1]];
    if (count != v.size()) return 3;
    std::memcpy(v.data(), src, count * 2);
    return 0;
}

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
