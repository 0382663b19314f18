/* example_c_api.c — the C ABI of include/hrb.h from plain C99: the call sequence of one interpolated frame
 * (HopperRender/HopperRender.cpp:953-957, 1180-1186) on synthetic NV12 frames.
 *
 *   gcc -std=c99 -Iinclude tools/example_c_api.c -o example -Lhopperrender_b200 -lhrb -Wl,-rpath,$PWD/hopperrender_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hrb.h"

static int fail(const char* what) {
    fprintf(stderr, "%s failed: %s\n", what, hrb_last_error());
    return 2;
}

int main(void) {
    const int W = 640, H = 360;
    const size_t frameBytes = (size_t)W * H * 3 / 2; /* NV12: H luma rows, H/2 interleaved chroma rows */
    unsigned char* in = (unsigned char*)malloc(frameBytes);
    unsigned char* out = (unsigned char*)malloc(frameBytes);
    hrb_ofc_desc desc;
    hrb_ofc_state st;
    hrb_ofc* h = NULL;
    int t, x, y;
    if (!in || !out) return 1;

    memset(&desc, 0, sizeof(desc));
    desc.frame_height = H;
    desc.frame_width = W;
    desc.input_stride = 0;  /* <= 0: the frame width, as in the reference constructor */
    desc.output_stride = 0;
    desc.delta_scalar = 8;
    desc.neighbor_scalar = 6;
    desc.black_level = 0.0f;
    desc.white_level = 255.0f;
    desc.max_calc_res = 270;
    desc.is_hdr = 0;
    desc.device_ordinal = 0;
    if (hrb_ofc_create(&h, &desc) != HRB_OK) return fail("hrb_ofc_create");

    for (t = 0; t < 3; ++t) { /* three frames before the first interpolation (m_frameCount >= 3) */
        for (y = 0; y < H; ++y)
            for (x = 0; x < W; ++x) in[(size_t)y * W + x] = (unsigned char)(((x + 4 * t) / 8 + y / 8) * 9);
        memset(in + (size_t)W * H, 128, (size_t)W * H / 2);
        if (hrb_ofc_update_frame(h, in) != HRB_OK) return fail("hrb_ofc_update_frame");
    }
    if (hrb_ofc_calculate_optical_flow(h) != HRB_OK) return fail("hrb_ofc_calculate_optical_flow");
    if (hrb_ofc_warp_frames(h, 0.5f, 2 /* BlendedFrame */) != HRB_OK) return fail("hrb_ofc_warp_frames");
    if (hrb_ofc_download_frame(h, out) != HRB_OK) return fail("hrb_ofc_download_frame");
    if (hrb_ofc_get_state(h, &st) != HRB_OK) return fail("hrb_ofc_get_state");
    printf("flow %dx%d, frame delta %u, flow time %.3f ms, first output bytes %u %u %u\n", st.flow_width, st.flow_height,
           st.total_frame_delta, st.ofc_calc_time * 1e3, out[0], out[1], out[2]);
    hrb_ofc_destroy(h);
    free(in);
    free(out);
    return 0;
}
