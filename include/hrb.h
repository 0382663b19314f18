/* hrb.h — C ABI of hopperrender_b200: HopperRender's optical-flow frame-interpolation hot path on B200.
 *
 * This is the drop-in boundary.  The reference has no FFI layer of its own: the boundary is the C++
 * class OpticalFlowCalc / OpticalFlowCalcSDR / OpticalFlowCalcHDR (HopperRender/opticalFlowCalc.h:24-138,
 * opticalFlowCalcSDR.h:13-56, opticalFlowCalcHDR.h:13-56) that CHopperRender::DeliverToRenderer drives
 * (HopperRender/HopperRender.cpp:918-924, 953-957, 1180-1186).  Each entry point below replaces one
 * member of that class; include/opticalFlowCalc*.h re-creates the class on top of this ABI so the
 * filter's host code compiles unchanged.
 *
 * Conventions: plain C, opaque handle, every call returns HRB_OK (0) or an error code; the message of
 * the last error on the calling thread is available from hrb_last_error().  No exception crosses this
 * boundary.  One handle = one video stream on one GPU, used by one thread at a time; distinct handles
 * are independent (own CUDA stream unless the caller supplies one).  There is no CPU fallback: every
 * call fails with HRB_ERR_CUDA when no sm_100 device is usable.
 *
 * Strides are in ELEMENTS (pixels), as in the reference (opticalFlowCalcHDR.cpp:20): bytes = elements
 * * (is_hdr ? 2 : 1).  Frames are NV12 (8-bit) or P010 (10 significant bits in the MSBs of 16):
 * `height` luma rows of `stride` elements, then height/2 rows of interleaved U,V.
 */
#ifndef HRB_H_
#define HRB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define HRB_API __declspec(dllexport)
#else
#define HRB_API __attribute__((visibility("default")))
#endif

/* error codes */
enum {
    HRB_OK = 0,
    HRB_ERR_INVALID_ARG = 1,   /* bad pointer / geometry / mode */
    HRB_ERR_CUDA = 2,          /* a CUDA runtime call failed (replaces CHECK_ERROR, opticalFlowCalc.h:15-22) */
    HRB_ERR_BLEND_RANGE = 3,   /* blendingScalar > 1 (replaces the throw at opticalFlowCalcSDR.cpp:143-146) */
    HRB_ERR_NO_DEVICE = 4,     /* no device with enough memory (replaces detectDevices, opticalFlowCalc.cpp:97-109) */
    HRB_ERR_STATE = 5          /* call not valid in the current state (e.g. tap read without tap mode) */
};

/* frameOutputMode values (HopperRender/HopperRender.h:10-18) */
enum {
    HRB_MODE_WARPED_FRAME_12 = 0,
    HRB_MODE_WARPED_FRAME_21 = 1,
    HRB_MODE_BLENDED_FRAME = 2,
    HRB_MODE_HSV_FLOW = 3,
    HRB_MODE_GREY_FLOW = 4,
    HRB_MODE_SIDE_BY_SIDE_1 = 5,
    HRB_MODE_SIDE_BY_SIDE_2 = 6
};

typedef struct hrb_ofc hrb_ofc;

/* Constructor arguments of OpticalFlowCalcSDR/HDR (opticalFlowCalcSDR.h:15-17) + target selection. */
typedef struct hrb_ofc_desc {
    int frame_height;
    int frame_width;
    int input_stride;    /* elements; <= 0 means frame_width (opticalFlowCalcSDR.cpp:212) */
    int output_stride;   /* elements; <= 0 means frame_width (opticalFlowCalcSDR.cpp:213) */
    int delta_scalar;
    int neighbor_scalar;
    float black_level;
    float white_level;
    int max_calc_res;    /* flow height is halved until <= this (opticalFlowCalcSDR.cpp:217-222) */
    int is_hdr;          /* 0: NV12 / OpticalFlowCalcSDR, 1: P010 / OpticalFlowCalcHDR */
    int device_ordinal;  /* CUDA device; replaces detectDevices (opticalFlowCalc.cpp:45-110) */
    void* cuda_stream;   /* optional cudaStream_t to run on (NULL: the handle creates its own) */
} hrb_ofc_desc;

/* Public data members of OpticalFlowCalc (opticalFlowCalc.h:26-50). Times are in seconds. */
typedef struct hrb_ofc_state {
    int frame_width, frame_height, input_stride, output_stride;
    float output_black_level, output_white_level;
    int res_scalar;              /* m_opticalFlowResScalar */
    int flow_width, flow_height; /* m_opticalFlowFrameWidth / Height */
    int search_radius;           /* m_opticalFlowSearchRadius */
    double ofc_calc_time, ofc_avg_calc_time, ofc_peak_calc_time;
    int ofc_calc_count;
    double ofc_calc_time_sum;
    double warp_calc_time;
    int delta_scalar, neighbor_bias_scalar;
    unsigned int total_frame_delta;
    unsigned int frame_count;
} hrb_ofc_state;

/* The members the filter writes on a live object (HopperRender.cpp:1386-1389, 1448, 1457). */
typedef struct hrb_ofc_params {
    int search_radius;        /* 5..16 (config.h:8-9) */
    int delta_scalar;
    int neighbor_bias_scalar;
    float black_level;
    float white_level;
} hrb_ofc_params;

/* ---- lifecycle --------------------------------------------------------------------------------- */
/* OpticalFlowCalcSDR::OpticalFlowCalcSDR / HDR ctor (opticalFlowCalcSDR.cpp:206-325, HDR :211-332) */
HRB_API int hrb_ofc_create(hrb_ofc** out, const hrb_ofc_desc* desc);
/* ~OpticalFlowCalcSDR / HDR (opticalFlowCalcSDR.cpp:185-204) */
HRB_API void hrb_ofc_destroy(hrb_ofc* h);

/* ---- the five virtuals ------------------------------------------------------------------------- */
/* updateFrame (opticalFlowCalcSDR.cpp:19-29): upload 1.5*H*input_stride elements, rotate the three
 * input slots, ++frame_count.  The host buffer is no longer read once the call returns. */
HRB_API int hrb_ofc_update_frame(hrb_ofc* h, const uint8_t* input_planes);
/* calculateOpticalFlow (opticalFlowCalcSDR.cpp:44-139): search ladder + blur, updates total_frame_delta
 * and the calc-time statistics; blocks until the flow is complete (the reference waits too, :119-120). */
HRB_API int hrb_ofc_calculate_optical_flow(hrb_ofc* h);
/* warpFrames (opticalFlowCalcSDR.cpp:141-168): asynchronous, result in the device output frame. */
HRB_API int hrb_ofc_warp_frames(hrb_ofc* h, float blending_scalar, int frame_output_mode);
/* warpFrames for up to HRB_WARP_BATCH_MAX output frames of the SAME source pair in one pass (beyond the reference, like the
 * asynchronous calls): output frame i is what hrb_ofc_warp_frames(h, blending_scalars[i], mode) would produce, bit for bit.
 * The forward / reverse flow of a sample is fetched once for all of them and both source frames cross HBM once.  The n
 * frames go to consecutive slots of the device output ring; fetch them with n calls of hrb_ofc_download_frame(_async) (in
 * the order given), with no warp_frames / copy_frame call in between. */
#define HRB_WARP_BATCH_MAX 8
HRB_API int hrb_ofc_warp_frames_batch(hrb_ofc* h, int n, const float* blending_scalars, int frame_output_mode);
/* copyFrame (opticalFlowCalcSDR.cpp:170-183) */
HRB_API int hrb_ofc_copy_frame(hrb_ofc* h);
/* downloadFrame (opticalFlowCalcSDR.cpp:31-42): blocking read of 1.5*H*output_stride elements; sets
 * warp_calc_time. */
HRB_API int hrb_ofc_download_frame(hrb_ofc* h, uint8_t* output_planes);

/* ---- public fields ----------------------------------------------------------------------------- */
/* get_state waits for every flow calculation in flight (total_frame_delta and the calc-time statistics are those of the
 * newest one); peek_state never blocks: it reflects the calculations that have finished so far.  A caller that polls
 * fields between calculate_optical_flow_async and the warps uses peek_state so that host and GPU keep overlapping. */
HRB_API int hrb_ofc_get_state(hrb_ofc* h, hrb_ofc_state* out);
HRB_API int hrb_ofc_peek_state(hrb_ofc* h, hrb_ofc_state* out);
HRB_API int hrb_ofc_set_params(hrb_ofc* h, const hrb_ofc_params* p);
/* m_frameCount = n; the filter writes 0 on seek (HopperRender.cpp:840) */
HRB_API int hrb_ofc_set_frame_count(hrb_ofc* h, unsigned int n);
HRB_API int hrb_ofc_reset(hrb_ofc* h); /* == hrb_ofc_set_frame_count(h, 0) */

/* ---- device-resident variants (benchmark / zero-copy callers) ---------------------------------- */
/* Same as hrb_ofc_update_frame but the source is already in device memory of the handle's GPU. */
HRB_API int hrb_ofc_update_frame_device(hrb_ofc* h, const void* device_planes);
/* Device address of the output frame most recently written by warp_frames / copy_frame (for a batch: its first frame;
 * the frames of a batch sit in consecutive slots of a ring that advances with every download). */
HRB_API int hrb_ofc_output_device_ptr(hrb_ofc* h, void** out);
/* Asynchronous variants for PINNED host memory.  Transfers run on their own streams: the upload of the next source
 * frame and the downloads of the current outputs overlap the kernels (the output frame is a ring of three on the
 * device, the upload has its own input slot).
 *   update_frame_async : the buffer must stay untouched until hrb_ofc_wait_upload (or synchronize) returns.
 *   download_frame_async: hands out a ticket; the buffer is valid after hrb_ofc_wait_download(ticket) (tickets may be
 *                         waited for in any order: a wait on an old ticket waits for the newest download that reuses its slot (tickets share 16
 *                         event slots), which proves the old copy has landed — it never returns early). */
HRB_API int hrb_ofc_update_frame_async(hrb_ofc* h, const uint8_t* pinned_input_planes);
HRB_API int hrb_ofc_wait_upload(hrb_ofc* h);
HRB_API int hrb_ofc_download_frame_async(hrb_ofc* h, uint8_t* pinned_output_planes, unsigned long long* ticket);
HRB_API int hrb_ofc_wait_download(hrb_ofc* h, unsigned long long ticket);
/* calculate_optical_flow without the final host wait (statistics are resolved at the next synchronize). */
HRB_API int hrb_ofc_calculate_optical_flow_async(hrb_ofc* h);
HRB_API int hrb_ofc_synchronize(hrb_ofc* h);
/* cudaStream_t the handle launches on */
HRB_API int hrb_ofc_stream(hrb_ofc* h, void** out);
/* Pin / unpin caller-owned frame memory once (e.g. a DirectShow allocator pool) so update/download DMA directly. */
HRB_API int hrb_host_register(void* ptr, size_t bytes);
HRB_API int hrb_host_unregister(void* ptr);
HRB_API int hrb_host_alloc(void** out, size_t bytes);
HRB_API int hrb_host_free(void* ptr);

/* ---- side data of the source frames ----------------------------------------------------------------------------- */
/* The filter reads the IMediaSideData blobs of the input sample (HDR10 / HDR10+ / Dolby Vision metadata and RPU, control
 * flags, content light level, EIA-608 captions, 3D offset: HopperRender.cpp:875-900) and sets them on EVERY output sample it
 * delivers for that source frame (:993-1022).  The library keeps them beside the frame: set_side_data attaches copies of
 * `count` blobs to the frame given to the LAST update_frame call (count 0 clears); get_side_data returns the blobs of that
 * frame for the output frames being delivered.  The pointers stay valid until the next set_side_data or destroy.  The
 * contents are opaque to the library. */
typedef struct hrb_side_data {
    uint8_t guid[16];   /* the interface id the blob was read with (IMediaSideData.h) */
    const void* data;
    size_t bytes;
} hrb_side_data;
HRB_API int hrb_ofc_set_side_data(hrb_ofc* h, const hrb_side_data* items, int count);
HRB_API int hrb_ofc_get_side_data(hrb_ofc* h, hrb_side_data* items, int capacity, int* count);

/* ---- spatial split of one stream over several GPUs ----------------------------------------------------------- */
/* Restrict warp_frames / copy_frame / download_frame to the luma rows [row_begin, row_end) (even bounds) and the
 * chroma rows that belong to them.  Every GPU of the split holds the full source frames and the full flow (the
 * search is replicated, it is not the PCIe-bound part); each produces and downloads only its stripe, into the same
 * offsets of a full-frame host buffer.  Default: the whole frame. */
HRB_API int hrb_ofc_set_output_stripe(hrb_ofc* h, int row_begin, int row_end);

/* ---- test taps --------------------------------------------------------------------------------- */
/* With tap mode on, calculate_optical_flow also records, per (iteration, step) pass: the window sums,
 * the winning layer per window and the offset field after the update.  Off by default (costs memory). */
HRB_API int hrb_ofc_set_tap_mode(hrb_ofc* h, int on);
HRB_API int hrb_ofc_num_passes(hrb_ofc* h, int* out);
HRB_API int hrb_ofc_pass_info(hrb_ofc* h, int pass, int* window_size, int* iteration, int* step, int* windows_x, int* windows_y);
enum {
    HRB_TAP_WINDOW_SUMS = 0, /* uint32 [R][windows_y][windows_x] == summedUpDeltaArray at the window representatives */
    HRB_TAP_WINDOW_LAYER = 1,/* uint8  [windows_y][windows_x]    == lowestLayerArray at the window representatives */
    HRB_TAP_OFFSETS = 2      /* int16  [2][flow_h][flow_w]       == offsetArray after adjustOffsetArrayKernel */
};
HRB_API int hrb_ofc_read_pass_tap(hrb_ofc* h, int pass, int which, void* dst, size_t bytes);
enum {
    HRB_BUF_OFFSET_ARRAY = 0,     /* int16 [2][flow_h][flow_w]: m_offsetArray after the last pass */
    HRB_BUF_FLOW_FOR_WARP = 1,    /* int16 [2][flow_h][flow_w]: m_blurredOffsetArray[0] (what warpFrames reads) */
    HRB_BUF_FLOW_LATEST = 2,      /* int16 [2][flow_h][flow_w]: m_blurredOffsetArray[1] (result of the last calculate) */
    HRB_BUF_OUTPUT_FRAME = 3,     /* m_outputFrameArray */
    HRB_BUF_RAW_FRAME_DELTA = 4,  /* uint32: the raw window sum total_frame_delta is derived from */
    HRB_BUF_FLOW_PEAK = 5         /* uint32 [2]: max |value| of the flow warpFrames reads, and of the latest one (the bound the warp kernel uses to skip the mirror) */
};
HRB_API int hrb_ofc_read_buffer(hrb_ofc* h, int which, void* dst, size_t bytes);
/* overwrite a blurred flow (HRB_BUF_FLOW_FOR_WARP / HRB_BUF_FLOW_LATEST) — lets tests drive warp_frames */
HRB_API int hrb_ofc_write_flow(hrb_ofc* h, int which, const int16_t* src, size_t count);

/* ---- measurement support ----------------------------------------------------------------------- */
typedef struct hrb_ofc_profile {
    /* accumulated since the last hrb_ofc_profile_reset, GPU time from CUDA events on the handle's stream */
    double ms_ingest, ms_search, ms_blur, ms_warp, ms_copy;
    uint64_t n_ingest, n_search, n_blur, n_warp, n_copy; /* kernel launches per class */
} hrb_ofc_profile;
HRB_API int hrb_ofc_set_profile(hrb_ofc* h, int on);
HRB_API int hrb_ofc_profile_read(hrb_ofc* h, hrb_ofc_profile* out); /* synchronizes */
HRB_API int hrb_ofc_profile_reset(hrb_ofc* h);
/* Kernel selection for the search ladder: 0 = automatic (tile kernel for windows >= 16, staged small-window kernel
 * below), 1 = the generic kernel for every pass, 2 = tile kernel without TMA (cp.async staging only) and down to
 * windows of 4, 3 = like 2 with the per-pixel fallback forced for every window, 4 = like 0 but windows of 2 and 4 use the
 * lane-per-pixel-column form with a butterfly reduction instead of one lane per window.
 * Results are identical; exists for A/B measurements and parity tests. */
HRB_API int hrb_ofc_set_search_variant(hrb_ofc* h, int variant);
/* hrb_ofc_calculate_optical_flow_async runs the search on its own stream, beside the warps issued after it (they read
 * the PREVIOUS flow, opticalFlowCalcSDR.cpp:113-123,150).  on = 0 keeps everything on the compute stream.  Default 1. */
HRB_API int hrb_ofc_set_flow_overlap(hrb_ofc* h, int on);
/* Orders the compute stream (hrb_ofc_stream) behind the flow calculation still running on the flow stream, without
 * blocking the host: work a caller enqueues on the compute stream afterwards sees the finished flow. */
HRB_API int hrb_ofc_join_flow(hrb_ofc* h);
/* Debug aid: record a per-CTA timeline of every tile search kernel of the following flow calculations (8 uint64 per
 * CTA: start, boxes landed, runs done, end [globaltimer ns], SM id, block x, block y, rounds; `words_per_pass` words for
 * each of up to 32 passes; 0 switches it off).  Used by tools/cta_timeline.py. */
HRB_API int hrb_ofc_debug_timeline(hrb_ofc* h, size_t words_per_pass);
HRB_API int hrb_ofc_debug_timeline_read(hrb_ofc* h, void* dst, size_t bytes);
/* number of kernels this library has launched in this process */
HRB_API uint64_t hrb_kernel_launch_count(void);
/* packed byte-SAD instruction peak of the device (VABSDIFF4.U8.ACC issue rate), in 1e9 byte-abs-diffs/s */
HRB_API int hrb_microbench_sad_peak(int device_ordinal, double* giga_absdiff_per_s);

HRB_API const char* hrb_last_error(void);
HRB_API const char* hrb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HRB_H_ */
