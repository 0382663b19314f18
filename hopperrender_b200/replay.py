"""Headless replay of the filter's delivery loop (CHopperRender::DeliverToRenderer,
HopperRender/HopperRender.cpp:944-1211) — the caller that turns the calculator into interpolated frames.

It reproduces, per source frame: the number of output frames (:944-948), the search-radius auto-tuner
(:951, :1438-1463), updateFrame / calculateOpticalFlow (:953-957) with the frame-delta history (:959-973),
the scene-change rule (:1126-1176), warpFrames-or-copyFrame (:1179-1183), downloadFrame (:1186) and the
blending-scalar accumulation (:1192-1197); the settings / metrics surface of SettingsInterface (iez.h:14-51:
GetCurrentSettings :1243-1350, UpdateUserSettings :1355-1390), the interpolation status (:819-831) and NewSegment
(:834-844).  DirectShow plumbing (samples, timestamps, media-type renegotiation) is out of scope;
`sink(frame_bytes, info)` stands in for m_pOutput->Deliver, and per-frame side data rides through as an opaque object.
"""
import math
from collections import deque

# HopperRender/config.h:8-9,14-15,28
MIN_SEARCH_RADIUS = 5
MAX_SEARCH_RADIUS = 16
UPPER_PERF_BUFFER = 1.4
LOWER_PERF_BUFFER = 1.6
DEFAULT_SCENE_CHANGE_THRESHOLD = 200

# ActiveState — HopperRender.h:20-25
DEACTIVATED, NOT_NEEDED, ACTIVE, TOO_SLOW = 0, 1, 2, 3

# 100-ns units used by the filter; 23.976 fps source, 144 / 60 Hz targets
SOURCE_FRAME_TIME_23976 = 417083
TARGET_FRAME_TIME_144 = 69444
TARGET_FRAME_TIME_60 = 166667


def num_int_frames(blending_scalar, target_frame_time, playback_frame_time):
    """m_iNumIntFrames — HopperRender.cpp:945."""
    return int(max(math.ceil((1.0 - blending_scalar) / (float(target_frame_time) / float(playback_frame_time))), 1.0))


def advance_blend(blending_scalar, target_frame_time, playback_frame_time):
    """HopperRender.cpp:1192-1197."""
    blending_scalar += float(target_frame_time) / float(playback_frame_time)
    if blending_scalar >= 1.0:
        blending_scalar -= 1.0
    return blending_scalar


def output_schedule(n_source_frames, target_frame_time=TARGET_FRAME_TIME_144, playback_frame_time=SOURCE_FRAME_TIME_23976,
                    blending_scalar=0.0):
    """For each source frame the list of blending scalars of its output frames (interpolation Active)."""
    sched = []
    b = blending_scalar
    for _ in range(n_source_frames):
        n = num_int_frames(b, target_frame_time, playback_frame_time)
        row = []
        for _ in range(n):
            row.append(b)
            b = advance_blend(b, target_frame_time, playback_frame_time)
        sched.append(row)
    return sched


class DeliveryLoop:
    """State of CHopperRender that drives the calculator (m_dBlendingScalar, histories, auto-tuner)."""

    def __init__(self, calc, source_frame_time=SOURCE_FRAME_TIME_23976, target_frame_time=TARGET_FRAME_TIME_144, frame_output=2,
                 scene_change_threshold=DEFAULT_SCENE_CHANGE_THRESHOLD, active=True, auto_adjust=True, buffer_frames=0):
        self.calc = calc
        self.rtSourceFrameTime = source_frame_time
        self.rtCurrPlaybackFrameTime = source_frame_time
        self.rtTargetFrameTime = target_frame_time
        self.iFrameOutput = frame_output
        self.iSceneChangeThreshold = scene_change_threshold
        self.iIntActiveState = ACTIVE if active else DEACTIVATED
        self.bUseDisplayFPS = False
        self.iBufferFrames = buffer_frames
        self.auto_adjust = auto_adjust
        self.dBlendingScalar = 0.0
        self.dTotalWarpDuration = 0.0
        self.frameDeltaHistory = deque()          # (frameNumber, totalDelta)
        self.sceneChangeDeltaHistory = deque()    # (frameNumber, delta1, delta2)
        self.iPeakSceneChangeDelta = 0
        self.iPeakSceneChangeDelta2 = 0
        self.log = []
        # True: the output frames of a source frame come from ONE warpFramesBatch pass (hrb_ofc_warp_frames_batch) instead of one
        # warpFrames call each; the delivered frames are the same.  Needs a calculator that has the batched entry point.
        self.batch = False

    @property
    def active(self):
        """m_iIntActiveState == Active, the condition DeliverToRenderer tests."""
        return self.iIntActiveState == ACTIVE

    def update_interpolation_status(self):
        """CHopperRender::UpdateInterpolationStatus — HopperRender.cpp:819-831."""
        if self.iIntActiveState and self.rtCurrPlaybackFrameTime > self.rtTargetFrameTime:
            self.iIntActiveState = ACTIVE
        elif self.iIntActiveState:
            self.iIntActiveState = NOT_NEEDED
        self.iPeakSceneChangeDelta = 0
        self.iPeakSceneChangeDelta2 = 0
        self.frameDeltaHistory.clear()
        self.sceneChangeDeltaHistory.clear()

    def new_segment(self, rate=1.0):
        """CHopperRender::NewSegment (seek or playback-speed change) — HopperRender.cpp:834-844."""
        self.rtCurrPlaybackFrameTime = int(float(self.rtSourceFrameTime) * (1.0 / rate))
        self.update_interpolation_status()
        self.calc.m_frameCount = 0

    def get_current_settings(self):
        """SettingsInterface::GetCurrentSettings — the 23 values of iez.h:14-37 as HopperRender.cpp:1325-1348 fills them
        once a calculator exists (times in milliseconds, frame rates from the 100-ns frame times)."""
        c = self.calc
        return {
            "bActivated": self.iIntActiveState != 0,
            "iFrameOutput": self.iFrameOutput,
            "dTargetFPS": 10000000.0 / float(self.rtTargetFrameTime),
            "bUseDisplayFPS": self.bUseDisplayFPS,
            "iDeltaScalar": int(c.m_deltaScalar),
            "iNeighborScalar": int(c.m_neighborBiasScalar),
            "iBlackLevel": int(c.m_outputBlackLevel),
            "iWhiteLevel": int(c.m_outputWhiteLevel),
            "iSceneChangeThreshold": self.iSceneChangeThreshold,
            "iIntActiveState": self.iIntActiveState,
            "dSourceFPS": 10000000.0 / float(self.rtCurrPlaybackFrameTime),
            "dOFCCalcTime": 1000.0 * c.m_ofcCalcTime,
            "dAVGOFCCalcTime": 1000.0 * c.m_ofcAvgCalcTime,
            "dPeakOFCCalcTime": 1000.0 * c.m_ofcPeakCalcTime,
            "dWarpCalcTime": 1000.0 * self.dTotalWarpDuration,
            "iDimX": int(c.m_frameWidth),
            "iDimY": int(c.m_frameHeight),
            "iLowDimX": int(c.m_opticalFlowFrameWidth),
            "iLowDimY": int(c.m_opticalFlowFrameHeight),
            "iTotalFrameDelta": self.iPeakSceneChangeDelta,
            "iTotalFrameDelta2": self.iPeakSceneChangeDelta2,
            "iBufferFrames": self.iBufferFrames,
            "iSearchRadius": int(c.m_opticalFlowSearchRadius),
        }

    def update_user_settings(self, bActivated, iFrameOutput, dTargetFPS, bUseDisplayFPS, iDeltaScalar, iNeighborScalar, iBlackLevel,
                             iWhiteLevel, iSceneChangeThreshold, iBufferFrames):
        """SettingsInterface::UpdateUserSettings — HopperRender.cpp:1355-1390.  The scalars and levels take effect at
        the calculator's next call, as in the reference; there is no display to query, so bUseDisplayFPS keeps the
        current target frame time (useDisplayRefreshRate, :1379)."""
        if not bActivated:
            self.iIntActiveState = DEACTIVATED
        elif not self.iIntActiveState:
            self.iIntActiveState = ACTIVE
        self.iFrameOutput = int(iFrameOutput)
        if dTargetFPS > 0.0 and not bUseDisplayFPS:
            self.rtTargetFrameTime = int((1.0 / float(dTargetFPS)) * 1e7)
        self.bUseDisplayFPS = bool(bUseDisplayFPS)
        self.iSceneChangeThreshold = int(iSceneChangeThreshold)
        self.iBufferFrames = int(iBufferFrames)
        self.update_interpolation_status()
        c = self.calc
        c.m_deltaScalar = int(iDeltaScalar)
        c.m_neighborBiasScalar = int(iNeighborScalar)
        c.m_outputBlackLevel = float(iBlackLevel)
        c.m_outputWhiteLevel = float(iWhiteLevel)

    def auto_adjust_settings(self):
        """CHopperRender::autoAdjustSettings — HopperRender.cpp:1438-1463."""
        source_frame_time_s = float(self.rtCurrPlaybackFrameTime) / 10000000.0
        curr_max = self.calc.m_ofcCalcTime + self.dTotalWarpDuration
        r = self.calc.m_opticalFlowSearchRadius
        if curr_max * UPPER_PERF_BUFFER > source_frame_time_s:
            if r > MIN_SEARCH_RADIUS:
                self.calc.m_opticalFlowSearchRadius = r - 1
        elif curr_max * LOWER_PERF_BUFFER < source_frame_time_s:
            if r < MAX_SEARCH_RADIUS:
                self.calc.m_opticalFlowSearchRadius = r + 1
        self.dTotalWarpDuration = 0.0

    def _scene_change(self):
        """HopperRender.cpp:1126-1176.  Returns sceneChangeDetected."""
        h = self.frameDeltaHistory
        if len(h) < 3:
            return False
        n = len(h)
        count = min(n - 2, 10)
        total = sum(h[n - 2 - i][1] for i in range(count))
        average = int(total // count)
        nxt = int(h[n - 1][1])
        cur = int(h[n - 2][1])
        d1 = cur - average
        d2 = cur - nxt
        if d1 > 0:
            frames_in_1s = int(1.0 * 10000000.0 / self.rtSourceFrameTime)
            fc = self.calc.m_frameCount
            self.sceneChangeDeltaHistory.append((fc, d1, d2 if d2 > 0 else 0))
            while self.sceneChangeDeltaHistory and (fc - self.sceneChangeDeltaHistory[0][0]) > frames_in_1s:
                self.sceneChangeDeltaHistory.popleft()
            self.iPeakSceneChangeDelta = 0
            self.iPeakSceneChangeDelta2 = 0
            for _, a, b in self.sceneChangeDeltaHistory:
                if a > self.iPeakSceneChangeDelta:
                    self.iPeakSceneChangeDelta = a
                    self.iPeakSceneChangeDelta2 = b
        return d1 >= self.iSceneChangeThreshold and d1 > 0 and d2 >= self.iSceneChangeThreshold and d2 > 0

    def deliver(self, in_buffer, out_buffer, sink=None, side_data=None):
        """One source frame through DeliverToRenderer.  Returns the number of frames delivered.  `side_data` (the
        opaque IMediaSideData blobs of the input sample, read at HopperRender.cpp:875-900 and set on every output sample at
        :997-1018) is handed to every output frame unchanged."""
        c = self.calc
        n_int = num_int_frames(self.dBlendingScalar, self.rtTargetFrameTime, self.rtCurrPlaybackFrameTime) if self.active else 1
        if self.auto_adjust:
            self.auto_adjust_settings()
        c.updateFrame(in_buffer)
        if self.active and c.m_frameCount >= 3:
            c.calculateOpticalFlow()
            frames_in_3s = int(3.0 * 10000000.0 / self.rtSourceFrameTime)
            fc = c.m_frameCount
            self.frameDeltaHistory.append((fc, c.m_totalFrameDelta))
            while self.frameDeltaHistory and (fc - self.frameDeltaHistory[0][0]) > frames_in_3s:
                self.frameDeltaHistory.popleft()
        batched = False
        for i in range(n_int):
            scene_change = self._scene_change()
            warped = self.active and c.m_frameCount >= 3 and not scene_change
            if warped and self.batch and i == 0 and n_int <= 8:
                # the scene-change decision depends on the delta history only: it is the same for every output frame of this
                # source frame, so all of them are warped, at the blend scalars the loop below steps through
                blends, b = [], self.dBlendingScalar
                for _ in range(n_int):
                    blends.append(b)
                    b = advance_blend(b, self.rtTargetFrameTime, self.rtCurrPlaybackFrameTime)
                c.warpFramesBatch(blends, self.iFrameOutput)
                batched = True
            if warped and batched:
                pass  # frame i of the batch is the next one downloadFrame fetches
            elif warped:
                c.warpFrames(self.dBlendingScalar, self.iFrameOutput)
            else:
                c.copyFrame()
            c.downloadFrame(out_buffer)
            self.dTotalWarpDuration += c.m_warpCalcTime
            info = {"source": c.m_frameCount, "index": i, "blend": self.dBlendingScalar, "warped": warped, "scene_change": scene_change,
                    "radius": c.m_opticalFlowSearchRadius, "side_data": side_data}
            self.log.append(info)
            if self.active:
                self.dBlendingScalar = advance_blend(self.dBlendingScalar, self.rtTargetFrameTime, self.rtCurrPlaybackFrameTime)
            if sink is not None:
                sink(out_buffer, info)
        return n_int
