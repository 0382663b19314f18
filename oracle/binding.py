"""ctypes binding of oracle/libhr_oracle.so (the CPU restatement of the reference path)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libhr_oracle.so")


class _State(C.Structure):
    _fields_ = [
        ("frameWidth", C.c_int), ("frameHeight", C.c_int), ("inputStride", C.c_int), ("outputStride", C.c_int),
        ("outputBlackLevel", C.c_float), ("outputWhiteLevel", C.c_float),
        ("resScalar", C.c_int), ("flowWidth", C.c_int), ("flowHeight", C.c_int), ("searchRadius", C.c_int),
        ("ofcCalcTime", C.c_double), ("ofcAvgCalcTime", C.c_double), ("ofcPeakCalcTime", C.c_double), ("warpCalcTime", C.c_double),
        ("deltaScalar", C.c_int), ("neighborBiasScalar", C.c_int), ("totalFrameDelta", C.c_uint), ("frameCount", C.c_uint),
    ]


_lib = None
_ref_lib = None


class _Api:
    """The calculator entry points of one library under a common name (oracle: orc_*, oracle/_ref: hrref_*)."""

    def __init__(self, lib, prefix):
        P = C.c_void_p
        sig = {
            "ofc_create": (P, [C.c_int] * 6 + [C.c_float, C.c_float, C.c_int, C.c_int]),
            "ofc_destroy": (None, [P]),
            "ofc_update_frame": (C.c_int, [P, P]),
            "ofc_download_frame": (C.c_int, [P, P]),
            "ofc_calculate_optical_flow": (C.c_int, [P]),
            "ofc_warp_frames": (C.c_int, [P, C.c_float, C.c_int]),
            "ofc_copy_frame": (C.c_int, [P]),
            "ofc_get_state": (None, [P, C.POINTER(_State)]),
            "ofc_set_params": (None, [P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]),
            "ofc_set_frame_count": (None, [P, C.c_uint]),
            "ofc_enable_taps": (None, [P, C.c_int]),
            "ofc_num_passes": (C.c_int, [P]),
            "ofc_pass_info": (C.c_int, [P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
            "ofc_read_pass_tap": (C.c_int, [P, C.c_int, C.c_int, P, C.c_size_t]),
            "ofc_read_buffer": (C.c_int, [P, C.c_int, P, C.c_size_t]),
            "ofc_write_flow": (C.c_int, [P, C.c_int, P, C.c_size_t]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, prefix + name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(lib_path()):
        subprocess.check_call(["make", "-C", _HERE, "libhr_oracle.so"])
    lib = C.CDLL(lib_path())
    P = C.c_void_p
    lib.api = _Api(lib, "orc_")
    lib.orc_calc_delta_sums.argtypes = [P, P, P, P] + [C.c_int] * 13
    lib.orc_determine_lowest_layer.argtypes = [P, P] + [C.c_int] * 4
    lib.orc_adjust_offset_array.argtypes = [P, P] + [C.c_int] * 5
    lib.orc_blur_flow.argtypes = [P, P, C.c_int, C.c_int]
    lib.orc_warp_frame.argtypes = [P, P, P, P, C.c_float, C.c_float] + [C.c_int] * 8 + [C.c_float, C.c_float, C.c_int, C.c_int]
    lib.orc_copy_frame_kernel.argtypes = [P, P] + [C.c_int] * 4 + [C.c_float, C.c_float, C.c_int, C.c_int]
    lib.orc_num_threads.restype = C.c_int
    _lib = lib
    return lib


def num_threads():
    return int(_load().orc_num_threads())


def set_num_threads(n):
    """Set the OpenMP thread count of the oracle explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    lib = _load()
    lib.orc_set_num_threads.restype = C.c_int
    lib.orc_set_num_threads.argtypes = [C.c_int]
    return int(lib.orc_set_num_threads(int(n)))


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


class kernels:
    """The six reference kernels as stateless functions on numpy arrays."""

    @staticmethod
    def calc_delta_sums(frame1, frame2, offsets, dimY, dimX, stride, ws, R, rs, iteration, step, deltaScalar, nbScalar, hdr):
        _, lh, lw = offsets.shape
        sums = np.zeros((R, lh, lw), np.uint32)
        _load().orc_calc_delta_sums(_p(sums), _p(frame1), _p(frame2), _p(offsets), dimY, dimX, stride, lh, lw, ws, R, rs, iteration, step,
                                    deltaScalar, nbScalar, int(hdr))
        return sums

    @staticmethod
    def determine_lowest_layer(sums, layers, ws):
        R, lh, lw = sums.shape
        _load().orc_determine_lowest_layer(_p(sums), _p(layers), ws, R, lh, lw)
        return layers

    @staticmethod
    def adjust_offset_array(offsets, layers, ws, R, step):
        _, lh, lw = offsets.shape
        _load().orc_adjust_offset_array(_p(offsets), _p(layers), ws, R, lh, lw, step)
        return offsets

    @staticmethod
    def blur_flow(offsets):
        _, lh, lw = offsets.shape
        out = np.empty_like(offsets)
        _load().orc_blur_flow(_p(offsets), _p(out), lh, lw)
        return out

    @staticmethod
    def warp_frame(src12, src21, flow, out, t12, t21, H, W, S, So, rs, mode, black, white, cz, hdr):
        _, lh, lw = flow.shape
        _load().orc_warp_frame(_p(src12), _p(src21), _p(flow), _p(out), t12, t21, lh, lw, H, W, S, So, rs, mode, black, white, cz, int(hdr))
        return out

    @staticmethod
    def copy_frame(src, out, H, W, S, So, black, white, cz, hdr):
        _load().orc_copy_frame_kernel(_p(src), _p(out), H, W, S, So, black, white, cz, int(hdr))
        return out


class OracleCalc:
    """Same surface as hopperrender_b200.OpticalFlowCalcSDR/HDR, computed on the CPU by the oracle."""

    def _api(self):
        return _load().api

    def __init__(self, frameHeight, frameWidth, inputStride, outputStride, deltaScalar, neighborScalar, blackLevel, whiteLevel,
                 maxCalcRes, hdr):
        self._lib = self._api()
        self.hdr = bool(hdr)
        self._h = C.c_void_p(self._lib.ofc_create(frameHeight, frameWidth, inputStride, outputStride, deltaScalar, neighborScalar,
                                                  blackLevel, whiteLevel, maxCalcRes, int(hdr)))
        if not self._h:
            raise RuntimeError("calculator construction failed: " + self._last_error())
        s = self.state()
        bpp = 2 if hdr else 1
        self.inputFrameBytes = (s.frameHeight * s.inputStride + (s.frameHeight // 2) * s.inputStride) * bpp
        self.outputFrameBytes = (s.frameHeight * s.outputStride + (s.frameHeight // 2) * s.outputStride) * bpp

    def _last_error(self):
        return ""

    def close(self):
        if self._h:
            self._lib.ofc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def state(self):
        s = _State()
        self._lib.ofc_get_state(self._h, C.byref(s))
        return s

    def _chk(self, rc, what):
        if rc:
            raise RuntimeError(f"{what} failed: {self._last_error()}")

    def updateFrame(self, a):
        assert a.nbytes >= self.inputFrameBytes
        self._chk(self._lib.ofc_update_frame(self._h, _p(a)), "updateFrame")

    def downloadFrame(self, a):
        assert a.nbytes >= self.outputFrameBytes
        self._chk(self._lib.ofc_download_frame(self._h, _p(a)), "downloadFrame")

    def calculateOpticalFlow(self):
        self._chk(self._lib.ofc_calculate_optical_flow(self._h), "calculateOpticalFlow")

    def warpFrames(self, t, mode):
        if self._lib.ofc_warp_frames(self._h, float(t), int(mode)):
            raise RuntimeError("[HopperRender] Error in function warpFrames " + self._last_error())

    def copyFrame(self):
        self._chk(self._lib.ofc_copy_frame(self._h), "copyFrame")

    def setParams(self, searchRadius=None, deltaScalar=None, neighborBiasScalar=None, black=None, white=None):
        s = self.state()
        self._lib.ofc_set_params(self._h, s.searchRadius if searchRadius is None else searchRadius,
                                     s.deltaScalar if deltaScalar is None else deltaScalar,
                                     s.neighborBiasScalar if neighborBiasScalar is None else neighborBiasScalar,
                                     s.outputBlackLevel if black is None else black, s.outputWhiteLevel if white is None else white)

    def setFrameCount(self, n):
        self._lib.ofc_set_frame_count(self._h, n)

    def enableTaps(self, on=True):
        self._lib.ofc_enable_taps(self._h, int(on))

    def numPasses(self):
        return self._lib.ofc_num_passes(self._h)

    def passInfo(self, p):
        v = [C.c_int() for _ in range(3)]
        assert self._lib.ofc_pass_info(self._h, p, *[C.byref(x) for x in v]) == 0
        return dict(zip(("windowSize", "iteration", "step"), (x.value for x in v)))

    def _shape(self):
        s = self.state()
        return s.flowHeight, s.flowWidth

    def readPassSums(self, p, R):
        lh, lw = self._shape()
        a = np.empty((R, lh, lw), np.uint32)
        assert self._lib.ofc_read_pass_tap(self._h, p, 0, _p(a), a.nbytes) == 0
        return a

    def readPassLayers(self, p):
        lh, lw = self._shape()
        a = np.empty((lh, lw), np.uint8)
        assert self._lib.ofc_read_pass_tap(self._h, p, 1, _p(a), a.nbytes) == 0
        return a

    def readPassOffsets(self, p):
        lh, lw = self._shape()
        a = np.empty((2, lh, lw), np.int16)
        assert self._lib.ofc_read_pass_tap(self._h, p, 2, _p(a), a.nbytes) == 0
        return a

    def _readFlow(self, which):
        lh, lw = self._shape()
        a = np.empty((2, lh, lw), np.int16)
        assert self._lib.ofc_read_buffer(self._h, which, _p(a), a.nbytes) == 0
        return a

    def readOffsetArray(self):
        return self._readFlow(0)

    def readFlow(self, latest=False):
        return self._readFlow(2 if latest else 1)

    def writeFlow(self, flow, latest=False):
        flow = np.ascontiguousarray(flow, np.int16)
        assert self._lib.ofc_write_flow(self._h, 2 if latest else 1, _p(flow), flow.size) == 0


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref: the UNMODIFIED reference classes + kernel strings on a real OpenCL device (see oracle/ref_build)
# ---------------------------------------------------------------------------------------------------------
def ref_lib_path():
    return os.path.join(_HERE, "_ref", "libhrref.so")


def _load_ref():
    global _ref_lib
    if _ref_lib is not None:
        return _ref_lib
    if not os.path.exists(ref_lib_path()):
        raise FileNotFoundError(f"{ref_lib_path()} not built (make -C oracle/ref_build, needs /root/reference)")
    lib = C.CDLL(ref_lib_path())
    lib.api = _Api(lib, "hrref_")
    lib.hrref_last_error.restype = C.c_char_p
    lib.hrref_opencl_library.restype = C.c_char_p
    lib.hrref_device_name.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    _ref_lib = lib
    return lib


class RefCalc(OracleCalc):
    """OpticalFlowCalcSDR / OpticalFlowCalcHDR of the reference itself, run through OpenCL."""

    def _api(self):
        return _load_ref().api

    def _last_error(self):
        return _load_ref().hrref_last_error().decode("utf-8", "replace")

    def deviceName(self):
        buf = C.create_string_buffer(256)
        _load_ref().hrref_device_name(self._h, buf, 256)
        return buf.value.decode()

    def openclLibrary(self):
        return _load_ref().hrref_opencl_library().decode()


def ref_available():
    """True when oracle/_ref is built and an OpenCL device accepts the reference."""
    try:
        c = RefCalc(32, 32, 0, 0, 8, 6, 0.0, 255.0, 270, False)
        c.close()
        return True
    except Exception:
        return False
