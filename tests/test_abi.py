"""The C-ABI library loads on a CPU-only machine and exports every symbol include/hrb.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hrb.h")).read()
    return sorted(set(re.findall(r"HRB_API\s+[\w\s\*]+?\b(hrb_\w+)\s*\(", src)))


def test_header_declares_the_reference_surface():
    names = declared_symbols()
    # the five virtuals + ctor/dtor of HopperRender/opticalFlowCalc.h:100-132
    for n in ["hrb_ofc_create", "hrb_ofc_destroy", "hrb_ofc_update_frame", "hrb_ofc_download_frame", "hrb_ofc_calculate_optical_flow",
              "hrb_ofc_warp_frames", "hrb_ofc_copy_frame", "hrb_ofc_get_state", "hrb_ofc_set_params", "hrb_last_error"]:
        assert n in names


def test_library_exports_every_declared_symbol():
    from hopperrender_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in declared_symbols():
        assert hasattr(lib, n), f"libhrb.so does not export {n}"
    assert set(declared_symbols()) == set(_lib.PROTOTYPES), "python prototypes and hrb.h differ"


def test_no_cpu_fallback_without_device():
    """Without a GPU the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import hopperrender_b200 as hr
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA error"):
        hr.OpticalFlowCalcSDR(64, 64, 0, 0, 8, 6, 0.0, 255.0, 270)


def test_product_does_not_reference_the_oracle():
    """Nothing under hopperrender_b200/ or include/ may import, link or call oracle/."""
    bad = []
    for base in ("hopperrender_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    if re.search(r"\boracle\b", txt) and not fn == "_lib.py":
                        bad.append(os.path.join(dp, fn))
                    if fn == "_lib.py" and re.search(r"import\s+oracle|from\s+oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_cpp_shim_compiles_and_links_against_the_library(tmp_path):
    """include/opticalFlowCalc*.h are the classes a HopperRender build would include instead of the reference's: the
    delivery-loop driver written against the reference's class surface must compile and link with nothing else."""
    import subprocess
    exe = tmp_path / "replay_check.bin"
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "replay.cpp"),
           "-o", str(exe), "-L" + os.path.join(ROOT, "hopperrender_b200"), "-lhrb", "-Wl,-rpath," + os.path.join(ROOT, "hopperrender_b200")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # no device here: the program must fail loudly (exception text from hrb_last_error), not fall back to anything
    raw = tmp_path / "f.raw"
    raw.write_bytes(bytes(64 * 48 * 3 // 2 * 3))
    run = subprocess.run([str(exe), str(raw), "64", "48", "0", "3", "166667", "270", "8"], capture_output=True, text=True)
    import torch
    if not torch.cuda.is_available():
        assert run.returncode != 0


@pytest.mark.parametrize("kw,needle", [
    (dict(frameHeight=0, frameWidth=0), "at least 16x16"),            # empty frame
    (dict(frameHeight=8, frameWidth=64), "at least 16x16"),
    (dict(frameHeight=48, frameWidth=63), "must be even"),            # NV12/P010 need even dimensions
    (dict(frameHeight=49, frameWidth=64), "must be even"),
    (dict(inputStride=32), "stride smaller"),                         # ragged: stride below the width
    (dict(outputStride=63), "stride smaller"),
    (dict(maxCalcRes=0), "max_calc_res"),
    (dict(frameHeight=4096, frameWidth=16, maxCalcRes=16), "below 4x4"),
    (dict(deltaScalar=32), "scalar outside"),
    (dict(neighborScalar=-1), "scalar outside"),
])
def test_create_rejects_bad_geometry_before_touching_the_device(kw, needle):
    """Argument errors are reported as such (HRB_ERR_INVALID_ARG with a message), with or without a GPU."""
    import hopperrender_b200 as hr
    args = dict(frameHeight=48, frameWidth=64, inputStride=0, outputStride=0, deltaScalar=8, neighborScalar=6, blackLevel=0.0,
                whiteLevel=255.0, maxCalcRes=270)
    args.update(kw)
    with pytest.raises(RuntimeError) as e:
        hr.OpticalFlowCalcSDR(args["frameHeight"], args["frameWidth"], args["inputStride"], args["outputStride"], args["deltaScalar"],
                              args["neighborScalar"], args["blackLevel"], args["whiteLevel"], args["maxCalcRes"])
    assert "hrb error 1" in str(e.value) and needle in str(e.value)


def test_null_handles_and_pointers_are_argument_errors():
    from hopperrender_b200 import _lib
    lib = _lib.load()
    assert lib.hrb_ofc_update_frame(None, None) == 1
    assert lib.hrb_ofc_calculate_optical_flow(None) == 1
    assert lib.hrb_ofc_warp_frames(None, 0.5, 2) == 1
    assert lib.hrb_ofc_download_frame(None, None) == 1
    assert lib.hrb_ofc_create(None, None) == 1
    assert b"null" in lib.hrb_last_error()


def test_header_is_plain_c_and_the_c_example_builds(tmp_path):
    """include/hrb.h is the FFI surface: it must compile as pedantic C99 (no C++-isms, no torch/CUDA types), and the C
    example of the call sequence must link against the library alone."""
    import subprocess
    exe = tmp_path / "example_c"
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tools", "example_c_api.c"), "-o", str(exe), "-L" + os.path.join(ROOT, "hopperrender_b200"), "-lhrb",
           "-Wl,-rpath," + os.path.join(ROOT, "hopperrender_b200")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and "flow 320x180" in run.stdout, run.stdout + run.stderr
    else:
        assert run.returncode == 2 and "no CUDA device" in run.stderr
