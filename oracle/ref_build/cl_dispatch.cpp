// cl_dispatch.cpp — binds the 24 OpenCL entry points the reference uses to a real OpenCL library at run time.
// Test infrastructure (oracle/_ref).  The library is chosen by $HRREF_OPENCL_LIB, else libOpenCL.so.1 (ICD
// loader), else libnvidia-opencl.so.1 (the vendor library, which also exports the entry points).
// clSetKernelArg / clEnqueueNDRangeKernel additionally report to an observer so that ref_harness.cpp can tap
// the reference's buffers after every search pass without touching the reference's sources.
#include <CL/cl.h>
#include <dlfcn.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

static void* g_lib = nullptr;
static std::string g_libName;

extern "C" const char* hrref_opencl_library() { return g_libName.c_str(); }

static void* resolve(const char* name) {
    if (!g_lib) {
        const char* env = getenv("HRREF_OPENCL_LIB");
        // a box with the NVIDIA driver but no /etc/OpenCL/vendors/*.icd: tell the ICD loader which vendor library to load
        if (!getenv("OCL_ICD_FILENAMES") && !getenv("OCL_ICD_VENDORS") && access("/etc/OpenCL/vendors/nvidia.icd", R_OK) != 0)
            setenv("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1", 1);
        const char* candidates[] = {env, "libOpenCL.so.1", "libnvidia-opencl.so.1", "libOpenCL.so"};
        for (const char* c : candidates) {
            if (!c || !*c) continue;
            void* lib = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (!lib) continue;
            // the ICD loader loads fine without any registered vendor (no /etc/OpenCL/vendors/*.icd) but then reports
            // no platform: skip it in that case and bind the vendor library directly
            typedef cl_int (*gp_t)(cl_uint, cl_platform_id*, cl_uint*);
            gp_t gp = (gp_t)dlsym(lib, "clGetPlatformIDs");
            cl_uint n = 0;
            if (!env && gp && (gp(0, nullptr, &n) != CL_SUCCESS || n == 0)) {
                dlclose(lib);
                continue;
            }
            g_lib = lib;
            g_libName = c;
            break;
        }
        if (!g_lib) {
            fprintf(stderr, "[hrref] no OpenCL library could be loaded: %s\n", dlerror());
            return nullptr;
        }
    }
    return dlsym(g_lib, name);
}

// observer hooks (set by ref_harness.cpp)
extern "C" {
void (*hrref_on_set_kernel_arg)(cl_kernel, cl_uint, size_t, const void*) = nullptr;
void (*hrref_on_enqueue_kernel)(cl_command_queue, cl_kernel) = nullptr;
}

#define FWD(ret, name, params, args, fail)                         \
    extern "C" ret name params {                                   \
        typedef ret(*fn_t) params;                                 \
        static fn_t fn = (fn_t)resolve(#name);                     \
        if (!fn) return fail;                                      \
        return fn args;                                            \
    }

#define ERR (-1001) /* CL_PLATFORM_NOT_FOUND_KHR */

FWD(cl_int, clGetPlatformIDs, (cl_uint a, cl_platform_id* b, cl_uint* c), (a, b, c), ERR)
FWD(cl_int, clGetDeviceIDs, (cl_platform_id a, cl_device_type b, cl_uint c, cl_device_id* d, cl_uint* e), (a, b, c, d, e), ERR)
FWD(cl_int, clGetDeviceInfo, (cl_device_id a, cl_device_info b, size_t c, void* d, size_t* e), (a, b, c, d, e), ERR)
FWD(cl_context, clCreateContext,
    (const cl_context_properties* a, cl_uint b, const cl_device_id* c, void (*d)(const char*, const void*, size_t, void*), void* e, cl_int* f),
    (a, b, c, d, e, f), nullptr)
FWD(cl_command_queue, clCreateCommandQueueWithProperties, (cl_context a, cl_device_id b, const cl_queue_properties* c, cl_int* d), (a, b, c, d), nullptr)
FWD(cl_mem, clCreateBuffer, (cl_context a, cl_mem_flags b, size_t c, void* d, cl_int* e), (a, b, c, d, e), nullptr)
FWD(cl_program, clCreateProgramWithSource, (cl_context a, cl_uint b, const char** c, const size_t* d, cl_int* e), (a, b, c, d, e), nullptr)
FWD(cl_int, clBuildProgram, (cl_program a, cl_uint b, const cl_device_id* c, const char* d, void (*e)(cl_program, void*), void* f), (a, b, c, d, e, f), ERR)
FWD(cl_int, clGetProgramBuildInfo, (cl_program a, cl_device_id b, cl_program_build_info c, size_t d, void* e, size_t* f), (a, b, c, d, e, f), ERR)
FWD(cl_kernel, clCreateKernel, (cl_program a, const char* b, cl_int* c), (a, b, c), nullptr)
FWD(cl_int, clReleaseProgram, (cl_program a), (a), ERR)
FWD(cl_int, clEnqueueWriteBuffer,
    (cl_command_queue a, cl_mem b, cl_bool c, size_t d, size_t e, const void* f, cl_uint g, const cl_event* h, cl_event* i), (a, b, c, d, e, f, g, h, i), ERR)
FWD(cl_int, clEnqueueReadBuffer, (cl_command_queue a, cl_mem b, cl_bool c, size_t d, size_t e, void* f, cl_uint g, const cl_event* h, cl_event* i),
    (a, b, c, d, e, f, g, h, i), ERR)
FWD(cl_int, clEnqueueFillBuffer,
    (cl_command_queue a, cl_mem b, const void* c, size_t d, size_t e, size_t f, cl_uint g, const cl_event* h, cl_event* i), (a, b, c, d, e, f, g, h, i), ERR)
FWD(cl_int, clWaitForEvents, (cl_uint a, const cl_event* b), (a, b), ERR)
FWD(cl_int, clGetEventProfilingInfo, (cl_event a, cl_profiling_info b, size_t c, void* d, size_t* e), (a, b, c, d, e), ERR)
FWD(cl_int, clFinish, (cl_command_queue a), (a), ERR)
FWD(cl_int, clReleaseMemObject, (cl_mem a), (a), ERR)
FWD(cl_int, clReleaseKernel, (cl_kernel a), (a), ERR)
FWD(cl_int, clReleaseCommandQueue, (cl_command_queue a), (a), ERR)
FWD(cl_int, clReleaseContext, (cl_context a), (a), ERR)
FWD(cl_int, clReleaseDevice, (cl_device_id a), (a), ERR)

extern "C" cl_int clSetKernelArg(cl_kernel k, cl_uint i, size_t n, const void* p) {
    typedef cl_int (*fn_t)(cl_kernel, cl_uint, size_t, const void*);
    static fn_t fn = (fn_t)resolve("clSetKernelArg");
    if (!fn) return ERR;
    if (hrref_on_set_kernel_arg) hrref_on_set_kernel_arg(k, i, n, p);
    return fn(k, i, n, p);
}

extern "C" cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint dim, const size_t* off, const size_t* g, const size_t* l, cl_uint ne,
                                         const cl_event* wl, cl_event* ev) {
    typedef cl_int (*fn_t)(cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint, const cl_event*, cl_event*);
    static fn_t fn = (fn_t)resolve("clEnqueueNDRangeKernel");
    if (!fn) return ERR;
    const cl_int rc = fn(q, k, dim, off, g, l, ne, wl, ev);
    if (rc == CL_SUCCESS && hrref_on_enqueue_kernel) hrref_on_enqueue_kernel(q, k);
    return rc;
}
