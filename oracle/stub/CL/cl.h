/*
 * Minimal stand-in for <CL/cl.h> (TEST INFRASTRUCTURE, oracle/ only).
 *
 * The image has no OpenCL headers or ICD.  This header declares
 * exactly the OpenCL types, constants and entry points that the reference's
 * encoder.c / decoder.c / OpenCLUtils.h name, so that those files compile
 * UNMODIFIED from /root/reference; oracle/ref_shim.c serves the calls on the
 * CPU.  Written from the OpenCL 1.x API's public signatures, not from any
 * reference file.
 */
#ifndef ORACLE_STUB_CL_H
#define ORACLE_STUB_CL_H
#include <stddef.h>
#include <stdint.h>

typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef float    cl_float;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef intptr_t cl_context_properties;

typedef struct orc_cl_platform *cl_platform_id;
typedef struct orc_cl_device   *cl_device_id;
typedef struct orc_cl_context  *cl_context;
typedef struct orc_cl_program  *cl_program;
typedef struct orc_cl_kernel   *cl_kernel;
typedef struct orc_cl_queue    *cl_command_queue;
typedef struct orc_cl_mem      *cl_mem;
typedef struct orc_cl_event    *cl_event;

#define CL_SUCCESS 0
#define CL_TRUE 1
#define CL_FALSE 0
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY  (1 << 2)

cl_context clCreateContext(const cl_context_properties *props, cl_uint ndev, const cl_device_id *devs,
                           void (*notify)(const char *, const void *, size_t, void *), void *user, cl_int *err);
cl_mem clCreateBuffer(cl_context ctx, cl_mem_flags flags, size_t size, void *host, cl_int *err);
cl_command_queue clCreateCommandQueue(cl_context ctx, cl_device_id dev, cl_command_queue_properties props, cl_int *err);
cl_kernel clCreateKernel(cl_program prog, const char *name, cl_int *err);
cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value);
cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t off, size_t size,
                            const void *ptr, cl_uint nwait, const cl_event *wait, cl_event *ev);
cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t off, size_t size,
                           void *ptr, cl_uint nwait, const cl_event *wait, cl_event *ev);
cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint dim, const size_t *goff,
                              const size_t *gsize, const size_t *lsize, cl_uint nwait, const cl_event *wait, cl_event *ev);
#endif
