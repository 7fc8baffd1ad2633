/*
 * CL/cl.h of the zero-source-change mode (host/clshim): the OpenCL 1.x types, constants and entry points that the
 * reference's C codec names (3d-DCT-video-encoding-OpenCL/encoder.c:148-254, decoder.c:153-292, OpenCLUtils.h:11-21),
 * declared from the public OpenCL API so that those files compile unmodified; host/clshim/clshim.c serves the calls with
 * libdct3d.so.  The build image has no OpenCL headers or ICD.
 */
#ifndef DCT3D_CLSHIM_CL_H
#define DCT3D_CLSHIM_CL_H
#include <stddef.h>
#include <stdint.h>

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef float cl_float;
typedef cl_uint cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef intptr_t cl_context_properties;

typedef struct clshim_platform *cl_platform_id;
typedef struct clshim_device *cl_device_id;
typedef struct clshim_context *cl_context;
typedef struct clshim_program *cl_program;
typedef struct clshim_kernel *cl_kernel;
typedef struct clshim_queue *cl_command_queue;
typedef struct clshim_mem *cl_mem;
typedef struct clshim_event *cl_event;

#define CL_SUCCESS 0
#define CL_INVALID_VALUE (-30)
#define CL_INVALID_KERNEL_NAME (-46)
#define CL_INVALID_KERNEL_ARGS (-52)
#define CL_OUT_OF_RESOURCES (-5)
#define CL_TRUE 1
#define CL_FALSE 0
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)

cl_context clCreateContext(const cl_context_properties *properties, cl_uint num_devices, const cl_device_id *devices,
                           void (*pfn_notify)(const char *, const void *, size_t, void *), void *user_data, cl_int *errcode_ret);
cl_mem clCreateBuffer(cl_context context, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *errcode_ret);
cl_command_queue clCreateCommandQueue(cl_context context, cl_device_id device, cl_command_queue_properties properties, cl_int *errcode_ret);
cl_kernel clCreateKernel(cl_program program, const char *kernel_name, cl_int *errcode_ret);
cl_int clSetKernelArg(cl_kernel kernel, cl_uint arg_index, size_t arg_size, const void *arg_value);
cl_int clEnqueueWriteBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking_write, size_t offset, size_t size, const void *ptr,
                            cl_uint num_events_in_wait_list, const cl_event *event_wait_list, cl_event *event);
cl_int clEnqueueReadBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking_read, size_t offset, size_t size, void *ptr,
                           cl_uint num_events_in_wait_list, const cl_event *event_wait_list, cl_event *event);
cl_int clEnqueueNDRangeKernel(cl_command_queue queue, cl_kernel kernel, cl_uint work_dim, const size_t *global_work_offset,
                              const size_t *global_work_size, const size_t *local_work_size, cl_uint num_events_in_wait_list,
                              const cl_event *event_wait_list, cl_event *event);
#endif
