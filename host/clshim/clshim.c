/* clshim.c -- zero-source-change mode: the eight cl* entry points and the four OpenCLUtils helpers that the
 * reference's C codec calls (3d-DCT-video-encoding-OpenCL/encoder.c:148-254, decoder.c:153-292, OpenCLUtils.h:13-21),
 * served by libdct3d.so.  The reference's encoder.c, decoder.c, main.c, ExpGolomb.c and CubeUtils.c compile UNMODIFIED against
 * host/clshim/CL/cl.h and link with this file (host/Makefile, target refgpu): their two-kernel sequences
 *
 *     dct_calculate_partial_sums  -> dct_aggregate_partial_sums       (3dDCT.cl:43-143)
 *     idct_calculate_partial_sums -> idct_aggregate_partial_sums      (3dDCT.cl:164-265)
 *
 * become ONE call each, dct3d_forward_f32 / dct3d_inverse_f32, on the same float cube-major slabs: the first kernel of a
 * pair only records which buffer feeds the pair, the second one runs the separable transform on the GPU and leaves the
 * coefficients (or the clamped pixels) in its output buffer.  Everything else the reference does -- readCubes, quantisation,
 * Exp-Golomb, zlib -- still runs in its own host code: this mode trades speed for not touching a line of it.
 * Buffers are page-locked host memory (dct3d_host_alloc); the transform calls move them over PCIe.
 */
#include <CL/cl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dct3d.h"

struct clshim_platform { int index; };
struct clshim_device { int index; };
struct clshim_context { int device; };
struct clshim_program { int unused; };
struct clshim_queue { int unused; };
struct clshim_event { int unused; };
struct clshim_mem { size_t size; void *data; struct clshim_mem *source; };   /* source: the buffer whose transform this one stands for */
struct clshim_kernel { int inverse, aggregate; int dims[3]; cl_mem in, out; };

static struct clshim_device the_device;
static struct clshim_context the_context;
static struct clshim_program the_program;
static struct clshim_queue the_queue;
static struct clshim_event the_event;
static dct3d_ctx *the_ctx;
static size_t the_ctx_cubes;
static int the_ctx_edge;

/* ---- OpenCLUtils.h:13-21 --------------------------------------------------------------------------------------- */
void printAvailablePlatforms(void)
{
    char buf[4096];
    if (dct3d_list_devices(buf, sizeof buf) < 0) printf("%s\n", dct3d_last_error(NULL));
    else printf("%s", buf);
}
cl_platform_id getPlatformIdForIndex(int platformIndex) { (void)platformIndex; return NULL; }
cl_device_id getDeviceId(int platformIndex) { the_device.index = platformIndex > 0 ? platformIndex - 1 : 0; return &the_device; }
size_t getMaxWorkGroupSize(cl_device_id deviceId) { (void)deviceId; return 1024; }
cl_program buildKernel(cl_context context, cl_device_id deviceId, const char *fileName) { (void)context; (void)deviceId; (void)fileName; return &the_program; }

/* ---- cl* -------------------------------------------------------------------------------------------------------- */
cl_context clCreateContext(const cl_context_properties *properties, cl_uint num_devices, const cl_device_id *devices,
                           void (*pfn_notify)(const char *, const void *, size_t, void *), void *user_data, cl_int *errcode_ret)
{
    (void)properties; (void)pfn_notify; (void)user_data;
    the_context.device = (num_devices && devices && devices[0]) ? devices[0]->index : 0;
    if (dct3d_device_count() <= the_context.device) { if (errcode_ret) *errcode_ret = CL_OUT_OF_RESOURCES; return NULL; }   /* no CPU fallback */
    if (errcode_ret) *errcode_ret = CL_SUCCESS;
    return &the_context;
}

cl_mem clCreateBuffer(cl_context context, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *errcode_ret)
{
    (void)context; (void)flags; (void)host_ptr;
    cl_mem m = (cl_mem)calloc(1, sizeof *m);
    if (m) m->size = size;                  /* storage is allocated on first use: the partial-sums buffer never needs any */
    if (errcode_ret) *errcode_ret = m ? CL_SUCCESS : CL_OUT_OF_RESOURCES;
    return m;
}

static void *storage(cl_mem m)
{
    if (!m->data) {
        m->data = dct3d_host_alloc(m->size);
        if (m->data) memset(m->data, 0, m->size);
    }
    return m->data;
}

cl_command_queue clCreateCommandQueue(cl_context context, cl_device_id device, cl_command_queue_properties properties, cl_int *errcode_ret)
{ (void)context; (void)device; (void)properties; if (errcode_ret) *errcode_ret = CL_SUCCESS; return &the_queue; }

cl_kernel clCreateKernel(cl_program program, const char *kernel_name, cl_int *errcode_ret)
{
    (void)program;
    static const char *names[4] = {"dct_calculate_partial_sums", "dct_aggregate_partial_sums", "idct_calculate_partial_sums", "idct_aggregate_partial_sums"};
    for (int i = 0; i < 4; i++) {
        if (strcmp(kernel_name, names[i])) continue;
        cl_kernel k = (cl_kernel)calloc(1, sizeof *k);
        if (!k) break;
        k->inverse = i >= 2;
        k->aggregate = i & 1;
        if (errcode_ret) *errcode_ret = CL_SUCCESS;
        return k;
    }
    if (errcode_ret) *errcode_ret = CL_INVALID_KERNEL_NAME;
    return NULL;
}

cl_int clSetKernelArg(cl_kernel kernel, cl_uint arg_index, size_t arg_size, const void *arg_value)
{
    (void)arg_size;
    if (!kernel) return CL_INVALID_VALUE;
    if (arg_index < 3) kernel->dims[arg_index] = *(const cl_int *)arg_value;      /* cubeWidth, cubeHeight, cubeDepth */
    else if (arg_index == 3) kernel->in = *(const cl_mem *)arg_value;
    else if (arg_index == 4) kernel->out = *(const cl_mem *)arg_value;
    /* 5: the work-group scratch size of the reduction, which no longer exists */
    return CL_SUCCESS;
}

cl_int clEnqueueWriteBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking_write, size_t offset, size_t size, const void *ptr,
                            cl_uint num_events_in_wait_list, const cl_event *event_wait_list, cl_event *event)
{
    (void)queue; (void)blocking_write; (void)num_events_in_wait_list; (void)event_wait_list; (void)event;
    if (!buffer || offset + size > buffer->size || !storage(buffer)) return CL_INVALID_VALUE;
    memcpy((char *)buffer->data + offset, ptr, size);
    buffer->source = NULL;
    return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking_read, size_t offset, size_t size, void *ptr,
                           cl_uint num_events_in_wait_list, const cl_event *event_wait_list, cl_event *event)
{
    (void)queue; (void)blocking_read; (void)num_events_in_wait_list; (void)event_wait_list; (void)event;
    if (!buffer || offset + size > buffer->size || !storage(buffer)) return CL_INVALID_VALUE;
    memcpy(ptr, (char *)buffer->data + offset, size);
    return CL_SUCCESS;
}

cl_int clEnqueueNDRangeKernel(cl_command_queue queue, cl_kernel kernel, cl_uint work_dim, const size_t *global_work_offset,
                              const size_t *global_work_size, const size_t *local_work_size, cl_uint num_events_in_wait_list,
                              const cl_event *event_wait_list, cl_event *event)
{
    (void)queue; (void)work_dim; (void)global_work_offset; (void)global_work_size; (void)local_work_size;
    (void)num_events_in_wait_list; (void)event_wait_list;
    if (!kernel || !kernel->in || !kernel->out) return CL_INVALID_KERNEL_ARGS;
    if (event) *event = &the_event;
    if (!kernel->aggregate) {               /* first kernel of the pair: its output stands for "the transform of kernel->in" */
        kernel->out->source = kernel->in;
        return CL_SUCCESS;
    }
    cl_mem src = kernel->in->source;        /* second kernel: transform the pair's real input into kernel->out */
    const int edge = kernel->dims[0];
    if (!src || (edge != 8 && edge != 4) || kernel->dims[1] != edge || kernel->dims[2] != edge) return CL_INVALID_KERNEL_ARGS;
    const size_t cube = (size_t)edge * edge * edge, ncubes = src->size / (sizeof(float) * cube);
    if (ncubes == 0 || kernel->out->size < src->size || !storage(src) || !storage(kernel->out)) return CL_INVALID_VALUE;
    if (!the_ctx || the_ctx_cubes != ncubes || the_ctx_edge != edge) {
        /* cube-major slabs carry no frame geometry: any frame of ncubes cubes will do (one cube per row) */
        if (the_ctx) dct3d_destroy(the_ctx);
        the_ctx = NULL;
        if (ncubes > (size_t)1 << 27 || dct3d_create(&the_ctx, the_context.device, edge, (int)(ncubes * edge), edge) != DCT3D_OK) {
            fprintf(stderr, "clshim: %s\n", dct3d_last_error(NULL));
            return CL_OUT_OF_RESOURCES;
        }
        the_ctx_cubes = ncubes;
        the_ctx_edge = edge;
    }
    const int rc = kernel->inverse ? dct3d_inverse_f32(the_ctx, (const float *)src->data, (float *)kernel->out->data, 1)
                                   : dct3d_forward_f32(the_ctx, (const float *)src->data, (float *)kernel->out->data, 1);
    if (rc != DCT3D_OK) { fprintf(stderr, "clshim: %s\n", dct3d_last_error(the_ctx)); return CL_OUT_OF_RESOURCES; }
    return CL_SUCCESS;
}
