/*
 * ref_shim.c -- CPU executor for the OpenCL calls made by the reference's
 * encoder.c / decoder.c (TEST INFRASTRUCTURE, oracle/ only).
 *
 * oracle/Makefile compiles the reference's C files UNMODIFIED from
 * /root/reference together with this shim into oracle/_ref/.  The shim serves
 *   - the 8 cl* entry points used at C/encoder.c:158-254, C/decoder.c:163-292
 *   - the 4 helpers declared in C/OpenCLUtils.h:13-21 (OpenCLUtils.c itself
 *     needs a real OpenCL platform and is not compiled)
 * and executes the four kernels of C/3dDCT.cl by name with their documented
 * semantics (float arithmetic, work-group tree reduction, cosf standing in for
 * the implementation-defined native_cos).  It is a restatement of the kernels'
 * behaviour written for this repo, not a copy of their source.
 *
 * Work-groups are independent, so they are run in parallel on pthreads
 * (ORC_SHIM_THREADS, default = online cores); the arithmetic inside one
 * work-group follows the kernel's order exactly.
 */
#include <CL/cl.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct orc_cl_platform { int unused; };
struct orc_cl_device { int unused; };
struct orc_cl_context { int unused; };
struct orc_cl_program { int unused; };
struct orc_cl_queue { int unused; };
struct orc_cl_event { int unused; };
struct orc_cl_mem { size_t size; void *data; };
struct orc_cl_kernel { int which; int dims[3]; cl_mem in, out; };

enum { K_DCT_CALC, K_DCT_AGG, K_IDCT_CALC, K_IDCT_AGG };

static struct orc_cl_device the_device;
static struct orc_cl_context the_context;
static struct orc_cl_program the_program;
static struct orc_cl_queue the_queue;
static struct orc_cl_event the_event;

/* ---- C/OpenCLUtils.h:13-21 ------------------------------------------------ */
void printAvailablePlatforms(void) { printf("1 - oracle CPU shim (no OpenCL platform in this image)\n"); }
cl_platform_id getPlatformIdForIndex(int platformIndex) { (void)platformIndex; return NULL; }
cl_device_id getDeviceId(int platformIndex) { (void)platformIndex; return &the_device; }
size_t getMaxWorkGroupSize(cl_device_id deviceId)
{
    (void)deviceId;
    const char *e = getenv("ORC_SHIM_MAX_WG"); /* lets tests exercise the partial-sums stage */
    return e ? (size_t)atol(e) : 1024;
}
cl_program buildKernel(cl_context context, cl_device_id deviceId, const char *fileName)
{ (void)context; (void)deviceId; (void)fileName; return &the_program; }

/* ---- cl* entry points ----------------------------------------------------- */
cl_context clCreateContext(const cl_context_properties *props, cl_uint ndev, const cl_device_id *devs,
                           void (*notify)(const char *, const void *, size_t, void *), void *user, cl_int *err)
{ (void)props; (void)ndev; (void)devs; (void)notify; (void)user; if (err) *err = CL_SUCCESS; return &the_context; }

cl_mem clCreateBuffer(cl_context ctx, cl_mem_flags flags, size_t size, void *host, cl_int *err)
{
    (void)ctx; (void)flags; (void)host;
    cl_mem m = (cl_mem)malloc(sizeof(*m));
    m->size = size; m->data = calloc(1, size);
    if (err) *err = CL_SUCCESS;
    return m;
}

cl_command_queue clCreateCommandQueue(cl_context ctx, cl_device_id dev, cl_command_queue_properties props, cl_int *err)
{ (void)ctx; (void)dev; (void)props; if (err) *err = CL_SUCCESS; return &the_queue; }

cl_kernel clCreateKernel(cl_program prog, const char *name, cl_int *err)
{
    (void)prog;
    int which = -1;
    if (!strcmp(name, "dct_calculate_partial_sums")) which = K_DCT_CALC;
    else if (!strcmp(name, "dct_aggregate_partial_sums")) which = K_DCT_AGG;
    else if (!strcmp(name, "idct_calculate_partial_sums")) which = K_IDCT_CALC;
    else if (!strcmp(name, "idct_aggregate_partial_sums")) which = K_IDCT_AGG;
    if (which < 0) { if (err) *err = -46; return NULL; }
    cl_kernel k = (cl_kernel)calloc(1, sizeof(*k));
    k->which = which;
    if (err) *err = CL_SUCCESS;
    return k;
}

cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value)
{
    (void)size;
    if (idx < 3) k->dims[idx] = *(const cl_int *)value;      /* cubeWidth, cubeHeight, cubeDepth */
    else if (idx == 3) k->in = *(const cl_mem *)value;
    else if (idx == 4) k->out = *(const cl_mem *)value;
    /* idx 5: __local scratch size; the executor allocates its own */
    return CL_SUCCESS;
}

cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t off, size_t size,
                            const void *ptr, cl_uint nwait, const cl_event *wait, cl_event *ev)
{ (void)q; (void)blocking; (void)nwait; (void)wait; (void)ev; memcpy((char *)buf->data + off, ptr, size); return CL_SUCCESS; }

cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t off, size_t size,
                           void *ptr, cl_uint nwait, const cl_event *wait, cl_event *ev)
{ (void)q; (void)blocking; (void)nwait; (void)wait; (void)ev; memcpy(ptr, (char *)buf->data + off, size); return CL_SUCCESS; }

/* In-place tree sum over a work-group's scratch, as the device function at
 * C/3dDCT.cl:11-22 does it: halve the stride each step, fold an odd tail into
 * element 0. */
static float group_tree_sum(float *v, size_t n)
{
    for (size_t stride = n / 2; stride > 0; stride /= 2, n /= 2) {
        for (size_t i = 0; i < stride; i++) v[i] = v[i] + v[i + stride];
        if (n % 2 != 0) v[0] = v[0] + v[n - 1];
    }
    return v[0];
}

typedef struct {
    cl_kernel k; size_t gsize, lsize; int inverse; int agg;
    const float *tw, *th, *td; volatile long *next; long ngroups;
} shim_job_t;

static void calc_group(const shim_job_t *j, long g, float *scratch)
{
    cl_kernel k = j->k;
    const int cw = k->dims[0], ch = k->dims[1], cd = k->dims[2];
    const int face = cw * ch, cs = face * cd;
    const size_t lsize = j->lsize;
    const float *in = (const float *)k->in->data;
    float *partial = (float *)k->out->data;
    const size_t groups_per_cube = (size_t)cs / lsize;
    size_t gid0 = (size_t)g * lsize;
    size_t cube = gid0 / cs, item0 = gid0 % cs, gidx = item0 / lsize;
    for (int a0 = 0; a0 < cd; a0++) for (int a1 = 0; a1 < ch; a1++) for (int a2 = 0; a2 < cw; a2++) {
        /* (a0,a1,a2) = output coefficient (forward) or output pixel (inverse) */
        for (size_t i = 0; i < lsize; i++) {
            int item = (int)(item0 + i);
            int b0 = item / face, b1 = (item % face) / cw, b2 = item % cw;
            float value = in[gid0 + i];
            if (!j->inverse) {
                scratch[i] = value * j->td[b0 * cd + a0] * j->th[b1 * ch + a1] * j->tw[b2 * cw + a2];
            } else {
                float c0 = b0 ? 1.0f : (float)M_SQRT1_2, c1 = b1 ? 1.0f : (float)M_SQRT1_2, c2 = b2 ? 1.0f : (float)M_SQRT1_2;
                scratch[i] = value * c0 * c1 * c2 * j->td[a0 * cd + b0] * j->th[a1 * ch + b1] * j->tw[a2 * cw + b2];
            }
        }
        float s = group_tree_sum(scratch, lsize);
        size_t idx = cube * cs + (size_t)a0 * face + (size_t)a1 * cw + a2;
        partial[idx * groups_per_cube + gidx] = s;
    }
}

static void agg_group(const shim_job_t *j, long g, float *scratch)
{
    cl_kernel k = j->k;
    const int cw = k->dims[0], ch = k->dims[1], cd = k->dims[2];
    const int face = cw * ch, cs = face * cd;
    const float scale = sqrtf(8.0f / (float)cs);
    const float *partial = (const float *)k->in->data;
    float *out = (float *)k->out->data;
    for (size_t i = 0; i < j->lsize; i++) scratch[i] = partial[(size_t)g * j->lsize + i];
    float s = group_tree_sum(scratch, j->lsize);
    if (!j->inverse) {
        int idx = (int)((size_t)g % cs);
        int k0 = idx / face, k1 = (idx % face) / cw, k2 = idx % cw;
        float c0 = k0 ? 1.0f : (float)M_SQRT1_2, c1 = k1 ? 1.0f : (float)M_SQRT1_2, c2 = k2 ? 1.0f : (float)M_SQRT1_2;
        out[g] = s * scale * c0 * c1 * c2;
    } else {
        float v = s * scale;
        if (v > 255) v = 255; else if (v < 0) v = 0;
        out[g] = v;
    }
}

static void *shim_worker(void *arg)
{
    const shim_job_t *j = (const shim_job_t *)arg;
    float *scratch = malloc(sizeof(float) * j->lsize);
    const long chunk = j->agg ? 4096 : 8;
    for (;;) {
        long g0 = __sync_fetch_and_add(j->next, chunk);
        if (g0 >= j->ngroups) break;
        long g1 = g0 + chunk < j->ngroups ? g0 + chunk : j->ngroups;
        for (long g = g0; g < g1; g++) { if (j->agg) agg_group(j, g, scratch); else calc_group(j, g, scratch); }
    }
    free(scratch);
    return NULL;
}

static void shim_run(shim_job_t *j)
{
    volatile long next = 0;
    j->next = &next;
    j->ngroups = (long)(j->gsize / j->lsize);
    const char *e = getenv("ORC_SHIM_THREADS");
    long nt = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 256) nt = 256;
    pthread_t th[256];
    for (long t = 0; t < nt; t++) pthread_create(&th[t], NULL, shim_worker, j);
    for (long t = 0; t < nt; t++) pthread_join(th[t], NULL);
}

static void run_calc(cl_kernel k, size_t gsize, size_t lsize, int inverse)
{
    const int cw = k->dims[0], ch = k->dims[1], cd = k->dims[2];
    const float piw = (float)M_PI / (float)cw, pih = (float)M_PI / (float)ch, pid = (float)M_PI / (float)cd;
    /* cos tables indexed [n][k], evaluated the way the kernel writes the angle */
    float *tw = malloc(sizeof(float) * cw * cw), *th = malloc(sizeof(float) * ch * ch), *td = malloc(sizeof(float) * cd * cd);
    for (int n = 0; n < cw; n++) for (int q = 0; q < cw; q++) tw[n * cw + q] = cosf(piw * (n + 0.5f) * q);
    for (int n = 0; n < ch; n++) for (int q = 0; q < ch; q++) th[n * ch + q] = cosf(pih * (n + 0.5f) * q);
    for (int n = 0; n < cd; n++) for (int q = 0; q < cd; q++) td[n * cd + q] = cosf(pid * (n + 0.5f) * q);
    shim_job_t j = { k, gsize, lsize, inverse, 0, tw, th, td, NULL, 0 };
    shim_run(&j);
    free(tw); free(th); free(td);
}

static void run_agg(cl_kernel k, size_t gsize, size_t lsize, int inverse)
{
    shim_job_t j = { k, gsize, lsize, inverse, 1, NULL, NULL, NULL, NULL, 0 };
    shim_run(&j);
}

cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint dim, const size_t *goff,
                              const size_t *gsize, const size_t *lsize, cl_uint nwait, const cl_event *wait, cl_event *ev)
{
    (void)q; (void)dim; (void)goff; (void)nwait; (void)wait;
    if (!k || !k->in || !k->out || !gsize || !lsize || *lsize == 0) return -52;
    switch (k->which) {
    case K_DCT_CALC:  run_calc(k, *gsize, *lsize, 0); break;
    case K_IDCT_CALC: run_calc(k, *gsize, *lsize, 1); break;
    case K_DCT_AGG:   run_agg(k, *gsize, *lsize, 0); break;
    case K_IDCT_AGG:  run_agg(k, *gsize, *lsize, 1); break;
    }
    if (ev) *ev = &the_event;
    return CL_SUCCESS;
}
