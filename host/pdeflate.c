/* pdeflate.c -- see pdeflate.h.
 *
 * Input is cut into blocks; every block is compressed by its own raw-deflate state (windowBits -15),
 * primed with the last 32 KiB of the previous block as dictionary so that matches across the cut are
 * not lost, and ended with Z_SYNC_FLUSH (an empty stored block: the output is byte aligned and the
 * deflate stream simply continues).  The last block ends with Z_FINISH.  zlib header, the blocks in
 * order, and the Adler-32 of everything (adler32_combine over the blocks) make one RFC 1950 stream.
 * Worker threads take blocks in order; a writer thread writes finished blocks in order.
 */
#include "pdeflate.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#define DICT_MAX 32768

typedef struct job {
    unsigned char *in, *out, *dict;
    size_t in_len, out_len, out_cap, dict_len;
    uLong adler;
    int final, done, failed;
} job;

struct pdeflate {
    FILE *out;
    int level, nthreads, failed, closing;
    size_t block;
    unsigned char *pend;            /* input not yet cut into a block */
    size_t pend_len;
    unsigned char dict[DICT_MAX];   /* tail of the previous block */
    size_t dict_len;
    /* ring of jobs: [written, taken) are with the workers or waiting for the writer, [taken, submitted) are queued */
    job *ring;
    size_t nring, submitted, taken, written;
    pthread_mutex_t mu;
    pthread_cond_t cv_work, cv_done, cv_space;
    pthread_t *workers, writer;
    uLong adler;
    unsigned long long in_total, out_total;
};

static void compress_job(const pdeflate *p, job *j)
{
    z_stream s;
    memset(&s, 0, sizeof s);
    if (deflateInit2(&s, p->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { j->failed = 1; return; }
    if (j->dict_len) deflateSetDictionary(&s, j->dict, (uInt)j->dict_len);
    j->out_cap = deflateBound(&s, (uLong)j->in_len) + 64;
    j->out = (unsigned char *)malloc(j->out_cap);
    if (!j->out) { j->failed = 1; deflateEnd(&s); return; }
    s.next_in = j->in;
    s.avail_in = (uInt)j->in_len;
    s.next_out = j->out;
    s.avail_out = (uInt)j->out_cap;
    const int zr = deflate(&s, j->final ? Z_FINISH : Z_SYNC_FLUSH);
    if ((j->final && zr != Z_STREAM_END) || (!j->final && (zr != Z_OK || s.avail_in != 0 || s.avail_out == 0))) j->failed = 1;
    j->out_len = j->out_cap - s.avail_out;
    j->adler = adler32(adler32(0L, Z_NULL, 0), j->in, (uInt)j->in_len);
    deflateEnd(&s);
}

static void *worker_main(void *arg)
{
    pdeflate *p = (pdeflate *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (p->taken == p->submitted && !p->closing) pthread_cond_wait(&p->cv_work, &p->mu);
        if (p->taken == p->submitted) { pthread_mutex_unlock(&p->mu); return NULL; }
        job *j = &p->ring[p->taken++ % p->nring];
        pthread_mutex_unlock(&p->mu);
        compress_job(p, j);
        pthread_mutex_lock(&p->mu);
        j->done = 1;
        pthread_cond_broadcast(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
}

static void *writer_main(void *arg)
{
    pdeflate *p = (pdeflate *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        job *j = &p->ring[p->written % p->nring];
        while (!(p->written < p->submitted && j->done) && !(p->closing && p->written == p->submitted))
            pthread_cond_wait(&p->cv_done, &p->mu);
        if (p->written == p->submitted) { pthread_mutex_unlock(&p->mu); return NULL; }
        pthread_mutex_unlock(&p->mu);
        int bad = j->failed;
        if (!bad && fwrite(j->out, 1, j->out_len, p->out) != j->out_len) bad = 1;
        pthread_mutex_lock(&p->mu);
        p->adler = adler32_combine(p->adler, j->adler, (z_off_t)j->in_len);
        p->in_total += j->in_len;
        p->out_total += j->out_len;
        if (bad) p->failed = 1;
        free(j->in); free(j->out); free(j->dict);
        memset(j, 0, sizeof *j);
        p->written++;
        pthread_cond_broadcast(&p->cv_space);
        pthread_cond_broadcast(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
}

pdeflate *pdeflate_open(FILE *out, int level, int threads, size_t block)
{
    pdeflate *p = (pdeflate *)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->out = out;
    p->level = level;
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads < 1) threads = 1;
    p->nthreads = threads;
    p->block = block ? block : (size_t)256 << 10;
    p->pend = (unsigned char *)malloc(p->block);
    p->nring = (size_t)threads * 4;
    p->ring = (job *)calloc(p->nring, sizeof(job));
    p->workers = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    p->adler = adler32(0L, Z_NULL, 0);
    pthread_mutex_init(&p->mu, NULL);
    pthread_cond_init(&p->cv_work, NULL);
    pthread_cond_init(&p->cv_done, NULL);
    pthread_cond_init(&p->cv_space, NULL);
    if (!p->pend || !p->ring || !p->workers) { free(p->pend); free(p->ring); free(p->workers); free(p); return NULL; }
    /* RFC 1950 header: CM = 8, CINFO = 7 (32 KiB window), FLEVEL from the level, no preset dictionary */
    const unsigned flevel = level >= 7 ? 3u : (level == 6 || level < 0) ? 2u : level >= 2 ? 1u : 0u;
    unsigned hdr = (0x78u << 8) | (flevel << 6);
    hdr += 31u - hdr % 31u;
    const unsigned char h[2] = {(unsigned char)(hdr >> 8), (unsigned char)hdr};
    if (fwrite(h, 1, 2, out) != 2) p->failed = 1;
    p->out_total = 2;
    for (int i = 0; i < threads; i++) pthread_create(&p->workers[i], NULL, worker_main, p);
    pthread_create(&p->writer, NULL, writer_main, p);
    return p;
}

/* hands the first `n` pending bytes to the workers as one block */
static int submit(pdeflate *p, size_t n, int final)
{
    job j;
    memset(&j, 0, sizeof j);
    j.in = (unsigned char *)malloc(n ? n : 1);
    j.dict = p->dict_len ? (unsigned char *)malloc(p->dict_len) : NULL;
    if (!j.in || (p->dict_len && !j.dict)) { free(j.in); free(j.dict); return -1; }
    memcpy(j.in, p->pend, n);
    j.in_len = n;
    if (p->dict_len) memcpy(j.dict, p->dict, p->dict_len);
    j.dict_len = p->dict_len;
    j.final = final;
    /* the tail of this block primes the next one */
    if (n >= DICT_MAX) { memcpy(p->dict, p->pend + n - DICT_MAX, DICT_MAX); p->dict_len = DICT_MAX; }
    else if (n) {
        const size_t keep = p->dict_len + n > DICT_MAX ? DICT_MAX - n : p->dict_len;
        memmove(p->dict, p->dict + p->dict_len - keep, keep);
        memcpy(p->dict + keep, p->pend, n);
        p->dict_len = keep + n;
    }
    pthread_mutex_lock(&p->mu);
    while (p->submitted - p->written >= p->nring) pthread_cond_wait(&p->cv_space, &p->mu);
    p->ring[p->submitted++ % p->nring] = j;
    pthread_cond_signal(&p->cv_work);
    const int bad = p->failed;
    pthread_mutex_unlock(&p->mu);
    return bad ? -1 : 0;
}

int pdeflate_write(pdeflate *p, const unsigned char *data, size_t n)
{
    while (n) {
        const size_t take = n < p->block - p->pend_len ? n : p->block - p->pend_len;
        memcpy(p->pend + p->pend_len, data, take);
        p->pend_len += take;
        data += take;
        n -= take;
        if (p->pend_len == p->block) {
            if (submit(p, p->pend_len, 0)) return -1;
            p->pend_len = 0;
        }
    }
    return 0;
}

int pdeflate_close(pdeflate *p, unsigned long long *in_bytes, unsigned long long *out_bytes)
{
    int rc = submit(p, p->pend_len, 1);         /* the final block, possibly empty */
    pthread_mutex_lock(&p->mu);
    p->closing = 1;
    pthread_cond_broadcast(&p->cv_work);
    pthread_cond_broadcast(&p->cv_done);
    pthread_mutex_unlock(&p->mu);
    for (int i = 0; i < p->nthreads; i++) pthread_join(p->workers[i], NULL);
    pthread_join(p->writer, NULL);
    const unsigned char t[4] = {(unsigned char)(p->adler >> 24), (unsigned char)(p->adler >> 16),
                                (unsigned char)(p->adler >> 8), (unsigned char)p->adler};
    if (fwrite(t, 1, 4, p->out) != 4) p->failed = 1;
    p->out_total += 4;
    if (p->failed) rc = -1;
    if (in_bytes) *in_bytes = p->in_total;
    if (out_bytes) *out_bytes = p->out_total;
    pthread_mutex_destroy(&p->mu);
    pthread_cond_destroy(&p->cv_work);
    pthread_cond_destroy(&p->cv_done);
    pthread_cond_destroy(&p->cv_space);
    free(p->pend); free(p->ring); free(p->workers);
    free(p);
    return rc;
}
