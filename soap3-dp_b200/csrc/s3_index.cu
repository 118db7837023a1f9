// s3_index.cu -- index upload and one-time re-layout for B200.
// Replaces GPUINDEXUpload / GPUINDEXFree (alignment.cu:27-115).
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

static thread_local char g_err[512] = "";
unsigned long long g_s3_launches = 0;
extern "C" unsigned long long s3_launch_count(void) { return g_s3_launches; }

void s3_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *s3_last_error(void) { return g_err; }

extern "C" int s3_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int s3_pipe_init(S3Pipe *p)
{
    if (p->ready) return S3_OK;
    S3_CUDA(cudaStreamCreateWithFlags(&p->in, cudaStreamNonBlocking));
    S3_CUDA(cudaStreamCreateWithFlags(&p->out, cudaStreamNonBlocking));
    for (int i = 0; i < S3_PIPE_CHUNKS; ++i) {
        S3_CUDA(cudaEventCreateWithFlags(&p->up[i], cudaEventDisableTiming));
        S3_CUDA(cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming));
    }
    p->ready = 1;
    return S3_OK;
}

void s3_pipe_destroy(S3Pipe *p)
{
    if (!p->ready) return;
    cudaStreamDestroy(p->in); cudaStreamDestroy(p->out);
    for (int i = 0; i < S3_PIPE_CHUNKS; ++i) { cudaEventDestroy(p->up[i]); cudaEventDestroy(p->done[i]); }
    p->ready = 0;
}

void s3_timing_mark(S3Timing *t, cudaStream_t st, int slot)
{
    if (!t->on || t->n >= S3_TIMING_MARKS) return;
    if (t->created <= t->n) { if (cudaEventCreate(&t->ev[t->n]) != cudaSuccess) return; t->created = t->n + 1; }
    cudaEventRecord(t->ev[t->n], st);
    t->slot[t->n] = slot;
    ++t->n;
}

// sums the marked intervals per slot since the last read, and forgets them
int s3_timing_read(S3Timing *t, cudaStream_t st, float *msPerSlot, int *launchesPerSlot)
{
    for (int k = 0; k < S3_TIMING_SLOTS; ++k) { msPerSlot[k] = 0.f; if (launchesPerSlot) launchesPerSlot[k] = 0; }
    S3_CUDA(cudaStreamSynchronize(st));
    for (int i = 1; i < t->n; ++i) {
        if (t->slot[i] < 0 || t->slot[i] >= S3_TIMING_SLOTS) continue;
        float ms = 0.f;
        S3_CUDA(cudaEventElapsedTime(&ms, t->ev[i - 1], t->ev[i]));
        msPerSlot[t->slot[i]] += ms;
        if (launchesPerSlot) launchesPerSlot[t->slot[i]] += 1;
    }
    t->n = 0;
    return S3_OK;
}

void s3_timing_destroy(S3Timing *t)
{
    for (int i = 0; i < t->created; ++i) cudaEventDestroy(t->ev[i]);
    t->created = 0; t->n = 0;
}

extern "C" int s3_index_set_timing(s3_index *ix, int on)
{
    if (!ix) { s3_set_error("s3_index_set_timing: NULL index"); return S3_EINVAL; }
    ix->timing.on = on ? 1 : 0; ix->timing.n = 0;
    return S3_OK;
}

extern "C" int s3_index_read_timing(s3_index *ix, float *msPerSlot, int *launchesPerSlot)
{
    if (!ix || !msPerSlot) { s3_set_error("s3_index_read_timing: NULL argument"); return S3_EINVAL; }
    S3_CUDA(cudaSetDevice(ix->device));
    return s3_timing_read(&ix->timing, ix->stream, msPerSlot, launchesPerSlot);
}

int s3_scratch(s3_index *ix, size_t bytes, void **out)
{
    if (bytes > ix->scratchBytes) {
        if (ix->scratch) { S3_CUDA(cudaStreamSynchronize(ix->stream)); S3_CUDA(cudaFree(ix->scratch)); ix->scratch = NULL; ix->scratchBytes = 0; }
        size_t want = bytes + bytes / 4;
        S3_CUDA(cudaMalloc(&ix->scratch, want));
        ix->scratchBytes = want;
    }
    *out = ix->scratch;
    return S3_OK;
}

int s3_pinned(s3_index *ix, size_t bytes, void **out)
{
    if (bytes > ix->pinnedBytes) {
        if (ix->pinned) { S3_CUDA(cudaStreamSynchronize(ix->stream)); S3_CUDA(cudaFreeHost(ix->pinned)); ix->pinned = NULL; ix->pinnedBytes = 0; }
        size_t want = bytes + bytes / 4;
        S3_CUDA(cudaMallocHost(&ix->pinned, want));
        ix->pinnedBytes = want;
    }
    *out = ix->pinned;
    return S3_OK;
}

// One thread per bucket.  The running counts at the bucket MIDDLE (64*b + 32) are taken from
// the reference's own sampled table (entry e = floor((64*b+32)/128), sample position 128*e,
// distance 32 or 96 bases) plus a direct count of those bases, i.e. exactly how the reference
// kernel itself would evaluate rank'(c, 64*b+32) (DV-Kernel.cu:256-280) -- no prefix scan
// needed.  The 64 bases become hi/lo bit planes (layout in s3_common.cuh).
__global__ void s3_relayout_kernel(const uint32_t *__restrict__ bwt, const uint32_t *__restrict__ occ,
                                   uint32_t textLength, uint32_t numWords, uint32_t numBuckets,
                                   uint4 *__restrict__ out)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= numBuckets) return;
    const uint32_t start = b * S3_BUCKET_BASES, mid = start + 32;
    const uint32_t e = mid >> 7;
    uint32_t cnt[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) cnt[c] = occ[4 * e + c];
    // bases [128e, mid): whole 16-base words (mid - 128e is 32 or 96); padding past the text = code 0,
    // counted here and by the kernel alike
    for (uint32_t w = (e << 7) >> 4; w < (mid >> 4); ++w) {
        const uint32_t word = w < numWords ? bwt[w] : 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j) cnt[(word >> (2 * (15 - j))) & 3]++;
    }
    // 4 source words (16 bases each, MSB first) -> 2 hi-plane + 2 lo-plane words (32 bases each, LSB first)
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t h = 0, l = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t wi = (start >> 4) + 2 * j + half;
            const uint32_t word = wi < numWords ? bwt[wi] : 0u;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const uint32_t code = (word >> (2 * (15 - t))) & 3;
                h |= (code >> 1) << (16 * half + t);
                l |= (code & 1) << (16 * half + t);
            }
        }
        hi[j] = h; lo[j] = l;
    }
    uint4 *o = out + (size_t)b * 2;
    o[0] = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
    o[1] = make_uint4(hi[0], lo[0], hi[1], lo[1]);
}

// One thread per K-mer: K exact LF-mapping steps from (firstL, n), exactly the loop of the search kernel.
__global__ void s3_seed_build_kernel(S3Half h, uint32_t textLength, uint32_t K, uint32_t firstL, uint2 *__restrict__ table)
{
    const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= (1u << (2 * K))) return;
    uint32_t lo = firstL, hi = textLength;
    for (uint32_t d = 0; d < K; ++d) {
        const uint32_t c = (key >> (2 * (K - 1 - d))) & 3u;
        uint32_t a[4], b[4];
        s3_rank4(h, lo, a);
        s3_rank4(h, hi + 1, b);
        lo = a[c] + 1; hi = b[c];
        if (lo > hi) { lo = 0xFFFFFFF0u | (d + 1); hi = 0; break; }
    }
    table[key] = make_uint2(lo, hi);
}

static int build_seed_tables(s3_index *ix)
{
    // K = floor(log4 n), in [4, 15]: a K-mer then has a handful of occurrences, so a pass is one lookup and
    // one or two steps away from a single suffix; at 3.1 Gbp the three tables are 3 x 8.6 GB, which a
    // 180 GB device has room for (K goes down until they fit in a quarter of the free memory).
    uint32_t K = 0;
    while (K < 15 && (1ull << (2 * (K + 1))) <= ix->textLength) ++K;
    if (K < 4) K = 4;
    size_t freeB = 0, totalB = 0;
    S3_CUDA(cudaMemGetInfo(&freeB, &totalB));
    while (K > 4 && 3 * (sizeof(uint2) << (2 * K)) > freeB / 4) --K;
    if (getenv("S3_SEED_K")) { const int k = atoi(getenv("S3_SEED_K")); if (k >= 4 && k <= 15) K = (uint32_t)k; }   // tuning experiments
    if (getenv("S3_NO_SEED_TABLES")) { ix->seed.K = 0; return S3_OK; }
    const size_t entries = (size_t)1 << (2 * K);
    const S3Half *half[3] = {&ix->fwd, &ix->fwd, &ix->rev};
    const uint32_t firstL[3] = {0u, 1u, 0u};
    for (int t = 0; t < 3; ++t) {
        S3_CUDA(cudaMalloc(&ix->d_seed[t], entries * sizeof(uint2)));
        s3_seed_build_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, ix->stream>>>(*half[t], ix->textLength, K, firstL[t], ix->d_seed[t]);
        S3_LAUNCHED(1);
        S3_CUDA(cudaGetLastError());
        ix->bytes += entries * sizeof(uint2);
    }
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    ix->seed.fwd0 = ix->d_seed[0]; ix->seed.fwd1 = ix->d_seed[1]; ix->seed.rev0 = ix->d_seed[2];
    ix->seed.K = K;
    return S3_OK;
}

static int upload_half(s3_index *ix, const uint32_t *bwt, const uint32_t *occ, uint32_t numOcc,
                       uint32_t textLength, uint4 **d_out, uint32_t *numBucketsOut)
{
    size_t numWords = ((size_t)textLength + 15) / 16;
    uint32_t numBuckets = textLength / S3_BUCKET_BASES + 1;
    uint32_t *d_bwt = NULL, *d_occ = NULL;
    S3_CUDA(cudaMalloc(&d_bwt, numWords * sizeof(uint32_t)));
    S3_CUDA(cudaMalloc(&d_occ, (size_t)numOcc * 4 * sizeof(uint32_t)));
    S3_CUDA(cudaMemcpyAsync(d_bwt, bwt, numWords * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
    S3_CUDA(cudaMemcpyAsync(d_occ, occ, (size_t)numOcc * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
    S3_CUDA(cudaMalloc(d_out, (size_t)numBuckets * 32));
    s3_relayout_kernel<<<(numBuckets + 255) / 256, 256, 0, ix->stream>>>(d_bwt, d_occ, textLength,
                                                                        (uint32_t)numWords, numBuckets, *d_out);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    S3_CUDA(cudaFree(d_bwt));
    S3_CUDA(cudaFree(d_occ));
    *numBucketsOut = numBuckets;
    ix->bytes += (size_t)numBuckets * 32;
    return S3_OK;
}

__global__ void s3_isa_kernel(const uint32_t *__restrict__ sa, uint32_t textLength, uint32_t *__restrict__ isa)
{
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row > textLength) return;
    const uint32_t pos = sa[row];
    if (pos < textLength) isa[pos] = (uint32_t)row;              // row 0 (the '$' suffix, position n or -1) has no entry
}

// Copies the suffix array (n + 1 rows) and the packed text onto the device and inverts the suffix array;
// enables check-and-extend in the search kernel.
static int attach_locate(s3_index *ix, const uint32_t *sa, const uint32_t *packedDNA, cudaMemcpyKind kind)
{
    if (ix->loc.sa) return S3_OK;
    const size_t n = ix->textLength, words = (n + 15) / 16 + 8;
    S3_CUDA(cudaMalloc(&ix->d_packedDNA, words * 4));
    S3_CUDA(cudaMemsetAsync(ix->d_packedDNA, 0, words * 4, ix->stream));
    S3_CUDA(cudaMemcpyAsync(ix->d_packedDNA, packedDNA, ((n + 15) / 16) * 4, kind, ix->stream));
    S3_CUDA(cudaMalloc(&ix->d_sa, (n + 1) * 4));
    S3_CUDA(cudaMemcpyAsync(ix->d_sa, sa, (n + 1) * 4, kind, ix->stream));
    S3_CUDA(cudaMalloc(&ix->d_isa, (n + 1) * 4));
    S3_CUDA(cudaMemsetAsync(ix->d_isa, 0, (n + 1) * 4, ix->stream));
    s3_isa_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ix->stream>>>(ix->d_sa, (uint32_t)n, ix->d_isa);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    ix->bytes += words * 4 + 2 * (n + 1) * 4;
    ix->loc.sa = ix->d_sa; ix->loc.isa = ix->d_isa; ix->loc.text = ix->d_packedDNA;
    return S3_OK;
}

// inside s3_index_upload, once the handle exists: a failed call releases what has been allocated so far (the struct is
// calloc'ed and s3_index_free checks every pointer)
#define S3_CUDA_IX(call)                                                                \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            s3_index_free(ix);                                                          \
            return S3_ECUDA;                                                            \
        }                                                                               \
    } while (0)
extern "C" int s3_index_upload(const uint32_t *bwt, const uint32_t *occ, const uint32_t *revBwt,
                               const uint32_t *revOcc, uint32_t numOcc, uint32_t inverseSa0,
                               uint32_t revInverseSa0, uint32_t textLength, const uint32_t *packedDNA,
                               const uint32_t *sa, int device, s3_index **out)
{
    if (!bwt || !occ || !revBwt || !revOcc || !out || textLength == 0) {
        s3_set_error("s3_index_upload: NULL array or empty text");
        return S3_EINVAL;
    }
    if (numOcc != (textLength + 127) / 128 + 1) {
        s3_set_error("s3_index_upload: numOcc %u does not match textLength %u (BGS-Build.cpp:141)", numOcc, textLength);
        return S3_EINVAL;
    }
    int ndev = s3_device_count();
    if (device < 0 || device >= ndev) {
        s3_set_error("s3_index_upload: CUDA device %d not available (%d devices); there is no CPU fallback", device, ndev);
        return S3_ECUDA;
    }
    S3_CUDA(cudaSetDevice(device));
    if (const char *g = getenv("S3_L2_FETCH_GRANULARITY")) {              // measurement knob (profiles/): 32 / 64 / 128
        size_t got = 0;
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "[soap3dp_b200] cudaLimitMaxL2FetchGranularity asked %s, is %zu\n", g, got);
    }
    s3_index *ix = (s3_index *)calloc(1, sizeof(s3_index));
    if (!ix) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    ix->device = device;
    ix->textLength = textLength;
    S3_CUDA_IX(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    S3_CUDA_IX(cudaMallocHost(&ix->pinnedCount, 64));
    {   // stream-ordered allocations (cudaMallocAsync) keep their memory between calls instead of returning it at every synchronize
        cudaMemPool_t pool;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    S3_CUDA_IX(cudaDeviceGetAttribute(&ix->numSms, cudaDevAttrMultiProcessorCount, device));
    ix->splitBudget = 256;
    S3_CUDA_IX(cudaMalloc(&ix->d_workCounter, 256));
    ix->searchSmem = (size_t)-1;
    int rc;
    uint32_t nb = 0;
    if ((rc = upload_half(ix, bwt, occ, numOcc, textLength, &ix->d_fwd, &nb)) != S3_OK) { s3_index_free(ix); return rc; }
    ix->fwd.buckets = ix->d_fwd; ix->fwd.inverseSa0 = inverseSa0; ix->fwd.numBuckets = nb;
    if ((rc = upload_half(ix, revBwt, revOcc, numOcc, textLength, &ix->d_rev, &nb)) != S3_OK) { s3_index_free(ix); return rc; }
    ix->rev.buckets = ix->d_rev; ix->rev.inverseSa0 = revInverseSa0; ix->rev.numBuckets = nb;
    if ((rc = build_seed_tables(ix)) != S3_OK) { s3_index_free(ix); return rc; }
    if (packedDNA && sa && (rc = attach_locate(ix, sa, packedDNA, cudaMemcpyHostToDevice)) != S3_OK) { s3_index_free(ix); return rc; }
    *out = ix;
    return S3_OK;
}

// ---- the reference's index files, straight from disk ------------------------------------------------------------------
// s3_index_load maps <prefix>.bwt, .fmv.gpu, .rev.bwt, .rev.fmv.gpu (and .sa, .pac for the text side) as soap3-dp-builder +
// BGS-Build write them -- a 5-word header (inverseSa0, cumulative frequencies of A C G T; 2bwt-lib/BWT.c:170-200,
// BGS-Build.cpp:139-160), then the payload; .sa: one more word, the sampling interval, which must be 1 (BWT.c:225-285); .pac: four
// bases per byte in n / 4 + 1 bytes, then one byte = n % 4, the bases the byte before it holds (TextConverter.c:666-720) -- and hands the mappings to
// s3_index_upload: no copy of the files is made on the host (the packed text, 0.25 byte per base, is the exception: its bytes are
// turned into the big-endian words of hsp->packedDNA).  The mappings are page-locked for the upload when the driver allows it
// (cudaHostRegister of a private mapping), so that the copies run at the link's rate.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string>

namespace {
struct S3Mapped {
    const uint32_t *words; size_t bytes; int registered;
    S3Mapped() : words(NULL), bytes(0), registered(0) {}
};
int map_file(const std::string &path, S3Mapped *m)
{
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { s3_set_error("s3_index_load: cannot open %s", path.c_str()); return S3_EINVAL; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 20) { close(fd); s3_set_error("s3_index_load: %s is too short", path.c_str()); return S3_EINVAL; }
    // a private, writable mapping (never written: no page is ever copied): the driver page-locks those, read-only file mappings it refuses
    void *p = mmap(NULL, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { s3_set_error("s3_index_load: mmap of %s failed", path.c_str()); return S3_ENOMEM; }
    m->words = (const uint32_t *)p; m->bytes = (size_t)st.st_size;
    if (cudaHostRegister(p, m->bytes, cudaHostRegisterPortable) == cudaSuccess) m->registered = 1;
    else cudaGetLastError();                          // not page-locked: the copies still work, staged by the driver
    if (getenv("S3_INDEX_LOAD_VERBOSE")) fprintf(stderr, "[s3_index_load] %s: %zu bytes mapped, %s\n", path.c_str(), m->bytes, m->registered ? "page-locked" : "not page-locked");
    return S3_OK;
}
void unmap_file(S3Mapped *m)
{
    if (!m->words) return;
    if (m->registered) cudaHostUnregister((void *)m->words);
    munmap((void *)m->words, m->bytes);
    *m = S3Mapped();
}
}  // namespace

extern "C" int s3_index_load(const char *prefix, int withText, int device, s3_index **out)
{
    if (!prefix || !out) { s3_set_error("s3_index_load: NULL argument"); return S3_EINVAL; }
    *out = NULL;
    if (device < 0 || device >= s3_device_count()) { s3_set_error("s3_index_load: CUDA device %d not available; there is no CPU fallback", device); return S3_ECUDA; }
    if (cudaSetDevice(device) != cudaSuccess) { s3_set_error("s3_index_load: cudaSetDevice failed"); return S3_ECUDA; }
    const std::string pre(prefix);
    S3Mapped f[6];               // bwt, fmv.gpu, rev.bwt, rev.fmv.gpu, sa, pac
    const char *ext[6] = {".bwt", ".fmv.gpu", ".rev.bwt", ".rev.fmv.gpu", ".sa", ".pac"};
    uint32_t *text = NULL;
    int rc = S3_OK;
    for (int i = 0; i < (withText ? 6 : 4) && rc == S3_OK; ++i) rc = map_file(pre + ext[i], &f[i]);
    if (rc == S3_OK) {
        const uint32_t n = f[0].words[4];
        const size_t numWords = ((size_t)n + 15) / 16, numOcc = ((size_t)n + 127) / 128 + 1;
        // the two files of a direction carry the same header; both directions index texts of one length
        if (memcmp(f[0].words, f[1].words, 20) || memcmp(f[2].words, f[3].words, 20) || f[2].words[4] != n || n == 0) {
            s3_set_error("s3_index_load: the headers of %s.* do not agree", prefix); rc = S3_EINVAL;
        } else if (f[0].bytes < 20 + numWords * 4 || f[2].bytes < 20 + numWords * 4 || f[1].bytes < 20 + numOcc * 16 || f[3].bytes < 20 + numOcc * 16) {
            s3_set_error("s3_index_load: a file of %s.* is shorter than its header says", prefix); rc = S3_EINVAL;
        }
        const uint32_t *sa = NULL;
        if (rc == S3_OK && withText) {
            if (memcmp(f[4].words, f[0].words, 20) || f[4].bytes < 24 + ((size_t)n + 1) * 4 || f[4].words[5] != 1) {
                s3_set_error("s3_index_load: %s.sa does not hold the full suffix array of this index (SaValueFreq must be 1)", prefix); rc = S3_EINVAL;
            } else sa = f[4].words + 6;
            const uint8_t *pac = (const uint8_t *)f[5].words;
            const size_t fileLen = f[5].bytes - 1, textLen = fileLen ? (fileLen - 1) * 4 + pac[fileLen] : 0;
            if (rc == S3_OK && textLen != n) { s3_set_error("s3_index_load: %s.pac holds %zu bases, the index %u", prefix, textLen, n); rc = S3_EINVAL; }
            if (rc == S3_OK) {
                text = (uint32_t *)calloc(numWords + 8, 4);
                if (!text) { s3_set_error("s3_index_load: out of host memory"); rc = S3_ENOMEM; }
                else for (size_t w = 0; w < numWords; ++w) {
                    uint32_t v = 0;
                    for (int b = 0; b < 4; ++b) { const size_t k = 4 * w + b; v = (v << 8) | (k < fileLen ? pac[k] : 0u); }
                    text[w] = v;
                }
            }
        }
        if (rc == S3_OK)
            rc = s3_index_upload(f[0].words + 5, f[1].words + 5, f[2].words + 5, f[3].words + 5, (uint32_t)numOcc, f[0].words[0], f[2].words[0], n,
                                 withText ? text : NULL, withText ? sa : NULL, device, out);
    }
    free(text);
    for (int i = 0; i < 6; ++i) unmap_file(&f[i]);
    return rc;
}

extern "C" int s3_index_set_locate_device(s3_index *ix, const uint32_t *d_sa, const uint32_t *d_packedDNA)
{
    if (!ix || !d_sa || !d_packedDNA) { s3_set_error("s3_index_set_locate_device: NULL argument"); return S3_EINVAL; }
    S3_CUDA(cudaSetDevice(ix->device));
    return attach_locate(ix, d_sa, d_packedDNA, cudaMemcpyDeviceToDevice);
}

extern "C" void s3_index_free(s3_index *ix)
{
    if (!ix) return;
    cudaSetDevice(ix->device);
    cudaStreamSynchronize(ix->stream);
    if (!ix->sharedArrays) {
        cudaFree(ix->d_fwd); cudaFree(ix->d_rev);
        for (int t = 0; t < 3; ++t) if (ix->d_seed[t]) cudaFree(ix->d_seed[t]);
        if (ix->d_packedDNA) cudaFree(ix->d_packedDNA);
        if (ix->d_sa) cudaFree(ix->d_sa);
        if (ix->d_isa) cudaFree(ix->d_isa);
    }
    s3_stage_ws_free(ix);
    s3_pipe_destroy(&ix->pipe);
    if (ix->d_workCounter) cudaFree(ix->d_workCounter);
    if (ix->d_hardItems) cudaFree(ix->d_hardItems);
    if (ix->d_itemStats) cudaFree(ix->d_itemStats);
    if (ix->d_heavy) cudaFree(ix->d_heavy);
    if (ix->side.ready) {
        cudaStreamSynchronize(ix->side.stream);
        if (ix->side.d_workCounter) cudaFree(ix->side.d_workCounter);
        if (ix->side.d_hardItems) cudaFree(ix->side.d_hardItems);
        if (ix->side.d_heavy) cudaFree(ix->side.d_heavy);
        cudaEventDestroy(ix->side.fork); cudaEventDestroy(ix->side.join);
        cudaStreamDestroy(ix->side.stream);
    }
    s3_timing_destroy(&ix->timing);
    if (ix->scratch) cudaFree(ix->scratch);
    if (ix->pinned) cudaFreeHost(ix->pinned);
    if (ix->pinnedCount) cudaFreeHost(ix->pinnedCount);
    cudaStreamDestroy(ix->stream);
    free(ix);
}

// A second handle on the same device arrays: its own stream, work queues and scratch, so that two batches can be in
// flight at once (one host thread each) -- the memory-bound search of one under the compute-bound DP of the other.  The
// reference gets the same overlap from its two engine threads (main thread searching batch k + 1 while a DP engine's GPU
// thread aligns batch k).  Free the clones before the handle they were made from.
extern "C" int s3_index_clone(s3_index *ix, s3_index **out)
{
    if (!ix || !out) { s3_set_error("s3_index_clone: NULL argument"); return S3_EINVAL; }
    S3_CUDA(cudaSetDevice(ix->device));
    s3_index *c = (s3_index *)calloc(1, sizeof(s3_index));
    if (!c) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    c->device = ix->device; c->fwd = ix->fwd; c->rev = ix->rev; c->seed = ix->seed; c->loc = ix->loc;
    for (int t = 0; t < 3; ++t) c->d_seed[t] = ix->d_seed[t];
    c->d_isa = ix->d_isa; c->textLength = ix->textLength; c->d_fwd = ix->d_fwd; c->d_rev = ix->d_rev;
    c->d_packedDNA = ix->d_packedDNA; c->d_sa = ix->d_sa; c->bytes = 0;
    c->numSms = ix->numSms; c->splitBudget = ix->splitBudget; c->searchSmem = (size_t)-1; c->sharedArrays = 1;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&c->d_workCounter, 256) != cudaSuccess ||
        cudaMallocHost(&c->pinnedCount, 64) != cudaSuccess) {
        s3_set_error("s3_index_clone: %s", cudaGetErrorString(cudaGetLastError()));
        s3_index_free(c);
        return S3_ECUDA;
    }
    *out = c;
    return S3_OK;
}

extern "C" size_t s3_index_device_bytes(const s3_index *ix) { return ix ? ix->bytes : 0; }
extern "C" void *s3_index_stream(const s3_index *ix) { return ix ? (void *)ix->stream : NULL; }

__global__ void s3_rank_probe_kernel(S3Half h, const uint32_t *__restrict__ idx, size_t n, uint32_t *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[4];
    s3_rank4(h, idx[i], r);
    reinterpret_cast<uint4 *>(out)[i] = make_uint4(r[0], r[1], r[2], r[3]);
}

extern "C" int s3_rank_probe(s3_index *ix, int which, const uint32_t *indices, size_t n, uint32_t *out)
{
    if (!ix || !indices || !out) { s3_set_error("s3_rank_probe: NULL argument"); return S3_EINVAL; }
    if (n == 0) return S3_OK;
    S3_CUDA(cudaSetDevice(ix->device));
    uint32_t *d_idx, *d_out;
    S3_CUDA(cudaMalloc(&d_idx, n * 4));
    S3_CUDA(cudaMalloc(&d_out, n * 16));
    S3_CUDA(cudaMemcpyAsync(d_idx, indices, n * 4, cudaMemcpyHostToDevice, ix->stream));
    s3_rank_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(which ? ix->rev : ix->fwd, d_idx, n, d_out);
    S3_CUDA(cudaGetLastError());
    S3_CUDA(cudaMemcpyAsync(out, d_out, n * 16, cudaMemcpyDeviceToHost, ix->stream));
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    cudaFree(d_idx); cudaFree(d_out);
    return S3_OK;
}

// ---- measurement: the memory system's random-sector rate ---------------------------------------------------------
// Independent random 32-byte reads (one LDG.E.256 each, the search's own access) over the forward bucket array, four in
// flight per thread: what bench.py reports the search launch's executed sectors against (the search is
// bound by random sectors, not by streaming bandwidth).  Replaces nothing in the reference.
template <int U>
__global__ void s3_random_sector_kernel(const uint4 *__restrict__ buckets, uint32_t numBuckets, uint32_t loads, uint32_t *__restrict__ sink)
{
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < loads; k += U) {
        uint32_t v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            x = x * 1664525u + 1013904223u;
            const uint4 *p = buckets + (size_t)(((unsigned long long)x * numBuckets) >> 32) * 2;
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]), "=r"(v[u][4]), "=r"(v[u][5]), "=r"(v[u][6]), "=r"(v[u][7]) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u][0] ^ v[u][7];
        x ^= acc & 1u;                      // the next addresses depend on nothing that waits: acc & 1 only keeps the loads alive
    }
    if (acc == 0x9E3779B9u) sink[0] = acc;
}

// Several shapes are tried (loads in flight per thread x threads per block) and the best rate is reported: the probe is
// meant to be a ceiling for kernels whose own access pattern this is.
extern "C" int s3_random_sector_probe(s3_index *ix, uint32_t loadsPerThread, float *ms, uint64_t *numLoads)
{
    if (!ix || !ms || !numLoads || loadsPerThread == 0) { s3_set_error("s3_random_sector_probe: bad argument"); return S3_EINVAL; }
    S3_CUDA(cudaSetDevice(ix->device));
    loadsPerThread = (loadsPerThread + 15) / 16 * 16;
    uint32_t *d_sink;
    S3_CUDA(cudaMalloc(&d_sink, 4));
    cudaEvent_t e0, e1;
    S3_CUDA(cudaEventCreate(&e0)); S3_CUDA(cudaEventCreate(&e1));
    double bestRate = 0;
    for (int shape = 0; shape < 6; ++shape) {
        const int U = shape % 3 == 0 ? 4 : shape % 3 == 1 ? 8 : 16, threads = shape < 3 ? 256 : 512;
        const unsigned blocks = (unsigned)ix->numSms * (shape < 3 ? 64 : 32);
        S3_CUDA(cudaEventRecord(e0, ix->stream));
        if (U == 4) s3_random_sector_kernel<4><<<blocks, threads, 0, ix->stream>>>(ix->d_fwd, ix->fwd.numBuckets, loadsPerThread, d_sink);
        else if (U == 8) s3_random_sector_kernel<8><<<blocks, threads, 0, ix->stream>>>(ix->d_fwd, ix->fwd.numBuckets, loadsPerThread, d_sink);
        else s3_random_sector_kernel<16><<<blocks, threads, 0, ix->stream>>>(ix->d_fwd, ix->fwd.numBuckets, loadsPerThread, d_sink);
        S3_LAUNCHED(1);
        S3_CUDA(cudaEventRecord(e1, ix->stream));
        S3_CUDA(cudaStreamSynchronize(ix->stream));
        float t = 0;
        S3_CUDA(cudaEventElapsedTime(&t, e0, e1));
        const uint64_t n = (uint64_t)blocks * threads * loadsPerThread;
        if ((double)n / t > bestRate) { bestRate = (double)n / t; *ms = t; *numLoads = n; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_sink);
    return S3_OK;
}

// ---- L2 access-policy window (measurement knob) --------------------------------------------------------------------
// Pins a window of one of the index arrays in the persisting part of L2 for the kernels launched on the index stream.
// region: 0 none (resets), 1 forward buckets, 2 reverse buckets, 3 seed table fwd1, 4 seed table rev0, 5 packed text,
// 6 suffix array.  windowBytes = 0: as much of the array as the device's largest window allows.  The window's hit ratio
// is set so that the persisting carve-out (persistBytes, capped at the device's maximum) is not over-subscribed.
// Nothing in the index is hot -- reads are uniform over the genome, the seed tables replace the top of the BWT -- so this
// exists to MEASURE that (profiles/), not because the search relies on it.
extern "C" int s3_index_set_l2_persist(s3_index *ix, int region, size_t windowBytes, size_t persistBytes)
{
    if (!ix) { s3_set_error("s3_index_set_l2_persist: NULL index"); return S3_EINVAL; }
    S3_CUDA(cudaSetDevice(ix->device));
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (region == 0) {
        attr.accessPolicyWindow.num_bytes = 0;
        S3_CUDA(cudaStreamSetAttribute(ix->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        S3_CUDA(cudaCtxResetPersistingL2Cache());
        S3_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
        return S3_OK;
    }
    const size_t n = ix->textLength;
    const size_t seedEntries = ix->seed.K ? ((size_t)1 << (2 * ix->seed.K)) : 0;
    void *base = NULL; size_t bytes = 0;
    switch (region) {
    case 1: base = ix->d_fwd; bytes = (size_t)ix->fwd.numBuckets * 32; break;
    case 2: base = ix->d_rev; bytes = (size_t)ix->rev.numBuckets * 32; break;
    case 3: base = ix->d_seed[1]; bytes = seedEntries * 8; break;
    case 4: base = ix->d_seed[2]; bytes = seedEntries * 8; break;
    case 5: base = ix->d_packedDNA; bytes = (n + 15) / 16 * 4; break;
    case 6: base = ix->d_sa; bytes = (n + 1) * 4; break;
    default: s3_set_error("s3_index_set_l2_persist: region %d", region); return S3_EINVAL;
    }
    if (!base) { s3_set_error("s3_index_set_l2_persist: the index does not hold region %d", region); return S3_EINVAL; }
    int maxWindow = 0, maxPersist = 0;
    S3_CUDA(cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, ix->device));
    S3_CUDA(cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ix->device));
    if (windowBytes == 0 || windowBytes > bytes) windowBytes = bytes;
    if (windowBytes > (size_t)maxWindow) windowBytes = (size_t)maxWindow;
    if (persistBytes == 0 || persistBytes > (size_t)maxPersist) persistBytes = (size_t)maxPersist;
    S3_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persistBytes));
    attr.accessPolicyWindow.base_ptr = base;
    attr.accessPolicyWindow.num_bytes = windowBytes;
    attr.accessPolicyWindow.hitRatio = windowBytes <= persistBytes ? 1.0f : (float)((double)persistBytes / (double)windowBytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    S3_CUDA(cudaStreamSetAttribute(ix->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return S3_OK;
}
