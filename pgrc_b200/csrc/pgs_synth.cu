// pgs_synth.cu — counter-based synthetic matcher inputs, bit-identical on the host and on the device.
//
// TEST / BENCH INFRASTRUCTURE (not part of the product library): built into pgrc_b200/libpgrc_synth.so.
// Every output byte is a pure function of (seed, index), so
//   * any sub-range of the text or of the reads can be generated on its own (a rank generates what it needs),
//   * the container without a GPU (OpenMP loop) and the B200 box (kernel) produce the same workload, which is what
//     lets tests/golden/ hold results the reference / the oracle computed HERE for inputs the GPU box re-creates.
// Shapes follow SURVEY.md §8(d): uniform-random genome; the pseudogenome-like text is `copies` concatenated passes
// over the genome cut into contigs (grid of `contig` bases with hashed jitter), each contig in a hashed orientation;
// reads start uniformly, half of them reverse-complemented, i.i.d. substitutions with probability err_q24 / 2^24 and
// at least one per read (the matcher only sees error-containing reads); packed 4 bases per byte, first base in the
// two most significant bits (SymbolsPackingFacility.cpp:168-185).
#include <cuda_runtime.h>
#include <stdint.h>

#define PGS_HD __host__ __device__ __forceinline__

struct pgs_params {
    uint64_t seed;
    uint64_t genome_len;
    uint64_t text_len;      // = full copies * genome_len + tail span
    uint32_t contig;        // grid of the contig cuts
    uint32_t read_len;
    uint32_t err_q24;       // substitution probability * 2^24
    uint32_t reserved;
};

namespace {

PGS_HD uint64_t mix(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
PGS_HD uint64_t h(uint64_t seed, uint64_t stream, uint64_t i) { return mix(mix(seed + stream * 0x9E3779B97F4A7C15ull) ^ (i * 0xD6E8FEB86659FD93ull)); }

PGS_HD uint32_t genome_code(const pgs_params &p, uint64_t i) { return (uint32_t)(h(p.seed, 0, i) >> 17) & 3u; }

// start of contig k of one pass over [0, span): grid position + jitter (contig 0 starts at 0)
PGS_HD uint64_t contig_start(const pgs_params &p, uint64_t pass, uint64_t k, uint64_t span) {
    if (k == 0) return 0;
    const uint64_t s = k * p.contig + h(p.seed, 16 + pass, k) % (p.contig / 2 + 1);
    return s < span ? s : span;
}

// code of text position q
PGS_HD uint32_t text_code(const pgs_params &p, uint64_t q) {
    const uint64_t G = p.genome_len;
    const uint64_t pass = q / G, off = q - pass * G;
    const uint64_t full = p.text_len / G;
    uint64_t span = G, lo = 0;
    if (pass >= full) {                 // the fractional last pass covers [lo, lo + span) of the genome
        span = p.text_len - full * G;
        lo = h(p.seed, 8, pass) % (G - span + 1);
    }
    uint64_t k = off / p.contig;
    if (off < contig_start(p, pass, k, span)) k--;
    const uint64_t s = contig_start(p, pass, k, span);
    uint64_t e = contig_start(p, pass, k + 1, span);
    if ((k + 1) * (uint64_t)p.contig >= span) e = span;
    const bool flip = (h(p.seed, 32 + pass, k) >> 20) & 1;
    const uint64_t src = flip ? s + (e - 1 - off) : off;
    const uint32_t c = genome_code(p, lo + src);
    return flip ? 3u - c : c;
}

PGS_HD void read_packed(const pgs_params &p, uint64_t r, uint8_t *dst) {
    const uint32_t L = p.read_len, plen = (L + 3) / 4;
    const uint64_t start = h(p.seed, 2, r) % (p.genome_len - L + 1);
    const bool flip = (h(p.seed, 3, r) >> 20) & 1;
    const uint32_t forced = (uint32_t)(h(p.seed, 5, r) % L);
    bool any = false;
    for (uint32_t b = 0; b < L; b++) any |= ((uint32_t)(h(p.seed, 4, r * 256 + b) >> 16) & 0xFFFFFFu) < p.err_q24;
    for (uint32_t byte = 0; byte < plen; byte++) {
        uint32_t v = 0;
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t b = byte * 4 + k;
            uint32_t c = 0;
            if (b < L) {
                c = flip ? 3u - genome_code(p, start + (L - 1 - b)) : genome_code(p, start + b);
                const uint64_t e = h(p.seed, 4, r * 256 + b);
                const bool sub = (((uint32_t)(e >> 16) & 0xFFFFFFu) < p.err_q24) || (!any && b == forced);
                if (sub) c = (c + 1u + (uint32_t)((e >> 44) % 3u)) & 3u;
            }
            v = (v << 2) | c;
        }
        dst[byte] = (uint8_t)v;
    }
}

__global__ void text_kernel(pgs_params p, uint8_t *out, uint64_t begin, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (uint8_t)((0x54474341u >> (8 * text_code(p, begin + i))) & 0xFF);
}
__global__ void reads_kernel(pgs_params p, uint8_t *out, uint64_t first, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) read_packed(p, first + i, out + i * ((p.read_len + 3) / 4));
}

bool is_device_ptr(const void *ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

} // namespace

extern "C" {

// ASCII text [begin, begin + count) into `out` (host or device memory of the current device); 0 = ok
int pgs_text(const pgs_params *p, uint8_t *out, uint64_t begin, uint64_t count) {
    if (!p || (!out && count) || begin + count > p->text_len || p->contig < 2 || p->genome_len < p->read_len + 1) return -1;
    if (!count) return 0;
    if (is_device_ptr(out)) {
        const uint64_t step = 1ull << 30;
        for (uint64_t o = 0; o < count; o += step) {
            const uint64_t n = count - o < step ? count - o : step;
            text_kernel<<<(unsigned)((n + 255) / 256), 256>>>(*p, out + o, begin + o, n);
        }
        return cudaGetLastError() == cudaSuccess ? 0 : -2;
    }
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)count; i++) out[i] = (uint8_t)((0x54474341u >> (8 * text_code(*p, begin + (uint64_t)i))) & 0xFF);
    return 0;
}

// packed reads [first, first + count) into `out` (count * ceil(read_len / 4) bytes)
int pgs_reads(const pgs_params *p, uint8_t *out, uint64_t first, uint64_t count) {
    if (!p || (!out && count) || p->read_len == 0 || p->read_len > 255 || p->genome_len < p->read_len + 1) return -1;
    if (!count) return 0;
    if (is_device_ptr(out)) {
        const uint64_t step = 1ull << 26;
        const uint32_t plen = (p->read_len + 3) / 4;
        for (uint64_t o = 0; o < count; o += step) {
            const uint64_t n = count - o < step ? count - o : step;
            reads_kernel<<<(unsigned)((n + 127) / 128), 128>>>(*p, out + o * plen, first + o, n);
        }
        return cudaGetLastError() == cudaSuccess ? 0 : -2;
    }
    const uint32_t plen = (p->read_len + 3) / 4;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)count; i++) read_packed(*p, first + (uint64_t)i, out + (uint64_t)i * plen);
    return 0;
}

} // extern "C"
