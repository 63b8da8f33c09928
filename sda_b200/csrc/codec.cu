// codec.cu -- the share wire codec either side of the hot path:
//   encode  client/src/crypto/encryption/sodium.rs:35-41   for share in shares { share.encode_var(..) }
//   decode  client/src/crypto/encryption/sodium.rs:83-90   while !reader.is_empty() { Share::decode_var(reader) }
// with `integer-encoding 1.0` semantics: zig-zag (v << 1) ^ (v >> 63), then unsigned LEB128,
// 7 bits per byte, least significant group first, MSB = continuation; values are concatenated with no length
// prefix.  A 61-bit share takes 9 bytes, an arbitrary i64 at most 10.
//
// Both directions are variable-length, so each CTA needs the total of every chunk before its own.  One launch per
// direction: a CTA takes the next chunk (a ticket, so that every earlier chunk is already running), sizes it, publishes
// its total, and gets its offset by looking back over its predecessors' published totals / running prefixes
// (`chunk_offset`, a chained scan); the input is read once.  The byte streams move through shared memory so that
// global traffic is 16-byte vectors in both directions whatever the byte alignment of a chunk's first value:
//   encode: read 8 n (i64), write <= 10 n bytes;   decode: read the bytes, write 8 n.
// The sealed box around the encoded bytes (libsodium) is out of scope and stays on the CPU.
#include "kernels.h"

namespace sda {

namespace {

constexpr int CTA = 256;
constexpr int EPT = 8;                          // encode: elements per thread
constexpr int ECHUNK = CTA * EPT;               // elements per CTA
constexpr int EBYTES = ECHUNK * 10 + 32;        // staging for a chunk's bytes (+ alignment phase + vector tail)
constexpr int DPT = 16;                         // decode: bytes per thread
constexpr int DCHUNK = CTA * DPT;               // bytes per CTA
constexpr int HALO = 16;                        // bytes before a chunk a value ending in it may start in (>= 9)

__device__ __forceinline__ uint64_t zigzag(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
__device__ __forceinline__ int varint_len(uint64_t z) { return (70 - __clzll(z | 1)) / 7; }   // ceil(bits / 7), >= 1

// block-wide exclusive scan of one value per thread (CTA = 256); returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t *total) {
    __shared__ uint32_t warp_sums[CTA / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < CTA / 32; w++) {
        const uint32_t s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    *total = tot;
    __syncthreads();
    return base + incl - x;
}

// ---- chained scan over chunks -------------------------------------------------------------------------------
// scratch: state[0 .. chunks) (zeroed before the launch), then the ticket counter (zeroed), then the grand total.
// A state word is a 62-bit count tagged TOTAL (this chunk alone) or PREFIX (this chunk and everything before it).
constexpr uint64_t TAG_TOTAL = 1ull << 62, TAG_PREFIX = 2ull << 62, COUNT_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_state(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// the next chunk in launch order (uniform across the CTA)
__device__ __forceinline__ uint32_t take_chunk(uint64_t *ticket) {
    __shared__ uint32_t chunk_s;
    if (threadIdx.x == 0) chunk_s = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long *>(ticket), 1ull);
    __syncthreads();
    return chunk_s;
}

// sum of the totals of chunks 0 .. chunk-1, given this chunk's own; called by every thread, the look-back runs in warp 0.
// The last chunk leaves the grand total in *grand.
__device__ __forceinline__ uint64_t chunk_offset(uint64_t *state, uint32_t chunk, uint32_t chunks, uint32_t my_total, uint64_t *grand) {
    __shared__ uint64_t offset_s;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0 && chunk > 0) st_state(state + chunk, TAG_TOTAL | my_total);
        uint64_t prefix = 0;
        for (int64_t first = (int64_t)chunk - 1; first >= 0; first -= 32) {
            const int64_t j = first - lane;
            uint64_t v;
            do {
                v = j >= 0 ? ld_state(state + j) : TAG_PREFIX;          // before chunk 0: an empty prefix
            } while (__any_sync(0xffffffffu, (v >> 62) == 0));
            const unsigned prefixes = __ballot_sync(0xffffffffu, (v >> 62) == 2);
            const int stop = prefixes ? __ffs(prefixes) - 1 : 31;        // the nearest chunk that already knows its prefix
            uint64_t part = lane <= stop ? (v & COUNT_MASK) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            prefix += part;
            if (prefixes) break;
        }
        if (lane == 0) {
            st_state(state + chunk, TAG_PREFIX | (prefix + my_total));
            offset_s = prefix;
            if (chunk + 1 == chunks) *grand = prefix + my_total;
        }
    }
    __syncthreads();
    return offset_s;
}

// ---- encode -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA)
varint_encode_kernel(const int64_t *__restrict__ in, size_t n, uint32_t chunks, uint64_t *state, uint8_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t stage[EBYTES];
    const uint32_t chunk = take_chunk(state + chunks);
    const size_t e0 = (size_t)chunk * ECHUNK + (size_t)threadIdx.x * EPT;
    uint64_t z[EPT];
    uint32_t bytes = 0;
#pragma unroll
    for (int i = 0; i < EPT; i++) {
        z[i] = e0 + i < n ? zigzag(__ldg(in + e0 + i)) : 0;
        if (e0 + i < n) bytes += varint_len(z[i]);
    }
    uint32_t total;
    const uint32_t local = block_exclusive_scan(bytes, &total);
    const uint64_t goff = chunk_offset(state, chunk, chunks, total, state + chunks + 1);
    const uint32_t phase = (uint32_t)((uintptr_t)(out + goff) & 15);    // stage with the destination's 16-byte phase
    uint32_t w = phase + local;
#pragma unroll
    for (int i = 0; i < EPT; i++) {
        if (e0 + i < n) {
            uint64_t v = z[i];
            while (v >= 0x80) {
                stage[w++] = (uint8_t)(v | 0x80);
                v >>= 7;
            }
            stage[w++] = (uint8_t)v;
        }
    }
    __syncthreads();
    // copy stage[phase .. phase + total) to out[goff ..): unaligned head and tail by bytes, the middle as uint4
    uint8_t *dst = out + goff - phase;                                   // 16-byte aligned
    const uint32_t end = phase + total;
    const uint32_t v0 = phase ? 16 : 0, v1 = end & ~15u;                 // vector range [v0, v1)
    if (v1 > v0) {
        for (uint32_t o = v0 + threadIdx.x * 16; o < v1; o += CTA * 16)
            *reinterpret_cast<uint4 *>(dst + o) = *reinterpret_cast<const uint4 *>(stage + o);
        for (uint32_t o = phase + threadIdx.x; o < v0 && o < end; o += CTA) dst[o] = stage[o];
        for (uint32_t o = v1 + threadIdx.x; o < end; o += CTA) dst[o] = stage[o];
    } else {
        for (uint32_t o = phase + threadIdx.x; o < end; o += CTA) dst[o] = stage[o];
    }
}

// ---- decode -------------------------------------------------------------------------------------
// a byte with the MSB clear ends a value: values = number of such bytes
__device__ __forceinline__ uint32_t terminators16(const uint4 w) {
    return __popc(~w.x & 0x80808080u) + __popc(~w.y & 0x80808080u) + __popc(~w.z & 0x80808080u) + __popc(~w.w & 0x80808080u);
}

// chunk bytes [c * DCHUNK, ..) with HALO bytes before them, staged at stage[HALO + i]; bytes outside the stream
// read as 0x00 before it (a terminator: the first value starts at 0) and are never consumed after it
__device__ __forceinline__ void stage_chunk(const uint8_t *__restrict__ buf, size_t len, size_t c0, uint8_t *stage) {
    const bool aligned = ((uintptr_t)buf & 15) == 0;
    const size_t o = c0 + (size_t)threadIdx.x * DPT;
    if (aligned && o + DPT <= len) {
        *reinterpret_cast<uint4 *>(stage + HALO + threadIdx.x * DPT) = __ldg(reinterpret_cast<const uint4 *>(buf + o));
    } else {
#pragma unroll
        for (int i = 0; i < DPT; i++) stage[HALO + threadIdx.x * DPT + i] = o + i < len ? __ldg(buf + o + i) : 0x80;
    }
    if (threadIdx.x < HALO) stage[threadIdx.x] = c0 >= HALO - threadIdx.x ? __ldg(buf + c0 - HALO + threadIdx.x) : 0x00;
}

// 7-bit groups of up to 10 little-endian bytes (lo = bytes 0..7, hi = bytes 8..9) -> 64-bit integer
__device__ __forceinline__ uint64_t squeeze_leb128(uint64_t lo, uint32_t hi) {
    uint64_t x = lo & 0x7f7f7f7f7f7f7f7full;
    x = ((x & 0x7f007f007f007f00ull) >> 1) | (x & 0x007f007f007f007full);      // 14 bits per 16-bit lane
    x = ((x & 0x3fff00003fff0000ull) >> 2) | (x & 0x00003fff00003fffull);      // 28 bits per 32-bit lane
    x = ((x & 0x0fffffff00000000ull) >> 4) | (x & 0x000000000fffffffull);      // 56 bits
    return x | ((uint64_t)(hi & 0x7f) << 56) | ((uint64_t)((hi >> 8) & 1) << 63);
}

// status bit 0: a value longer than 10 bytes; bit 1: the stream ends inside a value; bit 2: more values than `cap`
//
// Two phases per chunk so that neither diverges: every thread lists the terminator positions of its 16 bytes
// (prefix positions from a block scan), then thread v decodes value v of the chunk from a 10-byte window.
__global__ void __launch_bounds__(CTA)
varint_decode_kernel(const uint8_t *__restrict__ buf, size_t len, uint32_t chunks, uint64_t *state, size_t cap,
                     int64_t *__restrict__ out, unsigned *status) {
    __shared__ __align__(16) uint8_t stage[HALO + DCHUNK + 16];
    __shared__ uint16_t ends[DCHUNK];                       // chunk-relative position of every terminator, in order
    const uint32_t cid = take_chunk(state + chunks);
    const size_t c0 = (size_t)cid * DCHUNK;
    stage_chunk(buf, len, c0, stage);
    if (threadIdx.x < 16) stage[HALO + DCHUNK + threadIdx.x] = 0x80;
    __syncthreads();
    const uint8_t *chunk = stage + HALO;
    const uint4 w = *reinterpret_cast<const uint4 *>(chunk + threadIdx.x * DPT);
    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
    uint32_t mask = 0;                                      // bit i: byte i of my 16 ends a value
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t t = ~words[q] & 0x80808080u;
        mask |= (((t >> 7) & 1) | ((t >> 14) & 2) | ((t >> 21) & 4) | ((t >> 28) & 8)) << (4 * q);
    }
    uint32_t total;
    uint32_t slot = block_exclusive_scan(__popc(mask), &total);
    for (uint32_t m = mask; m; m &= m - 1) ends[slot++] = (uint16_t)(threadIdx.x * DPT + __ffs(m) - 1);
    __syncthreads();
    unsigned bad = 0;
    const uint64_t k0 = chunk_offset(state, cid, chunks, total, state + chunks + 1);     // values that end before this chunk
    for (uint32_t v = threadIdx.x; v < total; v += CTA) {
        const int end = ends[v];
        int start;
        if (v > 0) {
            start = ends[v - 1] + 1;
        } else {                                            // the first value may begin in the previous chunk
            start = end;
            while (end - start < 10 && (chunk[start - 1] & 0x80)) start--;
        }
        int n = end - start + 1;
        if (n > 10) {
            bad |= 1u;
            n = 10;
        }
        // unaligned 10-byte window at chunk[start]: three aligned words + the byte phase
        const uint8_t *a = chunk + start;
        const uint32_t *aw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(a) & ~(uintptr_t)3);
        const uint32_t sh = ((uint32_t)reinterpret_cast<uintptr_t>(a) & 3) * 8;
        const uint32_t w0 = aw[0], w1 = aw[1], w2 = aw[2], w3 = aw[3];
        uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
        // keep the n bytes of this value
        if (n < 4) b0 &= (1u << (8 * n)) - 1;
        if (n <= 4) b1 = 0;
        else if (n < 8) b1 &= (1u << (8 * (n - 4))) - 1;
        if (n <= 8) b2 = 0;
        else if (n == 9) b2 &= 0xffu;
        else b2 &= 0xffffu;
        const uint64_t z = squeeze_leb128(((uint64_t)b1 << 32) | b0, b2);
        const uint64_t k = k0 + v;
        if (k < cap) out[k] = (int64_t)(z >> 1) ^ -(int64_t)(z & 1);
        else bad |= 4u;
    }
    if (c0 + DCHUNK >= len && threadIdx.x == 0 && (chunk[len - 1 - c0] & 0x80)) bad |= 2u;   // this chunk holds the final byte
    if (bad) atomicOr(status, bad);
}

}  // namespace

// state per chunk, the ticket, the total
size_t varint_encode_scratch_elems(size_t n) { return (n + ECHUNK - 1) / ECHUNK + 2; }
size_t varint_decode_scratch_elems(size_t len) { return (len + DCHUNK - 1) / DCHUNK + 2; }

// out needs 10 n bytes; the encoded length lands in the last scratch element (device) for the caller to read back
cudaError_t launch_varint_encode(const LaunchCtx &lc, const int64_t *in, size_t n, uint8_t *out, uint64_t *scratch) {
    const size_t chunks = (n + ECHUNK - 1) / ECHUNK;
    if (chunks >> 31) return cudaErrorInvalidValue;
    const cudaError_t e = cudaMemsetAsync(scratch, 0, (chunks + 2) * sizeof(uint64_t), lc.stream);
    if (e != cudaSuccess || chunks == 0) return e;
    varint_encode_kernel<<<(unsigned)chunks, CTA, 0, lc.stream>>>(in, n, (uint32_t)chunks, scratch, out);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

// the value count lands in the last scratch element; *status (device, pre-cleared) gets the malformed-stream bits
cudaError_t launch_varint_decode(const LaunchCtx &lc, const uint8_t *buf, size_t len, int64_t *out, size_t cap,
                                 uint64_t *scratch, unsigned *status) {
    const size_t chunks = (len + DCHUNK - 1) / DCHUNK;
    if (chunks >> 31) return cudaErrorInvalidValue;
    const cudaError_t e = cudaMemsetAsync(scratch, 0, (chunks + 2) * sizeof(uint64_t), lc.stream);
    if (e != cudaSuccess || chunks == 0) return e;
    varint_decode_kernel<<<(unsigned)chunks, CTA, 0, lc.stream>>>(buf, len, (uint32_t)chunks, scratch, cap, out, status);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace sda
