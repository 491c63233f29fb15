// fasta.cu -- the FASTA front end of `builder` on the GPU (sm_100a).
//
// Replaces the host loop of the reference's build() (builder.cpp:203-262): getline per row, a row that
// starts with '>' closes the running record, every other row is appended to its sequence, and each
// non-empty sequence s becomes the document  reverse(s + '-' + revcomp(s)) = complement(s) + '-' +
// reverse(s)  (transform, builder.cpp:183-201) after normalize() (builder.cpp:60-104) has folded acgtn
// to upper case and turned everything outside ACGTN0123. into N.  The same rules, per byte and in
// parallel:
//   * a byte is a line start iff it is the first byte or follows '\n'; a line is a header line iff its
//     first byte is '>'; sequence bytes are the bytes of the other lines except '\n' ('\r' included:
//     the reference appends it to the sequence and then normalises it to N);
//   * S(i) = sequence bytes before i, R(i) = header lines starting at or before i: two prefix sums;
//     record k (k-th header, k = 0 for rows in front of the first header) owns the sequence bytes
//     [B[k], B[k+1]) with B[k] = S(header k);
//   * a record with L = B[k+1]-B[k] > 0 becomes a document of 2L+1 symbols plus its terminator at
//     offset 2*B[k] + 2*(non-empty records before k): a third prefix sum, over records.
// All passes stream the file with 16-byte loads; a tile is 4096 bytes.
#include "fasta.cuh"

namespace dsmfm {

namespace {

constexpr int kFaThreads = 256;
constexpr int kFaPer = 16;
constexpr int kFaTile = kFaThreads * kFaPer;
constexpr int kRecPer = 8; // records per thread in the record passes
constexpr int kRecTile = kFaThreads * kRecPer;

__device__ __forceinline__ int load16(const uint8_t *__restrict__ f, uint64_t i0, uint64_t m, uint8_t (&b)[kFaPer])
{
    if (i0 + kFaPer <= m) {
        const uint4 v = *reinterpret_cast<const uint4 *>(f + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < kFaPer; ++j) b[j] = (uint8_t)(w[j >> 2] >> ((j & 3) * 8));
        return kFaPer;
    }
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kFaPer; ++j) {
        b[j] = 0;
        if (i0 + j < m) {
            b[j] = f[i0 + j];
            cnt = j + 1;
        }
    }
    return cnt;
}

template <typename T> __device__ __forceinline__ T warp_max(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// Exclusive running maximum over the threads of the block (identity `none`).
__device__ __forceinline__ long long block_excl_max(long long v, long long none, long long *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o && t > incl) incl = t;
    }
    if (lane == 31) scratch[warp] = incl;
    long long prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev = none;
    __syncthreads();
    long long base = none;
    for (int w = 0; w < nw; ++w)
        if (w < warp && scratch[w] > base) base = scratch[w];
    __syncthreads();
    return prev > base ? prev : base;
}

// Exclusive block sums of two counters at once; totals of the block in tot[0], tot[1].
__device__ __forceinline__ void block_excl_sum2(uint32_t a, uint32_t b, uint32_t &ea, uint32_t &eb, uint32_t (*scratch)[2],
                                                uint32_t (&tot)[2])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) {
            ia += ta;
            ib += tb;
        }
    }
    if (lane == 31) {
        scratch[warp][0] = ia;
        scratch[warp][1] = ib;
    }
    __syncthreads();
    uint32_t ba = 0, bb = 0, sa = 0, sb = 0;
    for (int w = 0; w < nw; ++w) {
        if (w < warp) {
            ba += scratch[w][0];
            bb += scratch[w][1];
        }
        sa += scratch[w][0];
        sb += scratch[w][1];
    }
    __syncthreads();
    ea = ba + ia - a;
    eb = bb + ib - b;
    tot[0] = sa;
    tot[1] = sb;
}

// position of the last '\n' of every tile (-1: none)
__global__ void __launch_bounds__(kFaThreads) fa_last_nl_kernel(const uint8_t *__restrict__ f, uint64_t m,
                                                                long long *__restrict__ last_nl)
{
    __shared__ long long s_w[kFaThreads / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * kFaTile + (uint64_t)threadIdx.x * kFaPer;
    uint8_t b[kFaPer];
    const int cnt = load16(f, i0, m, b);
    long long last = -1;
#pragma unroll
    for (int j = 0; j < kFaPer; ++j)
        if (j < cnt && b[j] == '\n') last = (long long)(i0 + j);
    last = warp_max(last);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long v = -1;
        for (int w = 0; w < kFaThreads / 32; ++w) v = s_w[w] > v ? s_w[w] : v;
        last_nl[blockIdx.x] = v;
    }
}

// entry[t] = last '\n' in front of tile t (-1: none); one block
__global__ void __launch_bounds__(1024) fa_scan_max_kernel(const long long *__restrict__ last_nl, uint64_t ntiles,
                                                           long long *__restrict__ entry)
{
    __shared__ long long s_w[32];
    const uint64_t per = (ntiles + 1023) / 1024;
    const uint64_t a = (uint64_t)threadIdx.x * per, e = a + per < ntiles ? a + per : ntiles;
    long long mine = -1;
    for (uint64_t t = a; t < e; ++t) mine = last_nl[t] > mine ? last_nl[t] : mine;
    long long run = block_excl_max(mine, -1, s_w);
    for (uint64_t t = a; t < e; ++t) {
        entry[t] = run;
        run = last_nl[t] > run ? last_nl[t] : run;
    }
}

// What every per-byte pass does: walks the thread's 16 bytes knowing the line start that governs them.
// f(pos, byte, line_start, in_header) is called for every byte in range.
template <typename F>
__device__ __forceinline__ void walk_bytes(const uint8_t *__restrict__ f, uint64_t m, const long long *__restrict__ entry,
                                           long long *scratch, uint8_t (&b)[kFaPer], int &cnt, uint64_t &i0, F &&fn)
{
    i0 = (uint64_t)blockIdx.x * kFaTile + (uint64_t)threadIdx.x * kFaPer;
    cnt = load16(f, i0, m, b);
    long long last = -1;
#pragma unroll
    for (int j = 0; j < kFaPer; ++j)
        if (j < cnt && b[j] == '\n') last = (long long)(i0 + j);
    const long long before = block_excl_max(last, -1, scratch);
    const long long e = entry[blockIdx.x];
    uint64_t ls = (uint64_t)((before > e ? before : e) + 1); // start of the line the first byte lies in
    bool hdr = ls < m && (ls == i0 ? b[0] : f[ls]) == '>';
#pragma unroll
    for (int j = 0; j < kFaPer; ++j) {
        if (j < cnt) {
            const uint64_t pos = i0 + j;
            const bool start = pos == ls;
            if (start) hdr = b[j] == '>';
            fn(pos, b[j], start, hdr);
            if (b[j] == '\n') ls = pos + 1;
        }
    }
}

// sequence bytes and header lines per tile
__global__ void __launch_bounds__(kFaThreads)
fa_count_kernel(const uint8_t *__restrict__ f, uint64_t m, const long long *__restrict__ entry, uint32_t *__restrict__ cnt_seq,
                uint32_t *__restrict__ cnt_hdr)
{
    __shared__ long long s_m[kFaThreads / 32];
    __shared__ uint32_t s_s[kFaThreads / 32][2];
    uint8_t b[kFaPer];
    int cnt;
    uint64_t i0;
    uint32_t ns = 0, nh = 0;
    walk_bytes(f, m, entry, s_m, b, cnt, i0, [&](uint64_t, uint8_t c, bool start, bool hdr) {
        nh += start && hdr;
        ns += !hdr && c != '\n';
    });
    uint32_t es, eh, tot[2];
    block_excl_sum2(ns, nh, es, eh, s_s, tot);
    if (threadIdx.x == 0) {
        cnt_seq[blockIdx.x] = tot[0];
        cnt_hdr[blockIdx.x] = tot[1];
    }
}

// exclusive sums of up to two u32 arrays into u64 arrays, totals in total[0..1]; one block
__global__ void __launch_bounds__(1024) fa_scan_sum_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                                           uint64_t n, uint64_t *__restrict__ oa, uint64_t *__restrict__ ob,
                                                           uint64_t *__restrict__ total)
{
    __shared__ uint64_t s_a[1024], s_b[1024];
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = (uint64_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint64_t sa = 0, sb = 0;
    for (uint64_t t = lo; t < hi; ++t) {
        sa += a[t];
        if (b) sb += b[t];
    }
    s_a[threadIdx.x] = sa;
    s_b[threadIdx.x] = sb;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t ra = 0, rb = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint64_t ta = s_a[i], tb = s_b[i];
            s_a[i] = ra;
            s_b[i] = rb;
            ra += ta;
            rb += tb;
        }
        total[0] = ra;
        total[1] = rb;
    }
    __syncthreads();
    uint64_t ra = s_a[threadIdx.x], rb = s_b[threadIdx.x];
    for (uint64_t t = lo; t < hi; ++t) {
        oa[t] = ra;
        ra += a[t];
        if (b) {
            ob[t] = rb;
            rb += b[t];
        }
    }
}

// B[k] = sequence bytes in front of the k-th header line (k = 1..H); headers that hold nothing but
// blanks after '>' are counted (the reference's substr() throws std::out_of_range on them, builder.cpp:215)
__global__ void __launch_bounds__(kFaThreads)
fa_mark_kernel(const uint8_t *__restrict__ f, uint64_t m, const long long *__restrict__ entry,
               const uint64_t *__restrict__ off_seq, const uint64_t *__restrict__ off_hdr, uint64_t *__restrict__ B,
               unsigned long long *__restrict__ counters)
{
    __shared__ long long s_m[kFaThreads / 32];
    __shared__ uint32_t s_s[kFaThreads / 32][2];
    uint8_t b[kFaPer];
    int cnt;
    uint64_t i0;
    uint32_t ns = 0, nh = 0;
    uint32_t seqmask = 0, hdrmask = 0; // per byte: is a sequence byte / starts a header line
    walk_bytes(f, m, entry, s_m, b, cnt, i0, [&](uint64_t pos, uint8_t c, bool start, bool hdr) {
        const int j = (int)(pos - i0);
        if (start && hdr) {
            ++nh;
            hdrmask |= 1u << j;
        }
        if (!hdr && c != '\n') {
            ++ns;
            seqmask |= 1u << j;
        }
    });
    uint32_t es, eh, tot[2];
    block_excl_sum2(ns, nh, es, eh, s_s, tot);
    if (nh == 0) return; // no barrier follows
    uint64_t S = off_seq[blockIdx.x] + es, R = off_hdr[blockIdx.x] + eh;
#pragma unroll
    for (int j = 0; j < kFaPer; ++j) {
        if ((hdrmask >> j) & 1u) {
            ++R;
            B[R] = S;
            uint64_t q = i0 + j + 1;
            while (q < m && (f[q] == ' ' || f[q] == '\t')) ++q;
            if (q >= m || f[q] == '\n') atomicAdd(&counters[0], 1ull);
        }
        S += (seqmask >> j) & 1u;
    }
}

// non-empty records per record tile
__global__ void __launch_bounds__(kFaThreads) fa_rec_count_kernel(const uint64_t *__restrict__ B, uint64_t nrec,
                                                                  uint32_t *__restrict__ cnt)
{
    __shared__ uint32_t s_w[kFaThreads / 32];
    const uint64_t k0 = (uint64_t)blockIdx.x * kRecTile + (uint64_t)threadIdx.x * kRecPer;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kRecPer; ++j)
        if (k0 + j < nrec) c += B[k0 + j + 1] > B[k0 + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kFaThreads / 32; ++w) t += s_w[w];
        cnt[blockIdx.x] = t;
    }
}

// O[k] = offset of record k's document in the output = 2*B[k] + 2*(non-empty records before k)
__global__ void __launch_bounds__(kFaThreads) fa_rec_offset_kernel(const uint64_t *__restrict__ B, uint64_t nrec,
                                                                   const uint64_t *__restrict__ tile_off,
                                                                   uint64_t *__restrict__ O)
{
    __shared__ uint32_t s_s[kFaThreads / 32][2];
    const uint64_t k0 = (uint64_t)blockIdx.x * kRecTile + (uint64_t)threadIdx.x * kRecPer;
    uint32_t c = 0;
    bool ne[kRecPer];
#pragma unroll
    for (int j = 0; j < kRecPer; ++j) {
        ne[j] = k0 + j < nrec && B[k0 + j + 1] > B[k0 + j];
        c += ne[j];
    }
    uint32_t e, dummy, tot[2];
    block_excl_sum2(c, 0u, e, dummy, s_s, tot);
    uint64_t before = tile_off[blockIdx.x] + e;
#pragma unroll
    for (int j = 0; j < kRecPer; ++j) {
        if (k0 + j < nrec) O[k0 + j] = 2 * B[k0 + j] + 2 * before;
        before += ne[j];
    }
}

// the documents: complement(s) + '-' + reverse(s) + '\0' per non-empty record
__global__ void __launch_bounds__(kFaThreads)
fa_emit_kernel(const uint8_t *__restrict__ f, uint64_t m, const long long *__restrict__ entry,
               const uint64_t *__restrict__ off_seq, const uint64_t *__restrict__ off_hdr, const uint64_t *__restrict__ B,
               const uint64_t *__restrict__ O, uint8_t *__restrict__ out, uint32_t *__restrict__ bad_bitmap,
               unsigned long long *__restrict__ counters)
{
    __shared__ long long s_m[kFaThreads / 32];
    __shared__ uint32_t s_s[kFaThreads / 32][2];
    __shared__ uint8_t s_map[256], s_comp[256], s_ok[256];
    {
        // normalize() and complement(), builder.cpp:35-55, 60-104
        const int c = threadIdx.x;
        uint8_t v = 'N', ok = 0;
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N' || c == '0' || c == '1' || c == '2' || c == '3' ||
            c == '.') {
            v = (uint8_t)c;
            ok = 1;
        } else if (c == 'a' || c == 'c' || c == 'g' || c == 't' || c == 'n') {
            v = (uint8_t)(c - 'a' + 'A');
            ok = 1;
        }
        s_map[c] = v;
        s_ok[c] = ok;
        s_comp[c] = v == 'A' ? 'T' : v == 'T' ? 'A' : v == 'C' ? 'G' : v == 'G' ? 'C' : v;
    }
    uint8_t b[kFaPer];
    int cnt;
    uint64_t i0;
    uint32_t ns = 0, nh = 0;
    uint32_t seqmask = 0, hdrmask = 0; // per byte: is a sequence byte / starts a header line
    walk_bytes(f, m, entry, s_m, b, cnt, i0, [&](uint64_t pos, uint8_t c, bool start, bool hdr) {
        const int j = (int)(pos - i0);
        if (start && hdr) {
            ++nh;
            hdrmask |= 1u << j;
        }
        if (!hdr && c != '\n') {
            ++ns;
            seqmask |= 1u << j;
        }
    });
    uint32_t es, eh, tot[2];
    block_excl_sum2(ns, nh, es, eh, s_s, tot); // its barriers also publish the tables
    if (ns == 0) return;
    uint64_t S = off_seq[blockIdx.x] + es, R = off_hdr[blockIdx.x] + eh;
    uint64_t rec = ~0ull, b0 = 0, len = 0, base = 0;
#pragma unroll
    for (int j = 0; j < kFaPer; ++j) {
        if ((hdrmask >> j) & 1u) ++R;
        if ((seqmask >> j) & 1u) {
            if (rec != R) {
                rec = R;
                b0 = B[R];
                len = B[R + 1] - b0;
                base = O[R];
            }
            const uint64_t t = S - b0;
            const uint8_t raw = b[j];
            out[base + t] = s_comp[raw];
            out[base + 2 * len - t] = s_map[raw];
            if (t == 0) {
                out[base + len] = '-';
                out[base + 2 * len + 1] = 0;
            }
            if (!s_ok[raw]) {
                const uint32_t bit = 1u << (R & 31);
                if (!(atomicOr(&bad_bitmap[R >> 5], bit) & bit)) atomicAdd(&counters[1], 1ull);
                atomicMin(&counters[2], (unsigned long long)(i0 + j));
            }
            ++S;
        }
    }
}

} // namespace

uint64_t fasta_tiles(uint64_t m) { return div_up(m, kFaTile); }
uint64_t fasta_rec_tiles(uint64_t nrec) { return div_up(nrec, kRecTile); }

void launch_fasta_scan_lines(cudaStream_t st, const uint8_t *text, uint64_t m, long long *last_nl, long long *entry,
                             uint32_t *cnt_seq, uint32_t *cnt_hdr, uint64_t *off_seq, uint64_t *off_hdr, uint64_t *totals,
                             uint32_t *launches)
{
    const unsigned ntiles = (unsigned)fasta_tiles(m);
    fa_last_nl_kernel<<<ntiles, kFaThreads, 0, st>>>(text, m, last_nl);
    fa_scan_max_kernel<<<1, 1024, 0, st>>>(last_nl, ntiles, entry);
    fa_count_kernel<<<ntiles, kFaThreads, 0, st>>>(text, m, entry, cnt_seq, cnt_hdr);
    fa_scan_sum_kernel<<<1, 1024, 0, st>>>(cnt_seq, cnt_hdr, ntiles, off_seq, off_hdr, totals);
    DSM_LAUNCH_CHECK();
    if (launches) *launches += 4;
}

void launch_fasta_records(cudaStream_t st, const uint8_t *text, uint64_t m, const long long *entry, const uint64_t *off_seq,
                          const uint64_t *off_hdr, uint64_t nrec, uint64_t *B, uint64_t *O, uint32_t *rec_cnt,
                          uint64_t *rec_off, uint64_t *rec_total, unsigned long long *counters, uint32_t *launches)
{
    const unsigned ntiles = (unsigned)fasta_tiles(m);
    const unsigned rtiles = (unsigned)fasta_rec_tiles(nrec);
    fa_mark_kernel<<<ntiles, kFaThreads, 0, st>>>(text, m, entry, off_seq, off_hdr, B, counters);
    fa_rec_count_kernel<<<rtiles, kFaThreads, 0, st>>>(B, nrec, rec_cnt);
    fa_scan_sum_kernel<<<1, 1024, 0, st>>>(rec_cnt, nullptr, rtiles, rec_off, nullptr, rec_total);
    fa_rec_offset_kernel<<<rtiles, kFaThreads, 0, st>>>(B, nrec, rec_off, O);
    DSM_LAUNCH_CHECK();
    if (launches) *launches += 4;
}

void launch_fasta_emit(cudaStream_t st, const uint8_t *text, uint64_t m, const long long *entry, const uint64_t *off_seq,
                       const uint64_t *off_hdr, const uint64_t *B, const uint64_t *O, uint8_t *out, uint32_t *bad_bitmap,
                       unsigned long long *counters, uint32_t *launches)
{
    fa_emit_kernel<<<(unsigned)fasta_tiles(m), kFaThreads, 0, st>>>(text, m, entry, off_seq, off_hdr, B, O, out, bad_bitmap,
                                                                   counters);
    DSM_LAUNCH_CHECK();
    if (launches) ++*launches;
}

} // namespace dsmfm
