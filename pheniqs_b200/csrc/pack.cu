/*  pack.cu — the feed -> tile kernel (pack.cuh). Lane = read. */
#include "pack.cuh"

namespace phq {
namespace {

/* iupac.h:153-171 AsciiToAmbiguousBam, rows 0x30-0x7f; everything else is 15 */
__constant__ uint8_t ASCII_TO_BAM[80] = {
     1, 2, 4, 8, 15,15,15,15, 15,15,15,15, 15, 0,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15,
    15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15,
    15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
};
/* sequence.h BamToReverseComplementBam */
__constant__ uint8_t BAM_COMPLEMENT[16] = { 0x0, 0x8, 0x4, 0xc, 0x2, 0xa, 0x6, 0xe, 0x1, 0x9, 0x5, 0xd, 0x3, 0xb, 0x7, 0xf };

struct Base { uint32_t code; uint32_t quality; };

__device__ __forceinline__ long long read_begin(const RawSegmentView& v, long long r) {
    return v.offset != nullptr ? v.offset[v.first + r] : (v.first + r) * v.length;
}
__device__ __forceinline__ int read_length(const RawSegmentView& v, long long r) {
    return v.offset != nullptr ? static_cast< int >(v.offset[v.first + r + 1] - v.offset[v.first + r]) : static_cast< int >(v.length);
}
/* transform.h:65-80 */
__device__ __forceinline__ int absolute_end(const PackToken& t, int n) {
    if(t.end_terminated) {
        if(t.end < 0) { const int v = n + t.end; return v < 0 ? 0 : v; }
        return t.end > n ? n : t.end;
    }
    return n;
}
__device__ __forceinline__ int absolute_start(const PackToken& t, int n) {
    if(t.start < 0) { const int v = n + t.start; return v < 0 ? 0 : v; }
    return t.start > n ? 0 : t.start;
}

/* nucleotides Rule::apply appends to output segment s for read r (transform.h:142-169) */
__device__ __forceinline__ int observed_length(const PackPlan& plan, const uint8_t* table, long long r, int s) {
    int length = 0;
    for(int k = 0; k < plan.token_cardinality; ++k) {
        const PackToken& t = plan.token[k];
        if(t.output_segment != s) { continue; }
        const int n = read_length(plan.input[t.input_segment], r);
        const int size = absolute_end(t, n) - absolute_start(t, n);
        length += size > 0 ? size : 0;
    }
    return length;
}
/* nucleotide i (< observed length) of output segment s of read r */
__device__ __forceinline__ Base fetch(const PackPlan& plan, const uint8_t* table, long long r, int s, int i) {
    Base b;
    b.code = 0; b.quality = 0;
    int at = 0;
    for(int k = 0; k < plan.token_cardinality; ++k) {
        const PackToken& t = plan.token[k];
        if(t.output_segment != s) { continue; }
        const RawSegmentView& v = plan.input[t.input_segment];
        const int n = read_length(v, r);
        const int start = absolute_start(t, n);
        const int end = absolute_end(t, n);
        const int size = end - start;
        if(size <= 0) { continue; }
        if(i < at + size) {
            const int within = i - at;
            const long long source = read_begin(v, r) + (t.reverse_complement ? (end - within - 1) : (start + within));
            const uint32_t letter = v.sequence[source];
            uint32_t code = (letter >= 0x30u && letter < 0x80u) ? table[letter - 0x30u] : 15u;
            if(t.reverse_complement) { code = table[80 + code]; }
            b.code = code;
            b.quality = (static_cast< uint32_t >(v.quality[source]) - static_cast< uint32_t >(plan.phred_offset)) & 0xffu;     /* char arithmetic, fastq.h:73 */
            return b;
        }
        at += size;
    }
    return b;
}

__global__ void __launch_bounds__(256)
pack_kernel(const __grid_constant__ PackPlan plan, const long long n_reads, uint32_t* __restrict__ bases, uint16_t* __restrict__ nmask,
            uint32_t* __restrict__ quality, const long long pitch) {
    __shared__ uint8_t table[96];
    if(threadIdx.x < 80) { table[threadIdx.x] = ASCII_TO_BAM[threadIdx.x]; }
    else if(threadIdx.x < 96) { table[threadIdx.x] = BAM_COMPLEMENT[threadIdx.x - 80]; }
    __syncthreads();
    const int L = plan.nucleotide_cardinality;
    for(long long r = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; r < n_reads; r += static_cast< long long >(gridDim.x) * blockDim.x) {
        uint32_t lo = 0, hi = 0, ambiguous = 0;
        uint32_t phred[PHQ_MAX_NUCLEOTIDES / 4];
        #pragma unroll
        for(int w = 0; w < PHQ_MAX_NUCLEOTIDES / 4; ++w) { phred[w] = 0; }
        for(int s = 0; s < plan.segment_cardinality; ++s) {
            const int expected = plan.segment_offset[s + 1] - plan.segment_offset[s];
            const int length = observed_length(plan, table, r, s);
            for(int i = 0; i < expected; ++i) {
                const int j = plan.segment_offset[s] + i;
                Base b;
                if(i < length) {
                    b = fetch(plan, table, r, s, i);
                } else if(!plan.stale_semantics) {
                    /* absent position (phq_pack): quality PHQ_ABSENT_QUALITY, ambiguous with both base bits set */
                    lo |= 1u << j; hi |= 1u << j; ambiguous |= 1u << j;
                    phred[j >> 2] |= static_cast< uint32_t >(PHQ_ABSENT_QUALITY) << (8 * (j & 3));
                    continue;
                } else if(i == length) {
                    b.code = 0; b.quality = 0;                      /* the terminator the last append wrote */
                } else {
                    /* what an earlier read left there: the nearest one that reached position i */
                    b.code = plan.carry_code[j]; b.quality = plan.carry_quality[j];
                    for(long long earlier = r - 1; earlier >= 0; --earlier) {
                        const int reach = observed_length(plan, table, earlier, s);
                        if(reach == i) { b.code = 0; b.quality = 0; break; }
                        if(reach > i) { b = fetch(plan, table, earlier, s, i); break; }
                    }
                }
                switch(b.code) {
                    case 1: break;
                    case 2: lo |= 1u << j; break;
                    case 4: hi |= 1u << j; break;
                    case 8: lo |= 1u << j; hi |= 1u << j; break;
                    default: ambiguous |= 1u << j; break;
                }
                phred[j >> 2] |= b.quality << (8 * (j & 3));
            }
        }
        bases[r] = (lo & 0xffffu) | ((hi & 0xffffu) << 16);
        nmask[r] = static_cast< uint16_t >(ambiguous & 0xffffu);
        if(L > 16) {
            bases[pitch + r] = (lo >> 16) | (hi & 0xffff0000u);
            nmask[pitch + r] = static_cast< uint16_t >(ambiguous >> 16);
        }
        #pragma unroll
        for(int w = 0; w < PHQ_MAX_NUCLEOTIDES / 4; ++w) {
            if(4 * w < L) { quality[w * pitch + r] = phred[w]; }
        }
    }
}

}   /* namespace */

cudaError_t launch_pack(const PackPlan& plan, long long n_reads, uint32_t* bases, uint16_t* nmask, uint32_t* quality, long long pitch,
                        int multiprocessor_count, cudaStream_t stream) {
    if(n_reads <= 0) { return cudaSuccess; }
    const int threads = 256;
    const long long blocks = (n_reads + threads - 1) / threads;
    const long long resident = static_cast< long long >(multiprocessor_count) * 8;
    pack_kernel<<< static_cast< int >(blocks < resident ? blocks : resident), threads, 0, stream >>>(plan, n_reads, bases, nmask, quality, pitch);
    return cudaGetLastError();
}

}   /* namespace phq */
