// primitives.cu -- hand-written device-wide exclusive scan and stable LSD radix sort.
// Used by the task-list builder (offsets over C rows / C tiles), the line indices (CSR/CSC of a block table),
// assembly (unique tile extraction) and the Morton ordering of C.  HBM/latency-bound integer work.
#include "common.cuh"

namespace hbsm_b200 {

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
template <int THREADS>
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t* total, uint64_t* smem /*THREADS/32+1*/) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = warp_incl_scan(v);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = (lane < THREADS / 32) ? smem[lane] : 0;
        uint64_t wi = warp_incl_scan(w);
        if (lane < THREADS / 32) smem[lane] = wi - w;
        if (lane == 31) smem[THREADS / 32] = wi;
    }
    __syncthreads();
    uint64_t res = incl - v + smem[warp];
    *total = smem[THREADS / 32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const uint32_t* __restrict__ in, size_t n,
                                                                    uint64_t* __restrict__ block_sums) {
    __shared__ uint64_t sm[SCAN_THREADS / 32 + 1];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint64_t total;
    block_excl_scan<SCAN_THREADS>(s, &total, sm);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block, in-place exclusive scan of a u64 array; data[n] = total
__global__ void __launch_bounds__(1024) k_scan_u64_single(uint64_t* data, size_t n) {
    __shared__ uint64_t sm[1024 / 32 + 1];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += 1024) {
        size_t idx = base + threadIdx.x;
        uint64_t v = idx < n ? data[idx] : 0;
        uint64_t total;
        uint64_t ex = block_excl_scan<1024>(v, &total, sm);
        uint64_t carry = carry_s;
        if (idx < n) data[idx] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) data[n] = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in, size_t n,
                                                               const uint64_t* __restrict__ block_offsets,
                                                               uint64_t* __restrict__ out) {
    __shared__ uint64_t sm[SCAN_THREADS / 32 + 1];
    // thread t owns SCAN_ITEMS consecutive items so the per-thread prefix is a running sum
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    uint64_t total;
    uint64_t ex = block_excl_scan<SCAN_THREADS>(s, &total, sm) + block_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_offsets[gridDim.x];
}

// whole scan in ONE block (one launch instead of three): for the short arrays that dominate the launch count of the
// task-list builder and of the radix sort's digit histograms
__global__ void __launch_bounds__(1024) k_scan_small(const uint32_t* __restrict__ in, size_t n, uint64_t* __restrict__ out) {
    __shared__ uint64_t sm[1024 / 32 + 1];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    constexpr int ITEMS = 4;
    for (size_t base = 0; base < n; base += 1024 * ITEMS) {
        const size_t i0 = base + (size_t)threadIdx.x * ITEMS;
        uint32_t v[ITEMS];
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) { v[i] = (i0 + i < n) ? in[i0 + i] : 0; s += v[i]; }
        uint64_t total;
        uint64_t ex = block_excl_scan<1024>(s, &total, sm) + carry_s;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) { if (i0 + i < n) out[i0 + i] = ex; ex += v[i]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

}  // namespace

void exclusive_scan_u32(const uint32_t* d_in, uint64_t* d_out, size_t n) {
    if (n == 0) {
        HB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(uint64_t), engine().stream));
        return;
    }
    if (n <= 32 * 1024) {
        HB_LAUNCH(k_scan_small, 1, 1024, 0, d_in, n, d_out);
        return;
    }
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf<uint64_t> sums(nb + 1);
    HB_LAUNCH(k_scan_block_sums, (unsigned)nb, SCAN_THREADS, 0, d_in, n, sums.p);
    HB_LAUNCH(k_scan_u64_single, 1, 1024, 0, sums.p, nb);
    HB_LAUNCH(k_scan_apply, (unsigned)nb, SCAN_THREADS, 0, d_in, n, sums.p, d_out);
}

// ---------------------------------------------------------------------------------------------------
// radix sort: 8-bit digits, 3 kernels per pass (histogram, scan of digit-major histograms, stable scatter)
// ---------------------------------------------------------------------------------------------------
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;                       // rounds of RS_THREADS items per block
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;    // 2048 items per block
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, size_t n, int shift,
                                                          uint32_t* __restrict__ hist /*[RADIX][nblk]*/) {
    __shared__ uint32_t h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        size_t idx = base + (size_t)r * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & (RS_RADIX - 1)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in,
                                                             const uint32_t* __restrict__ vals_in, size_t n, int shift,
                                                             const uint64_t* __restrict__ offsets /*[RADIX][nblk]*/,
                                                             uint64_t* __restrict__ keys_out,
                                                             uint32_t* __restrict__ vals_out) {
    constexpr int NW = RS_THREADS / 32;
    __shared__ uint32_t warp_cnt[NW][RS_RADIX];   // per-round count of each digit in each warp
    __shared__ uint64_t digit_base[RS_RADIX];     // running global position of each digit for this block
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    digit_base[threadIdx.x] = offsets[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
    size_t base = (size_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        for (int w = 0; w < NW; ++w) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        size_t idx = base + (size_t)r * RS_THREADS + threadIdx.x;
        bool valid = idx < n;
        uint64_t key = valid ? keys_in[idx] : 0;
        uint32_t val = valid ? vals_in[idx] : 0;
        unsigned digit = (unsigned)((key >> shift) & (RS_RADIX - 1));
        unsigned active = __ballot_sync(0xffffffffu, valid);
        unsigned rank_in_warp = 0;
        if (valid) {
            unsigned peers = __match_any_sync(active, digit);
            rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (rank_in_warp == 0) warp_cnt[warp][digit] = __popc(peers);
        }
        __syncthreads();
        // position = digit_base + (items of this digit in earlier warps of this round) + rank within warp
        if (valid) {
            uint32_t before = 0;
            for (unsigned w = 0; w < warp; ++w) before += warp_cnt[w][digit];
            uint64_t pos = digit_base[digit] + before + rank_in_warp;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        {
            uint32_t tot = 0;
            for (int w = 0; w < NW; ++w) tot += warp_cnt[w][threadIdx.x];
            digit_base[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

}  // namespace

void radix_sort_pairs(uint64_t* d_keys, uint32_t* d_vals, size_t n, int key_bits) {
    radix_sort_pairs_bits(d_keys, d_vals, n, 0, key_bits);
}

void radix_sort_pairs_bits(uint64_t* d_keys, uint32_t* d_vals, size_t n, int lo_bit, int key_bits) {
    if (n <= 1) return;
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 64) key_bits = 64;
    if (lo_bit < 0 || lo_bit >= key_bits) lo_bit = 0;
    int passes = (key_bits - lo_bit + 7) / 8;
    size_t nblk = (n + RS_TILE - 1) / RS_TILE;
    DevBuf<uint64_t> keys_tmp(n);
    DevBuf<uint32_t> vals_tmp(n);
    DevBuf<uint32_t> hist((size_t)RS_RADIX * nblk);
    DevBuf<uint64_t> offs((size_t)RS_RADIX * nblk + 1);
    uint64_t* kin = d_keys; uint64_t* kout = keys_tmp.p;
    uint32_t* vin = d_vals; uint32_t* vout = vals_tmp.p;
    for (int p = 0; p < passes; ++p) {
        int shift = lo_bit + 8 * p;
        HB_LAUNCH(k_rs_hist, (unsigned)nblk, RS_THREADS, 0, kin, n, shift, hist.p);
        exclusive_scan_u32(hist.p, offs.p, (size_t)RS_RADIX * nblk);
        HB_LAUNCH(k_rs_scatter, (unsigned)nblk, RS_THREADS, 0, kin, vin, n, shift, offs.p, kout, vout);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    if (kin != d_keys) {
        HB_CUDA(cudaMemcpyAsync(d_keys, kin, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, engine().stream));
        HB_CUDA(cudaMemcpyAsync(d_vals, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, engine().stream));
    }
}

namespace {
__global__ void k_post_scalars(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, volatile uint64_t* mailbox) {
    mailbox[0] = a ? *a : 0;
    mailbox[1] = b ? *b : 0;
    __threadfence_system();
}
}  // namespace

std::pair<uint64_t, uint64_t> read_scalars(const uint64_t* d_a, const uint64_t* d_b) {
    Engine& e = engine();
    HB_LAUNCH(k_post_scalars, 1, 1, 0, d_a, d_b, e.mailbox);
    HB_CUDA(cudaStreamSynchronize(e.stream));
    return {e.mailbox[0], e.mailbox[1]};
}

}  // namespace hbsm_b200
