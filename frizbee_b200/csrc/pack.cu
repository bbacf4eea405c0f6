// pack.cu — Arrow-style (bytes, offsets) → tile-bucketed, unit-interleaved corpus (frz_device.cuh).
//
// Replaces the `&[S: AsRef<str>]` argument of Matcher::match_list (src/matcher/mod.rs:212):
// the reference chases one fat pointer per haystack; here the list is packed once and stays
// resident in HBM across queries.
#include "frz_device.cuh"
#include "frz_host.h"

namespace {

// ---- plan: one block per tile: bucket by unit count, emit slot metadata + group descriptors ----
__global__ void __launch_bounds__(256) k_pack_plan(const uint64_t* __restrict__ offsets, uint64_t n,
                                                   uint32_t* __restrict__ slot_meta, uint16_t* __restrict__ slot_of,
                                                   FrzGroupDesc* __restrict__ groups, uint64_t* __restrict__ tile_units,
                                                   unsigned int* __restrict__ err) {
    __shared__ uint32_t key[FRZ_TILE];
    __shared__ uint32_t len_s[FRZ_TILE];
    __shared__ uint32_t gun[FRZ_GROUPS_PER_TILE];
    const uint32_t tile = blockIdx.x;
    const uint64_t base = (uint64_t)tile * FRZ_TILE;
    for (int i = threadIdx.x; i < FRZ_TILE; i += blockDim.x) {
        uint64_t idx = base + i;
        if (idx < n) {
            uint64_t len = offsets[idx + 1] - offsets[idx];
            if (len > FRZ_MAX_HAY_LEN) { atomicOr(err, 1u); len = FRZ_MAX_HAY_LEN; }
            uint32_t units = (uint32_t)((len + FRZ_UNIT - 1) / FRZ_UNIT);
            len_s[i] = (uint32_t)len;
            key[i] = (units << FRZ_TILE_SHIFT) | (uint32_t)i;   // units < 2^18, stable via index bits
        } else {
            len_s[i] = 0;
            key[i] = 0xFFFFFFFFu;
        }
    }
    __syncthreads();
    // bitonic sort of 1024 keys (ascending)
    for (int k = 2; k <= FRZ_TILE; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < FRZ_TILE; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint32_t a = key[i], b = key[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int s = threadIdx.x; s < FRZ_TILE; s += blockDim.x) {
        uint32_t kk = key[s];
        if (kk == 0xFFFFFFFFu) {
            slot_meta[base + s] = FRZ_INVALID_SLOT;
        } else {
            uint32_t li = kk & (FRZ_TILE - 1);
            slot_meta[base + s] = (len_s[li] << FRZ_TILE_SHIFT) | li;
            slot_of[base + li] = (uint16_t)s;
        }
    }
    if (threadIdx.x < FRZ_GROUPS_PER_TILE) {
        // ascending order ⇒ the last valid lane of the group carries the group's max
        uint32_t g = threadIdx.x, mx = 0;
        for (int l = FRZ_GROUP - 1; l >= 0; l--) {
            uint32_t kk = key[g * FRZ_GROUP + l];
            if (kk != 0xFFFFFFFFu) { mx = kk >> FRZ_TILE_SHIFT; break; }
        }
        gun[g] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t off = 0;
        for (int g = 0; g < FRZ_GROUPS_PER_TILE; g++) {
            groups[tile * FRZ_GROUPS_PER_TILE + g] = FrzGroupDesc{off, gun[g]};
            off += gun[g] * FRZ_GROUP;
        }
        tile_units[tile] = off;
    }
}

// ---- exclusive scan of per-tile unit counts (single block; n_tiles is N/1024) ----
__global__ void __launch_bounds__(1024) k_scan_u64(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                   uint32_t n, uint64_t* __restrict__ total) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = i < n ? in[i] : 0;
        uint64_t x = v;
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (frz_lane() >= (uint32_t)d) x += y;
        }
        if (frz_lane() == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t w = warp_sum[threadIdx.x], xs = w;
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t y = __shfl_up_sync(0xffffffffu, xs, d);
                if (frz_lane() >= (uint32_t)d) xs += y;
            }
            warp_sum[threadIdx.x] = xs - w;  // exclusive
        }
        __syncthreads();
        uint64_t c = carry_s;
        uint64_t incl = c + warp_sum[threadIdx.x >> 5] + x;
        if (i < n) out[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__device__ __forceinline__ uint32_t load_word_safe(const uint8_t* bytes, uint64_t pos, uint64_t total_bytes) {
    // little-endian word at byte position `pos` (pos % 4 == alignment of bytes base assumed by caller)
    if (pos + 4 <= total_bytes) return *reinterpret_cast<const uint32_t*>(bytes + pos);
    uint32_t w = 0;
    for (int b = 0; b < 4; b++)
        if (pos + b < total_bytes) w |= (uint32_t)bytes[pos + b] << (8 * b);
    return w;
}

// ---- copy: one block per tile; thread per output unit so the 16-byte stores are coalesced ----
__global__ void __launch_bounds__(256) k_pack_copy(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ offsets,
                                                   uint64_t n, uint64_t total_bytes, const uint32_t* __restrict__ slot_meta,
                                                   const FrzGroupDesc* __restrict__ groups, const uint64_t* __restrict__ tile_base,
                                                   uint4* __restrict__ data) {
    __shared__ uint32_t goff[FRZ_GROUPS_PER_TILE + 1];
    const uint32_t tile = blockIdx.x;
    if (threadIdx.x < FRZ_GROUPS_PER_TILE) {
        FrzGroupDesc gd = groups[tile * FRZ_GROUPS_PER_TILE + threadIdx.x];
        goff[threadIdx.x] = gd.unit_off;
        if (threadIdx.x == FRZ_GROUPS_PER_TILE - 1) goff[FRZ_GROUPS_PER_TILE] = gd.unit_off + gd.gunits * FRZ_GROUP;
    }
    __syncthreads();
    const uint32_t total_units = goff[FRZ_GROUPS_PER_TILE];
    const uint64_t tb = tile_base[tile];
    const uint64_t misalign = reinterpret_cast<uintptr_t>(bytes) & 3;  // base pointer alignment
    const uint8_t* abase = bytes - misalign;                           // 4-byte aligned
    const uint64_t atotal = total_bytes + misalign;
    for (uint32_t u = threadIdx.x; u < total_units; u += blockDim.x) {
        // group of unit u: largest g with goff[g] <= u (groups with gunits == 0 share an offset)
        int lo = 0, hi = FRZ_GROUPS_PER_TILE - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (goff[mid] <= u) lo = mid; else hi = mid - 1;
        }
        // skip empty groups that start at the same offset: pick the one whose range contains u
        while (lo < FRZ_GROUPS_PER_TILE - 1 && goff[lo + 1] <= u) lo++;
        uint32_t rel = u - goff[lo];
        uint32_t k = rel >> 5, lane = rel & 31;
        uint32_t meta = slot_meta[(uint64_t)tile * FRZ_TILE + lo * FRZ_GROUP + lane];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (meta != FRZ_INVALID_SLOT) {
            uint32_t len = meta >> FRZ_TILE_SHIFT;
            uint64_t idx = (uint64_t)tile * FRZ_TILE + (meta & (FRZ_TILE - 1));
            uint32_t b0 = k * FRZ_UNIT;
            if (b0 < len) {
                uint64_t src = offsets[idx] + b0 + misalign;   // position in the aligned view
                uint32_t nvalid = min(len - b0, (uint32_t)FRZ_UNIT);
                uint64_t a = src & ~3ull;
                uint32_t sh = (uint32_t)(src & 3) * 8;
                uint32_t w[5];
#pragma unroll
                for (int j = 0; j < 5; j++) w[j] = load_word_safe(abase, a + 4 * j, atotal);
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    o[j] = __funnelshift_r(w[j], w[j + 1], sh);
                    int rem = (int)nvalid - 4 * j;
                    if (rem <= 0) o[j] = 0;
                    else if (rem < 4) o[j] &= (1u << (8 * rem)) - 1;
                }
                v = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        data[tb + u] = v;
    }
}

}  // namespace

// Builds the packed corpus from device-resident Arrow buffers.  Asynchronous on `stream`
// except for one small D2H copy (the packed size), which it has to wait for to allocate.
frz_status frz_pack_corpus_device(const uint8_t* d_bytes, const uint64_t* d_offsets, uint64_t n, uint64_t total_bytes,
                                  cudaStream_t stream, FrzCorpusStorage* out) {
    uint32_t n_tiles = (uint32_t)((n + FRZ_TILE - 1) / FRZ_TILE);
    out->n = n;
    out->n_tiles = n_tiles;
    out->total_bytes = total_bytes;
    if (n_tiles == 0) {
        out->total_units = 0;
        return FRZ_OK;
    }
    size_t slots = (size_t)n_tiles * FRZ_TILE;
    if (out->cap_tiles < n_tiles) {
        cudaFree(out->slot_meta); cudaFree(out->slot_of); cudaFree(out->groups); cudaFree(out->tile_base); cudaFree(out->scratch_tile_units);
        out->slot_meta = nullptr; out->slot_of = nullptr; out->groups = nullptr; out->tile_base = nullptr; out->scratch_tile_units = nullptr;
        out->cap_tiles = 0;
        FRZ_CUDA_TRY(cudaMalloc(&out->slot_meta, slots * sizeof(uint32_t)));
        FRZ_CUDA_TRY(cudaMalloc(&out->slot_of, slots * sizeof(uint16_t)));
        FRZ_CUDA_TRY(cudaMalloc(&out->groups, (size_t)n_tiles * FRZ_GROUPS_PER_TILE * sizeof(FrzGroupDesc)));
        FRZ_CUDA_TRY(cudaMalloc(&out->tile_base, (size_t)n_tiles * sizeof(uint64_t)));
        FRZ_CUDA_TRY(cudaMalloc(&out->scratch_tile_units, ((size_t)n_tiles + 2) * sizeof(uint64_t)));
        out->cap_tiles = n_tiles;
    }
    uint64_t* d_tile_units = out->scratch_tile_units;
    uint64_t* d_total = d_tile_units + out->cap_tiles;
    unsigned int* d_err = reinterpret_cast<unsigned int*>(d_total + 1);
    FRZ_CUDA_TRY(cudaMemsetAsync(d_total, 0, 16, stream));
    FRZ_CUDA_TRY(cudaMemsetAsync(out->slot_of, 0, slots * sizeof(uint16_t), stream));
    k_pack_plan<<<n_tiles, 256, 0, stream>>>(d_offsets, n, out->slot_meta, out->slot_of, out->groups, d_tile_units, d_err);
    k_scan_u64<<<1, 1024, 0, stream>>>(d_tile_units, out->tile_base, n_tiles, d_total);
    uint64_t h[2] = {0, 0};
    FRZ_CUDA_TRY(cudaMemcpyAsync(h, d_total, 16, cudaMemcpyDeviceToHost, stream));
    FRZ_CUDA_TRY(cudaStreamSynchronize(stream));
    unsigned int err = (unsigned int)(h[1] & 0xffffffffu);
    if (err) return frz_fail(FRZ_ERR_UNSUPPORTED, "haystack longer than 4 MiB");
    out->total_units = h[0];
    if (out->cap_units < out->total_units + 1) {
        cudaFree(out->data); out->data = nullptr; out->cap_units = 0;
        const uint64_t want = out->total_units + out->total_units / 16 + 1024;
        FRZ_CUDA_TRY(cudaMalloc(&out->data, (size_t)want * sizeof(uint4)));
        out->cap_units = want;
    }
    k_pack_copy<<<n_tiles, 256, 0, stream>>>(d_bytes, d_offsets, n, total_bytes, out->slot_meta, out->groups,
                                             out->tile_base, out->data);
    FRZ_CUDA_TRY(cudaGetLastError());
    return FRZ_OK;
}
