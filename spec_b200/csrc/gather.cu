// All-gather of the packed per-image output records over NVLink PEER MEMORY (SURVEY.md 8b `specb200_allgather_outputs`, 8e).
//
// The reference has no collective (single-process eval loop, spec/tester.py:143-167); BASELINE.json's multi-GPU configs shard
// the batch over ranks and gather the per-image records.  Round 1 did that with ncclAllGather, an SM kernel (many CTAs, tens
// of KB of shared memory each) that has to find room beside persistent one-CTA-per-SM conv kernels holding ~200 KB of
// shared memory: at 8 GPUs the median step grew from 7.9 to 9.8 ms (SCALE_r01.json).  Here every rank owns a receive region
// (slots x world x block bytes, cudaMalloc'ed by the library and exported as a CUDA IPC handle) that all peers map, and a
// gather is a PUT:
//   mode 0 (default)  world-1 cudaMemcpyAsync peer copies -> the COPY ENGINES move the block over NVLink, no SM is used;
//   mode 1            one push kernel: a few CTAs read the block once and store it to every peer with 16-byte vector stores;
// followed by a flag protocol on the same stream: `signal` (1 CTA) release-stores the step's sequence number into slot
// [rank] of every peer's flag array, `wait` (1 CTA) spins (bounded) until all peers have published it.  Both are 32/64-thread
// kernels without shared memory, so they co-reside with the persistent conv kernels.  Nothing here synchronises the host.
//
// Buffer-reuse safety is the caller's protocol (spec_b200/pipeline.py::PeerGatherer): with >= 3 slots, a peer's PUT into
// slot k of step i can only be issued after that peer observed this rank's signal of step i-1, which this rank's stream
// issues after everything that consumed slot k's previous content (step i-3) was enqueued.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/specb200.h"
#include "internal.h"

using namespace sb;

namespace {

constexpr int MAX_WORLD = 64;
constexpr unsigned SPIN_LIMIT = 1u << 27;          // bounded wait: a dead peer traps this rank instead of hanging the box

struct PeerTable { uint8_t* data[MAX_WORLD]; uint32_t* flags[MAX_WORLD]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// flags[peer][rank] = seq for every peer (own entry included: the wait kernel treats all ranks alike)
__global__ void gather_signal_kernel(PeerTable t, int world, int rank, int slot, int slots, uint32_t seq) {
    const int r = threadIdx.x;
    if (r < world) {
        __threadfence_system();
        st_release_sys(t.flags[r] + static_cast<size_t>(slot) * MAX_WORLD + rank, seq);
    }
}

// all peers have published seq for this slot (sequence numbers only grow; signed difference handles wrap-around)
__global__ void gather_wait_kernel(const uint32_t* my_flags, int world, int slot, uint32_t seq) {
    const int r = threadIdx.x;
    if (r < world) {
        const uint32_t* f = my_flags + static_cast<size_t>(slot) * MAX_WORLD + r;
        unsigned spins = 0;
        while (static_cast<int32_t>(ld_acquire_sys(f) - seq) < 0) {
            if (++spins > SPIN_LIMIT) { __trap(); }
            __nanosleep(200);
        }
    }
    __syncthreads();
}

// mode 1: every CTA reads its share of the local block once and stores it to all peers (NVLink posted writes, 16 B each);
// the last CTA to finish publishes the flags (threadfence-reduction pattern: fence.sys, then a device-scope counter).
__global__ void __launch_bounds__(256)
gather_push_kernel(PeerTable t, const uint4* __restrict__ src, size_t n16, size_t dst_off_bytes, int world, int rank, int slot,
                   uint32_t seq, unsigned* counter) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = src[i];
        for (int p = 0; p < world; ++p) {
            if (p == rank) continue;
            reinterpret_cast<uint4*>(t.data[p] + dst_off_bytes)[i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *counter = 0;
        if (threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(t.flags[threadIdx.x] + static_cast<size_t>(slot) * MAX_WORLD + rank, seq);
        }
    }
}

}  // namespace

struct specb200_gather {
    int rank = 0, world = 1, slots = 3;
    size_t block = 0;                 // bytes each rank contributes per step
    size_t data_bytes = 0;            // slots * world * block
    uint8_t* region = nullptr;        // [data | flags (slots * MAX_WORLD u32) | counter]
    PeerTable peers = {};
    std::vector<void*> opened;        // IPC mappings to close
    bool connected = false;
    int push_ctas = 16;
};

static size_t region_bytes(const specb200_gather* g) { return g->data_bytes + static_cast<size_t>(g->slots) * MAX_WORLD * 4 + 256; }
static uint32_t* region_flags(const specb200_gather* g, uint8_t* base) { return reinterpret_cast<uint32_t*>(base + g->data_bytes); }

extern "C" int specb200_gather_create(specb200_gather_t** out, int32_t rank, int32_t world, int64_t block_bytes, int32_t slots,
                                      uint8_t* ipc_handle_out) {
    if (!out || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || block_bytes <= 0 || (block_bytes & 15) || slots < 1 || slots > 8 ||
        !ipc_handle_out) {
        set_error("gather_create: bad arguments (block_bytes must be a positive multiple of 16, world <= 64, slots <= 8)");
        return 1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == SPECB200_IPC_HANDLE_BYTES, "IPC handle size");
    specb200_gather* g = new specb200_gather();
    g->rank = rank; g->world = world; g->slots = slots; g->block = static_cast<size_t>(block_bytes);
    g->data_bytes = static_cast<size_t>(slots) * world * g->block;
    if (!check_cuda(cudaMalloc(&g->region, region_bytes(g)), "gather: cudaMalloc") ||
        !check_cuda(cudaMemset(g->region, 0, region_bytes(g)), "gather: cudaMemset")) { delete g; return 1; }
    cudaIpcMemHandle_t h;
    if (!check_cuda(cudaIpcGetMemHandle(&h, g->region), "gather: cudaIpcGetMemHandle")) { cudaFree(g->region); delete g; return 1; }
    memcpy(ipc_handle_out, &h, sizeof(h));
    g->peers.data[rank] = g->region;
    g->peers.flags[rank] = region_flags(g, g->region);
    if (world == 1) g->connected = true;
    *out = g;
    return 0;
}

extern "C" int specb200_gather_connect(specb200_gather_t* g, const uint8_t* all_handles) {
    if (!g || !all_handles) { set_error("gather_connect: bad arguments"); return 1; }
    if (!check_cuda(cudaDeviceSynchronize(), "gather_connect: sync")) return 1;        // the memset of create has landed
    for (int p = 0; p < g->world; ++p) {
        if (p == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + static_cast<size_t>(p) * SPECB200_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        if (!check_cuda(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess), "gather_connect: cudaIpcOpenMemHandle (peer access over NVLink required)")) return 1;
        g->opened.push_back(ptr);
        g->peers.data[p] = static_cast<uint8_t*>(ptr);
        g->peers.flags[p] = region_flags(g, static_cast<uint8_t*>(ptr));
    }
    g->connected = true;
    return 0;
}

extern "C" void* specb200_gather_recv_ptr(specb200_gather_t* g, int32_t slot) {
    if (!g || slot < 0 || slot >= g->slots) return nullptr;
    return g->region + static_cast<size_t>(slot) * g->world * g->block;
}

extern "C" int specb200_allgather_outputs(specb200_gather_t* g, const void* src_dev, int32_t slot, uint32_t seq, int32_t mode, void* stream) {
    if (!g || !src_dev || slot < 0 || slot >= g->slots || mode < 0 || mode > 1) { set_error("allgather_outputs: bad arguments"); return 1; }
    if (!g->connected) { set_error("allgather_outputs: specb200_gather_connect has not been called"); return 1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NvtxRange nvtx("specb200:allgather_outputs (peer PUT)");
    const size_t off = (static_cast<size_t>(slot) * g->world + g->rank) * g->block;       // this rank's block inside a slot
    uint8_t* own = g->region + off;
    if (src_dev != own && !check_cuda(cudaMemcpyAsync(own, src_dev, g->block, cudaMemcpyDeviceToDevice, s), "allgather: local copy")) return 1;
    if (g->world == 1) return 0;
    uint32_t* my_flags = g->peers.flags[g->rank];
    if (mode == 0) {
        for (int d = 1; d < g->world; ++d) {                                              // staggered start: rank r begins with peer r+1
            const int p = (g->rank + d) % g->world;
            if (!check_cuda(cudaMemcpyAsync(g->peers.data[p] + off, own, g->block, cudaMemcpyDeviceToDevice, s), "allgather: peer copy")) return 1;
        }
        gather_signal_kernel<<<1, MAX_WORLD, 0, s>>>(g->peers, g->world, g->rank, slot, g->slots, seq);
    } else {
        unsigned* counter = reinterpret_cast<unsigned*>(g->region + g->data_bytes + static_cast<size_t>(g->slots) * MAX_WORLD * 4);
        gather_push_kernel<<<g->push_ctas, 256, 0, s>>>(g->peers, reinterpret_cast<const uint4*>(own), g->block / 16, off, g->world, g->rank, slot, seq, counter);
        // own flag: the push kernel's last CTA wrote it together with the peers'
    }
    gather_wait_kernel<<<1, MAX_WORLD, 0, s>>>(my_flags, g->world, slot, seq);
    return check_cuda(cudaGetLastError(), "allgather_outputs launch") ? 0 : 1;
}

extern "C" int specb200_gather_set_push_ctas(specb200_gather_t* g, int32_t ctas) {
    if (!g || ctas < 1 || ctas > 1024) { set_error("gather_set_push_ctas: bad arguments"); return 1; }
    g->push_ctas = ctas;
    return 0;
}

extern "C" void specb200_gather_destroy(specb200_gather_t* g) {
    if (!g) return;
    cudaDeviceSynchronize();
    for (void* p : g->opened) cudaIpcCloseMemHandle(p);
    if (g->region) cudaFree(g->region);
    delete g;
}
